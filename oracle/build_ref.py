#!/usr/bin/env python3
"""oracle/build_ref.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Compiles the reference's two Cython sources *where they lie* under /root/reference
(transform/helpers.pyx, transform/SVD.pyx) and its two Python modules of the path (transform/core.py,
transform/basic.py) into shared objects under oracle/_ref/.
Only the built .so files land in the repo tree (oracle/_ref/ is git-ignored but
travels to the GPU box); no reference source is copied.

Recipe: cythonize with language_level=2 -- with Cython 3's default
language_level=3 helpers.pyx does not compile (helpers.pyx:1678-1679).  The generated
.c files and object files go to a scratch directory under /tmp.

Usage:  python oracle/build_ref.py [--force]
Exit status 0 and a one-line note when /root/reference is absent (GPU box).
"""
import os
import shutil
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TFM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def build(force=False):
    srcs = {name: os.path.join(REF, "transform", name + ".pyx") for name in ("helpers", "SVD")}
    # the reference's two Python modules of the path, compiled the same way (Cython accepts .py): its own
    # Transform / Composition / Linear / Radial classes then travel to the GPU box as shared objects, so that
    # bench.py's reference arm can drive the reference's own Composition.invert there (ref_import.load_reference_compiled)
    pysrcs = {name: os.path.join(REF, "transform", name + ".py") for name in ("core", "basic")}
    if not all(os.path.exists(p) for p in list(srcs.values()) + list(pysrcs.values())):
        print("oracle/build_ref: reference sources not present at %s; nothing built" % REF)
        return False
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    targets = {name: os.path.join(OUT, name + suffix) for name in list(srcs) + list(pysrcs)}
    if not force and all(os.path.exists(t) for t in targets.values()):
        return True

    import numpy
    from Cython.Build import cythonize
    from setuptools import Extension
    from setuptools.dist import Distribution
    from setuptools.command.build_ext import build_ext

    os.makedirs(OUT, exist_ok=True)
    scratch = tempfile.mkdtemp(prefix="tfm_ref_build_")
    try:
        exts = [Extension(name, [path], include_dirs=[numpy.get_include()],
                          define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")],
                          extra_compile_args=["-O2", "-w"])
                for name, path in srcs.items()]
        exts = cythonize(exts, language_level=2, build_dir=os.path.join(scratch, "cy"),
                         quiet=True)
        # basic.py:1175 reads a local before assignment inside a bare try/except (Cubicpin's Cardano branch): Python
        # raises UnboundLocalError at run time and the except swallows it; Cython must do the same, not refuse
        import Cython.Compiler.Options as cyopts
        for name, path in pysrcs.items():
            cyopts.error_on_uninitialized = (name != "basic")
            exts = exts + cythonize([Extension(name, [path], extra_compile_args=["-O2", "-w"])], language_level=3,
                                    build_dir=os.path.join(scratch, "cy"), quiet=True)
        cyopts.error_on_uninitialized = True
        # Cython 3.3 emits `PyList_Reverse(((PyObject*)None))` for core.py:914 (`list(shape).reverse()`, the
        # reference's own quirk: the statement assigns None): the operand it dropped is the temporary that holds the
        # freshly built list.  The GENERATED C is repaired (the reference source is untouched); exactly one site.
        import re
        for ext in exts:
            for cfile in ext.sources:
                if not cfile.endswith("core.c"):
                    continue
                txt = open(cfile).read()
                bad = "PyList_Reverse(((PyObject*)None))"
                n = txt.count(bad)
                if n == 0:
                    continue
                if n != 1:
                    raise RuntimeError("oracle/build_ref: unexpected number of PyList_Reverse(None) sites: %d" % n)
                at = txt.index(bad)
                tmp = re.findall(r"(__pyx_t_\d+) = PySequence_List\(", txt[:at])[-1]
                open(cfile, "w").write(txt.replace(bad, "PyList_Reverse(%s)" % tmp))
        dist = Distribution({"ext_modules": exts})
        cmd = build_ext(dist)
        cmd.build_temp = os.path.join(scratch, "tmp")
        cmd.build_lib = os.path.join(scratch, "lib")
        cmd.finalize_options()
        cmd.run()
        for name, tgt in targets.items():
            shutil.copyfile(os.path.join(scratch, "lib", name + suffix), tgt)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    print("oracle/build_ref: built %s" % ", ".join(os.path.basename(t) for t in targets.values()))
    return True


if __name__ == "__main__":
    build(force="--force" in sys.argv)
