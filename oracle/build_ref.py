#!/usr/bin/env python3
"""oracle/build_ref.py -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

Compiles the reference's two Cython sources *where they lie* under /root/reference
(transform/helpers.pyx, transform/SVD.pyx) into shared objects under oracle/_ref/.
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
    if not all(os.path.exists(p) for p in srcs.values()):
        print("oracle/build_ref: reference sources not present at %s; nothing built" % REF)
        return False
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    targets = {name: os.path.join(OUT, name + suffix) for name in srcs}
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
