#!/usr/bin/env python3
"""oracle/build_ref_repaired.py -- TEST INFRASTRUCTURE ONLY.

The reference's anti-aliased resampler (interpND_grid -> jacobian -> interpND_jacobian,
helpers.pyx:1145-1803) cannot run as shipped.  This recipe builds the
reference's OWN code with the minimal named repairs applied, so that the oracle's restatement of that
path (oracle_np.interp_grid_jacobian, oracle_c.orc_interp_jacobian) can be pinned against it:

    R1   jacobian(): add the missing `return J`
    R2   jacobian(): the `jump_detect` keyword shadows the function of that name; the keyword is
         renamed and interpND_grid's float threshold reaches jump_detect() as its jump_thresh
    R3   interpND_grid(): return the result of interpND_jacobian()
    R4   interpND_jacobian(): delete `assert(0)`
    R5   interpND_jacobian(): `reg_size/2` on the undeclared (Python object) reg_size is a float under
         Python 3; integer halves, as the C-typed loop variables require
    R6   interpND_grid(): hand `bound` over as C ints (the signature says int[:])
    R11  jacobian(): `J[1:-1,1:-1,...].T /= wgt.clip(1)` raises; divide with explicit broadcasting
(R7, R8, R10 keep the code as it is; R9 and R12 are NOT
applied here -- they are the two places where the oracle deliberately departs from the text, and the
tests stay away from the mirror boundary and from non-square grids when they compare.)

The source is read where it lies under /root/reference, patched IN MEMORY, written to a scratch
directory under /tmp for Cython, and only the built module lands in oracle/_ref/
(helpers_repaired*.so).  Every substitution must match exactly once, or the build stops.

Usage:  python oracle/build_ref_repaired.py [--force]
"""
import os
import shutil
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TFM_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

REPAIRS = [
    # R2 (signature): the threshold is the only knob; jump detection is on when it is non-zero
    ("R2a", "def jacobian(index, \n             jump_detect=True,\n             jump_thresh=10.\n             ):",
            "def jacobian(index, \n             jump_thresh=10.\n             ):\n    do_jump = bool(jump_thresh)"),
    ("R2b", "    if(jump_detect):\n        jf=jump_detect(Js)\n",
            "    if(do_jump):\n        jf=jump_detect(Js, jump_thresh)\n"),
    # R11 + the 2-D branch's use of the keyword
    ("R11", "            J[1:-1,1:-1,...].T /= wgt.clip(1)\n        else:\n            J[1:-1,1:-1,...] /= 4",
            "            J[1:-1,1:-1,...] /= wgt.clip(1)[...,None,None]\n        else:\n            J[1:-1,1:-1,...] /= 4"),
    # R1: return the array (the text that follows the function is the @cython decorator of the next one)
    ("R1", "            J[ tuple(     list(repeat(np.s_[:], axis)) + [J.size[axis]-1] ) ] = (\n"
           "                J[ tuple( list(repeat(np.s_[:], axis)) + [J.size[axis]-2] ) ]\n                )\n",
           "            J[ tuple(     list(repeat(np.s_[:], axis)) + [J.size[axis]-1] ) ] = (\n"
           "                J[ tuple( list(repeat(np.s_[:], axis)) + [J.size[axis]-2] ) ]\n                )\n"
           "    return J\n"),
    # R2 (call site) / R3 / R6
    ("R2d", "    J = jacobian(index, jump_detect=jump_detect)\n", "    J = jacobian(index, jump_thresh=jump_detect)\n"),
    ("R6", "    bound = np.array(bound)    \n", "    bound = np.array(bound, dtype=np.intc)\n"),
    ("R3", "    interpND_jacobian(source.astype(RDTYPE),\n", "    return interpND_jacobian(source.astype(RDTYPE),\n"),
    # R5: reg_size is an undeclared (Python-object) variable, `reg_size/2` a float under Python 3
    ("R5a", "            for yr in range(-reg_size/2, reg_size/2 + 1):\n", "            for yr in range(-(reg_size//2), reg_size//2 + 1):\n"),
    ("R5b", "                for xr in range(-reg_size/2,reg_size/2+1):\n", "                for xr in range(-(reg_size//2), reg_size//2 + 1):\n"),
    # R4
    ("R4", "    xyf = np.empty([2])\n    \n    assert(0)\n", "    xyf = np.empty([2])\n    \n"),
]
# the remaining `if(jump_detect):` tests inside jacobian() (1-D, 2-D and N-D branches)
JUMP_IFS = ("        if(jump_detect):\n", "        if(do_jump):\n")


def patched_source():
    src = open(os.path.join(REF, "transform", "helpers.pyx")).read()
    for tag, old, new in REPAIRS:
        n = src.count(old)
        if n != 1:
            raise RuntimeError("oracle/build_ref_repaired: repair %s matches %d times (expected 1) -- "
                               "the reference text changed" % (tag, n))
        src = src.replace(old, new)
    a = src.index("def jacobian(index,")
    b = src.index("def interpND_grid(")
    body = src[a:b]
    if body.count(JUMP_IFS[0]) < 2:
        raise RuntimeError("oracle/build_ref_repaired: jump_detect tests not found in jacobian()")
    body = body.replace(JUMP_IFS[0], JUMP_IFS[1]).replace("            if(jump_detect):\n", "            if(do_jump):\n")
    return src[:a] + body + src[b:]


def build(force=False):
    if not os.path.exists(os.path.join(REF, "transform", "helpers.pyx")):
        print("oracle/build_ref_repaired: reference sources not present at %s; nothing built" % REF)
        return False
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    target = os.path.join(OUT, "helpers_repaired" + suffix)
    if not force and os.path.exists(target):
        return True
    import numpy
    from Cython.Build import cythonize
    from setuptools import Extension
    from setuptools.dist import Distribution
    from setuptools.command.build_ext import build_ext

    os.makedirs(OUT, exist_ok=True)
    scratch = tempfile.mkdtemp(prefix="tfm_ref_repaired_")
    try:
        pyx = os.path.join(scratch, "helpers_repaired.pyx")
        with open(pyx, "w") as f:
            f.write(patched_source())
        exts = [Extension("helpers_repaired", [pyx], include_dirs=[numpy.get_include()],
                          define_macros=[("NPY_NO_DEPRECATED_API", "NPY_1_7_API_VERSION")],
                          extra_compile_args=["-O2", "-w"])]
        exts = cythonize(exts, language_level=2, build_dir=os.path.join(scratch, "cy"), quiet=True)
        dist = Distribution({"ext_modules": exts})
        cmd = build_ext(dist)
        cmd.build_temp = os.path.join(scratch, "tmp")
        cmd.build_lib = os.path.join(scratch, "lib")
        cmd.finalize_options()
        cmd.run()
        shutil.copyfile(os.path.join(scratch, "lib", "helpers_repaired" + suffix), target)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    print("oracle/build_ref_repaired: built %s" % os.path.basename(target))
    return True


if __name__ == "__main__":
    build(force="--force" in sys.argv)
