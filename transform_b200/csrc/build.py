#!/usr/bin/env python3
"""Build libtfm_b200.so (sm_100a) in-tree with nvcc.

    python transform_b200/csrc/build.py [--force] [--verbose]

The shared library lands in transform_b200/lib/ (git-ignored, but it travels to the GPU box with
the gpurun snapshot).  Called by __graft_entry__.build().
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
OUT_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(HERE, "_obj")
OUT = os.path.join(OUT_DIR, "libtfm_b200.so")
SOURCES = ["tfm_capi.cu", "tfm_resample.cu", "tfm_resample_exact_f32.cu", "tfm_resample_exact_f64.cu",
           "tfm_resample_fast.cu", "tfm_apply.cu", "tfm_jacobian.cu", "tfm_jacobian_fast.cu", "tfm_fits.cu"]
HEADERS = ["tfm_chain.cuh", "tfm_common.cuh", "tfm_resample.cuh", "tfm_jacobian_common.cuh", os.path.join("..", "..", "include", "tfm_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build_variant(name, defines):
    """A/B builds for kernel experiments: lib/variants/libtfm_b200_<name>.so compiled with extra -D
    flags; select it at run time with TFM_B200_LIB=<path> (tools/ab_variants.py)."""
    vdir = os.path.join(OUT_DIR, "variants")
    odir = os.path.join(OBJ_DIR, "variant_" + name)
    os.makedirs(vdir, exist_ok=True)
    os.makedirs(odir, exist_ok=True)
    out = os.path.join(vdir, "libtfm_b200_%s.so" % name)

    def one(src):
        obj = os.path.join(odir, src + ".o")
        cmd = [NVCC] + FLAGS + ["-D" + d for d in defines] + ["-c", os.path.join(HERE, src), "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout))
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, SOURCES))
    cmd = [NVCC, "-shared", "-o", out] + objs + ["-Xlinker", "--exclude-libs,ALL", "-lcudart_static",
                                                  "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    return out


def build(force=False, verbose=False):
    srcs = [os.path.join(HERE, s) for s in SOURCES]
    deps = srcs + [os.path.join(HERE, h) for h in HEADERS] + [os.path.abspath(__file__)]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= _newest(deps):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdr_time = _newest(deps[len(srcs):])

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
        if (not force and os.path.exists(obj)
                and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_time)):
            return obj
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stdout))
        if verbose:
            print(r.stdout)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, srcs))
    tmp = OUT + ".tmp%d" % os.getpid()
    cmd = [NVCC, "-shared", "-o", tmp] + objs + ["-Xlinker", "--exclude-libs,ALL", "-lcudart_static",
                                                  "-lpthread", "-ldl", "-lrt"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s" % r.stdout)
    os.replace(tmp, OUT)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
