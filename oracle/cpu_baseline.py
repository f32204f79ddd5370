"""oracle/cpu_baseline.py -- timing harness for the reference's CPU path.  TEST INFRASTRUCTURE ONLY:
imported by bench.py's `cpu_baseline` leg and its `--impl reference` arm, never by transform_b200.

What is timed, per workload:
  linear   : Transform.resample(method='linear') as the reference executes it --
             pixel grid (core.py:574-578) -> Composition._reverse with np.matmul on [...,2,1]
             stacks (basic.py:160-176; oracle_np with MATVEC_IMPL='numpy') -> interpND.
             interpND is the REFERENCE'S OWN compiled module (oracle/_ref/helpers*.so, built
             from /root/reference/transform/helpers.pyx) when present, else the oracle port.
  jacobian : the repaired interpND_grid pipeline (oracle_np.resample_jacobian: NumPy jacobian()
             + the C restatement of the compiled interpND_jacobian loop).  The reference's own
             version cannot run, so this leg is always the port.
The reference is single-threaded; "all the host threads it can use" = independent row blocks
farmed out to os.cpu_count() forked workers (pixels are independent, so row tiling is exact).
"""
import multiprocessing as mp
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)

import oracle_np as o          # noqa: E402

_G = {}


def have_ref_helpers():
    try:
        import ref_import
        ref_import.load_ref_helpers()
        return True
    except Exception:
        return False


def affine_chain(n):
    """BASELINE config 2 scaled to n x n: Scale(1.3) o Rotation(30 deg) o Offset(-n/2)."""
    return o.Composition([o.Scale(1.3), o.Rotation(30, "deg"), o.Offset([-n / 2.0, -n / 2.0])])


def radial_chain(n):
    """BASELINE config 3: polar unwrap filling n x n."""
    c = (n - 1) / 2.0
    rmax = n / np.sqrt(2.0)
    return o.Composition([o.Scale([n / (2 * np.pi), n / rmax]), o.Radial(origin=[c, c])])


def synthetic_image(n, seed):
    return np.random.default_rng(seed).random((n, n), dtype=np.float32)


def _linear_rows(args):
    y0, y1 = args
    src, chain, n = _G["src"], _G["chain"], _G["n"]
    interp = _G["interp"]
    out = None
    for ya in range(y0, y1, 64):
        yb = min(y1, ya + 64)
        coords = chain.reverse(o.pixel_grid(n, n, ya, yb))
        blk = interp(src, coords, method="l", bound="t")
        out = blk if out is None else out
    return float(out[0, 0])


def _jacobian_rows(args):
    y0, y1 = args
    out = o.resample_jacobian(_G["chain"], _G["src"], method="g", bound="t", rows=(y0, y1))
    return float(out[0, 0])


def _run(fn, blocks, nproc):
    t0 = time.perf_counter()
    if nproc <= 1:
        for b in blocks:
            fn(b)
    else:
        ctx = mp.get_context("fork")
        with ctx.Pool(nproc) as pool:
            pool.map(fn, blocks, chunksize=1)
    return time.perf_counter() - t0


def time_sample(kind, n, rows, nproc, src=None, seed=0):
    """Time `rows` output rows (spread over the image) of an n x n resample on `nproc` workers.
    Returns (seconds, pixels, impl_kind) with impl_kind 'reference' or 'port'."""
    if src is None:
        src = synthetic_image(n, seed)
    _G["src"], _G["n"] = src, n
    impl = "port"
    if kind == "linear":
        _G["chain"] = affine_chain(n)
        o.MATVEC_IMPL = "numpy"
        if have_ref_helpers():
            import ref_import
            _G["interp"] = ref_import.load_ref_helpers().interpND
            impl = "reference"
        else:
            _G["interp"] = o.interpND
        fn = _linear_rows
    elif kind == "jacobian":
        _G["chain"] = radial_chain(n)
        o.clib()                      # build/load the C restatement before forking
        fn = _jacobian_rows
    else:
        raise ValueError(kind)
    nproc = max(1, int(nproc))
    per = max(1, rows // nproc)
    # row blocks spread evenly over the image height so the sample sees the whole map
    starts = np.linspace(0, n - per, nproc).astype(int)
    blocks = [(int(s), int(s) + per) for s in starts]
    try:
        dt = _run(fn, blocks, nproc)
    finally:
        o.MATVEC_IMPL = "fma_c"
    return dt, per * nproc * n, impl
