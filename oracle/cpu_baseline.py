"""oracle/cpu_baseline.py -- timing harness for the reference's CPU path.  TEST INFRASTRUCTURE ONLY:
imported by bench.py's `cpu_baseline` leg and its `--impl reference` arm, never by transform_b200.

What is timed, per workload:
  linear   : Transform.resample(method='linear') as the reference executes it --
             pixel grid (core.py:574-578) -> Composition.invert (core.py:1556-1559 over basic.py:160-176,
             np.matmul on [...,2,1] stacks) -> interpND.  All three are the REFERENCE'S OWN code when
             oracle/_ref holds its compiled modules (core / basic / helpers, built by oracle/build_ref.py from
             the files where they lie under /root/reference): impl kind 'reference'.  With only helpers*.so
             the chain is oracle_np's port with MATVEC_IMPL='numpy' ('reference-interp'); with nothing, the port.
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
    ref_chain = _G.get("ref_chain")
    out = None
    for ya in range(y0, y1, 64):
        yb = min(y1, ya + 64)
        if ref_chain is not None:
            # the reference's own code end to end: its integer pixel grid (core.py:574-578, rows ya:yb of it), its
            # Composition.invert (core.py:298-315, 1556-1559 -> basic.py:160-176), its compiled interpND
            grid = np.mgrid[0:n, ya:yb].transpose()
            coords = ref_chain.invert(grid)
        else:
            coords = chain.reverse(o.pixel_grid(n, n, ya, yb))
        blk = interp(src, coords, method="l", bound="t")
        out = blk if out is None else out
    return float(out[0, 0])


def _jacobian_rows(args):
    y0, y1 = args
    out = o.resample_jacobian(_G["jchain"], _G["src64"] if "src64" in _G else _G["src"], method="g", bound="t",
                              rows=(y0, y1))
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


def _blocks(n, rows, nproc):
    per = max(1, rows // nproc)
    # row blocks spread evenly over the image height so the sample sees the whole map
    starts = np.linspace(0, n - per, nproc).astype(int)
    return [(int(s), int(s) + per) for s in starts], per * nproc * n


def _setup(kind, n, src):
    """Set the globals a worker needs for `kind`; returns (row function, impl_kind)."""
    _G["src"], _G["n"] = src, n
    impl = "port"
    if kind == "linear":
        _G["chain"] = affine_chain(n)
        _G.pop("ref_chain", None)
        if have_ref_helpers():
            import ref_import
            _G["interp"] = ref_import.load_ref_helpers().interpND
            impl = "reference-interp"
            try:
                # the reference's own Transform classes, compiled from core.py / basic.py into oracle/_ref
                ref = ref_import.load_reference_compiled()
                _G["ref_chain"] = ref.Composition([ref.Scale(1.3, dim=2), ref.Rotation(30, unit="deg"),
                                                   ref.Offset([-n / 2.0, -n / 2.0])])
                _G["interp"] = ref.interpND
                impl = "reference"
            except Exception:
                pass
        else:
            _G["interp"] = o.interpND
        return _linear_rows, impl
    if kind == "jacobian":
        _G["jchain"] = radial_chain(n)
        # interpND_grid converts the frame to float64 once per call (helpers.pyx:1520); row blocks of ONE
        # call share that copy
        _G["src64"] = np.ascontiguousarray(src, dtype=np.float64)
        o.clib()                      # build/load the C restatement before forking
        return _jacobian_rows, impl
    raise ValueError(kind)


def time_sample(kind, n, rows, nproc, src=None, seed=0):
    """Time `rows` output rows (spread over the image) of an n x n resample on `nproc` workers.
    Returns (seconds, pixels, impl_kind) with impl_kind 'reference' or 'port'."""
    if src is None:
        src = synthetic_image(n, seed)
    fn, impl = _setup(kind, n, src)
    nproc = max(1, int(nproc))
    blocks, px = _blocks(n, rows, nproc)
    o.MATVEC_IMPL = "numpy"
    try:
        dt = _run(fn, blocks, nproc)
    finally:
        o.MATVEC_IMPL = "fma_c"
    return dt, px, impl


class CpuArm:
    """The reference's CPU path on a PERSISTENT pool of forked workers: the workers are created once,
    after the frame and both chains are in place, and every timed sample is one pool.map -- no
    per-sample fork, no re-pickling of the 268 MB frame (round 1's reference arm paid both on every
    step and reported 7.2 Mpix/s where the same box's one-shot leg measured 12.1)."""

    def __init__(self, n, nproc, src=None, seed=0):
        self.n, self.nproc = n, max(1, int(nproc))
        src = synthetic_image(n, seed) if src is None else src
        self.fn_lin, self.kind_lin = _setup("linear", n, src)
        self.fn_jac, self.kind_jac = _setup("jacobian", n, src)
        o.MATVEC_IMPL = "numpy"                # np.matmul on [...,2,1] stacks, as basic.py:150-152 does
        self.pool = mp.get_context("fork").Pool(self.nproc) if self.nproc > 1 else None
        if self.pool is not None:
            self.pool.map(_warm, range(self.nproc * 2), chunksize=1)

    def _time(self, fn, rows):
        blocks, px = _blocks(self.n, rows, self.nproc)
        t0 = time.perf_counter()
        if self.pool is None:
            for b in blocks:
                fn(b)
        else:
            self.pool.map(fn, blocks, chunksize=1)
        return time.perf_counter() - t0, px

    def linear(self, rows):
        return self._time(self.fn_lin, rows)

    def jacobian(self, rows):
        return self._time(self.fn_jac, rows)

    def close(self):
        o.MATVEC_IMPL = "fma_c"
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None


def _warm(_):
    return os.getpid()
