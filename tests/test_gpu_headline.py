"""The kernel variants behind the headline numbers, pinned against the oracle AT the BASELINE sizes.

bench.py's step is two calls on an 8192^2 float32 frame (linear through Scale o Rotation o Offset,
Jacobian-Gaussian through the polar unwrap); BASELINE config 2 is the same affine chain at 4096^2,
config 5 is Transform.apply on a long vector array.  Each test here
  (a) issues exactly the call the bench issues (same transform, size, precision, device-resident data),
  (b) asserts WHICH kernel variant the dispatcher chose (tfm_last_kernel), so that a silent change of
      the dispatch rules cannot leave the headline variant untested, and
  (c) compares >= 128 full output rows spread over the frame plus the four borders with the oracle
      (oracle.resample(rows=...) / oracle.resample_jacobian(rows=...), helpers.pyx:559-611,
      1542-1803 restated) under a tolerance whose outliers are capped in MAGNITUDE, not only counted.
"""
import numpy as np
import pytest

from conftest import to_oracle

pytestmark = pytest.mark.gpu

K2_TOL = 3e-4          # stated tolerance of the fp32 Jacobian method
K2_CAP = 0.3           # no sampled pixel (knife-edge ones included) may be off by more than this: a window shifted by
                       # one source pixel trades one row/column (~15 % of the weight) of uniform [0,1) samples for another


def sample_rows(n, k=128):
    rows = set(int(r) for r in np.linspace(0, n - 1, k))
    rows.update((0, 1, 2, n - 3, n - 2, n - 1))           # bottom and top borders (left/right: every row is full width)
    return sorted(rows)


def named_chain(tfm, n):
    return tfm.Composition([tfm.Scale(1.3, dim=2), tfm.Rotation(30, unit='deg'), tfm.Offset([-n / 2.0, -n / 2.0])])


def polar_chain(tfm, n):
    c = (n - 1) / 2.0
    return tfm.Composition([tfm.Scale([n / (2 * np.pi), n / (n / np.sqrt(2.0))]), tfm.Radial(origin=[c, c])])


def frame(n, seed):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)


def oracle_rows(oracle, otr, src, rows, **kw):
    return np.stack([oracle.resample(otr, src, rows=(y, y + 1), **kw)[0] for y in rows])


def err_stats(got, want):
    err = np.abs(np.asarray(got, dtype=np.float64) - np.asarray(want, dtype=np.float64))
    assert np.array_equal(np.isnan(got), np.isnan(want))
    err = np.nan_to_num(err)
    return err, float(err.max()), float(np.median(err))


# --------------------------------------------------------------------------------------------------
# K1 linear, fp32 fast path
# --------------------------------------------------------------------------------------------------
K1_VARIANTS = [  # (tuning, expected kernel name, None = whatever the default dispatch picks)
    (0, None),
    (16, "k_resample<trunc,affine>"),
    (256, "k_resample_tma<affine>"),
    (128, "k_resample_tex<affine>"),
    (8192, "k_resample_pipe<affine>"),
]


@pytest.mark.parametrize("n", [8192, 4096])
def test_linear_fp32_named_chain_all_variants(tfm, oracle, n):
    """bench step call 1 (8192^2) and BASELINE config 2 (4096^2): every fp32 linear kernel variant
    against the oracle, 1e-5 (north_star's fp32 tolerance), no outliers allowed."""
    from transform_b200 import _lib
    src = frame(n, 1000)
    tr = named_chain(tfm, n)
    rows = sample_rows(n)
    want = oracle_rows(oracle, to_oracle(tr), src.cpu().numpy(), rows, method='linear')
    default_name = None
    results = {}
    for tuning, name in K1_VARIANTS:
        out = tr.resample(src, method='linear', precision='fp32', tuning=tuning)
        k = _lib.last_kernel()
        if name is None:
            default_name = k
            # the named chain maps 58 % of the output outside the source: the texture gather is the default
            assert k == "k_resample_tex4<affine>", k
        else:
            assert k == name, (tuning, k)
        got = out[rows].cpu().numpy()
        err, emax, emed = err_stats(got, want)
        assert emax <= 1e-5, (n, tuning, k, emax, emed)
        results[tuning] = out
    # all variants round the fp32 fraction identically: bit-equal outputs (what sharding relies on)
    import torch
    for tuning, out in results.items():
        assert torch.equal(out, results[0]), (tuning, default_name)


@pytest.mark.parametrize("case", ["shift", "rot5", "rot45", "scale0p8", "scale2"])
def test_linear_fp32_full_coverage_maps(tfm, oracle, case):
    """The maps of bench.py's roofline breakdown (every source pixel is touched), 8192^2, default
    dispatch and every opt-in variant, against the oracle on sampled rows."""
    from transform_b200 import _lib
    import torch
    n = 8192
    src = frame(n, 7)
    tr = {"shift": tfm.Offset([3.25, -7.5]),
          "rot5": tfm.Wrap(tfm.Rotation(5, unit='deg'), tfm.Offset([-n / 2.0, -n / 2.0])),
          "rot45": tfm.Wrap(tfm.Rotation(45, unit='deg'), tfm.Offset([-n / 2.0, -n / 2.0])),
          "scale0p8": tfm.Scale(0.8, dim=2), "scale2": tfm.Scale(2.0, dim=2)}[case]
    rows = sample_rows(n, 64)
    want = oracle_rows(oracle, to_oracle(tr), src.cpu().numpy(), rows, method='linear')
    ref = None
    for tuning in (0, 16, 128, 256, 8192):
        out = tr.resample(src, method='linear', precision='fp32', tuning=tuning)
        k = _lib.last_kernel()
        if tuning == 0:
            # >= 55 % of the output maps inside the source: the persistent TMA pipeline is the default
            assert k == "k_resample_pipe<affine>", (case, k)
        err, emax, emed = err_stats(out[rows].cpu().numpy(), want)
        assert emax <= 1e-5, (case, tuning, k, emax)
        if ref is None:
            ref = out
        else:
            assert torch.equal(out, ref), (case, tuning, k)


@pytest.mark.parametrize("case", ["shift", "rot5", "rot45", "scale0p8"])
def test_cubic_fp32_full_coverage_maps(tfm, oracle, case):
    """interpND's cubic method (helpers.pyx:741-789) at 8192^2 on maps that cover the source: the cubic
    instantiation of the persistent TMA pipeline (default) and the load/store kernel (tuning bit 16) against the
    oracle on sampled rows (2e-5: sixteen fp32 taps, overshoot up to 1.3x the data range) and against each other
    bit for bit -- same coordinates, same floor/fraction, same Hermite order."""
    from transform_b200 import _lib
    import torch
    n = 8192
    src = frame(n, 11)
    tr = {"shift": tfm.Offset([3.25, -7.5]),
          "rot5": tfm.Wrap(tfm.Rotation(5, unit='deg'), tfm.Offset([-n / 2.0, -n / 2.0])),
          "rot45": tfm.Wrap(tfm.Rotation(45, unit='deg'), tfm.Offset([-n / 2.0, -n / 2.0])),
          "scale0p8": tfm.Scale(0.8, dim=2)}[case]
    rows = sample_rows(n, 48)
    want = oracle_rows(oracle, to_oracle(tr), src.cpu().numpy(), rows, method='cubic')
    out = tr.resample(src, method='cubic', precision='fp32')
    assert _lib.last_kernel() == "k_resample_pipe<affine,cubic>", (case, _lib.last_kernel())
    err, emax, emed = err_stats(out[rows].cpu().numpy(), want)
    assert emax <= 2e-5, (case, emax, emed)
    ref = tr.resample(src, method='cubic', precision='fp32', tuning=16)
    assert _lib.last_kernel() == "k_resample<trunc,affine>", _lib.last_kernel()
    assert torch.equal(out, ref), (case, float((out - ref).abs().max()))
    # a ragged output grid (not a multiple of the 64 x 32 tile) and a smaller frame
    m = 1000
    small = src[:m, :1024].contiguous()
    a = tr.resample(small, method='cubic', precision='fp32', shape=(m - 7, 1024 - 13))
    b = tr.resample(small, method='cubic', precision='fp32', shape=(m - 7, 1024 - 13), tuning=16)
    assert torch.equal(a, b), case


def test_nearest_fp32_config2(tfm, oracle):
    """BASELINE config 2, nearest: indices bit-exact.  In the fast path the host pre-composes the
    affine chain, which can move a coordinate by an ulp: a mismatch is legal only where the
    reference's coordinate sits within 1e-9 of a half-integer tie."""
    n = 4096
    src = frame(n, 7)
    tr = named_chain(tfm, n)
    rows = sample_rows(n)
    otr = to_oracle(tr)
    srcn = src.cpu().numpy()
    want = oracle_rows(oracle, otr, srcn, rows, method='nearest')
    got = tr.resample(src, method='nearest', precision='fp32')[rows].cpu().numpy()
    bad = np.argwhere(got != want)
    for (ri, x) in bad[:50]:
        c = otr.reverse(np.array([[float(x), float(rows[ri])]]))[0]
        frac = np.abs((c + 0.5) - np.round(c + 0.5))
        assert frac.min() < 1e-9, (rows[ri], x, c)
    assert len(bad) <= 4, len(bad)
    # parity mode: bit-exact everywhere
    exact = tr.resample(src, method='nearest')[rows].cpu().numpy()
    assert np.array_equal(exact, want)


# --------------------------------------------------------------------------------------------------
# parity mode (fp64) at full size
# --------------------------------------------------------------------------------------------------
def test_parity_mode_8192(tfm, oracle):
    """precision='fp64' at 8192^2: linear through the affine chain is bit-exact; linear through the
    polar chain holds 1e-12 relative (libm last bits); the Jacobian-Gaussian holds 1e-11."""
    from transform_b200 import _lib
    n = 8192
    src = frame(n, 1000)
    srcn = src.cpu().numpy()
    rows = sample_rows(n, 128)
    tr = named_chain(tfm, n)
    got = tr.resample(src, method='linear')
    assert _lib.last_kernel() == "k_resample<trunc>"
    want = oracle_rows(oracle, to_oracle(tr), srcn, rows, method='linear')
    assert np.array_equal(got[rows].cpu().numpy(), want)
    del got
    rad = polar_chain(tfm, n)
    got = rad.resample(src, method='linear')[rows].cpu().numpy()
    want = oracle_rows(oracle, to_oracle(rad), srcn, rows, method='linear')
    err, emax, emed = err_stats(got, want)
    assert emax <= 1e-9, emax            # |coordinate| ~ 8e3 x 2 ulp of libm sincos = 4e-12 px x gradient <= 1
    src64 = srcn.astype(np.float64)
    rows_j = sample_rows(n, 24)
    got = rad.resample(src, method='Gaussian')
    assert _lib.last_kernel() == "k_jacobian<exact64>"
    got = got[rows_j].cpu().numpy()
    cap = 2 * int(np.ceil(0.25 * n))
    want = np.stack([oracle.resample_jacobian(to_oracle(rad), src64, method='g', rows=(y, y + 1), max_reg=cap)[0]
                     for y in rows_j])
    err, emax, emed = err_stats(got, want)
    # Everything is 1e-11 except pixels that sit on a floor()/ceil() boundary of the footprint window,
    # where 2 ulp of sincos (CUDA libdevice vs glibc) move the window: the knife-edge columns on the
    # 45-degree rays (see test_jacobian_fp32_polar_8192) and a handful of ceil() boundaries of the
    # footprint size.  Magnitude cap: the weight of one window row/column (K2_CAP).
    knife = (np.arange(n) % (n // 8)) == 0
    assert emed <= 1e-13 and (err[:, ~knife] > 1e-10).sum() <= 8 and emax <= K2_CAP, (emax, emed, (err > 1e-10).sum())


# --------------------------------------------------------------------------------------------------
# K2 Jacobian-Gaussian, fp32 fast path
# --------------------------------------------------------------------------------------------------


def test_jacobian_fp32_polar_8192(tfm, oracle):
    """bench step call 2: the analytic-polar variant of k_jacobian<fast32> at 8192^2 -- every tap
    mode of it (texture gather = default, 8-byte and 4-byte loads) and the finite-difference kernel
    (tuning bit 512) -- against the oracle on sampled rows.

    Stated tolerance: 3e-4 absolute on uniform [0,1) data (K2_TOL), for EVERY sampled pixel except the
    knife-edge ones: with the origin at (n-1)/2 the output columns x = k * n/8 map onto the 45-degree
    rays, where source coordinates are exact integers up to the last bit of sin/cos.  floor() of such a
    coordinate -- and with it the hard-truncated footprint window -- is decided by rounding noise, in
    the reference as much as here (its result there changes with the libm).  Those 8 columns are
    compared under the magnitude cap only (K2_CAP: the weight of one window row/column)."""
    from transform_b200 import _lib
    n = 8192
    src = frame(n, 1000)
    rad = polar_chain(tfm, n)
    rows = sample_rows(n, 128)
    src64 = src.cpu().numpy().astype(np.float64)
    cap = 2 * int(np.ceil(0.25 * n))
    otr = to_oracle(rad)
    want = np.stack([oracle.resample_jacobian(otr, src64, method='g', rows=(y, y + 1), max_reg=cap)[0] for y in rows])
    knife = (np.arange(n) % (n // 8)) == 0
    ref = None
    # default = per-row spectral tables (the radius of this unwrap is constant along output rows) with two output
    # rows sharing each set of texture fetches; tuning bit 524288 = one pixel per set of fetches, 262144 = the
    # per-pixel closed form of the general polar op
    for tuning, name in ((0, "k_jacobian<fast32,analytic-polar-rows,tex,row-pairs>"),
                         (524288, "k_jacobian<fast32,analytic-polar-rows,tex>"),
                         (32768, "k_jacobian<fast32,analytic-polar-rows,ld64>"),
                         (4096, "k_jacobian<fast32,analytic-polar-rows,ld32>"),
                         (262144, "k_jacobian<fast32,analytic-polar,tex>"),
                         (262144 | 32768, "k_jacobian<fast32,analytic-polar,ld64>"),
                         (262144 | 4096, "k_jacobian<fast32,analytic-polar,ld32>"), (512, "k_jacobian<fast32,tex>"),
                         (512 | 4096, "k_jacobian<fast32>")):
        out = rad.resample(src, method='Gaussian', precision='fp32', tuning=tuning)
        assert _lib.last_kernel() == name, (tuning, _lib.last_kernel())
        got = out[rows].cpu().numpy()
        err, emax, emed = err_stats(got, want)
        assert emed < 2e-5, (name, emed)
        assert float(err[:, ~knife].max()) <= K2_TOL, (name, float(err[:, ~knife].max()))
        assert emax <= K2_CAP, (name, emax)
        if ref is None:
            ref = out
        elif not (tuning & 512):
            d = (out - ref).abs()
            d[:, torch_knife(n)] = 0
            assert float(d.max()) <= 2e-5, (name, float(d.max()))


def torch_knife(n):
    import torch
    return (torch.arange(n, device="cuda") % (n // 8)) == 0


def test_jacobian_fp32_affine_8192(tfm, oracle):
    """The constant-Jacobian variant (anti-aliased down-scaling + rotation) at 8192^2, every staging
    variant of it, against the oracle: 2e-5 everywhere (J is exact, no footprint-size flips)."""
    from transform_b200 import _lib
    import torch
    n = 8192
    src = frame(n, 5)
    tr = tfm.Composition([tfm.Offset([n / 2.0 + 0.3, n / 2.0 - 0.2]), tfm.Scale([0.6, 0.7]),
                          tfm.Rotation(25, unit='deg'), tfm.Offset([-n / 2.0, -n / 2.0])])
    rows = sample_rows(n, 96)
    src64 = src.cpu().numpy().astype(np.float64)
    cap = 2 * int(np.ceil(0.25 * n))
    otr = to_oracle(tr)
    want = np.stack([oracle.resample_jacobian(otr, src64, method='g', rows=(y, y + 1), max_reg=cap)[0] for y in rows])
    ref = None
    for tuning, name in ((0, "tex"), (4096, "ld32"), (32768, "ld64")):
        out = tr.resample(src, method='Gaussian', precision='fp32', tuning=tuning)
        k = _lib.last_kernel()
        assert k == "k_jacobian<fast32,const-J,%s>" % name, k
        err, emax, emed = err_stats(out[rows].cpu().numpy(), want)
        assert emax <= 2e-5, (tuning, k, emax, emed)
        if ref is None:
            ref = out
        else:
            assert float((out - ref).abs().max()) <= 2e-5, (tuning, k)


# --------------------------------------------------------------------------------------------------
# K3 apply (config 5) on a slice of the full job
# --------------------------------------------------------------------------------------------------
def test_apply_config5_slices(tfm, oracle):
    """BASELINE config 5's chains on 1e8 vectors (the 1e9 job is the same kernel looping longer):
    three 1e6-vector slices -- head, middle, ragged tail -- bit-exact (affine) / 1e-11 (Radial)."""
    import torch
    from transform_b200 import _lib
    nv = 10 ** 8 + 3
    g = torch.Generator(device="cuda").manual_seed(2)
    v = (torch.rand((nv, 2), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 8192
    aff = tfm.Composition([tfm.Scale([1.5, 2]), tfm.Rotation(30, unit='deg'), tfm.Offset([3, 4])]).inverse()
    rad = tfm.Composition([tfm.Scale([1.5, 2]), tfm.Radial(origin=[0.1, 0.2]), tfm.Offset([3, 4])]).inverse()
    sl = [slice(0, 10 ** 6), slice(nv // 2, nv // 2 + 10 ** 6), slice(nv - 10 ** 6 - 3, nv)]
    for tr, exact in ((aff, True), (rad, False)):
        out = tr.apply(v)
        assert _lib.last_kernel().startswith("k_apply"), _lib.last_kernel()
        for s in sl:
            want = oracle.apply(to_oracle(tr), v[s].cpu().numpy())
            got = out[s].cpu().numpy()
            if exact:
                assert np.array_equal(got, want)
            else:
                err = np.abs(got - want) / np.maximum(1.0, np.abs(want))
                assert float(err.max()) <= 1e-11, float(err.max())
        del out


def test_staged_rectangles_reproduce_the_full_frame_8192(tfm):
    """Multi-GPU shards at full size on one GPU: each output rectangle of a 2- and 3-way split (rows and
    columns) is resampled against the staged source rectangle the planner assigns it, exactly as a rank
    would, and must equal the same pixels of the unsharded call BIT FOR BIT (K1; K2 to 2e-6) -- for the kernels a shard
    runs (texture gather / TMA pipeline on a proven-sufficient rectangle, generic kernel otherwise).
    The named chain lands thousands of pixels on exact integer source columns, where the generic kernel
    once took the tap index from floor() and the fraction from the fixed-point floor."""
    import torch
    from transform_b200 import _engine, _lib, sharding
    n = 8192
    src = frame(n, 1000)
    c = (n - 1) / 2.0
    cases = [(named_chain(tfm, n), "linear", {}), (tfm.Offset([3.25, -7.5]), "linear", {}),
             (named_chain(tfm, n), "cubic", {}), (polar_chain(tfm, n), "Gaussian", {}),
             (tfm.Composition([tfm.Offset([c, c]), tfm.Scale([0.6, 0.7]), tfm.Rotation(25, unit='deg'),
                               tfm.Offset([-n / 2.0, -n / 2.0])]), "Gaussian", {})]
    for tr, method, kw in cases:
        want = tr.resample(src, method=method, precision="fp32", **kw)
        for world, split in ((2, "rows"), (3, "cols")):
            _, plans, _ = sharding.plan_rects(tr, (n, n), (n, n), method, "t", world, sharding._default_invert,
                                              split=split)
            for p in plans:
                x0, y0, x1, y1 = p["rect"]
                if p["whole"]:
                    sub, origin = src, (0, 0)
                elif p["box"] is None:
                    assert float(want[y0:y1, x0:x1].abs().max()) == 0.0
                    continue
                else:
                    bx0, by0, bx1, by1 = p["box"]
                    sub, origin = src[by0:by1, bx0:bx1].contiguous(), (bx0, by0)
                blk = sharding._default_resample_rect(tr, sub, origin, (n, n), p["rect"], (n, n), method, "t",
                                                      dict(precision="fp32", **kw))
                k = _lib.last_kernel()
                dmax = float((blk - want[y0:y1, x0:x1]).abs().max())
                if method == "Gaussian":
                    # footprints that cross the image edge are summed by the texture border path in the full
                    # frame and by the out-of-line edge path in a shard: same taps and weights, another order
                    assert dmax <= 2e-6, (method, world, split, p["rect"], k, dmax)
                else:
                    assert dmax == 0.0, (method, world, split, p["rect"], k, dmax)


def test_config4_tan_to_car_batch_4096(tfm, oracle):
    """BASELINE config 4 at its frame size: a batch of 4096^2 frames, TAN -> CAR (bench.config4_transform), fp32
    fast path -- the plane-batch texture kernel (k_resample_tex4_batch<car-tan>; 18 planes: two launches of 16 and
    2, groups of 8, 8, 2) and the one-launch-per-plane form, both against the oracle on sampled rows of the first,
    a middle and the last plane (1e-5 of the data range, the fp32 tolerance of linear interpolation) and against
    each other bit for bit on every pixel."""
    import sys
    import os
    import torch
    from transform_b200 import _lib
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    total, _, n = bench.config4_transform(tfm)
    nb = 18
    g = torch.Generator(device="cuda").manual_seed(44)
    frames = torch.rand((nb, n, n), generator=g, device="cuda", dtype=torch.float32)
    got = total.resample(frames, method="linear", precision="fp32", shape=(nb, n, n))
    assert _lib.last_kernel() == "k_resample_tex4_batch<car-tan>", _lib.last_kernel()
    ref = total.resample(frames, method="linear", precision="fp32", shape=(nb, n, n), tuning=1048576)
    assert _lib.last_kernel() == "k_resample_tex4<car-tan>", _lib.last_kernel()
    assert torch.equal(got, ref)
    rows = sample_rows(n, 48)
    otr = to_oracle(total)
    for k in (0, 9, nb - 1):
        src64 = frames[k].cpu().numpy().astype(np.float64)
        want = oracle_rows(oracle, otr, src64, rows, method="linear")
        err, emax, emed = err_stats(got[k][rows].cpu().numpy(), want)
        assert emax <= 1e-5, (k, emax, emed)
    assert float(got.abs().max()) > 0.5                          # the frames overlap: not a field of zeros
