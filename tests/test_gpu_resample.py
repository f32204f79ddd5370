"""GPU parity: K1 (fused chain + nearest/linear/cubic gather) against the oracle, through the
Python host side and the C ABI.  Bars (BASELINE.json north_star):
  nearest: bit-exact;  linear/cubic: <= 1e-12 relative in the fp64 mode (asserted bit-exact for
  affine chains, whose arithmetic convention is fully defined), <= 1e-5 relative in the fp32 path.
"""
import numpy as np
import pytest

from conftest import to_oracle, assert_close_rel

pytestmark = pytest.mark.gpu

BOUNDS = ['t', 'e', 'p', 'm', ['p', 'e'], ['m', 't'], 'extend', 'truncate']


def _affine(t, ny, nx):
    return t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit='deg'),
                          t.Offset([-nx / 2.0, -ny / 2.0])])


def _src(ny, nx, dtype, seed=0):
    return np.random.default_rng(seed).random((ny, nx)).astype(dtype)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("method", ['nearest', 'linear', 'cubic'])
@pytest.mark.parametrize("bound", BOUNDS)
def test_affine_exact_bitwise(tfm, oracle, dtype, method, bound):
    ny, nx = 67, 93                                   # ragged: not multiples of the 32x32 tile
    src = _src(ny, nx, dtype)
    tr = _affine(tfm, ny, nx)
    got = tr.resample(src, method=method, bound=bound)
    want = oracle.resample(to_oracle(tr), src, method=method, bound=bound)
    assert got.dtype == want.dtype
    assert got.shape == want.shape
    assert np.array_equal(got, want), "max diff %g" % np.abs(got - want).max()


@pytest.mark.parametrize("method", ['nearest', 'linear', 'cubic'])
def test_affine_fast32(tfm, oracle, method):
    ny, nx = 300, 411
    src = _src(ny, nx, np.float32, 3)
    tr = _affine(tfm, ny, nx)
    got = tr.resample(src, method=method, precision='fp32')
    want = oracle.resample(to_oracle(tr), src, method=method)
    assert got.dtype == np.float32
    if method == 'nearest':
        # The fast path pre-composes the affine chain on the host (the parity mode never does), so
        # a coordinate may round differently in its last bit; that can only flip an index when the
        # coordinate sits on a half-integer to within rounding.  Every mismatch must be such a tie.
        bad = np.argwhere(got != want.astype(np.float32))
        coords = to_oracle(tr).reverse(oracle.pixel_grid(ny, nx))
        for (y, x) in bad:
            t = coords[y, x] + 0.5
            assert np.abs(t - np.round(t)).min() < 1e-9, (y, x, coords[y, x])
        assert len(bad) < 1e-3 * got.size
    else:
        assert_close_rel(got, want, 1e-5, method)


@pytest.mark.parametrize("shape", [(5, 7), (40, 33), (64, 64), (1, 1), (200, 3)])
def test_output_shape_argument(tfm, oracle, shape):
    src = _src(37, 53, np.float64, 1)
    tr = tfm.Composition([tfm.Scale([0.7, 1.9]), tfm.Offset([-3.0, 2.0])])
    for m in ('n', 'l', 'c'):
        got = tr.resample(src, method=m, shape=shape)
        want = oracle.resample(to_oracle(tr), src, method=m, shape=shape)
        assert got.shape == tuple(shape)
        assert np.array_equal(got, want)


def test_empty_output(tfm):
    src = _src(8, 8, np.float32)
    out = tfm.Scale(2.0).resample(src, method='linear', shape=(0, 5))
    assert out.shape == (0, 5)


@pytest.mark.parametrize("kw", [dict(origin=[31.5, 25.5]), dict(origin=[31.5, 25.5], r0=2.0),
                                dict(ccw=True, pos_only=False, origin=[20.0, 30.0]),
                                dict(unit='deg', origin=[31.5, 25.5])])
@pytest.mark.parametrize("method", ['nearest', 'linear', 'cubic'])
def test_radial_unwrap(tfm, oracle, kw, method):
    ny, nx = 52, 64
    src = _src(ny, nx, np.float64, 2)
    rad = tfm.Radial(**kw)
    coeff = rad.params['ang_coeff']
    if kw.get('r0'):
        sc = tfm.Scale([nx / (2 * np.pi) * coeff, ny / 3.0])
    else:
        sc = tfm.Scale([nx / (2 * np.pi) * coeff, ny / 40.0])
    tr = tfm.Composition([sc, rad])
    got = tr.resample(src, method=method, bound='t')
    want = oracle.resample(to_oracle(tr), src, method=method, bound='t')
    if method == 'nearest':
        # CUDA's sincos/exp differ from glibc's by <= 2 ulp: an index can only flip within a hair
        # of a half-integer.  Count mismatches and demand that each is such a near-tie.
        bad = np.argwhere(got != want)
        coords = to_oracle(tr).reverse(oracle.pixel_grid(ny, nx))
        for (y, x) in bad:
            frac = np.abs((coords[y, x] + 0.5) - np.round(coords[y, x] + 0.5))
            assert frac.min() < 1e-9, "non-tie mismatch at %s: %s" % ((y, x), coords[y, x])
        assert len(bad) <= 4
    else:
        assert_close_rel(got, want, 1e-12, method)


def test_reference_kat_resample(tfm):
    """tests/test_01_core.py:259-294 of the reference, through the GPU path."""
    a = np.zeros([7, 5])
    a[2, 2] = 1
    b = tfm.Identity().resample(a, method='nearest')
    assert b.shape == (7, 5) and np.all(b == a)
    b = tfm.Identity().resample(a, shape=[5, 5], method='nearest')
    assert b.shape == (5, 5) and np.all(a[0:5, 0:5] == b)
    b = tfm.Scale(2.5, post=[2, 2], pre=[-2, -2]).resample(a, method='neares')
    chk = np.zeros([7, 5])
    chk[1:4, 1:4] = 1
    assert np.all(b == chk)
    b = tfm.Scale([1, 2.5], post=[2, 2], pre=[-2, -2]).resample(a, method='nearest')
    chk = np.zeros([7, 5])
    chk[1:4, 2] = 1
    assert np.all(b == chk)


def test_forbid_boundary_raises(tfm):
    src = _src(16, 16, np.float64)
    with pytest.raises(ValueError):
        tfm.Offset([5.0, 0.0]).resample(src, method='nearest', bound='f')
    out = tfm.Identity().resample(src, method='nearest', bound='f')
    assert np.array_equal(out, src)


def test_errors_match_reference(tfm):
    src = _src(16, 16, np.float64)
    with pytest.raises(ValueError):
        tfm.Scale(2.0).resample(src, method='quux')
    with pytest.raises(ValueError):
        tfm.Scale(2.0).resample(src, method='n', bound='x')
    with pytest.raises(ValueError):
        tfm.Scale(2.0).resample("not an array")
    with pytest.raises(AssertionError):
        tfm.Scale([0.0, 1.0]).resample(src)             # not invertible (basic.py:321)
    with pytest.raises(NotImplementedError):
        tfm.Scale(2.0).resample(src, method='fourier')   # outside the GPU path: loud, no fallback


def test_nan_and_absurd_coordinates(tfm, oracle):
    """interpND through the coordinate-lookup opcode with NaN / inf / 1e300 coordinates."""
    rng = np.random.default_rng(5)
    src = rng.random((37, 53))
    idx = rng.random((40, 45, 2)) * np.array([70, 60]) - 8
    idx[3, 4] = [np.nan, 2]
    idx[5, 6] = [1e300, -1e300]
    idx[7, 7] = [np.inf, 3]
    idx[8, 8] = [-0.5, 0.5]
    idx[9, 9] = [36.5, 52.5]
    for m in 'nlc':
        for b in ['t', 'e', 'p', 'm', ['p', 'e']]:
            for fv in (0, -9):
                got = tfm.interpND(src, idx, method=m, bound=b, fillvalue=fv)
                want = oracle.interpND(src, idx, method=m, bound=b, fillvalue=fv)
                same = (got == want) | (np.isnan(got) & np.isnan(want))
                assert same.all(), (m, b, fv, np.argwhere(~same)[:5])


def test_image_cube_planes(tfm, oracle):
    cube = np.random.default_rng(7).random((3, 40, 50)).astype(np.float32)
    tr = tfm.Composition([tfm.Rotation(0.3), tfm.Offset([-25.0, -20.0])])
    got = tr.resample(cube, method='linear')
    for k in range(3):
        want = oracle.resample(to_oracle(tr), cube[k], method='linear')
        assert np.array_equal(got[k], want)


def test_full_size_properties(tfm):
    """BASELINE config 2 size (4096^2): size-independent properties.
    identity map reproduces the image; a pure integer offset is a shifted copy; linearity."""
    import torch
    n = 4096
    g = torch.Generator(device='cuda').manual_seed(0)
    src = torch.rand((n, n), generator=g, device='cuda', dtype=torch.float32)
    out = tfm.Identity().resample(src, method='linear', precision='fp32')
    assert torch.equal(out, src)
    out = tfm.Offset([3.0, -5.0]).resample(src, method='nearest', precision='fp32')
    # out[y, x] = src[y + 5, x - 3]
    assert torch.equal(out[:n - 5, 3:], src[5:, :n - 3])
    assert float(out[n - 5:, :].abs().max()) == 0.0 and float(out[:, :3].abs().max()) == 0.0
    tr = tfm.Composition([tfm.Scale(1.3, dim=2), tfm.Rotation(30, unit='deg'), tfm.Offset([-2048, -2048])])
    a = tr.resample(src, method='linear', precision='fp32')
    b = tr.resample(2 * src, method='linear', precision='fp32')
    assert torch.allclose(b, 2 * a, rtol=1e-6, atol=1e-6)
    assert a.dtype == torch.float32 and a.shape == (n, n)


@pytest.mark.parametrize("dtype", [np.int16, np.uint8, np.int64, np.uint16])
def test_integer_sources_follow_reference_dtype_rules(tfm, oracle, dtype):
    """FITS images are often integers.  Reference rules (helpers.pyx:373-374, 598-611): nearest keeps
    the source dtype unless the truncation pass promotes it with the int64 fill; linear and cubic
    return float64.  Integer sources are carried as float64 on the device (exact below 2^53)."""
    info = np.iinfo(dtype)
    src = np.random.default_rng(3).integers(max(info.min, -30000), min(info.max, 30000), size=(41, 57)).astype(dtype)
    # cubic: the reference forms 3*(r2-r1) in the SOURCE integer dtype (helpers.pyx:774), which wraps
    # around for large int16 values; the device carries integers as float64 (documented deviation), so
    # the cubic comparison uses a range where the reference does not overflow
    src_small = (src // 16).astype(dtype)
    tr = tfm.Composition([tfm.Scale(1.2, dim=2), tfm.Rotation(0.3), tfm.Offset([-20.0, -15.0])])
    for m in ('nearest', 'linear', 'cubic'):
        for b in ('t', 'e', 'm'):
            if m == 'cubic' and dtype in (np.uint8, np.uint16):
                continue        # the reference's first Hermite pass wraps around in unsigned arithmetic
            data = src_small if m == 'cubic' else src
            got = tr.resample(data, method=m, bound=b)
            want = oracle.resample(to_oracle(tr), data, method=m, bound=b)
            assert got.dtype == want.dtype, (m, b, got.dtype, want.dtype)
            assert np.array_equal(got, want), (m, b)


def test_noncontiguous_and_odd_pitch_inputs(tfm, oracle):
    """Views with strides and widths that break the texture path's alignment rules must still work
    (the engine stages a contiguous copy; 485 float32 columns are not a multiple of 32 bytes)."""
    base = np.random.default_rng(4).random((64, 970)).astype(np.float32)
    view = base[:, ::2]                                     # 64 x 485, strided
    tr = tfm.Composition([tfm.Rotation(10, unit='deg'), tfm.Offset([-200.0, -30.0])])
    for prec, tol in (('fp64', 0.0), ('fp32', 1e-5)):
        got = tr.resample(view, method='linear', precision=prec)
        want = oracle.resample(to_oracle(tr), np.ascontiguousarray(view), method='linear')
        if tol == 0.0:
            assert np.array_equal(got, want)
        else:
            assert_close_rel(got, want, tol, prec)


# ---------------------------------------------------------------------------------------------
# windowed input-plane filters (helpers.pyx:703-735, 792-835)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m", list("szghtl"))
@pytest.mark.parametrize("bound", ['t', 'e', 'p', 'm', ['m', 'p']])
def test_filters_match_oracle(tfm, oracle, m, bound):
    rng = np.random.default_rng(77)
    src = rng.random((61, 83)).astype(np.float32)
    tr = tfm.Composition([tfm.Scale(1.7, dim=2), tfm.Rotation(25, unit='deg'), tfm.Offset([-20.0, -35.0])])
    for ob in (None, 2.2):
        if m == 'l' and ob is None:
            continue
        got = tr.resample(src, method=m, bound=bound, shape=(70, 90), oblur=ob, fillvalue=3)
        want = oracle.resample(to_oracle(tr), src, method=m, bound=bound, shape=(70, 90), oblur=ob, fillvalue=3)
        assert got.dtype == want.dtype == np.float64
        if m == 'l':
            assert np.array_equal(got, want)
        else:
            assert_close_rel(got, want, 1e-12, (m, bound, ob))


def test_filters_fast_mode_and_limits(tfm, oracle):
    rng = np.random.default_rng(78)
    src = rng.random((96, 128)).astype(np.float32)
    tr = tfm.Composition([tfm.Scale(0.8, dim=2), tfm.Rotation(-12, unit='deg'), tfm.Offset([-10.0, 5.0])])
    for m in "szghtl":
        ob = 1.5
        got = tr.resample(src, method=m, bound='t', oblur=ob, precision='fp32')
        want = oracle.resample(to_oracle(tr), src, method=m, bound='t', oblur=ob)
        assert got.dtype == np.float32
        assert_close_rel(got, want, 2e-5, m)
    # oblur = None: sinc and Lanczos weights by recurrence (one sinpi per axis) in fp32 mode
    for m in "sz":
        got = tr.resample(src, method=m, bound='m', precision='fp32')
        want = oracle.resample(to_oracle(tr), src, method=m, bound='m')
        assert_close_rel(got, want, 2e-5, m + " recurrence")
    # a constant image stays constant under every normalised window
    flat = np.full((40, 50), 2.5, dtype=np.float32)
    for m in "szght":
        got = tr.resample(flat, method=m, bound='e')
        assert np.allclose(got, 2.5, rtol=1e-13, atol=0)
    # oblur is rejected where the reference rejects it (helpers.pyx:543-544, 748-749)
    with pytest.raises(ValueError):
        tr.resample(src, method='n', oblur=2.0)
    with pytest.raises(ValueError):
        tr.resample(src, method='c', oblur=2.0)
    # the largest region the kernel holds (sinc, 16*7 = 112), and one beyond it
    big = tr.resample(src, method='s', bound='m', oblur=7.0, shape=(24, 24))
    want = oracle.resample(to_oracle(tr), src, method='s', bound='m', oblur=7.0, shape=(24, 24))
    assert_close_rel(big, want, 1e-11, "sinc oblur 7")
    with pytest.raises(NotImplementedError):
        tr.resample(src, method='s', oblur=7.5)


def test_filters_batched_and_f64(tfm, oracle):
    rng = np.random.default_rng(79)
    src = rng.random((3, 33, 47))
    tr = tfm.Wrap(tfm.Rotation(0.3), tfm.Offset([20.0, 15.0]))
    got = tr.resample(src, method='g', bound='p')
    for k in range(3):
        want = oracle.resample(to_oracle(tr), src[k], method='g', bound='p')
        assert_close_rel(got[k], want, 1e-12, k)


def test_filter_2x2_route_equals_general_kernel(tfm):
    """Hann / Tukey / tent regions of 2 x 2 run inside k_resample; tuning bit 64 forces the general
    filter kernel.  Same arithmetic, same order: identical bits in parity mode."""
    rng = np.random.default_rng(80)
    src = rng.random((50, 70)).astype(np.float32)
    tr = tfm.Composition([tfm.Scale(1.4, dim=2), tfm.Rotation(33, unit='deg'), tfm.Offset([-30.0, -20.0])])
    for m, ob in (("h", None), ("t", None), ("l", 0.7), ("h", 1.0)):
        for bound in ('t', 'p'):
            a = tr.resample(src, method=m, bound=bound, oblur=ob, shape=(64, 96))
            b = tr.resample(src, method=m, bound=bound, oblur=ob, shape=(64, 96), tuning=64)
            assert np.array_equal(a, b, equal_nan=True), (m, ob, bound)
        a = tr.resample(src, method=m, oblur=ob, precision='fp32')
        b = tr.resample(src, method=m, oblur=ob, precision='fp32', tuning=64)
        assert np.allclose(a, b, rtol=0, atol=2e-6), (m, ob)


def test_async_pipeline_matches_synchronous_calls(tfm):
    """transform_b200.AsyncResampler: same kernels on round-robin streams, copies left in flight;
    results must equal the synchronous API bit for bit, in any completion order."""
    import torch
    rng = np.random.default_rng(90)
    tr = tfm.Composition([tfm.Scale(1.2, dim=2), tfm.Rotation(15, unit='deg'), tfm.Offset([-300.0, -200.0])])
    frames = [torch.from_numpy(rng.random((512, 640)).astype(np.float32)).pin_memory() for _ in range(6)]
    outs = [torch.empty((512, 640), dtype=torch.float32).pin_memory() for _ in range(6)]
    pipe = tfm.AsyncResampler(depth=3)
    tickets = []
    for k, (f, o) in enumerate(zip(frames, outs)):
        m = ("linear", "Gaussian", "cubic")[k % 3]
        tickets.append((m, pipe.submit(tr, f, o, method=m, precision="fp32")))
    for k in (5, 0, 3, 1, 4, 2):
        m, tk = tickets[k]
        got = tk.result()
        want = tr.resample(frames[k], method=m, precision="fp32", out=torch.empty_like(outs[k]).pin_memory())
        assert torch.equal(got, want), (k, m)
        assert tk.done()
    pipe.drain()
    with pytest.raises(ValueError):
        pipe.submit(tr, torch.zeros((4, 4)), outs[0])        # pageable input
    with pytest.raises(ValueError):
        tr.resample(frames[0], method="linear", precision="fp32", out=torch.empty((512, 640)), sync=False)


def test_tma_staged_variant_is_bit_identical(tfm):
    """tuning bit 256: the TMA-staged shared-memory variant of the fp32 linear kernel (affine maps).
    Zero-filled out-of-range box parts are the truncate boundary; results equal the texture variant."""
    import torch
    rng = np.random.default_rng(91)
    src = torch.from_numpy(rng.random((700, 896)).astype(np.float32)).cuda()   # pitch % 32 == 0
    for tr in (tfm.Offset([3.25, -7.5]), tfm.Scale(0.45, dim=2),
               tfm.Composition([tfm.Scale(1.3, dim=2), tfm.Rotation(30, unit='deg'), tfm.Offset([-450.0, -350.0])]),
               tfm.Wrap(tfm.Rotation(45, unit='deg'), tfm.Offset([-450.0, -350.0]))):
        for shape in ((700, 896), (333, 517)):
            a = tr.resample(src, method='linear', precision='fp32', shape=shape)
            b = tr.resample(src, method='linear', precision='fp32', shape=shape, tuning=256)
            assert torch.equal(a, b), (str(tr), shape)
    assert tfm._lib.load().tfm_last_kernel().decode().startswith("k_resample_tma")


def test_c_example_matches_python_path(tfm, tmp_path):
    """examples/c_abi_example.c drives tfm_resample from plain C; same chain through the Python
    shim must give the same float64 value."""
    import subprocess
    from test_host_logic import _build_c_example
    exe = _build_c_example(tmp_path)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0, r.stdout
    val = float(r.stdout.strip().split("=")[-1])
    h, w = 96, 128
    src = ((np.arange(h * w) % 251) / np.float32(251.0)).astype(np.float32).reshape(h, w)
    tr = tfm.Composition([tfm.Scale(1.3, dim=2), tfm.Rotation(30, unit='deg'), tfm.Offset([-w / 2.0, -h / 2.0])])
    out = tr.resample(src, method='linear')
    assert abs(out[h // 2, w // 2] - val) <= 1e-12 * max(1.0, abs(val)), (val, out[h // 2, w // 2], r.stdout)


def test_device_out_is_validated(tfm):
    """A CUDA tensor passed as out= must be exactly what the kernel is about to write: the kernel takes
    pitch and element size from the EXPECTED layout, so a float32 buffer in parity mode (float64 output),
    a sliced view, a wrong shape or another leading axis would be written past its end or with the wrong
    stride.  Each is rejected with ValueError, for K1 and K2."""
    import torch
    src = torch.rand((40, 56), device="cuda", dtype=torch.float32)
    tr = tfm.Composition([tfm.Rotation(0.2), tfm.Offset([-20.0, -10.0])])
    for method in ("linear", "Gaussian"):
        good64 = torch.empty((40, 56), device="cuda", dtype=torch.float64)
        assert tr.resample(src, method=method, out=good64) is good64
        good32 = torch.empty((40, 56), device="cuda", dtype=torch.float32)
        assert tr.resample(src, method=method, precision="fp32", out=good32) is good32
        with pytest.raises(ValueError):                          # dtype: parity mode writes float64
            tr.resample(src, method=method, out=torch.empty((40, 56), device="cuda", dtype=torch.float32))
        with pytest.raises(ValueError):                          # dtype: fast mode writes float32
            tr.resample(src, method=method, precision="fp32", out=good64)
        with pytest.raises(ValueError):                          # shape differs from the output grid
            tr.resample(src, method=method, precision="fp32", out=torch.empty((40, 60), device="cuda"))
        with pytest.raises(ValueError):                          # non-contiguous view
            tr.resample(src, method=method, precision="fp32",
                        out=torch.empty((40, 112), device="cuda", dtype=torch.float32)[:, ::2])
        with pytest.raises(ValueError):                          # leading axis mismatch
            tr.resample(src, method=method, precision="fp32", out=torch.empty((2, 40, 56), device="cuda"))


def test_host_out_with_device_input_and_submit_many(tfm):
    """A CUDA frame with a host out= buffer fills the buffer (it used to be left untouched), and
    AsyncResampler.submit_many uploads a pinned frame once for several resamplings."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(5)
    src = torch.rand((300, 420), generator=g, device="cuda", dtype=torch.float32)
    tr = tfm.Composition([tfm.Scale(1.2, dim=2), tfm.Rotation(0.3), tfm.Offset([-200.0, -150.0])])
    rad = tfm.Composition([tfm.Scale([420 / (2 * np.pi), 1.5]), tfm.Radial(origin=[209.5, 149.5])])
    want_l = tr.resample(src, method="linear", precision="fp32")
    want_g = rad.resample(src, method="Gaussian", precision="fp32")
    host = np.zeros((300, 420), dtype=np.float32)
    ret = tr.resample(src, method="linear", precision="fp32", out=host)
    assert ret is host and np.array_equal(host, want_l.cpu().numpy())
    pipe = tfm.AsyncResampler(depth=2)
    h_src = torch.empty((300, 420), dtype=torch.float32).pin_memory()
    h_src.copy_(src)
    outs = [torch.zeros((300, 420), dtype=torch.float32).pin_memory() for _ in range(2)]
    tk = pipe.submit_many(h_src, [(tr, outs[0], "linear", {"precision": "fp32"}),
                                  (rad, outs[1], "Gaussian", {"precision": "fp32"})])
    res = tk.result()
    assert res[0] is outs[0] and torch.equal(outs[0], want_l.cpu()) and torch.equal(outs[1], want_g.cpu())
    # a 'forbid' boundary violated inside the pipeline surfaces at result(), not at submit
    far = tfm.Offset([1000.0, 0.0])
    tk = pipe.submit(far, h_src, outs[0], method="linear", bound="f", precision="fp32")
    with pytest.raises(ValueError):
        tk.result()
    pipe.drain()


def test_antialias_switch_through_resample(tfm, oracle):
    """resample(method='Gaussian', antialias=False) (or aa=False) falls back to the input-plane filter
    of the same letter, like interpND_grid(antialias=False) (helpers.pyx:1524-1530); both keywords are
    consumed whatever their values (the second used to survive and raise TypeError)."""
    src = np.random.default_rng(8).random((48, 64))
    tr = tfm.Composition([tfm.Scale(1.1, dim=2), tfm.Rotation(0.25), tfm.Offset([-30.0, -20.0])])
    want = oracle.resample(to_oracle(tr), src, method='g')
    for kw in ({"antialias": False}, {"aa": False}, {"antialias": False, "aa": True}, {"antialias": True, "aa": False}):
        got = tr.resample(src, method='Gaussian', jump_detect=4.0, **kw)
        assert_close_rel(got, want, 1e-12, str(kw))
    on = tr.resample(src, method='Gaussian', antialias=True, aa=True)
    assert not np.allclose(on, want)
    with pytest.raises(TypeError):
        tr.resample(src, method='Gaussian', antialias=False, bogus=1)


def test_repeated_device_calls_reuse_their_plan(tfm):
    """Repeated device-to-device calls reuse the argument block of the first call (_engine._PLANS).  The reuse
    must be invisible: changing the transform's parameters, the tensors or any option produces the result of a
    fresh call; a hit gives the bits of the first call."""
    import torch
    from transform_b200 import _engine
    n = 257
    g = torch.Generator(device="cuda").manual_seed(9)
    src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
    out = torch.empty_like(src)
    tr = tfm.Composition([tfm.Scale(1.3, dim=2), tfm.Rotation(30, unit='deg'), tfm.Offset([-n / 2.0, -n / 2.0])])
    _engine._PLANS.clear()
    for method in ("linear", "Gaussian"):
        first = tr.resample(src, method=method, precision="fp32", out=out).clone()
        nplans = len(_engine._PLANS)
        assert nplans >= 1
        again = tr.resample(src, method=method, precision="fp32", out=out)
        assert again is out and torch.equal(again, first) and len(_engine._PLANS) == nplans       # a hit
        # the source changes in place: same plan, new result
        src2 = src.clone()
        src.mul_(0.5)
        half = tr.resample(src, method=method, precision="fp32", out=out).clone()
        assert len(_engine._PLANS) == nplans
        assert torch.allclose(half, 0.5 * first, rtol=1e-6, atol=1e-7)
        src.copy_(src2)
        # another option, another tensor, mutated parameters: each a fresh call with the right answer
        other = tr.resample(src, method=method, precision="fp32", out=out, bound='e').clone()
        assert torch.equal(other, tr.resample(src.clone(), method=method, precision="fp32", bound='e'))
        tr.params['list'][1].params['matrix'][0, 1] *= 1.5        # in-place edit of a member's matrix
        edited = tr.resample(src, method=method, precision="fp32", out=out).clone()
        fresh = tr.resample(src.clone(), method=method, precision="fp32")
        assert torch.equal(edited, fresh) and not torch.equal(edited, first)
        tr.params['list'][1].params['matrix'][0, 1] /= 1.5
