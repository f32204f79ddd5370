"""GPU: Transform.remap (core.py:589-930) -- the reference's remap KATs, BASELINE config 1
(Rotation(45 deg).remap of a sample.fits-shaped frame, linear WCS) and config 4 (TAN -> CAR batch)."""
import numpy as np
import pytest

from conftest import to_oracle, assert_close_rel

pytestmark = pytest.mark.gpu

AHDR = {'SIMPLE': 'T', 'NAXIS': 2, 'NAXIS1': 7, 'NAXIS2': 7, 'CRPIX1': 4, 'CRPIX2': 4, 'CRVAL1': 0, 'CRVAL2': 0,
        'CDELT1': 1, 'CDELT2': 1, 'CUNIT1': 'pixel', 'CUNIT2': 'pixel', 'CTYPE1': 'X', 'CTYPE2': 'Y'}

# the WCS of the reference's sample.fits: 485 x 512, CD matrix, linear axes
SAMPLE_HDR = {'SIMPLE': 'T', 'BITPIX': -64, 'NAXIS': 2, 'NAXIS1': 485, 'NAXIS2': 512, 'CTYPE1': 'Solar-X',
              'CTYPE2': 'Solar-Y', 'CUNIT1': 'arcsec', 'CUNIT2': 'arcsec', 'CRPIX1': 243, 'CRPIX2': 256.5,
              'CRVAL1': -178.95728450798, 'CRVAL2': -469.237931837451, 'CD1_1': 0.831675615713575,
              'CD1_2': 0.0232308034180093, 'CD2_1': -0.0232308034180093, 'CD2_2': 0.831675615713575}


def _cross():
    a = np.zeros([7, 7])
    a[1:5, 3] = 1
    a[3, 0:5] = 1
    return a


def test_reference_remap_kats(tfm):
    """tests/test_01_core.py:297-394 of the reference."""
    a = _cross()
    b = tfm.Identity().remap({'data': a, 'header': AHDR}, method='nearest')
    assert np.all(b['data'] == a)
    b = tfm.PlusOne_().remap({'data': a, 'header': AHDR}, method='nearest')
    assert np.all(a == b['data'])
    assert b['header']['CRPIX1'] == 4 and b['header']['CRPIX2'] == 4
    assert b['header']['CRVAL1'] == 1 and b['header']['CRVAL2'] == 0
    assert b['header']['CDELT1'] == 1 and b['header']['CDELT2'] == 1
    b = tfm.Scale(3, dim=2).remap({'data': a, 'header': AHDR}, method='nearest')
    assert np.all(a == b['data'])
    assert b['header']['CRPIX1'] == 4 and b['header']['CRVAL1'] == 0
    assert np.isclose(b['header']['CDELT1'], 3, atol=1e-10) and np.isclose(b['header']['CDELT2'], 3, atol=1e-10)
    b = tfm.Identity().remap({'data': a, 'header': AHDR}, method='nearest', output_range=[[-3.5, 3.5], [-3.5, 3.5]])
    assert np.all(a == b['data']) and np.isclose(b['header']['CDELT1'], 1, atol=1e-10)
    expando = np.array([[0, 0, 1, 1, 1, 0, 0], [0, 0, 1, 1, 1, 0, 0], [1, 1, 1, 1, 1, 1, 1], [1, 1, 1, 1, 1, 1, 1],
                        [1, 1, 1, 1, 1, 1, 1], [0, 0, 1, 1, 1, 0, 0], [0, 0, 1, 1, 1, 0, 0]])
    b = tfm.Identity().remap({'data': a, 'header': AHDR}, method='nearest',
                             output_range=[[-3.5 / 3, 3.5 / 3], [-3.5 / 3, 3.5 / 3]])
    assert np.all(np.isclose(expando, b['data'], atol=1e-10))
    assert b['header']['CRPIX1'] == 4 and b['header']['CRVAL2'] == 0
    assert np.isclose(b['header']['CDELT1'], 1 / 3.0, atol=1e-10)
    b = tfm.Scale(3, dim=2).remap({'data': a, 'header': AHDR}, method='nearest',
                                  output_range=[[-3.5, 3.5], [-3.5, 3.5]])
    assert np.all(np.isclose(expando, b['data'], atol=1e-10)) and np.isclose(b['header']['CDELT2'], 1, atol=1e-10)


def test_notebook_remap_header_kat(tfm):
    """docs/Transform-docs.ipynb cell 31: Composition([Rotation(30 deg), Scale(10)]).remap(7x7,
    shape=[70,70]) -> CRPIX 35.5, CDELT 1.3660254037844, CRVAL 0."""
    a = _cross()
    tr = tfm.Composition([tfm.Rotation(30, unit='deg'), tfm.Scale(10)])
    b = tr.remap({'data': a, 'header': AHDR}, shape=[70, 70], method='linear')
    assert b['data'].shape == (70, 70)
    assert b['header']['CRPIX1'] == 35.5 and b['header']['CRPIX2'] == 35.5
    assert np.isclose(b['header']['CDELT1'], 1.3660254037844, atol=1e-12)
    assert np.isclose(b['header']['CRVAL1'], 0, atol=1e-12)


def test_config1_rotation45_remap_linear(tfm, oracle):
    """BASELINE config 1: t.Rotation(45,'deg').remap(sample.fits, method='linear').  The frame here
    has sample.fits's shape and WCS with synthetic data (the reference's file does not travel to
    the GPU box).  Autoscaled output WCS: the values derived in SURVEY.md section 8d by following
    core.py:874-902 literally, axis-order mix included."""
    data = np.random.default_rng(11).random((512, 485))
    out = tfm.Rotation(45, unit='deg').remap((data, SAMPLE_HDR), method='linear')
    h = out.header
    assert out.data.shape == (512, 485) and out.data.dtype == np.float64
    assert np.isclose(h['CRPIX1'], 256.5) and np.isclose(h['CRPIX2'], 243.0)
    assert np.isclose(h['CRVAL1'], 221.137667800, atol=1e-6) and np.isclose(h['CRVAL2'], -458.786752794, atol=1e-6)
    assert np.isclose(h['CDELT1'], 1.146020970701, atol=1e-9) and np.isclose(h['CDELT2'], 1.207991128727, atol=1e-9)
    # pixel values: the composed chain out_wcs^-1 o Rotation o in_wcs through the oracle
    total = tfm.Composition([tfm.WCS(out.wcs).inverse(), tfm.Rotation(45, unit='deg'), tfm.WCS(SAMPLE_HDR)])
    want = oracle.resample(to_oracle(total), data, method='linear', shape=(512, 485))
    assert_close_rel(out.data, want, 1e-12, "config 1")
    assert np.count_nonzero(out.data) > 0.3 * out.data.size


def test_config4_tan_to_car_batch(tfm, oracle):
    """BASELINE config 4 in miniature: a batch of TAN frames remapped onto a CAR template,
    Identity().remap(frame, template=hdr_car, method='linear').  PARITY UNPINNED by the reference
    (wcslib absent): checked against the oracle's independent restatement of paper II."""
    n = 96
    tan = {'NAXIS': 2, 'NAXIS1': n, 'NAXIS2': n, 'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN', 'CUNIT1': 'deg',
           'CUNIT2': 'deg', 'CRPIX1': n / 2 + 0.5, 'CRPIX2': n / 2 + 0.5, 'CRVAL1': 150.0, 'CRVAL2': 20.0,
           'CDELT1': -0.02, 'CDELT2': 0.02}
    car = dict(tan, CTYPE1='RA---CAR', CTYPE2='DEC--CAR')
    frames = np.random.default_rng(12).random((3, n, n)).astype(np.float32)
    out0 = tfm.Identity().remap((frames[0], tan), template=car, method='linear')
    assert out0.data.shape == (n, n) and out0.header['CTYPE1'] == 'RA---CAR'
    total = tfm.Composition([tfm.WCS(car).inverse(), tfm.WCS(tan)])
    batch = total.resample(frames, method='linear', shape=(3, n, n))          # frames batched over planes
    for k in range(3):
        want = oracle.resample(to_oracle(total), frames[k], method='linear', shape=(n, n))
        assert_close_rel(batch[k], want, 1e-10, "frame %d" % k)
    assert_close_rel(out0.data, batch[0], 1e-15, "remap == composed chain")
    assert np.count_nonzero(batch[0]) > 0.8 * n * n                           # >= 80 % overlap


@pytest.mark.parametrize("src_proj,dst_proj", [("ARC", "CEA"), ("STG", "AIT"), ("SIN", "ZEA"), ("MER", "SFL")])
def test_remap_between_added_projections(tfm, oracle, src_proj, dst_proj):
    """the other projections of paper II through the same remap path
    (parity and fp32 mode; the fused TAN/CAR pair op must not be applied to these)."""
    n = 80
    a = {'NAXIS': 2, 'NAXIS1': n, 'NAXIS2': n, 'CTYPE1': 'HPLN-' + src_proj, 'CTYPE2': 'HPLT-' + src_proj,
         'CUNIT1': 'deg', 'CUNIT2': 'deg', 'CRPIX1': n / 2 + 0.5, 'CRPIX2': n / 2 + 0.5, 'CRVAL1': 20.0,
         'CRVAL2': 10.0, 'CDELT1': 0.05, 'CDELT2': 0.05}
    b = dict(a, CTYPE1='HPLN-' + dst_proj, CTYPE2='HPLT-' + dst_proj)
    frame = np.random.default_rng(13).random((n, n)).astype(np.float32)
    out = tfm.Identity().remap((frame, a), template=b, method='linear')
    assert out.header['CTYPE1'] == 'HPLN-' + dst_proj
    total = tfm.Composition([tfm.WCS(b).inverse(), tfm.WCS(a)])
    want = oracle.resample(to_oracle(total), frame, method='linear', shape=(n, n))
    assert_close_rel(out.data, want, 1e-9, "parity")
    fast = total.resample(frame, method='linear', shape=(n, n), precision='fp32')
    assert_close_rel(fast, want, 2e-5, "fp32")
    assert np.count_nonzero(out.data) > 0.7 * n * n
    with pytest.raises(NotImplementedError):
        tfm.WCS(dict(a, PV2_3=0.5))                               # a projection parameter outside the GPU path


@pytest.mark.parametrize("rot", [0.0, 17.0])
def test_tan_to_car_fast_path_tables(tfm, oracle, rot):
    """fp32 fast path of the TAN -> CAR remap: the fused WCS pair takes sin/cos of the CAR frame's native
    longitude / latitude from per-tile angle-addition tables (tfm_chain.cuh, wcs_car_tables) -- batch of
    planes (LDG kernel), single frame (texture-gather kernel), an axis-aligned and a rotated CD matrix,
    an output grid that is not a multiple of the tile.  Against the oracle (2e-5, the fp32 tolerance) and
    against the parity mode of the same library (coordinates agree to 1e-9 px)."""
    n, m = 150, 171
    c, s = np.cos(np.radians(rot)), np.sin(np.radians(rot))
    tan = {'NAXIS': 2, 'NAXIS1': m, 'NAXIS2': n, 'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN', 'CUNIT1': 'deg',
           'CUNIT2': 'deg', 'CRPIX1': m / 2 + 0.5, 'CRPIX2': n / 2 + 0.5, 'CRVAL1': 150.0, 'CRVAL2': 20.0,
           'CD1_1': -0.02 * c, 'CD1_2': 0.02 * s, 'CD2_1': 0.02 * s, 'CD2_2': 0.02 * c}
    car = dict(tan, CTYPE1='RA---CAR', CTYPE2='DEC--CAR')
    frames = np.random.default_rng(14).random((3, n, m)).astype(np.float32)
    total = tfm.Composition([tfm.WCS(car).inverse(), tfm.WCS(tan)])
    from transform_b200 import _lib
    batch = total.resample(frames, method='linear', shape=(3, n, m), precision='fp32')
    assert _lib.last_kernel() == "k_resample<trunc>", _lib.last_kernel()
    single = total.resample(np.ascontiguousarray(np.pad(frames[1], ((0, 0), (0, 5)))), method='linear',
                            shape=(n, m), precision='fp32')             # 176 columns: 32-byte pitch -> texture path
    assert _lib.last_kernel() == "k_resample_tex4<car-tan>", _lib.last_kernel()
    for k in range(3):
        want = oracle.resample(to_oracle(total), frames[k], method='linear', shape=(n, m))
        assert_close_rel(batch[k], want, 2e-5, "frame %d rot %g" % (k, rot))
    want1 = oracle.resample(to_oracle(total), np.pad(frames[1], ((0, 0), (0, 5))), method='linear', shape=(n, m))
    assert_close_rel(single, want1, 2e-5, "texture path rot %g" % rot)
    assert np.count_nonzero(batch[0]) > 0.5 * n * m


@pytest.mark.parametrize("kind", ["car-tan", "affine", "polar"])
def test_texture_batch_kernel(tfm, oracle, kind):
    """A batch of planes through the texture path runs k_resample_tex4_batch: a CTA computes the coordinates of
    its tile once and serves up to 8 planes with them (16 texture objects per launch).  Every plane must equal
    the one-launch-per-plane form (tuning bit 1048576) bit for bit -- same coordinates, same gathers -- for plane
    counts that end inside a group of 8 and inside a launch of 16, on an output grid that is not a multiple of
    the tile; and the oracle to the fp32 tolerance."""
    import torch
    from transform_b200 import _lib
    n, m = 160, 384                                              # plane pitch 160 * 384 * 4 bytes: a multiple of 512
    if kind == "car-tan":
        tan = {'NAXIS': 2, 'NAXIS1': m, 'NAXIS2': n, 'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN', 'CUNIT1': 'deg',
               'CUNIT2': 'deg', 'CRPIX1': m / 2 + 0.5, 'CRPIX2': n / 2 + 0.5, 'CRVAL1': 150.0, 'CRVAL2': 20.0,
               'CDELT1': -0.02, 'CDELT2': 0.02}
        tr = tfm.Composition([tfm.WCS(dict(tan, CTYPE1='RA---CAR', CTYPE2='DEC--CAR')).inverse(), tfm.WCS(tan)])
        name = "k_resample_tex4_batch<car-tan>"
    elif kind == "affine":
        tr = tfm.Composition([tfm.Scale(1.3, dim=2), tfm.Rotation(30, unit='deg'), tfm.Offset([-m / 2.0, -n / 2.0])])
        name = "k_resample_tex4_batch<affine>"
    else:
        tr = tfm.Composition([tfm.Scale([m / (2 * np.pi), 0.5]), tfm.Radial(origin=[m / 2.0 - 0.3, n / 2.0 + 0.2])])
        name = "k_resample_tex4_batch"
    g = torch.Generator(device="cuda").manual_seed(77)
    for nz in (2, 8, 11, 17, 33):
        frames = torch.rand((nz, n, m), device="cuda", dtype=torch.float32, generator=g)
        shape = (nz, n - 3, m - 5)
        got = tr.resample(frames, method='linear', precision='fp32', shape=shape)
        assert _lib.last_kernel() == name, (_lib.last_kernel(), nz)
        ref = tr.resample(frames, method='linear', precision='fp32', shape=shape, tuning=1048576)
        assert _lib.last_kernel() == name.replace("_batch", ""), _lib.last_kernel()
        assert torch.equal(got, ref), (kind, nz, float((got - ref).abs().max()))
        for k in (0, nz - 1):
            want = oracle.resample(to_oracle(tr), frames[k].cpu().numpy(), method='linear', shape=shape[1:])
            assert_close_rel(got[k].cpu().numpy(), want, 2e-5, "%s plane %d of %d" % (kind, k, nz))
