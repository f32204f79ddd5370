"""GPU: SIP / TPV distortions (TFM_OP_SIP, TFM_OP_TPV), CEA's lambda and the slant orthographic projection through K3 (apply / invert), K1 (resample) and K2
(Jacobian resample) -- SURVEY.md section 8f rank 4.

PARITY UNPINNED by the reference (its WCS arithmetic is astropy's, absent: core.py:1727, 1742).  The
CUDA op (nested Horner form, per-point fixed-point inversion) is checked against the oracle's
independent restatement (explicit power sums, batch iteration), itself checked on the CPU against hand
values and a root finder (tests/test_sip.py)."""
import numpy as np
import pytest

from conftest import to_oracle, assert_close_rel
from test_sip import sip_header

pytestmark = pytest.mark.gpu


def test_sip_apply_invert(tfm, oracle):
    """all_pix2world / all_world2pix of a TAN-SIP header over a 4096^2 detector: world coordinates to
    1e-10 degree (4e-7 arcsec), pixels back to 1e-7; the kernel that ran is the streaming apply."""
    from transform_b200 import _lib
    a = tfm.WCS(sip_header())
    assert a.otype == ['RA---TAN-SIP', 'DEC--TAN-SIP']
    oa = to_oracle(a)
    p = np.random.default_rng(1).random((20000, 2)) * 4096 - 0.5
    w = a.apply(p)
    assert _lib.last_kernel().startswith("k_apply")
    ow = oracle.apply(oa, p)
    dlon = ((w[:, 0] - ow[:, 0] + 180.0) % 360.0 - 180.0) * np.cos(np.radians(ow[:, 1]))
    assert np.abs(dlon).max() < 1e-10 and np.abs(w[:, 1] - ow[:, 1]).max() < 1e-10
    # the distortion is really in: the undistorted header lands up to ~50 px * 1.4e-5 deg away
    plain = tfm.WCS({k: v for k, v in sip_header().items() if not k[:2] in ('A_', 'B_')})
    assert np.abs(plain.apply(p)[:, 1] - w[:, 1]).max() > 2e-4
    back = a.invert(w)
    assert np.abs(back - p).max() < 1e-7
    assert np.abs(back - oracle.apply(oa, ow, invert=True)).max() < 1e-7
    # extra components pass through (core.py:268-292), NaN stays NaN and poisons nothing else
    q = np.concatenate([p[:50], np.arange(50.0)[:, None]], axis=1)
    q[7, 0] = np.nan
    r = a.invert(a.apply(q))
    assert np.isnan(r[7, :2]).all() and np.array_equal(r[:, 2], q[:, 2])
    ok = np.arange(50) != 7
    assert np.abs(r[ok, :2] - q[ok, :2]).max() < 1e-7
    # the coefficient table is uploaded once per object and device
    assert len(a.params['wcs']._sip_dev) == 1
    # float32 vectors (fp32 storage, fast arithmetic): pixel scale
    w32 = a.apply(p[:2000].astype(np.float32), precision='fp32')
    assert w32.dtype == np.float32 and np.abs(w32 - ow[:2000]).max() < 2e-4


@pytest.mark.parametrize("order", [2, 5, 9])
def test_sip_orders(tfm, oracle, order):
    """Every order up to the convention's maximum (9) through the same op: random coefficients scaled so
    that each term moves the corner of a 1024^2 frame by a few pixels."""
    rng = np.random.default_rng(order)
    hdr = {'NAXIS': 2, 'NAXIS1': 1024, 'NAXIS2': 1024, 'CTYPE1': 'RA---TAN-SIP', 'CTYPE2': 'DEC--TAN-SIP',
           'CRPIX1': 512.5, 'CRPIX2': 512.5, 'CRVAL1': 10.0, 'CRVAL2': -35.0, 'CDELT1': -1e-4, 'CDELT2': 1e-4,
           'A_ORDER': order, 'B_ORDER': order}
    for L_ in 'AB':
        for p in range(order + 1):
            for q in range(order + 1 - p):
                if p + q >= 2:
                    hdr['%s_%d_%d' % (L_, p, q)] = float(rng.normal() * 2.0 / 512.0 ** (p + q))
    a = tfm.WCS(hdr)
    oa = to_oracle(a)
    pix = rng.random((5000, 2)) * 1024 - 0.5
    w, ow = a.apply(pix), oracle.apply(oa, pix)
    assert np.abs(w - ow).max() < 1e-10
    back, oback = a.invert(w), oracle.apply(oa, ow, invert=True)
    ok = np.abs(oracle.apply(oa, oback) - ow).max(axis=1) < 1e-9        # points where the iteration converges
    assert ok.mean() > 0.7
    assert np.abs(back[ok] - oback[ok]).max() < 1e-7


def test_sip_resample(tfm, oracle):
    """A TAN-SIP frame resampled onto the undistorted TAN grid (what remap does with a template): the output
    grid goes world -> pixel through the SIP inversion inside K1; parity mode 1e-9, fp32 mode 2e-5; and the
    anti-aliased method through K2 (finite-difference Jacobian of the SIP chain)."""
    n = 160
    # 0.8 arcmin pixels: world coordinates of ~150 degrees round at 3e-14 degree = 2e-12 pixel, so that the two
    # derivations of the spherical part (unit vectors here, angles in the oracle) agree far below the tolerance
    sip = sip_header(NAXIS1=n, NAXIS2=n, CRPIX1=n / 2 + 0.5, CRPIX2=n / 2 - 3.25,
                     CD1_1=-1.39e-2, CD1_2=1.1e-3, CD2_1=1.2e-3, CD2_2=1.38e-2)
    for k in list(sip):                                                  # same distortion in px on a 160^2 frame
        if k[:2] in ('A_', 'B_') and not k.endswith('ORDER'):
            p, q = int(k[2]), int(k[4])
            sip[k] = sip[k] * (4096.0 / n) ** (p + q - 1)
    plain = {k: v for k, v in sip.items() if not k[:2] in ('A_', 'B_')}
    plain['CTYPE1'], plain['CTYPE2'] = 'RA---TAN', 'DEC--TAN'
    frame = np.random.default_rng(3).random((n, n)).astype(np.float32)
    total = tfm.Composition([tfm.WCS(plain).inverse(), tfm.WCS(sip)])
    ot = to_oracle(total)
    idx = ot.reverse(oracle.pixel_grid(n, n))
    moved = np.hypot(*(idx - oracle.pixel_grid(n, n)).reshape(-1, 2).T)
    assert 3.0 < moved.max() < 30.0                                      # the map is a real distortion
    want = oracle.resample(ot, frame, method='linear', shape=(n, n))
    got = total.resample(frame, method='linear', shape=(n, n))
    assert_close_rel(got, want, 1e-9, "parity")
    fast = total.resample(frame, method='linear', shape=(n, n), precision='fp32')
    assert_close_rel(fast, want, 2e-5, "fp32")
    out = tfm.Identity().remap((frame, sip), template=plain, method='linear')
    assert_close_rel(out.data, got, 1e-15, "remap == composed chain")
    # (sic) the input header's SIP keywords survive in the output header: wcs2head removes the linear terms
    # only (core.py:1229 "TODO: Find nonlinear terms and remove them also")
    assert out.header['CTYPE1'] == 'RA---TAN' and out.header['A_ORDER'] == 3
    back = tfm.Identity().remap((frame, plain), template=sip, method='linear')       # onto the distorted grid
    # (sic) astropy's default to_header() (core.py:1208) writes standard keywords only and strips '-SIP'
    assert 'A_ORDER' not in back.header and back.header['CTYPE1'] == 'RA---TAN'
    tb = tfm.Composition([tfm.WCS(sip).inverse(), tfm.WCS(plain)])
    wantb = oracle.resample(to_oracle(tb), frame, method='linear', shape=(n, n))
    assert_close_rel(back.data, wantb, 1e-9, "onto the SIP grid")
    wantj = oracle.resample_jacobian(ot, frame, method='g', shape=(n, n))
    gotj = total.resample(frame, method='G', shape=(n, n))
    assert_close_rel(gotj, wantj, 1e-9, "jacobian parity")
    fastj = total.resample(frame, method='G', shape=(n, n), precision='fp32')
    assert_close_rel(fastj, wantj, 2e-4, "jacobian fp32")


def test_cea_lambda(tfm, oracle):
    """PV2_1 of the cylindrical equal-area projection (paper II eq. 80-81) and LONPOLE given as PV1_3."""
    hdr = {'NAXIS': 2, 'NAXIS1': 3600, 'NAXIS2': 2400, 'CTYPE1': 'CRLN-CEA', 'CTYPE2': 'CRLT-CEA', 'CUNIT1': 'deg',
           'CUNIT2': 'deg', 'CRPIX1': 1800.5, 'CRPIX2': 1200.5, 'CRVAL1': 180.0, 'CRVAL2': 0.0, 'CDELT1': 0.1,
           'CDELT2': 0.1, 'PV2_1': 0.6, 'PV1_3': 5.0}
    a = tfm.WCS(hdr)
    oa = to_oracle(a)
    assert oa.cea_lambda == 0.6 and oa.phi_p == 5.0
    p = np.random.default_rng(2).random((8000, 2)) * [3400, 2400] + [100, 0]     # clear of the +-180 degree seam
    w, ow = a.apply(p), oracle.apply(oa, p)
    assert np.array_equal(np.isnan(w), np.isnan(ow))
    ok = ~np.isnan(ow).any(axis=1)
    assert ok.mean() > 0.6                                               # rows beyond |sin(theta)| = 1 are off the sphere
    dlon = ((w[ok, 0] - ow[ok, 0] + 180.0) % 360.0 - 180.0) * np.cos(np.radians(ow[ok, 1]))
    assert np.abs(dlon).max() < 1e-10 and np.abs(w[ok, 1] - ow[ok, 1]).max() < 1e-10
    assert np.abs(a.invert(w[ok]) - p[ok]).max() < 1e-7
    one = tfm.WCS(dict(hdr, PV2_1=1.0))
    assert np.nanmax(np.abs(one.apply(p)[:, 1] - w[:, 1])) > 1.0         # lambda matters


def test_tpv_apply_invert_resample(tfm, oracle):
    """A '-TPV' header (SCAMP's polynomial on the intermediate world coordinates, all 40 slots in use somewhere):
    K3 both ways over the 2048^2 detector against the oracle's explicit term list, then K1 / K2 resampling of the
    frame onto the undistorted TAN grid.  PARITY UNPINNED (wcslib absent)."""
    from test_sip import tpv_header
    from transform_b200 import _lib
    a = tfm.WCS(tpv_header())
    oa = to_oracle(a)
    ops = a._chain(False)
    assert [op.kind for op in ops] == [_lib.OP_WCS, _lib.OP_TPV, _lib.OP_WCS]
    assert [op.kind for op in a._chain(True)] == [_lib.OP_WCS, _lib.OP_TPV, _lib.OP_WCS]
    p = np.random.default_rng(21).random((20000, 2)) * 2048 - 0.5
    w, ow = a.apply(p), oracle.apply(oa, p)
    dlon = ((w[:, 0] - ow[:, 0] + 180.0) % 360.0 - 180.0) * np.cos(np.radians(ow[:, 1]))
    assert np.abs(dlon).max() < 1e-10 and np.abs(w[:, 1] - ow[:, 1]).max() < 1e-10
    plain = tfm.WCS({k: v for k, v in tpv_header().items() if not k.startswith('PV')})
    assert 5e-5 < np.abs(plain.apply(p)[:, 1] - w[:, 1]).max() < 2e-3       # a fraction of an arcsecond .. arcseconds
    back = a.invert(w)
    assert np.abs(back - p).max() < 1e-6
    assert np.abs(back - oracle.apply(oa, ow, invert=True)).max() < 1e-6
    # resampling: a coarser plate scale so that world-coordinate rounding stays far below the tolerance
    n = 144
    hdr = tpv_header(NAXIS1=n, NAXIS2=n, CRPIX1=n / 2 + 0.5, CRPIX2=n / 2 + 2.0, CD1_1=-3.0e-3, CD1_2=0.0,
                     CD2_1=0.0, CD2_2=3.0e-3)
    tan = {k: v for k, v in hdr.items() if not k.startswith('PV')}
    tan['CTYPE1'], tan['CTYPE2'] = 'RA---TAN', 'DEC--TAN'
    frame = np.random.default_rng(22).random((n, n)).astype(np.float32)
    total = tfm.Composition([tfm.WCS(tan).inverse(), tfm.WCS(hdr)])
    ot = to_oracle(total)
    g = oracle.pixel_grid(n, n)
    moved = np.hypot(*(ot.reverse(g) - g).reshape(-1, 2).T)
    assert 1.0 < moved.max() < 20.0                                      # 0.43-degree field: the high orders bite
    want = oracle.resample(ot, frame, method='linear', shape=(n, n))
    assert_close_rel(total.resample(frame, method='linear', shape=(n, n)), want, 1e-9, "parity")
    assert_close_rel(total.resample(frame, method='linear', shape=(n, n), precision='fp32'), want, 2e-5, "fp32")
    wantj = oracle.resample_jacobian(ot, frame, method='g', shape=(n, n))
    assert_close_rel(total.resample(frame, method='G', shape=(n, n)), wantj, 1e-9, "jacobian parity")
    assert_close_rel(total.resample(frame, method='G', shape=(n, n), precision='fp32'), wantj, 2e-4, "jacobian fp32")
    out = tfm.Identity().remap((frame, hdr), template=tan, method='linear')
    assert_close_rel(out.data, total.resample(frame, method='linear', shape=(n, n)), 1e-15, "remap == composed chain")


@pytest.mark.parametrize("slant", [(0.2, -0.3), (0.0, 0.8390996311772799), (-0.05, 0.0)])
def test_slant_orthographic(tfm, oracle, slant):
    """SIN with (xi, eta) = (PV2_1, PV2_2) (paper II eq. 61-62; the middle case is the NCP projection of a field at
    declination 50 degrees): a 46-degree field so that the slant terms matter, both ways, NaN beyond the limb."""
    hdr = {'NAXIS': 2, 'NAXIS1': 2048, 'NAXIS2': 2048, 'CTYPE1': 'RA---SIN', 'CTYPE2': 'DEC--SIN', 'CUNIT1': 'deg',
           'CUNIT2': 'deg', 'CRPIX1': 1024.5, 'CRPIX2': 1024.5, 'CRVAL1': 83.0, 'CRVAL2': 50.0, 'CDELT1': -0.0225,
           'CDELT2': 0.0225, 'PV2_1': slant[0], 'PV2_2': slant[1]}
    a = tfm.WCS(hdr)
    oa = to_oracle(a)
    assert oa.sin_slant == (slant[0], slant[1])
    p = np.random.default_rng(23).random((8000, 2)) * 2048
    w, ow = a.apply(p), oracle.apply(oa, p)
    assert np.array_equal(np.isnan(w), np.isnan(ow))
    ok = ~np.isnan(ow).any(axis=1)
    assert ok.mean() > 0.9
    dlon = ((w[ok, 0] - ow[ok, 0] + 180.0) % 360.0 - 180.0) * np.cos(np.radians(ow[ok, 1]))
    assert np.abs(dlon).max() < 1e-10 and np.abs(w[ok, 1] - ow[ok, 1]).max() < 1e-10
    back = a.invert(w[ok])
    assert np.abs(back - p[ok]).max() < 1e-7
    straight = tfm.WCS({k: v for k, v in hdr.items() if not k.startswith('PV')})
    assert np.nanmax(np.abs(straight.apply(p)[:, 1] - w[:, 1])) > 0.05       # the slant is really in
    # the far side of the sky is NaN in the world -> pixel direction, as in the oracle
    sky = np.stack((np.random.default_rng(24).random(4000) * 360, np.random.default_rng(25).random(4000) * 180 - 90), -1)
    bp, obp = a.invert(sky), oracle.apply(oa, sky, invert=True)
    assert np.array_equal(np.isnan(bp), np.isnan(obp)) and 0.2 < np.isnan(obp[:, 0]).mean() < 0.8
    vis = ~np.isnan(obp[:, 0])
    assert np.abs(bp[vis] - obp[vis]).max() < 1e-7
