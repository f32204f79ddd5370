"""GPU parity: K3 (streaming Transform.apply) against the oracle."""
import numpy as np
import pytest

from conftest import to_oracle, assert_close_rel

pytestmark = pytest.mark.gpu


def _chains(t):
    aff = t.Composition([t.Scale([1.5, 2]), t.Rotation(30, unit='deg'), t.Offset([3, 4])])
    rad = t.Composition([t.Scale([1.5, 2]), t.Radial(origin=[0.1, 0.2]), t.Offset([3, 4])])
    return {"affine": aff, "affine_inv": aff.inverse(), "radial": rad, "radial_inv": rad.inverse(),
            "linear_general": t.Linear(matrix=np.array([[1.2, 0.3], [-0.4, 0.9]]),
                                       pre=np.array([1., 2.]), post=np.array([-3., 4.])),
            "wrap": t.Wrap(t.Rotation(0.4), t.Offset([10.0, -7.0])),
            "radial_log": t.Radial(origin=[0.1, 0.2], r0=3.0, ccw=True, pos_only=False)}


@pytest.mark.parametrize("name", ["affine", "affine_inv", "linear_general", "wrap"])
@pytest.mark.parametrize("n", [0, 1, 7, 1000, 100003])
def test_apply_affine_bitwise(tfm, oracle, name, n):
    tr = _chains(tfm)[name]
    v = np.random.default_rng(2).random((n, 2)) * 8192 - 4096
    for inv in (False, True):
        got = tr.apply(v, invert=inv)
        want = oracle.apply(to_oracle(tr), v, invert=inv)
        assert got.shape == want.shape and got.dtype == np.float64
        assert np.array_equal(got, want)


@pytest.mark.parametrize("name", ["radial", "radial_inv", "radial_log"])
def test_apply_radial(tfm, oracle, name):
    tr = _chains(tfm)[name]
    v = np.random.default_rng(3).random((5000, 2)) * 40 - 20
    if name != "radial":
        v = np.abs(v) * np.array([0.15, 0.5])           # (theta, r) style inputs for inverse chains
    for inv in (False, True):
        got = tr.apply(v, invert=inv)
        want = oracle.apply(to_oracle(tr), v, invert=inv)
        assert_close_rel(got, want, 1e-13, name)


def test_extra_components_pass_through(tfm, oracle):
    """core.py:271-276, 287-292: components beyond the transform's dimension are untouched."""
    tr = _chains(tfm)["affine"]
    v = np.random.default_rng(4).random((6, 5, 4)) * 100
    got = tr.apply(v)
    want = oracle.apply(to_oracle(tr), v)
    assert got.shape == v.shape
    assert np.array_equal(got, want)
    assert np.array_equal(got[..., 2:], v[..., 2:])


def test_roundtrip_and_fast32(tfm):
    tr = _chains(tfm)["radial"]
    v = np.random.default_rng(5).random((4096, 2)) * 200 + 5
    back = tr.invert(tr.apply(v))
    assert np.allclose(back, v, atol=1e-9)
    v32 = v.astype(np.float32)
    got = tr.apply(v32, precision='fp32')
    assert got.dtype == np.float32
    assert np.allclose(got, tr.apply(v), rtol=2e-6, atol=1e-4)


def test_reference_kats(tfm):
    """tests/test_02_basic.py:104-113, 139-153, 209-223 of the reference."""
    assert all(tfm.Scale(np.array([1, 2])).apply(np.array([10, 20])) == np.array([10, 40]))
    assert all(tfm.Scale(3).apply(np.array([5, 50])) == np.array([15, 150]))
    a = tfm.Rotation(90, unit='deg')
    d0 = np.array([[1, 0], [0, 1]])
    d1 = a.apply(d0)
    assert np.all(np.isclose(d1, np.array([[0, 1], [-1, 0]]), atol=1e-15))
    assert np.all(np.isclose(a.invert(d1), d0, atol=1e-15))
    x = np.arange(4.).reshape(2, 2)
    b = tfm.Radial(origin=[130, 130])
    assert np.all(np.isclose(np.array([[131.0, 130.0], [128.75, 127.27]]), b.invert(x), atol=1e-2))
    assert np.all(np.isclose(np.array([[2.3600555, 183.14202], [2.360116, 180.31362]]), b.apply(x), atol=1e-2))
    cd = tfm.Radial(origin=[130, 130], r0=5)
    assert np.allclose(np.array([[2.3600555, 3.600824], [2.360116, 3.5852597]]), cd.apply(x), atol=1e-2)
    assert np.allclose(np.array([[143.59141, 130], [88.207337, 38.681365]]), cd.invert(x), atol=1e-2)
    g = np.mgrid[0:7, 0:7].T - 3
    tr = tfm.Radial()
    bb = tr.apply(g)
    assert np.isclose(bb[0, -1, 0], 0.25 * np.pi, atol=1e-5) and np.isclose(bb[-1, 0, 0], 1.25 * np.pi, atol=1e-5)
    assert np.all(np.isclose(g, tr.invert(bb), atol=1e-10))


def test_direction_and_dimension_errors(tfm):
    with pytest.raises(AssertionError):
        tfm.Scale([0.0, 1.0]).invert(np.zeros((3, 2)))
    with pytest.raises(ValueError):
        tfm.Scale(2.0).apply(np.zeros((3, 1)))


def test_wcs_linear_pin(tfm, oracle):
    """tests/test_01_core.py:160-169: WCS(sample.fits).apply([[0,0]]) ~ [[-386.15825, -676.1092]].
    The header values are those of the reference's sample.fits."""
    hdr = {'NAXIS': 2, 'NAXIS1': 485, 'NAXIS2': 512, 'CTYPE1': 'Solar-X', 'CTYPE2': 'Solar-Y',
           'CUNIT1': 'arcsec', 'CUNIT2': 'arcsec', 'CRPIX1': 243, 'CRPIX2': 256.5,
           'CRVAL1': -178.95728450798, 'CRVAL2': -469.237931837451,
           'CD1_1': 0.831675615713575, 'CD1_2': 0.0232308034180093,
           'CD2_1': -0.0232308034180093, 'CD2_2': 0.831675615713575}
    a = tfm.WCS(hdr)
    assert a.idim == 2 and a.odim == 2
    assert a.itype == ['X', 'Y'] and a.iunit == ['Pixels', 'Pixels']
    assert a.ounit == ['arcsec', 'arcsec'] and a.otype == ['Solar-X', 'Solar-Y']
    got = a.apply([[0, 0]], 0)
    assert np.all(np.isclose(got, np.array([[-386.15825, -676.1092]]), atol=1e-4))
    p = np.random.default_rng(0).random((1000, 2)) * 500
    w = a.apply(p)
    assert np.array_equal(w, oracle.apply(to_oracle(a), p))
    assert np.allclose(a.invert(w), p, atol=1e-9)


@pytest.mark.parametrize("proj", ["TAN", "CAR", "STG", "SIN", "ARC", "ZEA", "CEA", "MER", "SFL", "AIT"])
@pytest.mark.parametrize("crval", [(150.0, 2.2), (10.0, -45.0), (0.3, 0.0), (200.0, 60.0)])
@pytest.mark.parametrize("cdelt", [0.0004, 0.012])
def test_wcs_projections(tfm, oracle, proj, crval, cdelt):
    """Spherical projections: PARITY UNPINNED by the reference (astropy absent); GPU (unit-vector
    form) vs the independent angle-form restatement in the oracle, 1e-11 degree / 1e-8 pixel, on a
    1.6-degree and on a 49-degree field."""
    hdr = {'NAXIS': 2, 'NAXIS1': 4096, 'NAXIS2': 4096, 'CTYPE1': 'RA---' + proj, 'CTYPE2': 'DEC--' + proj,
           'CRPIX1': 2048.5, 'CRPIX2': 2048.5, 'CRVAL1': crval[0], 'CRVAL2': crval[1],
           'CDELT1': -cdelt, 'CDELT2': cdelt, 'CUNIT1': 'deg', 'CUNIT2': 'deg'}
    a = tfm.WCS(hdr)
    p = np.random.default_rng(1).random((4000, 2)) * 4096
    w = a.apply(p)
    ow = oracle.apply(to_oracle(a), p)
    assert np.array_equal(np.isnan(w), np.isnan(ow))
    ok = ~np.isnan(ow).any(axis=1)
    assert ok.mean() > 0.9
    # longitude is ill-conditioned at the celestial poles: compare the arc it subtends
    dlon = ((w[ok, 0] - ow[ok, 0] + 180.0) % 360.0 - 180.0) * np.cos(np.radians(ow[ok, 1]))
    assert np.abs(dlon).max() < 1e-10 and np.abs(w[ok, 1] - ow[ok, 1]).max() < 1e-10
    back = a.invert(w[ok])
    assert np.abs(back - p[ok]).max() < 1e-7
    assert np.abs(back - oracle.apply(to_oracle(a), ow[ok], invert=True)).max() < 1e-7


def test_pincushion_kats_and_modes(tfm, oracle):
    """The reference's own tables (tests/test_02_basic.py:331-419) through the CUDA path -- idim = 0
    transforms act on every component of a 3x3 array -- plus fp32 mode and error behaviour."""
    x = np.arange(9.).reshape(3, 3)
    a = tfm.Quadpin()
    d1 = a.apply(x)
    assert np.allclose(d1, [[0, 1, 2.1818182], [3.5454545, 5.0909091, 6.8181818],
                            [8.7272727, 10.818182, 13.090909]], atol=1e-2)
    assert np.allclose(a.invert(d1), x, atol=1e-2)
    a = tfm.Quadpin(origin=5.0)
    d8 = a.apply(x)
    assert np.allclose(d8, [[-1.8181818, -0.090909091, 1.4545455], [2.8181818, 4, 5],
                            [6, 7.1818182, 8.5454545]], atol=1e-2)
    assert np.allclose(a.invert(d8), x, atol=1e-2)
    a = tfm.Cubicpin(strength=2)
    d1 = a.apply(x)
    assert np.allclose(d1, [[0, 0.6, 3.6], [11.4, 26.4, 51], [87.6, 138.6, 206.4]], atol=1e-2)
    assert np.allclose(a.invert(d1), [[0, 0.79501675, 1.6595098], [2.5116355, 3.3596201, 4.2058325],
                                      [5.0511344, 5.8959087, 6.7403507]], atol=1e-2)
    a = tfm.Cubicpin(length=2.0)
    assert np.allclose(a.apply(x), x, atol=1e-2)
    assert np.isnan(a.invert(x)).all()
    assert np.allclose(tfm.Poly().apply(x), x, atol=1e-2)
    a = tfm.Poly(length=3.0, strength=2.0)
    assert np.allclose(a.apply(x), [[0., 0.55555556, 1.55555556], [3., 4.88888889, 7.22222222],
                                    [10., 13.22222222, 16.88888889]], atol=1e-2)
    with pytest.raises(AssertionError):
        a.invert(x)                                          # Poly has no inverse (basic.py:1296)
    assert str(tfm.Quarticpin()) == "Transform( Poly/Quarticpin )"
    # large random batch, both precisions, against the oracle
    rng = np.random.default_rng(4)
    v = rng.random((100003, 2)) * 300 - 150
    for tr in (tfm.Quadpin(idim=2, odim=2, strength=0.2, length=80.0, origin=np.array([5.0, -3.0])),
               tfm.Cubicpin(idim=2, odim=2, strength=0.3, length=90.0),
               tfm.Quarticpin(idim=2, odim=2, length=120.0)):
        otr = to_oracle(tr)
        got = tr.apply(v)
        assert_close_rel(got, oracle.apply(otr, v), 1e-12, str(tr))
        f32 = tr.apply(v.astype(np.float32), precision='fp32')
        assert f32.dtype == np.float32
        assert_close_rel(f32, oracle.apply(otr, v), 2e-6 * 300, str(tr))
        if not tr.no_reverse:
            assert_close_rel(tr.invert(v), oracle.apply(otr, v, invert=True), 1e-12, str(tr))
    # integer input + dim=: the reference's write-back truncation, reproduced
    iv = np.arange(40).reshape(20, 2)
    q = tfm.Quadpin(dim=2, strength=0.3, length=10.0)
    assert np.array_equal(q.apply(iv), oracle.apply(to_oracle(q), iv))
    assert np.array_equal(q.apply(iv), np.trunc(q.apply(iv.astype(float))))


def test_linear_nd_reference_kats(tfm):
    """The reference's own 3-D Rotation tests (tests/test_02_basic.py:170-186) through the K3-ND kernel."""
    from transform_b200 import _lib
    d0 = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]])
    a = tfm.Rotation([[1, 2, 90], [0, 1, 90]], unit='deg')
    assert a.idim == 3
    d1 = a.apply(d0)
    assert _lib.last_kernel() == "k_apply_linear_nd<f64>"
    assert np.all(np.isclose(d1, np.array([[0, 0, 1], [-1, 0, 0], [0, -1, 0]]), atol=1e-15))
    a = tfm.Rotation([[0, 1, 90], [1, 2, 90]], unit='deg')
    b = tfm.Rotation(euler=np.array([90, 0, 90]), unit='deg')
    d1a, d1b = a.apply(d0), b.apply(d0)
    assert np.all(np.isclose(d1a, d1b, atol=1e-15))
    assert np.all(np.isclose(d1a, np.array([[0, 1, 0], [0, 0, 1], [1, 0, 0]]), atol=1e-15))
    assert np.all(np.isclose(b.invert(d1b), d0, atol=1e-15))
    assert "%s" % a == "Transform( Linear/Rotation )"
    with pytest.raises(ValueError):
        a.apply(np.zeros((4, 2)))                                # core.py:284-285: too few components


@pytest.mark.parametrize("dim", [3, 4, 5, 8])
def test_linear_nd_random_chains(tfm, oracle, dim):
    """Linear-family chains of dimension 3..8 (Scale o Linear o Offset, forward and inverted, pass-through
    components, device tensors) against the oracle: bit for bit in 3-D, where the oracle's matrix-product order
    is pinned against NumPy; within 4 ulp of sum |m_ik v_k| per op beyond."""
    import torch
    rng = np.random.default_rng(100 + dim)
    m = rng.standard_normal((dim, dim)) + 2.0 * np.eye(dim)
    tr = tfm.Composition([tfm.Scale(rng.uniform(0.5, 2.0, size=dim)),
                          tfm.Linear(matrix=m, pre=rng.uniform(-1, 1, size=dim), post=rng.uniform(-1, 1, size=dim)),
                          tfm.Offset(rng.uniform(-3, 3, size=dim))])
    cases = [tr, tr.inverse()]
    if dim == 3:
        cases.append(tfm.Composition([tfm.Rotation(euler=rng.uniform(-180, 180, size=3), unit='deg'),
                                      tfm.Offset([1.0, -2.0, 0.5])]))
    for n, extra in ((0, 0), (1, 0), (1000, 0), (777, 2)):
        v = rng.uniform(-100.0, 100.0, size=(n, dim + extra))
        for c in cases:
            for inv in (False, True):
                got = c.apply(v, invert=inv)
                want = oracle.apply(to_oracle(c), v, invert=inv)
                assert got.shape == want.shape and got.dtype == np.float64
                if dim == 3:
                    assert np.array_equal(got, want), (n, extra, inv)
                else:
                    assert np.allclose(got, want, rtol=1e-13, atol=1e-12), (dim, n, inv, np.abs(got - want).max())
                if extra:
                    assert np.array_equal(got[:, dim:], v[:, dim:])
    vt = torch.from_numpy(rng.uniform(-10, 10, size=(500, dim))).cuda()
    out = tr.apply(vt)
    assert out.is_cuda and out.dtype == torch.float64
    assert np.array_equal(out.cpu().numpy(), tr.apply(vt.cpu().numpy()))


def test_linear_nd_matches_reference_golden(tfm):
    """K3-ND against golden vectors of the unmodified reference (tests/golden/linear_nd_golden.npz): bit for bit in
    3-D (Rotation from Euler angles / axis pairs, Scale o Rotation o Offset, a general matrix with offsets; forward
    and inverted; pass-through components), 1e-13 relative for the 4-D chain."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_linear_nd_golden as mk
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "linear_nd_golden.npz"))
    a = {k[4:]: G[k] for k in G.files if k.startswith("arg/")}
    for key in [k for k in G.files if not k.startswith("arg/")]:
        case, vn, d = key.split("/")
        got = mk.build(tfm, case, a).apply(a[vn].copy(), invert=(d == "inv"))
        if case == "chain4":
            assert np.allclose(got, G[key], rtol=1e-13, atol=1e-12), key
        else:
            assert np.array_equal(got, G[key]), key
