"""CPU: the oracle (oracle/oracle_np.py + oracle_c.c) against golden vectors recorded FROM THE REAL
REFERENCE (tests/golden/make_golden.py, run in the authoring container) -- this is what pins the
oracle.  Bars: bit-exact wherever the arithmetic convention is fully defined (interpND on given
indices; affine chains); <= 1e-12 relative for chains through libm transcendentals (Radial), whose
last bit may differ between machines."""
import importlib.util
import os
import sys

import numpy as np
import pytest

from conftest import to_oracle, assert_close_rel

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = np.load(os.path.join(HERE, "golden", "reference_golden.npz"))


def _chains(module):
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.chains(module)


@pytest.fixture(scope="module")
def chains(tfm):
    return _chains(tfm)


RESAMPLE_KEYS = [k for k in GOLD.files if k.startswith("resample/") and not k.endswith("/shape")]


def _near_tie_only(got, want, coords):
    bad = np.argwhere(got != want)
    for (y, x) in bad:
        t = coords[y, x] + 0.5
        assert np.abs(t - np.round(t)).min() < 1e-9, (y, x, coords[y, x])
    assert len(bad) <= 4


@pytest.mark.parametrize("key", RESAMPLE_KEYS)
def test_resample_matches_reference(oracle, chains, key):
    _, name, m, b, sname = key.split("/")
    src = GOLD["src32"] if sname == "f32" else GOLD["src64"]
    otr = to_oracle(chains[name])
    got = oracle.resample(otr, src, method=m, bound=b)
    want = GOLD[key]
    assert got.dtype == want.dtype and got.shape == want.shape
    if "radial" in name:
        if m == 'n':
            _near_tie_only(got, want, otr.reverse(oracle.pixel_grid(*src.shape)))
        else:
            assert_close_rel(got, want, 1e-12, key)
    else:
        assert np.array_equal(got, want), key


def test_resample_shape_argument(oracle, chains):
    for name, tr in chains.items():
        got = oracle.resample(to_oracle(tr), GOLD["src64"], method='linear', shape=(20, 37))
        want = GOLD["resample/%s/shape" % name]
        assert got.shape == (20, 37)
        if "radial" in name:
            assert_close_rel(got, want, 1e-12, name)
        else:
            assert np.array_equal(got, want), name


def test_apply_matches_reference(oracle, chains):
    for name, tr in chains.items():
        otr = to_oracle(tr)
        v = GOLD["apply/%s/in" % name]
        fwd = oracle.apply(otr, v)
        rev = oracle.apply(otr, fwd, invert=True)
        f3 = oracle.apply(otr, GOLD["apply/%s/in3" % name])
        if "radial" in name:
            assert_close_rel(fwd, GOLD["apply/%s/fwd" % name], 1e-13, name)
            assert_close_rel(rev, GOLD["apply/%s/rev" % name], 1e-12, name)
            assert_close_rel(f3, GOLD["apply/%s/fwd3" % name], 1e-13, name)
        else:
            assert np.array_equal(fwd, GOLD["apply/%s/fwd" % name]), name
            assert np.array_equal(rev, GOLD["apply/%s/rev" % name]), name
            assert np.array_equal(f3, GOLD["apply/%s/fwd3" % name]), name


INTERP_KEYS = [k for k in GOLD.files if k.startswith("interp/") and k != "interp/index"]


@pytest.mark.parametrize("key", INTERP_KEYS)
def test_interpND_matches_reference(oracle, key):
    _, m, b, fv = key.split("/")
    bound = b.split("+") if "+" in b else b
    got = oracle.interpND(GOLD["src32"], GOLD["interp/index"], method=m, bound=bound, fillvalue=int(fv))
    want = GOLD[key]
    assert got.dtype == want.dtype
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (key, np.argwhere(~same)[:4])


FILT_KEYS = [k for k in GOLD.files if k.startswith("filt/")]


@pytest.mark.parametrize("key", FILT_KEYS)
def test_interpND_filters_match_reference(oracle, key):
    """sinc / Lanczos / Gaussian / Hann / Tukey / tent-with-oblur (helpers.pyx:703-735, 792-835):
    bit-exact, including NumPy's reduction orders."""
    _, m, b, ob = key.split("/")
    bound = b.split("+") if "+" in b else b
    ob = None if ob == "None" else float(ob)
    got = oracle.interpND(GOLD["src32"], GOLD["interp/index"], method=m, bound=bound, fillvalue=-9, oblur=ob)
    want = GOLD[key]
    assert got.dtype == want.dtype
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (key, np.argwhere(~same)[:4])


def test_filters_through_resample_match_reference(oracle, chains):
    for m in "szghtl":
        got = oracle.interpND(GOLD["src64"], GOLD["interp/index"], method=m, bound='m', oblur=1.3)
        want = GOLD["filt64/%s" % m]
        assert ((got == want) | (np.isnan(got) & np.isnan(want))).all(), m
    for m in "szght":
        got = oracle.resample(to_oracle(chains["affine"]), GOLD["src32"], method=m, bound='t')
        assert np.array_equal(got, GOLD["fresample/affine/%s/t/f32" % m], equal_nan=True), m
        got = oracle.resample(to_oracle(chains["radial"]), GOLD["src64"], method=m, bound='p', shape=(20, 37))
        assert_close_rel(got, GOLD["fresample/radial/%s/p/f64" % m], 1e-11, m)
    with pytest.raises(ValueError):
        oracle.interpND(GOLD["src32"], GOLD["interp/index"], method='n', oblur=2.0)
    with pytest.raises(ValueError):
        oracle.interpND(GOLD["src32"], GOLD["interp/index"], method='c', oblur=2.0)


def _pin_chains(module):
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.pin_chains(module)


PIN_NAMES = sorted({k.split("/")[1] for k in GOLD.files if k.startswith("pin/") and k != "pin/in"})


@pytest.mark.parametrize("name", PIN_NAMES)
def test_pincushion_matches_reference(oracle, tfm, name):
    """Quadpin / Cubicpin / Poly / Quarticpin (basic.py:846-1462), quirks included: bit-exact."""
    tr = _pin_chains(tfm)[name]
    otr = to_oracle(tr)
    v = GOLD["pin/in"]
    if "pin/%s/fwd" % name in GOLD.files:
        assert np.array_equal(oracle.apply(otr, v), GOLD["pin/%s/fwd" % name], equal_nan=True)
    if "pin/%s/rev" % name in GOLD.files:
        assert np.array_equal(oracle.apply(otr, v, invert=True), GOLD["pin/%s/rev" % name], equal_nan=True)
    if "pin/%s/lin" % name in GOLD.files:
        assert np.array_equal(oracle.resample(otr, GOLD["src32"], method='l'), GOLD["pin/%s/lin" % name], equal_nan=True)
        assert np.array_equal(oracle.resample(otr, GOLD["src64"], method='n', bound='m'), GOLD["pin/%s/near" % name],
                              equal_nan=True)


def test_jacobian_pieces_match_reference(oracle):
    # square grid: simple_jacobian and jump_detect bit-identical to the reference
    sq = GOLD["jac/coords_sq"]
    Js = oracle.simple_jacobian(sq)
    assert np.array_equal(oracle.jump_detect(Js, 4.0), GOLD["jac/jump4_sq"])
    assert np.array_equal(oracle.jump_detect(Js, 1.2), GOLD["jac/jump1p2_sq"])
    assert (GOLD["jac/jump4_sq"] == 0).any() and (GOLD["jac/jump1p2_sq"] == 0).any()
    # non-square grid: simple_jacobian identical; jump_detect differs ONLY through repair R12 --
    # the reference writes its upper-edge results to index shape[-1]-1 of the moved axis
    # (helpers.pyx:1136) instead of shape[0]-1, i.e. onto an interior row/column (or nowhere),
    # and leaves the true upper edge at 1.
    coords = GOLD["jac/coords"]
    Js = oracle.simple_jacobian(coords)
    assert np.array_equal(Js, GOLD["jac/simple"])
    ch, cw = Js.shape[0], Js.shape[1]
    for thr, key in ((4.0, "jac/jump4"), (1.2, "jac/jump1p2")):
        mine, ref = oracle.jump_detect(Js, thr), GOLD[key]
        diff = np.argwhere(mine != ref)
        for (r, c) in diff:
            assert r in (ch - 1, cw - 1) or c in (cw - 1, ch - 1), (thr, r, c)
        assert np.all(ref[ch - 1, 1:-1] == 1) or ch == cw       # the reference never touched its top edge


def test_svd_matches_reference(oracle):
    for k in range(GOLD["svd/M"].shape[0]):
        U, s, V = oracle.svd2x2(GOLD["svd/M"][k])
        assert np.allclose(U, GOLD["svd/U"][k], rtol=0, atol=1e-15)
        assert np.allclose(s, GOLD["svd/s"][k], rtol=0, atol=1e-15)
        assert np.allclose(V, GOLD["svd/V"][k], rtol=0, atol=1e-15)
        assert np.allclose(oracle.svc2x2(U, s, V), GOLD["svd/M"][k], atol=1e-14)
        assert s[0] >= s[1] >= 0


SVDPYX = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "svdpyx_golden.npz"))
SVDPYX_CASES = sorted(k.split("/")[0] for k in SVDPYX.files if k.endswith("/params"))


@pytest.mark.parametrize("case", SVDPYX_CASES)
@pytest.mark.parametrize("phot", ["radiance", "flux"])
def test_hann_window_and_flux_match_reference_svdpyx(oracle, case, phot):
    """The Hann-window filter ('w') and phot='flux' against the reference's WORKING DeForest-2004
    implementation, SVD.pyx::map_coordinates (SVD.pyx:193-292; window 110-111, flux rule 285-286), on
    the maps where the two algorithms coincide tap for tap (tests/golden/make_svdpyx_golden.py:
    integer source positions, diagonal Jacobian with scales >= 2, 'extend' boundary, pad_pow large so
    that interpND_grid's padded inverse singular values equal map_coordinates' plain 1/s).
    Tolerance 1e-14 absolute on values of order 1 (measured 3e-16 / 2e-15)."""
    sx, sy, ox, oy = SVDPYX[case + "/params"]
    want = SVDPYX[case + "/" + phot]
    otr = oracle.Linear(matrix=np.diag([1 / sx, 1 / sy]), matinv=np.diag([sx, sy]), pre=[-ox, -oy])
    got = oracle.resample_jacobian(otr, SVDPYX["src64"], method='w', bound='e', shape=want.shape,
                                   pad_pow=64, jump_detect=0, phot=phot)
    assert float(np.abs(got - want).max()) <= 1e-14 * max(1.0, float(np.abs(want).max()))
    # the window as coded in helpers.pyx:1708-1717 ('h', cos^2) is a different function: it must NOT agree
    other = oracle.resample_jacobian(otr, SVDPYX["src64"], method='h', bound='e', shape=want.shape,
                                     pad_pow=64, jump_detect=0, phot=phot)
    assert float(np.abs(other - want).max()) > 1e-3


AA_KEYS = [k for k in GOLD.files if k.startswith("aa/") and not k.endswith("/coords")]


def _aa_kwargs(key):
    parts = key.split("/")
    kw = {"method": parts[2], "bound": parts[3]}
    if len(parts) > 4 and parts[4] == "nojump":
        kw["jump_detect"] = 0
    if len(parts) > 4 and parts[4] == "oblur2":
        kw["oblur"] = 2.0
    return parts[1], kw


@pytest.mark.parametrize("key", AA_KEYS)
def test_antialiased_path_matches_repaired_reference(oracle, key):
    """The anti-aliased resampler against the reference's OWN code built with the named repairs
    R1-R6, R11 (oracle/build_ref_repaired.py): identical bits, for every filter shape, boundary,
    with and without jump detection."""
    name, kw = _aa_kwargs(key)
    got = oracle.interp_grid_jacobian(GOLD["src64"], GOLD["aa/%s/coords" % name], **kw)
    assert np.array_equal(got, GOLD[key], equal_nan=True), key


def _linear_nd_cases():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_linear_nd_golden as mk
    G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "linear_nd_golden.npz"))
    a = {k[4:]: G[k] for k in G.files if k.startswith("arg/")}
    return mk, G, a


def test_linear_nd_oracle_matches_reference_golden(oracle):
    """Linear-family transforms in 3-D / 4-D (basic.py:141-176, 338-506) through Transform.apply: the oracle against
    golden vectors of the unmodified reference (tests/golden/make_linear_nd_golden.py) -- bit for bit in 3-D, where the
    row order of np.matmul is pinned; 4 ulp of sum |m_ik v_k| for the 4-D chain."""
    import transform_b200 as t
    from conftest import to_oracle
    mk, G, a = _linear_nd_cases()
    for key in [k for k in G.files if not k.startswith("arg/")]:
        case, vn, d = key.split("/")
        q = to_oracle(mk.build(t, case, a))
        got = oracle.apply(q, a[vn].copy(), invert=(d == "inv"))
        if case == "chain4":
            assert np.allclose(got, G[key], rtol=1e-13, atol=1e-12), key
        else:
            assert np.array_equal(got, G[key]), key
