"""GPU: the CUDA path against golden vectors recorded from the REAL reference
(tests/golden/reference_golden.npz).  Same bars as the oracle's own pinning: bit-exact for affine
chains and for interpND on given indices, <= 1e-12 relative through transcendental ops."""
import numpy as np
import pytest

from conftest import to_oracle, assert_close_rel
from test_oracle_golden import GOLD, RESAMPLE_KEYS, INTERP_KEYS, FILT_KEYS, _chains

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def chains(tfm):
    return _chains(tfm)


@pytest.mark.parametrize("key", RESAMPLE_KEYS)
def test_resample_matches_reference(tfm, oracle, chains, key):
    _, name, m, b, sname = key.split("/")
    src = GOLD["src32"] if sname == "f32" else GOLD["src64"]
    got = chains[name].resample(src, method=m, bound=b)
    want = GOLD[key]
    assert got.dtype == want.dtype and got.shape == want.shape
    if "radial" in name:
        if m == 'n':
            coords = to_oracle(chains[name]).reverse(oracle.pixel_grid(*src.shape))
            bad = np.argwhere(got != want)
            for (y, x) in bad:
                tt = coords[y, x] + 0.5
                assert np.abs(tt - np.round(tt)).min() < 1e-9
            assert len(bad) <= 4
        else:
            assert_close_rel(got, want, 1e-12, key)
    else:
        assert np.array_equal(got, want), key


def test_apply_matches_reference(tfm, chains):
    for name, tr in chains.items():
        fwd = tr.apply(GOLD["apply/%s/in" % name])
        rev = tr.invert(fwd)
        f3 = tr.apply(GOLD["apply/%s/in3" % name])
        if "radial" in name:
            assert_close_rel(fwd, GOLD["apply/%s/fwd" % name], 1e-13, name)
            assert_close_rel(rev, GOLD["apply/%s/rev" % name], 1e-12, name)
            assert_close_rel(f3, GOLD["apply/%s/fwd3" % name], 1e-13, name)
        else:
            assert np.array_equal(fwd, GOLD["apply/%s/fwd" % name]), name
            assert np.array_equal(rev, GOLD["apply/%s/rev" % name]), name
            assert np.array_equal(f3, GOLD["apply/%s/fwd3" % name]), name


@pytest.mark.parametrize("key", INTERP_KEYS)
def test_interpND_matches_reference(tfm, key):
    _, m, b, fv = key.split("/")
    bound = b.split("+") if "+" in b else b
    got = tfm.interpND(GOLD["src32"], GOLD["interp/index"], method=m, bound=bound, fillvalue=int(fv))
    want = GOLD[key]
    assert got.dtype == want.dtype
    same = (got == want) | (np.isnan(got) & np.isnan(want))
    assert same.all(), (key, np.argwhere(~same)[:4])


@pytest.mark.parametrize("key", FILT_KEYS)
def test_interpND_filters_match_reference(tfm, key):
    """Collapsible filters (helpers.pyx:703-735, 792-835) against the real reference: the tent has no
    transcendental and is bit-exact; the others go through sin / cos / exp, whose last-ulp
    differences between CUDA's and the host's libm leave <= 1e-12."""
    _, m, b, ob = key.split("/")
    bound = b.split("+") if "+" in b else b
    ob = None if ob == "None" else float(ob)
    got = tfm.interpND(GOLD["src32"], GOLD["interp/index"], method=m, bound=bound, fillvalue=-9, oblur=ob)
    want = GOLD[key]
    assert got.dtype == want.dtype and got.shape == want.shape
    if m == 'l':
        assert ((got == want) | (np.isnan(got) & np.isnan(want))).all(), key
    else:
        assert_close_rel(got, want, 1e-12, key)


def test_filters_through_resample_match_reference(tfm, chains):
    for m in "szghtl":
        got = tfm.interpND(GOLD["src64"], GOLD["interp/index"], method=m, bound='m', oblur=1.3)
        assert_close_rel(got, GOLD["filt64/%s" % m], 1e-12, m)
    for m, name in (("s", "sinc"), ("z", "zlanczos"), ("g", "gaussian"), ("h", "hann"), ("t", "tukey")):
        got = chains["affine"].resample(GOLD["src32"], method=name, bound='t')
        want = GOLD["fresample/affine/%s/t/f32" % m]
        assert got.dtype == want.dtype
        assert_close_rel(got, want, 1e-12, m)
        got = chains["radial"].resample(GOLD["src64"], method=name, bound='p', shape=(20, 37))
        assert_close_rel(got, GOLD["fresample/radial/%s/p/f64" % m], 1e-11, m)


def test_svd_matches_reference(tfm):
    for k in range(GOLD["svd/M"].shape[0]):
        U, s, V = np.zeros((2, 2)), np.zeros(2), np.zeros((2, 2))
        tfm.svd2x2(GOLD["svd/M"][k], U, s, V)
        assert np.allclose(U, GOLD["svd/U"][k], rtol=0, atol=1e-14)
        assert np.allclose(s, GOLD["svd/s"][k], rtol=0, atol=1e-14)
        assert np.allclose(V, GOLD["svd/V"][k], rtol=0, atol=1e-14)


# ---------------------------------------------------------------------------------------------
# per-axis pincushion family (basic.py:846-1462)
# ---------------------------------------------------------------------------------------------
from test_oracle_golden import PIN_NAMES, _pin_chains  # noqa: E402


@pytest.mark.parametrize("name", PIN_NAMES)
def test_pincushion_matches_reference(tfm, name):
    """Quadpin is add/mul/div/sqrt only -> bit-exact; Cubicpin and Poly go through pow(), whose last
    bit differs between CUDA and glibc -> 1e-12 (resampled values: bilinear weights inherit it)."""
    tr = _pin_chains(tfm)[name]
    exact = name.startswith("quad")
    v = GOLD["pin/in"]

    def check(got, want, what):
        assert got.shape == want.shape and got.dtype == want.dtype, (what, got.dtype, want.dtype)
        if exact:
            assert np.array_equal(got, want, equal_nan=True), what
        else:
            assert_close_rel(got, want, 1e-12, what)

    if "pin/%s/fwd" % name in GOLD.files:
        check(tr.apply(v), GOLD["pin/%s/fwd" % name], name + " fwd")
    if "pin/%s/rev" % name in GOLD.files:
        check(tr.invert(v), GOLD["pin/%s/rev" % name], name + " rev")
    if "pin/%s/lin" % name in GOLD.files:
        got = tr.resample(GOLD["src32"], method='linear')
        want = GOLD["pin/%s/lin" % name]
        if exact:
            assert np.array_equal(got, want, equal_nan=True), name
        else:
            # a 1-ulp coordinate difference moves a bilinear value by ~1e-13; at an exact integer
            # coordinate it can flip floor(), which leaves the value continuous all the same
            assert_close_rel(got, want, 1e-11, name + " lin")
        got = tr.resample(GOLD["src64"], method='nearest', bound='m')
        want = GOLD["pin/%s/near" % name]
        assert got.dtype == want.dtype
        assert (got != want).sum() <= (0 if exact else 3), name


# ---------------------------------------------------------------------------------------------
# the anti-aliased path against the reference's own code built with the named repairs
# ---------------------------------------------------------------------------------------------
from test_oracle_golden import AA_KEYS, _aa_kwargs  # noqa: E402


@pytest.mark.parametrize("key", AA_KEYS)
def test_antialiased_path_matches_repaired_reference(tfm, key):
    """K2 (through interpND_grid on the recorded coordinate arrays) against outputs of the reference's
    own interpND_grid -> jacobian -> interpND_jacobian compiled with repairs R1-R6, R11
    (oracle/build_ref_repaired.py, fixtures by tests/golden/make_golden.py).  Parity mode: same
    summation order, libm last bits only; fp32 mode: tolerance, isolated footprint-size flips."""
    name, kw = _aa_kwargs(key)
    coords = GOLD["aa/%s/coords" % name]
    want = GOLD[key]
    got = tfm.interpND_grid(GOLD["src64"], coords, sv_limit=100.0, **kw)
    assert got.dtype == want.dtype and got.shape == want.shape
    assert_close_rel(got, want, 1e-11, key)
    if kw["method"] == "g":
        fast = tfm.interpND_grid(GOLD["src64"].astype(np.float32), coords, sv_limit=100.0, precision='fp32', **kw)
        err = np.abs(fast - want)
        assert np.median(err) < 2e-5 and (err > 5e-4).sum() <= max(3, err.size // 100), (key, err.max())
