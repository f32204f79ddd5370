"""CPU: the oracle against the LIVE reference (imported from /root/reference through
oracle/ref_import.py) on seeded random cases -- beyond the committed golden vectors.  Runs only where
the reference tree is mounted (the authoring container); skipped on the GPU box."""
import os
import sys
import warnings

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/transform"),
                                reason="the reference tree is not mounted on this machine")

BOUNDS = ['t', 'e', 'p', 'm']


@pytest.fixture(scope="module")
def ref():
    import ref_import
    warnings.simplefilter("ignore")
    return ref_import.load_reference()


def _pair(ref, o, rng, ny, nx, kind):
    """The same random transform as a reference object and as an oracle object."""
    if kind == "affine":
        sc = rng.uniform(0.5, 2.0, size=2)
        ang = rng.uniform(-180, 180)
        off = rng.uniform(-0.6, 0.1, size=2) * np.array([nx, ny])
        r = ref.Composition([ref.Scale(sc), ref.Rotation(ang, unit='deg'), ref.Offset(off)])
        q = o.Composition([o.Scale(sc), o.Rotation(ang, unit='deg'), o.Offset(off)])
        return r, q
    if kind == "radial":
        org = [rng.uniform(0.2, 0.8) * nx, rng.uniform(0.2, 0.8) * ny]
        ccw, pos = bool(rng.random() < 0.5), bool(rng.random() < 0.5)
        r0 = None if rng.random() < 0.6 else rng.uniform(0.5, 3.0)
        pre = [nx / (2 * np.pi), ny / (3.0 if r0 else 0.7 * max(nx, ny))]
        r = ref.Composition([ref.Scale(pre), ref.Radial(origin=np.array(org), ccw=ccw, pos_only=pos, r0=r0)])
        q = o.Composition([o.Scale(pre), o.Radial(origin=org, ccw=ccw, pos_only=pos, r0=r0)])
        return r, q
    st, ln = rng.uniform(-0.4, 0.4), rng.uniform(20, 80)
    org = rng.uniform(0.3, 0.7, size=2) * np.array([nx, ny])
    if rng.random() < 0.5:
        return (ref.Quadpin(idim=2, odim=2, strength=st, length=ln, origin=org),
                o.Quadpin(idim=2, strength=st, length=ln, origin=org))
    deg = int(rng.integers(2, 5))
    return (ref.Poly(idim=2, odim=2, strength=st / 2, length=ln, degree=deg, origin=org).inverse(),
            o.Inverse(o.Poly(idim=2, strength=st / 2, length=ln, degree=deg, origin=org)))


@pytest.mark.parametrize("seed", range(30))
def test_resample_random_cases(ref, oracle, seed):
    rng = np.random.default_rng(7000 + seed)
    ny, nx = int(rng.integers(3, 60)), int(rng.integers(3, 80))
    kind = ("affine", "radial", "pin")[seed % 3]
    r, q = _pair(ref, oracle, rng, ny, nx, kind)
    dt = (np.float32, np.float64, np.int16)[int(rng.integers(0, 3))]
    src = (rng.random((ny, nx)) * 100).astype(dt)
    shape = None if rng.random() < 0.5 else (int(rng.integers(1, 70)), int(rng.integers(1, 90)))
    bound = BOUNDS[int(rng.integers(0, 4))]
    if rng.random() < 0.3:
        bound = [BOUNDS[int(rng.integers(0, 4))], BOUNDS[int(rng.integers(0, 4))]]
    for m in ('nearest', 'linear', 'cubic', 'hann', 'gaussian', 'zlanczos'):
        if m == 'cubic' and np.dtype(dt).kind in 'iu':
            continue
        want = r.resample(src, method=m, bound=bound, shape=shape)
        got = oracle.resample(q, src, method=m, bound=bound, shape=shape)
        what = (seed, kind, m, bound, shape, np.dtype(dt).name)
        assert got.shape == want.shape and got.dtype == want.dtype, what
        if kind == "radial":
            if m == 'nearest':
                assert (got != want).sum() <= max(2, got.size // 500), what
            else:
                ok = np.isclose(got, want, rtol=0, atol=1e-10 * max(1.0, float(np.nanmax(np.abs(want)))), equal_nan=True)
                assert ok.all(), what
        else:
            assert np.array_equal(got, want, equal_nan=True), what


@pytest.mark.parametrize("seed", range(12))
def test_apply_random_cases(ref, oracle, seed):
    rng = np.random.default_rng(8000 + seed)
    kind = ("affine", "radial", "pin")[seed % 3]
    r, q = _pair(ref, oracle, rng, 64, 80, kind)
    v = rng.uniform(1.0, 60.0, size=(int(rng.integers(1, 400)), int(rng.integers(2, 5))))
    for invert in (False, True):
        if (invert and r.no_reverse) or (not invert and r.no_forward):
            continue
        want = r.apply(v.copy(), invert=invert)
        got = oracle.apply(q, v.copy(), invert=invert)
        if kind == "radial":
            assert np.allclose(got, want, rtol=1e-13, atol=1e-12), (seed, invert)
        else:
            assert np.array_equal(got, want, equal_nan=True), (seed, kind, invert)


@pytest.mark.parametrize("seed", range(8))
def test_jacobian_pieces_random_cases(ref, oracle, seed):
    """simple_jacobian / jump_detect (square grids: no R12 ambiguity) / svd2x2 against the live
    reference's compiled helpers."""
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.integers(5, 40))
    # a smooth map with a branch cut now and then: forward polar coordinates of a shifted grid
    g = np.mgrid[range(n), range(n)].transpose().astype(float)
    org = rng.uniform(0.2, 0.8, size=2) * n
    d = g - org
    coords = np.stack([np.arctan2(d[..., 1], d[..., 0]) * rng.uniform(1, 20), np.hypot(d[..., 0], d[..., 1])], -1)
    coords += rng.normal(0, 1e-3, coords.shape)
    Js_ref = ref.simple_jacobian(coords)
    Js = oracle.simple_jacobian(coords)
    assert np.array_equal(Js, Js_ref)
    for thr in (1.2, 4.0, 10.0):
        assert np.array_equal(oracle.jump_detect(Js, thr), ref.jump_detect(Js_ref, thr)), (seed, thr)
    for _ in range(10):
        M = rng.uniform(-3, 3, size=(2, 2))
        U, s, V = np.zeros((2, 2)), np.zeros(2), np.zeros((2, 2))
        ref.svd2x2(M, U, s, V)
        U2, s2, V2 = oracle.svd2x2(M)
        assert np.allclose(U2, U, atol=1e-14) and np.allclose(s2, s, atol=1e-14) and np.allclose(V2, V, atol=1e-14)


@pytest.mark.parametrize("seed", range(6))
def test_antialiased_path_random_cases(oracle, seed):
    """oracle.interp_grid_jacobian against the repaired build of the reference's own interpND_grid
    (oracle/build_ref_repaired.py) on random smooth maps; square grids, no mirror boundary (R9 and R12
    are the oracle's two deliberate departures)."""
    import ref_import
    hr = ref_import.load_ref_helpers_repaired()
    rng = np.random.default_rng(9500 + seed)
    n = int(rng.integers(12, 40))
    src = rng.random((int(rng.integers(10, 50)), int(rng.integers(10, 50))))
    g = np.mgrid[range(n), range(n)].transpose().astype(float)
    A = rng.uniform(-1.5, 1.5, size=(2, 2)) + np.eye(2)
    idx = (g - n / 2) @ A.T + np.array(src.shape[::-1]) / 2 + 0.05 * np.sin(g[..., ::-1] / 3.0)
    for m in "glhtz":
        for b in ("t", "e", "p"):
            jd = (4.0, 0, 1.5)[int(rng.integers(0, 3))]
            ob = (1.0, 1.7)[int(rng.integers(0, 2))]
            want = np.asarray(hr.interpND_grid(src, idx.copy(), method=m, bound=b, jump_detect=jd, oblur=ob, strict=True))
            got = oracle.interp_grid_jacobian(src, idx.copy(), method=m, bound=b, jump_detect=jd, oblur=ob)
            assert np.array_equal(got, want, equal_nan=True), (seed, m, b, jd, ob)


def test_mirror_boundary_departure_R9(oracle):
    """R9: under the mirror boundary the reference's footprint loop reflects with `s-1-i` after the
    mod-2s step (helpers.pyx:1768-1769, 1789-1790) where its own apply_boundary uses `2s-1-i`
    (helpers.pyx:172).  The oracle (and K2) follow apply_boundary.  The two agree wherever no tap
    leaves the image and differ where one does -- and there the reference's rule can index out of the
    array's lower end (negative index), which is why it is not reproduced."""
    import ref_import
    hr = ref_import.load_ref_helpers_repaired()
    rng = np.random.default_rng(77)
    src = rng.random((40, 40))
    n = 16
    g = np.mgrid[range(n), range(n)].transpose().astype(float)
    inner = g * 0.9 + 12.3                                   # footprints stay inside the image
    a = np.asarray(hr.interpND_grid(src, inner.copy(), method='g', bound='m', strict=True))
    b = oracle.interp_grid_jacobian(src, inner.copy(), method='g', bound='m')
    assert np.array_equal(a, b)
    edge = g * 0.9 + 30.2                                    # footprints cross the upper edges
    a = np.asarray(hr.interpND_grid(src, edge.copy(), method='g', bound='m', strict=True))
    b = oracle.interp_grid_jacobian(src, edge.copy(), method='g', bound='m')
    crossing = (edge.max(axis=-1) + 3 >= 40)
    assert np.array_equal(a[~crossing], b[~crossing])
    assert not np.array_equal(a[crossing], b[crossing])
    # the oracle's rule is apply_boundary's: same result as mirroring the image explicitly
    big = np.pad(src, 40, mode='symmetric')
    c = oracle.interp_grid_jacobian(big, edge + 40.0, method='g', bound='t')
    assert np.allclose(b, c, rtol=0, atol=1e-12)          # shifted coordinates round their fractions differently


@pytest.mark.parametrize("seed", range(6))
def test_apply_linear_3d_random_cases(ref, oracle, seed):
    """Linear-family transforms beyond two dimensions (basic.py:141-176, 338-506: Rotation from axis pairs or
    Euler angles, Scale, Offset, a general matrix with pre/post offsets) through Transform.apply, forward and
    inverted, with pass-through components: the oracle equals the live reference bit for bit in 3-D."""
    import transform_b200 as t                                   # host-side mirror: params only (no GPU here)
    from conftest import to_oracle
    rng = np.random.default_rng(9100 + seed)
    eul = rng.uniform(-180, 180, size=3)
    sc = rng.uniform(0.5, 2.0, size=3)
    off = rng.uniform(-5, 5, size=3)
    m = rng.standard_normal((3, 3))
    pre, post = rng.uniform(-1, 1, size=3), rng.uniform(-1, 1, size=3)
    cases = {
        "euler": lambda p: p.Rotation(euler=eul, unit='deg'),
        "pairs": lambda p: p.Rotation([[1, 2, eul[0]], [0, 1, eul[1]]], unit='deg'),
        "chain": lambda p: p.Composition([p.Scale(sc), p.Rotation(euler=eul, unit='deg'), p.Offset(off)]),
        "general": lambda p: p.Linear(matrix=m, pre=pre, post=post),
    }
    v = rng.uniform(-50.0, 50.0, size=(int(rng.integers(1, 300)), int(rng.integers(3, 6))))
    for name, make in cases.items():
        r, mine = make(ref), make(t)
        q = to_oracle(mine)
        for invert in (False, True):
            want = r.apply(v.copy(), invert=invert)
            got = oracle.apply(q, v.copy(), invert=invert)
            assert got.shape == want.shape
            assert np.array_equal(got, want), (seed, name, invert, np.abs(got - want).max())


def test_compiled_reference_classes_equal_the_source_import(tmp_path):
    """oracle/build_ref.py also compiles the reference's core.py / basic.py (where they lie) into oracle/_ref, so
    that bench.py's reference arm can drive the reference's OWN Composition.invert on the GPU box, where
    /root/reference does not exist.  The compiled classes must be the reference: same results, bit for bit, as
    the source import (run in a subprocess: both register a package named `transform`)."""
    import subprocess
    import ref_import
    code = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
import ref_import
ref = ref_import.%s()
n = 96
src = np.random.default_rng(0).random((n, n)).astype(np.float32)
v = np.random.default_rng(1).random((200, 3)) * 60
aff = ref.Composition([ref.Scale(1.3, dim=2), ref.Rotation(30, unit="deg"), ref.Offset([-n / 2.0, -n / 2.0])])
rad = ref.Composition([ref.Scale([n / (2 * np.pi), 1.4]), ref.Radial(origin=[47.5, 47.5])])
out = {"lin": aff.resample(src, method="linear"), "cub": aff.resample(src, method="cubic"),
       "near": rad.resample(src, method="nearest"), "inv": rad.invert(v), "fwd": aff.apply(v),
       "pin": ref.Cubicpin(idim=2, odim=2, strength=0.1, length=30.0).apply(v[:, :2]),
       "rot3": ref.Rotation(euler=np.array([10.0, 20.0, 30.0]), unit="deg").apply(v)}
np.savez(%r, **out)
print(str(aff))
'''
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle")
    res = {}
    for which in ("load_reference_compiled", "load_reference"):
        path = str(tmp_path / (which + ".npz"))
        r = subprocess.run([sys.executable, "-c", code % (here, which, path)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        res[which] = (np.load(path), r.stdout.strip())
    a, b = res["load_reference_compiled"], res["load_reference"]
    assert a[1] == b[1]                                           # the stringifier protocol too
    for k in b[0].files:
        assert np.array_equal(a[0][k], b[0][k], equal_nan=True), k


def test_cpu_arm_linear_leg_is_the_reference():
    """The linear leg of the CPU baseline runs the reference's own classes when oracle/_ref holds them."""
    import cpu_baseline as cb
    arm = cb.CpuArm(512, 2, seed=0)
    try:
        assert arm.kind_lin == "reference", arm.kind_lin
        t, px = arm.linear(64)
        assert px == 64 * 512 and t > 0
    finally:
        arm.close()
