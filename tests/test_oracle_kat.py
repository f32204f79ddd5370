"""CPU: the reference's own known-answer tests (tests/test_0[123]_*.py of the reference, cited per
test) restated against the oracle -- the second leg of the oracle's pinning."""
import numpy as np

import oracle_np as o


def test_boundary_tie_rounding():
    """tests/test_03_helpers.py:22-54: floor(v+0.5) then per-mode boundary on a 3-wide axis."""
    v = np.array([0, 2, 3, 4, -1, -2, -0.49, -0.51, -0.1, -0.9])
    idx = np.floor(v + 0.5).astype(np.int64)
    assert list(o.boundary_index(idx, 3, 't')) == [0, 2, -1, -1, -1, -1, 0, -1, 0, -1]
    assert list(o.boundary_index(idx, 3, 'e')) == [0, 2, 2, 2, 0, 0, 0, 0, 0, 0]
    assert list(o.boundary_index(idx, 3, 'p')) == [0, 2, 0, 1, 2, 1, 0, 2, 0, 2]
    assert list(o.boundary_index(idx, 3, 'm')) == [0, 2, 2, 1, 0, 1, 0, 0, 0, 0]
    try:
        o.boundary_index(idx, 3, 'f')
        assert False
    except ValueError:
        pass
    assert list(o.boundary_index(np.array([2, 2]), 3, 'f')) == [2, 2]
    # SURVEY.md section 8c probe fact: ties round up
    assert list(np.floor(np.array([-1.5, -0.5, 0.5, 1.5, 2.5]) + 0.5).astype(int)) == [-1, 0, 1, 2, 3]


def test_nearest_and_chunks():
    """tests/test_03_helpers.py:84-94, 146-159, 221-228."""
    data = (np.mgrid[0:5, 0:50:10].transpose()).sum(axis=-1)
    assert o.interpND(data, np.array([3, 4]), 'n', bound='f') == 43
    assert list(o.interpND(data, np.array([[3, 4], [1, 2], [0.4, 1.2]]), 'n', bound='f')) == [43, 21, 10]
    assert o.interpND(data, np.array([1.2, 2.8]), 'n') == 31
    assert o.interpND(data, np.array([[-1, -1]]), 'n', bound=['p', 'e'])[0] == 4
    d2 = (np.mgrid[0:2, 0:20:10].transpose()).sum(axis=-1)
    gx, gy = np.meshgrid(np.arange(-1, 3), np.arange(-1, 3))
    idx = np.stack([gx, gy], -1).astype(float)
    assert np.array_equal(o.interpND(d2, idx, 'n', bound='p'),
                          [[11, 10, 11, 10], [1, 0, 1, 0], [11, 10, 11, 10], [1, 0, 1, 0]])
    assert np.array_equal(o.interpND(d2, idx, 'n', bound='t'),
                          [[0, 0, 0, 0], [0, 0, 1, 0], [0, 10, 11, 0], [0, 0, 0, 0]])
    assert np.array_equal(o.interpND(d2, idx, 'n', bound='m'),
                          [[0, 0, 1, 1], [0, 0, 1, 1], [10, 10, 11, 11], [10, 10, 11, 11]])


def test_linear_table():
    """tests/test_03_helpers.py:230-260."""
    data = np.mgrid[0:5, 0:50:10].transpose().sum(axis=-1)
    assert np.isclose(o.interpND(data, np.array([1, 3]), 'l'), 31, atol=1e-5)
    assert np.isclose(o.interpND(data, np.array([1.2, 3]), 'l'), 31.2, atol=1e-5)
    assert np.isclose(o.interpND(data, np.array([1, 2.8]), 'l'), 29, atol=1e-5)
    assert np.isclose(o.interpND(data, np.array([1.2, 2.8]), 'l'), 29.2, atol=1e-5)
    data = np.mgrid[0:5, 0:500:100].transpose().sum(axis=-1)
    dex = np.mgrid[1:2.1:0.2, 2:3.1:0.2].transpose()
    a = o.interpND(data, dex, 'l')
    want = np.array([[201 + 20 * j + 0.2 * i for i in range(6)] for j in range(6)])
    assert a.shape == (6, 6) and np.all(np.isclose(a, want, atol=1e-4))


def test_cubic_impulse_and_edges():
    """tests/test_03_helpers.py:262-297 (2-D image with one row standing in for the 1-D case) and
    the probe facts of SURVEY.md section 8c on a [1..5] ramp."""
    row = np.array([0, 0, 1, 0, 0], dtype=float)
    img = np.tile(row, (6, 1))
    x = np.arange(0, 4.01, 0.1)
    idx = np.stack([x, np.full_like(x, 2.0)], -1)
    y = o.interpND(img, idx, 'c', bound='e')
    assert np.all(np.isclose(y[[0, 10, 20, 30, 40]], [0, 0, 1, 0, 0], atol=1e-5))
    assert np.all(y[1:10] < 0) and np.all(y[11:30] > 0) and np.all(y[31:40] < 0)
    ramp = np.tile(np.arange(1.0, 6.0), (6, 1))
    xs = np.array([-0.5, -0.3, 4.3, 4.5, 4.7])
    idx = np.stack([xs, np.full_like(xs, 2.0)], -1)
    assert np.allclose(o.interpND(ramp, idx, 'l', bound='t'), [0.5, 0.7, 3.5, 2.5, 1.5])
    assert np.allclose(o.interpND(ramp, idx, 'c', bound='t'), [0.4375, 0.6685, 3.7835, 2.5625, 1.3215], atol=5e-5)
    assert np.allclose(o.interpND(ramp, idx, 'l', bound='e'), [1, 1, 5, 5, 5])
    xs = np.array([-1.0, -0.5, -0.3])
    idx = np.stack([xs, np.full_like(xs, 2.0)], -1)
    assert np.allclose(o.interpND(ramp, idx, 'l', bound='t', fillvalue=-9), [-9, -4, -2])
    nanidx = np.array([[np.nan, 2.0]])
    assert o.interpND(ramp, nanidx, 'n', bound='t', fillvalue=7)[0] == 7


def test_resample_kats():
    """tests/test_01_core.py:259-294."""
    a = np.zeros([7, 5])
    a[2, 2] = 1
    assert np.all(o.resample(o.Identity(), a, 'nearest') == a)
    assert np.all(o.resample(o.Identity(), a, 'nearest', shape=[5, 5]) == a[0:5, 0:5])
    b = o.resample(o.Scale(2.5, post=[2, 2], pre=[-2, -2]), a, 'neares')
    chk = np.zeros([7, 5])
    chk[1:4, 1:4] = 1
    assert np.all(b == chk)
    b = o.resample(o.Scale([1, 2.5], post=[2, 2], pre=[-2, -2]), a, 'nearest')
    chk = np.zeros([7, 5])
    chk[1:4, 2] = 1
    assert np.all(b == chk)


def test_rotation_and_radial_kats():
    """tests/test_02_basic.py:139-153, 209-283."""
    r = o.Rotation(90, 'deg')
    assert np.all(np.isclose(r.matrix, [[0, -1], [1, 0]], atol=1e-15))
    d0 = np.array([[1., 0], [0, 1]])
    d1 = r.forward(d0)
    assert np.all(np.isclose(d1, [[0, 1], [-1, 0]], atol=1e-15))
    assert np.all(np.isclose(r.reverse(d1), d0, atol=1e-15))
    x = np.arange(4.).reshape(2, 2)
    b = o.Radial(origin=[130, 130])
    assert np.all(np.isclose([[131.0, 130.0], [128.75, 127.27]], b.reverse(x), atol=1e-2))
    assert np.all(np.isclose([[2.3600555, 183.14202], [2.360116, 180.31362]], b.forward(x), atol=1e-2))
    cd = o.Radial(origin=[130, 130], r0=5)
    assert np.allclose([[2.3600555, 3.600824], [2.360116, 3.5852597]], cd.forward(x), atol=1e-2)
    assert np.allclose([[143.59141, 130], [88.207337, 38.681365]], cd.reverse(x), atol=1e-2)
    a = (np.mgrid[0:7, 0:7].T - 3).astype(float)
    tr = o.Radial()
    bb = tr.forward(a)
    assert np.isclose(bb[0, -1, 0], 0.25 * np.pi, atol=1e-5) and np.isclose(bb[0, 0, 0], 0.75 * np.pi, atol=1e-5)
    assert np.isclose(bb[-1, 0, 0], 1.25 * np.pi, atol=1e-5) and np.isclose(bb[-1, -1, 0], 1.75 * np.pi, atol=1e-5)
    assert np.all(np.isclose(bb[[0, -1, 0, -1], [0, -1, -1, 0], 1], 3 * np.sqrt(2), atol=1e-5))
    assert np.all(np.isclose(a, tr.reverse(bb), atol=1e-10))
    tr2 = o.Radial(pos_only=False)
    b2 = tr2.forward(a)
    assert np.isclose(b2[-1, 0, 0], -0.75 * np.pi, atol=1e-5) and np.isclose(b2[-1, -1, 0], -0.25 * np.pi, atol=1e-5)
    assert np.all(np.isclose(a, tr2.reverse(b2), atol=1e-10))


def test_simple_jacobian_and_jump_kats():
    """tests/test_03_helpers.py:380-402, 632-645."""
    grid = np.zeros([4, 4, 2])
    grid[1, 2, 0] = 1
    grid[2, 2, 1] = 0.5
    J = o.simple_jacobian(grid)
    assert J.shape == (3, 3, 2, 2)
    assert np.all(J[..., 0, 0] == [[0, 0.5, -0.5], [0, 0.5, -0.5], [0, 0, 0]])
    assert np.all(J[..., 0, 1] == [[0, 0.5, 0.5], [0, -0.5, -0.5], [0, 0, 0]])
    assert np.all(J[..., 1, 0] == [[0, 0, 0], [0, 0.25, -0.25], [0, 0.25, -0.25]])
    assert np.all(J[..., 1, 1] == [[0, 0, 0], [0, 0.25, 0.25], [0, -0.25, -0.25]])
    grid = np.zeros([5, 5, 2])
    grid[1, 2, 0] = 1
    grid[2, 2, 1] = 0.5
    jf = o.jump_detect(o.simple_jacobian(grid))
    assert np.all(jf == [[1, 1, 1, 1], [1, 1, 1, 1], [1, 0, 0, 1], [1, 1, 1, 1]])


def test_jacobian_interior_stencil():
    """without jumps the grid-centred Jacobian is the Sobel-like stencil
    (the CUDA fast path uses that form); edges copy their neighbours (helpers.pyx:1244-1247)."""
    rng = np.random.default_rng(0)
    I = rng.random((9, 11, 2)) * 3
    J = o.jacobian(I, jump_thresh=0)
    y, x = 4, 5
    ddx = (I[y + 1, x + 1] - I[y + 1, x - 1] + 2 * (I[y, x + 1] - I[y, x - 1]) + I[y - 1, x + 1] - I[y - 1, x - 1]) / 8
    ddy = (I[y + 1, x + 1] - I[y - 1, x + 1] + 2 * (I[y + 1, x] - I[y - 1, x]) + I[y + 1, x - 1] - I[y - 1, x - 1]) / 8
    assert np.allclose(J[y, x, :, 0], ddx, atol=1e-14) and np.allclose(J[y, x, :, 1], ddy, atol=1e-14)
    assert np.array_equal(J[0], J[1]) and np.array_equal(J[:, -1], J[:, -2]) and np.array_equal(J[0, 0], J[1, 1])


def test_svd_kats():
    """tests/test_03_helpers.py:667-711."""
    U, s, V = o.svd2x2(np.eye(2))
    assert np.allclose(U, np.eye(2)) and np.allclose(V, np.eye(2)) and np.allclose(s, 1)
    M = np.array(((0, 2.0), (0.333, 0)))
    U, s, V = o.svd2x2(M)
    assert np.isclose(s[0], 2) and np.isclose(s[1], 0.333)
    assert np.allclose(o.svc2x2(U, s, V), M, atol=1e-8)
    M = np.array(((1.0, 2.0), (4, 3)))
    U, s, V = o.svd2x2(M)
    assert np.allclose(o.svc2x2(U, s, V), M, atol=1e-8)
    assert np.allclose(np.matmul(M, o.svc2x2(V, 1.0 / s, U)), np.eye(2), atol=1e-8)


def test_wcs_linear_pin_and_identity_gaussian():
    """tests/test_01_core.py:160-169 (sample.fits CD-matrix WCS); SURVEY.md section 8c probe: identity map
    with the Gaussian gives s' = (3+1)^(-1/8) = 0.841, reg_size 4 -> 5x5 taps."""
    w = o.WCS([243, 256.5], [[0.831675615713575, 0.0232308034180093], [-0.0232308034180093, 0.831675615713575]],
              [-178.95728450798, -469.237931837451])
    assert np.all(np.isclose(w.forward(np.array([[0., 0.]])), [[-386.15825, -676.1092]], atol=1e-4))
    img = np.zeros((9, 9))
    img[4, 4] = 1.0
    out = o.resample_jacobian(o.Identity(), img, method='g', bound='e')
    assert np.count_nonzero(out) == 25 and np.isclose(out.sum(), 1.0)
    s = (3 + 1) ** (-1 / 8.0)
    w0 = 1.0 / sum(np.exp(-(s * a) ** 2 - (s * b) ** 2) for a in range(-2, 3) for b in range(-2, 3))
    assert np.isclose(out[4, 4], w0, rtol=1e-12)


def test_pincushion_kats(oracle):
    """tests/test_02_basic.py:331-419 (test_009_Quadpin, test_010_Cubicpin, test_011_Poly), same
    inputs, expected tables and tolerances."""
    x = np.arange(9.).reshape(3, 3)
    a = oracle.Quadpin()
    d1 = oracle.apply(a, x)
    assert np.allclose(d1, [[0, 1, 2.1818182], [3.5454545, 5.0909091, 6.8181818],
                            [8.7272727, 10.818182, 13.090909]], atol=1e-2)
    assert np.allclose(oracle.apply(a, d1, invert=True), x, atol=1e-2)
    a = oracle.Quadpin(length=5.0)
    d6 = np.array([[0, 0.92727273, 1.8909091], [2.8909091, 3.9272727, 5], [6.1090909, 7.2545455, 8.4363636]])
    assert np.allclose(oracle.apply(a, x), d6, atol=1e-2)
    assert np.allclose(oracle.apply(a, d6, invert=True), x, atol=1e-2)
    a = oracle.Quadpin(origin=5.0)
    d8 = oracle.apply(a, x)
    assert np.allclose(d8, [[-1.8181818, -0.090909091, 1.4545455], [2.8181818, 4, 5],
                            [6, 7.1818182, 8.5454545]], atol=1e-2)
    assert np.allclose(oracle.apply(a, d8, invert=True), x, atol=1e-2)
    a = oracle.Quadpin(strength=6.0)
    d12 = np.array([[0, 1, 3.7142857], [8.1428571, 14.285714, 22.142857], [31.714286, 43, 56]])
    assert np.allclose(oracle.apply(a, x), d12, atol=1e-2)
    assert np.allclose(oracle.apply(a, d12, invert=True), x, atol=1e-2)

    a = oracle.Cubicpin(strength=2)
    d1 = oracle.apply(a, x)
    assert np.allclose(d1, [[0, 0.6, 3.6], [11.4, 26.4, 51], [87.6, 138.6, 206.4]], atol=1e-2)
    assert np.allclose(oracle.apply(a, d1, invert=True),
                       [[0, 0.79501675, 1.6595098], [2.5116355, 3.3596201, 4.2058325],
                        [5.0511344, 5.8959087, 6.7403507]], atol=1e-2)
    a = oracle.Cubicpin(length=2.0)                         # strength 0: forward identity, reverse all-NaN
    assert np.allclose(oracle.apply(a, x), x, atol=1e-2)
    assert np.isnan(oracle.apply(a, x, invert=True)).all()

    assert np.allclose(oracle.apply(oracle.Poly(), x), x, atol=1e-2)
    a = oracle.Poly(length=3.0, strength=2.0)
    assert np.allclose(oracle.apply(a, x), [[0., 0.55555556, 1.55555556], [3., 4.88888889, 7.22222222],
                                            [10., 13.22222222, 16.88888889]], atol=1e-2)


def test_projection_closed_forms(oracle):
    """Calabretta & Greisen 2002: radii / ordinates of the added projections at angles where the
    formulas have closed-form values, and project(deproject()) round trips."""
    r0 = 180.0 / np.pi
    z = np.zeros(1)
    for proj, R in (("STG", 2 * r0), ("SIN", r0), ("ARC", 90.0), ("ZEA", r0 * np.sqrt(2.0))):
        x, y = oracle._project(proj, z + 90.0, z)              # phi = 90, theta = 0: on the +x axis
        assert np.allclose([x[0], y[0]], [R, 0.0], atol=1e-12), proj
        x, y = oracle._project(proj, z, z + 90.0)              # the native pole maps to the origin
        assert np.allclose([x[0], y[0]], [0.0, 0.0], atol=1e-12), proj
    x, y = oracle._project("TAN", z + 180.0, z + 45.0)         # R = r0 cot(45) = r0, phi = 180: +y
    assert np.allclose([x[0], y[0]], [0.0, r0], atol=1e-12)
    assert np.allclose(oracle._project("CEA", z + 30.0, z + 90.0), [[30.0], [r0]])
    assert np.allclose(oracle._project("MER", z + 30.0, z + 45.0)[1], r0 * np.log(np.tan(np.radians(67.5))))
    assert np.allclose(oracle._project("SFL", z + 60.0, z + 60.0), [[30.0], [60.0]])
    x, y = oracle._project("AIT", z + 180.0, z)                # end of the equator: x = 2 sqrt(2) r0
    assert np.allclose([x[0], y[0]], [2 * np.sqrt(2.0) * r0, 0.0])
    x, y = oracle._project("AIT", z, z + 90.0)                 # pole: y = sqrt(2) r0
    assert np.allclose([x[0], y[0]], [0.0, np.sqrt(2.0) * r0])
    rng = np.random.default_rng(8)
    phi = rng.random(500) * 340 - 170
    theta = rng.random(500) * 160 - 80
    for proj in ("TAN", "STG", "SIN", "ARC", "ZEA", "CAR", "CEA", "MER", "SFL", "AIT"):
        th = np.abs(theta) + 5 if proj in ("TAN", "SIN") else theta       # visible hemisphere only
        x, y = oracle._project(proj, phi, th)
        p2, t2 = oracle._deproject(proj, x, y)
        assert np.allclose(p2, phi, atol=1e-9) and np.allclose(t2, th, atol=1e-9), proj


def test_matvec_rounding_matches_numpy(oracle):
    """The order in which np.matmul(m, v[..., None]) accumulates a row (basic.py:150-152, 168-170), restated in
    oracle_c.c::orc_matvec_fma and used by the CUDA fp64 mode (K1/K3 for 2 x 2, K3-ND beyond): bit for bit for
    2 x 2 and 3 x 3 matrices in both memory orders and for transposed 4 x 4 views.  Re-measured here against
    the NumPy of whatever machine runs the CPU suite."""
    rng = np.random.default_rng(20)
    for n, orders in ((2, (False, True)), (3, (False, True)), (4, (True,))):
        for f_order in orders:
            for _ in range(200):
                m = rng.standard_normal((n, n))
                v = rng.standard_normal((7, n)) * 10.0 ** rng.integers(-3, 4)
                mm = np.ascontiguousarray(m.T).T if f_order else m
                want = np.matmul(mm, v[..., None])[..., 0]
                got = oracle._matvec(m, v, f_order)
                assert np.array_equal(got, want), (n, f_order)
    # larger products are not pinned: a few ulp of sum |m_ik v_k|
    for n in (4, 5, 8):
        m = rng.standard_normal((n, n))
        v = rng.standard_normal((50, n))
        want = np.matmul(m, v[..., None])[..., 0]
        got = oracle._matvec(m, v, False)
        scale = np.abs(m) @ np.abs(v).T
        assert np.all(np.abs(got - want) <= 4 * np.finfo(float).eps * scale.T)
