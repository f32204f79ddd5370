"""GPU: seeded random sweeps of the whole option space of K1 / K2 / K3 against the oracle -- random
affine / radial / pincushion chains, ragged shapes, every boundary (also mixed per axis), fill
values, source dtypes, methods.  Parity mode: bit-exact for affine chains, 1e-12 through
transcendental ops; fp32 mode within the stated tolerances."""
import numpy as np
import pytest

from conftest import to_oracle, assert_close_rel

pytestmark = pytest.mark.gpu

BOUNDS = ['t', 'e', 'p', 'm']


def _random_chain(tfm, rng, ny, nx, kind):
    ang = rng.uniform(-180, 180)
    sc = rng.uniform(0.5, 2.0, size=2) if rng.random() < 0.5 else rng.uniform(0.5, 2.0)
    off = rng.uniform(-0.6, 0.1, size=2) * np.array([nx, ny])
    parts = [tfm.Scale(sc, dim=2) if np.ndim(sc) == 0 else tfm.Scale(sc), tfm.Rotation(ang, unit='deg'),
             tfm.Offset(off)]
    if kind == "affine":
        rng.shuffle(parts)
        tr = tfm.Composition(parts[:rng.integers(1, 4)])
        return tr.inverse() if rng.random() < 0.3 else tr
    if kind == "radial":
        rad = tfm.Radial(origin=[rng.uniform(0.2, 0.8) * nx, rng.uniform(0.2, 0.8) * ny],
                         ccw=bool(rng.random() < 0.5), pos_only=bool(rng.random() < 0.5),
                         r0=(None if rng.random() < 0.6 else rng.uniform(0.5, 3.0)))
        pre = tfm.Scale([nx / (2 * np.pi), ny / (3.0 if rad.params['r0'] else 0.7 * max(nx, ny))])
        chain = [pre, rad]
        if rng.random() < 0.5:
            chain.append(tfm.Offset(rng.uniform(-3, 3, size=2)))
        return tfm.Composition(chain)
    pin = (tfm.Quadpin(idim=2, odim=2, strength=rng.uniform(-0.4, 0.4), length=rng.uniform(20, 80),
                       origin=rng.uniform(0.3, 0.7, size=2) * np.array([nx, ny]))
           if rng.random() < 0.5 else
           tfm.Poly(idim=2, odim=2, strength=rng.uniform(-0.2, 0.2), length=rng.uniform(40, 90),
                    degree=int(rng.integers(2, 5)), origin=rng.uniform(0.3, 0.7, size=2) * np.array([nx, ny])).inverse())
    return tfm.Composition([tfm.Offset(rng.uniform(-2, 2, size=2)), pin]) if rng.random() < 0.5 else pin


@pytest.mark.parametrize("seed", range(64))
def test_k1_random_cases(tfm, oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    ny, nx = int(rng.integers(3, 90)), int(rng.integers(3, 120))
    kind = ("affine", "affine", "radial", "pin")[seed % 4]
    tr = _random_chain(tfm, rng, ny, nx, kind)
    dt = (np.float32, np.float64, np.int16, np.uint8)[int(rng.integers(0, 4))]
    src = (rng.random((ny, nx)) * 200).astype(dt)
    shape = None if rng.random() < 0.4 else (int(rng.integers(1, 100)), int(rng.integers(1, 130)))
    bound = BOUNDS[int(rng.integers(0, 4))]
    if rng.random() < 0.3:
        bound = [BOUNDS[int(rng.integers(0, 4))], BOUNDS[int(rng.integers(0, 4))]]
    fill = int(rng.integers(-5, 6))
    for m in ('nearest', 'linear', 'cubic', 'hann', 'gaussian'):
        if m == 'cubic' and np.dtype(dt).kind in 'iu':
            continue                                   # integer cubic overflows in the reference's own dtype
        got = tr.resample(src, method=m, bound=bound, shape=shape, fillvalue=fill)
        want = oracle.resample(to_oracle(tr), src, method=m, bound=bound, shape=shape, fillvalue=fill)
        what = (seed, kind, m, bound, shape, np.dtype(dt).name)
        assert got.shape == want.shape and got.dtype == want.dtype, what
        if kind == "affine" and m in ('nearest', 'linear', 'cubic'):
            assert np.array_equal(got, want, equal_nan=True), what
        elif m == 'nearest':
            assert (got != want).sum() <= max(2, got.size // 500), what      # ulp-level ties only
        else:
            assert_close_rel(got, want, 1e-10, what)
        if dt == np.float32 and m in ('linear', 'cubic', 'hann'):
            fast = tr.resample(src, method=m, bound=bound, shape=shape, fillvalue=fill, precision='fp32')
            assert_close_rel(fast, want, 3e-5, ("fp32",) + what)


@pytest.mark.parametrize("seed", range(8))
def test_k3_random_cases(tfm, oracle, seed):
    rng = np.random.default_rng(2000 + seed)
    kind = ("affine", "radial", "pin", "affine")[seed % 4]
    tr = _random_chain(tfm, rng, 64, 80, kind)
    n = int(rng.integers(1, 5000))
    width = int(rng.integers(2, 5))
    v = rng.uniform(1.0, 60.0, size=(n, width))
    for invert in (False, True):
        if (invert and tr.no_reverse) or (not invert and tr.no_forward):
            continue
        got = tr.apply(v, invert=invert)
        want = oracle.apply(to_oracle(tr), v, invert=invert)
        if kind == "affine":
            assert np.array_equal(got, want), (seed, invert)
        else:
            assert_close_rel(got, want, 1e-11, (seed, kind, invert))
        assert np.array_equal(got[:, 2:], v[:, 2:])          # extra components pass through


@pytest.mark.parametrize("seed", range(8))
def test_k2_random_cases(tfm, oracle, seed):
    rng = np.random.default_rng(3000 + seed)
    ny, nx = int(rng.integers(24, 70)), int(rng.integers(24, 90))
    # anisotropic on purpose ("Degenerate SVD")
    tr = tfm.Composition([tfm.Offset(rng.uniform(0.3, 0.7, size=2) * np.array([nx, ny])),
                          tfm.Scale(rng.uniform(0.35, 1.6, size=2) * np.array([1.0, 1.23])),
                          tfm.Rotation(rng.uniform(-90, 90), unit='deg'),
                          tfm.Offset(-rng.uniform(0.3, 0.7, size=2) * np.array([nx, ny]))])
    src = rng.random((ny, nx)).astype(np.float32)
    bound = BOUNDS[int(rng.integers(0, 4))]
    cap = max(2, 2 * int(np.ceil(0.25 * max(ny, nx))))
    for letter, filt in (('G', 'g'), ('H', 'h'), ('R', 't'), ('L', 'l')):
        got = tr.resample(src, method=letter, bound=bound)
        want = oracle.resample_jacobian(to_oracle(tr), src, method=filt, bound=bound, max_reg=cap)
        assert_close_rel(got, want, 1e-10, (seed, letter, bound))
        fast = tr.resample(src, method=letter, bound=bound, precision='fp32')
        err = np.abs(fast - want)
        # isolated pixels may take the neighbouring footprint size (ceil of an fp32 quantity): their number is
        # bounded AND so is their size -- a window one tap wider or narrower trades one row/column of the weight
        # (at most ~1/3 of it for the smallest, 3 x 3, window) of uniform [0,1) samples for another
        assert np.median(err) < 2e-5 and (err > 5e-4).sum() <= max(3, err.size // 200), (seed, letter, bound, err.max())
        assert err.max() <= 0.3, (seed, letter, bound, err.max())
