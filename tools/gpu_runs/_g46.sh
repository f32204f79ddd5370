cd $GRAFT_REPO_ROOT
timeout 300 python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0,'.')
import transform_b200 as tfm
from transform_b200 import _lib
for seed in range(6):
    rng = np.random.default_rng(500 + seed)
    sh, sw = int(rng.integers(160, 400)), 8 * int(rng.integers(24, 60))
    ny, nx = int(rng.integers(3, 300)), int(rng.integers(700, 1500))
    src = rng.random((sh, sw)).astype(np.float32)
    c = [sw / 2.0 + rng.uniform(-3, 3), sh / 2.0 + rng.uniform(-3, 3)]
    rmax = 0.45 * min(sh, sw)
    tr = tfm.Composition([tfm.Scale([nx / (2 * np.pi), max(ny, 8) / rmax * rng.uniform(1.05, 1.6)]),
                          tfm.Radial(origin=c, ccw=bool(seed & 1))])
    dsrc = torch.from_numpy(src).cuda()
    for phot in ("radiance", "flux"):
        got = tr.resample(dsrc, method='G', precision='fp32', shape=(ny, nx), phot=phot)
        k1 = _lib.last_kernel()
        one = tr.resample(dsrc, method='G', precision='fp32', shape=(ny, nx), phot=phot, tuning=524288)
        k2 = _lib.last_kernel()
        print(seed, phot, (ny, nx), k1, "|", k2, "ops", [ (o.kind, round(o.p[3],3)) for o in tr._int_chain(True)])
PY
