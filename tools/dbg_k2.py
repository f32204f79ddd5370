"""Debug: compare K2 tap modes with each other and with the oracle on a few rows (GPU box)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import transform_b200 as t
from transform_b200 import _lib
import oracle_np as o
from conftest import to_oracle
n = 8192
g = torch.Generator(device="cuda").manual_seed(1000)
src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
c = (n - 1) / 2.0
rad = t.Composition([t.Scale([n / (2 * np.pi), n / (n / np.sqrt(2.0))]), t.Radial(origin=[c, c])])
outs = {}
for tun in (0, 32768, 4096, 512):
    outs[tun] = rad.resample(src, method='Gaussian', precision='fp32', tuning=tun)
    print(tun, _lib.last_kernel())
for a in (32768, 4096, 512):
    d = (outs[a] - outs[0]).abs()
    m = float(d.max()); idx = int(d.argmax()); y, x = idx // n, idx % n
    print("tuning", a, "vs 0: max", m, "at", (y, x), "count>1e-4", int((d > 1e-4).sum()), "vals", float(outs[a][y, x]), float(outs[0][y, x]))
    ys, xs = torch.nonzero(d > 1e-4, as_tuple=True)
    if len(ys):
        print("   rows range", int(ys.min()), int(ys.max()), "cols range", int(xs.min()), int(xs.max()), "sample", [(int(ys[i]), int(xs[i])) for i in range(0, len(ys), max(1, len(ys)//8))][:8])
src64 = src.cpu().numpy().astype(np.float64)
otr = to_oracle(rad)
cap = 2 * int(np.ceil(0.25 * n))
for y in (0, 1, 700, 3000, 5793, 6500, 8191):
    want = o.resample_jacobian(otr, src64, method='g', rows=(y, y + 1), max_reg=cap)[0]
    for tun in outs:
        got = outs[tun][y].cpu().numpy()
        e = np.abs(got - want)
        print("row", y, "tuning", tun, "max err %.3e median %.3e n>3e-4 %d argmax %d" % (e.max(), np.median(e), (e > 3e-4).sum(), e.argmax()))
