#!/usr/bin/env python3
"""GPU check of the K2 fast-path variants against each other (per-row tables vs per-pixel closed form of the
analytic polar kernel, tuning bit 262144), row blocks against the full frame, and timings at 8192^2."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import transform_b200 as t
from transform_b200 import _lib, _engine

def chains(n, c):
    return {
        "polar": t.Composition([t.Scale([n / (2 * np.pi), np.sqrt(2.0)]), t.Radial(origin=[c, c - 0.34])]),
        "polar_ccw": t.Composition([t.Scale([n / (2 * np.pi), 1.3]), t.Radial(origin=[c, c - 0.34], ccw=True)]),
        "polar_a2_ccw": t.Composition([t.Scale([n / (2 * np.pi), 1.3]), t.Radial(origin=[c, c - 0.34], ccw=True),
                                       t.Scale([0.9, 1.1]), t.Offset([5.0, -3.0])]),
        "logpolar": t.Composition([t.Scale([n / (2 * np.pi), n / 6.0]), t.Radial(origin=[c, c - 0.34], r0=1.0)]),
    }

n = 768
c = (n - 1) / 2.0 + 0.123
g = torch.Generator(device="cuda").manual_seed(3)
src = torch.rand((n, n), device="cuda", dtype=torch.float32, generator=g)
for name, tr in chains(n, c).items():
    for kw in ({}, {"phot": "flux"}):
        a = tr.resample(src, method="G", precision="fp32", **kw); ka = _lib.last_kernel()
        b = tr.resample(src, method="G", precision="fp32", tuning=524288, **kw); kb = _lib.last_kernel()
        d = (a - b).abs()
        print(name, kw, ka, "|", kb, "max %.3g" % float(d.max()), "n>1e-5", int((d > 1e-5).sum()), "n>1e-4", int((d > 1e-4).sum()))
    full = tr.resample(src, method="G", precision="fp32")
    ops = tr._int_chain(True)
    oks = []
    for (y0, y1) in ((0, 5), (5, 400), (400, 768), (767, 768), (1, 2), (33, 97)):
        blk = _engine.resample_jacobian(ops, src, (y1 - y0, n), 'g', 't', precision='fp32', dst_origin=(0, y0), grid_full=(n, n))
        oks.append(bool(torch.equal(blk, full[y0:y1])))
    print("   row blocks equal the full frame:", oks)
srcn = src.clone(); srcn[300, 500] = float("nan"); srcn[100, 77] = float("inf")
trn = chains(768, c)["polar"]
a = trn.resample(srcn, method="G", precision="fp32"); b = trn.resample(srcn, method="G", precision="fp32", tuning=524288)
print("non-finite sets equal:", bool(torch.equal(torch.isfinite(a), torch.isfinite(b))), int((~torch.isfinite(a)).sum()), int((~torch.isfinite(b)).sum()))
n = 8192
src = torch.rand((n, n), device="cuda", dtype=torch.float32, generator=g)
dst = torch.empty_like(src)
cc = (n - 1) / 2.0
tr = t.Composition([t.Scale([n / (2 * np.pi), n / (n / np.sqrt(2.0))]), t.Radial(origin=[cc, cc])])
down = t.Composition([t.Offset([n / 2.0 + 0.3, n / 2.0 - 0.2]), t.Scale([0.6, 0.7]), t.Rotation(25, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
def timed(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for tun in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,131072,524288,262144").split(",")]:
    ms = timed(lambda: tr.resample(src, method="G", precision="fp32", out=dst, tuning=tun))
    print("8192 polar tuning", tun, _lib.last_kernel(), "%.4f ms" % ms)
ms = timed(lambda: down.resample(src, method="G", precision="fp32", out=dst))
print("8192 downscale", _lib.last_kernel(), "%.4f ms" % ms)
a = tr.resample(src, method="G", precision="fp32"); b = tr.resample(src, method="G", precision="fp32", tuning=524288)
d = (a - b).abs(); print("8192 default vs one-pixel rows: max %.3g" % float(d.max()), "n>1e-5", int((d > 1e-5).sum()), "n>1e-4", int((d > 1e-4).sum()))
