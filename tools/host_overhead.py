"""Host-side cost of one resample call: a 256^2 frame makes the kernel negligible (GPU box)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import transform_b200 as t
n = 256
src = torch.rand((n, n), device="cuda", dtype=torch.float32)
dst = torch.empty_like(src)
tr = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
for m, kw in (("linear", {}), ("nearest", {}), ("Gaussian", {})):
    for _ in range(50):
        tr.resample(src, method=m, precision="fp32", out=dst, **kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    N = 2000
    for _ in range(N):
        tr.resample(src, method=m, precision="fp32", out=dst, **kw)
    torch.cuda.synchronize()
    print(m, "%.1f us per call" % ((time.perf_counter() - t0) / N * 1e6))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(2000):
    tr.resample(src, method="linear", precision="fp32", out=dst)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
