"""K2 timings per filter shape (GPU box): 8192^2 float32, polar unwrap and an affine downsample."""
import sys
sys.path.insert(0, '.')
import numpy as np, torch
import transform_b200 as t
n = 8192
src = torch.rand((n, n), device="cuda", dtype=torch.float32)
dst = torch.empty_like(src)
c = (n - 1) / 2.0
maps = {"polar": t.Composition([t.Scale([n / (2 * np.pi), np.sqrt(2.0)]), t.Radial(origin=[c, c])]),
        "affine0.6": t.Composition([t.Offset([n / 2.0, n / 2.0]), t.Scale([0.6, 0.7]), t.Rotation(20, unit="deg"),
                                    t.Offset([-n / 2.0, -n / 2.0])])}
for name, tr in maps.items():
    for m in ("Gaussian", "Hanning", "Rounded", "Linear"):
        f = lambda: tr.resample(src, method=m, precision="fp32", out=dst)
        f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            f()
        e1.record(); torch.cuda.synchronize()
        print(name, m, "%.2f ms" % (e0.elapsed_time(e1) / 3), flush=True)
