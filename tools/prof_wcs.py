"""Run the TAN -> CAR remap kernel a few times (for ncu): python tools/prof_wcs.py [batch|single]"""
import sys
import torch
sys.path.insert(0, '.')
import transform_b200 as t, bench
mode = sys.argv[1] if len(sys.argv) > 1 else "batch"
total, nb, n = bench.config4_transform(t)
nb = 8 if mode == "batch" else 1
g = torch.Generator(device="cuda").manual_seed(7)
frames = torch.rand((nb, n, n) if nb > 1 else (n, n), generator=g, device="cuda", dtype=torch.float32)
out = torch.empty_like(frames)
for _ in range(4):
    total.resample(frames, method="linear", precision="fp32", out=out, shape=tuple(frames.shape))
torch.cuda.synchronize()
print("done", float(out.sum()))
