"""Run the bench's Jacobian step a few times (for ncu): python tools/prof_k2.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import transform_b200 as t
n = 8192
g = torch.Generator(device="cuda").manual_seed(7)
src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
dst = torch.empty_like(src)
tr = t.Composition([t.Scale([n / (2 * np.pi), np.sqrt(2.0)]), t.Radial(origin=[(n - 1) / 2.0] * 2)])
for _ in range(4):
    tr.resample(src, method='G', precision="fp32", out=dst)
torch.cuda.synchronize()
print("done", float(dst.sum()))
