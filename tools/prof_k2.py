"""Run a Jacobian-filter step a few times (for ncu): python tools/prof_k2.py [polar|affine|named] [fp32|fp64]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import transform_b200 as t
n = 8192
g = torch.Generator(device="cuda").manual_seed(7)
src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
dst = torch.empty_like(src)
case = sys.argv[1] if len(sys.argv) > 1 else "polar"
prec = sys.argv[2] if len(sys.argv) > 2 else "fp32"
tr = {"polar": t.Composition([t.Scale([n / (2 * np.pi), np.sqrt(2.0)]), t.Radial(origin=[(n - 1) / 2.0] * 2)]),
      "affine": t.Composition([t.Scale([0.6, 0.7]), t.Rotation(25, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])]),
      "named": t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])}[case]
if prec != "fp32":
    dst = dst.double()
for _ in range(4):
    tr.resample(src, method='G', precision=prec, out=dst)
torch.cuda.synchronize()
print("done", float(dst.sum()))
