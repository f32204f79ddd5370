"""Run one resample case a few times (for ncu): python tools/prof_case.py shift|rot45|named|polar [method] [tuning]"""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import transform_b200 as t
case = sys.argv[1] if len(sys.argv) > 1 else "shift"
method = sys.argv[2] if len(sys.argv) > 2 else "linear"
tuning = int(sys.argv[3]) if len(sys.argv) > 3 else 0
n = 8192
g = torch.Generator(device="cuda").manual_seed(7)
src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
dst = torch.empty_like(src)
tr = {"shift": t.Offset([3.25, -7.5]),
      "rot45": t.Wrap(t.Rotation(45, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])),
      "named": t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])]),
      "polar": t.Composition([t.Scale([n / (2 * np.pi), np.sqrt(2.0)]), t.Radial(origin=[(n - 1) / 2.0] * 2)])}[case]
for _ in range(6):
    tr.resample(src, method=method, precision="fp32", out=dst, tuning=tuning)
torch.cuda.synchronize()
print("done", float(dst.sum()))
