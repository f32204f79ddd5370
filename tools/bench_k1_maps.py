#!/usr/bin/env python3
"""Linear resampling at 8192^2 float32 through several affine maps, per kernel variant (GPU box).
tuning 0 = auto, 128 = first texture-gather form, 256 = TMA-staged shared-memory taps (one box per CTA),
16 = plain LDG gather."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import transform_b200 as t  # noqa: E402

n = 8192
g = torch.Generator(device="cuda").manual_seed(7)
src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
dst = torch.empty_like(src)
ref = torch.empty_like(src)
maps = {"named": t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])]),
        "shift": t.Offset([3.25, -7.5]), "scale0p8": t.Scale(0.8, dim=2), "scale2": t.Scale(2.0, dim=2),
        "scale0p4": t.Scale(0.4, dim=2),
        "rot5": t.Wrap(t.Rotation(5, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])),
        "rot45": t.Wrap(t.Rotation(45, unit="deg"), t.Offset([-n / 2.0, -n / 2.0]))}
peak = 6547.5
res = {}
for name, tr in maps.items():
    row = {}
    for tun in (0, 8192, 128, 256, 16):
        f = lambda: tr.resample(src, method="linear", precision="fp32", out=dst, tuning=tun)
        f(); f(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            f()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        row[tun] = ms
        if tun == 0:
            ref.copy_(dst)
        else:
            assert torch.equal(ref, dst), (name, tun)
    res[name] = row
    print(name, " ".join("tuning %d: %.4f ms (%.0f%%)" % (k, v, 100 * 2 * n * n * 4 / (v * 1e-3) / 1e9 / peak) for k, v in row.items()), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/k1_maps.json", "w"), indent=1)
