#!/usr/bin/env python3
"""Kernel-only timing of the windowed-filter variant of K1 on the bench geometry (GPU box).
    python tools/bench_filters.py [N]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import transform_b200 as t  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
src = torch.rand((n, n), device="cuda", dtype=torch.float32)
tr = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit='deg'), t.Offset([-n / 2.0, -n / 2.0])])
ident = t.Offset([0.25, 0.25])
res = {}
for prec in ("fp32", "fp64"):
    out = torch.empty((n, n), device="cuda", dtype=torch.float32 if prec == "fp32" else torch.float64)
    for name, trf in (("affine", tr), ("shift", ident)):
        for m, ob in (("h", None), ("t", None), ("l", 1.0), ("z", None), ("g", None), ("s", None), ("h", 2.0)):
            reps = 3 if m in "sg" and prec == "fp64" else 5
            trf.resample(src, method=m, precision=prec, out=out, oblur=ob)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                trf.resample(src, method=m, precision=prec, out=out, oblur=ob)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            res["%s/%s/%s/%s" % (prec, name, m, ob)] = round(ms, 4)
            print(prec, name, m, ob, "%.3f ms" % ms, "%.0f Mpix/s" % (n * n / ms / 1e3), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/filters.json", "w"), indent=1)

# K0: FITS decode throughput (8192^2 samples)
import transform_b200 as t2  # noqa: E402
peak = 6547.5
nn = n * n
for bitpix, out_b in ((-32, 4), (16, 8), (-64, 8)):
    raw = torch.randint(0, 255, (nn * abs(bitpix) // 8,), device="cuda", dtype=torch.uint8)
    t2.decode_fits_data(raw, bitpix, (n, n))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        t2.decode_fits_data(raw, bitpix, (n, n))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gbs = nn * (abs(bitpix) // 8 + out_b) / (ms * 1e-3) / 1e9
    print("fits decode BITPIX %d: %.3f ms  %.0f GB/s (%.0f%% of %.0f)" % (bitpix, ms, gbs, 100 * gbs / peak, peak), flush=True)
    del raw
