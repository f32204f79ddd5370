#!/usr/bin/env python3
"""A/B timing of kernel build variants (run on the GPU box).

    python tools/ab_variants.py k2 name1=path1.so name2=path2.so ...

Each variant is timed in its own process (the library is loaded once per process)."""
import json
import os
import subprocess
import sys

CHILD = r'''
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import transform_b200 as t
n = 8192
g = torch.Generator(device="cuda").manual_seed(7)
src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
dst = torch.empty_like(src)
c = (n - 1) / 2.0
tr_rad = t.Composition([t.Scale([n / (2 * np.pi), n / (n / np.sqrt(2.0))]), t.Radial(origin=[c, c])])
tr_aff = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
def timed(fn, reps):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
what = sys.argv[1]
out = {}
if what == "k2":
    from transform_b200 import _lib
    tr_dn = t.Composition([t.Offset([n / 2.0 + 0.3, n / 2.0 - 0.2]), t.Scale([0.6, 0.7]), t.Rotation(25, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
    for tun in [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "0").split(",")]:
        out["radial_t%d_ms" % tun] = round(timed(lambda: tr_rad.resample(src, method="G", precision="fp32", out=dst, tuning=tun), 10), 4)
        out["radial_t%d_k" % tun] = _lib.last_kernel()
        out["radial_t%d_sum" % tun] = float(dst.double().sum().item())      # variants must agree bit for bit
        out["radial_t%d_absdiff_checksum" % tun] = float((dst[1:] - dst[:-1]).abs().double().sum().item())
        out["named_t%d_ms" % tun] = round(timed(lambda: tr_aff.resample(src, method="G", precision="fp32", out=dst, tuning=tun), 10), 4)
        out["down_t%d_ms" % tun] = round(timed(lambda: tr_dn.resample(src, method="G", precision="fp32", out=dst, tuning=tun), 10), 4)
else:
    maps = {"named": tr_aff, "shift": t.Offset([3.25, -7.5]), "scale0p8": t.Scale(0.8, dim=2),
            "rot5": t.Wrap(t.Rotation(5, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])),
            "rot45": t.Wrap(t.Rotation(45, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])),
            "polar": tr_rad}
    for name, tr in maps.items():
        out["linear_%s_fp32_ms" % name] = round(timed(lambda: tr.resample(src, method="linear", precision="fp32", out=dst), 30), 4)
print(json.dumps(out))
'''

what = sys.argv[1]
tunings = "0"
specs = sys.argv[2:]
if specs and specs[0].startswith("tunings="):
    tunings = specs[0].split("=", 1)[1]
    specs = specs[1:]
res = {}
for spec in specs:
    name, path = spec.split("=", 1)
    env = dict(os.environ)
    if path != "default":
        env["TFM_B200_LIB"] = os.path.abspath(path)
    r = subprocess.run([sys.executable, "-c", CHILD, what, tunings], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-400:]
    print(name, line, flush=True)
    try:
        res[name] = json.loads(line)
    except Exception:
        res[name] = {"error": line}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/ab_%s.json" % what, "w"), indent=1)
