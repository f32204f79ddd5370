#!/usr/bin/env python3
"""One kernel of the library, a few launches (target of `ncu -k regex:... -c 1`):

    python tools/prof_all.py <case>

cases: linear_shift linear_named linear_rot45 nearest cubic cubic_shift cubic_ldg hann lanczos linear_polar linear_parity
       jacobian_polar jacobian_polar_rows jacobian_polar_single jacobian_constj jacobian_fd jacobian_parity jacobian_hanning
       wcs_batch wcs_single apply_affine apply_radial apply_nd fits_m32 fits_i16"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
import transform_b200 as t  # noqa: E402
from transform_b200 import _lib  # noqa: E402

case = sys.argv[1]
n = 8192
g = torch.Generator(device="cuda").manual_seed(7)
c = (n - 1) / 2.0
named = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
polar = t.Composition([t.Scale([n / (2 * np.pi), np.sqrt(2.0)]), t.Radial(origin=[c, c])])
down = t.Composition([t.Offset([n / 2.0 + 0.3, n / 2.0 - 0.2]), t.Scale([0.6, 0.7]), t.Rotation(25, unit="deg"),
                      t.Offset([-n / 2.0, -n / 2.0])])
reps = 4
if case.startswith(("linear", "nearest", "cubic", "hann", "lanczos", "jacobian")):
    src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
    f32 = torch.empty_like(src)
    f64 = torch.empty((n, n), device="cuda", dtype=torch.float64) if case.endswith("parity") else None
    calls = {
        "linear_shift": lambda: t.Offset([3.25, -7.5]).resample(src, method="linear", precision="fp32", out=f32),
        "linear_named": lambda: named.resample(src, method="linear", precision="fp32", out=f32),
        "linear_rot45": lambda: t.Wrap(t.Rotation(45, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])).resample(
            src, method="linear", precision="fp32", out=f32),
        "nearest": lambda: named.resample(src, method="nearest", precision="fp32", out=f32),
        "cubic": lambda: named.resample(src, method="cubic", precision="fp32", out=f32),
        "cubic_shift": lambda: t.Offset([3.25, -7.5]).resample(src, method="cubic", precision="fp32", out=f32),
        "cubic_ldg": lambda: named.resample(src, method="cubic", precision="fp32", out=f32, tuning=16),
        "hann": lambda: named.resample(src, method="hann", precision="fp32", out=f32),
        "lanczos": lambda: named.resample(src, method="zlanczos", precision="fp32", out=f32),
        "linear_polar": lambda: polar.resample(src, method="linear", precision="fp32", out=f32),
        "linear_parity": lambda: named.resample(src, method="linear", out=f64),
        "jacobian_polar": lambda: polar.resample(src, method="Gaussian", precision="fp32", out=f32),
        "jacobian_polar_single": lambda: polar.resample(src, method="Gaussian", precision="fp32", out=f32, tuning=262144),
        "jacobian_polar_rows": lambda: polar.resample(src, method="Gaussian", precision="fp32", out=f32, tuning=524288),
        "jacobian_constj": lambda: down.resample(src, method="Gaussian", precision="fp32", out=f32),
        "jacobian_fd": lambda: polar.resample(src, method="Gaussian", precision="fp32", out=f32, tuning=512),
        "jacobian_parity": lambda: polar.resample(src, method="Gaussian", out=f64),
        "jacobian_hanning": lambda: polar.resample(src, method="Hanning", precision="fp32", out=f32),
    }
    fn = calls[case]
    if case == "jacobian_parity":
        reps = 2
elif case.startswith("wcs"):
    import bench
    total, nb, m = bench.config4_transform(t)
    nb = 8 if case == "wcs_batch" else 1
    frames = torch.rand((nb, m, m) if nb > 1 else (m, m), generator=g, device="cuda", dtype=torch.float32)
    out = torch.empty_like(frames)
    fn = lambda: total.resample(frames, method="linear", precision="fp32", out=out, shape=tuple(frames.shape))  # noqa: E731
elif case == "apply_nd":
    v = (torch.rand((2 * 10 ** 8, 3), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 8192
    tr = t.Composition([t.Scale([1.5, 2, 0.5]), t.Rotation(euler=np.array([10.0, 20.0, 30.0]), unit="deg"), t.Offset([3, 4, 5])])
    fn = lambda: tr.apply(v)  # noqa: E731
elif case.startswith("apply"):
    v = (torch.rand((2 * 10 ** 8, 2), generator=g, device="cuda", dtype=torch.float64) - 0.5) * 8192
    mid = t.Rotation(30, unit="deg") if case == "apply_affine" else t.Radial(origin=[0.1, 0.2])
    tr = t.Composition([t.Scale([1.5, 2]), mid, t.Offset([3, 4])]).inverse()
    fn = lambda: tr.apply(v)  # noqa: E731
elif case.startswith("fits"):
    bitpix = -32 if case == "fits_m32" else 16
    nbytes = n * n * abs(bitpix) // 8
    raw = torch.randint(0, 255, (nbytes,), generator=g, device="cuda", dtype=torch.uint8)
    fn = lambda: t.decode_fits_data(raw, bitpix, (n, n))  # noqa: E731
else:
    raise SystemExit("unknown case " + case)
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print(case, _lib.last_kernel())
