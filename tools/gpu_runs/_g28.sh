cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_remap.py -q -m gpu -x 2>&1 | tail -12
timeout 600 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0,'.')
import transform_b200 as t, bench
from transform_b200 import _lib
total, nb, n = bench.config4_transform(t)
g = torch.Generator(device="cuda").manual_seed(7)
frames = torch.rand((nb, n, n), generator=g, device="cuda", dtype=torch.float32)
outb = torch.empty_like(frames)
def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
for tun in (0, 1048576):
    ms = timed(lambda: total.resample(frames, method="linear", precision="fp32", out=outb, shape=(nb, n, n), tuning=tun), 5)
    print("config4 tuning", tun, _lib.last_kernel(), "ms %.4f" % ms, "frac %.4f" % (2*nb*n*n*4/(ms*1e-3)/1e9/6547.5))
aff = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
for tun in (0, 1048576):
    ms = timed(lambda: aff.resample(frames, method="linear", precision="fp32", out=outb, shape=(nb, n, n), tuning=tun), 5)
    print("affine batch tuning", tun, _lib.last_kernel(), "ms %.4f" % ms, "frac %.4f" % (2*nb*n*n*4/(ms*1e-3)/1e9/6547.5))
PY
