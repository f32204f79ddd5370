cd $GRAFT_REPO_ROOT
timeout 600 python tools/check_k2.py 0,2048,131072,133120 2>&1 | tail -9
timeout 300 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0,'.')
import transform_b200 as t
from transform_b200 import _lib
def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
n=8192
src = torch.rand((n, n), device="cuda", dtype=torch.float32); dst=torch.empty_like(src)
tr = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
for tun in (0, 8192):
    ms = timed(lambda: tr.resample(src, method="cubic", precision="fp32", out=dst, tuning=tun))
    print("cubic named", tun, _lib.last_kernel(), "%.4f ms" % ms)
PY
