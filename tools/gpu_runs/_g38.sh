cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_resample.py tests/test_gpu_fuzz.py tests/test_gpu_golden.py -q -m gpu -x 2>&1 | tail -4
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
for name, tr in (("shift", t.Offset([3.25,-7.5])), ("rot5", t.Wrap(t.Rotation(5, unit='deg'), t.Offset([-n/2.0,-n/2.0]))), ("rot45", t.Wrap(t.Rotation(45, unit='deg'), t.Offset([-n/2.0,-n/2.0])))):
    for tun in (0, 16):
        ms = timed(lambda: tr.resample(src, method="cubic", precision="fp32", out=dst, tuning=tun))
        print("cubic", name, tun, _lib.last_kernel(), "%.4f ms  %.1f %% of roofline" % (ms, 100*2*n*n*4/(ms*1e-3)/1e9/6547.5))
PY
