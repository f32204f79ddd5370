cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_remap.py -q -m gpu -x 2>&1 | tail -12
timeout 600 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0,'.')
import transform_b200 as t, bench
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
ms = timed(lambda: total.resample(frames, method="linear", precision="fp32", out=outb, shape=(nb, n, n)), 3)
print("config4 ms", ms, "frac", 2*nb*n*n*4/(ms*1e-3)/1e9/6547.5)
one = frames[0].contiguous(); o1 = torch.empty_like(one)
ms1 = timed(lambda: total.resample(one, method="linear", precision="fp32", out=o1), 10)
from transform_b200 import _lib
print("single frame ms", ms1, _lib.last_kernel(), "frac", 2*n*n*4/(ms1*1e-3)/1e9/6547.5)
PY
