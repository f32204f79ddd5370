cd $GRAFT_REPO_ROOT
timeout 900 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0,'.')
import transform_b200 as t
from transform_b200 import _lib
def timed(fn, reps):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
g = torch.Generator(device="cuda").manual_seed(7)
for n in (1024, 8192):
    src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
    c = (n - 1) / 2.0
    rad = t.Composition([t.Scale([n / (2 * np.pi), n / (n / np.sqrt(2.0))]), t.Radial(origin=[c, c])])
    down = t.Composition([t.Offset([n / 2.0 + 0.3, n / 2.0 - 0.2]), t.Scale([0.6, 0.7]), t.Rotation(25, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
    for name, tr in (("polar", rad), ("down", down)):
        for meth in ("Gaussian", "Hanning") if n == 1024 else ("Gaussian",):
            for bound in ("t", "m"):
                for ob in (1.0, 1.5) if n == 1024 else (1.0,):
                    a = tr.resample(src, method=meth, bound=bound, oblur=ob)
                    b = tr.resample(src, method=meth, bound=bound, oblur=ob, tuning=2097152)
                    c4 = tr.resample(src, method=meth, bound=bound, oblur=ob, tuning=4194304)
                    print('   one exp per tap vs general loop:', 'bit-identical' if torch.equal(c4, b) else 'DIFF', ' recurrence rel diff %.3g' % float(((a-b).abs()/(b.abs()+1e-300)).max()))
                    print(n, name, meth, bound, ob, _lib.last_kernel())
    if n == 8192:
        o = torch.empty((n, n), device="cuda", dtype=torch.float64)
        for tun in (0, 4194304, 2097152):
            print("parity polar 8192 tuning", tun, "%.3f ms" % timed(lambda: rad.resample(src, method="Gaussian", out=o, tuning=tun), 3))
            print("parity down 8192 tuning", tun, "%.3f ms" % timed(lambda: down.resample(src, method="Gaussian", out=o, tuning=tun), 3))
PY
timeout 900 python -m pytest tests/test_gpu_jacobian.py tests/test_gpu_golden.py tests/test_gpu_headline.py -q -m gpu -x 2>&1 | tail -3
