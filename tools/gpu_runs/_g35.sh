cd $GRAFT_REPO_ROOT
timeout 600 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0,'.')
import transform_b200 as t
from transform_b200 import _lib
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
n=8192
g = torch.Generator(device="cuda").manual_seed(7)
src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
dst = torch.empty_like(src)
c=(n-1)/2.0
rad = t.Composition([t.Scale([n / (2 * np.pi), n / (n / np.sqrt(2.0))]), t.Radial(origin=[c, c])])
pin = t.Composition([t.Quadpin(idim=2, odim=2, strength=0.2, length=n/2.0, origin=[c, c]), t.Scale([0.8, 0.7])])
for name, tr in (("polar", rad), ("pincushion", pin)):
    for tun in (512, 512|4096):
        ms = timed(lambda: tr.resample(src, method="Gaussian", precision="fp32", out=dst, tuning=tun))
        print(name, tun, _lib.last_kernel(), "%.3f ms" % ms)
    a = tr.resample(src, method="Gaussian", precision="fp32", tuning=512)
    b = tr.resample(src, method="Gaussian", precision="fp32", tuning=512|4096)
    d=(a-b).abs(); print("  tex vs ldg: max %.3g n>1e-5 %d" % (float(d.max()), int((d>1e-5).sum())))
    for bound in ("e","m"):
        a = tr.resample(src, method="Gaussian", precision="fp32", tuning=512, bound=bound); ka=_lib.last_kernel()
        b = tr.resample(src, method="Gaussian", precision="fp32", tuning=512|4096, bound=bound)
        d=(a-b).abs(); print("  bound", bound, ka, "max %.3g" % float(d.max()))
PY
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_jacobian.py tests/test_gpu_fuzz.py tests/test_gpu_golden.py tests/test_gpu_sharding.py tests/test_gpu_remap.py -q -m gpu -x 2>&1 | tail -4
