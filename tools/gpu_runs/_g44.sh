cd $GRAFT_REPO_ROOT
for v in "" transform_b200/lib/variants/libtfm_b200_tex4_mb5.so transform_b200/lib/variants/libtfm_b200_tex4_mb6.so; do
echo "=== lib ${v:-default}"
TFM_B200_LIB=${v:+$PWD/$v} timeout 300 python - <<'PY'
import sys, torch, numpy as np
sys.path.insert(0,'.')
import transform_b200 as t
from transform_b200 import _lib
def timed(fn, reps=20):
    fn(); fn(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b)/reps
n=8192
src = torch.rand((n, n), device="cuda", dtype=torch.float32); dst=torch.empty_like(src)
c=(n-1)/2.0
rad = t.Composition([t.Scale([n / (2 * np.pi), n / (n / np.sqrt(2.0))]), t.Radial(origin=[c, c])])
pin = t.Composition([t.Quadpin(idim=2, odim=2, strength=0.2, length=n/2.0, origin=[c, c]), t.Scale([0.8, 0.7])])
for name, tr in (("polar", rad), ("pincushion", pin)):
    ms = timed(lambda: tr.resample(src, method="linear", precision="fp32", out=dst))
    print(name, _lib.last_kernel(), "%.4f ms  %.1f %%" % (ms, 100*2*n*n*4/(ms*1e-3)/1e9/6547.5))
PY
done
