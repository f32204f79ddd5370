cd $GRAFT_REPO_ROOT
timeout 300 python tools/host_overhead.py 2>&1 | head -40
timeout 300 python - <<'PY'
import sys, time, torch, numpy as np
sys.path.insert(0,'.')
import transform_b200 as t
n=4096
src = torch.rand((n, n), device="cuda", dtype=torch.float32); dst=torch.empty_like(src)
tr = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
for m in ("linear","nearest"):
    for _ in range(20): tr.resample(src, method=m, precision="fp32", out=dst)
    torch.cuda.synchronize()
    evs=[(torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)) for _ in range(50)]
    t0=time.perf_counter()
    for a,b in evs:
        a.record(); tr.resample(src, method=m, precision="fp32", out=dst); b.record()
    torch.cuda.synchronize()
    wall=(time.perf_counter()-t0)/50*1e6
    ker=sorted(a.elapsed_time(b)*1e3 for a,b in evs)
    print(m, "wall per call %.1f us; event-bracketed per call: median %.1f us min %.1f us" % (wall, ker[25], ker[0]))
PY
