"""torchrun --nproc-per-node N tools/sharded_nccl_demo.py : one 8192^2 frame resident on rank 0, the output
grid split across N GPUs (transform_b200.sharding.ShardedResampler over NCCL); checks the result against
the unsharded call and prints where the time goes (scatter / compute / gather / assemble)."""
import os
import sys
import time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import transform_b200 as t
from transform_b200 import sharding, _lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 8192
g = torch.Generator(device="cuda").manual_seed(1000)
src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32) if rank == 0 else None
c = (n - 1) / 2.0
cases = {"linear/affine": (t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])]), "linear"),
         "linear/shift": (t.Offset([3.25, -7.5]), "linear"),
         "Gaussian/polar": (t.Composition([t.Scale([n / (2 * np.pi), np.sqrt(2.0)]), t.Radial(origin=[c, c])]), "Gaussian")}
for name, (tr, method) in cases.items():
    for split in ("rows", "cols"):
        for nsub in (1, 4):
            sr = sharding.ShardedResampler(tr, (n, n), method=method, bound="t", nsub=nsub, split=split, precision="fp32")
            for _ in range(3):
                out = sr(src)
            dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
            reps = 10
            for _ in range(reps):
                out = sr(src, sync=False)
            torch.cuda.synchronize(); dist.barrier(); dt = (time.perf_counter() - t0) / reps
            if rank == 0:
                want = tr.resample(src, method=method, precision="fp32")
                kern = _lib.last_kernel()
                d = (out - want).abs()
                bad = int((d > 0).sum())
                msg = ""
                if bad:
                    ys, xs = torch.nonzero(d > 0, as_tuple=True)
                    msg = " max %.3e rows %d..%d cols %d..%d" % (float(d.max()), int(ys.min()), int(ys.max()), int(xs.min()), int(xs.max()))
                print("%-16s split=%-4s(%s) nsub=%d world=%d: %.3f ms  staged %.2f frames  differing pixels %d%s  [single-GPU kernel %s]"
                      % (name, split, sr.split, nsub, world, dt * 1e3, sr.staged_pixels / float(n * n), bad, msg, kern), flush=True)
            del sr
# phase timing (synchronising between phases: not overlapped, shows the parts)
for name, (tr, method) in cases.items():
    sr = sharding.ShardedResampler(tr, (n, n), method=method, bound="t", nsub=1, precision="fp32")
    sr.profile = True
    for _ in range(3):
        sr(src)
    if rank == 0:
        print(name, "split", sr.split, "phases (ms, synchronised):", {k: round(v, 3) for k, v in sr.phase_ms.items()}, flush=True)
    del sr
dist.destroy_process_group()
