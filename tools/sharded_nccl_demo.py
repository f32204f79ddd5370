"""torchrun --nproc-per-node N tools/sharded_nccl_demo.py : one 8192^2 frame, output rows split across
N GPUs over NCCL (scatter of inverse-mapped sub-rectangles, all_gather of row blocks); checks the
result against the unsharded call and prints timings (end to end, including the transfers)."""
import os
import sys
import time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import transform_b200 as t
from transform_b200 import sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 8192
src = np.random.default_rng(0).random((n, n), dtype=np.float32) if rank == 0 else None
tr = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
for method in ("linear", "Gaussian"):
    for rep in range(2):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        out = sharding.resample_sharded(tr, src, method=method, bound="t", precision="fp32",
                                        src_shape=(n, n), src_dtype=np.float32)
        dist.barrier(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if rank == 0:
        want = tr.resample(src, method=method, precision="fp32")
        print("%s world=%d sharded (host frame in/out) %.1f ms equal=%s" % (method, world, dt * 1e3, np.array_equal(out, want)))
    # resident frame: device tensor in, device tensor out
    dsrc = torch.from_numpy(src).cuda() if rank == 0 else None
    for rep in range(3):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        dout = sharding.resample_sharded(tr, dsrc, method=method, bound="t", precision="fp32",
                                         src_shape=(n, n), src_dtype=np.float32)
        dist.barrier(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    if rank == 0:
        print("%s world=%d sharded (resident frame) %.1f ms equal=%s" % (method, world, dt * 1e3,
                                                                        np.array_equal(dout.cpu().numpy(), want)))
dist.destroy_process_group()
