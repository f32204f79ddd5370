#!/usr/bin/env python3
"""Host <-> device copy ceiling of the box, all ranks at once (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/host_ceiling.py

Each rank copies the bench's per-step traffic (one 268 MB frame up, two down) between pinned host buffers and
its GPU with NO kernels in between: uploads only, downloads only, and both directions at once on two streams.
The aggregate over ranks is the most the end-to-end number of bench.py could reach on this box
(e2e Mpix/s <= 134.2 Mpix per step x steps/s), with and without NUMA-local host buffers."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    numa = None
    if not os.environ.get("TFM_BENCH_NO_NUMA"):
        from transform_b200 import _device as tdev
        numa = tdev.bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 8192
    h_src = torch.empty((n, n), dtype=torch.float32).pin_memory()
    h_out = [torch.empty((n, n), dtype=torch.float32).pin_memory() for _ in range(2)]
    h_src.zero_()
    for b in h_out:
        b.zero_()
    d_src = torch.empty((n, n), device="cuda", dtype=torch.float32)
    d_out = [torch.empty((n, n), device="cuda", dtype=torch.float32) for _ in range(2)]
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(up, down, reps=20):
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s_up):
                    d_src.copy_(h_src, non_blocking=True)
            if down:
                with torch.cuda.stream(s_dn):
                    h_out[0].copy_(d_out[0], non_blocking=True)
                    h_out[1].copy_(d_out[1], non_blocking=True)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
        nbytes = reps * ((n * n * 4 if up else 0) + (2 * n * n * 4 if down else 0)) * world
        return nbytes / dt / 1e9, reps * world / dt

    run(True, True, 3)
    res = {"world": world, "numa_local_buffers": numa is not None}
    res["h2d_only_GBs"], _ = run(True, False)
    res["d2h_only_GBs"], _ = run(False, True)
    both, steps = run(True, True)
    res["both_GBs"] = both
    res["bench_steps_per_s_ceiling"] = steps
    res["e2e_ceiling_mpix_s"] = steps * 2 * n * n / 1e6
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
