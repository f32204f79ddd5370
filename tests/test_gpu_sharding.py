"""GPU: transform_b200.sharding with the REAL kernels.  Two ranks (gloo for the plumbing) share the
one GPU of the test box: each computes its inverse-mapped bounding box with K3, receives only its
sub-rectangle, resamples its row block with K1/K2 against the staged rectangle (global origins, tap
containment checked by the kernel's status word), and rank 0 gathers.  Result must equal the
unsharded call bit for bit."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, case, ret):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import transform_b200 as t
    from transform_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        method, bound, kw = case
        ny, nx = 160, 224
        src = np.random.default_rng(0).random((ny, nx)).astype(np.float32)
        if method == 'G':
            tr = t.Composition([t.Scale([nx / (2 * np.pi), ny / 140.0]), t.Radial(origin=[111.5, 79.5])])
        else:
            tr = t.Composition([t.Scale(1.3, dim=2), t.Rotation(25, unit='deg'), t.Offset([-nx / 2.0, -ny / 2.0])])
        out = sharding.resample_sharded(tr, src if rank == 0 else None, method=method, bound=bound,
                                        src_shape=src.shape, src_dtype=src.dtype, **kw)
        if rank == 0:
            want = tr.resample(src, method=method, bound=bound, **kw)
            ret["ok"] = bool(out.shape == want.shape and out.dtype == want.dtype and np.array_equal(out, want))
            ret["maxdiff"] = float(np.abs(out.astype(np.float64) - want).max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [("linear", "t", {}), ("cubic", "t", {}), ("nearest", "e", {}),
                                  ("linear", "t", {"precision": "fp32"}), ("G", "t", {}),
                                  ("G", "t", {"precision": "fp32"}), ("hann", "t", {}),
                                  ("gaussian", "t", {"oblur": 1.5}), ("zlanczos", "t", {"precision": "fp32"})])
def test_sharded_equals_unsharded(case):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29700 + (os.getpid() % 250)
    mp.spawn(_worker, args=(2, port, case, ret), nprocs=2, join=True)
    assert ret["ok"], dict(ret)
