"""CPU, world_size 2, gloo: the multi-rank plumbing of transform_b200.sharding -- row-block plan,
inverse-mapped bounding boxes with halos, scatter of sub-rectangles, gather of row blocks.
The two device steps are injected: the oracle stands in for the kernels (tests may do that; the
product never does), and every pixel a rank was NOT sent is poisoned with NaN, so a bounding box
that is too small shows up as NaN in the result."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, case, ret):
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_np as o
    import transform_b200 as t
    from transform_b200 import sharding
    from conftest import to_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ny, nx, method, bound, shape = case[:5]
        extra = {"oblur": case[5]} if len(case) > 5 else {}
        src = np.random.default_rng(0).random((ny, nx))
        tr = t.Composition([t.Scale(1.3, dim=2), t.Rotation(25, unit='deg'), t.Offset([-nx / 2.0, -ny / 2.0])])

        def invert_points(transform, pts):
            return to_oracle(transform).reverse(np.asarray(pts, dtype=float))

        def resample_block(transform, sub, origin, full, rows, shp, m, b, kw):
            frame = np.full(full, np.nan)                       # poison what was not staged
            frame[origin[1]:origin[1] + sub.shape[0], origin[0]:origin[0] + sub.shape[1]] = sub
            return o.resample(to_oracle(transform), frame, method=m, bound=b, shape=shp, rows=rows,
                              oblur=kw.get("oblur"))

        out = sharding.resample_sharded(tr, src if rank == 0 else None, method=method, bound=bound, shape=shape,
                                        invert_points=invert_points, resample_block=resample_block,
                                        src_shape=src.shape, src_dtype=src.dtype, **extra)
        if rank == 0:
            want = o.resample(to_oracle(tr), src, method=method, bound=bound, shape=shape, **extra)
            ret["ok"] = bool(out.shape == want.shape and np.array_equal(out, want))
            ret["nan"] = int(np.isnan(out).sum())
            plans = sharding.plan(tr, shape or src.shape, src.shape, method, bound, world, invert_points, **extra)
            ret["boxes"] = [p["box"] for p in plans]
            ret["whole"] = [p["whole"] for p in plans]
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [(96, 128, 'linear', 't', None), (96, 128, 'cubic', 't', (61, 77)),
                                  (64, 64, 'nearest', 't', (5, 9)), (96, 128, 'linear', 'p', None),
                                  # windowed filters: the halo follows the region size (8 and 2*ceil(3*2.5))
                                  (96, 128, 'gaussian', 't', None), (96, 128, 'zlanczos', 't', (61, 77), 2.5)])
def test_sharded_resample_world2(case):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29600 + (os.getpid() % 300)
    mp.spawn(_worker, args=(2, port, case, ret), nprocs=2, join=True)
    assert ret["ok"], dict(ret)
    assert ret["nan"] == 0
    if case[3] == 't':
        # truncating boundary: each rank is staged a proper sub-rectangle, not the frame
        assert not any(ret["whole"]) and all(b is not None for b in ret["boxes"])
        ny, nx = case[0], case[1]
        assert all((b[2] - b[0]) * (b[3] - b[1]) < ny * nx for b in ret["boxes"])
    else:
        assert all(ret["whole"])                                # wrapped taps: whole frame per rank


def test_row_blocks_and_boxes():
    sys.path.insert(0, ROOT)
    from transform_b200 import sharding
    assert sharding.row_blocks(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert sharding.row_blocks(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]          # ragged: empty tails
    pts = sharding.border_points(3, 6, 40)
    assert pts[:, 1].min() == 3 and pts[:, 1].max() == 5 and pts[:, 0].max() == 39
    assert sharding.source_box(np.array([[np.nan, 1.0]]), 10, 10, 1) is None
    assert sharding.source_box(np.array([[-50.0, -50.0]]), 10, 10, 1) is None
    assert sharding.source_box(np.array([[2.2, 3.7], [5.1, 4.0]]), 100, 100, 2, margin=0) == (0, 1, 16, 7)       # columns rounded out to multiples of 8


# --------------------------------------------------------------------------------------------------
# ShardedResampler: cached plan, row / column split, strips gathered while the next one is computed
# --------------------------------------------------------------------------------------------------
def _worker_sr(rank, world, port, case, ret):
    import torch
    import torch.distributed as dist
    for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_np as o
    import transform_b200 as t
    from transform_b200 import sharding
    from conftest import to_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        kind, method, split, nsub, shape = case
        ny, nx = 96, 128
        src = np.random.default_rng(0).random((ny, nx)).astype(np.float32)
        if kind == "polar":
            tr = t.Composition([t.Scale([nx / (2 * np.pi), ny / 70.0]), t.Radial(origin=[63.5, 47.5])])
        else:
            tr = t.Composition([t.Scale(1.3, dim=2), t.Rotation(25, unit='deg'), t.Offset([-nx / 2.0, -ny / 2.0])])
        aa = method[0].isupper()

        def invert_points(transform, pts):
            return to_oracle(transform).reverse(np.asarray(pts, dtype=float))

        def resample_rect(transform, sub, origin, full, rect, shp, m, b, kw, out=None):
            frame = np.full(full, np.nan)                       # poison what was not staged
            frame[origin[1]:origin[1] + sub.shape[0], origin[0]:origin[0] + sub.shape[1]] = sub
            x0, y0, x1, y1 = rect
            if aa:
                blk = o.resample_jacobian(to_oracle(transform), frame, method=m[0].lower(), bound=b, shape=shp,
                                          rows=(y0, y1), max_reg=24)
            else:
                blk = o.resample(to_oracle(transform), frame, method=m, bound=b, shape=shp, rows=(y0, y1))
            return np.ascontiguousarray(blk[:, x0:x1])

        sr = sharding.ShardedResampler(tr, src.shape, shape=shape, method=method, bound='t', nsub=nsub, split=split,
                                       invert_points=invert_points, resample_rect=resample_rect,
                                       **({"max_half": 12} if aa else {}))
        for rep in range(2):                                     # second call: cached plan and buffers
            out = sr(src if rank == 0 else None)
        shp = shape or src.shape
        if aa:
            want = o.resample_jacobian(to_oracle(tr), src, method=method[0].lower(), bound='t', shape=shp, max_reg=24)
        else:
            want = o.resample(to_oracle(tr), src, method=method, bound='t', shape=shp)
        got = out.numpy()
        ret["ok%d" % rank] = bool(got.shape == want.shape and np.array_equal(got, want, equal_nan=False))
        ret["nan%d" % rank] = int(np.isnan(got).sum())
        if rank == 0:
            ret["split"] = sr.split
            ret["staged"] = int(sr.staged_pixels)
            ret["whole"] = [p["whole"] for p in sr.plans]
            ret["bytes"] = (int(sr.bytes_scatter), int(sr.bytes_gather))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", [("affine", "linear", "rows", 1, None), ("affine", "linear", "cols", 3, None),
                                  ("affine", "cubic", "auto", 2, (61, 77)), ("polar", "linear", "auto", 2, None),
                                  ("polar", "Gaussian", "auto", 2, None), ("affine", "Gaussian", "rows", 3, (50, 70))])
def test_sharded_resampler_world2(case):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_worker_sr, args=(2, port, case, ret), nprocs=2, join=True)
    r = dict(ret)
    assert r["ok0"] and r["ok1"], r                               # every rank holds the full, exact result
    assert r["nan0"] == 0 and r["nan1"] == 0, r                   # no tap fell on a pixel that was not staged
    assert r["staged"] < 2 * 96 * 128, r                          # less than two whole frames were sent
    assert r["bytes"][1] > 0
    if case[0] == "polar":
        assert r["split"] == "cols", r                            # wedges beat annuli


@pytest.mark.parametrize("case", [("affine", "linear", "cols", 2, (61, 77)), ("polar", "Gaussian", "auto", 2, None)])
def test_sharded_resampler_world3(case):
    """Three ranks: ragged rectangles (the last rank's part is narrower, 77 and 128 columns do not divide by 3), the
    same exactness and staging checks as with two."""
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29800 + (os.getpid() % 90)
    mp.spawn(_worker_sr, args=(3, port, case, ret), nprocs=3, join=True)
    r = dict(ret)
    assert all(r["ok%d" % k] for k in range(3)), r
    assert all(r["nan%d" % k] == 0 for k in range(3)), r
    assert r["bytes"][0] > 0 and r["bytes"][1] > 0
