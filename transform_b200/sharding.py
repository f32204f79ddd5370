"""Multi-GPU sharding of one resample call: output row blocks across ranks.

One process per GPU (`torch.distributed`, NCCL on the GPU box, gloo in the CPU tests).  The path
shards naturally -- every output pixel is independent -- so the only communication is plumbing:
  1. every rank pushes the border of its output row block through the inverse chain (on its GPU,
     through the K3 kernel) and takes the bounding box of the result, inflated by the filter halo;
  2. rank 0, which owns the input frame, sends each rank its sub-rectangle (or the whole frame when
     the box covers most of it, or a boundary mode wraps/clamps in global coordinates);
  3. each rank resamples its block against the staged sub-rectangle (global origins are passed to
     the kernel, which also verifies that no tap falls outside the staged data);
  4. the row blocks are gathered on rank 0.
No collective sits inside the data path of a kernel; for batches of frames (BASELINE config 4) use
whole frames per rank and no communication at all.

The two device-side steps are injectable (`invert_points`, `resample_block`) so that the plan and
the scatter/gather plumbing are testable with gloo on CPU-only machines.
"""
import numpy as np

HALO = {"n": 1, "l": 2, "c": 3}          # taps reach floor-1 .. floor+2 (cubic); +1 of slack


def halo_of(method, oblur=None):
    """Reach of the interpolation footprint around floor(v), in pixels (None: whole frame needed)."""
    from . import _engine
    m = method[0]
    if m in _engine._FILTERS and (m != "l" or oblur is not None):
        return _engine.filter_size(m, oblur) // 2 + 1
    return HALO.get(m)


def row_blocks(height, world):
    """Contiguous row blocks [y0, y1) per rank, equal size except the tail (possibly empty)."""
    per = -(-int(height) // int(world))
    return [(min(r * per, height), min((r + 1) * per, height)) for r in range(world)]


def border_points(y0, y1, width, step=16):
    """Grid points on the border of rows [y0,y1) x [0,width), sampled every `step` pixels (corners
    always included).  An affine map sends the block to a parallelogram, so its border bounds the
    image of the block exactly; for smooth non-affine maps `step` trades cost for safety."""
    if y1 <= y0 or width <= 0:
        return np.zeros((0, 2))
    xs = np.unique(np.concatenate([np.arange(0, width, step), [width - 1]])).astype(np.float64)
    ys = np.unique(np.concatenate([np.arange(y0, y1, step), [y1 - 1]])).astype(np.float64)
    pts = [np.stack([xs, np.full_like(xs, ys[0])], -1), np.stack([xs, np.full_like(xs, ys[-1])], -1),
           np.stack([np.full_like(ys, xs[0]), ys], -1), np.stack([np.full_like(ys, xs[-1]), ys], -1)]
    return np.concatenate(pts, 0)


def interior_points(y0, y1, width, n=9):
    """A coarse n x n lattice inside the block: catches maps whose extrema are not on the border."""
    if y1 <= y0 or width <= 0:
        return np.zeros((0, 2))
    xs = np.linspace(0, width - 1, n)
    ys = np.linspace(y0, y1 - 1, n)
    g = np.stack(np.meshgrid(xs, ys), -1).reshape(-1, 2)
    return g


def source_box(coords, src_h, src_w, halo, margin=2):
    """Bounding box (x0, y0, x1, y1), clipped to the image, of the inverse-mapped sample points,
    inflated by the filter halo plus a safety margin.  Returns None when nothing maps inside.
    Non-finite coordinates are ignored (they produce the fill value whatever is staged)."""
    c = np.asarray(coords, dtype=np.float64).reshape(-1, 2)
    ok = np.isfinite(c).all(axis=1)
    if not ok.any():
        return None
    c = c[ok]
    x0 = int(np.floor(c[:, 0].min())) - halo - margin
    x1 = int(np.ceil(c[:, 0].max())) + halo + margin + 1
    y0 = int(np.floor(c[:, 1].min())) - halo - margin
    y1 = int(np.ceil(c[:, 1].max())) + halo + margin + 1
    x0, y0 = max(x0, 0), max(y0, 0)
    x1, y1 = min(x1, src_w), min(y1, src_h)
    if x1 <= x0 or y1 <= y0:
        return None
    return (x0, y0, x1, y1)


def plan(transform, shape, src_shape, method, bound, world, invert_points, smooth_margin=8, oblur=None):
    """Per-rank plan: list of dicts {rows, box | None, whole}.  `whole` = stage the entire frame
    (non-truncating boundary modes act in global coordinates; or the box is most of the frame)."""
    ny, nx = int(shape[-2]), int(shape[-1])
    sh, sw = int(src_shape[-2]), int(src_shape[-1])
    halo = halo_of(method, oblur)
    bl = bound if isinstance(bound, (list, tuple)) else [bound]
    wrapping = any(b[0] != 't' for b in bl)
    out = []
    for (y0, y1) in row_blocks(ny, world):
        entry = {"rows": (y0, y1), "box": None, "whole": False}
        if y1 > y0:
            if halo is None or wrapping:
                entry["whole"] = True                    # Jacobian footprints / wrapped taps: whole frame
            else:
                pts = np.concatenate([border_points(y0, y1, nx), interior_points(y0, y1, nx)], 0)
                box = source_box(invert_points(transform, pts), sh, sw, halo, margin=smooth_margin)
                if box is not None and (box[2] - box[0]) * (box[3] - box[1]) > 0.8 * sh * sw:
                    entry["whole"] = True
                else:
                    entry["box"] = box
        out.append(entry)
    return out


def _default_invert(transform, pts):
    return np.asarray(transform.invert(pts))


def _default_resample(transform, sub, origin, full, rows, shape, method, bound, kw):
    """Resample output rows `rows` against the staged sub-rectangle `sub` whose [0][0] sits at
    global (origin_x, origin_y) of a frame of extent `full`.  Runs K1 / K2 through the engine."""
    from . import _engine
    y0, y1 = rows
    blk_shape = (y1 - y0, int(shape[-1]))
    ops = transform._int_chain(True)
    m = method[0]
    whole = (origin == (0, 0) and tuple(full) == tuple(sub.shape[-2:]))
    if m in _engine._AA_METHODS:
        return _engine.resample_jacobian(ops, sub, blk_shape, _engine._AA_METHODS[m], bound,
                                         dst_origin=(0, y0), grid_full=(int(shape[-2]), int(shape[-1])),
                                         src_origin=origin, src_full=None if whole else tuple(full), **kw)
    return _engine.resample(ops, sub, blk_shape, m, bound, dst_origin=(0, y0), src_origin=origin,
                            src_full=None if whole else tuple(full), **kw)


def resample_sharded(transform, data, method='n', bound='t', shape=None, *, group=None, root=0,
                     invert_points=None, resample_block=None, **kw):
    """Transform.resample with the output rows split across the ranks of `group`.

    `data` (NumPy array or CUDA tensor) is needed on `root` only (other ranks may pass None; they
    must pass `src_shape=` and `src_dtype=` through kw in that case).  Returns the full result on
    `root` (a CUDA tensor when the input was one), None elsewhere."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    invert_points = invert_points or _default_invert
    resample_block = resample_block or _default_resample

    src_shape = kw.pop("src_shape", None)
    src_dtype = kw.pop("src_dtype", None)
    dev_in = data is not None and isinstance(data, torch.Tensor) and data.is_cuda
    if data is not None:
        if dev_in:
            src_shape, src_dtype = tuple(data.shape), np.dtype(str(data.dtype).replace("torch.", ""))
        else:
            data = np.asarray(data)
            src_shape, src_dtype = data.shape, data.dtype
            if backend == "nccl":
                # one upload of the frame; the per-rank sub-rectangles are cut on the device and go out
                # over NVLink (slicing 8192^2 on the host cost ~0.2 s per call)
                data = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
    if src_shape is None:
        raise ValueError("ranks without the data must be told src_shape= and src_dtype=")
    if shape is None:
        shape = src_shape
    full = (int(src_shape[-2]), int(src_shape[-1]))
    # every rank derives the SAME plan (deterministic host logic + identical kernels)
    plans = plan(transform, shape, src_shape, method, bound, world, invert_points, oblur=kw.get("oblur"))
    mine = plans[rank]

    # ---- scatter: root -> rank r : its sub-rectangle ----
    def cut(p):
        if p["whole"]:
            return data, (0, 0)
        if p["box"] is None:
            return None, (0, 0)
        x0, y0, x1, y1 = p["box"]
        if isinstance(data, torch.Tensor):
            return data[..., y0:y1, x0:x1].contiguous(), (x0, y0)
        return np.ascontiguousarray(data[..., y0:y1, x0:x1]), (x0, y0)

    tdt = torch.from_numpy(np.zeros(0, dtype=src_dtype)).dtype
    sub, origin = None, (0, 0)
    reqs = []
    if rank == root:
        for r, p in enumerate(plans):
            piece, org = cut(p)
            if r == root:
                sub, origin = piece, org
            elif piece is not None:
                tp = piece if isinstance(piece, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(piece))
                reqs.append(dist.isend(tp.to(dev).contiguous(), dst=r, group=group))
    else:
        p = mine
        if p["whole"] or p["box"] is not None:
            if p["whole"]:
                shp, origin = tuple(src_shape), (0, 0)
            else:
                x0, y0, x1, y1 = p["box"]
                shp, origin = tuple(src_shape[:-2]) + (y1 - y0, x1 - x0), (x0, y0)
            buf = torch.empty(shp, dtype=tdt, device=dev)
            dist.recv(buf, src=root, group=group)
            sub = buf.cpu().numpy() if backend != "nccl" else buf
    for q in reqs:
        q.wait()

    # ---- compute this rank's row block ----
    y0, y1 = mine["rows"]
    lead = tuple(int(s) for s in shape[:-2])
    if y1 > y0 and sub is not None:
        blk = resample_block(transform, sub, origin, full, (y0, y1), shape, method, bound, kw)
    elif y1 > y0:
        blk = None                                      # nothing maps inside: the block is all fill
    else:
        blk = None

    # ---- gather the row blocks on root ----
    per = -(-int(shape[-2]) // world)
    # the result dtype is decided by the compute step; agree on it (0 = f32, 1 = f64)
    code = torch.tensor([-1], dtype=torch.int64, device=dev)
    if blk is not None:
        bdt = blk.dtype if isinstance(blk, np.ndarray) else np.dtype(str(blk.dtype).replace("torch.", ""))
        code[0] = 0 if bdt == np.float32 else 1
    dist.all_reduce(code, op=dist.ReduceOp.MAX, group=group)
    out_np = np.float32 if int(code[0]) == 0 else np.float64
    out_t = torch.float32 if int(code[0]) == 0 else torch.float64
    pad = torch.zeros(lead + (per, int(shape[-1])), dtype=out_t, device=dev)
    fill = float(kw.get("fillvalue", 0))
    if y1 > y0:
        if blk is None:
            pad[..., :y1 - y0, :] = fill
        else:
            tb = blk if not isinstance(blk, np.ndarray) else torch.from_numpy(np.ascontiguousarray(blk))
            pad[..., :y1 - y0, :] = tb.to(device=dev, dtype=out_t)
    gathered = [torch.empty_like(pad) for _ in range(world)] if rank == root else None
    if backend == "nccl":
        # NCCL has no gather-to-root primitive exposed uniformly; all_gather of equal blocks is one
        # NVSwitch collective and keeps every rank's buffers symmetric
        gathered = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(gathered, pad, group=group)
    else:
        dist.gather(pad, gathered, dst=root, group=group)
    if rank != root:
        return None
    full_out = torch.cat(gathered, dim=-2)[..., :int(shape[-2]), :]
    if dev_in:
        return full_out                                  # device frame in, device frame out
    return full_out.cpu().numpy().astype(out_np, copy=False)
