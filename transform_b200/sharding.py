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
import time

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
    # columns on multiples of 8 pixels: the staged rectangle then has a 32-byte row pitch and 32-byte
    # aligned rows, which the texture / 8-byte-load tap paths of the kernels need (else they fall back)
    x0 = (x0 // 8) * 8
    x1 = -(-x1 // 8) * 8
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
    # send / recv take GLOBAL ranks; `rank`, `root` and the plan index are group-local
    g_rank = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)

    src_shape = kw.pop("src_shape", None)
    src_dtype = kw.pop("src_dtype", None)
    dev_in = data is not None and isinstance(data, torch.Tensor) and data.is_cuda
    if data is not None:
        if dev_in:
            src_shape, src_dtype = tuple(data.shape), np.dtype(str(data.dtype).replace("torch.", ""))
        else:
            data = np.asarray(data)
            if data.dtype not in (np.float32, np.float64):
                # integer (FITS) frames travel as float64, exactly as the unsharded path carries them
                data = data.astype(np.float64)
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
                reqs.append(dist.isend(tp.to(dev).contiguous(), dst=g_rank(r), group=group))
    else:
        p = mine
        if p["whole"] or p["box"] is not None:
            if p["whole"]:
                shp, origin = tuple(src_shape), (0, 0)
            else:
                x0, y0, x1, y1 = p["box"]
                shp, origin = tuple(src_shape[:-2]) + (y1 - y0, x1 - x0), (x0, y0)
            buf = torch.empty(shp, dtype=tdt, device=dev)
            dist.recv(buf, src=g_rank(root), group=group)
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
        dist.gather(pad, gathered, dst=g_rank(root), group=group)
    if rank != root:
        return None
    full_out = torch.cat(gathered, dim=-2)[..., :int(shape[-2]), :]
    if dev_in:
        return full_out                                  # device frame in, device frame out
    return full_out.cpu().numpy().astype(out_np, copy=False)


# --------------------------------------------------------------------------------------------------
# Device-resident, plan-cached form: one frame on `root`, output split across the ranks of a group
# --------------------------------------------------------------------------------------------------
def footprint_halo(transform, pts, invert_points, oblur=1.0, pad_pow=8.0, padval=3.0, cap=None):
    """Reach (pixels) of the Jacobian-adaptive footprint over the sample points `pts`: the kernel's
    half-width is ceil(oblur / s'_min), s' = (padval + s^p)^(-1/p) (helpers.pyx:1661-1666), with s the
    singular values of the local Jacobian -- here from +-1-pixel central differences of the
    inverse-mapped sample points.  Non-finite Jacobians are ignored (those pixels produce 0)."""
    pts = np.asarray(pts, dtype=np.float64).reshape(-1, 2)
    if len(pts) == 0:
        return 2
    ex, ey = np.array([1.0, 0.0]), np.array([0.0, 1.0])
    dx = 0.5 * (np.asarray(invert_points(transform, pts + ex)) - np.asarray(invert_points(transform, pts - ex)))
    dy = 0.5 * (np.asarray(invert_points(transform, pts + ey)) - np.asarray(invert_points(transform, pts - ey)))
    g00 = dx[:, 0] ** 2 + dy[:, 0] ** 2
    g11 = dx[:, 1] ** 2 + dy[:, 1] ** 2
    g01 = dx[:, 0] * dx[:, 1] + dy[:, 0] * dy[:, 1]
    lam = 0.5 * (g00 + g11) + np.sqrt((0.5 * (g00 - g11)) ** 2 + g01 ** 2)       # largest eigenvalue of J J^T
    lam = lam[np.isfinite(lam)]
    if len(lam) == 0:
        return 2
    smax = float(np.sqrt(lam.max())) * 1.05                                        # between-sample growth
    half = int(np.ceil(float(oblur) * (padval + smax ** pad_pow) ** (1.0 / pad_pow))) + 1
    if cap is not None:
        half = min(half, int(cap))
    return half + 1


def rect_points(x0, y0, x1, y1, step=16, n=9):
    """Border (every `step` pixels) plus a coarse interior lattice of the grid rectangle [x0,x1) x [y0,y1)."""
    if x1 <= x0 or y1 <= y0:
        return np.zeros((0, 2))
    xs = np.unique(np.concatenate([np.arange(x0, x1, step), [x1 - 1]])).astype(np.float64)
    ys = np.unique(np.concatenate([np.arange(y0, y1, step), [y1 - 1]])).astype(np.float64)
    pts = [np.stack([xs, np.full_like(xs, ys[0])], -1), np.stack([xs, np.full_like(xs, ys[-1])], -1),
           np.stack([np.full_like(ys, xs[0]), ys], -1), np.stack([np.full_like(ys, xs[-1]), ys], -1),
           np.stack(np.meshgrid(np.linspace(x0, x1 - 1, n), np.linspace(y0, y1 - 1, n)), -1).reshape(-1, 2)]
    return np.concatenate(pts, 0)


def plan_rects(transform, shape, src_shape, method, bound, world, invert_points, split="auto", smooth_margin=8,
               oblur=None, aa_kw=None):
    """Per-rank plan for a split of the output grid into `world` rectangles along rows or columns.

    Returns (split, [ {rect:(x0,y0,x1,y1), box:(x0,y0,x1,y1)|None, whole:bool} ... ], staged_pixels).
    'auto' evaluates both splits and keeps the one that stages fewer source pixels: a polar unwrap cut
    by rows needs annuli (whose bounding boxes are whole discs), cut by columns it needs wedges."""
    from . import _engine
    ny, nx = int(shape[-2]), int(shape[-1])
    sh, sw = int(src_shape[-2]), int(src_shape[-1])
    m = method[0]
    aa = m in _engine._AA_METHODS
    bl = bound if isinstance(bound, (list, tuple)) else [bound]
    wrapping = any(b[0] != 't' for b in bl)

    def one(split_axis):
        out, staged = [], 0
        blocks = row_blocks(ny if split_axis == "rows" else nx, world)
        for (a, b) in blocks:
            rect = (0, a, nx, b) if split_axis == "rows" else (a, 0, b, ny)
            entry = {"rect": rect, "box": None, "whole": False}
            if b > a:
                if wrapping:
                    entry["whole"] = True
                else:
                    pts = rect_points(*rect)
                    if aa:
                        kw = aa_kw or {}
                        filt = _engine._AA_METHODS[m]
                        halo = footprint_halo(transform, pts, invert_points, oblur=kw.get("oblur", 1.0),
                                              pad_pow=kw.get("pad_pow", 8), padval=3.0 if filt in "zg" else 1.0,
                                              cap=kw.get("max_half"))
                        margin = smooth_margin + 2          # the Jacobian stencil reads coordinates, not data
                    else:
                        halo, margin = halo_of(method, oblur), smooth_margin
                    box = source_box(invert_points(transform, pts), sh, sw, halo, margin=margin)
                    if box is not None and (box[2] - box[0]) * (box[3] - box[1]) > 0.9 * sh * sw:
                        entry["whole"] = True
                    else:
                        entry["box"] = box
                if entry["whole"]:
                    staged += sh * sw
                elif entry["box"] is not None:
                    staged += (entry["box"][2] - entry["box"][0]) * (entry["box"][3] - entry["box"][1])
            out.append(entry)
        return out, staged

    if split in ("rows", "cols"):
        pl, staged = one(split)
        return split, pl, staged
    pr, sr = one("rows")
    pc, sc = one("cols")
    return ("rows", pr, sr) if sr <= sc else ("cols", pc, sc)


def _default_resample_rect(transform, sub, origin, full, rect, shape, method, bound, kw, out=None):
    """Resample the output rectangle `rect` = (x0, y0, x1, y1) of the grid `shape` against the staged
    sub-rectangle `sub` (global origin `origin`, full source extent `full`) with K1 / K2."""
    from . import _engine
    x0, y0, x1, y1 = rect
    blk_shape = (y1 - y0, x1 - x0)
    ops = transform._int_chain(True)
    m = method[0]
    whole = (tuple(origin) == (0, 0) and tuple(full) == tuple(sub.shape[-2:]))
    if m in _engine._AA_METHODS:
        return _engine.resample_jacobian(ops, sub, blk_shape, _engine._AA_METHODS[m], bound, dst_origin=(x0, y0),
                                         grid_full=(int(shape[-2]), int(shape[-1])), src_origin=origin,
                                         src_full=None if whole else tuple(full), out=out, **kw)
    return _engine.resample(ops, sub, blk_shape, m, bound, dst_origin=(x0, y0), src_origin=origin,
                            src_full=None if whole else tuple(full), out=out, **kw)


class ShardedResampler:
    """`Transform.resample` of ONE frame with the output grid split across the ranks of a process group.

    The frame lives on `root` (a CUDA tensor with NCCL, a CPU tensor / array with gloo).  Per call:
      scatter   root cuts each rank's source rectangle on the device and sends it (grouped P2P);
      compute   every rank resamples its output rectangle against what it was sent (`nsub` sub-blocks);
      gather    all_gather of the sub-blocks, issued asynchronously as each one is finished, so the
                gather of sub-block k overlaps the kernel of sub-block k+1.
    The plan (split axis, rectangles, source boxes with filter / footprint halos) is computed once in
    the constructor.  Every rank returns the full result.  Attributes after a call: `bytes_scatter`,
    `bytes_gather` (what crossed the links, all ranks), `bytes_root_out`, `bytes_rank_in` (the busiest
    GPU's egress / ingress: what the link bandwidth bounds)."""

    def __init__(self, transform, src_shape, shape=None, method='n', bound='t', *, group=None, root=0, nsub=1,
                 split="auto", src_dtype=np.float32, invert_points=None, resample_rect=None, **kw):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.transform, self.method, self.bound, self.kw = transform, method, bound, dict(kw)
        self.group, self.root = group, int(root)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        self.dev = (torch.device("cuda", torch.cuda.current_device()) if self.backend == "nccl"
                    else torch.device("cpu"))
        self.g_rank = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
        self.src_shape = tuple(int(v) for v in src_shape)
        self.shape = tuple(int(v) for v in (shape if shape is not None else src_shape))
        if len(self.shape) != 2 or len(self.src_shape) != 2:
            raise NotImplementedError("ShardedResampler: one 2-D frame per call (batches shard by frame)")
        self.src_tdt = torch.from_numpy(np.zeros(0, dtype=src_dtype)).dtype
        self.invert_points = invert_points or _default_invert
        self.resample_rect = resample_rect or _default_resample_rect
        self.nsub = max(1, int(nsub))
        aa_kw = {k: self.kw[k] for k in ("oblur", "pad_pow") if k in self.kw}
        self.split, self.plans, self.staged_pixels = plan_rects(
            transform, self.shape, self.src_shape, method, bound, self.world, self.invert_points, split=split,
            oblur=self.kw.get("oblur"), aa_kw=aa_kw)
        self._recv = None
        self._out = None
        self._out_t = None
        self._slabs = self._gath = None
        self.profile = False                # True: synchronise between phases and record phase_ms
        self.phase_ms = {}
        self.bytes_scatter = self.bytes_gather = self.bytes_root_out = self.bytes_rank_in = 0

    # -- what rank r needs of the source: (view-slices, origin) --
    def _piece_geom(self, p):
        if p["whole"]:
            return (0, 0, self.src_shape[1], self.src_shape[0])
        return p["box"]

    def _sub_rects(self, rect):
        """`nsub` strips of a rank's rectangle, along the axis that is NOT the split axis' complement:
        rows of the rectangle (strips stay contiguous in the gathered output either way)."""
        x0, y0, x1, y1 = rect
        return [(x0, a + y0, x1, b + y0) for (a, b) in row_blocks(y1 - y0, self.nsub) if b > a]

    def __call__(self, frame=None, sync=True):
        torch, dist = self.torch, self.dist
        rank, world, root = self.rank, self.world, self.root
        full = self.src_shape
        mine = self.plans[rank]
        es = torch.empty(0, dtype=self.src_tdt).element_size()
        # ---- scatter ----
        ops, sub, origin = [], None, (0, 0)
        self.bytes_scatter = 0
        if rank == root:
            if frame is None:
                raise ValueError("ShardedResampler: the root rank must pass the frame")
            if not torch.is_tensor(frame):
                frame = torch.from_numpy(np.ascontiguousarray(frame))
            frame = frame.to(device=self.dev, dtype=self.src_tdt)
            for r, p in enumerate(self.plans):
                g = self._piece_geom(p)
                if g is None or p["rect"][2] <= p["rect"][0] or p["rect"][3] <= p["rect"][1]:
                    continue
                x0, y0, x1, y1 = g
                piece = frame if (x0, y0, x1, y1) == (0, 0, full[1], full[0]) else frame[y0:y1, x0:x1].contiguous()
                if r == root:
                    sub, origin = piece, (x0, y0)
                else:
                    ops.append(dist.P2POp(dist.isend, piece, self.g_rank(r), group=self.group))
                    self.bytes_scatter += piece.numel() * es
        else:
            g = self._piece_geom(mine)
            if g is not None and mine["rect"][2] > mine["rect"][0] and mine["rect"][3] > mine["rect"][1]:
                x0, y0, x1, y1 = g
                if self._recv is None or tuple(self._recv.shape) != (y1 - y0, x1 - x0):
                    self._recv = torch.empty((y1 - y0, x1 - x0), dtype=self.src_tdt, device=self.dev)
                sub, origin = self._recv, (x0, y0)
                ops.append(dist.P2POp(dist.irecv, sub, self.g_rank(root), group=self.group))
        if self.profile:
            torch.cuda.synchronize(); dist.barrier(group=self.group); _t0 = time.perf_counter()
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        if self.profile:
            torch.cuda.synchronize(); dist.barrier(group=self.group); _t1 = time.perf_counter()
            self.phase_ms["scatter"] = (_t1 - _t0) * 1e3
        # ---- compute + gather, strip by strip: the all_gather of strip k is in flight while strip k+1
        # is being resampled (NCCL runs it on its own stream behind an event of the compute stream) ----
        x0, y0, x1, y1 = mine["rect"]
        ny, nx = self.shape
        rows_split = self.split == "rows"
        per = -(-(ny if rows_split else nx) // world)
        slab_h, slab_w = (per, nx) if rows_split else (ny, per)
        hs = -(-slab_h // self.nsub)
        strips = [(k * hs, min((k + 1) * hs, slab_h)) for k in range(self.nsub) if k * hs < slab_h]
        fill = float(self.kw.get("fillvalue", 0))
        works, gathered = [], []
        if self._out_t is None:
            # first call: agree on the result dtype.  Decided by the compute step (precision / method rules
            # of the engine), so one small block is resampled and its dtype reduced over the ranks.
            code = -1
            if sub is not None and y1 > y0 and x1 > x0:
                probe = self.resample_rect(self.transform, sub if self.backend == "nccl" else sub.numpy(), origin, full,
                                           (x0, y0, min(x0 + 8, x1), min(y0 + 8, y1)), self.shape, self.method,
                                           self.bound, self.kw)
                pdt = probe.dtype if torch.is_tensor(probe) else torch.from_numpy(np.zeros(0, probe.dtype)).dtype
                code = 0 if pdt == torch.float32 else 1
            codet = torch.tensor([code], dtype=torch.int64, device=self.dev)
            dist.all_reduce(codet, op=dist.ReduceOp.MAX, group=self.group)
            self._out_t = torch.float32 if int(codet[0]) <= 0 else torch.float64
        if self._out is None:
            # persistent buffers: the full result, one slab and one gather target per strip
            self._out = torch.empty((world * per, nx) if rows_split else (ny, world * per), dtype=self._out_t,
                                    device=self.dev)
            self._slabs = [torch.empty((b - a, slab_w), dtype=self._out_t, device=self.dev) for (a, b) in strips]
            direct = rows_split and len(strips) == 1           # rank blocks are contiguous in the result
            self._gath = [self._out if direct else
                          torch.empty((world * (b - a), slab_w), dtype=self._out_t, device=self.dev) for (a, b) in strips]
        out = self._out
        for k, (a, b) in enumerate(strips):
            # the part of this strip that lies inside the rank's real rectangle (ragged tails are fill)
            ry0, ry1 = y0 + a, min(y0 + b, y1)
            rect = (x0, ry0, x1, ry1)
            slab = self._slabs[k]
            whole_strip = (ry1 - ry0 == b - a) and (x1 - x0 == slab_w)
            if not whole_strip or sub is None:
                slab.fill_(fill)
            if sub is not None and ry1 > ry0 and x1 > x0:
                dst = slab if (whole_strip and self.backend == "nccl") else None
                blk = self.resample_rect(self.transform, sub if self.backend == "nccl" else sub.numpy(), origin,
                                         full, rect, self.shape, self.method, self.bound, self.kw, out=dst)
                if dst is None:
                    blk = blk if torch.is_tensor(blk) else torch.from_numpy(np.ascontiguousarray(blk))
                    slab[:ry1 - ry0, :x1 - x0] = blk.to(device=self.dev, dtype=self._out_t)
            g = self._gath[k]
            works.append(dist.all_gather_into_tensor(g, slab, group=self.group, async_op=True))
            gathered.append(((a, b), g, slab))
        if self.profile:
            torch.cuda.synchronize(); dist.barrier(group=self.group); _t2 = time.perf_counter()
            self.phase_ms["compute+gather"] = (_t2 - _t1) * 1e3
        for w, ((a, b), g, _) in zip(works, gathered):
            w.wait()
            if g is out:
                continue
            g3 = g.view(world, b - a, slab_w)
            if rows_split:
                out.view(world, per, nx)[:, a:b, :] = g3
            else:
                out.view(ny, world, per)[a:b] = g3.permute(1, 0, 2)
        eso = out.element_size()
        self.bytes_gather = world * (world - 1) * slab_h * slab_w * eso
        self.bytes_root_out = self.bytes_scatter + (world - 1) * slab_h * slab_w * eso
        self.bytes_rank_in = (world - 1) * slab_h * slab_w * eso
        out = out[:ny] if rows_split else out[:, :nx]
        if self.profile:
            torch.cuda.synchronize(); _t3 = time.perf_counter()
            self.phase_ms["assemble"] = (_t3 - _t2) * 1e3
        if sync and self.backend == "nccl":
            torch.cuda.current_stream().synchronize()
        return out
