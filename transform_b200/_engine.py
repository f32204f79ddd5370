"""Host-side driver for libtfm_b200.so: argument marshalling, dtype rules, error mapping.

This module is the only place that calls the C ABI.  It restates *no* arithmetic of the path --
only the reference's bookkeeping rules that decide shapes and dtypes:
  * method / boundary are parsed by first character (core.py:557, helpers.pyx:137, 529-533)
  * per-axis boundary lists are in (X,Y) order (helpers.pyx:113-120)
  * nearest keeps the source dtype unless the truncation pass of sampleND promotes it; linear
    and cubic return float64 (helpers.pyx:373-374, 598-611, 770-778)
"""
import ctypes

import numpy as np

from . import _device as dev
from . import _lib as L

_METHODS = {"n": L.NEAREST, "l": L.LINEAR, "c": L.CUBIC}
# capital-letter methods of Transform.resample (core.py:495-508) -> interpND_grid filter letters
_AA_METHODS = {"G": "g", "H": "h", "R": "t", "Z": "z", "L": "l", "W": "w"}
# the collapsible input-plane filters of interpND (helpers.pyx:703-735, 792-835) and their region size
_FILTERS = {"s": (L.SINC, 16), "z": (L.LANCZOS, 6), "g": (L.GAUSSIAN, 8), "h": (L.HANN, 2),
            "t": (L.TUKEY, 2), "l": (L.TENT, 2)}
_UNSUPPORTED_INTERP = ("f",)


def filter_size(m, oblur):
    """Edge of the sampled region (helpers.pyx:708-711)."""
    size = _FILTERS[m][1]
    if oblur is not None:
        size = int(2 * np.ceil(size / 2 * oblur))
    return size


def parse_precision(precision):
    if precision in (None, "fp64", "f64", "exact", "exact64", L.PREC_EXACT64):
        return L.PREC_EXACT64
    if precision in ("fp32", "f32", "fast", "fast32", L.PREC_FAST32):
        return L.PREC_FAST32
    raise ValueError("precision must be 'fp64' (parity mode) or 'fp32' (fast path)")


def bound_pair(bound, ndim=2):
    """-> (code_x, code_y); the reference's parsing rules (helpers.pyx:113-120, 137, 174-176)."""
    if not isinstance(bound, (tuple, list)):
        bl = [bound] * ndim
    elif len(bound) == 1:
        bl = [bound[0]] * ndim
    else:
        bl = list(bound)
    if len(bl) != ndim:
        raise ValueError("apply_boundary: boundary list must match vec dims")
    try:
        return tuple(L.BOUND_CODE[b[0]] for b in bl)
    except (KeyError, IndexError, TypeError):
        raise ValueError("apply_boundary: boundaries are 'f', 't', 'e', 'p', or 'm'.")


def trunc_pass_applies(bound):
    """helpers.pyx:373: `any(map(lambda s: s[0]=='t', bound))`; a bare string iterates by character."""
    seq = bound if isinstance(bound, (tuple, list)) else list(bound)
    return any(s[0] == "t" for s in seq)


def _image(t, np_dtype, x0=0, y0=0, full=None):
    """CUDA tensor [..., H, W] (contiguous) -> TfmImage."""
    h, w = int(t.shape[-2]), int(t.shape[-1])
    planes = 1
    for s in t.shape[:-2]:
        planes *= int(s)
    item = np.dtype(np_dtype).itemsize
    img = L.TfmImage()
    img.data = t.data_ptr() if t.numel() else None
    img.dtype = L.F32 if np.dtype(np_dtype) == np.float32 else L.F64
    img.h, img.w = h, w
    img.pitch = w * item
    img.x0, img.y0 = x0, y0
    img.full_h, img.full_w = (h, w) if full is None else full
    img.n_planes = max(planes, 1)
    img.plane_pitch = h * w * item
    return img


def _device_out(out, shape, ddt, src):
    """A caller-supplied CUDA tensor as the kernel destination: it must be exactly what the kernel is
    about to write -- same leading axes and grid, the output dtype of this call, dense rows, same
    device.  (The kernel takes pitch and element size from the EXPECTED layout: anything else would
    be written with the wrong stride, or past its end.)"""
    if tuple(int(v) for v in out.shape) != tuple(int(v) for v in shape):
        raise ValueError("out= must have shape %s, not %s" % (tuple(shape), tuple(out.shape)))
    if dev.np_dtype_of(out) != np.dtype(ddt):
        raise ValueError("out= must have dtype %s for this call, not %s" % (np.dtype(ddt).name, out.dtype))
    if not out.is_contiguous():
        raise ValueError("out= must be a contiguous CUDA tensor")
    if out.device != src.device:
        raise ValueError("out= is on %s but the data is on %s" % (out.device, src.device))
    return out


def _src_to_device(data, want_f32):
    """-> (cuda tensor, numpy dtype on device, original numpy dtype, came_from_numpy)."""
    if dev.is_cuda_tensor(data):
        npdt = dev.np_dtype_of(data)
        if npdt is None:
            raise NotImplementedError("transform_b200: CUDA tensors must be float32 or float64")
        t = data.contiguous()
        if want_f32 and npdt != np.float32:
            t = t.float()
            return t, np.dtype(np.float32), npdt, False
        return t, npdt, npdt, False
    if dev.is_tensor(data):
        npdt = dev.np_dtype_of(data)
        if npdt is not None and data.is_pinned() and data.is_contiguous() and \
                (npdt == np.float32 or not want_f32):
            # pinned host tensor: one asynchronous DMA, no staging copy
            dev.require_cuda()
            return data.to("cuda", non_blocking=True), npdt, npdt, True
        data = data.numpy()
    a = np.asarray(data)
    orig = a.dtype
    if want_f32:
        return dev.to_device(a, np.float32), np.dtype(np.float32), orig, True
    if orig == np.float32:
        return dev.to_device(a, np.float32), np.dtype(np.float32), orig, True
    if orig.kind not in "fiub":
        raise ValueError("transform_b200: data must be a real numeric array")
    # every other real dtype is carried as float64 on the device (exact for |v| < 2^53)
    return dev.to_device(a, np.float64), np.dtype(np.float64), orig, True


_PLANS = {}          # prepared argument blocks of repeated device-to-device resample calls (see resample)
_PLANS_MAX = 64


def resample(chain_ops, data, shape, method, bound, fillvalue=0, precision=None, tuning=0,
             dst_origin=(0, 0), src_origin=(0, 0), src_full=None, out=None, oblur=None, sync=True,
             status_host=None):
    """K1 driver.  `chain_ops` is the inverse chain in execution order (list of TfmOp)."""
    lib = L.load()
    # Repeated device-to-device calls (the same lowered chain objects, the same tensors, the same options):
    # the argument block of the first call is reused -- validation, dtype rules and ctypes marshalling cost
    # ~28 us per call in Python, more than the kernel of a 4096^2 frame.
    plan_key = None
    if out is not None and status_host is None and dev.is_cuda_tensor(data) and dev.is_cuda_tensor(out):
        try:
            plan_key = (tuple(map(id, chain_ops)), data.data_ptr(), tuple(data.shape), data.dtype, data.stride(),
                        out.data_ptr(), tuple(out.shape), out.dtype, out.stride(), tuple(shape), method,
                        bound if isinstance(bound, str) else tuple(bound), fillvalue, precision, int(tuning),
                        tuple(dst_origin), tuple(src_origin), src_full, oblur, data.device.index, out.device.index)
            hit = _PLANS.get(plan_key)
        except TypeError:                        # an unhashable option: no plan
            plan_key, hit = None, None
        if hit is not None:
            rc = lib.tfm_resample(hit[0], dev.stream_ptr())
            L.raise_for_status(rc, "tfm_resample")
            return out
    prec = parse_precision(precision)
    m = method[0]
    if m not in _METHODS and m not in _FILTERS:
        if m in _UNSUPPORTED_INTERP:
            raise NotImplementedError(
                "transform_b200: interpND method '%s' (Fourier interpolation) is outside the GPU "
                "hot path; there is no CPU fallback" % m)
        raise ValueError("interpND: valid methods are ('n','l','c','f','s','z','g','h','t')")
    # helpers.pyx:543-544, 748-749
    if oblur is not None:
        if m == "n":
            raise ValueError("interpND: oblur is not allowed with nearest-neighbor interpolation.")
        if m == "c":
            raise ValueError("interpND: oblur is not allowed with cubic-spline interpolation.")
        oblur = float(oblur)
        if not oblur > 0:
            raise ValueError("interpND: oblur must be positive")
    is_filter = m in _FILTERS and (m != "l" or oblur is not None)
    if is_filter and filter_size(m, oblur) > L.FILTER_MAX_SIZE:
        raise NotImplementedError("transform_b200: filter region of %d pixels exceeds the kernel "
                                  "limit of %d" % (filter_size(m, oblur), L.FILTER_MAX_SIZE))
    bx, by = bound_pair(bound)
    fast = prec == L.PREC_FAST32
    src, sdt, orig_dt, from_numpy = _src_to_device(data, fast)
    if src.dim() < 2:
        raise ValueError("map: Transform idim must match input data shape")
    ny, nx = int(shape[-2]), int(shape[-1])
    lead = tuple(int(s) for s in shape[:-2])
    if lead != tuple(int(s) for s in src.shape[:-2]):
        raise NotImplementedError("transform_b200: leading (broadcast) axes of `shape` must equal "
                                  "those of the data")
    # output dtype rules
    if fast:
        ddt = np.dtype(np.float32)
        final_dt = ddt
    elif m == "n":
        fv_dt = np.array(fillvalue).dtype
        final_dt = np.result_type(orig_dt, fv_dt) if trunc_pass_applies(bound) else np.dtype(orig_dt)
        if sdt == np.float32 and final_dt == np.float32:
            ddt = np.dtype(np.float32)
        else:
            ddt = np.dtype(np.float64)
    else:
        ddt = np.dtype(np.float64)
        final_dt = ddt
    host_out = None
    if out is not None and dev.is_cuda_tensor(out):
        dst = _device_out(out, lead + (ny, nx), ddt, src)
    else:
        host_out = out
        dst = dev.empty(lead + (ny, nx), ddt)
    flags = 0
    if (m == "c" and not fast and sdt == np.float32 and not trunc_pass_applies(bound)):
        flags |= L.F_REGION_SRC_DTYPE
    subrect = src_full is not None
    need_status = subrect or bx == L.BOUND_CODE["f"] or by == L.BOUND_CODE["f"]
    status = dev.status_word() if need_status else None

    arr, n = L.make_chain(chain_ops)
    a = L.TfmResampleArgs()
    a.abi_version = L.ABI_VERSION
    a.n_ops = n
    a.chain = ctypes.cast(arr, ctypes.POINTER(L.TfmOp))
    a.src = _image(src, sdt, src_origin[0], src_origin[1], src_full)
    a.dst = _image(dst, ddt, dst_origin[0], dst_origin[1])
    a.method = _FILTERS[m][0] if is_filter else _METHODS[m]
    a.bound_x, a.bound_y = bx, by
    a.precision = prec
    a.flags = flags
    a.tuning = int(tuning)
    a.fillvalue = float(fillvalue)
    a.oblur = float(oblur) if (is_filter and oblur is not None) else 0.0
    a.status_dev = status.data_ptr() if status is not None else None
    rc = lib.tfm_resample(ctypes.byref(a), dev.stream_ptr())
    L.raise_for_status(rc, "tfm_resample")
    _check_status(status, sync, status_host)
    if plan_key is not None and status is None and dst is out and src is data:
        # (no status word: 'forbid' bounds and staged rectangles always take the full path)
        if len(_PLANS) >= _PLANS_MAX:
            _PLANS.clear()
        _PLANS[plan_key] = (ctypes.byref(a), a, arr, list(chain_ops))     # the ops stay alive: their ids stay theirs
    if host_out is not None:
        return _into_host(dst, host_out, sync)
    if not from_numpy:
        return dst
    res = dev.to_host(dst)
    if res.dtype != final_dt:
        res = res.astype(final_dt)
    return res


def _into_host(dst, host_out, sync=True):
    """Device result -> caller-provided host buffer (numpy array or CPU tensor; pinned = one DMA).
    sync=False (pipeline.AsyncResampler): the copy is left in flight on the current stream -- only
    valid for pinned targets, and the caller owns the synchronisation."""
    torch = dev.torch_mod()
    tgt = host_out if dev.is_tensor(host_out) else torch.from_numpy(host_out)
    if tuple(tgt.shape) != tuple(dst.shape) or tgt.dtype != dst.dtype:
        raise ValueError("out= must have shape %s and dtype %s" % (tuple(dst.shape), dst.dtype))
    if not sync and not tgt.is_pinned():
        raise ValueError("asynchronous results need a pinned host buffer")
    tgt.copy_(dst, non_blocking=True)
    if sync:
        torch.cuda.current_stream().synchronize()
    return host_out


def raise_for_status_bits(bits):
    """Status word of a kernel (include/tfm_b200.h, tfm_resample_args.status_dev) -> exception."""
    if bits & 1:
        raise ValueError("apply_boundary: boundary violation with 'forbid' condition")
    if bits & 2:
        raise RuntimeError("transform_b200: a tap fell outside the staged source sub-rectangle "
                           "(sharding bounding box too small)")


def _check_status(status, sync=True, status_host=None):
    """Read the device status word.  Blocking by default (`.item()` is a device sync).  With sync=False
    and a pinned int32 `status_host` the word is copied asynchronously on the current stream instead,
    and whoever waits for the call (pipeline.Ticket.result) passes it to raise_for_status_bits."""
    if status is None:
        return
    if not sync and status_host is not None:
        status_host.copy_(status, non_blocking=True)
        return
    raise_for_status_bits(int(status.item()))


def flux_requested(phot):
    """core.py:521-533: 'flux' / 'extensive' ask for summed-value preservation; 'radiance' /
    'intensive' (default) for local-value preservation.  Only the first character is tested."""
    if not isinstance(phot, str) or not phot:
        raise ValueError("phot must be 'radiance'/'intensive' or 'flux'/'extensive'")
    c = phot[0].lower()
    if c in ("f", "e"):
        return True
    if c in ("r", "i"):
        return False
    raise ValueError("phot must be 'radiance'/'intensive' or 'flux'/'extensive'")


def resample_jacobian(chain_ops, data, shape, method, bound, fillvalue=0, oblur=1.0, iblur=1.0,
                      pad_pow=8, sv_limit=0.25, jump_detect=4.0, precision=None, tuning=0,
                      dst_origin=(0, 0), grid_full=None, src_origin=(0, 0), src_full=None,
                      max_reg=None, out=None, phot='radiance', sync=True, status_host=None):
    """K2 driver: interpND_grid(antialias=True) semantics (helpers.pyx:1293-1530) with the
    coordinates generated in-kernel from `chain_ops`."""
    lib = L.load()
    plan_key = None                              # repeated device-to-device calls: see resample()
    if out is not None and status_host is None and dev.is_cuda_tensor(data) and dev.is_cuda_tensor(out):
        try:
            plan_key = ("J", tuple(map(id, chain_ops)), data.data_ptr(), tuple(data.shape), data.dtype, data.stride(),
                        out.data_ptr(), tuple(out.shape), out.dtype, out.stride(), tuple(shape), method,
                        bound if isinstance(bound, str) else tuple(bound), fillvalue, oblur, iblur, pad_pow, sv_limit,
                        jump_detect, precision, int(tuning), tuple(dst_origin), grid_full, tuple(src_origin), src_full,
                        max_reg, phot, data.device.index, out.device.index)
            hit = _PLANS.get(plan_key)
        except TypeError:
            plan_key, hit = None, None
        if hit is not None:
            rc = lib.tfm_resample_jacobian(hit[0], dev.stream_ptr())
            L.raise_for_status(rc, "tfm_resample_jacobian")
            return out
    prec = parse_precision(precision)
    try:
        filt = L.FILTER_CODE[method[0]]
    except (KeyError, IndexError, TypeError):
        raise ValueError("interpND_grid: method must be 'l','z','h','t', or 'g' (or 'w', the Hann window)")
    try:
        bx, by = bound_pair(bound)
    except ValueError:
        raise ValueError("interpND_grid: each bound must be 'f','t','e','p', or 'm'")
    fast = prec == L.PREC_FAST32
    src, sdt, orig_dt, from_numpy = _src_to_device(data, fast)
    if src.dim() != 2 and len(shape) != src.dim():
        raise ValueError("interpND_grid: index broadcast dims must match source dims")
    if src.dim() < 2:
        raise ValueError("interpND_grid: anti-aliasing is only implemented for 2-D at present.")
    ny, nx = int(shape[-2]), int(shape[-1])
    lead = tuple(int(s) for s in shape[:-2])
    if lead != tuple(int(s) for s in src.shape[:-2]):
        raise NotImplementedError("transform_b200: leading axes of `shape` must equal those of the data")
    ddt = np.dtype(np.float32) if fast else np.dtype(np.float64)
    host_out = None
    if out is not None and dev.is_cuda_tensor(out):
        dst = _device_out(out, lead + (ny, nx), ddt, src)
    else:
        host_out = out
        dst = dev.empty(lead + (ny, nx), ddt)
    subrect = src_full is not None
    need_status = subrect or bx == L.BOUND_CODE["f"] or by == L.BOUND_CODE["f"]
    status = dev.status_word() if need_status else None
    if max_reg is None:
        # the reference accepts sv_limit ("fraction of the input array that may be averaged over by
        # a single output pixel", helpers.pyx:1441-1445) but never uses it; here it bounds the
        # footprint so that a singular map cannot hang the GPU.
        sh = src_full if src_full is not None else (int(src.shape[-2]), int(src.shape[-1]))
        max_reg = max(2, 2 * int(np.ceil(float(sv_limit) * max(sh))))
    gfull = (ny, nx) if grid_full is None else grid_full
    if gfull[0] < 3 or gfull[1] < 3:
        raise ValueError("interpND_grid: the output grid must be at least 3x3 for the Jacobian stencil")

    arr, n = L.make_chain(chain_ops)
    a = L.TfmJacobianArgs()
    a.abi_version = L.ABI_VERSION
    a.n_ops = n
    a.chain = ctypes.cast(arr, ctypes.POINTER(L.TfmOp))
    a.src = _image(src, sdt, src_origin[0], src_origin[1], src_full)
    a.dst = _image(dst, ddt, dst_origin[0], dst_origin[1], gfull)
    a.filter = filt
    a.bound_x, a.bound_y = bx, by
    a.precision = prec
    a.flags = L.F_FLUX if flux_requested(phot) else 0
    a.tuning = int(tuning)
    a.fillvalue = float(fillvalue)
    a.oblur, a.iblur, a.pad_pow = float(oblur), float(iblur), float(pad_pow)
    a.jump_thresh = float(jump_detect) if jump_detect else 0.0
    a.max_reg = int(max_reg)
    a.status_dev = status.data_ptr() if status is not None else None
    rc = lib.tfm_resample_jacobian(ctypes.byref(a), dev.stream_ptr())
    L.raise_for_status(rc, "tfm_resample_jacobian")
    _check_status(status, sync, status_host)
    if plan_key is not None and status is None and dst is out and src is data:
        if len(_PLANS) >= _PLANS_MAX:
            _PLANS.clear()
        _PLANS[plan_key] = (ctypes.byref(a), a, arr, list(chain_ops))
    if host_out is not None:
        return _into_host(dst, host_out, sync)
    if not from_numpy:
        return dst
    return dev.to_host(dst)


def apply(chain_ops, data, precision=None, tuning=0):
    """K3 driver: data [..., L>=2] -> same shape; first two components transformed."""
    lib = L.load()
    prec = parse_precision(precision)
    fast = prec == L.PREC_FAST32
    from_numpy = not dev.is_cuda_tensor(data)
    if from_numpy:
        a_np = np.asarray(data)
        if fast and a_np.dtype == np.float32:
            t, npdt = dev.to_device(a_np, np.float32), np.dtype(np.float32)
        else:
            t, npdt = dev.to_device(a_np, np.float64), np.dtype(np.float64)
    else:
        t = data.contiguous()
        npdt = dev.np_dtype_of(t)
        if npdt is None or (npdt == np.float32 and not fast):
            t = t.double()
            npdt = np.dtype(np.float64)
    vec_len = int(t.shape[-1])
    n_vec = t.numel() // max(vec_len, 1)
    out = dev.torch_mod().empty_like(t)
    arr, n = L.make_chain(chain_ops)
    a = L.TfmApplyArgs()
    a.abi_version = L.ABI_VERSION
    a.n_ops = n
    a.chain = ctypes.cast(arr, ctypes.POINTER(L.TfmOp))
    a.inp = t.data_ptr() if t.numel() else None
    a.out = out.data_ptr() if t.numel() else None
    a.n_vec = n_vec
    a.vec_len = vec_len
    a.dtype = L.F32 if npdt == np.float32 else L.F64
    a.precision = prec
    a.tuning = int(tuning)
    rc = lib.tfm_apply(ctypes.byref(a), dev.stream_ptr())
    L.raise_for_status(rc, "tfm_apply")
    return dev.to_host(out) if from_numpy else out


def apply_linear_nd(nd_ops, dim, data):
    """K3-ND driver: data [..., L >= dim] through a chain of Linear-family ops of dimension `dim` (3..8);
    the first `dim` components are transformed, the rest copied through.  float64 arithmetic and result
    (the reference's np.matmul on float64 stacks)."""
    lib = L.load()
    from_numpy = not dev.is_cuda_tensor(data)
    if from_numpy:
        t = dev.to_device(np.asarray(data), np.float64)
    else:
        t = data.contiguous()
        if dev.np_dtype_of(t) != np.float64:
            t = t.double()
    vec_len = int(t.shape[-1])
    n_vec = t.numel() // max(vec_len, 1)
    out = dev.torch_mod().empty_like(t)
    n = len(nd_ops)
    if n > L.MAX_OPS:
        raise NotImplementedError("transform_b200: chains longer than %d elements are not supported" % L.MAX_OPS)
    arr = (L.TfmLinearNdOp * max(n, 1))()
    for i, o in enumerate(nd_ops):
        arr[i] = o
    a = L.TfmApplyNdArgs()
    a.abi_version = L.ABI_VERSION
    a.n_ops = n
    a.chain = ctypes.cast(arr, ctypes.POINTER(L.TfmLinearNdOp))
    a.dim = int(dim)
    a.vec_len = vec_len
    a.inp = t.data_ptr() if t.numel() else None
    a.out = out.data_ptr() if t.numel() else None
    a.n_vec = n_vec
    a.dtype = L.F64
    rc = lib.tfm_apply_linear_nd(ctypes.byref(a), dev.stream_ptr())
    L.raise_for_status(rc, "tfm_apply_linear_nd")
    return dev.to_host(out) if from_numpy else out


def svd2x2(M):
    """Device svd2x2 on one matrix (test hook, helpers.pyx:1807-1808) -> U, s, V."""
    lib = L.load()
    dev.require_cuda()
    dp = ctypes.POINTER(ctypes.c_double)
    Mi = np.ascontiguousarray(M, dtype=np.float64)
    U, s, V = np.zeros((2, 2)), np.zeros(2), np.zeros((2, 2))
    rc = lib.tfm_svd2x2(Mi.ctypes.data_as(dp), U.ctypes.data_as(dp), s.ctypes.data_as(dp),
                        V.ctypes.data_as(dp))
    L.raise_for_status(rc, "tfm_svd2x2")
    return U, s, V


def svc2x2(U, s, V):
    lib = L.load()
    dev.require_cuda()
    dp = ctypes.POINTER(ctypes.c_double)
    Ui = np.ascontiguousarray(U, dtype=np.float64)
    si = np.ascontiguousarray(s, dtype=np.float64)
    Vi = np.ascontiguousarray(V, dtype=np.float64)
    M = np.zeros((2, 2))
    rc = lib.tfm_svc2x2(Ui.ctypes.data_as(dp), si.ctypes.data_as(dp), Vi.ctypes.data_as(dp),
                        M.ctypes.data_as(dp))
    L.raise_for_status(rc, "tfm_svc2x2")
    return M
