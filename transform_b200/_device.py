"""Device-buffer plumbing (PyTorch is used for allocation, streams and copies only).

Nothing here computes: every arithmetic step of the hot path runs in libtfm_b200.so.
"""
import numpy as np

_pinned_cache = {}


class NoCudaDevice(RuntimeError):
    pass


def torch_mod():
    import torch
    return torch


def require_cuda():
    torch = torch_mod()
    if not torch.cuda.is_available():
        raise NoCudaDevice("transform_b200: no CUDA device is visible and there is no CPU fallback "
                           "(the reference's NumPy path is not part of this package)")
    return torch


def is_cuda_tensor(x):
    try:
        import torch
    except ImportError:
        return False
    return isinstance(x, torch.Tensor) and x.is_cuda


def is_tensor(x):
    try:
        import torch
    except ImportError:
        return False
    return isinstance(x, torch.Tensor)


def stream_ptr():
    """Raw handle of the current CUDA stream.  torch.cuda.current_stream() costs ~15 us of Python per
    call (half of a whole resample call on a small frame); the C accessor behind it does not."""
    torch = torch_mod()
    try:
        return int(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))
    except AttributeError:
        return int(torch.cuda.current_stream().cuda_stream)


def _pinned(nbytes, slot):
    """Reusable pinned staging buffer (cudaHostAlloc is slow; keep one per slot, grow only)."""
    torch = torch_mod()
    buf = _pinned_cache.get(slot)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
        _pinned_cache[slot] = buf
    return buf


def to_device(arr, np_dtype, slot="h2d"):
    """numpy array -> contiguous CUDA tensor of np_dtype, staged through pinned memory."""
    torch = require_cuda()
    a = np.ascontiguousarray(arr, dtype=np_dtype)
    nbytes = a.nbytes
    stage = _pinned(nbytes, slot)
    view = stage[:nbytes].numpy().view(np_dtype).reshape(a.shape) if nbytes else None
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(np_dtype)]
    dev = torch.empty(a.shape, dtype=tdt, device="cuda")
    if nbytes:
        np.copyto(view, a)
        dev.copy_(torch.from_numpy(view), non_blocking=True)
        # the staging buffer is reused by the next call: make sure the copy has left it
        torch.cuda.current_stream().synchronize()
    return dev


def to_host(t, slot="d2h"):
    """CUDA tensor -> fresh numpy array, staged through pinned memory."""
    torch = torch_mod()
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    npdt = {torch.float32: np.float32, torch.float64: np.float64}[t.dtype]
    out = np.empty(tuple(t.shape), dtype=npdt)
    if nbytes:
        stage = _pinned(nbytes, slot)
        view = stage[:nbytes].view(t.dtype).reshape(t.shape)
        view.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        np.copyto(out, view.numpy())
    return out


def empty(shape, np_dtype):
    torch = require_cuda()
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(np_dtype)]
    return torch.empty(tuple(shape), dtype=tdt, device="cuda")


def status_word():
    torch = require_cuda()
    return torch.zeros(1, dtype=torch.int32, device="cuda")


def np_dtype_of(t):
    torch = torch_mod()
    return {torch.float32: np.dtype(np.float32), torch.float64: np.dtype(np.float64)}.get(t.dtype)


def bind_to_gpu_numa(device_index):
    """Pin the calling process to the CPU cores closest to a GPU (NVML's ideal affinity mask), so that
    pinned host buffers allocated afterwards land on that GPU's NUMA node.  With one process per GPU
    all streaming host<->device copies otherwise funnel through whichever node the allocator picked.
    Best effort: returns the CPU set, or None when NVML / the syscall is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None
