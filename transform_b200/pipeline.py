"""Stream-pipelined resampling of host frames: upload, kernel and download of consecutive calls overlap.

A synchronous `Transform.resample(host_frame, out=host_buffer)` is a serial chain -- H2D copy, kernel,
D2H copy -- and for the bandwidth-bound methods the two PCIe copies take 50x longer than the kernel.
`AsyncResampler` gives each submitted call its own CUDA stream (round-robin over `depth` streams):
the H2D copy of call k+1, the kernel of call k and the D2H copy of call k-1 run concurrently, and the
link is used in both directions at once.  Same kernels, same results; only the scheduling differs.

    pipe = AsyncResampler(depth=4)
    tickets = [pipe.submit(tr, frame, out=buf, method='linear', precision='fp32')
               for frame, buf in zip(pinned_frames, pinned_outputs)]
    for t in tickets:
        result = t.result()          # blocks until that call's D2H copy has landed in `buf`
    pipe.drain()

Inputs and outputs must be pinned CPU tensors (`torch.empty(...).pin_memory()`): pageable memory
would turn every copy into a synchronous staged one.
"""
from . import _device as dev


class Ticket:
    """Handle of one submitted call: `.result()` waits for its output buffer to be complete."""

    def __init__(self, event, out, status_host=None):
        self._event = event
        self._out = out
        self._status = status_host

    def done(self):
        return self._event.query()

    def result(self):
        """Wait for the call; raises what the synchronous call would have raised ('forbid' boundary
        violated -> ValueError), read from the status word that travelled back with the result."""
        self._event.synchronize()
        if self._status is not None:
            from . import _engine
            _engine.raise_for_status_bits(int(self._status[0]))
        return self._out


class AsyncResampler:
    def __init__(self, depth=4, device=None):
        torch = dev.require_cuda()
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self._torch = torch
        self._device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._streams = [torch.cuda.Stream(device=self._device) for _ in range(depth)]
        self._events = [None] * depth
        # one pinned status word per slot: a 'forbid' boundary does not serialise the pipeline
        self._status = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(depth)]
        self._next = 0

    def submit(self, transform, data, out, method='n', bound='t', **kw):
        """Enqueue transform.resample(data, method, bound, out=out, **kw); returns a Ticket."""
        torch = self._torch
        if not (dev.is_tensor(data) and not data.is_cuda and data.is_pinned()):
            raise ValueError("AsyncResampler.submit: `data` must be a pinned CPU tensor")
        if not (dev.is_tensor(out) and not out.is_cuda and out.is_pinned()):
            raise ValueError("AsyncResampler.submit: `out` must be a pinned CPU tensor")
        slot = self._next
        self._next = (slot + 1) % len(self._streams)
        stream = self._streams[slot]
        prev = self._events[slot]
        if prev is not None:
            prev.synchronize()               # bound the queue: at most `depth` calls in flight
        # work submitted by the caller on the current stream (e.g. filling `data`) comes first
        stream.wait_stream(torch.cuda.current_stream(self._device))
        status_host = self._status[slot]
        status_host.zero_()
        with torch.cuda.stream(stream):
            transform.resample(data, method=method, bound=bound, out=out, sync=False,
                               status_host=status_host, **kw)
            ev = torch.cuda.Event()
            ev.record(stream)
        self._events[slot] = ev
        return Ticket(ev, out, status_host)

    def submit_many(self, data, calls):
        """One upload, several resamplings of the same frame: `calls` is a list of
        (transform, out, method, kwargs).  The frame crosses PCIe once; each result goes back into its
        own pinned buffer.  Returns one Ticket (its `.result()` is the list of outputs)."""
        torch = self._torch
        if not (dev.is_tensor(data) and not data.is_cuda and data.is_pinned()):
            raise ValueError("AsyncResampler.submit_many: `data` must be a pinned CPU tensor")
        for c in calls:
            if not (dev.is_tensor(c[1]) and not c[1].is_cuda and c[1].is_pinned()):
                raise ValueError("AsyncResampler.submit_many: every `out` must be a pinned CPU tensor")
        slot = self._next
        self._next = (slot + 1) % len(self._streams)
        stream = self._streams[slot]
        prev = self._events[slot]
        if prev is not None:
            prev.synchronize()
        stream.wait_stream(torch.cuda.current_stream(self._device))
        status_host = self._status[slot]
        status_host.zero_()
        with torch.cuda.stream(stream):
            frame = data.to(self._device, non_blocking=True)          # the one H2D copy of this step
            for (transform, out, method, kw) in calls:
                kw = dict(kw or {})
                bound = kw.pop("bound", "t")
                transform.resample(frame, method=method, bound=bound, out=out, sync=False,
                                   status_host=status_host, **kw)
            ev = torch.cuda.Event()
            ev.record(stream)
            frame.record_stream(stream)
        self._events[slot] = ev
        return Ticket(ev, [c[1] for c in calls], status_host)

    def drain(self):
        for ev in self._events:
            if ev is not None:
                ev.synchronize()
