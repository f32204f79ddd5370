"""Where does the end-to-end step go?  Raw pinned copies vs the pipelined API, several depths (GPU box)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import transform_b200 as t
n = 8192
g = torch.Generator(device="cuda").manual_seed(1)
src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
h_src = torch.empty((n, n), dtype=torch.float32).pin_memory(); h_src.copy_(src)
h_out = [torch.empty((n, n), dtype=torch.float32).pin_memory() for _ in range(4)]
d = [torch.empty((n, n), device="cuda", dtype=torch.float32) for _ in range(4)]
def wall(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
mb = n * n * 4 / 1e6
ms = wall(lambda: d[0].copy_(h_src, non_blocking=True)); print("H2D alone      %.2f ms  %.1f GB/s" % (ms, mb / ms))
ms = wall(lambda: h_out[0].copy_(d[0], non_blocking=True)); print("D2H alone      %.2f ms  %.1f GB/s" % (ms, mb / ms))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d[1].copy_(h_src, non_blocking=True)
    with torch.cuda.stream(s2): h_out[1].copy_(d[2], non_blocking=True); h_out[2].copy_(d[3], non_blocking=True)
ms = wall(both); print("1 H2D + 2 D2H concurrent  %.2f ms  (D2H %.1f GB/s, H2D %.1f GB/s)" % (ms, 2 * mb / ms, mb / ms))
c = (n - 1) / 2.0
tr_aff = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
tr_rad = t.Composition([t.Scale([n / (2 * np.pi), np.sqrt(2.0)]), t.Radial(origin=[c, c])])
for depth in (2, 3, 4, 6):
    pipe = t.AsyncResampler(depth=depth)
    k = [0]
    def step():
        i = k[0] & 1; k[0] += 1
        pipe.submit_many(h_src, [(tr_aff, h_out[i], "linear", {"precision": "fp32"}), (tr_rad, h_out[2 + i], "Gaussian", {"precision": "fp32"})])
    for _ in range(3): step()
    pipe.drain(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(12): step()
    pipe.drain(); torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / 12 * 1e3
    print("submit_many depth %d: %.2f ms per step -> %.0f Mpix/s" % (depth, ms, 2 * n * n / ms / 1e3))
