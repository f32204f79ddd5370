import sys, os
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import transform_b200 as t
from transform_b200 import _lib, _engine, sharding
import oracle_np as o
from conftest import to_oracle
n = 8192
g = torch.Generator(device="cuda").manual_seed(1000)
src = torch.rand((n, n), generator=g, device="cuda", dtype=torch.float32)
tr = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit="deg"), t.Offset([-n / 2.0, -n / 2.0])])
want = tr.resample(src, method="linear", precision="fp32")
print("full:", _lib.last_kernel())
ops = tr._int_chain(True)
split, plans, staged = sharding.plan_rects(tr, (n, n), (n, n), "linear", "t", 2, sharding._default_invert, split="rows")
for p in plans:
    x0, y0, x1, y1 = p["rect"]
    bx0, by0, bx1, by1 = p["box"]
    sub = src[by0:by1, bx0:bx1].contiguous()
    blk = _engine.resample(ops, sub, (y1 - y0, x1 - x0), "l", "t", dst_origin=(x0, y0), src_origin=(bx0, by0),
                           src_full=(n, n), precision="fp32")
    print("rect", p["rect"], "box", p["box"], _lib.last_kernel())
    d = (blk - want[y0:y1, x0:x1]).abs()
    ys, xs = torch.nonzero(d > 0, as_tuple=True)
    print("  differing", len(ys))
    for i in range(min(6, len(ys))):
        y, x = int(ys[i]) + y0, int(xs[i]) + x0
        c = to_oracle(tr).reverse(np.array([[float(x), float(y)]]))[0]
        ov = o.resample(to_oracle(tr), src.cpu().numpy(), method='linear', rows=(y, y + 1))[0, x]
        print("   pixel", (x, y), "coords", c, "sharded", float(blk[y - y0, x - x0]), "full", float(want[y, x]), "oracle", ov)
# whole frame through the generic kernel (bound 'e' on one axis forces it? no: use a 1-pixel-smaller staged rect)
