"""Small ragged-shape invocations of K1 (all variants), K2 (both modes), K3 and the helper kernels for
compute-sanitizer (memcheck).  Shapes are chosen to exercise partial tiles, the last-column tile,
out-of-image footprints, every boundary mode, planes, sub-rectangles and the texture path."""
import sys
import numpy as np
sys.path.insert(0, '.')
import transform_b200 as t
from transform_b200 import _engine

rng = np.random.default_rng(0)
for (ny, nx) in ((33, 65), (70, 97), (1, 1), (5, 130)):
    src32 = rng.random((ny, nx)).astype(np.float32)
    src64 = rng.random((ny, nx))
    aff = t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit='deg'), t.Offset([-nx / 2.0, -ny / 2.0])])
    for m in ('n', 'l', 'c'):
        for b in ('t', 'e', 'p', 'm', ['p', 't']):
            aff.resample(src32, method=m, bound=b)
            aff.resample(src64, method=m, bound=b, shape=(ny + 3, nx + 5))
        aff.resample(src32, method=m, precision='fp32')
        aff.resample(src32, method=m, precision='fp32', fillvalue=2.5)
    if ny >= 3 and nx >= 3:
        rad = t.Composition([t.Scale([nx / (2 * np.pi), ny / (0.7 * nx)]), t.Radial(origin=[nx / 2.0, ny / 2.0])])
        for prec in ('fp64', 'fp32'):
            rad.resample(src32, method='G', precision=prec)
            rad.resample(src32, method='H', precision=prec, bound='m')
            rad.inverse().resample(src32, method='G', precision=prec, bound='p')
        rad.resample(src32, method='linear', precision='fp32')
        t.jacobian(rad.invert(np.mgrid[0:nx, 0:ny].T.astype(float)))
cube = rng.random((3, 40, 50)).astype(np.float32)
t.Rotation(0.3).resample(cube, method='linear')
t.Rotation(0.3).resample(cube, method='linear', precision='fp32')
big = rng.random((96, 128)).astype(np.float32)
ops = t.Offset([0.25, 0.25])._chain(True)
for m in ('n', 'l', 'c'):       # staged sub-rectangle [10:60, 20:100] of a 96 x 128 frame, output block at (25, 15)
    _engine.resample(ops, np.ascontiguousarray(big[10:60, 20:100]), (40, 70), m, 't', dst_origin=(25, 15),
                     src_origin=(20, 10), src_full=(96, 128))
try:                            # a deliberately too small staging rectangle must be reported, not read past
    _engine.resample(t.Rotation(0.2)._chain(True), np.ascontiguousarray(big[10:60, 20:100]), (96, 128), 'l', 't',
                     src_origin=(20, 10), src_full=(96, 128))
    raise SystemExit("staged-rectangle miss was not reported")
except RuntimeError:
    pass
for n in (0, 1, 7, 1001):
    v = rng.random((n, 2))
    t.Rotation(0.3).apply(v)
    t.Radial().apply(v.astype(np.float32), precision='fp32')
    t.Rotation(0.3).apply(rng.random((n, 5)))
print("sanitize cases done")
