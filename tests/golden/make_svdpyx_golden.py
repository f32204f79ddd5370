#!/usr/bin/env python3
"""Golden vectors from the reference's WORKING anti-aliaser, SVD.pyx::map_coordinates (SVD.pyx:193-292),
on maps where its algorithm and the repaired interpND_grid coincide tap for tap.

Run in the authoring container (needs /root/reference; oracle/build_ref.py compiles SVD.pyx where it
lies into oracle/_ref/).  Output: tests/golden/svdpyx_golden.npz.

When do the two coincide?  map_coordinates evaluates its window at INTEGER offsets from the exact source
position and samples the source there with nearest-neighbour (order=0); interpND_grid evaluates it at
the integer taps around floor(position).  With integer source positions both are the same taps and
the same window arguments.  map_coordinates inverts the UNPADDED singular values and applies the
inverse of Ji^T (its Ji holds d(source)/d(x) in row 0): for a diagonal Jacobian with both scales > 1
that equals interpND_grid's padded inverse once pad_pow is large ((1 + s^64)^(-1/64) = 1/s to the last
bit for s >= 2).  Its window is the Hann window (cos(pi x) + 1)/2 per axis (SVD.pyx:110-111), its
boundary 'extend' (clip of the rounded index), its flux rule |det Ji| (SVD.pyx:285-286).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402

CASES = {            # name: (scale_x, scale_y, offset_x, offset_y, target shape)
    "s2x3": (2, 3, 3, 1, (14, 26)),
    "s3x2": (3, 2, 0, 2, (20, 18)),
    "s4x2": (4, 2, 1, 0, (22, 12)),
}


def main():
    svd = ref_import.load_ref_svd()
    src = np.load(os.path.join(HERE, "reference_golden.npz"))["src64"]
    out = {"src64": src}
    for name, (sx, sy, ox, oy, shape) in CASES.items():
        def ci(c, sx=sx, sy=sy, ox=ox, oy=oy):
            r = np.array(c, dtype=np.float64)
            r[..., 0] = r[..., 0] * sx + ox
            r[..., 1] = r[..., 1] * sy + oy
            return r
        for flux in (0, 1):
            tgt = np.zeros(shape)
            svd.map_coordinates(src, tgt, ci, conserve_flux=flux, order=0)
            out["%s/%s" % (name, "flux" if flux else "radiance")] = tgt
        out["%s/params" % name] = np.array([sx, sy, ox, oy], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "svdpyx_golden.npz"), **out)
    print("wrote", sorted(out))


if __name__ == "__main__":
    main()
