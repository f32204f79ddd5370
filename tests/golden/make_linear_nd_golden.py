#!/usr/bin/env python3
"""Golden vectors for Linear-family transforms beyond two dimensions (basic.py:141-176, 338-506) from the
UNMODIFIED reference: Rotation from Euler angles and from axis pairs, Scale o Rotation o Offset chains and a
general matrix with pre / post offsets in 3-D, plus one 4-D chain; forward and inverted, with pass-through
components.  Run in the authoring container (needs /root/reference).  Output: tests/golden/linear_nd_golden.npz.

Each case stores the constructor arguments as arrays, so that the tests rebuild the SAME transform from the
same numbers with transform_b200 (GPU) and with the oracle (CPU)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402


def build(p, case, a):
    """The transform of `case` from package `p` (the reference, transform_b200, ...) and the argument arrays `a`."""
    if case == "euler":
        return p.Rotation(euler=a["euler"], unit="deg")
    if case == "pairs":
        return p.Rotation([[1, 2, float(a["euler"][0])], [0, 1, float(a["euler"][1])]], unit="deg")
    if case == "chain":
        return p.Composition([p.Scale(a["scale3"]), p.Rotation(euler=a["euler"], unit="deg"), p.Offset(a["offset3"])])
    if case == "general":
        return p.Linear(matrix=a["matrix3"], pre=a["pre3"], post=a["post3"])
    if case == "chain4":
        return p.Composition([p.Scale(a["scale4"]), p.Linear(matrix=a["matrix4"], post=a["post4"])])
    raise KeyError(case)


CASES = ("euler", "pairs", "chain", "general", "chain4")


def main():
    ref = ref_import.load_reference()
    rng = np.random.default_rng(77)
    a = {"euler": rng.uniform(-180, 180, size=3), "scale3": rng.uniform(0.5, 2.0, size=3),
         "offset3": rng.uniform(-5, 5, size=3), "matrix3": rng.standard_normal((3, 3)) + 2 * np.eye(3),
         "pre3": rng.uniform(-1, 1, size=3), "post3": rng.uniform(-1, 1, size=3),
         "scale4": rng.uniform(0.5, 2.0, size=4), "matrix4": rng.standard_normal((4, 4)) + 2 * np.eye(4),
         "post4": rng.uniform(-1, 1, size=4),
         "v3": rng.uniform(-50, 50, size=(257, 3)), "v5": rng.uniform(-50, 50, size=(64, 5))}
    out = {"arg/" + k: v for k, v in a.items()}
    for case in CASES:
        tr = build(ref, case, a)
        for vn in ("v3", "v5"):
            if case == "chain4" and vn == "v3":
                continue
            for inv in (False, True):
                out["%s/%s/%s" % (case, vn, "inv" if inv else "fwd")] = tr.apply(a[vn].copy(), invert=inv)
    np.savez_compressed(os.path.join(HERE, "linear_nd_golden.npz"), **out)
    print("wrote linear_nd_golden.npz: %d arrays" % len(out))


if __name__ == "__main__":
    main()
