#!/usr/bin/env python3
"""tests/golden/make_golden.py -- regenerate the golden fixtures FROM THE REAL REFERENCE.

Runs only in the authoring container (needs /root/reference): imports the UNMODIFIED reference
package through oracle/ref_import.py (its helpers.pyx compiled where it lies, a stub for the absent
astropy) and records inputs + outputs of the hot path as small .npz files next to this script.
The tests (CPU: oracle vs golden; GPU: CUDA vs golden) only read the .npz files.

    python tests/golden/make_golden.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_import  # noqa: E402


def chains(t):
    """name -> (constructor spec used by tests/golden_specs.py, reference Transform)."""
    return {
        "affine": t.Composition([t.Scale(1.3, dim=2), t.Rotation(30, unit='deg'), t.Offset([-32.0, -24.0])]),
        "aniso": t.Scale([0.7, 1.9], post=[2.0, -1.0], pre=[-3.0, 0.5]),
        "linear": t.Linear(matrix=np.array([[1.2, 0.3], [-0.4, 0.9]]), pre=np.array([1., 2.]),
                           post=np.array([-3., 4.])),
        "wrap": t.Wrap(t.Rotation(0.4), t.Offset([30.0, 20.0])),
        "inverse": t.Composition([t.Scale(0.8, dim=2), t.Offset([3.5, -2.25])]).inverse(),
        "radial": t.Composition([t.Scale([64 / (2 * np.pi), 48 / 40.0]), t.Radial(origin=[31.5, 23.5])]),
        "radial_log": t.Composition([t.Scale([64 / (2 * np.pi), 48 / 3.0]),
                                     t.Radial(origin=[31.5, 23.5], r0=2.0, ccw=True, pos_only=False)]),
    }


def pin_chains(t):
    """Per-axis pincushion family (basic.py:846-1462); the forward-only Poly is used inverted, which
    is how it serves resample()."""
    return {
        # `dim=` given: the reference writes the result back into its copy of the INTEGER pixel grid,
        # so resample() sees truncated coordinates (quirk reproduced, see oracle_np._Pin._join)
        "quad": t.Quadpin(dim=2, strength=-0.3, length=40.0, origin=np.array([30.0, 20.0])),
        "cubic": t.Cubicpin(dim=2, strength=0.4, length=60.0, origin=np.array([31.5, 23.5])),
        "poly_inv": t.Poly(dim=2, length=50.0, strength=0.2, degree=4, origin=np.array([31.5, 23.5])).inverse(),
        "quad_dim1": t.Quadpin(dim=1, strength=0.5, length=25.0),
        # idim/odim given instead: float coordinates all the way (the usable spelling)
        "quad_f": t.Quadpin(idim=2, odim=2, strength=-0.3, length=40.0, origin=np.array([30.0, 20.0])),
        "quad_vec": t.Quadpin(idim=2, odim=2, strength=np.array([0.2, -0.1]), length=np.array([30., 50.])),
        "cubic_f": t.Cubicpin(idim=2, odim=2, strength=0.4, length=60.0, origin=np.array([31.5, 23.5])),
        "cubic_neg": t.Cubicpin(idim=2, odim=2, strength=-0.2, length=80.0),
        "poly_f": t.Poly(idim=2, odim=2, length=50.0, strength=0.2, degree=4,
                         origin=np.array([31.5, 23.5])).inverse(),
        "quartic_inv": t.Quarticpin(idim=2, odim=2, length=60.0).inverse(),
        "quad_chain": t.Composition([t.Offset([2.0, -1.0]), t.Quadpin(dim=2, strength=0.15, length=35.0,
                                                                      origin=np.array([31.5, 23.5])),
                                     t.Scale(1.1, dim=2)]),
        "poly_den0": t.Poly(idim=2, odim=2, length=45.0, strength=-0.2, degree=5, den=0).inverse(),
    }


def main():
    warnings.simplefilter("ignore")
    t = ref_import.load_reference()
    rng = np.random.default_rng(20201108)
    src32 = rng.random((48, 64)).astype(np.float32)
    src64 = rng.random((48, 64))
    out = {"src32": src32, "src64": src64}
    ch = chains(t)
    bounds = ['t', 'e', 'p', 'm']
    for name, tr in ch.items():
        for m in ('nearest', 'linear', 'cubic'):
            for b in bounds:
                for sname, src in (("f32", src32), ("f64", src64)):
                    if name not in ("affine", "radial") and (b not in ('t', 'm') or sname == "f64"):
                        continue                      # keep the fixture small
                    out["resample/%s/%s/%s/%s" % (name, m[0], b, sname)] = tr.resample(src, method=m, bound=b)
        out["resample/%s/shape" % name] = tr.resample(src64, method='linear', shape=(20, 37))
        v = rng.random((257, 2)) * 80 - 20
        if "radial" in name:
            v = np.abs(v) * np.array([0.07, 0.4]) + 0.05
        out["apply/%s/in" % name] = v
        out["apply/%s/fwd" % name] = tr.apply(v)
        out["apply/%s/rev" % name] = tr.invert(tr.apply(v))
        v3 = rng.random((5, 7, 4)) * 10 + 1
        out["apply/%s/in3" % name] = v3
        out["apply/%s/fwd3" % name] = tr.apply(v3)
    # interpND with explicit indices, incl. NaN / inf / ties / per-axis boundaries / fillvalue
    idx = rng.random((21, 17, 2)) * np.array([80, 60]) - 8
    idx[3, 4] = [np.nan, 2]
    idx[5, 6] = [1e300, -1e300]
    idx[7, 7] = [np.inf, 3]
    idx[8, 8] = [-0.5, 0.5]
    idx[9, 9] = [47.5, 63.5]
    idx[10, :5, 0] = [-1.5, -0.5, 0.5, 1.5, 2.5]
    out["interp/index"] = idx
    for m in 'nlc':
        for b in ('t', 'e', 'p', 'm', ['p', 'e'], ['m', 't'], 'extend', 'truncate'):
            for fv in (0, -9):
                key = "interp/%s/%s/%s" % (m, b if isinstance(b, str) else "+".join(b), fv)
                out[key] = t.interpND(src32, idx, method=m, bound=b, fillvalue=fv)
    # the collapsible input-plane filters (helpers.pyx:703-735, 792-835), with and without oblur
    for m in 'szghtl':
        for b in ('t', 'e', 'p', 'm', ['p', 'e']):
            for ob in (None, 1.7, 0.6):
                if m == 'l' and ob is None:
                    continue
                if ob == 0.6 and b != 't':
                    continue
                key = "filt/%s/%s/%s" % (m, b if isinstance(b, str) else "+".join(b), ob)
                out[key] = t.interpND(src32, idx, method=m, bound=b, fillvalue=-9, oblur=ob)
        out["filt64/%s" % m] = t.interpND(src64, idx, method=m, bound='m', oblur=1.3)
    for m in ('sinc', 'zlanczos', 'gaussian', 'hann', 'tukey'):
        out["fresample/affine/%s/t/f32" % m[0]] = ch["affine"].resample(src32, method=m, bound='t')
        out["fresample/radial/%s/p/f64" % m[0]] = ch["radial"].resample(src64, method=m, bound='p', shape=(20, 37))
    # per-axis pincushion transforms
    vp = rng.random((301, 2)) * 160 - 60
    out["pin/in"] = vp
    for name, tr in pin_chains(t).items():
        if not tr.no_forward:
            out["pin/%s/fwd" % name] = tr.apply(vp.copy())
        if not tr.no_reverse:
            out["pin/%s/rev" % name] = tr.invert(vp.copy())
            if name != "quad_dim1":
                out["pin/%s/lin" % name] = tr.resample(src32, method='linear')
                out["pin/%s/near" % name] = tr.resample(src64, method='nearest', bound='m')
    # Jacobian pieces that the reference can run (helpers.pyx:846-958, 960-1141, 1807-1810)
    # forward through the polar transform: atan2 wraps at the branch cut, which is what jump_detect
    # exists for (the inverse direction, cos/sin, is continuous)
    coords = ch["radial"].apply(np.mgrid[range(40), range(36)].transpose())
    out["jac/coords"] = coords
    sq = ch["radial"].apply(np.mgrid[range(40), range(40)].transpose())      # square grid: no R12 ambiguity
    out["jac/coords_sq"] = sq
    out["jac/jump4_sq"] = t.jump_detect(t.simple_jacobian(sq), 4.0)
    out["jac/jump1p2_sq"] = t.jump_detect(t.simple_jacobian(sq), 1.2)
    Js = t.simple_jacobian(coords)
    out["jac/simple"] = Js
    out["jac/jump4"] = t.jump_detect(Js, 4.0)
    out["jac/jump1p2"] = t.jump_detect(Js, 1.2)
    Ms = rng.random((16, 2, 2)) * 4 - 2
    Us, ss, Vs = np.zeros((16, 2, 2)), np.zeros((16, 2)), np.zeros((16, 2, 2))
    for k in range(16):
        t.svd2x2(Ms[k], Us[k], ss[k], Vs[k])
    out["svd/M"], out["svd/U"], out["svd/s"], out["svd/V"] = Ms, Us, ss, Vs
    # the reference's WORKING alternative anti-aliaser (SVD.pyx::map_coordinates) on an affine
    # downsample: a cross-check (different variant of the same algorithm), not a bit pin
    svd = ref_import.load_ref_svd()
    tgt = np.zeros((24, 32))
    svd.map_coordinates(src64, tgt, lambda c: t.Scale([2.0, 2.0]).apply(c))
    out["svdpyx/target"] = tgt
    # the reference's OWN anti-aliased path, built with the named repairs (oracle/build_ref_repaired.py):
    # square output grids (R12 is not applied there), no mirror boundary (R9 is not applied there)
    hr = ref_import.load_ref_helpers_repaired()
    grid48 = np.mgrid[range(48), range(48)].transpose()
    aa_maps = {"affine": t.Composition([t.Offset([30.0, 22.0]), t.Scale([0.55, 0.8]), t.Rotation(25, unit='deg'),
                                        t.Offset([-24.0, -24.0])]),
               "polar": t.Composition([t.Scale([48 / (2 * np.pi), 48 / 30.0]), t.Radial(origin=[31.5, 23.5])])}
    for name, tr in aa_maps.items():
        coords = tr.invert(grid48)
        out["aa/%s/coords" % name] = coords
        for m in "glht":
            for b in ("t", "e", "p"):
                out["aa/%s/%s/%s" % (name, m, b)] = np.asarray(hr.interpND_grid(src64, coords.copy(), method=m, bound=b, strict=True))
        out["aa/%s/g/t/nojump" % name] = np.asarray(hr.interpND_grid(src64, coords.copy(), method='g', bound='t',
                                                                      jump_detect=0, strict=True))
        out["aa/%s/g/e/oblur2" % name] = np.asarray(hr.interpND_grid(src64, coords.copy(), method='g', bound='e',
                                                                      oblur=2.0, strict=True))
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    print("wrote %d arrays, %d bytes" % (len(out), os.path.getsize(os.path.join(HERE, "reference_golden.npz"))))


if __name__ == "__main__":
    main()
