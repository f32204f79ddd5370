"""Shared test plumbing.

Markers: `gpu` = needs a CUDA device (run on the B200 box with `-m gpu`); everything else runs on
CPU.  The oracle (oracle/) is imported here and ONLY under tests/, bench.py's baseline legs and
__graft_entry__.smoke().
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (B200); deselected on CPU runs")


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def to_oracle(tr):
    """Build the oracle twin of a transform object from its `params` dictionary.

    Works on transform_b200 objects AND on objects of the real reference package (same params
    layout by design), so product, oracle and reference always share parameter VALUES -- including
    the memory order of matrix/matinv, which decides NumPy's gemv rounding."""
    import oracle_np as o
    name = type(tr).__name__
    p = tr.params
    if name in ("Linear", "Scale", "Rotation", "Offset"):
        return o.Linear(matrix=p.get("matrix"), pre=p.get("pre"), post=p.get("post"),
                        matinv=p.get("matinv"))
    if name == "PlusOne_":
        return o.Linear(pre=[1.0, 0.0])
    if name == "Radial":
        return o.Radial(origin=np.asarray(p["origin"], dtype=float)[0:2].reshape(-1)[0:2],
                        r0=p["r0"], ang_coeff=p["ang_coeff"], ccw=p["ccw"], pos_only=p["pos"])
    if name == "Quadpin":
        return o.Quadpin(origin=p["origin"], length=p["length"], strength=p["strength"],
                         dim=p["dim"], idim=tr.idim)
    if name == "Cubicpin":
        return o.Cubicpin(origin=p["origin"], length=p["length"], strength=p["strength"],
                          dim=p["dim"], idim=tr.idim)
    if name in ("Poly", "Quarticpin"):
        return o.Poly(origin=p["origin"], length=p["length"], strength=p["strength"], degree=p["degree"],
                      dim=p["dim"], idim=tr.idim, den=p["den"])
    if name == "Identity":
        return o.Identity()
    if name == "Inverse":
        return o.Inverse(to_oracle(p["transform"]))
    if name in ("Composition", "Wrap"):
        return o.Composition([to_oracle(x) for x in p["list"]])
    if name == "WCS":
        w = p["wcs"]
        proj = w.projection()
        from transform_b200.fitswcs import _PROJ_NAME
        code = None if proj is None else _PROJ_NAME[proj[0]]
        if w.tpv is not None:
            core = o.WCS_TPV(w.crpix, w.linear_matrix(), w.crval, w.tpv, w.lonpole, w.latpole)
        else:
            core = o.WCS(w.crpix, w.linear_matrix(), w.crval, code, w.lonpole, w.latpole,
                         cea_lambda=w.cea_lambda, sin_slant=w.sin_slant)
        if w.has_sip():
            return o.WCS_SIP(core, o.SIP(w.crpix, w.sip_a, w.sip_b))
        return core
    raise NotImplementedError(name)


@pytest.fixture(scope="session")
def oracle():
    import oracle_np
    return oracle_np


@pytest.fixture(scope="session")
def tfm():
    import transform_b200
    return transform_b200


def assert_close_rel(got, want, rel, what=""):
    """max |got-want| <= rel * max(1, max|want|)   (NaN must match NaN)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), "%s: NaN pattern differs (%d vs %d)" % (what, nan_g.sum(), nan_w.sum())
    if got.size == 0:
        return
    scale = max(1.0, float(np.nanmax(np.abs(want))) if (~nan_w).any() else 1.0)
    err = float(np.nanmax(np.abs(got - want))) if (~nan_w).any() else 0.0
    assert err <= rel * scale, "%s: max abs err %.3e > %.1e * %.3g" % (what, err, rel, scale)
