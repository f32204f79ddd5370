"""SIP distortion and projection parameters of FITS WCS headers -- CPU side.

PARITY UNPINNED: the reference hands both to astropy (core.py:1727, 1742), absent here.  What can be
checked without it: the oracle's restatement against hand-computed polynomials, against an independent
root finder for the inverse, the header reader, and that the product lowers a SIP header to
[TFM_OP_SIP, TFM_OP_WCS] only when a CUDA device can hold the coefficient table (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_np as o
import transform_b200 as t
from transform_b200 import _lib as L
from transform_b200 import fitswcs as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# coefficients of the size found in real instrument headers (HST/ACS-like: ~50 px at the corners of 4096^2)
SIP_KEYS = {'A_ORDER': 3, 'B_ORDER': 4,
            'A_2_0': 8.2e-6, 'A_1_1': -4.7e-6, 'A_0_2': 2.1e-6, 'A_3_0': -4.4e-10, 'A_1_2': 3.3e-10, 'A_0_3': 1.2e-10,
            'B_2_0': -2.9e-6, 'B_1_1': 6.4e-6, 'B_0_2': -7.9e-6, 'B_3_0': 9.0e-11, 'B_2_1': -5.2e-10,
            'B_0_3': 4.1e-10, 'B_4_0': 1.0e-14, 'B_2_2': -2.0e-14}


def sip_header(proj='TAN', **more):
    h = {'NAXIS': 2, 'NAXIS1': 4096, 'NAXIS2': 4096, 'CTYPE1': 'RA---%s-SIP' % proj, 'CTYPE2': 'DEC--%s-SIP' % proj,
         'CRPIX1': 2048.5, 'CRPIX2': 2000.25, 'CRVAL1': 150.1, 'CRVAL2': 2.3,
         'CD1_1': -1.39e-5, 'CD1_2': 1.1e-6, 'CD2_1': 1.2e-6, 'CD2_2': 1.38e-5}
    h.update(SIP_KEYS)
    h.update(more)
    return h


def test_sip_oracle_closed_forms():
    """x' = x + A_2_0 u^2 + A_1_1 u v, y' = y + B_0_2 v^2 with u = x + 1 - CRPIX1 (pixels are 0-based here,
    CRPIX is FITS 1-based), by hand."""
    a = np.zeros((3, 3)); a[2, 0] = 1e-3; a[1, 1] = -2e-3
    b = np.zeros((3, 3)); b[0, 2] = 5e-4
    s = o.SIP([11.0, 21.0], a, b)
    p = np.array([[14.0, 17.0], [10.0, 20.0], [0.0, 0.0]])
    u, v = p[:, 0] + 1 - 11.0, p[:, 1] + 1 - 21.0
    want = np.stack((p[:, 0] + 1e-3 * u * u - 2e-3 * u * v, p[:, 1] + 5e-4 * v * v), -1)
    assert np.allclose(s.forward(p), want, rtol=0, atol=1e-13)
    assert np.array_equal(s.forward(p)[1], p[1])                       # the reference pixel does not move
    # terms beyond the order are ignored: a[2, 1] belongs to order 3
    a2 = a.copy(); a2[2, 1] = 7.0
    assert np.array_equal(o.SIP([11.0, 21.0], a2, b).forward(p), s.forward(p))
    # constant and linear terms are legal SIP terms
    a3 = np.zeros((2, 2)); a3[0, 0] = 0.5; a3[1, 0] = 0.01
    got = o.SIP([1.0, 1.0], a3, np.zeros((1, 1))).forward(np.array([[10.0, 3.0]]))
    assert np.allclose(got, [[10.0 + 0.5 + 0.1, 3.0]], atol=1e-14)


def test_sip_oracle_inverse_is_the_fixed_point():
    """reverse() = the point p with p + f(p) = v (what all_world2pix iterates to): round trips to 1e-10 pixel
    over a 4096^2 detector, and equal to an independent root finder (scipy, hybrid Powell) to 1e-9."""
    w = F.WCSParams(F.Header(sip_header()))
    s = o.SIP(w.crpix, w.sip_a, w.sip_b)
    rng = np.random.default_rng(5)
    p = rng.random((5000, 2)) * 4096 - 0.5
    f = s.forward(p)
    assert 20.0 < np.abs(f - p).max() < 200.0                          # a real distortion, not a perturbation
    assert np.abs(s.reverse(f) - p).max() < 1e-10
    assert np.abs(s.forward(s.reverse(p)) - p).max() < 1e-10
    from scipy.optimize import root
    for v in p[:12]:
        sol = root(lambda q: s.forward(q[None, :])[0] - v, v, tol=1e-13)
        assert sol.success and np.abs(sol.x - s.reverse(v[None, :])[0]).max() < 1e-9
    # NaN in, NaN out; nothing else is touched
    q = p[:4].copy(); q[1, 0] = np.nan
    r = s.reverse(q)
    assert np.isnan(r[1]).all() and not np.isnan(r[[0, 2, 3]]).any()


def test_sip_world_round_trip_and_size():
    """Through the whole pixel -> world -> pixel chain (TAN-SIP); the distortion moves world coordinates by
    the polynomial times the plate scale."""
    w = F.WCSParams(F.Header(sip_header()))
    core = o.WCS(w.crpix, w.linear_matrix(), w.crval, 'TAN', w.lonpole, w.latpole)
    full = o.WCS_SIP(core, o.SIP(w.crpix, w.sip_a, w.sip_b))
    p = np.random.default_rng(6).random((3000, 2)) * 4096
    sky = full.forward(p)
    assert np.abs(full.reverse(sky) - p).max() < 1e-7
    shift = np.hypot((sky[:, 0] - core.forward(p)[:, 0]) * np.cos(np.radians(sky[:, 1])), sky[:, 1] - core.forward(p)[:, 1])
    pix = np.hypot(*(o.SIP(w.crpix, w.sip_a, w.sip_b).forward(p) - p).T)
    assert np.allclose(shift, pix * 1.39e-5, rtol=0.12)


def test_sip_header_reader():
    w = F.WCSParams(F.Header(sip_header()))
    assert w.has_sip() and w.sip_a.shape == (4, 4) and w.sip_b.shape == (5, 5)
    assert w.sip_a[2, 0] == 8.2e-6 and w.sip_a[1, 2] == 3.3e-10 and w.sip_b[2, 2] == -2.0e-14 and w.sip_a[0, 0] == 0.0
    assert w.projection()[0] == L.WCS_PROJ_TAN                          # 'RA---TAN-SIP' is TAN
    tab = w.sip_table()
    assert tab.shape == (2, L.SIP_DIM, L.SIP_DIM) and tab[0, 2, 0] == 8.2e-6 and tab[1, 4, 0] == 1.0e-14
    assert np.count_nonzero(tab) == len(SIP_KEYS) - 2
    # astropy applies the coefficients with or without the -SIP suffix
    assert F.WCSParams(F.Header(dict(sip_header(), CTYPE1='RA---TAN', CTYPE2='DEC--TAN'))).has_sip()
    # to_header() as core.py:1208 calls it is astropy's default: standard keywords only, '-SIP' taken off CTYPE;
    # relax=True round-trips the distortion
    h0 = w.to_header()
    assert 'A_ORDER' not in h0 and h0['CTYPE1'] == 'RA---TAN' and h0['CTYPE2'] == 'DEC--TAN'
    w2 = F.WCSParams(w.to_header(relax=True))
    assert np.array_equal(w2.sip_a, w.sip_a) and np.array_equal(w2.sip_b, w.sip_b) and w2.ctype == w.ctype
    c = w.copy()
    assert c.sip_a is not w.sip_a and np.array_equal(c.sip_a, w.sip_a) and c._sip_dev == {}
    import pickle
    w._sip_dev = {'stand-in for a device tensor': object}               # the cache never travels with the object
    c = pickle.loads(pickle.dumps(w))
    assert c._sip_dev == {} and np.array_equal(c.sip_b, w.sip_b)
    w._sip_dev = {}
    with pytest.raises(ValueError):                                     # astropy: both orders or none
        F.WCSParams(F.Header({k: v for k, v in sip_header().items() if k != 'B_ORDER'}))
    with pytest.raises(ValueError):
        F.WCSParams(F.Header(dict(sip_header(), A_ORDER=10)))
    bad = F.WCSParams(F.Header(sip_header()))
    bad.sip_a = np.ones((3, 3))                                         # a_2_2 with order 2
    with pytest.raises(ValueError):
        bad.sip_table()
    for key in ('CPDIS1', 'D2IMDIS1', 'DP1'):                           # Paper IV look-up tables: loudly out of scope
        with pytest.raises(NotImplementedError):
            F.WCSParams(F.Header(dict(sip_header(), **{key: 'LOOKUP'})))
    assert not F.WCSParams(F.Header({'NAXIS': 2, 'CTYPE1': 'RA---TAN', 'CTYPE2': 'DEC--TAN'})).has_sip()


def test_foreign_wcs_object_keeps_its_distortion():
    """An astropy-like WCS object (anything with to_header) handed to DataWrapper / WCS: converted through
    to_header(relax=True) so that its SIP coefficients survive; objects whose to_header takes no argument still work."""
    from transform_b200.remap import _as_wcs
    inner = F.WCSParams(F.Header(sip_header()))

    class Relaxed:
        pixel_shape = (4096, 4096)

        def to_header(self, relax=None):
            return dict(inner.to_header(relax=relax).items())

    class Plain:
        def to_header(self):
            return dict(inner.to_header().items())

    w = _as_wcs(Relaxed())
    assert w.has_sip() and np.array_equal(w.sip_a, inner.sip_a) and w.pixel_shape == [4096, 4096]
    assert w.ctype == ['RA---TAN-SIP', 'DEC--TAN-SIP']
    w = _as_wcs(Plain())
    assert not w.has_sip() and w.ctype == ['RA---TAN', 'DEC--TAN']
    assert t.WCS(Relaxed()).params['wcs'].has_sip()


def test_projection_parameters():
    base = {'NAXIS': 2, 'CTYPE1': 'CRLN-CEA', 'CTYPE2': 'CRLT-CEA', 'CRPIX1': 100.0, 'CRPIX2': 120.0,
            'CRVAL1': 30.0, 'CRVAL2': 0.0, 'CDELT1': 0.1, 'CDELT2': 0.1}
    w = F.WCSParams(F.Header(dict(base, PV2_1=0.5, LONPOLE=10.0, PV1_3=25.0, PV1_4=80.0)))
    assert w.cea_lambda == 0.5 and w.lonpole == 25.0 and w.latpole == 80.0    # PVi_3 / PVi_4 override the keywords
    op = w.lower(L.FWD)[0]
    assert op.kind == L.OP_WCS and (op.flags & 0xff) == L.WCS_PROJ_CEA and op.p[21] == 0.5
    assert w.to_header()['PV2_1'] == 0.5
    assert F.WCSParams(F.Header(dict(base, PV2_1=1.0, PV1_0=0.0, PV1_1=0.0, PV1_2=0.0))).cea_lambda == 1.0
    for bad in ({'PV2_2': 0.3}, {'PV1_1': 5.0}, {'PV1_0': 1.0}, {'PV3_1': 2.0}):
        with pytest.raises(NotImplementedError):
            F.WCSParams(F.Header(dict(base, **bad)))
    with pytest.raises(ValueError):
        F.WCSParams(F.Header(dict(base, PV2_1=1.5)))
    # oracle: y = r0 sin(theta) / lambda (paper II eq. 80) and its inverse
    r0 = 180.0 / np.pi
    z = np.zeros(1)
    assert np.allclose(o._project("CEA", z + 30.0, z + 30.0, 0.5), [[30.0], [r0 * 0.5 / 0.5]])
    th = np.linspace(-85, 85, 41)
    x, y = o._project("CEA", th * 2, th, 0.25)
    p2, t2 = o._deproject("CEA", x, y, 0.25)
    assert np.allclose(p2, th * 2, atol=1e-10) and np.allclose(t2, th, atol=1e-10)
    ow = o.WCS(w.crpix, w.linear_matrix(), w.crval, 'CEA', w.lonpole, w.latpole, cea_lambda=0.5)
    p = np.random.default_rng(2).random((500, 2)) * 200
    assert np.abs(ow.reverse(ow.forward(p)) - p).max() < 1e-8
    assert np.abs(ow.forward(p) - o.WCS(w.crpix, w.linear_matrix(), w.crval, 'CEA', w.lonpole, w.latpole).forward(p)).max() > 0.1


def test_sip_lowering_needs_a_device_and_abi_constants():
    hdr = open(os.path.join(ROOT, "include", "tfm_b200.h")).read()
    assert int(re.search(r"TFM_OP_SIP\s*=\s*(\d+)", hdr).group(1)) == L.OP_SIP
    assert int(re.search(r"#define TFM_SIP_MAX_ORDER\s+(\d+)", hdr).group(1)) == L.SIP_MAX_ORDER
    assert int(re.search(r"#define TFM_SIP_DIM\s+(\d+)", hdr).group(1)) == L.SIP_DIM == L.SIP_MAX_ORDER + 1
    assert int(re.search(r"#define TFM_SIP_MAXITER\s+(\d+)", hdr).group(1)) == L.SIP_MAXITER == o.SIP.MAXITER
    import torch
    a = t.WCS(sip_header())
    if not torch.cuda.is_available():
        from transform_b200._device import NoCudaDevice
        with pytest.raises(NoCudaDevice):                               # the table lives on the device: no CPU path
            a.apply(np.zeros((3, 2)))
    # a SIP opcode without a coefficient table, or with an order out of range, is refused by the C ABI before
    # any device work (the in / out pointers are never dereferenced for a refused chain)
    lib = L.load()
    ops = (L.TfmOp * 1)()
    ops[0].kind, ops[0].dir = L.OP_SIP, L.FWD
    buf = (ctypes.c_double * 8)()
    v = L.TfmApplyArgs()
    v.abi_version, v.n_ops, v.chain = L.ABI_VERSION, 1, ops
    v.inp, v.out = ctypes.addressof(buf), ctypes.addressof(buf)
    v.vec_len, v.n_vec, v.dtype, v.precision = 2, 4, L.F64, L.PREC_EXACT64
    assert lib.tfm_apply(ctypes.byref(v), None) == L.EVALUE             # null table
    ops[0].p[0] = 4.94e-324                                             # some non-null bit pattern
    ops[0].p[1] = 12.0
    assert lib.tfm_apply(ctypes.byref(v), None) == L.EVALUE             # A_ORDER > 9
    ops[0].p[1], ops[0].p[2] = 3.0, -1.0
    assert lib.tfm_apply(ctypes.byref(v), None) == L.EVALUE             # B_ORDER < 0
    if not torch.cuda.is_available():
        ops[0].p[2] = 4.0                                               # well-formed: passes validation, then
        assert lib.tfm_apply(ctypes.byref(v), None) == L.ECUDA          # fails for want of a device, loudly


def _host_build_of_device_ops(tmp_path):
    """The text of the cold CUDA ops (sip_poly .. tpv_cold in csrc/tfm_chain.cuh and the slant-SIN pair), compiled
    by g++ with the CUDA qualifiers and intrinsics stubbed out.  Test infrastructure: a check of the kernels'
    arithmetic that needs no GPU; the GPU tests repeat it through the library."""
    import subprocess
    src = open(os.path.join(ROOT, "transform_b200", "csrc", "tfm_chain.cuh")).read()
    i0 = src.index("__device__ __forceinline__ double sip_poly(")
    i1 = src.index("// Internal opcode (never crosses the C ABI): a WCS pixel->world op")
    j0 = src.index("struct SlantVec {")
    j1 = src.index("template <class A, int LVL>\n__device__ __forceinline__ void op_wcs(")
    harness = r'''
#include <cmath>
#include <cstring>
#include "%s"
#define __device__
#define __forceinline__ inline
#define __noinline__
#define TFM_R2D 57.29577951308232
static inline double __ldg(const double *p) { return *p; }
static inline long long __double_as_longlong(double d) { long long b; memcpy(&b, &d, 8); return b; }
using std::fma;
%s
%s
extern "C" void run_sip(const tfm_op *op, double *xy, long n) { for (long k = 0; k < n; ++k) op_sip(*op, xy[2 * k], xy[2 * k + 1]); }
extern "C" void run_tpv(const tfm_op *op, double *xy, long n) { for (long k = 0; k < n; ++k) op_tpv(*op, xy[2 * k], xy[2 * k + 1]); }
extern "C" void run_slant(double xi, double eta, int fwd, const double *in, double *out, long n) {
    for (long k = 0; k < n; ++k) {
        if (fwd) wcs_sin_slant_deproject(xi, eta, in[3 * k], in[3 * k + 1], out[3 * k], out[3 * k + 1], out[3 * k + 2]);
        else { wcs_sin_slant_project(xi, eta, in[3 * k], in[3 * k + 1], in[3 * k + 2], out[3 * k], out[3 * k + 1]); out[3 * k + 2] = 0.0; }
    }
}
''' % (os.path.join(ROOT, "include", "tfm_b200.h"), src[i0:i1], src[j0:j1])
    cpp = tmp_path / "ops_host.cpp"
    cpp.write_text(harness)
    so = tmp_path / "ops_host.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", str(cpp), "-o", str(so)], check=True)
    lib = ctypes.CDLL(str(so))
    for f in (lib.run_sip, lib.run_tpv):
        f.argtypes = [ctypes.POINTER(L.TfmOp), ctypes.c_void_p, ctypes.c_long]
    lib.run_slant.argtypes = [ctypes.c_double, ctypes.c_double, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    return lib


def _ptr_as_double(arr):
    return np.array([arr.ctypes.data], dtype=np.int64).view(np.float64)[0]


def test_sip_device_functions_compiled_for_the_host(tmp_path):
    """op_sip against the oracle: forward 1e-11 px, inverse 1e-9 px, orders 2..9."""
    lib = _host_build_of_device_ops(tmp_path)
    for order in (2, 3, 5, 9):
        rng = np.random.default_rng(100 + order)
        hdr = {'NAXIS': 2, 'CTYPE1': 'RA---TAN-SIP', 'CTYPE2': 'DEC--TAN-SIP', 'CRPIX1': 500.5, 'CRPIX2': 520.0,
               'A_ORDER': order, 'B_ORDER': max(order - 1, 2)}
        for L_, m in (('A', order), ('B', max(order - 1, 2))):
            for p in range(m + 1):
                for q in range(m + 1 - p):
                    if p + q >= 2:
                        hdr['%s_%d_%d' % (L_, p, q)] = float(rng.normal() * 1.5 / 512.0 ** (p + q))
        w = F.WCSParams(F.Header(hdr))
        tab = np.ascontiguousarray(w.sip_table())
        s = o.SIP(w.crpix, w.sip_a, w.sip_b)
        pix = rng.random((4000, 2)) * 1024 - 0.5
        for direction in (L.FWD, L.REV):
            op = L.TfmOp()
            op.kind, op.dir = L.OP_SIP, direction
            op.p[0] = _ptr_as_double(tab)
            op.p[1], op.p[2] = float(w.sip_a.shape[0] - 1), float(w.sip_b.shape[0] - 1)
            op.p[3], op.p[4] = float(w.crpix[0]), float(w.crpix[1])
            xy = pix.copy()
            lib.run_sip(ctypes.byref(op), xy.ctypes.data, len(xy))
            if direction == L.FWD:
                assert np.abs(xy - s.forward(pix)).max() < 1e-11, order
            else:
                want = s.reverse(pix)
                ok = np.abs(s.forward(want) - pix).max(axis=1) < 1e-9          # where the iteration converges
                assert ok.mean() > 0.7 and np.abs(xy[ok] - want[ok]).max() < 1e-9, order


def tpv_header(**more):
    """A '-TPV' header of the kind SCAMP writes: a third-order polynomial with radial terms (about 0.2 % of
    the field), plus single terms of every higher degree so that all 40 slots are exercised somewhere."""
    h = {'NAXIS': 2, 'NAXIS1': 2048, 'NAXIS2': 2048, 'CTYPE1': 'RA---TPV', 'CTYPE2': 'DEC--TPV', 'CUNIT1': 'deg',
         'CUNIT2': 'deg', 'CRPIX1': 1024.5, 'CRPIX2': 1000.25, 'CRVAL1': 210.3, 'CRVAL2': -12.7,
         'CD1_1': -7.3e-5, 'CD1_2': 2.0e-6, 'CD2_1': 2.1e-6, 'CD2_2': 7.3e-5,
         'PV1_0': 1.2e-5, 'PV1_1': 1.0007, 'PV1_2': 3.0e-4, 'PV1_3': 2.0e-4, 'PV1_4': 1.1e-3, 'PV1_5': -2.0e-3,
         'PV1_6': 7.0e-4, 'PV1_7': -4.0e-2, 'PV1_8': 1.0e-2, 'PV1_9': -3.5e-2, 'PV1_10': 5.0e-3, 'PV1_11': 2.0e-2,
         'PV1_14': 0.3, 'PV1_17': 2.0, 'PV1_23': -1.0, 'PV1_27': 10.0, 'PV1_31': 50.0, 'PV1_39': -30.0,
         'PV2_0': -2.3e-5, 'PV2_1': 0.9995, 'PV2_2': -2.0e-4, 'PV2_4': -9.0e-4, 'PV2_5': 1.5e-3, 'PV2_6': -4.0e-4,
         'PV2_7': -4.1e-2, 'PV2_9': -3.9e-2, 'PV2_11': -1.0e-2, 'PV2_12': 0.2, 'PV2_16': -0.1, 'PV2_22': 1.5,
         'PV2_24': 8.0, 'PV2_30': -6.0, 'PV2_38': 40.0}
    h.update(more)
    return h


def test_tpv_oracle_and_header():
    """The TPV polynomial by hand (every kind of term: powers of the own axis, cross terms, odd powers of r, with
    the roles of xi and eta exchanged on axis 2), its inverse as a root, the header reader and the lowering order."""
    tab = np.zeros((2, 40))
    tab[0, [0, 1, 2, 3, 5, 11, 19, 39]] = [1e-3, 1.01, 2e-2, 3e-2, 0.5, 0.7, 1.1, 1.3]
    tab[1, [0, 1, 2, 6, 8, 23, 30]] = [-2e-3, 0.98, -1e-2, 0.4, 0.6, 0.9, 1.2]
    t = o.TPV(tab)
    xi, eta = 0.31, -0.17
    r = np.hypot(xi, eta)
    want = [1e-3 + 1.01 * xi + 2e-2 * eta + 3e-2 * r + 0.5 * xi * eta + 0.7 * r ** 3 + 1.1 * xi ** 3 * eta ** 2 + 1.3 * r ** 7,
            -2e-3 + 0.98 * eta - 1e-2 * xi + 0.4 * xi ** 2 + 0.6 * eta ** 2 * xi + 0.9 * r ** 5 + 1.2 * xi ** 6]
    assert np.allclose(t.forward(np.array([[xi, eta]]))[0], want, rtol=0, atol=1e-15)
    w = F.WCSParams(F.Header(tpv_header()))
    assert w.is_tpv() and w.projection()[0] == L.WCS_PROJ_TAN and w.tpv.shape == (2, 40)
    assert w.tpv[0, 1] == 1.0007 and w.tpv[0, 39] == -30.0 and w.tpv[1, 3] == 0.0 and w.tpv[1, 38] == 40.0
    assert w.lonpole is None and w.cea_lambda == 1.0            # PV1_3 is a coefficient here, not LONPOLE
    t = o.TPV(w.tpv)
    v = np.random.default_rng(9).random((4000, 2)) * 0.15 - 0.075          # the field of the header, degrees
    f = t.forward(v)
    assert 1e-4 < np.abs(f - v).max() < 2e-3                    # up to ~7 arcsec = 25 pixels
    assert np.abs(t.reverse(f) - v).max() < 1e-13
    from scipy.optimize import root
    for q in f[:10]:
        sol = root(lambda z: t.forward(z[None, :])[0] - q, q, tol=1e-15)
        assert sol.success and np.abs(sol.x - t.reverse(q[None, :])[0]).max() < 1e-12
    full = o.WCS_TPV(w.crpix, w.linear_matrix(), w.crval, w.tpv)
    plain = o.WCS(w.crpix, w.linear_matrix(), w.crval, 'TAN')
    p = np.random.default_rng(10).random((2000, 2)) * 2048
    sky = full.forward(p)
    assert np.abs(full.reverse(sky) - p).max() < 1e-7
    assert 1e-4 < np.abs(sky - plain.forward(p))[:, 1].max() < 2e-3
    # defaults: a -TPV header without PV keywords is plain TAN
    bare = F.WCSParams(F.Header({k: v for k, v in tpv_header().items() if not k.startswith('PV')}))
    assert bare.tpv is None and bare.is_tpv()
    h2 = w.to_header()
    assert h2['PV1_39'] == -30.0 and h2['PV2_1'] == 0.9995 and 'PV2_3' not in h2 and h2['CTYPE1'] == 'RA---TPV'
    assert np.array_equal(F.WCSParams(h2).tpv, w.tpv)
    with pytest.raises(ValueError):
        F.WCSParams(F.Header(tpv_header(PV1_40=1.0)))
    with pytest.raises(ValueError):
        F.WCSParams(F.Header(tpv_header(PV3_1=1.0)))
    import torch
    if not torch.cuda.is_available():
        from transform_b200._device import NoCudaDevice
        with pytest.raises(NoCudaDevice):
            t_ = __import__("transform_b200").WCS(tpv_header())
            t_.apply(np.zeros((2, 2)))
    hdr = open(os.path.join(ROOT, "include", "tfm_b200.h")).read()
    assert int(re.search(r"TFM_OP_TPV\s*=\s*(\d+)", hdr).group(1)) == L.OP_TPV
    assert int(re.search(r"#define TFM_TPV_NCOEF\s+(\d+)", hdr).group(1)) == L.TPV_NCOEF == len(o.TPV.TERMS)
    assert int(re.search(r"#define TFM_TPV_MAXITER\s+(\d+)", hdr).group(1)) == L.TPV_MAXITER == o.TPV.MAXITER


def test_tpv_and_slant_device_functions_compiled_for_the_host(tmp_path):
    """op_tpv (forward 1e-15 degree, inverse 1e-13) and the slant orthographic pair (unit-vector form on the
    device, angle form in the oracle: 1e-12) against the oracle."""
    lib = _host_build_of_device_ops(tmp_path)
    w = F.WCSParams(F.Header(tpv_header()))
    tab = np.ascontiguousarray(w.tpv)
    t = o.TPV(tab)
    Minv = np.linalg.inv(np.array([[tab[0, 1], tab[0, 2]], [tab[1, 2], tab[1, 1]]]))
    v = np.random.default_rng(11).random((5000, 2)) * 0.15 - 0.075
    for direction in (L.FWD, L.REV):
        op = L.TfmOp()
        op.kind, op.dir = L.OP_TPV, direction
        op.p[0] = _ptr_as_double(tab)
        op.p[1], op.p[2], op.p[3], op.p[4] = Minv.ravel()
        xy = v.copy()
        lib.run_tpv(ctypes.byref(op), xy.ctypes.data, len(xy))
        want = t.forward(v) if direction == L.FWD else t.reverse(v)
        assert np.abs(xy - want).max() < (1e-15 if direction == L.FWD else 1e-13)
    r0 = 180.0 / np.pi
    rng = np.random.default_rng(12)
    phi, theta = rng.random(4000) * 360 - 180, rng.random(4000) * 100 - 10
    for sl in ((0.2, -0.3), (0.0, 1.0 / np.tan(np.radians(50.0))), (-0.05, 0.0)):       # the middle one is "NCP"
        ph, th = np.radians(phi), np.radians(theta)
        m = np.ascontiguousarray(np.stack((np.cos(th) * np.cos(ph), np.cos(th) * np.sin(ph), np.sin(th)), -1))
        out = np.zeros_like(m)
        lib.run_slant(sl[0], sl[1], 0, m.ctypes.data, out.ctypes.data, len(m))
        x, y = o._project("SIN", phi, theta, 1.0, sl)
        assert np.array_equal(np.isnan(out[:, 0]), np.isnan(x)) and 0.5 < np.isnan(x).mean() * 10 < 9
        ok = ~np.isnan(x)
        assert np.abs(out[ok, 0] - x[ok]).max() < 1e-12 and np.abs(out[ok, 1] - y[ok]).max() < 1e-12
        xy = np.ascontiguousarray(np.stack((x[ok], y[ok], np.zeros(ok.sum())), -1))
        back = np.zeros_like(xy)
        lib.run_slant(sl[0], sl[1], 1, xy.ctypes.data, back.ctypes.data, len(xy))
        p2, t2 = o._deproject("SIN", x[ok], y[ok], 1.0, sl)
        got_phi = np.degrees(np.arctan2(back[:, 1], back[:, 0]))
        got_theta = np.degrees(np.arctan2(back[:, 2], np.hypot(back[:, 0], back[:, 1])))
        near = np.abs(t2 - theta[ok]) < 1e-6                     # the nearer root is the point projected
        assert near.mean() > 0.8
        assert np.abs(got_theta - t2)[near].max() < 1e-9
        assert np.abs(((got_phi - p2 + 180) % 360 - 180) * np.cos(np.radians(t2)))[near].max() < 1e-9
    assert np.allclose(o._project("SIN", phi, np.abs(theta) + 1, 1.0, (0.0, 0.0)), o._project("SIN", phi, np.abs(theta) + 1))
