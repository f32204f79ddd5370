"""oracle/oracle_np.py -- CPU restatement of the reference's resample / apply hot path.

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module; transform_b200 never does.

Every function cites the reference lines (under /root/reference/transform/) it restates.
It is a *restatement*, not an import: whole-array NumPy like the reference, but written
per tap instead of through the reference's [...,2,2,2] index expansion, so that 8192^2
inputs can be row-tiled in bounded memory (results are per-pixel independent).

Pinning (see tests/test_oracle_golden.py and tests/golden/make_golden.py):
  * nearest / linear / cubic + five boundaries, Linear/Scale/Rotation/Offset, Radial,
    simple_jacobian, jump_detect, svd2x2/svc2x2: pinned against the UNMODIFIED reference
    imported in the authoring container (golden .npz fixtures committed) and against the
    reference's own known-answer tests (tests/test_0[123]_*.py vectors restated in
    tests/test_oracle_kat.py).
  * the windowed filters of interpND (sinc, Lanczos, Gaussian, Hann, Tukey, tent with oblur) and
    the pincushion family (Quadpin, Cubicpin, Poly, Quarticpin): pinned bit-for-bit the same way,
    including NumPy's reduction orders and the integer write-back quirk of `dim=` pincushions.
  * linear WCS: pinned by tests/test_01_core.py:160-169 (sample.fits value).
  * the anti-aliased (Jacobian) resampler: the shipped interpND_grid / jacobian / interpND_jacobian
    cannot run, so the restatement here follows helpers.pyx with the
    repair list R1-R12 enumerated in `interp_grid_jacobian` below -- and is PINNED BIT FOR BIT against
    the reference's OWN code compiled with the minimal repairs R1-R6, R11
    (oracle/build_ref_repaired.py -> oracle/_ref/helpers_repaired*.so): all four usable filter shapes
    x boundaries t/e/p x jump detection on/off x oblur, golden fixtures `aa/*` and live random sweeps.
    R9 (mirror boundary) and R12 (non-square jump_detect) are the two deliberate departures and are
    outside that pin; it is also cross-checked against SVD.pyx::map_coordinates on affine maps.
  * PARITY UNPINNED: the spherical projections (TAN, CAR, STG, SIN, ARC, ZEA, CEA, MER, SFL, AIT) and
    the SIP / TPV distortions (classes SIP, TPV) -- astropy/wcslib is absent and unpinned (setup.py:27-32) -- and phot='flux' (documented at
    core.py:521-533 but not implemented by the reference).

Arithmetic conventions that the CUDA path reproduces bit-for-bit in its fp64-exact mode:
  * a 2x2 matrix-vector product is fma(m00, x, fl(m01*y)) -- what np.matmul on [...,2,1] stacks
    (basic.py:150-152) evaluates to with NumPy 2.3.5 / OpenBLAS 0.3.30 on x86-64 (measured
    bit-exact on 8000/8000 samples).  A different BLAS build may round the last bit differently;
    such differences can only flip a nearest-neighbour index within 1 ulp of a half-integer
    (DESIGN.md "near-ties").
  * nearest index = floor(v + 0.5) in float64 (helpers.pyx:127).
"""
import numpy as np

TWO_PI = 2 * np.pi          # basic.py:684 uses 2*np.pi


# --------------------------------------------------------------------------------------
# Transform algebra (core.py:218-315, 1438-1559; basic.py:141-176, 670-711)
# --------------------------------------------------------------------------------------
class Op:
    """One element of a flattened chain.  forward/reverse act on float64 [...,2] arrays."""
    idim = 2
    odim = 2

    def forward(self, v):
        raise AssertionError("abstract")       # core.py:388-390

    def reverse(self, v):
        raise AssertionError("abstract")       # core.py:417-419


_CLIB = None

# "fma_c": the defined convention (C fma, bit-reproducible everywhere) -- used for parity.
# "numpy": np.matmul on [...,N,1] stacks exactly as basic.py:150-152 executes it -- used by the
#          timed CPU-baseline legs so that the baseline costs what the reference costs.
MATVEC_IMPL = "fma_c"


def clib():
    """ctypes handle on oracle/_build/liboracle_c.so (built on demand with gcc)."""
    global _CLIB
    if _CLIB is None:
        import ctypes
        import os
        import sys
        here = os.path.dirname(os.path.abspath(__file__))
        if here not in sys.path:
            sys.path.insert(0, here)
        import build_oracle
        lib = ctypes.CDLL(build_oracle.build())
        dp = ctypes.POINTER(ctypes.c_double)
        i64 = ctypes.c_int64
        lib.orc_matvec_fma.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.c_int, dp, i64, i64,
                                       dp, i64]
        lib.orc_matvec_fma.restype = None
        lib.orc_svd2x2.argtypes = [dp, dp, dp, dp]
        lib.orc_svd2x2.restype = None
        lib.orc_svc2x2.argtypes = [dp, dp, dp, dp]
        lib.orc_svc2x2.restype = None
        lib.orc_interp_jacobian.argtypes = [dp, i64, i64, dp, dp, i64, i64,
                                            ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                            ctypes.c_double, i64, dp]
        lib.orc_interp_jacobian.restype = ctypes.c_int
        _CLIB = lib
    return _CLIB


def _dptr(a):
    import ctypes
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _matvec(m, v, f_order=False):
    """np.matmul(m, v[...,None])[...,0] (basic.py:150-152, 168-170) with a *defined* rounding that
    reproduces NumPy 2.3.5/OpenBLAS 0.3.30 on x86-64 (see oracle_c.c::orc_matvec_fma):
      C-contiguous m:            acc = fma(m[i,0], v[0], fl(m[i,1]*v[1])); acc = fma(m[i,k], v[k], acc), k = 2..
      transposed view (f_order): acc = fl(m[i,0]*v[0]); acc = fma(m[i,k], v[k], acc), k ascending (<= 3 columns;
                                 4 columns: as C-contiguous)
    pinned bit for bit for 2 x 2, 3 x 3 (both orders) and transposed 4 x 4; larger products are unpinned."""
    if MATVEC_IMPL == "numpy":
        mm = np.asarray(m, dtype=np.float64)
        if f_order and mm.flags.c_contiguous:
            mm = np.asfortranarray(mm)
        return np.matmul(mm, np.expand_dims(v, -1))[..., :, 0]
    m = np.ascontiguousarray(m, dtype=np.float64)
    vv = np.ascontiguousarray(v, dtype=np.float64)
    out = np.empty(vv.shape[:-1] + (m.shape[0],), dtype=np.float64)
    n = int(np.prod(vv.shape[:-1], dtype=np.int64))
    if n:
        clib().orc_matvec_fma(_dptr(m), m.shape[0], m.shape[1], int(bool(f_order)), _dptr(vv), n,
                              vv.shape[-1], _dptr(out), m.shape[0])
    return out


def _is_f_view(m):
    """True when the reference would hand np.matmul a transposed (non-C-contiguous) matrix."""
    return m is not None and isinstance(m, np.ndarray) and m.ndim == 2 and not m.flags.c_contiguous


class Linear(Op):
    """out = post + matrix.(in + pre)   (basic.py:141-158);
    reverse: matinv.(in - post) - pre   (basic.py:160-176).  Each piece optional."""

    def __init__(self, matrix=None, pre=None, post=None, matinv=None):
        self.matrix_f = _is_f_view(matrix)
        self.matinv_f = _is_f_view(matinv)
        self.matrix = None if matrix is None else np.array(matrix, dtype=np.float64, order="C")
        self.pre = None if pre is None else np.asarray(pre, dtype=np.float64)
        self.post = None if post is None else np.asarray(post, dtype=np.float64)
        if matinv is None and self.matrix is not None:
            matinv = np.linalg.inv(self.matrix)            # basic.py:115-119
        self.matinv = None if matinv is None else np.array(matinv, dtype=np.float64, order="C")

    def forward(self, v):
        if self.pre is not None:
            v = v + self.pre
        if self.matrix is not None:
            v = _matvec(self.matrix, v, self.matrix_f)
        if self.post is not None:
            v = v + self.post
        return v

    def reverse(self, v):
        if self.post is not None:
            v = v - self.post
        if self.matinv is not None:
            v = _matvec(self.matinv, v, self.matinv_f)
        if self.pre is not None:
            v = v - self.pre
        return v


def Scale(scale, dim=2, pre=None, post=None):
    """basic.py:309-332: diagonal matrix, matinv = diag(1/scale)."""
    s = np.zeros(dim) + np.asarray(scale, dtype=np.float64)
    return Linear(matrix=np.diag(s), matinv=np.diag(1.0 / s), pre=pre, post=post)


def Rotation(angle, unit="rad", pre=None, post=None):
    """basic.py:470-502 for the scalar (axis 0 -> axis 1) case: matinv = matrix^T."""
    coeff = {"r": 1.0, "d": np.pi / 180.0}[unit[0]]
    a = angle * coeff
    c, s = np.cos(a), np.sin(a)
    m = np.array([[c, -s], [s, c]])
    return Linear(matrix=m, matinv=m.transpose(), pre=pre, post=post)   # a VIEW, as basic.py:501


def Offset(offset):
    """basic.py:542-547: pre=offset, no matrix."""
    return Linear(pre=np.asarray(offset, dtype=np.float64))


class Radial(Op):
    """Cartesian -> (theta, r | ln r/r0) and back (basic.py:670-711)."""

    def __init__(self, origin=(0.0, 0.0), r0=None, ang_coeff=1.0, ccw=False, pos_only=True):
        self.origin = np.asarray(origin, dtype=np.float64).reshape(-1)[0:2]
        self.r0 = r0
        self.ang_coeff = float(ang_coeff)
        self.ccw = bool(ccw)
        self.pos = bool(pos_only)

    def forward(self, v):
        out = np.empty(v.shape, dtype=np.float64)
        if not np.all(self.origin == 0):                       # basic.py:674-676
            v = v - self.origin
        y = v[..., 1] if self.ccw else -v[..., 1]              # basic.py:678-681
        out[..., 0] = np.arctan2(y, v[..., 0]) / self.ang_coeff
        if self.pos:
            out[..., 0] %= TWO_PI                              # basic.py:683-684 (after unit division)
        r2 = v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]     # np.sum(data*data,-1), k ascending
        if self.r0:                                            # basic.py:686 (truthiness)
            out[..., 1] = 0.5 * np.log(r2 / (self.r0 * self.r0))
        else:
            out[..., 1] = np.sqrt(r2)
        return out

    def reverse(self, v):
        d0 = v[..., 0] * self.ang_coeff                        # basic.py:694
        out = np.empty(v.shape, dtype=np.float64)
        out[..., 0] = np.cos(d0)
        out[..., 1] = np.sin(d0) if self.ccw else -np.sin(d0)
        if self.r0 is not None:                                # basic.py:703-706
            out = out * (self.r0 * np.exp(v[..., 1, np.newaxis]))
        else:
            out = out * v[..., 1, np.newaxis]
        if not np.all(self.origin == 0):
            out = out + self.origin
        return out


class _Pin(Op):
    """Shared parameter handling of Quadpin / Cubicpin / Poly (basic.py:899-929, 1064-1094,
    1281-1315): origin, length and strength are scalars or per-axis vectors; `dim` limits the
    transform to the first `dim` components (idim = odim = dim)."""

    def __init__(self, origin=0, length=1, strength=0.0, dim=None, idim=0):
        self.origin, self.length, self.strength, self.dim = origin, length, strength, dim
        self.idim = self.odim = int(idim) if dim is None else int(dim)          # basic.py:912-914

    def _split(self, data):
        data = np.array(data)                                  # `data = data.copy()`: dtype kept
        out = data[..., :self.dim].copy() if self.dim is not None else data.copy()
        return data, out

    def _join(self, data, out):
        # Quirk kept (basic.py:962-965 and twins): with `dim` given the result is written back INTO
        # the copy of the input, so integer input -- the np.mgrid pixel grid that resample() feeds the
        # first reverse op (core.py:574-578) -- truncates the transformed coordinates to int64.
        # Without `dim` (e.g. idim=2, odim=2) the float result is returned as is.
        if self.dim is not None:
            with np.errstate(invalid="ignore"):
                data[..., :self.dim] = out
            return data
        return out


class Quadpin(_Pin):
    """Quadratic pincushion, per axis (basic.py:846-1017)."""

    def __init__(self, origin=0, length=1, strength=0.1, dim=None, idim=0):
        super().__init__(origin, length, strength, dim, idim)

    def forward(self, v):                                      # basic.py:932-967
        data, out = self._split(v)
        origin, strength, length = self.origin, self.strength, self.length
        shifted = not np.all(origin == 0)
        if shifted:
            origin = np.array([origin])
            out = out - origin
        out = out + strength * (out * np.abs(out)) / length
        out = out / (np.abs(strength) + 1)
        if shifted:
            out = out + origin
        return self._join(data, out)

    def reverse(self, v):                                      # basic.py:970-1010
        data, out = self._split(v)
        origin, strength, length = self.origin, self.strength, self.length
        od = out.copy()
        shifted = not np.all(origin == 0)
        if shifted:
            origin = np.array([origin])
            od = out - origin
        cmp = (data[..., :self.dim] if self.dim is not None else data) < origin   # `data < origin`, unshifted
        with np.errstate(invalid="ignore"):
            out = -1.0 + np.sqrt(1.0 + 4.0 * strength / length * np.abs(od) * (1.0 + np.abs(strength)))
            out = out * length / (2.0 * strength)
            out = out * (1.0 - 2.0 * cmp)
        if shifted:
            out = out + origin
        return self._join(data, out)


class Cubicpin(_Pin):
    """Cubic pincushion, per axis (basic.py:1020-1175).  As coded: the forward divides by
    strength^2 + 1 where the reverse multiplies by strength + 1, and the reverse evaluates BOTH
    Cardano roots (the first `try` dies on an unbound name after computing the negative branch;
    the bare `except` computes the positive one) -- reproduced, not repaired."""

    def __init__(self, origin=0, length=1, strength=0.0, dim=None, idim=0):
        super().__init__(origin, length, strength, dim, idim)

    def forward(self, v):                                      # basic.py:1097-1128
        data, out = self._split(v)
        origin, strength, length = self.origin, self.strength, self.length
        shifted = not np.all(origin == 0)
        if shifted:
            origin = np.array([origin])
            out = out - origin
        dl0 = out / length
        out = out + strength * out * dl0 * dl0
        out = out / ((strength * strength) + 1)
        if shifted:
            out = out + origin
        return self._join(data, out)

    def reverse(self, v):                                      # basic.py:1130-1170
        data, out = self._split(v)
        origin, strength, length = self.origin, self.strength, self.length
        shifted = not np.all(origin == 0)
        if shifted:
            origin = np.array([origin])
            out = out - origin
        out = out * (strength + 1)
        A = strength / length / length
        D = -out
        alpha = 27 * A * A * D
        beta = 3 * A * 1
        with np.errstate(all="ignore"):
            inner = np.sqrt(alpha * alpha + 4.0 * beta * beta * beta)
            an = 0.5 * (alpha - inner)
            cn = (1 - 2 * (an < 0)) * (abs(an) ** (1 / 3))
            ap = 0.5 * (alpha + inner)
            cp = (1 - 2 * (ap < 0)) * (abs(ap) ** (1 / 3))
            cuberoot = cp + cn
            try:
                out = (-1 / (3 * A)) * cuberoot
            except ZeroDivisionError:
                out = (-1 / (3 * np.nan)) * cuberoot
        if shifted:
            out = out + origin
        return self._join(data, out)


class Poly(_Pin):
    """Forward-only polynomial pincushion (basic.py:1178-1362); Quarticpin = Poly(degree=4, strength=0.1)."""

    def __init__(self, origin=0, length=1, strength=0.0, degree=2, dim=None, den=1, idim=0):
        super().__init__(origin, length, strength, dim, idim)
        self.degree, self.den = degree, den

    def forward(self, v):                                      # basic.py:1318-1358
        data, out = self._split(v)
        origin, strength, length = self.origin, self.strength, self.length
        shifted = not np.all(origin == 0)
        if shifted:
            origin = np.array([origin])
            out = out - origin
        a = out
        if self.den != 1:
            sc = 1.0 / (np.abs(strength) + 1)
        else:
            sc = 1.0 / (np.abs(strength) ** (self.degree - 1) + 1)
        for k in range(2, self.degree + 1):
            a = a + strength * (out ** k) / (length ** (k - 1))
        out = sc * a
        if shifted:
            out = out + origin
        return self._join(data, out)

    def reverse(self, v):
        raise AssertionError("Poly is invalid in the reverse direction")


class Identity(Op):
    idim = 0
    odim = 0

    def forward(self, v):
        return v

    def reverse(self, v):
        return v


class Inverse(Op):
    """core.py:1478-1482."""

    def __init__(self, t):
        self.t = t

    def forward(self, v):
        return self.t.reverse(v)

    def reverse(self, v):
        return self.t.forward(v)


class Composition(Op):
    """core.py:1551-1559: forward walks the list right-to-left, reverse left-to-right calling
    each member's inverse; nested compositions are flattened at construction (1534-1537)."""

    def __init__(self, ops):
        flat = []
        for o in ops:
            flat.extend(o.ops if isinstance(o, Composition) else [o])
        self.ops = flat

    def forward(self, v):
        for o in self.ops[::-1]:
            v = o.forward(v)
        return v

    def reverse(self, v):
        for o in self.ops:
            v = o.reverse(v)
        return v


def _first_leaf_is_dim_pin(t, invert):
    """Is the first leaf op that `t` applies (in the given direction) a pincushion op with `dim`?"""
    while True:
        if isinstance(t, Inverse):
            t, invert = t.t, not invert
        elif isinstance(t, Composition):
            t = t.ops[0] if invert else t.ops[-1]
        else:
            return isinstance(t, _Pin) and t.dim is not None


def linear_dim(t):
    """Dimension of a transform made of Linear ops only (None otherwise): the size of its matrices / offsets."""
    if isinstance(t, Linear):
        for a in (t.matrix, t.matinv):
            if a is not None:
                return int(a.shape[0])
        for a in (t.pre, t.post):
            if a is not None and np.ndim(a) == 1:
                return int(np.shape(a)[0])
        return None
    if isinstance(t, Inverse):
        return linear_dim(t.t)
    if isinstance(t, Composition):
        dims = [linear_dim(x) for x in t.ops]
        if not dims or any(d is None for d in dims):
            return None
        return max(dims)
    return None


def apply(t, data, invert=False):
    """Transform.apply for 2-D transforms (core.py:218-294): extra trailing components are
    passed through unchanged (core.py:271-276, 287-292)."""
    data = np.asarray(data)
    d = data.astype(np.float64)
    if data.dtype.kind in "iu" and _first_leaf_is_dim_pin(t, invert):
        d = data                                 # integer input reaches a pincushion op with `dim`: see _Pin._join
    fn = t.reverse if invert else t.forward
    n = 2
    if isinstance(t, Identity):
        return d
    nd = linear_dim(t)                           # Linear-family chains of dimension 3..8 (basic.py:141-176)
    if nd is not None and nd > 2:
        n = nd
    if isinstance(t, _Pin):                      # idim = dim, or 0 = every component (core.py:271-292)
        n = t.idim
        if n == 0:
            return fn(d)
    if d.shape[-1] < n:
        raise ValueError("requires %d dimensions; data have %d" % (n, d.shape[-1]))
    if d.shape[-1] > n:
        return np.append(fn(d[..., 0:n]), d[..., n:], axis=-1)
    return fn(d)


# --------------------------------------------------------------------------------------
# Boundary conditions and taps (helpers.pyx:34-181, 183-376)
# --------------------------------------------------------------------------------------
def _bound_chars(bound, ndim=2):
    """helpers.pyx:113-120 and 137: scalar or per-axis list in (X,Y) order; first char only."""
    if not isinstance(bound, (tuple, list)):
        bound = [bound] * ndim
    elif len(bound) == 1:
        bound = [bound[0]] * ndim
    if len(bound) != ndim:
        raise ValueError("apply_boundary: boundary list must match vec dims")
    return [b[0] for b in bound]


def boundary_index(i, size, b):
    """Integer index -> boundary-conditioned index for one axis (helpers.pyx:136-176).
    't' marks out-of-range with -1; 'f' raises."""
    i = np.asarray(i, dtype=np.int64)
    if b == "f":
        if not np.all((i >= 0) & (i < size)):
            raise ValueError("apply_boundary: boundary violation with 'forbid' condition")
        return i
    if b == "t":
        return np.where((i < 0) | (i >= size), -1, i)
    if b == "e":
        return np.clip(i, 0, size - 1)
    if b == "p":
        return i % size                                        # Python-style (floored) modulo
    if b == "m":
        j = i % (2 * size)
        return np.where(j >= size, (2 * size - 1) - j, j)      # helpers.pyx:170-172
    raise ValueError("apply_boundary: boundaries are 'f', 't', 'e', 'p', or 'm'.")


def _trunc_pass_applies(bound):
    """helpers.pyx:373: `any(map(lambda s: s[0]=='t', bound))` -- a bare string is iterated by
    character, so 'extend' and 'truncate' qualify while 'e' does not.  When it applies, the
    np.where at helpers.pyx:374 also promotes the gathered region with the fill array's dtype
    (np.array(0) is int64, so a float32 region becomes float64 *before* any arithmetic)."""
    seq = bound if isinstance(bound, (tuple, list)) else list(bound)
    return any(s[0] == "t" for s in seq)


def _tap(source, ix, iy, bx, by, fillvalue, trunc_pass):
    """source[...,iy,ix] with boundary conditions (helpers.pyx:356-374).  Returned in the
    reference's *region dtype*: the source dtype, unless the truncation pass promoted it."""
    h, w = source.shape[-2], source.shape[-1]
    jx = boundary_index(ix, w, bx)
    jy = boundary_index(iy, h, by)
    val = source[jy, jx]                          # negative (-1) indexes wrap; overwritten next
    if trunc_pass:
        val = np.where((jx < 0) | (jy < 0), np.array(fillvalue), val)
    return val


I64_MIN = np.iinfo(np.int64).min


def _cast_i64(f):
    """`.astype('int')` of an integer-valued float64 array as x86-64 performs it
    (helpers.pyx:127): NaN, +-inf and |f| >= 2^63 all give INT64_MIN (cvttsd2si "integer
    indefinite").  Spelled out so the oracle is warning-free and platform independent; the
    CUDA path implements the same rule."""
    f = np.asarray(f, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        ok = np.abs(f) < 2.0 ** 63
    return np.where(ok, np.where(ok, f, 0.0).astype(np.int64), I64_MIN)


def _tap_index(fl, k):
    """Index of tap k along one axis: the reference adds the chunk offset in floating point
    (helpers.pyx:347-352) and only then rounds with floor(v+0.5) (helpers.pyx:127)."""
    with np.errstate(invalid="ignore"):
        return _cast_i64(np.floor((fl + k) + 0.5))


def region_dtype(src_dtype, bound, fillvalue=0):
    """dtype of gathered taps (and of the nearest-neighbour result): see _trunc_pass_applies."""
    if _trunc_pass_applies(bound):
        return np.result_type(src_dtype, np.array(fillvalue).dtype)
    return np.dtype(src_dtype)


def interp_nearest(source, index, bound="t", fillvalue=0):
    """helpers.pyx:541-549 -> sampleND: idx = floor(v+0.5) (helpers.pyx:127)."""
    bx, by = _bound_chars(bound)
    index = np.asarray(index, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        ix = _cast_i64(np.floor(index[..., 0] + 0.5))
        iy = _cast_i64(np.floor(index[..., 1] + 0.5))
    return _tap(source, ix, iy, bx, by, fillvalue, _trunc_pass_applies(bound))


def interp_linear(source, index, bound="t", fillvalue=0):
    """helpers.pyx:559-611: 2x2 region at floor(v), each corner boundary-conditioned
    independently; weight = prod(alpha or 1-alpha); sum(region*weight)."""
    bx, by = _bound_chars(bound)
    index = np.asarray(index, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        fx = np.floor(index[..., 0])
        fy = np.floor(index[..., 1])
        ax = index[..., 0] - fx
        ay = index[..., 1] - fy
    ix0, ix1 = _tap_index(fx, 0), _tap_index(fx, 1)
    iy0, iy1 = _tap_index(fy, 0), _tap_index(fy, 1)
    bxw = 1 - ax
    byw = 1 - ay
    # numpy reduces region*weight over axes (-1,-2) of the contiguous 2x2 block sequentially:
    # ((r00*w00 + r01*w01) + r10*w10) + r11*w11 with x fastest.  The CUDA fp64 path uses this order.
    tp = _trunc_pass_applies(bound)
    r00 = _tap(source, ix0, iy0, bx, by, fillvalue, tp)
    r01 = _tap(source, ix1, iy0, bx, by, fillvalue, tp)
    r10 = _tap(source, ix0, iy1, bx, by, fillvalue, tp)
    r11 = _tap(source, ix1, iy1, bx, by, fillvalue, tp)
    return ((r00 * (bxw * byw) + r01 * (ax * byw)) + r10 * (bxw * ay)) + r11 * (ax * ay)


def _hermite(r0, r1, r2, r3, b):
    """helpers.pyx:755-778: cubic Hermite with central-difference slopes, in the reference's
    operation order.  NOTE the dtype flow: r* arrive in the region dtype, so for a float32 source
    whose region was not promoted (see _trunc_pass_applies) the differences d, s0, s1 and the two
    bracketed combinations are evaluated in float32; the products with b (float64) promote."""
    a0 = r1
    d = r2 - r1
    s0 = (r2 - r0) * 0.5
    s1 = (r3 - r1) * 0.5
    return a0 + b * (s0 + b * ((3 * d - 2 * s0 - s1) + b * (s1 + s0 - 2 * d)))


def interp_cubic(source, index, bound="t", fillvalue=0):
    """helpers.pyx:703-787: 4x4 region at floor(v)-1, X axis collapsed first, then Y."""
    bx, by = _bound_chars(bound)
    index = np.asarray(index, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        fx = np.floor(index[..., 0])
        fy = np.floor(index[..., 1])
        b_x = index[..., 0] - fx
        b_y = index[..., 1] - fy
    ixs = [_tap_index(fx, k) for k in range(-1, 3)]
    tp = _trunc_pass_applies(bound)
    rows = []
    for dy in range(-1, 3):
        iy = _tap_index(fy, dy)
        t = [_tap(source, ix, iy, bx, by, fillvalue, tp) for ix in ixs]
        rows.append(_hermite(t[0], t[1], t[2], t[3], b_x))
    return _hermite(rows[0], rows[1], rows[2], rows[3], b_y)


FILTER_SIZE = {"l": 2, "s": 16, "z": 6, "g": 8, "h": 2, "t": 2}       # helpers.pyx:708


def filter_window(m, of):
    """Window value at offset `of` (in units of oblur-scaled input pixels), helpers.pyx:809-826."""
    if m == "s":
        return np.sinc(of)
    if m == "z":
        bb = np.clip(of, -3, 3)
        return 3 * np.sinc(bb) * np.sinc(bb / 3)
    if m == "g":
        return np.exp(-of * of / 0.5 / 0.5)
    if m == "h":
        return 1 + np.cos(np.pi * np.clip(of, -1, 1))
    if m == "t":
        return 1 + np.cos(np.pi * np.clip(np.abs(of) * 2 - 0.5, 0, 1))
    if m == "l":
        return 1 - np.clip(np.abs(of), 0, 1)
    raise AssertionError("This can't happen")


def interp_filter(source, index, method, bound="t", fillvalue=0, oblur=None):
    """The collapsible input-plane filters of interpND (helpers.pyx:703-735, 792-835): sinc, Lanczos,
    Gaussian, Hann, Tukey and the tent with oblur.  A size x size region at floor(v) - (size/2 - 1) is
    gathered with per-tap boundary conditions, then collapsed one axis at a time, X first:
        region <- sum(k * region, axis) / sum(k, axis),   k = window((tap - sample)/oblur)."""
    bx, by = _bound_chars(bound)
    m = method[0]
    index = np.asarray(index, dtype=np.float64)
    size = FILTER_SIZE[m]
    if oblur is not None:
        size = int(2 * np.ceil(size / 2 * oblur))                          # helpers.pyx:711
    else:
        oblur = 1.0
    offset = int(size / 2) - 1
    with np.errstate(invalid="ignore"):
        fx, fy = np.floor(index[..., 0]), np.floor(index[..., 1])
        b_x, b_y = index[..., 0] - fx, index[..., 1] - fy
    tp = _trunc_pass_applies(bound)
    ixs = [_tap_index(fx - offset, k) for k in range(size)]
    # of = ((1 - b) - (offset + 1) + i) / oblur    (helpers.pyx:800-802)
    kx = [filter_window(m, ((1 - b_x) - (offset + 1) + i) / oblur) for i in range(size)]
    ky = [filter_window(m, ((1 - b_y) - (offset + 1) + i) / oblur) for i in range(size)]
    # Summation orders, pinned bit-for-bit against the compiled reference: the first-axis kernel `k`
    # inherits the transposed-mgrid memory order, so its normaliser k.sum(axis=-1) runs along a
    # strided axis (sequential adds); the other three reductions run along a contiguous axis
    # (NumPy's 8-accumulator pairwise order, see _npsum).
    kxs = _seqsum(kx)
    rows = []
    with np.errstate(invalid="ignore"):
        for j in range(size):
            iy = _tap_index(fy - offset, j)
            taps = [_tap(source, ix, iy, bx, by, fillvalue, tp) for ix in ixs]
            rows.append(_npsum([k * t for k, t in zip(kx, taps)]) / kxs)
        return _npsum([k * r for k, r in zip(ky, rows)]) / _npsum(ky)


def _seqsum(terms):
    acc = terms[0]
    for t in terms[1:]:
        acc = acc + t
    return acc


def _npsum(terms):
    """Sum of a list of arrays in the order NumPy's reduction over a contiguous last axis uses
    (pairwise_sum in umath: sequential below 8 terms; 8 interleaved accumulators, combined as
    ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the remainder sequentially, up to 128 terms)."""
    n = len(terms)
    if n < 8:
        acc = terms[0]
        for t in terms[1:]:
            acc = acc + t
        return acc
    r = list(terms[0:8])
    i = 8
    while i + 8 <= n:
        for k in range(8):
            r[k] = r[k] + terms[i + k]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    for t in terms[i:]:
        res = res + t
    return res


def interpND(source, index, method="n", bound="t", fillvalue=0, oblur=None):
    """2-D subset of helpers.pyx:379-844: nearest, linear, cubic and the collapsible filters
    (everything but the Fourier method)."""
    m = method[0]
    source = np.asarray(source)
    if m == "n":
        if oblur is not None:
            raise ValueError("interpND: oblur is not allowed with nearest-neighbor interpolation.")
        return interp_nearest(source, index, bound, fillvalue)
    if m == "l" and oblur is None:
        return interp_linear(source, index, bound, fillvalue)
    if m == "c":
        if oblur is not None:
            raise ValueError("interpND: oblur is not allowed with cubic-spline interpolation.")
        return interp_cubic(source, index, bound, fillvalue)
    if m in FILTER_SIZE:
        return interp_filter(source, index, m, bound, fillvalue, oblur)
    raise ValueError("interpND: valid methods are ('n','l','c','f','s','z','g','h','t')")


# --------------------------------------------------------------------------------------
# resample driver (core.py:422-586)
# --------------------------------------------------------------------------------------
def pixel_grid(ny, nx, y0=0, y1=None):
    """core.py:574-578: np.mgrid[range(NX),range(NY)].transpose() -> [NY,NX,2], [y,x,:]=(x,y).
    Rows y0:y1 only (exact tiling: pixels are independent)."""
    y1 = ny if y1 is None else y1
    g = np.empty((y1 - y0, nx, 2), dtype=np.float64)
    g[..., 0] = np.arange(nx, dtype=np.float64)[None, :]
    g[..., 1] = np.arange(y0, y1, dtype=np.float64)[:, None]
    return g


def _grid_for(t, ny, nx, y0=0, y1=None):
    """The grid as the first reverse op sees it: np.mgrid is int64 (core.py:574-578), which matters
    only to a pincushion op with `dim` (it writes its result back into the integer array)."""
    g = pixel_grid(ny, nx, y0, y1)
    return g.astype(np.int64) if _first_leaf_is_dim_pin(t, True) else g


def resample(t, data, method="n", bound="t", shape=None, fillvalue=0, rows=None, tile=256, oblur=None):
    """Transform.resample (core.py:422-586) for 2-D data.  `rows=(y0,y1)` restricts the output to
    a row block (used for row-tiled / multi-process CPU baselines)."""
    data = np.asarray(data)
    if shape is None:
        shape = data.shape
    ny, nx = int(shape[-2]), int(shape[-1])
    y0, y1 = (0, ny) if rows is None else rows
    m = method[0]
    out = None
    for ya in range(y0, y1, tile):
        yb = min(y1, ya + tile)
        coords = t.reverse(_grid_for(t, ny, nx, ya, yb))
        blk = interpND(data, coords, method=m, bound=bound, fillvalue=fillvalue, oblur=oblur)
        if out is None:
            out = np.empty((y1 - y0, nx), dtype=blk.dtype)
        out[ya - y0:yb - y0] = blk
    if out is None:
        dt = region_dtype(data.dtype, bound, fillvalue) if m == "n" else np.float64
        out = np.empty((0, nx), dtype=dt)
    return out


# --------------------------------------------------------------------------------------
# Jacobian estimation (helpers.pyx:846-958, 960-1141, 1145-1290), 2-D branches
# --------------------------------------------------------------------------------------
def simple_jacobian(index):
    """helpers.pyx:896-906: cell-centred finite differences, output [NY-1,NX-1,2,2],
    J[...,i,j] = d coord_i / d pix_j (j=0: x direction, j=1: y direction)."""
    I = np.asarray(index, dtype=np.float64)
    if I.ndim != 3 or I.shape[-1] != 2:
        raise ValueError("simple_jacobian (oracle): 2-D grids of 2-vectors only")
    if I.shape[0] < 2 or I.shape[1] < 2:
        raise ValueError("SimpleJacobian: must have at least two vectors along each grid axis")
    a, b, c, d = I[1:, 1:], I[1:, :-1], I[:-1, 1:], I[:-1, :-1]
    J = np.empty((I.shape[0] - 1, I.shape[1] - 1, 2, 2))
    J[..., :, 0] = (a - b + c - d) / 2
    J[..., :, 1] = (a + b - c - d) / 2
    return J


def _kth_of_stack(stack, k):
    s = np.sort(np.stack(stack, axis=-1), axis=-1)
    return s[..., k]


def jump_detect(Js, jump_thresh=10):
    """helpers.pyx:1012-1141, N-D branch evaluated for N=2.
    flag = Jmag2 <= (k-th smallest of the neighbourhood) * jump_thresh, with
      interior: 3x3 neighbourhood, k = int(9/3+0.5) = 3          (1045-1072)
      edges   : 2x3 neighbourhood (edge line + the next one), k = int(6/3+0.5) = 2 (1079-1139)
      corners : left at 1.
    R12 (deviation from the text): the upper-edge target slice is written with shape[-1]
    (helpers.pyx:1136) where shape[0] is meant; identical for square grids, and for non-square
    grids the shipped line raises (broadcast into an empty slice).  Restated with shape[0]."""
    Js = np.asarray(Js, dtype=np.float64)
    m2 = (Js * Js).sum(axis=(-2, -1))
    ny, nx = m2.shape
    flag = np.ones((ny, nx), dtype=np.int64)
    if ny >= 3 and nx >= 3:
        lags = [m2[a:a + ny - 2, b:b + nx - 2] for a in range(3) for b in range(3)]
        flag[1:-1, 1:-1] = m2[1:-1, 1:-1] <= _kth_of_stack(lags, 3) * jump_thresh
    if nx >= 3 and ny >= 2:
        lo = [m2[a:a + 1, b:b + nx - 2] for a in range(2) for b in range(3)]
        flag[0:1, 1:-1] = m2[0:1, 1:-1] <= _kth_of_stack(lo, 2) * jump_thresh
        hi = [m2[a + ny - 2:a + ny - 1, b:b + nx - 2] for a in range(2) for b in range(3)]
        flag[ny - 1:ny, 1:-1] = m2[ny - 1:ny, 1:-1] <= _kth_of_stack(hi, 2) * jump_thresh
    if ny >= 3 and nx >= 2:
        lo = [m2[b:b + ny - 2, a:a + 1] for a in range(2) for b in range(3)]
        flag[1:-1, 0:1] = m2[1:-1, 0:1] <= _kth_of_stack(lo, 2) * jump_thresh
        hi = [m2[b:b + ny - 2, a + nx - 2:a + nx - 1] for a in range(2) for b in range(3)]
        flag[1:-1, nx - 1:nx] = m2[1:-1, nx - 1:nx] <= _kth_of_stack(hi, 2) * jump_thresh
    return flag


def jacobian(index, jump_thresh=4.0):
    """helpers.pyx:1145-1290, 2-D branch (1225-1247), with repairs
      R1  `return J` (the shipped function falls off the end),
      R2  the `jump_detect` keyword shadows the function (1146 vs 1206): un-shadowed, and the
          threshold reaches jump_detect() as jump_thresh (interpND_grid passes it at 1498),
      R11 `J[1:-1,1:-1,...].T /= wgt.clip(1)` (1239) raises; restated as a broadcast divide.
    jump_thresh == 0/False/None disables jump detection (helpers.pyx:1459)."""
    I = np.asarray(index, dtype=np.float64)
    Js = simple_jacobian(I)
    ny, nx = I.shape[0], I.shape[1]
    J = np.empty((ny, nx, 2, 2))
    if jump_thresh:
        jf = jump_detect(Js, jump_thresh)
        Js = np.where(jf[..., None, None] != 0, Js, 0.0)                   # 1207
    if ny >= 3 and nx >= 3:
        acc = Js[:-1, :-1] + Js[1:, :-1] + Js[:-1, 1:] + Js[1:, 1:]        # 1227-1231
        if jump_thresh:
            wgt = jf[:-1, :-1] + jf[1:, :-1] + jf[:-1, 1:] + jf[1:, 1:]    # 1234-1238
            acc = acc / np.clip(wgt, 1, None)[..., None, None]
        else:
            acc = acc / 4
        J[1:-1, 1:-1] = acc
    J[0] = J[1]                                                              # 1244-1247
    J[-1] = J[-2]
    J[:, 0] = J[:, 1]
    J[:, -1] = J[:, -2]
    return J


# --------------------------------------------------------------------------------------
# 2x2 SVD (helpers.pyx:1845-1916) -- NumPy restatement; the C twin lives in oracle_c.c
# --------------------------------------------------------------------------------------
def svd2x2(M):
    """Returns U, s, V with M = U diag(s) V^T (helpers.pyx:1845-1894)."""
    a, b, c, d = M[0][0], M[0][1], M[1][0], M[1][1]
    th = 0.5 * np.arctan2(2 * (a * c + b * d), (a * a + b * b) - (c * c + d * d))
    ph = 0.5 * np.arctan2(2 * (a * b + c * d), (a * a + c * c) - (b * b + d * d))
    st, ct, sp, cp = np.sin(th), np.cos(th), np.sin(ph), np.cos(ph)
    cs00 = (ct * a + st * c) * cp + (ct * b + st * d) * sp
    cs11 = (st * a - ct * c) * sp - (st * b - ct * d) * cp
    c00 = 1.0 if cs00 >= 0 else -1.0
    c11 = 1.0 if cs11 >= 0 else -1.0
    U = np.array([[ct, -st], [st, ct]])
    V = np.array([[cp * c00, -sp * c11], [sp * c00, cp * c11]])
    return U, np.array([abs(cs00), abs(cs11)]), V


def svc2x2(U, s, V):
    """M = U diag(s) V^T (helpers.pyx:1901-1916)."""
    u00, u10, u01, u11 = U[0][0] * s[0], U[1][0] * s[0], U[0][1] * s[1], U[1][1] * s[1]
    return np.array([[u00 * V[0][0] + u01 * V[0][1], u00 * V[1][0] + u01 * V[1][1]],
                     [u10 * V[0][0] + u11 * V[0][1], u10 * V[1][0] + u11 * V[1][1]]])


# --------------------------------------------------------------------------------------
# Anti-aliased (Jacobian-adaptive) resampling: interpND_grid (helpers.pyx:1293-1530) driving
# interpND_jacobian (helpers.pyx:1542-1803, C restatement in oracle_c.c)
# --------------------------------------------------------------------------------------
_METHOD_CODE = {"l": 1, "z": 2, "h": 3, "t": 4, "g": 5,          # helpers.pyx:1501
                "w": 6}                                          # the Hann window of SVD.pyx:110-111
_BOUND_CODE = {"f": 1, "t": 2, "e": 3, "p": 4, "m": 5}           # helpers.pyx:1511


def interp_jacobian(source, index, J, method="g", bound="t", fillvalue=0, oblur=1.0, iblur=1.0,
                    pad_pow=8, max_reg=0, phot="radiance"):
    """interpND_jacobian (helpers.pyx:1542-1803) via oracle_c.orc_interp_jacobian.
    phot='flux' (documented at core.py:521-533, not implemented there): scale by |det J| of the
    unpadded Jacobian, the rule of the reference's other anti-aliaser (SVD.pyx:285-286)."""
    try:
        mcode = _METHOD_CODE[method[0]]
    except KeyError:
        raise ValueError("interpND_grid: method must be 'l','z','h','t', or 'g'")
    try:
        bx, by = (_BOUND_CODE[c] for c in _bound_chars(bound))
    except KeyError:
        raise ValueError("interpND_grid: each bound must be 'f','t','e','p', or 'm'")
    src = np.ascontiguousarray(source, dtype=np.float64)               # helpers.pyx:1520
    idx = np.ascontiguousarray(index, dtype=np.float64)
    Jc = np.ascontiguousarray(J, dtype=np.float64)
    ny, nx = idx.shape[0], idx.shape[1]
    dest = np.empty((ny, nx), dtype=np.float64)
    rc = clib().orc_interp_jacobian(_dptr(src), src.shape[0], src.shape[1], _dptr(idx), _dptr(Jc),
                                    ny, nx, mcode, bx, by, float(fillvalue), float(oblur),
                                    float(iblur), float(pad_pow), int(max_reg), _dptr(dest))
    if rc == 2:
        raise ValueError("value out of bounds with Forbid boundary in effect")
    if rc != 0:
        raise ValueError("interpND_jacobian: invalid interpolation method")
    if phot[0].lower() in ("f", "e"):
        dest = dest * np.abs(Jc[..., 0, 0] * Jc[..., 1, 1] - Jc[..., 0, 1] * Jc[..., 1, 0])
    return dest


def interp_grid_jacobian(source, index, method="g", bound="t", fillvalue=0, oblur=1.0, iblur=1.0,
                         pad_pow=8, jump_detect=4.0, max_reg=0, phot="radiance"):
    """interpND_grid with antialias=True (helpers.pyx:1480-1530), 2-D only.

    The shipped code cannot run.  Repairs applied, each a named
    deviation from the text of helpers.pyx:
      R1  jacobian(): add the missing `return J` (after 1290)
      R2  jacobian(): the `jump_detect` kwarg shadows the function of that name (1146/1206);
          un-shadowed, and interpND_grid's float threshold (1498) reaches it as jump_thresh
      R3  interpND_grid: return the result of interpND_jacobian (1520-1530 drops it)
      R4  interpND_jacobian: delete `assert(0)` (1640)
      R5  reg_size is an integer; `-reg_size/2 .. reg_size/2` (1678-1679) with reg_size even
      R6  bound codes passed as C ints (1517 builds int64, 1546 declares int[:])
      R7  kept as coded: a truncated tap adds its filter weight but not its value (1740 vs
          1793-1796), so edges fade to 0 and `fillvalue` is unused
      R8  kept as coded: s_i <- (padval + s_i^pad_pow)^(-1/pad_pow), padval 1 (l,h,t) or 3 (z,g)
          (1612-1615, 1661-1662) -- the docstring formula at 1434-1435 differs; code wins
      R9  mirror boundary: 2s-1-i as apply_boundary (172), not s-1-i (1768-1769, 1789-1790)
      R10 the judged filter is the Gaussian exp(-(xf^2+yf^2)) (1734-1735)
      R11 jacobian(): `J[1:-1,1:-1,...].T /= wgt.clip(1)` (1239) raises; broadcast divide instead
      R12 jump_detect(): upper-edge slice uses shape[0] where 1136 says shape[-1] (non-square)
    `max_reg` (0 = off) is an explicit cap on reg_size standing in for the unused sv_limit."""
    source = np.asarray(source)
    index = np.asarray(index, dtype=np.float64)
    if index.shape[-1] != source.ndim:
        raise ValueError("interpND_grid: source shape must match index vector length")
    if source.ndim != index.ndim - 1:
        raise ValueError("interpND_grid: index broadcast dims must match source dims")
    if index.shape[-1] != 2:
        raise ValueError("interpND_grid: anti-aliasing is only implemented for 2-D at present.")
    J = jacobian(index, jump_thresh=jump_detect)
    return interp_jacobian(source, index, J, method, bound, fillvalue, oblur, iblur, pad_pow,
                           max_reg, phot)


def resample_jacobian(t, data, method="g", bound="t", shape=None, rows=None, **kw):
    """What Transform.resample would do for the capital-letter methods (core.py:495-508,
    581-584) once dispatched to interpND_grid: grid -> inverse chain -> Jacobian filter.
    `rows=(y0,y1)`: compute only that row block, but with the Jacobian of the FULL grid
    semantics (2 extra rows of coordinates on each side, edge rules by global position)."""
    data = np.asarray(data)
    if shape is None:
        shape = data.shape
    ny, nx = int(shape[-2]), int(shape[-1])
    if rows is None:
        coords = t.reverse(pixel_grid(ny, nx))
        return interp_grid_jacobian(data, coords, method=method, bound=bound, **kw)
    y0, y1 = rows
    ya, yb = max(0, y0 - 3), min(ny, y1 + 3)
    coords = t.reverse(pixel_grid(ny, nx, ya, yb))
    full = _jacobian_rows(coords, ya, yb, ny, kw.get("jump_detect", 4.0))
    sub = slice(y0 - ya, y1 - ya)
    kw2 = {k: v for k, v in kw.items() if k != "jump_detect"}
    return interp_jacobian(data, coords[sub], full[sub], method, bound, **kw2)


def _jacobian_rows(coords, ya, yb, ny, jump_thresh):
    """jacobian() of a row block [ya,yb) of a grid with ny rows: identical to slicing the full-grid
    result as long as the block carries >= 3 rows of margin on interior sides (the stencil reaches
    2 rows; edge rules only fire on true grid edges)."""
    if ya == 0 and yb == ny:
        return jacobian(coords, jump_thresh)
    J = jacobian(coords, jump_thresh)
    # rows whose edge rules fired spuriously (block edges that are not grid edges) are inside the
    # margin and are discarded by the caller; nothing to fix up.
    return J


# --------------------------------------------------------------------------------------
# FITS WCS (core.py:1719-1747 -> astropy.wcs.WCS.all_pix2world / all_world2pix, origin 0)
# --------------------------------------------------------------------------------------
_R0 = 180.0 / np.pi


def _deproject(proj, x, y, lam=1.0, slant=(0.0, 0.0)):
    """Projection-plane (x, y) [deg] -> native (phi, theta) [deg]; Calabretta & Greisen 2002, in the
    paper's own angle form (the CUDA path works on unit vectors: the two derivations cross-check)."""
    with np.errstate(all="ignore"):
        if proj == "SIN" and tuple(slant) != (0.0, 0.0):                     # slant orthographic, eq. 61-62
            # xi-eta form of paper II section 5.1.5: with X = x / r0, Y = y / r0 and s = 1 - sin(theta),
            # (xi^2 + eta^2 + 1) s^2 - 2 (X xi + Y eta + 1) s + X^2 + Y^2 = 0; the root nearer the pole
            xi, eta = slant
            X, Y = x / _R0, y / _R0
            a, b, c = xi * xi + eta * eta + 1.0, X * xi + Y * eta + 1.0, X * X + Y * Y
            s = c / (b + np.sqrt(b * b - a * c))                             # = (b - sqrt) / a; NaN beyond the limb
            theta = 90.0 - 2.0 * np.degrees(np.arcsin(np.sqrt(s / 2.0)))     # 1 - sin(theta) = s, stable at the pole
            phi = np.degrees(np.arctan2(X - xi * s, -(Y - eta * s)))
            return phi, theta
        if proj in ("TAN", "STG", "SIN", "ARC", "ZEA"):                      # zenithal, eq. 14-15
            R = np.hypot(x, y)
            phi = np.where(R == 0, 0.0, np.degrees(np.arctan2(x, -y)))
            if proj == "TAN":
                theta = np.degrees(np.arctan2(_R0, R))                       # eq. 55
            elif proj == "STG":
                theta = 90.0 - 2.0 * np.degrees(np.arctan(R / (2.0 * _R0)))  # eq. 57
            elif proj == "SIN":
                theta = np.degrees(np.arccos(R / _R0))                       # eq. 60 (NaN beyond the limb)
            elif proj == "ARC":
                theta = np.where(R <= 180.0, 90.0 - R, np.nan)               # eq. 68
            else:
                theta = 90.0 - 2.0 * np.degrees(np.arcsin(R / (2.0 * _R0)))  # eq. 70
            return phi, theta
        if proj == "CAR":
            return x, y                                                      # eq. 83-84
        if proj == "CEA":
            return x, np.degrees(np.arcsin(lam * y / _R0))                   # eq. 81 (lambda = PV2_1, default 1)
        if proj == "MER":
            return x, 2.0 * np.degrees(np.arctan(np.exp(y / _R0))) - 90.0    # eq. 86
        if proj == "SFL":
            c = np.cos(np.radians(y))
            phi = np.where(c > 0, x / np.where(c > 0, c, 1.0), np.where(x == 0, 0.0, np.nan))   # eq. 88
            phi = np.where((np.abs(y) <= 90.0) & (np.abs(phi) <= 180.0 * (1 + 1e-12)), phi, np.nan)
            return phi, y
        if proj == "AIT":
            z2 = 1.0 - (x / (4.0 * _R0)) ** 2 - (y / (2.0 * _R0)) ** 2       # eq. 109-110
            z2 = np.where(z2 >= 0.5 * (1 - 1e-12), z2, np.nan)
            Z = np.sqrt(z2)
            phi = 2.0 * np.degrees(np.arctan2((x / (2.0 * _R0)) * Z, 2.0 * z2 - 1.0))
            theta = np.degrees(np.arcsin((y / _R0) * Z))
            return phi, theta
    raise ValueError(proj)


def _project(proj, phi, theta, lam=1.0, slant=(0.0, 0.0)):
    """Native (phi, theta) [deg] -> projection-plane (x, y) [deg]."""
    ph, th = np.radians(phi), np.radians(theta)
    with np.errstate(all="ignore"):
        if proj == "SIN" and tuple(slant) != (0.0, 0.0):                     # eq. 61-62
            xi, eta = slant
            vis = theta >= -np.degrees(np.arctan(xi * np.sin(ph) - eta * np.cos(ph)))   # wcslib's bounds check
            x = _R0 * (np.cos(th) * np.sin(ph) + xi * (1.0 - np.sin(th)))
            y = -_R0 * (np.cos(th) * np.cos(ph) - eta * (1.0 - np.sin(th)))
            return np.where(vis, x, np.nan), np.where(vis, y, np.nan)
        if proj in ("TAN", "STG", "SIN", "ARC", "ZEA"):
            if proj == "TAN":
                R = np.where(theta > 0, _R0 / np.tan(th), np.nan)            # eq. 54
            elif proj == "STG":
                R = np.where(theta > -90, 2.0 * _R0 * np.tan(np.radians(90.0 - theta) / 2.0), np.nan)
            elif proj == "SIN":
                R = np.where(theta >= 0, _R0 * np.cos(th), np.nan)           # eq. 59
            elif proj == "ARC":
                R = 90.0 - theta                                             # eq. 67
            else:
                R = np.where(theta > -90, 2.0 * _R0 * np.sin(np.radians(90.0 - theta) / 2.0), np.nan)
            return R * np.sin(ph), -R * np.cos(ph)                           # eq. 12-13
        if proj == "CAR":
            return phi, theta
        if proj == "CEA":
            return phi, _R0 * np.sin(th) / lam                               # eq. 80
        if proj == "MER":
            return phi, np.where(np.abs(theta) < 90, _R0 * np.log(np.tan(np.radians(90.0 + theta) / 2.0)), np.nan)
        if proj == "SFL":
            return phi * np.cos(th), theta
        if proj == "AIT":
            g = np.sqrt(2.0 / (1.0 + np.cos(th) * np.cos(ph / 2.0)))         # eq. 108
            return 2.0 * _R0 * g * np.cos(th) * np.sin(ph / 2.0), _R0 * g * np.sin(th)
    raise ValueError(proj)


class WCS(Op):
    """pixel (origin 0) <-> world for two axes.

    The reference delegates to astropy/wcslib, absent from /root/reference and unpinned
    (setup.py:27-32).  Restated from Greisen & Calabretta 2002 (paper I: linear part) and
    Calabretta & Greisen 2002 (paper II: spherical rotation eq. 2 and 5, TAN eq. 54-55, CAR
    eq. 83-84, LONPOLE/LATPOLE defaults and the celestial pole, eq. 8-10).
      * linear part: PINNED by tests/test_01_core.py:160-169 (sample.fits, CD matrix).
      * TAN / CAR:   PARITY UNPINNED (nothing in the reference's tests exercises them).
    Deliberately written with the explicit Euler-angle formulas per point -- not with the 3x3
    rotation matrix the CUDA path uses -- so that the two derivations check each other.
    proj: None (linear axes), 'TAN', 'CAR', or one of STG SIN ARC ZEA CEA MER SFL AIT (default
    projection parameters, except CEA's lambda = PV2_1).  cd = CDELT_i * PC_ij."""

    def __init__(self, crpix, cd, crval, proj=None, lonpole=None, latpole=None, cea_lambda=1.0, sin_slant=(0.0, 0.0)):
        self.cea_lambda = float(cea_lambda)
        self.sin_slant = (float(sin_slant[0]), float(sin_slant[1]))
        self.crpix = np.asarray(crpix, dtype=np.float64)
        self.cd = np.asarray(cd, dtype=np.float64)
        self.cdinv = np.linalg.inv(self.cd)
        self.crval = np.asarray(crval, dtype=np.float64)
        self.proj = proj
        if proj is not None:
            phi0, theta0 = (0.0, 90.0) if proj in ("TAN", "STG", "SIN", "ARC", "ZEA") else (0.0, 0.0)
            a0, d0 = self.crval
            self.phi_p = lonpole if lonpole is not None else phi0 + (0.0 if d0 >= theta0 else 180.0)
            theta_p = 90.0 if latpole is None else latpole
            if theta0 == 90.0:
                self.alpha_p, self.delta_p = a0, d0
            else:
                dphi = np.radians(self.phi_p - phi0)
                t0, dd = np.radians(theta0), np.radians(d0)
                a = np.degrees(np.arctan2(np.sin(t0), np.cos(t0) * np.cos(dphi)))
                b = np.degrees(np.arccos(np.clip(
                    np.sin(dd) / np.sqrt(1.0 - (np.cos(t0) * np.sin(dphi)) ** 2), -1, 1)))
                cands = [c for c in (a + b, a - b) if -90.0 - 1e-10 <= c <= 90.0 + 1e-10]
                self.delta_p = float(np.clip(min(cands, key=lambda c: abs(c - theta_p)), -90, 90))
                dp = np.radians(self.delta_p)
                if abs(abs(d0) - 90.0) < 1e-12:
                    self.alpha_p = a0
                elif abs(self.delta_p - 90.0) < 1e-12:
                    self.alpha_p = a0 + self.phi_p - phi0 - 180.0
                elif abs(self.delta_p + 90.0) < 1e-12:
                    self.alpha_p = a0 - self.phi_p + phi0
                else:
                    self.alpha_p = a0 - np.degrees(np.arctan2(
                        np.sin(dphi) * np.cos(t0) / np.cos(dd),
                        (np.sin(t0) - np.sin(dp) * np.sin(dd)) / (np.cos(dp) * np.cos(dd))))

    def forward(self, v):
        v = np.asarray(v, dtype=np.float64)
        d = (v + 1.0) - self.crpix                                # FITS pixels are 1-based
        ix = self.cd[0, 0] * d[..., 0] + self.cd[0, 1] * d[..., 1]
        iy = self.cd[1, 0] * d[..., 0] + self.cd[1, 1] * d[..., 1]
        out = np.empty(v.shape, dtype=np.float64)
        if self.proj is None:
            out[..., 0] = ix + self.crval[0]
            out[..., 1] = iy + self.crval[1]
            return out
        phi, theta = _deproject(self.proj, ix, iy, self.cea_lambda, self.sin_slant)
        ph, th, dp = np.radians(phi - self.phi_p), np.radians(theta), np.radians(self.delta_p)
        # eq. (2); the latitude is taken with atan2(z, hypot(x, y)) instead of asin(z) -- the
        # well-conditioned form wcslib switches to near the poles (sphx2s)
        xx = np.sin(th) * np.cos(dp) - np.cos(th) * np.sin(dp) * np.cos(ph)
        yy = -np.cos(th) * np.sin(ph)
        zz = np.sin(th) * np.sin(dp) + np.cos(th) * np.cos(dp) * np.cos(ph)
        lon = self.alpha_p + np.degrees(np.arctan2(yy, xx))
        lat = np.degrees(np.arctan2(zz, np.hypot(xx, yy)))
        lon = np.where(lon < 0, lon + 360.0, lon)
        lon = np.where(lon >= 360.0, lon - 360.0, lon)
        out[..., 0] = lon
        out[..., 1] = lat
        return out

    def reverse(self, v):
        v = np.asarray(v, dtype=np.float64)
        if self.proj is None:
            ix = v[..., 0] - self.crval[0]
            iy = v[..., 1] - self.crval[1]
        else:
            al, de, dp = np.radians(v[..., 0] - self.alpha_p), np.radians(v[..., 1]), np.radians(self.delta_p)
            xx = np.sin(de) * np.cos(dp) - np.cos(de) * np.sin(dp) * np.cos(al)      # eq. (5)
            yy = -np.cos(de) * np.sin(al)
            zz = np.sin(de) * np.sin(dp) + np.cos(de) * np.cos(dp) * np.cos(al)
            phi = self.phi_p + np.degrees(np.arctan2(yy, xx))
            theta = np.degrees(np.arctan2(zz, np.hypot(xx, yy)))
            phi = np.where(phi > 180.0, phi - 360.0, phi)
            phi = np.where(phi <= -180.0, phi + 360.0, phi)
            ix, iy = _project(self.proj, phi, theta, self.cea_lambda, self.sin_slant)
        px = self.cdinv[0, 0] * ix + self.cdinv[0, 1] * iy
        py = self.cdinv[1, 0] * ix + self.cdinv[1, 1] * iy
        out = np.empty(v.shape, dtype=np.float64)
        out[..., 0] = (px + self.crpix[0]) - 1.0
        out[..., 1] = (py + self.crpix[1]) - 1.0
        return out


class SIP(Op):
    """SIP polynomial distortion (Shupe et al. 2005, "The SIP convention for representing distortion
    in FITS image headers"): the pix2foc stage of astropy's all_pix2world / all_world2pix, which
    core.py:1727 and 1742 call.  PARITY UNPINNED (astropy absent from /root/reference and from
    this image); restated from the convention and from the documented algorithm of all_world2pix.

    a[p, q] = A_p_q, b[p, q] = B_p_q (terms with p + q above the order are ignored), crpix 1-based.
      forward (pixel, origin 0 -> distortion-corrected pixel):  u = x + 1 - CRPIX1, v = y + 1 - CRPIX2,
              x' = x + sum A_p_q u^p v^q,  y' = y + sum B_p_q u^p v^q
      reverse: all_world2pix's fixed-point iteration pix <- pix0 - f(pix) on the FORWARD polynomials
              (AP_/BP_ are not used), run to convergence per point (step < 1e-12 pixel, at most 20
              evaluations = astropy's maxiter).  astropy stops when the worst point of the batch
              moves less than 1e-4 pixel, so it agrees with this only to that tolerance.
    Written as an explicit sum of powers -- the CUDA op is a nested Horner form -- so that the
    two cross-check."""

    MAXITER = 20

    def __init__(self, crpix, a, b):
        self.crpix = np.asarray(crpix, dtype=np.float64)
        self.a = np.asarray(a, dtype=np.float64)
        self.b = np.asarray(b, dtype=np.float64)

    @staticmethod
    def _poly(c, u, v):
        m = c.shape[0] - 1
        out = np.zeros(np.broadcast(u, v).shape, dtype=np.float64)
        for p in range(m + 1):
            for q in range(m + 1 - p):
                if c[p, q] != 0.0:
                    out = out + c[p, q] * np.power(u, p) * np.power(v, q)
        return out

    def _f(self, v):
        uu = (v[..., 0] + 1.0) - self.crpix[0]
        vv = (v[..., 1] + 1.0) - self.crpix[1]
        return np.stack((self._poly(self.a, uu, vv), self._poly(self.b, uu, vv)), axis=-1)

    def forward(self, v):
        v = np.asarray(v, dtype=np.float64)
        return v + self._f(v)

    def reverse(self, v):
        pix0 = np.asarray(v, dtype=np.float64)
        pix = pix0.copy()
        active = np.ones(pix.shape[:-1], dtype=bool)
        with np.errstate(all="ignore"):
            for _ in range(self.MAXITER):
                new = pix0 - self._f(pix)
                d2 = np.sum((new - pix) ** 2, axis=-1)
                pix = np.where(active[..., None], new, pix)
                active = active & (d2 >= 1e-24)
                if not active.any():
                    break
        return pix


def WCS_SIP(wcs, sip):
    """all_pix2world / all_world2pix of a header with a SIP distortion: forward = SIP then the core
    WCS, reverse = core WCS then the SIP inversion."""
    return Composition([wcs, sip])


class TPV(Op):
    """The TPV polynomial distortion ('RA---TPV' / 'DEC--TPV' headers, PV1_m / PV2_m, m = 0..39; the
    convention of the FITS WCS registry written by SCAMP): wcslib applies it to the intermediate world
    coordinates (xi, eta) [deg] between the linear part and the gnomonic projection, inside the wcsp2s /
    wcss2p that astropy's all_pix2world / all_world2pix call (core.py:1727, 1742).  PARITY UNPINNED.

    tab[0][m] = PV1_m, tab[1][m] = PV2_m.  With r = sqrt(xi^2 + eta^2):
      xi'  = PV1_0 + PV1_1 xi + PV1_2 eta + PV1_3 r + PV1_4 xi^2 + PV1_5 xi eta + PV1_6 eta^2 + PV1_7 xi^3 + ...
             + PV1_11 r^3 + ... + PV1_23 r^5 + ... + PV1_39 r^7
      eta' = the same with the roles of xi and eta exchanged and PV2_m.
    Written as an explicit list of (power of own axis, power of other axis, power of r) -- the CUDA op
    accumulates each degree in Horner fashion -- so that the two cross-check.  reverse: Newton-like iteration on
    the constant linear part, to convergence (wcslib iterates to its own tolerance on the same root)."""

    TERMS = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1),
             (2, 0, 0), (1, 1, 0), (0, 2, 0),
             (3, 0, 0), (2, 1, 0), (1, 2, 0), (0, 3, 0), (0, 0, 3),
             (4, 0, 0), (3, 1, 0), (2, 2, 0), (1, 3, 0), (0, 4, 0),
             (5, 0, 0), (4, 1, 0), (3, 2, 0), (2, 3, 0), (1, 4, 0), (0, 5, 0), (0, 0, 5),
             (6, 0, 0), (5, 1, 0), (4, 2, 0), (3, 3, 0), (2, 4, 0), (1, 5, 0), (0, 6, 0),
             (7, 0, 0), (6, 1, 0), (5, 2, 0), (4, 3, 0), (3, 4, 0), (2, 5, 0), (1, 6, 0), (0, 7, 0), (0, 0, 7)]
    MAXITER = 30

    def __init__(self, tab):
        self.tab = np.asarray(tab, dtype=np.float64)
        assert self.tab.shape == (2, 40) and len(self.TERMS) == 40

    def _axis(self, c, own, other):
        r = np.hypot(own, other)
        out = np.zeros(np.broadcast(own, other).shape)
        for m, (a, b, k) in enumerate(self.TERMS):
            if c[m] != 0.0:
                out = out + c[m] * np.power(own, a) * np.power(other, b) * np.power(r, k)
        return out

    def forward(self, v):
        v = np.asarray(v, dtype=np.float64)
        return np.stack((self._axis(self.tab[0], v[..., 0], v[..., 1]),
                         self._axis(self.tab[1], v[..., 1], v[..., 0])), axis=-1)

    def reverse(self, v):
        t = np.asarray(v, dtype=np.float64)
        M = np.array([[self.tab[0, 1], self.tab[0, 2]], [self.tab[1, 2], self.tab[1, 1]]])
        Mi = np.linalg.inv(M)
        x = (t - np.array([self.tab[0, 0], self.tab[1, 0]])) @ Mi.T
        active = np.ones(x.shape[:-1], dtype=bool)
        with np.errstate(all="ignore"):
            for _ in range(self.MAXITER):
                step = (self.forward(x) - t) @ Mi.T
                x = np.where(active[..., None], x - step, x)
                active = active & (np.sum(step * step, axis=-1) >= 1e-26)
                if not active.any():
                    break
        return x


def WCS_TPV(crpix, cd, crval, tab, lonpole=None, latpole=None):
    """A '-TPV' header: linear part -> TPV polynomial -> TAN -> celestial rotation."""
    lin = WCS(crpix, cd, [0.0, 0.0], None)
    prj = WCS([1.0, 1.0], np.eye(2), crval, "TAN", lonpole, latpole)
    return Composition([prj, TPV(tab), lin])
