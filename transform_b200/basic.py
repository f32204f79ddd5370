"""Linear / Scale / Rotation / Offset and Radial: parameter objects for the GPU opcode chain.

Host-side mirror of /root/reference/transform/basic.py:12-716 -- same constructors, the same
`params` dictionaries (pre, post, matrix, matinv, ...), the same validation errors.  The
arithmetic of _forward/_reverse (basic.py:141-176, 670-711) is not restated here: `_lower`
packs the parameters into a tfm_op and the CUDA library evaluates it.
"""
import numpy as np

from . import _lib as L
from .core import Transform
from .units import angular_coefficient


def _fingerprint(*vals):
    """Cheap identity-and-content token of a few small parameters: lowering is memoised per instance
    and direction, and redone whenever a parameter was replaced OR edited in place (the reference
    reads `params` at call time, so in-place edits must keep working)."""
    out = []
    for v in vals:
        if isinstance(v, np.ndarray):
            out.append((id(v), v.strides, v.tobytes()))
        else:
            out.append(v)
    return tuple(out)


def _vec2(v, fill):
    """1- or 2-vector -> (x, y) with the missing component neutral."""
    a = np.asarray(v, dtype=np.float64).reshape(-1)
    return (float(a[0]), float(a[1]) if a.size > 1 else fill)


class Linear(Transform):
    """out = post + matrix x (in + pre)   (basic.py:12-176).  pre, post, matrix all optional."""

    def __init__(self, *, pre=None, post=None, matrix=None, iunit=None, ounit=None,
                 itype=None, otype=None):
        idim = odim = None
        if matrix is not None:
            if not isinstance(matrix, np.ndarray):
                raise ValueError("Linear: matrix must be a 2D numpy.ndarray or None")
            if matrix.ndim != 2:
                raise ValueError("Linear: matrix must be 2D")
            odim, idim = matrix.shape
        idim = self._parse_prepost(idim, pre, 'Linear', 'Pre-offset', 'idim')
        odim = self._parse_prepost(odim, post, 'Linear', 'Post-offset', 'odim')
        if matrix is not None:
            if odim is None:
                odim = idim
            elif idim is None:
                idim = odim
            elif idim != odim:
                raise ValueError("Linear: idim and odim must match if no matrix is supplied")
        idim = 2 if idim is None else idim
        odim = 2 if odim is None else odim
        matinv = None
        if matrix is not None:
            try:
                matinv = np.linalg.inv(matrix)           # basic.py:115-119 (C-contiguous result)
            except Exception:
                matinv = None
        self.idim = idim
        self.odim = odim
        self.no_forward = False
        self.no_reverse = matinv is None and matrix is not None
        self.iunit, self.ounit, self.itype, self.otype = iunit, ounit, itype, otype
        self.params = {'pre': pre, 'post': post, 'matrix': matrix, 'matinv': matinv}

    def __str__(self):
        if not hasattr(self, '_strtmp'):
            self._strtmp = 'Linear'
        return super().__str__()

    def _parse_prepost(self, dimspec, vec, objname, prepostname, dimname):
        """basic.py:180-218."""
        if dimspec is not None and not isinstance(dimspec, np.ndarray):
            dimspec = np.array(dimspec)
        if dimspec is not None and dimspec.size > 1:
            raise ValueError("%s: %s must be a scalar" % (objname, dimname))
        if vec is not None:
            if not isinstance(vec, (np.ndarray, list, tuple)):
                raise ValueError("%s: %s must be a 1-D numpy.ndarray vector or None" % (objname, prepostname))
            if len(np.shape(vec)) != 1:
                raise ValueError("%s: %s must be a 1-D vector" % (objname, prepostname))
            if dimspec is not None and dimspec != np.shape(vec)[0]:
                raise ValueError("%s: %s size must match %s" % (objname, prepostname, dimname))
            dimspec = np.shape(vec)[0]
        return dimspec

    def _lower(self, direction):
        if self.idim != self.odim or self.idim > L.LIN_ND_MAX:
            raise NotImplementedError("transform_b200: non-square Linear transforms (or dimension > %d) "
                                      "are outside the GPU hot path" % L.LIN_ND_MAX)
        if self.idim > 2:
            return self._lower_nd(direction)
        p = self.params
        m = p['matrix'] if direction == L.FWD else p['matinv']
        fp = _fingerprint(m, p['pre'], p['post'])
        memo = self.__dict__.setdefault('_lower_memo', {})
        hit = memo.get(direction)
        if hit is not None and hit[0] == fp:
            return hit[1]
        memo[direction] = (fp, self._lower_uncached(direction, p, m))
        return memo[direction][1]

    def _lower_nd(self, direction):
        """Dimension 3..8 (basic.py:141-176): one tfm_linear_nd_op for the K3-ND kernel; only Transform.apply
        takes these (image resampling is 2-D)."""
        p = self.params
        d = self.idim
        op = L.TfmLinearNdOp()
        op.reserved = d                          # host-side note of the op's dimension (the C side ignores it)
        flags = 0
        m = p['matrix'] if direction == L.FWD else p['matinv']
        if m is not None:
            m = np.asarray(m, dtype=np.float64)
            flags |= L.LIN_HAS_MATRIX
            if not m.flags.c_contiguous:         # a transposed view (Rotation.matinv): np.matmul rounds differently
                flags |= L.LIN_MAT_FORDER
            for i in range(d):
                for j in range(d):
                    op.matrix[i * d + j] = float(m[i, j])
        # forward: +pre, matrix, +post; reverse: -post, matinv, -pre (x - p == x + (-p) exactly)
        before, after, sign = (p['pre'], p['post'], 1.0) if direction == L.FWD else (p['post'], p['pre'], -1.0)
        if before is not None:
            flags |= L.LIN_HAS_PRE
            v = np.broadcast_to(np.asarray(before, dtype=np.float64), (d,))
            for i in range(d):
                op.add_before[i] = sign * float(v[i])
        if after is not None:
            flags |= L.LIN_HAS_POST
            v = np.broadcast_to(np.asarray(after, dtype=np.float64), (d,))
            for i in range(d):
                op.add_after[i] = sign * float(v[i])
        op.flags = flags
        return [op] if flags else []

    @staticmethod
    def _lower_uncached(direction, p, m):
        op = L.TfmOp()
        op.kind, op.dir = L.OP_LINEAR, direction
        flags = 0
        if m is not None:
            m = np.asarray(m)
            flags |= L.LIN_HAS_MATRIX
            # np.matmul rounds differently for a transposed view (e.g. Rotation's matinv):
            # carry the memory order so that the fp64-exact kernel reproduces it
            if m.ndim == 2 and not m.flags.c_contiguous:
                flags |= L.LIN_MAT_FORDER
            mm = np.eye(2)
            mm[:m.shape[0], :m.shape[1]] = m
            op.p[0], op.p[1], op.p[2], op.p[3] = (float(mm[0, 0]), float(mm[0, 1]),
                                                  float(mm[1, 0]), float(mm[1, 1]))
        if p['pre'] is not None:
            flags |= L.LIN_HAS_PRE
            op.p[4], op.p[5] = _vec2(p['pre'], 0.0)
        if p['post'] is not None:
            flags |= L.LIN_HAS_POST
            op.p[6], op.p[7] = _vec2(p['post'], 0.0)
        op.flags = flags
        return [op] if flags else []


class Scale(Linear):
    """Diagonal Linear (basic.py:224-336)."""

    def __init__(self, /, scale, dim=None, *, post=None, pre=None, iunit=None, ounit=None,
                 itype=None, otype=None):
        dim = self._parse_prepost(dim, pre, 'Scale', 'Pre-offset', 'dim')
        dim = self._parse_prepost(dim, post, 'Scale', 'Post-offset', 'dim')
        if pre is not None:
            pre = pre + np.array(0)
        if post is not None:
            post = post + np.array(0)
        if not isinstance(scale, np.ndarray):
            scale = np.array(scale) if isinstance(scale, (list, tuple)) else np.array([scale])
        if scale.ndim > 1:
            raise ValueError('Scale: scale parameter must be scalar or vector')
        if dim is not None:
            dim = int(dim)
            if scale.shape[0] != dim and scale.shape[0] != 1:
                raise ValueError('Scale d must agree with size of scale vector')
            if scale.shape[0] == 1:
                scale = scale + np.zeros(dim)
        elif scale.shape[0] == 1:
            dim = 2
            scale = scale + np.zeros(2)
        else:
            dim = scale.shape[0]
        m = np.zeros([dim, dim])
        m1 = np.zeros([dim, dim])
        nonzero = all(scale != 0)
        for i in range(dim):
            m[i, i] = scale[i]
            if nonzero:
                m1[i, i] = 1.0 / scale[i]
        self.idim = dim
        self.odim = dim
        self.no_forward = False
        self.no_reverse = bool(any(scale == 0))
        self.iunit, self.ounit, self.itype, self.otype = iunit, ounit, itype, otype
        self.params = {'pre': pre, 'post': post, 'matrix': m, 'matinv': m1, 'scale': scale}

    def __str__(self):
        self._strtmp = "Linear/Scale (%s)" % (self.params['scale'],)
        return super().__str__()


class Rotation(Linear):
    """Rotation matrices from axis pairs / Euler angles (basic.py:338-506)."""

    def __init__(self, rot=None, *, post=None, pre=None, euler=None, unit='rad', iunit=None,
                 ounit=None, itype=None, otype=None):
        d_offs = self._parse_prepost(None, pre, 'Rotation', 'Pre-offset', 'd')
        d_offs = self._parse_prepost(d_offs, post, 'Rotation', 'Post-offset', 'd')
        if rot is None:
            if euler is None:
                raise ValueError("Rotation: either rot or euler angles must be specified")
            if len(euler) != 3:
                raise ValueError("Rotation: euler angles must have 3 components (axial vector)")
            rot = np.array([[0, 1, euler[2]], [2, 0, euler[1]], [1, 2, euler[0]]])
        else:
            assert euler is None, "Rotation: must specify only one of rot and euler angles"
            if not isinstance(rot, np.ndarray):
                rot = np.array(rot)
        if rot.ndim > 2:
            raise ValueError("Rotation: rot parameter must be a collection of 3-vectors")
        if rot.size == 1:
            rot = np.array([[0, 1, rot]])
        if rot.ndim == 1:
            rot = np.expand_dims(rot, 0)
        fr_axes = rot[:, 0].astype(int)
        to_axes = rot[:, 1].astype(int)
        if any(fr_axes == to_axes):
            raise ValueError('Rotation: invalid axis-to-self rotation is not allowed')
        coeff, _ = angular_coefficient(unit, "Radial")       # the reference's messages say "Radial"
        angs = rot[:, 2] * coeff
        d_fr = int(max(fr_axes.max(), to_axes.max())) + 1
        if d_offs is None:
            d = d_fr
        elif d_offs < d_fr:
            raise ValueError('Rotation: offset vectors must have at least the dims of the rotation')
        else:
            d = int(d_offs)
        out = np.eye(d)
        for i in range(fr_axes.shape[0])[::-1]:              # last listed rotation acts first
            m = np.eye(d)
            c, s = np.cos(angs[i]), np.sin(angs[i])
            m[fr_axes[i], fr_axes[i]] = c
            m[to_axes[i], to_axes[i]] = c
            m[to_axes[i], fr_axes[i]] = s
            m[fr_axes[i], to_axes[i]] = -s
            out = np.matmul(m, out)
        self.idim = d
        self.odim = d
        self.no_forward = False
        self.no_reverse = False
        self.iunit, self.ounit, self.itype, self.otype = iunit, ounit, itype, otype
        # matinv is a transposed VIEW, as in the reference (basic.py:501): its memory order
        # decides how NumPy's gemv rounds, and _lower carries that into the opcode
        self.params = {'pre': pre, 'post': post, 'matrix': out, 'matinv': out.transpose()}

    def __str__(self):
        self._strtmp = "Linear/Rotation"
        return super().__str__()


class Offset(Linear):
    """data + offset (basic.py:508-551)."""

    def __init__(self, offset):
        if not isinstance(offset, np.ndarray):
            offset = offset + np.array(0)
        if len(np.shape(offset)) > 1:
            raise ValueError("Offset: input must be a vector")
        d = np.shape(offset)[0]
        self.idim = d
        self.odim = d
        self.no_forward = False
        self.no_reverse = False
        self.iunit = self.ounit = self.itype = self.otype = None
        self.params = {'pre': offset, 'post': None, 'matrix': None, 'matinv': None}

    def __str__(self):
        self._strtmp = "Linear/Offset"
        return super().__str__()


class Radial(Transform):
    """Cartesian -> (azimuth, radius | ln radius/r0) and back (basic.py:553-716)."""

    def __init__(self, *, r0=None, iunit=None, ounit=None, itype=None, otype=None, idim=2, odim=2,
                 origin=np.zeros(2), unit='radian', ccw=False, pos_only=True):
        if not isinstance(origin, np.ndarray):
            origin = np.array([origin])                      # basic.py:620-621 (note the extra axis)
        if r0 is not None:
            otype = ["Azimuth", "Ln radius"]
            r0_sq = r0 * r0
        else:
            otype = ["Azimuth", "Radius"]
            r0_sq = None
        coeff, uname = angular_coefficient(unit, "Radial")
        if ounit is None:
            ounit = ["rad", None]                            # unit of the converted quantity (basic.py:646-647)
        elif ounit[0] is None:
            ounit[0] = "rad"
        self.idim = idim
        self.odim = odim
        self.no_forward = False
        self.no_reverse = False
        self.iunit, self.ounit, self.itype, self.otype = iunit, ounit, itype, otype
        self.params = {'origin': origin, 'r0': r0, 'r0_sq': r0_sq, 'ang_coeff': coeff,
                       'ccw': ccw, 'pos': pos_only}

    def _lower(self, direction):
        p = self.params
        if self.idim != 2 or self.odim != 2:
            raise NotImplementedError("transform_b200: Radial is on the GPU path for 2-D only")
        origin = np.asarray(p['origin'], dtype=np.float64)
        o2 = origin[0:2]                                     # basic.py:674, 708 (slices the FIRST axis)
        ox, oy = _vec2(o2.reshape(-1)[0:2], 0.0)
        op = L.TfmOp()
        op.kind, op.dir = L.OP_RADIAL, direction
        flags = 0
        if p['ccw']:
            flags |= L.RAD_CCW
        if p['pos']:
            flags |= L.RAD_POS_ONLY
        if p['r0'] is not None:
            flags |= L.RAD_R0_NOT_NONE
            if p['r0']:
                flags |= L.RAD_R0_TRUTHY
            op.p[2] = float(p['r0'])
        if not np.all(o2 == 0):
            flags |= L.RAD_ORIGIN_NONZERO
        op.flags = flags
        op.p[0], op.p[1] = ox, oy
        op.p[3] = float(p['ang_coeff'])
        return [op]

    def __str__(self):
        if not hasattr(self, '_strtmp'):
            self._strtmp = 'Radial'
        return super().__str__()


# ---------------------------------------------------------------------------------------------
# per-axis pincushion family (basic.py:846-1462)
# ---------------------------------------------------------------------------------------------
def _axis_value(v, a):
    """Component `a` of a scalar-or-vector parameter, keeping its Python / NumPy scalar type so that
    the constant folding below rounds exactly as the reference's own expressions do."""
    if np.ndim(v) == 0:
        return v
    flat = np.asarray(v).reshape(-1)
    return flat[a] if flat.size > 1 else flat[0]


class _Pincushion(Transform):
    """Shared constructor and lowering of Quadpin / Cubicpin / Poly.  The reference's `_forward`
    mutates idim/odim on first use (`self.idim, self.odim = data.shape`, basic.py:936-937), which
    only works for 2-D arrays and is not reproduced: idim = 0 means "every component"."""
    _elementwise = True
    _name = "Pincushion"
    _kind = None

    def _init(self, iunit, ounit, itype, otype, idim, odim, dim, no_reverse, params):
        if dim is not None:                                   # basic.py:912-914
            idim = dim
            odim = dim
        self.idim, self.odim = idim, odim
        self.no_forward = False
        self.no_reverse = no_reverse
        self.iunit, self.ounit, self.itype, self.otype = iunit, ounit, itype, otype
        self.params = params

    def _consts(self, direction, strength, length):
        raise NotImplementedError

    def _lower(self, direction):
        p = self.params
        ncomp = self.idim if self.idim else 2
        if ncomp > 2:
            raise NotImplementedError("transform_b200: only 1-D and 2-D transforms are on the GPU path")
        op = L.TfmOp()
        op.kind, op.dir = L.OP_PIN, direction
        flags = self._kind
        origin = p['origin']
        if not np.all(origin == 0):                           # basic.py:949, 987: shift only then
            flags |= L.PIN_SHIFTED
        if ncomp == 1:
            flags |= L.PIN_DIM1
        op.flags = flags
        op.reserved = int(p.get('degree', 0))
        op._writes_back = p['dim'] is not None                # see Transform._grid_chain
        with np.errstate(all="ignore"):
            for a in range(2):
                base = 12 * a
                op.p[base] = float(_axis_value(origin, a))
                consts = self._consts(direction, _axis_value(p['strength'], a), _axis_value(p['length'], a))
                for k, c in enumerate(consts):
                    op.p[base + 1 + k] = float(c)
        return [op]

    def __str__(self):
        if not hasattr(self, '_strtmp'):
            self._strtmp = self._name
        return super().__str__()


class Quadpin(_Pincushion):
    """Quadratic pincushion scaling per axis, with inverse (basic.py:846-1017)."""
    _name = "Quadpin"
    _kind = L.PIN_QUAD

    def __init__(self, *, iunit=None, ounit=None, itype=None, otype=None, idim=0, odim=0,
                 origin=0, length=1, strength=0.1, dim=None):
        self._init(iunit, ounit, itype, otype, idim, odim, dim, False,
                   {'origin': origin, 'length': length, 'strength': strength, 'dim': dim})

    def _consts(self, direction, strength, length):
        if direction == L.FWD:                                # basic.py:957-958
            return (strength, length, np.abs(strength) + 1)
        # basic.py:994-996
        return (4.0 * strength / length, 1.0 + np.abs(strength), length, 2.0 * strength)


class Cubicpin(_Pincushion):
    """Cubic pincushion scaling per axis (basic.py:1020-1175), as coded: the two directions use
    different normalisations (strength^2+1 vs strength+1) and are not exact inverses."""
    _name = "Cubicpin"
    _kind = L.PIN_CUBIC

    def __init__(self, *, iunit=None, ounit=None, itype=None, otype=None, idim=0, odim=0,
                 origin=0, length=1, strength=0.0, dim=None):
        self._init(iunit, ounit, itype, otype, idim, odim, dim, False,
                   {'origin': origin, 'length': length, 'strength': strength, 'dim': dim})

    def _consts(self, direction, strength, length):
        if direction == L.FWD:                                # basic.py:1117-1119
            return (strength, length, (strength * strength) + 1)
        Avar = strength / length / length                     # basic.py:1148-1153
        Betavar = 3 * Avar * 1
        try:
            coef = -1 / (3 * Avar)                            # basic.py:1170-1173
        except ZeroDivisionError:
            coef = -1 / (3 * np.nan)
        return (strength + 1, 27 * Avar * Avar, 4.0 * Betavar * Betavar * Betavar, coef)


class Poly(_Pincushion):
    """Forward-only polynomial pincushion (basic.py:1178-1362)."""
    _name = "Poly"
    _kind = L.PIN_POLY

    def __init__(self, *, iunit=None, ounit=None, itype=None, otype=None, idim=0, odim=0,
                 origin=0, length=1, strength=0.0, degree=2, dim=None, den=1):
        self._init(iunit, ounit, itype, otype, idim, odim, dim, True,
                   {'origin': origin, 'length': length, 'strength': strength, 'degree': degree,
                    'dim': dim, 'den': den})

    def _consts(self, direction, strength, length):
        degree = int(self.params['degree'])
        if degree > L.PIN_MAX_DEGREE:
            raise NotImplementedError("transform_b200: Poly degree > %d" % L.PIN_MAX_DEGREE)
        if self.params['den'] != 1:                           # basic.py:1341-1344
            sc = 1.0 / (np.abs(strength) + 1)
        else:
            sc = 1.0 / (np.abs(strength) ** (degree - 1) + 1)
        lk = [length ** (k - 1) for k in range(2, degree + 1)]                # basic.py:1347
        return (strength, sc, 0.0, 0.0) + tuple(lk)


class Quarticpin(Poly):
    """Poly with degree 4, den 1, strength 0.1 preselected (basic.py:1365-1452)."""
    _name = "Poly/Quarticpin"

    def __init__(self, *, iunit=None, ounit=None, itype=None, otype=None, idim=0, odim=0,
                 origin=0, length=1, strength=0.1, degree=4, dim=None, den=1):
        super().__init__(iunit=iunit, ounit=ounit, itype=itype, otype=otype, idim=idim, odim=odim,
                         origin=origin, length=length, strength=strength, degree=degree, dim=dim, den=den)
