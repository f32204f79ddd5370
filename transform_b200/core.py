"""Transform base class and the structural transforms (Identity, Inverse, Composition, Wrap).

Host-side mirror of /root/reference/transform/core.py for the resample / remap / apply path:
same class names, constructor arguments, attribute names (idim, odim, no_forward, no_reverse,
params, ...), keyword arguments and exceptions -- but no NumPy arithmetic.  A Transform here is a
*description*; `apply`, `invert` and `resample` lower it to a flat, direction-resolved opcode
chain (include/tfm_b200.h: tfm_op) and hand it to the CUDA library.  There is no CPU fallback.
"""
import copy

import numpy as np

from . import _engine
from . import _lib as L


class Transform:
    """Composable, invertible coordinate transform (reference: core.py:16-586).

    Subclasses describe themselves through `_lower(direction)`, which returns the list of
    opcodes (execution order) that realises `_forward` (direction FWD) or `_reverse` (REV).
    """

    def __init__(self):
        # core.py:165-168
        raise AssertionError("generic transforms must be subclassed (e.g. Transform.identity)")

    # -- stringification protocol of the reference (core.py:172-214) --
    def __str__(self):
        try:
            s = self._strtmp
        except AttributeError:
            return "Generic Transform stringifier - you should never see this"
        del self._strtmp
        flag = self.__dict__.pop("_str_not_top_tmp", False)
        return s if flag else "Transform( %s )" % s

    # -- lowering --
    def _lower(self, direction):
        # core.py:388-390 / 417-419: the abstract _forward/_reverse raise AssertionError
        name = "_forward" if direction == L.FWD else "_reverse"
        raise AssertionError("Transform.%s should always be overloaded by a subclass." % name)

    def _chain(self, invert):
        """Validity checks of apply() (core.py:264-266, 280-282) + lowering."""
        if invert:
            if self.no_reverse:
                raise AssertionError("This %s is invalid in the reverse direction" % self.__str__())
            return self._lower(L.REV)
        if self.no_forward:
            raise AssertionError("This Transform (%s) is invalid in the forward direction" % self.__str__())
        return self._lower(L.FWD)

    def _int_chain(self, invert):
        """Chain as INTEGER input sees it -- the np.mgrid pixel grid of resample() (core.py:574-578)
        or an integer array given to apply().  Only one thing differs from `_chain`: a pincushion op
        constructed with `dim=` that comes first writes its result back into the integer copy of its
        input (basic.py:962-965 and twins), i.e. truncates it to int64; reproduced by an opcode flag."""
        ops = self._chain(invert)
        if ops and isinstance(ops[0], L.TfmOp) and ops[0].kind == L.OP_PIN and getattr(ops[0], "_writes_back", False):
            ops[0].flags |= L.PIN_TRUNC
        return ops

    # -- vector path (core.py:218-315) --
    def apply(self, data, invert=False, *, precision=None):
        """Apply the transform to [..., N] vectors (K3 streaming kernel).

        Same contract as the reference: the last axis is the vector; components beyond the
        transform's dimension are passed through unchanged (core.py:271-276, 287-292).
        `precision`: 'fp64' (default, parity mode) or 'fp32' (fast path; float32 in/out)."""
        is_dev = _engine.dev.is_cuda_tensor(data)
        if not is_dev and not isinstance(data, np.ndarray):
            data = np.array(data)
        int_input = (not is_dev) and data.dtype.kind in "iu"
        ops = self._int_chain(invert) if int_input else self._chain(invert)
        dim = self.odim if invert else self.idim
        nlast = int(data.shape[-1]) if len(data.shape) else 0
        if len(data.shape) == 0 or nlast < dim:
            raise ValueError("This %s requires %d dimensions; data have %d" % (self.__str__(), dim, nlast))
        if dim > 2:
            # Linear-family chains of dimension 3..8 (Linear / Scale / Offset / Rotation from axis pairs or Euler
            # angles, basic.py:141-176, 338-506): the K3-ND kernel
            if not all(isinstance(o, L.TfmLinearNdOp) and o.reserved == dim for o in ops):
                raise NotImplementedError("transform_b200: beyond 2-D only Linear-family transforms of one "
                                          "dimension are on the GPU path")
            if not ops:
                return data
            return _engine.apply_linear_nd(ops, dim, data)
        if not ops:
            return data                         # Identity._forward returns its argument (core.py:1381-1385)
        if dim == 0 and getattr(self, "_elementwise", False) and nlast != 2:
            # idim = 0: the transform acts on every component (core.py:274-276, 290-292); an
            # elementwise op with per-axis-identical parameters runs on the flattened data as pairs
            if is_dev:
                raise NotImplementedError("transform_b200: idim=0 transforms on device tensors")
            if any(np.ndim(self.params[k]) != 0 for k in ('origin', 'length', 'strength')):
                raise ValueError("vector parameters need vectors of matching length (2 on the GPU path)")
            arr = np.asarray(data, dtype=np.float64)
            flat = arr.reshape(-1)
            if flat.size % 2:
                flat = np.concatenate([flat, np.zeros(1)])
            res = _engine.apply(ops, flat.reshape(-1, 2), precision=precision)
            return res.reshape(-1)[:arr.size].reshape(arr.shape)
        if nlast == 1:                          # 1-D transform on 1-vectors: pad a dummy component
            if is_dev:
                raise NotImplementedError("transform_b200: 1-vectors on device tensors")
            pad = np.concatenate([np.asarray(data, dtype=np.float64), np.zeros(data.shape)], axis=-1)
            return _engine.apply(ops, pad, precision=precision)[..., 0:1]
        return _engine.apply(ops, data, precision=precision)

    def invert(self, data, invert=False, **kw):
        """core.py:298-315."""
        return self.apply(data, invert=not invert, **kw)

    def composition(self, target=None):
        """core.py:318-341."""
        if isinstance(target, (list, tuple)):
            return Composition([self] + list(target))
        return Composition([self, target])

    def inverse(self):
        """core.py:344-361."""
        return Inverse(self)

    # -- image path (core.py:422-586) --
    def resample(self, data, /, method='n', bound='t', phot='radiance', shape=None, *,
                 precision=None, fillvalue=0, tuning=0, out=None, sync=True, **aa):
        """Resample gridded data through the transform (reference: core.py:422-586).

        The output grid is enumerated implicitly, mapped back through the inverse chain in
        registers and interpolated in ONE CUDA kernel.  `method` (first character):
        'n'earest, 'l'inear, 'c'ubic and the input-plane filters 's'inc, 'z' (Lanczos),
        'g'aussian, 'h'ann, 't'ukey use interpND semantics (helpers.pyx:541-611, 703-835);
        the capital letters the reference advertises but never dispatches (core.py:495-508,
        581-584) -- 'G'aussian, 'H'anning, 'R'ounded (Tukey), 'Z'Lanczos -- run the
        Jacobian-adaptive anti-aliasing kernel with interpND_grid's keyword arguments
        (oblur, iblur, pad_pow, sv_limit, jump_detect) taken from **aa.

        `phot` ('radiance' default / 'flux'): the reference documents but never implements it
        (core.py:521-533) and silently resamples intensively.  Here the anti-aliased methods honour
        'flux' by scaling each output pixel by |det J| (as SVD.pyx:285-286 does); the plain
        interpolators ignore it like the reference.

        Extra keyword-only arguments (not in the reference): precision ('fp64' parity mode,
        default; 'fp32' fast path), fillvalue (interpND's, default 0), out (a CUDA tensor to
        write into, or a host array / pinned CPU tensor that receives the result), sync (False
        leaves the copy into a pinned `out` in flight; used by pipeline.AsyncResampler)."""
        if isinstance(data, np.ndarray) or _engine.dev.is_tensor(data):
            data0 = data
        elif hasattr(data, 'data') and isinstance(data.data, np.ndarray):
            data0 = data.data
        else:
            raise ValueError('Transform.map requires a numpy array or an object containing one')
        if not isinstance(method, str) or not method:
            raise ValueError("interpND: valid methods are ('n','l','c','f','s','z','g','h','t')")
        if shape is None:
            shape = tuple(data0.shape)
        shape = tuple(int(s) for s in shape)
        # core.py:566-571
        if len(shape) < self.odim:
            raise ValueError('map: Transform odim must match output data shape')
        if len(data0.shape) < self.idim:
            raise ValueError('map: Transform idim must match input data shape')
        if len(data0.shape) - self.idim != len(shape) - self.odim:
            raise ValueError('map: shape and source dimensions must match')
        if self.idim > 2 or self.odim > 2:
            raise NotImplementedError("transform_b200: only 1-D and 2-D transforms are on the GPU path")
        if len(shape) < 2:
            raise NotImplementedError("transform_b200: the GPU path resamples 2-D planes "
                                      "(optionally batched over leading axes)")
        ops = self._int_chain(True)             # self.invert(grid), core.py:574 (the grid is int64)
        m = method[0]
        status_host = aa.pop('status_host', None)
        # interpND_grid's two spellings of the anti-aliasing switch (helpers.pyx:1297-1298); both are
        # consumed whatever their values
        antialias = bool(aa.pop('antialias', True)) & bool(aa.pop('aa', True))
        if m in _engine._AA_METHODS:
            if antialias:
                return _engine.resample_jacobian(ops, data0, shape, _engine._AA_METHODS[m], bound,
                                                 fillvalue=fillvalue, precision=precision,
                                                 tuning=tuning, out=out, phot=phot, sync=sync,
                                                 status_host=status_host, **aa)
            # anti-aliasing switched off: the capital letter falls back to its input-plane filter, as
            # interpND_grid(antialias=False) does (helpers.pyx:1524-1530); only oblur carries over
            m = _engine._AA_METHODS[m]
            for k in ('iblur', 'pad_pow', 'sv_limit', 'jump_detect', 'max_reg'):
                aa.pop(k, None)
        oblur = aa.pop('oblur', None)           # interpND's keyword (helpers.pyx:509-516)
        if aa:
            raise TypeError("resample() got unexpected keyword arguments %s" % sorted(aa))
        return _engine.resample(ops, data0, shape, m, bound, fillvalue=fillvalue,
                                precision=precision, tuning=tuning, out=out, oblur=oblur, sync=sync,
                                status_host=status_host)

    def remap(self, data, /, method='n', bound='t', phot='radiance', shape=None, template=None,
              input_range=None, output_range=None, justify=False, rectify=True, wcs=None, **kw):
        """Resample in scientific (WCS) coordinates (reference: core.py:589-930)."""
        from .remap import remap as _remap
        return _remap(self, data, method=method, bound=bound, phot=phot, shape=shape,
                      template=template, input_range=input_range, output_range=output_range,
                      justify=justify, rectify=rectify, wcs=wcs, **kw)


class Identity(Transform):
    """core.py:1345-1401."""

    def __init__(self, **kwargs):
        self.idim = 0
        self.odim = 0
        self.no_forward = 0
        self.no_reverse = 0
        self.iunit = ""
        self.ounit = ""
        self.itype = ""
        self.otype = ""
        self.params = kwargs

    def _lower(self, direction):
        return []

    def __str__(self):
        s = "Identity"
        if len(self.params):
            s += " (%d param%s)" % (len(self.params), "" if len(self.params) == 1 else "s")
        self._strtmp = s
        return super().__str__()

    def inverse(self):
        return self


class PlusOne_(Transform):
    """core.py:1404-1435: 1-D test transform adding one to the first component (lowered as a
    1-D offset embedded in the 2-D opcode)."""

    def __init__(self):
        self.idim = 1
        self.odim = 1
        self.no_forward = False
        self.no_reverse = False
        self.iunit = None
        self.ounit = None
        self.itype = None
        self.otype = None
        self.params = {}

    def _lower(self, direction):
        op = L.TfmOp()
        op.kind, op.dir, op.flags = L.OP_LINEAR, direction, L.LIN_HAS_PRE
        op.p[4], op.p[5] = 1.0, 0.0
        return [op]

    def __str__(self):
        self._strtmp = "_PlusOne"
        return super().__str__()


class Inverse(Transform):
    """core.py:1438-1491."""

    def __init__(self, t):
        self.idim = t.odim
        self.odim = t.idim
        self.no_forward = t.no_reverse
        self.no_reverse = t.no_forward
        self.iunit = t.ounit
        self.ounit = t.iunit
        self.itype = t.otype
        self.otype = t.itype
        self.params = {'transform': copy.copy(t)}

    def _lower(self, direction):
        inner = self.params['transform']
        # Inverse._forward -> inner.apply(invert=True) re-checks validity (core.py:1478-1482)
        return inner._chain(direction == L.FWD)

    def __str__(self):
        self.params['transform']._str_not_top_tmp = True
        self._strtmp = "Inverse " + self.params['transform'].__str__()
        return super().__str__()

    def inverse(self):
        return self.params['transform']


class Composition(Transform):
    """core.py:1495-1572: members applied right-to-left forward, left-to-right (each inverted)
    in reverse; nested compositions are flattened at construction."""

    def __init__(self, translist):
        if not isinstance(translist, list):
            raise ValueError("Composition requires a list of Transforms")
        if len(translist) < 1:
            raise ValueError("Composition requires at least one Transform")
        members = []
        idim = 0
        odim = 0
        for trans in translist:
            if not isinstance(trans, Transform):
                raise AssertionError("transform.Composition: got something that's not a Transform")
            if idim == 0:
                idim = trans.idim
            if trans.odim != 0:
                odim = trans.odim
            if isinstance(trans, Composition):
                members.extend(copy.copy(trans.params['list']))
            else:
                members.append(copy.copy(trans))
        self.idim = idim
        self.odim = odim
        self.no_forward = any(m.no_forward for m in members)
        self.no_reverse = any(m.no_reverse for m in members)
        self.iunit = members[-1].iunit
        self.ounit = members[0].ounit
        self.itype = members[-1].itype
        self.otype = members[0].otype
        self.params = {'list': members}

    def _lower(self, direction):
        ops = []
        if direction == L.FWD:                   # core.py:1551-1554
            for xf in self.params['list'][::-1]:
                ops.extend(xf._chain(False))
        else:                                    # core.py:1556-1559
            for xf in self.params['list']:
                ops.extend(xf._chain(True))
        return ops

    def __str__(self):
        strings = []
        for xf in self.params['list']:
            xf._str_not_top_tmp = True
            strings.append(xf.__str__())
        self._strtmp = '( (' + ') o ('.join(strings) + ') )'
        return super().__str__()


class Wrap(Composition):
    """core.py:1576-1600: W^-1 o T o W."""

    def __init__(self, T, W):
        super().__init__([W.inverse(), T, W])
