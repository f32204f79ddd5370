"""DataWrapper, the WCS transform and Transform.remap.

Host-side mirror of /root/reference/transform/core.py:589-930 (remap), 941-1260 (DataWrapper) and
1650-1747 (WCS).  remap() is metadata work on 121 sample points plus ONE call of resample(); it
keeps the reference's logic -- including its (Y,X)/(X,Y) mix when autoscaling non-square images
(core.py:876, 900, 902) -- so that the composed chain handed to the GPU
is the one the reference would have built.  astropy objects are replaced by fitswcs.Header /
fitswcs.WCSParams (astropy is absent here); if a caller passes astropy objects they are read
through their mapping / to_header() interfaces.
"""
import copy

import numpy as np

from . import _lib as L
from .core import Transform, Identity, Composition
from .fitswcs import Header, WCSParams, HDUList, PrimaryHDU, FITSHeaderFromDict


def _as_header(obj):
    if isinstance(obj, Header):
        return obj
    if hasattr(obj, 'to_header') and not isinstance(obj, WCSParams):     # astropy WCS
        return Header(dict(obj.to_header()))
    return Header(obj)


def _as_wcs(obj):
    if isinstance(obj, WCSParams):
        return obj
    if hasattr(obj, 'to_header'):                                         # astropy WCS (duck-typed)
        try:
            hdr = obj.to_header(relax=True)       # the conversion must not lose a SIP distortion the object carries
        except TypeError:
            hdr = obj.to_header()
        w = WCSParams(Header(dict(hdr)))
        if getattr(obj, 'pixel_shape', None) is not None:
            w.pixel_shape = list(obj.pixel_shape)
        return w
    return WCSParams(_as_header(obj))


class DataWrapper:
    """(data, header, wcs) triple (core.py:941-1260).  Indexable by 0/1/2 and 'data'/'header'/'wcs'."""

    def __init__(self, this, template=None):
        data = header = wcs = None
        if isinstance(this, DataWrapper):
            data, header, wcs = this.data, copy.copy(this.header), copy.copy(this.wcs)
        else:
            if isinstance(this, HDUList) or (isinstance(this, list) and this
                                             and hasattr(this[0], 'header') and hasattr(this[0], 'data')):
                this = this[0]                                            # core.py:991-992
            if isinstance(this, (tuple, list)):
                try:
                    data = copy.copy(this[0])
                    header = copy.copy(this[1])
                except Exception:
                    raise ValueError("DataWrapper: tuple requies both data and header")
                if len(this) > 2:
                    wcs = copy.copy(this[2])
            elif isinstance(this, PrimaryHDU):
                data = copy.copy(this.data)
                header = copy.copy(this.header)
                wcs = WCSParams(header)
            elif isinstance(this, Header):
                header = copy.copy(this)
            elif isinstance(this, WCSParams):
                wcs = copy.copy(this)
            elif isinstance(this, np.ndarray):
                data = copy.copy(this)
            elif isinstance(this, dict):
                if 'data' in this or 'header' in this or 'wcs' in this:
                    data = this.get('data')
                    header = this.get('header')
                    wcs = this.get('wcs')
                elif 'NAXIS' in this or 'SIMPLE' in this:
                    header = copy.copy(this)
                else:
                    raise ValueError("DataWrapper: dict requires either: FITS header tags; or "
                                     "'data', 'header' and/or 'wcs'")
            elif isinstance(this, str):
                raise ValueError("DataWrapper: couldn't parse meaningful data from the supplied %s."
                                 % this.__class__)
            elif hasattr(this, 'data') or hasattr(this, 'header') or hasattr(this, 'wcs'):
                data = copy.copy(getattr(this, 'data', None))
                header = copy.copy(getattr(this, 'header', None))
                wcs = copy.copy(getattr(this, 'wcs', None))
            else:
                raise ValueError("DataWrapper: couldn't parse meaningful data from the supplied %s."
                                 % this.__class__)

        if header is None:
            if wcs is not None:
                wcs = _as_wcs(wcs)
                header = wcs.to_header()
                header['NAXIS'] = wcs.naxis
                if wcs.pixel_shape is None and data is None:
                    raise ValueError("DataWrapper: couldn't determine size of the pixel canvas "
                                     "(no header; wcs object has no pixel_bound)")
                if data is not None:
                    for ii in range(len(data.shape)):
                        header["NAXIS%d" % (len(data.shape) - ii)] = data.shape[ii]
                else:
                    for ii in range(wcs.naxis):
                        header["NAXIS%d" % (ii + 1)] = wcs.pixel_shape[ii]
            elif data is not None:
                if not isinstance(data, np.ndarray):
                    raise ValueError("DataWrapper: data must be a numpy array")
                h = {'NAXIS': len(data.shape)}                            # core.py:1095-1115
                names = {1: 'X axis', 2: 'Y axis', 3: 'Z axis'}
                for ii in range(len(data.shape)):
                    axis = len(data.shape) - ii
                    h["NAXIS%d" % axis] = data.shape[ii]
                    h["CDELT%d" % axis] = 1.0
                    h["CUNIT%d" % axis] = "Pixel"
                    h["CTYPE%d" % axis] = names.get(axis, "axis %d" % axis)
                    h["CRPIX%d" % axis] = 1
                    h["CRVAL%d" % axis] = 0
                h["COMMENT"] = "Header autogenerated by transform.DataWrapper"
                header = FITSHeaderFromDict(h)
                wcs = WCSParams(header)
            else:
                raise ValueError("DataWrapper: requires data, header, and/or WCS object.")
        else:
            header = _as_header(header)
            wcs = WCSParams(header) if wcs is None else _as_wcs(wcs)

        self.wcs = wcs
        self.header = header
        self.data = data
        self._WCS_ = None

        if template is not None:
            if isinstance(template, DataWrapper):
                template = template.header
            if isinstance(template, WCSParams) or (hasattr(template, 'to_header')
                                                   and not isinstance(template, (Header, dict))):
                self.wcs = _as_wcs(template)
                self.wcs2head()
            else:
                if isinstance(template, (Header, dict)):
                    self.wcs = WCSParams(_as_header(template))
                elif (isinstance(template, (list, tuple))
                      or (isinstance(template, np.ndarray) and template.ndim == 1)):
                    if len(template) != self.header['NAXIS']:
                        raise ValueError("DataWrapper: shape format template must match data axis count")
                    for i in range(self.header['NAXIS']):
                        self.header["NAXIS%d" % i] = template[-1 - i]     # (sic) core.py:1175
                    self.wcs.pixel_shape = [template[-1 - i] for i in range(self.header['NAXIS'])]
                else:
                    raise ValueError("DataWrapper: template must be a FITS header, WCS object, or shape")
                self.wcs2head()

    def WCS(self):
        if self._WCS_ is None:
            self._WCS_ = WCS(self)
        return self._WCS_

    def head2wcs(self):
        self.wcs = WCSParams(self.header)
        self._WCS_ = None

    def wcs2head(self):
        """core.py:1204-1245."""
        hdr = self.wcs.to_header()
        if self.header is None:
            self.header = hdr
        else:
            header = self.header
            n = hdr['WCSAXES']
            for i in range(n):
                for stem in ("CDELT%d", "CRPIX%d", "CRVAL%d"):
                    if (stem % (i + 1)) in header:
                        del header[stem % (i + 1)]
                for j in range(n):
                    for stem in ("CD%d_%d", "PC%d_%d"):
                        if (stem % (i + 1, j + 1)) in header:
                            del header[stem % (i + 1, j + 1)]
            if "CROTA2" in header:
                del header["CROTA2"]
            for ky in hdr.keys():
                header.set(ky, hdr[ky])
        if 'NAXIS' not in self.header:
            self.header['NAXIS'] = self.wcs.naxis
        if self.wcs.pixel_shape is not None:
            for i in range(self.wcs.naxis):
                self.header['NAXIS%d' % (i + 1)] = self.wcs.pixel_shape[i]

    def __getitem__(self, index):
        if index == 0 or index == 'data':
            return self.data
        if index == 1 or index == 'header':
            return self.header
        if index == 2 or index == 'wcs':
            return self.wcs
        if index == 3 or index == 'WCS':
            return self.WCS
        raise IndexError(index)


class WCS(Transform):
    """Pixel (origin 0) <-> world coordinates from a FITS header (core.py:1650-1747).

    Where the reference calls astropy's all_pix2world / all_world2pix (core.py:1727, 1742), this
    lowers the header to a TFM_OP_WCS opcode: linear part, plus one of ten spherical projections with
    the celestial rotation when CTYPE says so, preceded (pixel -> world) or followed (world -> pixel)
    by a TFM_OP_SIP when the header carries a SIP polynomial distortion (A_ORDER / B_ORDER ...)."""

    def __init__(self, dingus):
        if isinstance(dingus, DataWrapper):
            if not isinstance(dingus.wcs, WCSParams):
                dingus.head2wcs()
            wcs_obj = dingus.wcs
        elif isinstance(dingus, WCSParams):
            wcs_obj = dingus
        elif hasattr(dingus, 'to_header') and not isinstance(dingus, (Header, dict)):
            wcs_obj = _as_wcs(dingus)              # an astropy.wcs.WCS (core.py:1688-1689), duck-typed
        else:
            wcs_obj = DataWrapper(dingus).wcs
        n = wcs_obj.naxis
        names = ['X', 'Y', 'Z']
        itype = names[0:min(len(names), n)]
        while len(itype) < n:
            itype.append('Coord %d' % len(itype))
        self.idim = n + 0
        self.odim = n + 0
        self.iunit = ['Pixels'] * n
        self.ounit = ["%s" % a for a in wcs_obj.cunit]
        self.itype = itype
        self.otype = ["%s" % a for a in wcs_obj.ctype]
        self.no_forward = False
        self.no_reverse = False
        self.params = {'wcs': wcs_obj}

    def _lower(self, direction):
        return self.params['wcs'].lower(direction)

    def __str__(self):
        self._strtmp = "WCS"
        return super().__str__()


def remap(self, data, method='n', bound='t', phot='radiance', shape=None, template=None,
          input_range=None, output_range=None, justify=False, rectify=True, wcs=None, **kw):
    """Transform.remap (core.py:589-930)."""
    data = DataWrapper(data, template=wcs)
    if not isinstance(data.data, np.ndarray):
        raise ValueError("remap: data must be a NumPy array")
    input_trans = WCS(data) if data.wcs is not None else Identity()
    idim = len(data.data.shape) if self.idim == 0 else self.idim

    out_template = DataWrapper(template) if template is not None else None

    if shape is None:                                                     # core.py:818-835
        if out_template is None:
            shape = data.data.shape
        elif out_template.wcs is None or out_template.wcs.pixel_shape is None:
            shape = data.data.shape if out_template.data is None else out_template.data.shape
        else:
            shape = list(reversed(out_template.wcs.pixel_shape))

    if out_template is None or out_template.wcs is None:
        if output_range is not None:                                      # core.py:841-849
            if not isinstance(output_range, np.ndarray):
                output_range = np.array(output_range)
            if output_range.ndim != 2 or (self.odim != 0 and output_range.shape[0] != self.odim):
                raise ValueError("remap: output_range must be Nx2 and match Transform dims")
        else:
            n = 11
            if input_range is not None:                                   # core.py:857-870
                try:
                    ishape = input_range.shape
                except AttributeError:
                    raise ValueError('remap: input_range parameter must be an Nx2 NumPy array')
                if len(ishape) != 2 or ishape[0] != idim:
                    raise ValueError("remap: input_range must be Nx2 and match Transform dims")
                isamp = np.mgrid[tuple(slice(0, n) for _ in range(idim))].T / (n - 1.0)
                isamp = isamp * (input_range[..., 1] - input_range[..., 0]) + input_range[..., 0]
            else:                                                         # core.py:872-878
                isamp = np.mgrid[tuple(slice(0, n) for _ in range(idim))].T / (n - 1.0)
                isamp = isamp * data.data.shape        # (sic) (X,Y) vector times (NY,NX): core.py:876
                isamp = isamp - 0.5
                isamp = WCS(data).apply(isamp)
            osamp = np.asarray(self.apply(isamp))                         # core.py:883
            axes = tuple(range(osamp.ndim - 1))
            output_range = np.stack((np.amin(osamp, axis=axes), np.amax(osamp, axis=axes)), 1)
        assert len(shape) == output_range.shape[0]                       # core.py:895
        otwcs = WCSParams(None, naxis=len(shape))                         # core.py:899-902
        otwcs.crpix = [shape[i] / 2 + 0.5 for i in range(len(shape))]
        otwcs.crval = list((output_range[..., 0] + output_range[..., 1]) * 0.5)
        otwcs.cdelt = [(output_range[i, 1] - output_range[i, 0]) / shape[i] for i in range(len(shape))]
        otwcs.pixel_shape = None                                          # (sic) list.reverse() is None: core.py:914
        out_template = DataWrapper(data.header, template=otwcs)           # core.py:916

    output_trans = out_template.WCS().inverse()                           # core.py:920
    total_trans = Composition([output_trans, self, input_trans])          # core.py:923
    resampled = total_trans.resample(data.data, method=method, bound=bound, phot=phot,
                                     shape=shape, **kw)                    # core.py:924
    return DataWrapper((resampled, data.header), template=out_template)
