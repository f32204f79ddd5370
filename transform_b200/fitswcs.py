"""Minimal FITS header / FITS image reader / WCS parameter holder.

The reference hands all of this to astropy (`astropy.io.fits.Header`, `astropy.wcs.WCS`,
core.py:991-1023, 1126-1131, 1727, 1742), which is neither pinned by the reference
(setup.py:27-32) nor installed here.  This module provides the small subset the resample / remap
path needs, restated from the FITS standard and from Greisen & Calabretta 2002 (A&A 395, 1061,
"paper I") and Calabretta & Greisen 2002 (A&A 395, 1077, "paper II") -- the papers the reference
itself cites (core.py:38-42, 1655-1658):

  * Header       ordered, dict-like FITS header with the astropy methods core.py uses
  * read_fits    primary-HDU reader (BITPIX 8/16/32/64/-32/-64, big-endian, BSCALE/BZERO)
  * WCSParams    CRPIX/CRVAL/CDELT + CD | PC | CROTA2 linear part (paper I), ten spherical
                 projections with LONPOLE/LATPOLE defaults (paper II), CEA's lambda, SIN's slant, the
                 SIP (Shupe et al. 2005) and TPV polynomial distortions; `lower()` packs them into
                 the TFM_OP_WCS (+ TFM_OP_SIP / TFM_OP_TPV) opcodes evaluated by the CUDA library.

PARITY: the linear part is pinned by the reference's own test (tests/test_01_core.py:160-169,
sample.fits).  The spherical projections, SIP and TPV are PARITY UNPINNED (no astropy/wcslib to diff against).
"""
import copy
import math

import numpy as np

from . import _lib as L

_COMMENTARY = ('', 'HISTORY', 'COMMENT')


class Header:
    """Ordered FITS header: the subset of astropy.io.fits.Header that core.py touches
    (item access, `in`, `del`, keys(), set(), get(), copy)."""

    def __init__(self, cards=None):
        self._keys = []
        self._vals = {}
        self._commentary = {k: [] for k in _COMMENTARY}
        if cards is None:
            return
        if isinstance(cards, Header):
            self._keys = list(cards._keys)
            self._vals = dict(cards._vals)
            self._commentary = {k: list(v) for k, v in cards._commentary.items()}
            return
        items = cards.items() if hasattr(cards, 'items') else cards
        for k, v in items:
            self.set(k, v)

    @staticmethod
    def _norm(key):
        return key.upper() if isinstance(key, str) else key

    def set(self, key, value=None):
        key = self._norm(key)
        if key in _COMMENTARY:
            if isinstance(value, (list, tuple)):
                self._commentary[key].extend(str(v) for v in value)
            else:
                self._commentary[key].extend(str(value).split("\n"))
            return
        if key not in self._vals:
            self._keys.append(key)
        self._vals[key] = value

    def __setitem__(self, key, value):
        self.set(key, value)

    def __getitem__(self, key):
        key = self._norm(key)
        if key in _COMMENTARY:
            return list(self._commentary[key])
        return self._vals[key]

    def __delitem__(self, key):
        key = self._norm(key)
        if key in _COMMENTARY:
            self._commentary[key] = []
            return
        del self._vals[key]
        self._keys.remove(key)

    def __contains__(self, key):
        key = self._norm(key)
        if key in _COMMENTARY:
            return bool(self._commentary[key])
        return key in self._vals

    def get(self, key, default=None):
        return self[key] if key in self else default

    def keys(self):
        return list(self._keys) + [k for k in _COMMENTARY if self._commentary[k]]

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._keys)

    def copy(self):
        return Header(self)

    def __copy__(self):
        return Header(self)

    def __eq__(self, other):
        try:
            return dict(self.items()) == dict(other.items())
        except AttributeError:
            return NotImplemented

    def __repr__(self):
        return "Header(%r)" % (self.items(),)


def FITSHeaderFromDict(HeaderDict):
    """core.py:1262-1324: dict -> header, tolerating multi-line '' / HISTORY / COMMENT cards."""
    if isinstance(HeaderDict, Header):
        return HeaderDict
    return Header(HeaderDict)


class PrimaryHDU:
    def __init__(self, data, header):
        self.data = data
        self.header = header


class HDUList(list):
    pass


def _parse_card_value(text):
    text = text.strip()
    if not text:
        return None
    if text[0] == "'":
        end = 1
        out = []
        while end < len(text):
            if text[end] == "'":
                if end + 1 < len(text) and text[end + 1] == "'":
                    out.append("'")
                    end += 2
                    continue
                break
            out.append(text[end])
            end += 1
        return "".join(out).rstrip()
    if "/" in text:
        text = text.split("/", 1)[0].strip()
    if text in ("T", "F"):
        return text == "T"
    try:
        return int(text)
    except ValueError:
        pass
    try:
        return float(text.replace("D", "E").replace("d", "e"))
    except ValueError:
        return text


def decode_fits_data(raw_dev, bitpix, shape, bscale=1, bzero=0):
    """FITS data unit on the device -> CUDA tensor of `shape` (K0, csrc/tfm_fits.cu).

    raw_dev: 1-D uint8 CUDA tensor holding the big-endian samples as they sit in the file.
    Returns float32 for BITPIX = -32 (a byte swap; scaling, if any, in float32 as NumPy does for a
    float32 array and Python scalars), float64 otherwise -- integer samples convert exactly (the
    resamplers carry integer images as float64 anyway) and scaling is `stored * BSCALE + BZERO`
    with NumPy's two roundings."""
    from . import _device as dev
    torch = dev.require_cuda()
    lib = L.load()
    n = 1
    for s_ in shape:
        n *= int(s_)
    nbytes = n * (abs(int(bitpix)) // 8)
    if raw_dev.dtype != torch.uint8 or not raw_dev.is_cuda or raw_dev.numel() < nbytes:
        raise ValueError("decode_fits_data: raw_dev must be a uint8 CUDA tensor of at least %d bytes" % nbytes)
    f32 = int(bitpix) == -32
    out = torch.empty(tuple(int(s_) for s_ in shape), dtype=torch.float32 if f32 else torch.float64,
                      device=raw_dev.device)
    rc = lib.tfm_fits_decode(raw_dev.data_ptr(), int(bitpix), n, float(bscale), float(bzero),
                             L.F32 if f32 else L.F64, out.data_ptr(), dev.stream_ptr())
    L.raise_for_status(rc, "tfm_fits_decode")
    return out


def read_fits(path, device=False):
    """Read the primary HDU of a FITS file -> HDUList([PrimaryHDU(data, header)]).

    device=True: the data unit is uploaded as raw bytes and decoded on the GPU (`decode_fits_data`);
    `data` is then a CUDA tensor that resample()/remap() take without another copy."""
    with open(path, "rb") as f:
        raw = f.read()
    hdr = Header()
    pos = 0
    done = False
    while not done:
        block = raw[pos:pos + 2880]
        if len(block) < 2880:
            raise ValueError("read_fits: truncated header in %s" % path)
        pos += 2880
        for i in range(0, 2880, 80):
            card = block[i:i + 80].decode("ascii", errors="replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if key in ("COMMENT", "HISTORY") or (key == "" and card.strip()):
                hdr.set(key, card[8:].rstrip())
            elif card[8:10] == "= ":
                hdr.set(key, _parse_card_value(card[10:]))
    naxis = int(hdr.get("NAXIS", 0))
    data = None
    if naxis > 0:
        shape = tuple(int(hdr["NAXIS%d" % (naxis - i)]) for i in range(naxis))
        bitpix = int(hdr["BITPIX"])
        dt = {8: ">u1", 16: ">i2", 32: ">i4", 64: ">i8", -32: ">f4", -64: ">f8"}[bitpix]
        n = int(np.prod(shape, dtype=np.int64))
        if device:
            from . import _device as dev
            torch = dev.require_cuda()
            nbytes = n * (abs(bitpix) // 8)
            if len(raw) < pos + nbytes:
                raise ValueError("read_fits: truncated data unit in %s" % path)
            stage = dev._pinned(nbytes, "fits")
            stage[:nbytes].numpy()[:] = np.frombuffer(raw, dtype=np.uint8, count=nbytes, offset=pos)
            raw_dev = stage[:nbytes].to("cuda", non_blocking=True)
            data = decode_fits_data(raw_dev, bitpix, shape, hdr.get("BSCALE", 1), hdr.get("BZERO", 0))
            torch.cuda.current_stream().synchronize()        # the staging buffer is reused
            return HDUList([PrimaryHDU(data, hdr)])
        arr = np.frombuffer(raw, dtype=dt, count=n, offset=pos).reshape(shape)
        data = arr.astype(arr.dtype.newbyteorder("="))
        bscale, bzero = hdr.get("BSCALE", 1), hdr.get("BZERO", 0)
        if bscale != 1 or bzero != 0:
            data = data * bscale + bzero
    return HDUList([PrimaryHDU(data, hdr)])


# ------------------------------------------------------------------------------------------
# WCS
# ------------------------------------------------------------------------------------------
_LNG_PREFIX = ("RA--", "GLON", "ELON", "SLON", "HLON", "HPLN", "HGLN", "CRLN", "SOLX")
_LAT_PREFIX = ("DEC-", "GLAT", "ELAT", "SLAT", "HLAT", "HPLT", "HGLT", "CRLT", "SOLY")
# code, phi0, theta0 (paper II table 13: zenithal projections have their reference point at the
# native pole, the cylindrical / pseudo-cylindrical / Hammer-Aitoff ones at (0, 0))
_PROJ = {"TAN": (L.WCS_PROJ_TAN, 0.0, 90.0), "CAR": (L.WCS_PROJ_CAR, 0.0, 0.0),
         "STG": (L.WCS_PROJ_STG, 0.0, 90.0), "SIN": (L.WCS_PROJ_SIN, 0.0, 90.0),
         "ARC": (L.WCS_PROJ_ARC, 0.0, 90.0), "ZEA": (L.WCS_PROJ_ZEA, 0.0, 90.0),
         "CEA": (L.WCS_PROJ_CEA, 0.0, 0.0), "MER": (L.WCS_PROJ_MER, 0.0, 0.0),
         "SFL": (L.WCS_PROJ_SFL, 0.0, 0.0), "AIT": (L.WCS_PROJ_AIT, 0.0, 0.0)}
_PROJ_NAME = {v[0]: k for k, v in _PROJ.items()}
_PROJ["TPV"] = _PROJ["TAN"]        # TAN with the PVi_m polynomial distortion of the TPV convention (TFM_OP_TPV)


def _sph_to_vec(lon_deg, lat_deg):
    lo, la = math.radians(lon_deg), math.radians(lat_deg)
    return np.array([math.cos(la) * math.cos(lo), math.cos(la) * math.sin(lo), math.sin(la)])


def _native_to_celestial_point(phi, theta, alpha_p, delta_p, phi_p):
    """Paper II eq. (2): native (phi, theta) -> celestial (alpha, delta), degrees."""
    ph, th = math.radians(phi - phi_p), math.radians(theta)
    dp = math.radians(delta_p)
    alpha = alpha_p + math.degrees(math.atan2(-math.cos(th) * math.sin(ph),
                                              math.sin(th) * math.cos(dp)
                                              - math.cos(th) * math.sin(dp) * math.cos(ph)))
    sd = math.sin(th) * math.sin(dp) + math.cos(th) * math.cos(dp) * math.cos(ph)
    delta = math.degrees(math.asin(max(-1.0, min(1.0, sd))))
    return alpha, delta


class WCSParams:
    """The pieces of an astropy.wcs.WCS that core.py reads or writes, for 2 axes.

    Attributes mirror astropy's `wcs.wcs.*` names: naxis, crpix, crval, cdelt, ctype, cunit,
    pc (the matrix, CD folded in as PC with CDELT=1 the way astropy does), pixel_shape (X,Y)."""

    def __init__(self, header=None, naxis=2):
        self.naxis = naxis
        self.crpix = [0.0] * naxis
        self.crval = [0.0] * naxis
        self.cdelt = [1.0] * naxis
        self.ctype = [''] * naxis
        self.cunit = [''] * naxis
        self.pc = np.eye(naxis)
        self.lonpole = None
        self.latpole = None
        self.pixel_shape = None
        self.cea_lambda = 1.0          # PV2_1 of a CEA projection (paper II eq. 80)
        self.sin_slant = (0.0, 0.0)    # (xi, eta) = (PV2_1, PV2_2) of a SIN projection (paper II eq. 61-62)
        self.tpv = None                # TPV: float64 [2][40], tpv[0][m] = PV1_m, tpv[1][m] = PV2_m
        self.sip_a = None              # SIP: (A_ORDER+1)^2 array, sip_a[p, q] = A_p_q (Shupe et al. 2005)
        self.sip_b = None
        self._sip_dev = {}             # (device, table bytes) -> resident coefficient table
        if header is not None:
            self._from_header(header)

    # astropy spells these wcs.wcs.crpix etc.; core.py uses both levels
    @property
    def wcs(self):
        return self

    def _from_header(self, h):
        def get(key, default=None):
            return h[key] if key in h else default
        naxis = get('WCSAXES', get('NAXIS', None))
        if naxis is None:
            naxis = 0
            for k in h.keys():
                for stem in ('CTYPE', 'CRPIX', 'CRVAL', 'CDELT'):
                    if isinstance(k, str) and k.startswith(stem) and k[len(stem):].isdigit():
                        naxis = max(naxis, int(k[len(stem):]))
        naxis = int(naxis) if naxis else 2
        self.__init__(None, naxis)
        for i in range(naxis):
            a = i + 1
            self.crpix[i] = float(get('CRPIX%d' % a, 0.0))
            self.crval[i] = float(get('CRVAL%d' % a, 0.0))
            self.cdelt[i] = float(get('CDELT%d' % a, 1.0))
            self.ctype[i] = str(get('CTYPE%d' % a, '')).strip()
            self.cunit[i] = str(get('CUNIT%d' % a, '')).strip()
        self._read_sip(h)
        for k in h.keys():
            if isinstance(k, str) and (k[:5] in ('CPDIS', 'CQDIS', 'D2IMD') or k[:4] == 'D2IM'
                                       or (k[:2] in ('DP', 'DQ') and k[2:3].isdigit())):
                raise NotImplementedError("transform_b200: table-lookup distortion keyword %s "
                                          "(Paper IV / D2IM) is not on the GPU path" % k)
        has_cd = any(('CD%d_%d' % (i + 1, j + 1)) in h for i in range(naxis) for j in range(naxis))
        has_pc = any(('PC%d_%d' % (i + 1, j + 1)) in h for i in range(naxis) for j in range(naxis))
        if has_cd:
            # astropy/wcslib: CDi_j is carried as PCi_j with CDELT = 1 (missing elements are 0)
            m = np.zeros((naxis, naxis))
            for i in range(naxis):
                for j in range(naxis):
                    m[i, j] = float(get('CD%d_%d' % (i + 1, j + 1), 0.0))
            self.pc = m
            self.cdelt = [1.0] * naxis
        elif has_pc:
            m = np.eye(naxis)
            for i in range(naxis):
                for j in range(naxis):
                    m[i, j] = float(get('PC%d_%d' % (i + 1, j + 1), 1.0 if i == j else 0.0))
            self.pc = m
        elif 'CROTA2' in h and naxis >= 2:
            rho = math.radians(float(h['CROTA2']))      # paper I / AIPS convention
            c, s = math.cos(rho), math.sin(rho)
            m = np.eye(naxis)
            m[0, 0], m[0, 1] = c, -s * self.cdelt[1] / self.cdelt[0]
            m[1, 0], m[1, 1] = s * self.cdelt[0] / self.cdelt[1], c
            self.pc = m
        if 'LONPOLE' in h:
            self.lonpole = float(h['LONPOLE'])
        if 'LATPOLE' in h:
            self.latpole = float(h['LATPOLE'])
        self._read_pv(h)               # after LONPOLE / LATPOLE: PVi_3 / PVi_4 override them
        if all(('NAXIS%d' % (i + 1)) in h for i in range(naxis)):
            self.pixel_shape = [int(h['NAXIS%d' % (i + 1)]) for i in range(naxis)]

    def _read_pv(self, h):
        """PVi_m projection parameters (paper II section 2.3 and table 13).  On the GPU path: lambda of
        CEA (PV<lat>_1), LONPOLE / LATPOLE given as PV<lng>_3 / PV<lng>_4 (they override the
        keywords, as in wcslib's wcsset) and explicit statements of a projection's default
        (phi0, theta0) in PV<lng>_1 / PV<lng>_2; anything else raises."""
        pv = {}
        for k in h.keys():
            if isinstance(k, str) and k.startswith('PV') and '_' in k and k[2:].replace('_', '').isdigit():
                i, m = k[2:].split('_')
                pv[(int(i), int(m))] = float(h[k])
        if not pv:
            return
        try:
            proj = self.projection()
        except NotImplementedError:
            proj = None
        name = _PROJ_NAME.get(proj[0]) if proj is not None else None
        lng, lat = 1, 2
        if self.is_tpv():
            # every PV1_m / PV2_m of a -TPV header is a coefficient of the distortion polynomial
            tab = np.zeros((2, L.TPV_NCOEF))
            tab[0, 1] = tab[1, 1] = 1.0
            for (i, m), val in pv.items():
                if i not in (1, 2) or not (0 <= m < L.TPV_NCOEF):
                    raise ValueError("WCS: PV%d_%d is not a term of the TPV polynomial (m = 0..39 on axes 1, 2)" % (i, m))
                tab[i - 1, m] = val
            self.tpv = tab
            return
        for (i, m), val in sorted(pv.items()):
            ok = False
            if proj is not None and i == lng:
                if m == 0:
                    ok = (val == 0.0)                      # PVi_0: no offset of the fiducial point
                elif m == 1:
                    ok = (val == proj[3])
                elif m == 2:
                    ok = (val == proj[4])
                elif m == 3:
                    self.lonpole, ok = val, True
                elif m == 4:
                    self.latpole, ok = val, True
            elif proj is not None and i == lat:
                if name == 'CEA' and m == 1:
                    if not (0.0 < val <= 1.0):
                        raise ValueError("WCS: CEA needs 0 < PV2_1 <= 1, got %r" % val)
                    self.cea_lambda, ok = val, True
                elif name == 'SIN' and m in (1, 2):
                    sl = list(self.sin_slant)
                    sl[m - 1] = val
                    self.sin_slant, ok = tuple(sl), True
                else:
                    ok = (val == 0.0) and name in ('TAN', 'CAR', 'STG', 'SIN', 'ARC', 'ZEA', 'MER', 'SFL', 'AIT')
            else:
                ok = (val == 0.0)
            if not ok:
                raise NotImplementedError("transform_b200: projection parameter PV%d_%d = %r is not on "
                                          "the GPU path (CEA's lambda, SIN's slant, the TPV polynomial, "
                                          "LONPOLE/LATPOLE as PVi_3/PVi_4 and default values are)" % (i, m, val))

    def _read_sip(self, h):
        """A_ORDER / B_ORDER / A_p_q / B_p_q (Shupe et al. 2005).  astropy applies them whenever
        the keywords are present, with or without the '-SIP' suffix of CTYPE; AP_/BP_ (the
        inverse polynomials) are not used by all_world2pix and are ignored here too."""
        has_a, has_b = 'A_ORDER' in h, 'B_ORDER' in h
        if not (has_a or has_b):
            return
        if not (has_a and has_b):
            raise ValueError("WCS: Both A_ORDER and B_ORDER must be provided for a SIP distortion")
        tabs = []
        for L_ in ('A', 'B'):
            m = int(h[L_ + '_ORDER'])
            if not (0 <= m <= L.SIP_MAX_ORDER):
                raise ValueError("WCS: %s_ORDER = %d is outside 0..%d" % (L_, m, L.SIP_MAX_ORDER))
            t = np.zeros((m + 1, m + 1))
            for p in range(m + 1):
                for q in range(m + 1 - p):
                    key = '%s_%d_%d' % (L_, p, q)
                    if key in h:
                        t[p, q] = float(h[key])
            tabs.append(t)
        self.sip_a, self.sip_b = tabs

    def has_sip(self):
        return self.sip_a is not None and self.sip_b is not None

    def is_tpv(self):
        return self.naxis >= 2 and all(len(c) >= 8 and c[4] == '-' and c[5:8].upper() == 'TPV' for c in self.ctype[:2])

    def _table_on_device(self, tab, slot):
        """Coefficient table resident on the current device: uploaded once per device and table contents;
        the tensor lives as long as this object."""
        from . import _device as dev
        torch = dev.require_cuda()
        key = (int(torch.cuda.current_device()), slot, tab.tobytes())
        t = self._sip_dev.get(key)
        if t is None:
            t = dev.to_device(tab, np.float64, slot=slot)
            self._sip_dev = {k: v for k, v in self._sip_dev.items() if k[:2] != key[:2]}
            self._sip_dev[key] = t
        return t

    def _tpv_op(self, direction):
        """TFM_OP_TPV (include/tfm_b200.h): the polynomial table on the device, the inverse of its linear part."""
        import struct
        tab = np.ascontiguousarray(self.tpv, dtype=np.float64)
        if tab.shape != (2, L.TPV_NCOEF):
            raise ValueError("WCS: the TPV table must be [2][%d]" % L.TPV_NCOEF)
        M = np.array([[tab[0, 1], tab[0, 2]], [tab[1, 2], tab[1, 1]]])
        if abs(np.linalg.det(M)) < 1e-300:
            raise ValueError("WCS: the linear part of the TPV polynomial is singular")
        Minv = np.linalg.inv(M)
        t = self._table_on_device(tab, "tpv")
        op = L.TfmOp()
        op.kind, op.dir, op.flags = L.OP_TPV, direction, 0
        op.p[0] = struct.unpack("d", struct.pack("q", int(t.data_ptr())))[0]
        op.p[1], op.p[2], op.p[3], op.p[4] = (float(Minv[0, 0]), float(Minv[0, 1]),
                                              float(Minv[1, 0]), float(Minv[1, 1]))
        return op

    def sip_table(self):
        """float64 [2][SIP_DIM][SIP_DIM]: the layout TFM_OP_SIP reads (include/tfm_b200.h)."""
        t = np.zeros((2, L.SIP_DIM, L.SIP_DIM))
        for k, c in enumerate((self.sip_a, self.sip_b)):
            c = np.asarray(c, dtype=np.float64)
            if c.ndim != 2 or c.shape[0] != c.shape[1] or c.shape[0] > L.SIP_DIM:
                raise ValueError("WCS: SIP coefficient arrays must be square, order <= %d" % L.SIP_MAX_ORDER)
            m = c.shape[0] - 1
            for p in range(m + 1):
                for q in range(m + 1):
                    if p + q <= m:
                        t[k, p, q] = c[p, q]
                    elif c[p, q] != 0.0:
                        raise ValueError("WCS: SIP term %s_%d_%d exceeds the polynomial order %d"
                                         % ('AB'[k], p, q, m))
        return t

    def _sip_op(self, direction):
        """TFM_OP_SIP with its coefficient table resident on the current device (uploaded once per
        device and table contents; the tensor lives as long as this object)."""
        import struct
        t = self._table_on_device(self.sip_table(), "sip")
        op = L.TfmOp()
        op.kind, op.dir, op.flags = L.OP_SIP, direction, 0
        op.p[0] = struct.unpack("d", struct.pack("q", int(t.data_ptr())))[0]
        op.p[1], op.p[2] = float(np.asarray(self.sip_a).shape[0] - 1), float(np.asarray(self.sip_b).shape[0] - 1)
        op.p[3], op.p[4] = float(self.crpix[0]), float(self.crpix[1])
        return op

    def __getstate__(self):
        # the device-resident coefficient tables are a cache: never pickled (another process uploads its own)
        st = dict(self.__dict__)
        st['_sip_dev'] = {}
        return st

    def __deepcopy__(self, memo):
        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = {} if k == '_sip_dev' else copy.deepcopy(v, memo)
        return new

    def copy(self):
        return copy.deepcopy(self)

    def linear_matrix(self):
        """M_ij = CDELT_i * PC_ij (paper I eq. 1-3; wcslib's `piximg`)."""
        return np.asarray(self.cdelt, dtype=np.float64)[:, None] * np.asarray(self.pc, dtype=np.float64)

    def projection(self):
        """-> None for purely linear axes, else (code, lng_axis, lat_axis, phi0, theta0)."""
        if self.naxis < 2:
            return None
        c0, c1 = self.ctype[0].upper(), self.ctype[1].upper()
        def split(c):
            if len(c) >= 8 and c[4] == '-':
                return c[:4], c[5:8]
            return None, None
        p0, j0 = split(c0)
        p1, j1 = split(c1)
        if p0 is None or p1 is None:
            return None
        if j0 != j1:
            raise ValueError("WCS: CTYPE1/CTYPE2 name different projections (%s, %s)" % (c0, c1))
        if j0 not in _PROJ:
            raise NotImplementedError("transform_b200: WCS projection '%s' is not on the GPU path "
                                      "(%s are)" % (j0, ", ".join(sorted(_PROJ))))
        is_lng0 = p0 in _LNG_PREFIX or p0.endswith("LN") or p0.endswith("LON")
        is_lat1 = p1 in _LAT_PREFIX or p1.endswith("LT") or p1.endswith("LAT")
        if not (is_lng0 and is_lat1):
            raise NotImplementedError("transform_b200: WCS axes must be (longitude, latitude) in that order")
        code, phi0, theta0 = _PROJ[j0]
        return code, 0, 1, phi0, theta0

    def celestial_pole(self):
        """(alpha_p, delta_p, phi_p): paper II section 2.4, eq. (8)-(10), LONPOLE/LATPOLE defaults."""
        code, _, _, phi0, theta0 = self.projection()
        alpha0, delta0 = self.crval[0], self.crval[1]
        phi_p = self.lonpole
        if phi_p is None:
            phi_p = phi0 + (0.0 if delta0 >= theta0 else 180.0)
        theta_p = 90.0 if self.latpole is None else self.latpole
        if theta0 == 90.0:
            return alpha0, delta0, phi_p
        # eq. (8)
        dphi = math.radians(phi_p - phi0)
        t0, d0 = math.radians(theta0), math.radians(delta0)
        a = math.degrees(math.atan2(math.sin(t0), math.cos(t0) * math.cos(dphi)))
        denom = math.sqrt(max(0.0, 1.0 - (math.cos(t0) * math.sin(dphi)) ** 2))
        arg = math.sin(d0) / denom if denom > 0 else 2.0
        if abs(arg) > 1.0 + 1e-12:
            raise ValueError("WCS: invalid LONPOLE/CRVAL combination (no solution for the celestial pole)")
        b = math.degrees(math.acos(max(-1.0, min(1.0, arg))))
        cands = [c for c in (a + b, a - b) if -90.0 - 1e-10 <= c <= 90.0 + 1e-10]
        if not cands:
            cands = [((a + b + 90.0) % 360.0) - 90.0]
        delta_p = min(cands, key=lambda c: abs(c - theta_p))
        delta_p = max(-90.0, min(90.0, delta_p))
        # eq. (9), (10)
        if abs(abs(delta0) - 90.0) < 1e-12:
            alpha_p = alpha0
        elif abs(delta_p - 90.0) < 1e-12:
            alpha_p = alpha0 + phi_p - phi0 - 180.0
        elif abs(delta_p + 90.0) < 1e-12:
            alpha_p = alpha0 - phi_p + phi0
        else:
            dp = math.radians(delta_p)
            ya = math.sin(dphi) * math.cos(t0) / math.cos(d0)
            xa = (math.sin(t0) - math.sin(dp) * math.sin(d0)) / (math.cos(dp) * math.cos(d0))
            alpha_p = alpha0 - math.degrees(math.atan2(ya, xa))
        return alpha_p, delta_p, phi_p

    def rotation_matrix(self):
        """3x3 R with l = R.m: native unit vector m -> celestial unit vector l (paper II eq. 2/5).
        Built from the images of the three native basis vectors under eq. (2)."""
        alpha_p, delta_p, phi_p = self.celestial_pole()
        R = np.zeros((3, 3))
        for k, (phi, theta) in enumerate(((0.0, 0.0), (90.0, 0.0), (0.0, 90.0))):
            a, d = _native_to_celestial_point(phi, theta, alpha_p, delta_p, phi_p)
            R[:, k] = _sph_to_vec(a, d)
        return R

    def lower(self, direction):
        """-> [TfmOp]: fwd = pixel (origin 0) -> world, rev = world -> pixel (core.py:1719-1747)."""
        if self.naxis != 2:
            raise NotImplementedError("transform_b200: WCS transforms are on the GPU path for 2 axes")
        op = L.TfmOp()
        op.kind, op.dir = L.OP_WCS, direction
        M = self.linear_matrix()
        Minv = np.linalg.inv(M)
        op.p[0], op.p[1] = float(self.crpix[0]), float(self.crpix[1])
        op.p[2], op.p[3], op.p[4], op.p[5] = (float(M[0, 0]), float(M[0, 1]),
                                              float(M[1, 0]), float(M[1, 1]))
        op.p[6], op.p[7], op.p[8], op.p[9] = (float(Minv[0, 0]), float(Minv[0, 1]),
                                              float(Minv[1, 0]), float(Minv[1, 1]))
        op.p[10], op.p[11] = float(self.crval[0]), float(self.crval[1])
        proj = self.projection()
        if proj is None:
            op.flags = L.WCS_PROJ_NONE
        else:
            op.flags = proj[0]
            R = self.rotation_matrix()
            for i in range(3):
                for j in range(3):
                    op.p[12 + 3 * i + j] = float(R[i, j])
            if proj[0] == L.WCS_PROJ_CEA:
                op.p[21] = float(self.cea_lambda)
            elif proj[0] == L.WCS_PROJ_SIN:
                op.p[21], op.p[22] = float(self.sin_slant[0]), float(self.sin_slant[1])
        ops = [op]
        if self.tpv is not None and proj is not None:
            # wcslib applies the TPV polynomial to the intermediate world coordinates, between the linear part
            # and the gnomonic projection: {linear part with CRVAL = 0 and no projection, TPV, projection part
            # with CRPIX = 1 and a unit matrix}
            lin, prj = L.TfmOp(), L.TfmOp()
            lin.kind, lin.dir, lin.flags = L.OP_WCS, direction, L.WCS_PROJ_NONE
            prj.kind, prj.dir, prj.flags = L.OP_WCS, direction, op.flags
            for k in range(L.OP_NPARAM):
                lin.p[k] = op.p[k] if k < 10 else 0.0
                prj.p[k] = op.p[k] if k >= 10 else 0.0
            prj.p[0] = prj.p[1] = 1.0
            prj.p[2] = prj.p[5] = prj.p[6] = prj.p[9] = 1.0
            ops = [lin, self._tpv_op(direction), prj] if direction == L.FWD else [prj, self._tpv_op(direction), lin]
        if self.has_sip():
            # all_pix2world: pix2foc (SIP) then the core; all_world2pix: the core, then the SIP
            # fixed-point inversion (core.py:1727, 1742)
            return [self._sip_op(direction)] + ops if direction == L.FWD else ops + [self._sip_op(direction)]
        return ops

    def to_header(self, relax=None):
        """The keys astropy's WCS.to_header() emits that core.py relies on (core.py:1074, 1208-1234).

        As in astropy, the default (the way core.py calls it) writes standard keywords only: the SIP
        coefficients are left out and the '-SIP' suffix is taken off CTYPE -- so the header remap()
        writes for a SIP template describes the undistorted frame (a property of the reference, kept).
        relax=True writes A_ORDER / B_ORDER / A_p_q / B_p_q and keeps the suffix."""
        do_sip = bool(relax) and self.has_sip()
        h = Header()
        h['WCSAXES'] = self.naxis
        for i in range(self.naxis):
            h['CRPIX%d' % (i + 1)] = float(self.crpix[i])
        pc = np.asarray(self.pc)
        for i in range(self.naxis):
            for j in range(self.naxis):
                if pc[i, j] != (1.0 if i == j else 0.0):
                    h['PC%d_%d' % (i + 1, j + 1)] = float(pc[i, j])
        for i in range(self.naxis):
            h['CDELT%d' % (i + 1)] = float(self.cdelt[i])
        for i in range(self.naxis):
            if self.cunit[i]:
                h['CUNIT%d' % (i + 1)] = self.cunit[i]
        for i in range(self.naxis):
            if self.ctype[i]:
                c = self.ctype[i]
                if self.has_sip() and not do_sip and c.upper().endswith('-SIP'):
                    c = c[:-4]
                h['CTYPE%d' % (i + 1)] = c
        for i in range(self.naxis):
            h['CRVAL%d' % (i + 1)] = float(self.crval[i])
        if self.projection() is not None:
            _, _, phi_p = self.celestial_pole()
            h['LONPOLE'] = float(phi_p)
            h['LATPOLE'] = float(90.0 if self.latpole is None else self.latpole)
            if self.projection()[0] == L.WCS_PROJ_CEA and self.cea_lambda != 1.0:
                h['PV2_1'] = float(self.cea_lambda)
            if self.projection()[0] == L.WCS_PROJ_SIN and self.sin_slant != (0.0, 0.0):
                h['PV2_1'], h['PV2_2'] = float(self.sin_slant[0]), float(self.sin_slant[1])
            if self.tpv is not None:
                for i in range(2):
                    for m in range(L.TPV_NCOEF):
                        if self.tpv[i][m] != (1.0 if m == 1 else 0.0):
                            h['PV%d_%d' % (i + 1, m)] = float(self.tpv[i][m])
        if do_sip:
            for L_, c in (('A', np.asarray(self.sip_a)), ('B', np.asarray(self.sip_b))):
                m = c.shape[0] - 1
                h['%s_ORDER' % L_] = m
                for p in range(m + 1):
                    for q in range(m + 1 - p):
                        if c[p, q] != 0.0:
                            h['%s_%d_%d' % (L_, p, q)] = float(c[p, q])
        return h
