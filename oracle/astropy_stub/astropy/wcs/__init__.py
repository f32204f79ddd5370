"""astropy.wcs stand-in: class name only (the real arithmetic is wcslib; not available)."""


class WCS:
    def __init__(self, *a, **k):
        raise NotImplementedError("astropy.wcs is not available in this image (stub)")


class _NS:
    pass


wcs = _NS()
wcs.WCS = WCS
