"""astropy.io.fits stand-in: class names only, so isinstance() checks evaluate False."""


class _NS:
    pass


class Header(dict):
    pass


class HDUList(list):
    pass


class PrimaryHDU:
    pass


class _HeaderCommentaryCards(list):
    pass


header = _NS()
header.Header = Header
header._HeaderCommentaryCards = _HeaderCommentaryCards
hdu = _NS()
hdu.hdulist = _NS()
hdu.hdulist.HDUList = HDUList
hdu.image = _NS()
hdu.image.PrimaryHDU = PrimaryHDU
