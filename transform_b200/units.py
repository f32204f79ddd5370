"""Angular-unit parsing for Rotation / Radial.

The reference delegates to astropy.units (basic.py:444-458, 630-644): a unit *name* is looked up
with getattr(astropy.units, name), turned into a Quantity and converted to radians.  astropy is
not a dependency here; the angular units a transform can meaningfully take are tabulated, and
astropy Unit / Quantity objects are still accepted by duck typing (anything with `.to`).
"""
import math

_RAD_PER_UNIT = {
    "rad": 1.0, "radian": 1.0, "radians": 1.0,
    "deg": math.pi / 180.0, "degree": math.pi / 180.0, "degrees": math.pi / 180.0,
    "arcmin": math.pi / 180.0 / 60.0, "arcminute": math.pi / 180.0 / 60.0,
    "arcsec": math.pi / 180.0 / 3600.0, "arcsecond": math.pi / 180.0 / 3600.0,
    "mas": math.pi / 180.0 / 3600.0e3, "uas": math.pi / 180.0 / 3600.0e6,
    "hourangle": math.pi / 12.0, "cycle": 2 * math.pi,
}
_CANON = {"radian": "rad", "radians": "rad", "degree": "deg", "degrees": "deg",
          "arcminute": "arcmin", "arcsecond": "arcsec"}


def angular_coefficient(unit, who="Radial"):
    """Return (radians per `unit`, canonical unit name).  Raises ValueError like basic.py:453, 458."""
    if hasattr(unit, "to"):                       # astropy Unit or Quantity (duck-typed)
        try:
            import astropy.units as u
            q = unit if hasattr(unit, "value") else 1.0 * unit
            q = q.to(u.radian)
            return float(q.value), str(getattr(unit, "unit", unit))
        except Exception:
            raise ValueError("%s: '%s' doesn't appear to be an angular unit or quantity" % (who, unit))
    if not isinstance(unit, str):
        raise ValueError("%s: couldn't convert '%s' to an astropy unit object" % (who, unit))
    if unit not in _RAD_PER_UNIT:
        raise ValueError("%s: couldn't convert '%s' to an astropy unit object" % (who, unit))
    return _RAD_PER_UNIT[unit], _CANON.get(unit, unit)
