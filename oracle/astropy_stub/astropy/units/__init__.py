"""astropy.units stand-in: angular units only (see ../__init__.py)."""
import math


class Unit:
    def __init__(self, name, to_rad):
        self.name = name
        self._to_rad = to_rad

    def __rmul__(self, other):
        return Quantity(other, self)

    def __mul__(self, other):
        return Quantity(other, self)

    def __str__(self):
        return self.name

    def __repr__(self):
        return "Unit(%r)" % self.name


class Quantity:
    def __init__(self, value, unit):
        self.value = value
        self.unit = unit

    def to(self, unit):
        if unit._to_rad is None or self.unit._to_rad is None:
            raise ValueError("not an angular unit")
        return Quantity(self.value * self.unit._to_rad / unit._to_rad, unit)


radian = rad = Unit("rad", 1.0)
deg = degree = Unit("deg", math.pi / 180.0)
arcmin = arcminute = Unit("arcmin", math.pi / 180.0 / 60.0)
arcsec = arcsecond = Unit("arcsec", math.pi / 180.0 / 3600.0)
meter = m = Unit("m", None)


class _NS:
    pass


core = _NS()
core.Unit = Unit
quantity = _NS()
quantity.Quantity = Quantity
