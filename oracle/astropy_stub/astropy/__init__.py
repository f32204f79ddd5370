"""Minimal stand-in for the `astropy` package -- TEST INFRASTRUCTURE ONLY.

astropy is not installed in this image (and there is no network).  The reference's
core.py/basic.py import `astropy`, `astropy.io.fits`, `astropy.wcs`, `astropy.units`
at module scope (core.py:6-9, basic.py:9) but the resample()/apply() hot path only
touches `astropy.units` for angular-unit parsing (basic.py:444-458, 630-644).  This stub
provides exactly that surface so the *unmodified* reference can be imported in the
authoring container to generate golden vectors (tests/golden/make_golden.py).  It is
never imported by transform_b200.
"""
from . import units      # noqa: F401
from . import io         # noqa: F401
from . import wcs        # noqa: F401
