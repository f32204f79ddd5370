"""transform_b200 -- B200-native drop-in for the image-resampling hot path of drzowie/transform.

Same public surface as the reference package for this path (transform/__init__.py:6-13 star-imports
core, helpers and basic): Transform, Identity, Inverse, Composition, Wrap, Linear, Scale, Rotation,
Offset, Radial, WCS, DataWrapper, interpND, interpND_grid, svd2x2, svc2x2.  Every arithmetic step
runs in libtfm_b200.so (hand-written CUDA for sm_100a); importing the package needs neither a GPU
nor the library, calling into the path does -- and fails loudly otherwise.
"""
from .core import Transform, Identity, PlusOne_, Inverse, Composition, Wrap       # noqa: F401
from .basic import (Linear, Scale, Rotation, Offset, Radial, Quadpin, Cubicpin,   # noqa: F401
                    Poly, Quarticpin)
from .helpers import (interpND, interpND_grid, svd2x2, svc2x2, simple_jacobian,   # noqa: F401
                      jump_detect, jacobian)
from .fitswcs import (Header, WCSParams, read_fits, FITSHeaderFromDict,           # noqa: F401
                      decode_fits_data)
from .remap import DataWrapper, WCS                                               # noqa: F401
from .pipeline import AsyncResampler                                              # noqa: F401

__version__ = "0.1.0"
