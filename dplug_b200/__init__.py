"""dplug_b200 — B200-native (sm_100a) PBR compositor + mipmap pyramid behind Dplug's ICompositor boundary.

Host-side mirror (Python, for tests and benchmarks) of the reference interface for this one path:
`Mipmap`, `PBRCompositor`, `tileAreas` ... all backed by the CUDA library libb2pbr.so through its C ABI
(include/b2pbr.h).  There is no CPU implementation in this package.
"""
from ._lib import B2Error, LIB_PATH, box2i, BoxArray  # noqa: F401
from .mipmap import Mipmap, Quality, RGBA8, L16  # noqa: F401
from .compositor import PBRCompositor, Layer, default_params, liftGammaGainContrastTable  # noqa: F401
from .resizer import ImageResizer  # noqa: F401
from .boxlist import tileAreas, removeOverlappingAreas, impactOnNextLevel, growRect, haveNoOverlap, DirtyRectList  # noqa: F401
