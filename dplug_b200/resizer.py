"""Host-side mirror of dplug:graphics ImageResizer (graphics/dplug/graphics/resizer.d) over libb2pbr.so: the resampling
runs on the GPU (csrc/resizer.cu), with the semantics of the reference's in-tree stb_image_resize port."""
import ctypes as C

import numpy as np

from . import _lib

RESIZE_DIFFUSE, RESIZE_MATERIAL, RESIZE_DEPTH, RESIZE_GENERIC_L16, RESIZE_SMOOTHER = range(5)


class ImageResizer:
    """`resizer.resizeImageDiffuse(input, output)` etc.: `output` is a preallocated array of the wanted size ((h, w, 4)
    uint8 for the RGBA functions, (h, w) uint16 for the L16 ones); it is filled and returned."""

    def __init__(self, device=0):
        self._l = _lib.lib()
        self._device = device

    def _resize(self, kind, src, out, dtype):
        src = np.ascontiguousarray(src, dtype=dtype)
        assert out.dtype == dtype and out.flags.c_contiguous and out.shape[2:] == src.shape[2:]
        _lib.check(self._l.b2_resize_image(self._device, kind, src.ctypes.data_as(C.c_void_p), src.shape[1], src.shape[0], src.strides[0],
                                           out.ctypes.data_as(C.c_void_p), out.shape[1], out.shape[0], out.strides[0]))
        return out

    def resizeImageDiffuse(self, input, output):        # resizer.d:346-372
        return self._resize(RESIZE_DIFFUSE, input, output, np.uint8)

    def resizeImageMaterial(self, input, output):       # resizer.d:374-376 (alias of resizeImageGeneric)
        return self._resize(RESIZE_MATERIAL, input, output, np.uint8)

    def resizeImageDepth(self, input, output):          # resizer.d:171-197
        return self._resize(RESIZE_DEPTH, input, output, np.uint16)

    def resizeImageGeneric(self, input, output):        # resizer.d:39-64 (RGBA), 144-169 (L16)
        if np.asarray(input).ndim == 2:
            return self._resize(RESIZE_GENERIC_L16, input, output, np.uint16)
        return self._resize(RESIZE_MATERIAL, input, output, np.uint8)

    def resizeImageSmoother(self, input, output):       # resizer.d:92-117
        return self._resize(RESIZE_SMOOTHER, input, output, np.uint8)
