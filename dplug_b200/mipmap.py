"""Device-resident Mipmap!RGBA / Mipmap!L16 (mirror of graphics/dplug/graphics/mipmap.d:30-646)."""
import ctypes as C
import numpy as np
from . import _lib

RGBA8 = 0
L16 = 1


class Quality:  # mipmap.d:36-46
    box = 0
    cubic = 1
    boxAlphaCovIntoPremul = 2


class Mipmap:
    """`Mipmap(color, maxLevel, w, h)`: level l is (w>>l) x (h>>l); pixels live in HBM."""

    def __init__(self, color, maxLevel, w, h, device=0, band=None, _handle=None):
        self._l = _lib.lib()
        self.color = color
        self._owned = _handle is None
        if _handle is not None:
            self._h = C.c_void_p(_handle)
        else:
            h_ = C.c_void_p()
            if band is None:
                _lib.check(self._l.b2mip_create(C.byref(h_), device, color, maxLevel, w, h))
            else:
                y0, y1, up, down = band
                _lib.check(self._l.b2mip_create_band(C.byref(h_), device, color, maxLevel, w, h, y0, y1, up, down))
            self._h = h_

    def close(self):
        if self._owned and self._h:
            self._l.b2mip_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_stream(self, cuda_stream):
        _lib.check(self._l.b2mip_set_stream(self._h, C.c_void_p(cuda_stream)))

    def numLevels(self):
        return self._l.b2mip_num_levels(self._h)

    def width(self):
        return self._l.b2mip_width(self._h)

    def height(self):
        return self._l.b2mip_height(self._h)

    def levelInfo(self, level):
        ref = _lib.imageref()
        first = C.c_int()
        _lib.check(self._l.b2mip_level(self._h, level, C.byref(ref), C.byref(first)))
        return ref.w, ref.h, ref.pitch, ref.pixels, first.value

    def _dtype(self):
        return np.uint8 if self.color == RGBA8 else np.uint16

    def _shape(self, level):
        w, h, _, _, _ = self.levelInfo(level)
        return (h, w, 4) if self.color == RGBA8 else (h, w)

    def upload(self, level, image, rects=None):
        """host -> device for `rects` (default: the whole level).  `image` is the full level: (h,w,4) u8 or (h,w) u16."""
        image = np.ascontiguousarray(image, dtype=self._dtype())
        assert image.shape == self._shape(level), (image.shape, self._shape(level))
        h, w = image.shape[0], image.shape[1]
        if rects is None:
            rects = [(0, 0, w, h)]
        pitch = image.strides[0]
        _lib.check(self._l.b2mip_upload(self._h, level, image.ctypes.data_as(C.c_void_p), pitch, _lib.boxes(rects), len(rects)))
        _lib.check(self._l.b2mip_sync(self._h))  # the numpy temporary may go away

    def uploadRows(self, level, rows, row0, x0=0):
        """host -> device for a block of rows: `rows[k]` is global row `row0 + k` of the level (band objects)."""
        rows = np.ascontiguousarray(rows, dtype=self._dtype())
        n, w = rows.shape[0], rows.shape[1]
        if n == 0:
            return
        pitch = rows.strides[0]
        origin = rows.ctypes.data - row0 * pitch - x0 * (4 if self.color == RGBA8 else 2)   # address of global pixel (0, 0)
        _lib.check(self._l.b2mip_upload(self._h, level, C.c_void_p(origin), pitch, _lib.boxes([(x0, row0, x0 + w, row0 + n)]), 1))
        _lib.check(self._l.b2mip_sync(self._h))

    def downloadRows(self, level, row0, row1):
        w, h, _, _, _ = self.levelInfo(level)
        out = np.zeros((row1 - row0, w, 4) if self.color == RGBA8 else (row1 - row0, w), dtype=self._dtype())
        if row1 > row0:
            origin = out.ctypes.data - row0 * out.strides[0]
            _lib.check(self._l.b2mip_download(self._h, level, C.c_void_p(origin), out.strides[0], _lib.boxes([(0, row0, w, row1)]), 1))
        return out

    def download(self, level, rects=None, out=None):
        shape = self._shape(level)
        if out is None:
            out = np.zeros(shape, dtype=self._dtype())
        if rects is None:
            rects = [(0, 0, shape[1], shape[0])]
        _lib.check(self._l.b2mip_download(self._h, level, out.ctypes.data_as(C.c_void_p), out.strides[0], _lib.boxes(rects), len(rects)))
        return out

    def generateNextLevel(self, quality, updateRectPreviousLevel, level):
        out = _lib.box2i()
        _lib.check(self._l.b2mip_generate_next_level(self._h, int(quality), _lib.box2i(*updateRectPreviousLevel), level, C.byref(out)))
        return out.astuple()

    def generateLevel(self, level, quality, updateRect):
        _lib.check(self._l.b2mip_generate_level(self._h, level, int(quality), _lib.box2i(*updateRect)))

    def generateMipmaps(self, quality):
        _lib.check(self._l.b2mip_generate_mipmaps(self._h, int(quality)))

    def generateChain(self, firstQuality, cubicFromLevel):
        _lib.check(self._l.b2mip_generate_chain(self._h, int(firstQuality), int(cubicFromLevel)))

    def generateLevels(self, firstLevel, qualities, updateRectPreviousLevel):
        """len(qualities) successive generateNextLevel calls in fused launches; returns the updated rectangles."""
        n = len(qualities)
        q = (C.c_int * max(n, 1))(*[int(v) for v in qualities])
        out = (_lib.box2i * max(n, 1))()
        _lib.check(self._l.b2mip_generate_levels(self._h, int(firstLevel), n, q, _lib.box2i(*updateRectPreviousLevel), out))
        return [out[k].astuple() for k in range(n)]

    def _sample(self, kind, level, x, y):
        level = np.ascontiguousarray(np.broadcast_to(np.asarray(level, dtype=np.float32), np.shape(x)), dtype=np.float32).ravel()
        x = np.ascontiguousarray(x, dtype=np.float32).ravel()
        y = np.ascontiguousarray(y, dtype=np.float32).ravel()
        out = np.zeros((x.size, 4), dtype=np.float32)
        _lib.check(self._l.b2mip_sample(self._h, kind, x.size, level.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p),
                                        y.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return out if self.color == RGBA8 else out[:, 0]

    def linearSample(self, level, x, y):
        return self._sample(0, level, x, y)

    def cubicSample(self, level, x, y):
        return self._sample(1, level, x, y)

    def linearMipmapSample(self, level, x, y):
        return self._sample(2, level, x, y)

    def sync(self):
        _lib.check(self._l.b2mip_sync(self._h))

    def kernelLaunches(self):
        return int(self._l.b2mip_kernel_launches(self._h))
