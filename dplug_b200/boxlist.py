"""Rectangle-list helpers of the tile scheduler, mirroring gui/dplug/gui/boxlist.d (host logic inside libb2pbr)."""
from . import _lib


def tileAreas(areas, maxWidth, maxHeight):
    """boxlist.d:139-167 — split each box into tiles of at most maxWidth x maxHeight, row-major."""
    l = _lib.lib()
    arr = _lib.boxes(areas)
    n = l.b2_tile_areas(arr, len(areas), maxWidth, maxHeight, None, 0)
    out = (_lib.box2i * max(1, n))()
    l.b2_tile_areas(arr, len(areas), maxWidth, maxHeight, out, n)
    return [out[i].astuple() for i in range(n)]


def removeOverlappingAreas(boxes):
    """boxlist.d:54-118 — same coverage, no overlaps."""
    l = _lib.lib()
    arr = _lib.boxes(boxes)
    cap = 16 * max(1, len(boxes)) + 64
    out = (_lib.box2i * cap)()
    n = l.b2_remove_overlapping_areas(arr, len(boxes), out, cap)
    return [out[i].astuple() for i in range(min(n, cap))]


def impactOnNextLevel(quality, area, w, h):
    """Mipmap.impactOnNextLevel — graphics/dplug/graphics/mipmap.d:623-645."""
    return _lib.lib().b2_impact_on_next_level(int(quality), _lib.box2i(*area), w, h).astuple()


def growRect(rect, margin, width, height):
    """GUIGraphics.convertPBRLayerRectToRawLayerRect — gui/dplug/gui/graphics.d:981-1002."""
    return _lib.lib().b2_grow_rect(_lib.box2i(*rect), margin, width, height).astuple()


def haveNoOverlap(boxes):
    """boxlist.d:171-186"""
    return bool(_lib.lib().b2_have_no_overlap(_lib.boxes(boxes), len(boxes)))


class DirtyRectList:
    """gui/dplug/gui/boxlist.d:200-303 — the widgets' dirty-rectangle accumulator (never holds overlapping boxes)."""

    def __init__(self):
        import ctypes as C
        self._l = _lib.lib()
        h = C.c_void_p()
        _lib.check(self._l.b2_dirty_rects_create(C.byref(h)))
        self._h = h

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.b2_dirty_rects_destroy(self._h)
            self._h = None

    def addRect(self, rect):
        _lib.check(self._l.b2_dirty_rects_add(self._h, _lib.box2i(*rect)))

    def isEmpty(self):
        return bool(self._l.b2_dirty_rects_is_empty(self._h))

    def pullAllRectangles(self):
        n = self._l.b2_dirty_rects_pull_all(self._h, None, 0)
        out = (_lib.box2i * max(1, n))()
        n = self._l.b2_dirty_rects_pull_all(self._h, out, n)
        return [out[i].astuple() for i in range(n)]
