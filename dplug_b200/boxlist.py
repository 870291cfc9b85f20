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
