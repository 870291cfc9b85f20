"""GPU parity of the FUSED mipmap chain (dplug_b200/csrc/mip_fused.cu, b2mip_generate_levels): up to three
generateLevel calls per launch must give exactly what the reference's level-by-level chain gives
(graphics/dplug/graphics/mipmap.d:534-618 called per level from gui/dplug/gui/graphics.d:1378-1419) — on full frames,
on dirty rectangles, and when the stored levels are stale outside the rectangles.  Bit-exact."""
import itertools

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BOX, CUBIC, PREMUL = 0, 1, 2


def _rand(rng, color, w, h, sparse_alpha=False):
    if color == 0:
        img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        if sparse_alpha:   # emissive channel mostly 0, as on a real UI: exercises the all-zero-alpha shortcut
            lit = rng.integers(0, 20, size=(max(h // 8, 1) + 1, max(w // 8, 1) + 1)) == 0
            mask = np.kron(lit, np.ones((8, 8), dtype=bool))[:h, :w]
            img[..., 3] = np.where(mask, img[..., 3], 0)
        return img
    return rng.integers(0, 65536, size=(h, w), dtype=np.uint16)


def _pair(b2, oracle, color, maxLevel, w, h):
    return b2.Mipmap(color, maxLevel, w, h), oracle.Mipmap(color, maxLevel, w, h)


def _fill_stale(rng, g, o, color):
    """random (mutually inconsistent) content in every level of both pyramids"""
    for level in range(g.numLevels()):
        shape = g.download(level).shape
        img = _rand(rng, color, shape[1], shape[0])
        g.upload(level, img); o.upload(level, img)


def _oracle_chain(o, first, qualities, rect):
    rects = []
    for k, q in enumerate(qualities):
        rect = o.generateNextLevel(q, rect, first + k)
        rects.append(rect)
    return rects


def _assert_levels_equal(g, o, what):
    for level in range(1, g.numLevels()):
        a, b = g.download(level), o.download(level)
        if not np.array_equal(a, b):
            bad = np.argwhere((a != b).reshape(a.shape[0], a.shape[1], -1).any(axis=-1))
            raise AssertionError(f"{what}: level {level} differs at {len(bad)} texels, first (y, x) = {bad[0].tolist()}")


SIZES = [(620, 330), (515, 389), (64, 48), (37, 21), (5, 9), (3, 2), (1, 7), (255, 257), (130, 66), (1030, 777), (129, 1025)]


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("color,recipe", [
    (0, (PREMUL, CUBIC, CUBIC, CUBIC, CUBIC)),           # diffuse, graphics.d:1380-1384
    (1, (BOX, BOX, CUBIC, CUBIC)),                        # depth, graphics.d:1414
    (0, (BOX, BOX, CUBIC, CUBIC, CUBIC, CUBIC, CUBIC)),   # generateMipmaps(box), mipmap.d:517-528 (skybox)
    (0, (CUBIC, BOX, PREMUL)),
    (1, (CUBIC, CUBIC, BOX, CUBIC, BOX)),
])
def test_full_frame_recipes(b2, oracle, w, h, color, recipe):
    rng = np.random.default_rng(w * 31 + h * 7 + color + len(recipe))
    g, o = _pair(b2, oracle, color, len(recipe), w, h)
    assert g.numLevels() == o.numLevels()
    n = g.numLevels() - 1
    if n < 1:
        return
    src = _rand(rng, color, w, h, sparse_alpha=(w % 2 == 0))
    _fill_stale(rng, g, o, color)
    g.upload(0, src); o.upload(0, src)
    q = list(recipe[:n])
    assert g.generateLevels(1, q, (0, 0, w, h)) == _oracle_chain(o, 1, q, (0, 0, w, h))
    _assert_levels_equal(g, o, (w, h, color, recipe))


@pytest.mark.parametrize("color", [0, 1])
@pytest.mark.parametrize("n", [1, 2, 3])
def test_every_quality_combination(b2, oracle, color, n):
    """all quality tuples of length n (one launch each), odd sizes so the edge-extension of the tiles is exercised"""
    quals = (BOX, CUBIC, PREMUL) if color == 0 else (BOX, CUBIC)
    rng = np.random.default_rng(1000 + 10 * color + n)
    w, h = 301, 167
    src = _rand(rng, color, w, h)
    for q in itertools.product(quals, repeat=n):
        g, o = _pair(b2, oracle, color, n, w, h)
        _fill_stale(rng, g, o, color)
        g.upload(0, src); o.upload(0, src)
        assert g.generateLevels(1, list(q), (0, 0, w, h)) == _oracle_chain(o, 1, list(q), (0, 0, w, h))
        _assert_levels_equal(g, o, q)


@pytest.mark.parametrize("color,recipe", [
    (0, (PREMUL, CUBIC, CUBIC, CUBIC, CUBIC)),
    (1, (BOX, BOX, CUBIC, CUBIC)),
    (0, (BOX, CUBIC, BOX, CUBIC)),
])
def test_dirty_rects_on_stale_pyramids(b2, oracle, color, recipe):
    """Random dirty rectangles; every level starts with random content, so a texel outside a level's update rectangle
    must be read as stored (not recomputed) by the next step, and nothing outside the rectangles may change."""
    rng = np.random.default_rng(77 + color + len(recipe))
    w, h = 333, 210
    g, o = _pair(b2, oracle, color, len(recipe), w, h)
    for it in range(40):
        _fill_stale(rng, g, o, color)
        x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
        if it % 4 == 0:    # small rectangles (a knob redraw)
            r = (x0, y0, min(w, x0 + int(rng.integers(1, 40))), min(h, y0 + int(rng.integers(1, 40))))
        elif it % 4 == 1:  # rectangles touching the right / bottom edge (odd last column and row)
            r = (x0, y0, w, h)
        else:
            r = (x0, y0, int(rng.integers(x0, w + 1)), int(rng.integers(y0, h + 1)))   # may be degenerate
        q = list(recipe)
        assert g.generateLevels(1, q, r) == _oracle_chain(o, 1, q, r), r
        _assert_levels_equal(g, o, (r, recipe))


def test_start_above_level_one(b2, oracle):
    rng = np.random.default_rng(8)
    w, h = 411, 275
    for color in (0, 1):
        g, o = _pair(b2, oracle, color, 6, w, h)
        _fill_stale(rng, g, o, color)
        r = (13, 9, 97, 88)    # a rectangle of level 2
        q = [CUBIC, BOX, CUBIC, CUBIC]
        assert g.generateLevels(3, q, r) == _oracle_chain(o, 3, q, r)
        _assert_levels_equal(g, o, color)


def test_fused_equals_level_by_level_large(b2):
    """Size-independent check at a size the oracle would take long on: fused launches == one launch per level."""
    rng = np.random.default_rng(3)
    from dplug_b200 import _lib
    l = _lib.lib()
    for color, recipe, (w, h) in [(0, (PREMUL, CUBIC, CUBIC, CUBIC, CUBIC, CUBIC, CUBIC, CUBIC), (4099, 2050)),
                                   (1, (BOX, BOX, CUBIC, CUBIC), (3840, 2160))]:
        src = _rand(rng, color, w, h, sparse_alpha=True)
        a, b = b2.Mipmap(color, len(recipe), w, h), b2.Mipmap(color, len(recipe), w, h)
        a.upload(0, src); b.upload(0, src)
        was = l.b2_set_mip_fusion(2)   # fuse every level, also the large ones
        before = a.kernelLaunches()
        ra = a.generateLevels(1, list(recipe), (0, 0, w, h))
        assert a.kernelLaunches() - before == (len(recipe) + 2) // 3
        l.b2_set_mip_fusion(0)
        try:
            rb = b.generateLevels(1, list(recipe), (0, 0, w, h))
            assert b.kernelLaunches() == len(recipe)
        finally:
            l.b2_set_mip_fusion(was)
        assert ra == rb
        for level in range(1, a.numLevels()):
            assert np.array_equal(a.download(level), b.download(level)), (color, level)


def test_generate_levels_errors(b2):
    g = b2.Mipmap(0, 3, 64, 64)
    with pytest.raises(b2.B2Error):
        g.generateLevels(1, [BOX, BOX, BOX, BOX], (0, 0, 64, 64))   # only levels 1..3 exist
    with pytest.raises(b2.B2Error):
        g.generateLevels(1, [7], (0, 0, 64, 64))
    z = b2.Mipmap(1, 3, 64, 64)
    with pytest.raises(b2.B2Error):
        z.generateLevels(1, [BOX, PREMUL], (0, 0, 64, 64))
    assert g.generateLevels(1, [], (0, 0, 64, 64)) == []
