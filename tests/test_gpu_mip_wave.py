"""GPU parity of the whole-pyramid WAVEFRONT kernel (dplug_b200/csrc/mip_wave.cu): every level of a
Mipmap.generateMipmaps / GUIGraphics.regenerateMipmaps chain (graphics/dplug/graphics/mipmap.d:515-618,
gui/dplug/gui/graphics.d:1378-1419) in ONE persistent launch whose levels advance together down the image, must give
exactly what the reference's level-by-level chain gives — on full frames, on large dirty rectangles of stale pyramids,
when the same plan is launched again and again (the kernel resets its own counters), and at sizes the oracle would take
long on (against the one-launch-per-level kernels).  Bit-exact."""
import numpy as np
import pytest

from tests.test_gpu_mip_fused import BOX, CUBIC, PREMUL, _assert_levels_equal, _fill_stale, _oracle_chain, _pair, _rand

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _wave_on():
    """the wavefront launch is an opt-in (measured slower than the per-level launches): switched on for these tests"""
    from dplug_b200 import _lib
    was = _lib.lib().b2_set_mip_wave(1)
    yield
    _lib.lib().b2_set_mip_wave(was)

DIFFUSE = (PREMUL, CUBIC, CUBIC, CUBIC, CUBIC)           # graphics.d:1380-1384
DEPTH = (BOX, BOX, CUBIC, CUBIC)                         # graphics.d:1414
SKYBOX = (BOX, BOX, CUBIC, CUBIC, CUBIC, CUBIC, CUBIC)   # generateMipmaps(box), mipmap.d:517-528


@pytest.mark.parametrize("w,h", [(1400, 900), (1237, 1011), (2051, 769), (1030, 1030)])
@pytest.mark.parametrize("color,recipe", [(0, DIFFUSE), (1, DEPTH), (0, SKYBOX), (0, (CUBIC, BOX, PREMUL, CUBIC)), (1, (CUBIC, CUBIC, BOX, CUBIC, BOX))])
def test_full_frame_vs_oracle(b2, oracle, w, h, color, recipe):
    rng = np.random.default_rng(w * 13 + h * 5 + color + len(recipe))
    g, o = _pair(b2, oracle, color, len(recipe), w, h)
    src = _rand(rng, color, w, h, sparse_alpha=(w % 2 == 0))
    _fill_stale(rng, g, o, color)
    g.upload(0, src); o.upload(0, src)
    before = g.kernelLaunches()
    assert g.generateLevels(1, list(recipe), (0, 0, w, h)) == _oracle_chain(o, 1, list(recipe), (0, 0, w, h))
    assert g.kernelLaunches() - before == 1, "the chain did not run as one wavefront launch"
    _assert_levels_equal(g, o, (w, h, color, recipe))


@pytest.mark.parametrize("color,recipe", [(0, DIFFUSE), (1, DEPTH)])
def test_large_dirty_rects_on_stale_pyramids(b2, oracle, color, recipe):
    """Dirty rectangles large enough for the wavefront path; every level starts with random content, so texels outside a
    level's update rectangle must be read as stored, and nothing outside the rectangles may change.  More distinct
    rectangles than plan slots, so slots are recycled; every third rectangle is repeated (plan re-use)."""
    rng = np.random.default_rng(5 + color)
    w, h = 1700, 1300
    g, o = _pair(b2, oracle, color, len(recipe), w, h)
    rects = []
    for it in range(12):
        x0, y0 = int(rng.integers(0, 200)), int(rng.integers(0, 150))
        x1, y1 = int(rng.integers(1500, w + 1)), int(rng.integers(1200, h + 1))
        rects.append((x0, y0, x1, y1) if it % 4 else (x0, y0, w, h))
    waves = 0
    for it, r in enumerate(rects + rects[::3]):
        _fill_stale(rng, g, o, color)
        before = g.kernelLaunches()
        assert g.generateLevels(1, list(recipe), r) == _oracle_chain(o, 1, list(recipe), r), r
        waves += (g.kernelLaunches() - before == 1)
        _assert_levels_equal(g, o, (it, r, recipe))
    assert waves == len(rects) + len(rects[::3])


def test_start_above_level_one(b2, oracle):
    rng = np.random.default_rng(9)
    w, h = 2600, 2100
    for color in (0, 1):
        g, o = _pair(b2, oracle, color, 6, w, h)
        _fill_stale(rng, g, o, color)
        r = (3, 5, 1290, 1040)    # a rectangle of level 1
        q = [CUBIC, BOX, CUBIC, CUBIC]
        before = g.kernelLaunches()
        assert g.generateLevels(2, q, r) == _oracle_chain(o, 2, q, r)
        assert g.kernelLaunches() - before == 1
        _assert_levels_equal(g, o, color)


def test_same_plan_launched_repeatedly_and_knob(b2):
    """wavefront launch == one launch per level, at sizes the oracle would take long on; the same plan is launched four
    times on changing content (the kernel leaves its counters zeroed) and the knob switches the path off"""
    rng = np.random.default_rng(3)
    from dplug_b200 import _lib
    l = _lib.lib()
    for color, recipe, (w, h) in [(0, (PREMUL,) + (CUBIC,) * 7, (4099, 2050)), (1, DEPTH, (3840, 2160)), (0, DIFFUSE, (3840, 2160))]:
        a, b = b2.Mipmap(color, len(recipe), w, h), b2.Mipmap(color, len(recipe), w, h)
        for rep in range(4):
            src = _rand(rng, color, w, h, sparse_alpha=(rep % 2 == 0))
            a.upload(0, src); b.upload(0, src)
            before = a.kernelLaunches()
            ra = a.generateLevels(1, list(recipe), (0, 0, w, h))
            assert a.kernelLaunches() - before == 1
            l.b2_set_mip_wave(0)
            try:
                before = b.kernelLaunches()
                rb = b.generateLevels(1, list(recipe), (0, 0, w, h))
                assert b.kernelLaunches() - before > 1
            finally:
                l.b2_set_mip_wave(1)
            assert ra == rb
            for level in range(1, a.numLevels()):
                assert np.array_equal(a.download(level), b.download(level)), (color, rep, level)
