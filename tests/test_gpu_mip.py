"""GPU parity of the mipmap kernels against the CPU oracle, through the C ABI.  Bit-exact (integer paths and
the IEEE-exact alpha-premultiplying box)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [(64, 48), (37, 21), (5, 9), (2, 2), (3, 2), (620, 330), (1, 7), (255, 257), (130, 66)]


def _rand(rng, color, w, h):
    if color == 0:
        return rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    return rng.integers(0, 65536, size=(h, w), dtype=np.uint16)


@pytest.mark.parametrize("w,h", SIZES)
@pytest.mark.parametrize("color,quality", [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1)])
def test_generate_next_level_full(b2, oracle, w, h, color, quality):
    rng = np.random.default_rng(w * 1000 + h * 10 + color * 3 + quality)
    src = _rand(rng, color, w, h)
    g = b2.Mipmap(color, 1, w, h)
    o = oracle.Mipmap(color, 1, w, h)
    assert g.numLevels() == o.numLevels()
    if g.numLevels() < 2:
        return
    g.upload(0, src)
    o.upload(0, src)
    rg = g.generateNextLevel(quality, (0, 0, w, h), 1)
    ro = o.generateNextLevel(quality, (0, 0, w, h), 1)
    assert rg == ro
    assert np.array_equal(g.download(1), o.download(1))


@pytest.mark.parametrize("color,quality", [(0, 0), (0, 1), (0, 2), (1, 0), (1, 1)])
def test_generate_next_level_dirty_rects(b2, oracle, color, quality):
    """Random dirty rectangles: only the impacted rectangle of the next level may change, and it must match."""
    rng = np.random.default_rng(99 + color * 7 + quality)
    w, h = 203, 131
    src = _rand(rng, color, w, h)
    g = b2.Mipmap(color, 1, w, h)
    o = oracle.Mipmap(color, 1, w, h)
    g.upload(0, src); o.upload(0, src)
    sentinel = _rand(rng, color, w // 2, h // 2)
    for _ in range(25):
        g.upload(1, sentinel); o.upload(1, sentinel)
        x0, y0 = int(rng.integers(0, w)), int(rng.integers(0, h))
        r = (x0, y0, int(rng.integers(x0, w + 1)), int(rng.integers(y0, h + 1)))
        assert g.generateNextLevel(quality, r, 1) == o.generateNextLevel(quality, r, 1)
        assert np.array_equal(g.download(1), o.download(1)), r


def test_cubic_wide_rectangles(b2, oracle):
    """Rectangles at least 256 texels wide take the four-texels-per-thread cubic kernel: ragged left / right edges,
    the clamped taps at both image borders, odd source width."""
    rng = np.random.default_rng(31)
    w, h = 1301, 97
    src = _rand(rng, 0, w, h)
    g, o = b2.Mipmap(0, 1, w, h), oracle.Mipmap(0, 1, w, h)
    g.upload(0, src); o.upload(0, src)
    sentinel = _rand(rng, 0, w // 2, h // 2)
    rects = [(0, 0, w, h), (1, 0, w, h), (0, 0, w - 1, h), (3, 5, 700, 60), (517, 2, 1301, 97), (2, 0, 1299, 1), (5, 7, 521, 9)]
    for _ in range(12):
        x0 = int(rng.integers(0, w - 600)); y0 = int(rng.integers(0, h - 1))
        rects.append((x0, y0, int(rng.integers(x0 + 560, w + 1)), int(rng.integers(y0 + 1, h + 1))))
    for r in rects:
        g.upload(1, sentinel); o.upload(1, sentinel)
        assert g.generateNextLevel(1, r, 1) == o.generateNextLevel(1, r, 1)
        assert np.array_equal(g.download(1), o.download(1)), r


def test_premul_box_rounding_known_answer(b2, oracle):
    """A 2x2 quad whose green mean lands exactly on x.5 before the +0.5: ((A+B)+C)+D of singly rounded products gives
    43, a fused multiply-add anywhere in the chain (or another summation order) gives 42.  Found by the fused-chain
    tests; the expected bytes are the reference's arithmetic (mipmap.d:869-894) evaluated in numpy float32."""
    quad = np.array([[[117, 203, 90, 151], [252, 120, 196, 176]], [[233, 153, 28, 71], [115, 55, 0, 210]]], dtype=np.uint8)
    f = np.float32
    inv = f(1.0) / f(255.0)
    lin = np.zeros((4, 4), dtype=np.float32)
    for t, p in enumerate(quad.reshape(4, 4)):
        a = f(p[3]) * inv
        for k in range(3):
            lin[t, k] = f(p[k]) * inv * f(p[k]) * inv * a
        lin[t, 3] = a
    expect = [int(((lin[0, k] + lin[1, k]) + lin[2, k] + lin[3, k]) * f(0.25) * f(255.0) + f(0.5)) for k in range(4)]
    assert expect == [76, 43, 31, 152]
    # the quad at several positions (vector path, ragged edges) of a larger image
    img = np.tile(quad, (9, 11, 1))
    h, w = img.shape[:2]
    for x0 in (0, 1):
        g, o = b2.Mipmap(0, 1, w, h), oracle.Mipmap(0, 1, w, h)
        g.upload(0, img); o.upload(0, img)
        blank = np.zeros((h // 2, w // 2, 4), dtype=np.uint8)
        g.upload(1, blank); o.upload(1, blank)
        g.generateNextLevel(2, (2 * x0, 0, w, h), 1); o.generateNextLevel(2, (2 * x0, 0, w, h), 1)
        out = g.download(1)
        assert np.array_equal(out, o.download(1))
        assert (out[:, x0:] == np.array(expect, dtype=np.uint8)).all()


def test_premul_box_large_random(b2, oracle):
    """2 M random quads against the oracle (rounding-boundary cases occur about once per 10^4 texels)."""
    rng = np.random.default_rng(123)
    w, h = 4096, 2048
    img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    g, o = b2.Mipmap(0, 1, w, h), oracle.Mipmap(0, 1, w, h)
    g.upload(0, img); o.upload(0, img)
    g.generateNextLevel(2, (0, 0, w, h), 1); o.generateNextLevel(2, (0, 0, w, h), 1)
    assert np.array_equal(g.download(1), o.download(1))


def test_full_chain_diffuse_and_depth_recipes(b2, oracle):
    """regenerateMipmaps' recipes (graphics.d:1378-1419) level by level at the C1 size, odd dimensions included."""
    rng = np.random.default_rng(5)
    W, H = 620, 330
    d = _rand(rng, 0, W, H); z = _rand(rng, 1, W, H)
    gd, od = b2.Mipmap(0, 5, W, H), oracle.Mipmap(0, 5, W, H)
    gz, oz = b2.Mipmap(1, 4, W, H), oracle.Mipmap(1, 4, W, H)
    gd.upload(0, d); od.upload(0, d); gz.upload(0, z); oz.upload(0, z)
    r1 = r2 = (0, 0, W, H)
    for level in range(1, 6):
        q = 2 if level == 1 else 1
        r1 = gd.generateNextLevel(q, r1, level); r2 = od.generateNextLevel(q, r2, level)
        assert r1 == r2
        assert np.array_equal(gd.download(level), od.download(level)), level
    r1 = r2 = (0, 0, W, H)
    for level in range(1, 5):
        q = 1 if level >= 3 else 0
        r1 = gz.generateNextLevel(q, r1, level); r2 = oz.generateNextLevel(q, r2, level)
        assert np.array_equal(gz.download(level), oz.download(level)), level
    # generateChain = same recipe over the full rectangle in one call
    gd2 = b2.Mipmap(0, 5, W, H); gd2.upload(0, d); gd2.generateChain(2, 2)
    gz2 = b2.Mipmap(1, 4, W, H); gz2.upload(0, z); gz2.generateChain(0, 3)
    for level in range(1, 6):
        assert np.array_equal(gd2.download(level), od.download(level)), level
    for level in range(1, 5):
        assert np.array_equal(gz2.download(level), oz.download(level)), level


def test_generate_mipmaps_skybox_recipe(b2, oracle):
    """Mipmap(12, image).generateMipmaps(box): box for levels 1,2 then cubic (mipmap.d:517-528)."""
    from dplug_b200 import synth
    sky = synth.skybox(256, 192)
    g, o = b2.Mipmap(0, 12, 256, 192), oracle.Mipmap(0, 12, 256, 192)
    g.upload(0, sky); o.upload(0, sky)
    g.generateMipmaps(0); o.generateMipmaps(0)
    assert g.numLevels() == o.numLevels() == 8
    for level in range(g.numLevels()):
        assert np.array_equal(g.download(level), o.download(level)), level


def test_box_idempotent_on_constant_and_linear(b2):
    """Size-independent properties at a large size: a constant image stays constant through every level and
    kernel; box is linear for even values (a checksum-of-checksums style check without the oracle)."""
    W = H = 2048
    img = np.full((H, W, 4), 77, dtype=np.uint8)
    g = b2.Mipmap(0, 8, W, H)
    g.upload(0, img)
    g.generateMipmaps(0)
    for level in range(1, g.numLevels()):
        assert (g.download(level) == 77).all()
    z = np.full((H, W), 40000, dtype=np.uint16)
    gz = b2.Mipmap(1, 8, W, H)
    gz.upload(0, z)
    gz.generateChain(0, 3)
    for level in range(1, gz.numLevels()):
        assert (gz.download(level) == 40000).all()


def test_samplers(b2, oracle):
    """linearSample / cubicSample / linearMipmapSample on the device vs the oracle (bit-exact: same IEEE ops)."""
    rng = np.random.default_rng(21)
    W, H = 96, 70
    for color in (0, 1):
        src = _rand(rng, color, W, H)
        g, o = b2.Mipmap(color, 5, W, H), oracle.Mipmap(color, 5, W, H)
        g.upload(0, src); o.upload(0, src)
        g.generateMipmaps(0); o.generateMipmaps(0)
        n = 400
        x = rng.uniform(-3, W + 3, n).astype(np.float32)
        y = rng.uniform(-3, H + 3, n).astype(np.float32)
        x[:8] = [0, 0.5, W, W - 0.5, W / 2, 1e-3, W + 2, -1]
        y[:8] = [0, 0.5, H, H - 0.5, H / 2, 1e-3, H + 2, -1]
        for level in (0, 1, 2, 4, 5, 9, -1):
            assert np.array_equal(g.linearSample(level, x, y), o.linearSample(level, x, y)), (color, level)
        for level in (0, 3, 4, 5):
            xs, ys = np.clip(x, 0, W), np.clip(y, 0, H)
            assert np.array_equal(g.cubicSample(level, xs, ys), o.cubicSample(level, xs, ys)), (color, level)
        lv = rng.uniform(0, 6.5, n).astype(np.float32)
        lv[:4] = [0.0, 1.0, 2.5, 5.0]
        assert np.array_equal(g.linearMipmapSample(lv, x, y), o.linearMipmapSample(lv, x, y)), color


def test_errors_are_reported(b2):
    g = b2.Mipmap(0, 2, 16, 16)
    with pytest.raises(b2.B2Error):
        g.generateNextLevel(0, (0, 0, 16, 16), 7)
    with pytest.raises(b2.B2Error):
        g.generateLevel(1, 0, (0, 0, 100, 100))
    z = b2.Mipmap(1, 2, 16, 16)
    with pytest.raises(b2.B2Error):
        z.generateNextLevel(2, (0, 0, 16, 16), 1)   # boxAlphaCovIntoPremul is RGBA-only (mipmap.d:594-595)
    e = b2.Mipmap(0, 4, 0, 0)   # zero-sized is supported (image.d:657-669)
    assert e.numLevels() == 0


def test_cubic_tma_kernel_on_every_rectangle():
    """The TMA-fed cubic kernel normally takes only rectangles that fill the GPU; B2_CUBIC_TMA=2 (read once per process)
    forces it on every cubic RGBA step, so the small-size, odd-size and dirty-rectangle cases above run through its edge
    handling (clamped outer columns, ragged strips, partial stores) in a child process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, B2_CUBIC_TMA="2")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", "not tma_kernel_on_every",
                        "tests/test_gpu_mip.py::test_generate_next_level_full", "tests/test_gpu_mip.py::test_generate_next_level_dirty_rects",
                        "tests/test_gpu_mip.py::test_cubic_wide_rectangles", "tests/test_gpu_mip.py::test_full_chain_diffuse_and_depth_recipes"],
                       cwd=root, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-1000:]
