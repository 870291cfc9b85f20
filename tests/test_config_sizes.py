"""Oracle parity at the BASELINE config sizes (GPU): C3 8192x8192 8-level chain, C4 2048x1024 full frame, and bands of
the C5 16384x16384 canvas — the sizes the benchmark runs at, not only the small frames of the other tests."""
import numpy as np
import pytest
from dplug_b200 import synth

pytestmark = pytest.mark.gpu


def test_c3_8192_pyramid_vs_oracle(b2, oracle):
    """BASELINE configs[2]: Mipmap(8) on 8192x8192 RGBA, level 1 boxAlphaCovIntoPremul, levels 2..8 cubic
    (graphics.d:1378-1392): every level bit-exact against the oracle."""
    W = H = 8192
    blk = synth.diffuse(W, 1024, synth.SEED + 3)
    rng = np.random.default_rng(11)
    blk2 = rng.integers(0, 256, size=(1024, W, 4), dtype=np.uint8)       # dense random alpha: every texel takes the arithmetic path
    m = b2.Mipmap(b2.RGBA8, 8, W, H)
    o = oracle.Mipmap(oracle.RGBA8, 8, W, H)
    lvl0 = o.level(0)
    for k, y0 in enumerate(range(0, H, 1024)):
        src = blk if k % 2 == 0 else np.roll(blk2, 13 * k, axis=1)
        m.uploadRows(0, src, y0)
        lvl0[y0:y0 + 1024] = src
    m.generateChain(2, 2)
    rect = (0, 0, W, H)
    for level in range(1, 9):
        rect = o.generateNextLevel(oracle.Q_PREMUL if level == 1 else oracle.Q_CUBIC, rect, level)
    assert m.numLevels() == 9
    for level in range(1, 9):
        assert np.array_equal(m.download(level), o.download(level)), level


def test_c4_2048x1024_full_frame_vs_oracle(b2, oracle):
    """BASELINE configs[3]: one plug-in-instance UI at 2048x1024, the whole frame against the oracle (+-1 LSB lighting,
    bit-exact mips), through the host call the D shim makes."""
    W, H = 2048, 1024
    d, m, z = synth.frame(W, H, synth.SEED + 21)
    sky = synth.skybox(256, 192)
    c = b2.PBRCompositor(0); c.resizeBuffers(W, H); c.setSkybox(sky)
    wfb = np.zeros((H, W, 4), np.uint8)
    c.compositeTile(wfb, b2.tileAreas([(0, 0, W, H)], 64, 64), d, m, z)
    ref = oracle.Graphics(W, H, numThreads=16); ref.setSkybox(sky); ref.setInputs(d, m, z)
    want = ref.fullFrame()
    for l in range(1, 6):
        assert np.array_equal(c.diffuseMap.download(l), ref.diffuseMap.download(l)), l
    for l in range(1, 5):
        assert np.array_equal(c.depthMap.download(l), ref.depthMap.download(l)), l
    diff = np.abs(wfb.astype(np.int16) - want.astype(np.int16))
    assert diff.max() <= 1 and (diff > 0).mean() < 0.02, (diff.max(), (diff > 0).mean())


def test_c5_16384_canvas_bands_vs_oracle(b2, oracle):
    """BASELINE configs[4]: three 64-row bands of the 16384x16384 canvas (top edge, across the 2048-row boundary of the
    8-GPU split, bottom edge) against the oracle run on the same canvas size (global coordinates: toEye, blue noise and
    the edge clamps depend on them).  Only the rows the bands depend on are filled and regenerated on either side."""
    W = H = 16384
    sky = synth.skybox(256, 192)
    c = b2.PBRCompositor(0); c.resizeBuffers(W, H); c.setSkybox(sky)
    ref = oracle.Graphics(W, H, numThreads=16); ref.setSkybox(sky)
    from dplug_b200 import sharding as sh
    for (y0, y1) in ((0, 64), (2016, 2080), (H - 64, H)):
        lo, hi = sh.need_rows(W, H, y0, y1, sh.DIFFUSE, 0)        # the widest footprint (bloom), SURVEY Appendix C
        lo, hi = max(0, lo - 2), min(H, hi + 2)
        d, m, z = (synth.diffuse(W, H, synth.SEED + 5, lo, hi), synth.material(W, H, synth.SEED + 5, lo, hi), synth.depth(W, H, synth.SEED + 5, lo, hi))
        c.diffuseMap.uploadRows(0, d, lo); c.materialMap.uploadRows(0, m, lo); c.depthMap.uploadRows(0, z, lo)
        ref.setInputRows(d, m, z, lo)
        rect = [(0, lo, W, hi)]
        c.regenerateMipmaps(rect); ref.regenerateMipmaps(rect)
        band = [(0, y0, W, y1)]
        c.compositeGUI(band); ref.compositeGUI(band)
        got = c.downloadRows(y0, y1)
        want = np.array(ref.output())[y0:y1]
        diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
        assert diff.max() <= 1 and (diff > 0).mean() < 0.02, ((y0, y1), diff.max(), (diff > 0).mean())
