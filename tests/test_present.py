"""the fused post-composite chain (doDraw stages E-H) against its CPU restatement."""
import numpy as np
import pytest
from dplug_b200 import synth


def test_lggc_table_properties(b2):
    """setLiftGammaGainContrastRGB: neutral settings give the identity table (colorcorrection.d:34-41, 69-116)."""
    ident = b2.liftGammaGainContrastTable()
    assert np.array_equal(ident, np.tile(np.arange(256, dtype=np.uint8), (3, 1)))
    t = b2.liftGammaGainContrastTable((0.05, 1.2, 0.95, 0.3), (0.0, 0.9, 1.0, 0.0), (0.1, 1.0, 1.1, 0.5))
    assert t.shape == (3, 256) and (np.diff(t.astype(int), axis=1) >= 0).all()      # monotonic transfer curves
    assert t[2, -1] == 255 and t[0, 0] > 0                                            # gain > 1 saturates, lift raises black


def test_oracle_present_identity(oracle):
    """RGBA8, no table, user area == window: the window receives the composited rects with alpha forced to 255."""
    rng = np.random.default_rng(0)
    comp = rng.integers(0, 256, size=(40, 60, 4), dtype=np.uint8)
    rend = np.zeros_like(comp); win = np.zeros_like(comp)
    oracle.present(comp, rend, 0, win, (0, 0), [(0, 0, 60, 40)])
    assert np.array_equal(win[..., :3], comp[..., :3]) and (win[..., 3] == 255).all()


@pytest.mark.gpu
@pytest.mark.parametrize("pf", [0, 1, 2])
def test_present_parity(b2, oracle, pf):
    W, H = 300, 180
    d, m, z = synth.frame(W, H)
    c = b2.PBRCompositor(0); c.resizeBuffers(W, H); c.setSkybox(synth.skybox(64, 48))
    c.diffuseMap.upload(0, d); c.materialMap.upload(0, m); c.depthMap.upload(0, z)
    c.regenerateMipmaps([(0, 0, W, H)]); c.compositeGUI([(0, 0, W, H)])
    comp = c.download()
    table = b2.liftGammaGainContrastTable((0.03, 1.1, 0.97, 0.2), (0.0, 0.95, 1.0, 0.1), (0.06, 1.0, 1.05, 0.0))
    c.setColorCorrection(table)
    # window larger than the UI: user area centred (graphics.d:1063-1084), first draw repaints the borders
    WW, WH = 340, 200
    ux, uy = (WW - W) // 2, (WH - H) // 2
    win_g = np.zeros((WH, WW, 4), dtype=np.uint8); win_o = np.zeros_like(win_g); rend = np.zeros_like(comp)
    c.present(win_g, pf, (ux, uy), [(0, 0, W, H)], redrawBorders=True)
    oracle.present(comp, rend, pf, win_o, (ux, uy), [(0, 0, W, H)], table, redrawBorders=True)
    assert np.array_equal(win_g, win_o)
    # later frames: only some rects are displayed; everything else in the window must stay as it was
    rects = [(10, 20, 90, 70), (128, 64, 300, 180), (0, 0, 5, 5)]
    comp2 = comp.copy()
    d2 = synth.diffuse(W, H, synth.SEED + 1)
    c.diffuseMap.upload(0, d2); c.regenerateMipmaps([(0, 0, W, H)]); c.compositeGUI([(0, 0, W, H)])
    comp2 = c.download()
    c.present(win_g, pf, (ux, uy), rects)
    oracle.present(comp2, rend, pf, win_o, (ux, uy), rects, table)
    assert np.array_equal(win_g, win_o)
    # no colour correction, window smaller than the UI (cropped user area, graphics.d:1068-1082)
    c.setColorCorrection(None)
    small_g = np.zeros((100, 200, 4), dtype=np.uint8); small_o = np.zeros_like(small_g)
    c.present(small_g, pf, (0, 0), [(0, 0, W, H)], redrawBorders=True)
    oracle.present(comp2, np.zeros_like(comp2), pf, small_o, (0, 0), [(0, 0, W, H)], None, redrawBorders=True)
    assert np.array_equal(small_g, small_o)
