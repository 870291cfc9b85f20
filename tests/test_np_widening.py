"""The C++ restatement of the widened SURVEY §8(f) rows (oracle/ref_present.cpp: post-composite chain, widget blend,
screenshot encoders) against a second restatement written independently from the D sources in numpy
(tools/np_widening.py).  Byte for byte; no GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import np_widening as npw  # noqa: E402


def test_blend_color_every_operand_pair(oracle):
    """blendColor(RGBA) over all 256 x 256 (colour, alpha) pairs against a background sweep; blendColor(L16) over a grid
    that includes the products that wrap a 32-bit int (color.d:522-526)."""
    fgv, av = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for bgval in (0, 1, 77, 128, 254, 255):
        h, w = fgv.shape
        src = (np.repeat(fgv[..., None], 4, axis=-1).copy(), np.zeros((h, w), np.uint16), np.repeat(fgv[..., None], 4, axis=-1).copy())
        dst = [np.full((h, w, 4), bgval, np.uint8), np.zeros((h, w), np.uint16), np.full((h, w, 4), 255 - bgval, np.uint8)]
        op = (av.copy(), np.zeros((h, w), np.uint8), av.copy())
        want0 = npw.blend_with_alpha(src[0], dst[0].copy(), op[0]); want2 = npw.blend_with_alpha(src[2], dst[2].copy(), op[2])
        oracle.draw_layer(dst, src, op, 0, 0, [(0, 0, w, h)], 1)
        assert np.array_equal(dst[0], want0) and np.array_equal(dst[2], want2), bgval
    rng = np.random.default_rng(0)
    h, w = 256, 512
    fg = rng.integers(0, 65536, size=(h, w), dtype=np.uint16); bg = rng.integers(0, 65536, size=(h, w), dtype=np.uint16)
    fg[:8] = 65535; bg[:4] = 65535                                           # bright texels, high opacity: the wrapping products
    a = np.tile(np.arange(256, dtype=np.uint8), (w // 256) * h).reshape(h, w)
    want = npw.blend_with_alpha(fg, bg.copy(), a)
    dst = [np.zeros((h, w, 4), np.uint8), bg.copy(), np.zeros((h, w, 4), np.uint8)]
    src = (np.zeros((h, w, 4), np.uint8), fg, np.zeros((h, w, 4), np.uint8))
    oracle.draw_layer(dst, src, (np.zeros((h, w), np.uint8), a, np.zeros((h, w), np.uint8)), 0, 0, [(0, 0, w, h)], 1)
    assert np.array_equal(dst[1], want)


@pytest.mark.parametrize("pf", [0, 1, 2])
def test_present_chain(oracle, pf):
    rng = np.random.default_rng(pf)
    H, W = 90, 140
    comp = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
    table = rng.integers(0, 256, size=(3, 256), dtype=np.uint8)
    for (WW, WH, ux, uy) in [(W, H, 0, 0), (W + 30, H + 20, 15, 10), (100, 60, 0, 0)]:
        for tab in (None, table):
            rend_o, rend_n = np.zeros_like(comp), np.zeros_like(comp)
            win_o = rng.integers(0, 256, size=(WH, WW, 4), dtype=np.uint8); win_n = win_o.copy()
            full = [(0, 0, W, H)]
            oracle.present(comp, rend_o, pf, win_o, (ux, uy), full, tab, redrawBorders=True)
            npw.present(comp, rend_n, pf, win_n, (ux, uy), full, tab, redraw_borders=True)
            assert np.array_equal(win_o, win_n) and np.array_equal(rend_o, rend_n)
            comp2 = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
            rects = [(10, 20, 90, 70), (100, 64, W, H), (0, 0, 5, 5)]
            oracle.present(comp2, rend_o, pf, win_o, (ux, uy), rects, tab)
            npw.present(comp2, rend_n, pf, win_n, (ux, uy), rects, tab)
            assert np.array_equal(win_o, win_n) and np.array_equal(rend_o, rend_n)


@pytest.mark.parametrize("pf", [0, 1, 2])
def test_screenshot_encoders(oracle, pf):
    rng = np.random.default_rng(10 + pf)
    H, W = 37, 53
    rend = rng.integers(0, 256, size=(H, W, 4), dtype=np.uint8)
    depth = rng.integers(0, 65536, size=(H, W), dtype=np.uint16)
    depth[0, :16] = np.arange(16) * 4096; depth[1, :16] = np.arange(16) * 4096 + 4095
    assert oracle.encodeScreenshotAsQB(rend, pf, depth) == npw.encode_screenshot_qb(rend, depth)
    assert np.array_equal(oracle.screenshotPixels(rend, pf), npw.screenshot_png_pixels(rend, pf))
