"""the screenshot encoders (gui/dplug/gui/screencap.d:21-136) — encodeScreenshotAsQB built on the
device, the pixel pass of encodeScreenshotAsPNG, and the PNG container — against their CPU restatement, taken on the
rendered buffer of doDraw stage G.bis (gui/dplug/gui/graphics.d:843-857)."""
import struct
import zlib

import numpy as np
import pytest

from dplug_b200 import synth


def test_oracle_qb_known_answer(oracle):
    """the file layout of screencap.d:33-82 on a 3 x 2 image written out by hand"""
    rend = np.arange(3 * 2 * 4, dtype=np.uint8).reshape(2, 3, 4)
    depth = np.array([[0, 4095, 4096], [32768, 65535, 61439]], dtype=np.uint16)
    qb = oracle.encodeScreenshotAsQB(rend, 0, depth)
    head = bytes([1, 1, 0, 0]) + struct.pack("<5I", 0, 0, 0, 0, 1) + bytes([1, ord("0")]) + struct.pack("<3I", 3, 26, 2) + struct.pack("<3i", 0, 0, 0)
    assert len(head) == 50 and qb[:50] == head
    vox = np.frombuffer(qb[50:], dtype=np.uint8).reshape(2, 26, 3, 4)     # z (image row), y (depth), x, rgba
    assert len(qb) == 50 + 3 * 26 * 2 * 4
    for z in range(2):
        for x in range(3):
            d = 10 + (16 * int(depth[z, x])) // 65536
            assert (vox[z, :, x, :3] == rend[z, x, :3]).all()
            assert (vox[z, : d + 1, x, 3] == 255).all() and (vox[z, d + 1:, x, 3] == 0).all()
    assert [10 + (16 * int(v)) // 65536 for v in depth.ravel()] == [10, 10, 11, 18, 25, 24]


def test_oracle_png_pixel_pass(oracle):
    """screencap.d:106-131: BGRA swaps bytes 0 and 2, ARGB rotates a to the end"""
    px = np.array([[[1, 2, 3, 4], [10, 20, 30, 40]]], dtype=np.uint8)
    assert np.array_equal(oracle.screenshotPixels(px, 0), px)
    assert oracle.screenshotPixels(px, 1).tolist() == [[[3, 2, 1, 4], [30, 20, 10, 40]]]
    assert oracle.screenshotPixels(px, 2).tolist() == [[[2, 3, 4, 1], [20, 30, 40, 10]]]


def _decode_png(png):
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, ihdr = 8, b"", None
    while pos < len(png):
        n, = struct.unpack(">I", png[pos:pos + 4])
        typ, data = png[pos + 4:pos + 8], png[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", png[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + data) == crc, typ
        if typ == b"IHDR":
            ihdr = struct.unpack(">IIBBBBB", data)
        elif typ == b"IDAT":
            idat += data
        pos += 12 + n
        if typ == b"IEND":
            break
    assert pos == len(png)
    w, h, depth, ctype, comp, filt, inter = ihdr
    assert (depth, ctype, comp, filt, inter) == (8, 6, 0, 0, 0)
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, w * 4 + 1)
    assert (raw[:, 0] == 0).all()                                          # filter type 0 on every scanline
    return raw[:, 1:].reshape(h, w, 4)


@pytest.mark.gpu
@pytest.mark.parametrize("pf", [0, 1, 2])
@pytest.mark.parametrize("w,h", [(300, 180), (517, 263)])
def test_screenshot_parity(b2, oracle, pf, w, h):
    d, m, z = synth.frame(w, h)
    z = z.copy()
    z[::7, ::5] = np.random.default_rng(pf).integers(0, 65536, size=z[::7, ::5].shape, dtype=np.uint16)   # every depth bucket, 4096 boundaries
    z[0, :16] = np.arange(16, dtype=np.uint16) * 4096
    z[1, :16] = np.arange(16, dtype=np.uint32).astype(np.uint32) * 4096 + 4095
    c = b2.PBRCompositor(0); c.resizeBuffers(w, h); c.setSkybox(synth.skybox(64, 48))
    c.diffuseMap.upload(0, d); c.materialMap.upload(0, m); c.depthMap.upload(0, z)
    c.regenerateMipmaps([(0, 0, w, h)]); c.compositeGUI([(0, 0, w, h)])
    comp = c.download()
    for table in (None, b2.liftGammaGainContrastTable((0.03, 1.1, 0.97, 0.2), (0.0, 0.95, 1.0, 0.1), (0.06, 1.0, 1.05, 0.0))):
        c.setColorCorrection(table)
        # the reference's rendered buffer at stage G.bis: blit + colour correction + reorderComponents(pf)
        rend = np.zeros_like(comp); win = np.zeros_like(comp)
        oracle.present(comp, rend, pf, win, (0, 0), [(0, 0, w, h)], table, redrawBorders=True)
        qb_g, qb_o = c.encodeScreenshotAsQB(pf), oracle.encodeScreenshotAsQB(rend, pf, z)
        assert len(qb_g) == len(qb_o) == 50 + w * 26 * h * 4
        assert qb_g == qb_o
        want = oracle.screenshotPixels(rend, pf)
        assert np.array_equal(c.screenshotPixels(pf), want)
        assert np.array_equal(_decode_png(c.encodeScreenshotAsPNG(pf)), want)
