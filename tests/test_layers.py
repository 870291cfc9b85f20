"""widget cache -> level-0 maps on the device (blit / opacity-masked blend), bit-exact
against the CPU restatement, then through the whole frame."""
import numpy as np
import pytest
from dplug_b200 import synth


def test_oracle_blend_semantics(oracle):
    """blendColor known properties (color.d:453-526): alpha 255 -> foreground, alpha 0 -> untouched, RGBA rounds down;
    the L16 path reproduces the reference's 32-bit wrap (65535 blended over anything at full opacity gives 65534)."""
    z = lambda v, dt, shape=(1, 1): np.full(shape, v, dtype=dt)
    dst = [np.full((1, 1, 4), 10, np.uint8), z(1000, np.uint16), np.full((1, 1, 4), 20, np.uint8)]
    src = [np.full((1, 1, 4), 200, np.uint8), z(65535, np.uint16), np.full((1, 1, 4), 100, np.uint8)]
    op = [z(255, np.uint8), z(255, np.uint8), z(0, np.uint8)]
    oracle.draw_layer(dst, src, op, 0, 0, [(0, 0, 1, 1)], 1)
    assert (dst[0] == 200).all() and (dst[2] == 20).all()
    assert dst[1][0, 0] == 65534          # int overflow in the reference: (65535*65535) wraps, /65535 truncates to -2
    dst[0][...] = 10
    op[0][...] = 128
    oracle.draw_layer(dst, src, op, 0, 0, [(0, 0, 1, 1)], 1)
    assert dst[0][0, 0, 0] == (200 * 128 + 10 * 127) // 255


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1])
def test_draw_layer_parity(b2, oracle, mode):
    W, H = 320, 200
    rng = np.random.default_rng(4 + mode)
    d, m, z = synth.frame(W, H)
    c = b2.PBRCompositor(0); c.resizeBuffers(W, H); c.setSkybox(synth.skybox(64, 48))
    c.diffuseMap.upload(0, d); c.materialMap.upload(0, m); c.depthMap.upload(0, z)
    dst = [d.copy(), z.copy(), m.copy()]
    for trial in range(4):
        w, h = int(rng.integers(20, 150)), int(rng.integers(20, 120))
        x, y = int(rng.integers(0, W - w + 1)), int(rng.integers(0, H - h + 1))
        src = [rng.integers(0, 256, (h, w, 4), dtype=np.uint8), rng.integers(0, 65536, (h, w), dtype=np.uint16),
               rng.integers(0, 256, (h, w, 4), dtype=np.uint8)]
        if trial == 0:
            src[1][...] = 65535                                   # exercises the wrapping L16 products
        op = [rng.integers(0, 256, (h, w), dtype=np.uint8) for _ in range(3)]
        for o in op:
            o[rng.random((h, w)) < 0.3] = 0
            o[rng.random((h, w)) < 0.3] = 255
        layer = b2.Layer(w, h, withOpacity=(mode == 1))
        for k in range(3):
            layer.upload(k, src[k])
            if mode == 1:
                layer.upload(3 + k, op[k])
        rects = [(0, 0, w // 2, h), (w // 2, h // 3, w, h)]
        c.drawLayer(layer, x, y, rects, mode)
        oracle.draw_layer(dst, src, op if mode == 1 else None, x, y, rects, mode)
        assert np.array_equal(c.diffuseMap.download(0), dst[0])
        assert np.array_equal(c.depthMap.download(0), dst[1])
        assert np.array_equal(c.materialMap.download(0), dst[2])
    # the frame composited from maps produced on the device equals the frame from the oracle-produced maps
    c.regenerateMipmaps([(0, 0, W, H)]); c.compositeGUI([(0, 0, W, H)])
    ref = oracle.Graphics(W, H, numThreads=8); ref.setSkybox(synth.skybox(64, 48)); ref.setInputs(dst[0], dst[2], dst[1])
    diff = np.abs(c.download().astype(np.int16) - ref.fullFrame().astype(np.int16))
    assert diff.max() <= 1
