"""SURVEY §8(f) row 2, second half: ImageResizer (graphics/dplug/graphics/resizer.d) on the device against the CPU
restatement of the reference's in-tree stb_image_resize port (graphics/dplug/graphics/stb_image_resize.d, the
`DPLUG_USE_STB_IMAGE_RESIZE_V2 = false` branches).  Bit-exact.  The restatement itself is held against hand-computed
filter weights, against the port's own invariants (stb_image_resize.d:1027-1028, 1103-1104: weights sum to 1) and
against a float64 model of separable filtering written from the filter definitions."""
import numpy as np
import pytest

from dplug_b200 import synth

KINDS = {0: "diffuse (MKS 2013, 86 %)", 1: "material (Catmull-Rom / Mitchell)", 2: "depth (MKS 2021)", 3: "generic L16", 4: "smoother (cubic B-spline)"}
SIZE_PAIRS = [((124, 66), (62, 33)), ((124, 66), (93, 50)), ((62, 33), (124, 66)), ((100, 41), (137, 23)), ((37, 91), (35, 200)),
              ((64, 64), (64, 64)), ((200, 10), (13, 7)), ((9, 5), (40, 31)), ((1240, 660), (620, 330))]


def _src(rng, kind, w, h):
    if kind in (2, 3):
        z = synth.depth(w, h).copy()
        z[::5, ::3] = rng.integers(0, 65536, size=z[::5, ::3].shape, dtype=np.uint16)
        return z
    img = synth.diffuse(w, h).copy()
    img[::4, ::6] = rng.integers(0, 256, size=img[::4, ::6].shape, dtype=np.uint8)
    return img


# ---- the float64 model: weights straight from the filter definitions (stb_image_resize.d:659-774), normalised per output
def _kernel(kind_filter, x):
    x = np.abs(x)
    if kind_filter == "catmullrom":
        return np.where(x < 1, 1 - x * x * (2.5 - 1.5 * x), np.where(x < 2, 2 - x * (4 + x * (0.5 * x - 2.5)), 0.0))
    if kind_filter == "mitchell":
        return np.where(x < 1, (16 + x * x * (21 * x - 36)) / 18, np.where(x < 2, (32 + x * (-60 + x * (36 - 7 * x))) / 18, 0.0))
    if kind_filter == "cubic":
        return np.where(x < 1, (4 + x * x * (3 * x - 6)) / 6, np.where(x < 2, (8 + x * (-12 + x * (6 - x))) / 6, 0.0))
    mk = np.where(x < 0.5, 0.75 - x * x, np.where(x < 1.5, 0.5 * (x - 1.5) ** 2, 0.0))
    mks13 = np.where(x < 0.5, 17 / 16 - 7 * x * x / 4, np.where(x < 1.5, 0.25 * (4 * x * x - 11 * x + 7), np.where(x < 2.5, -0.125 * (x - 2.5) ** 2, 0.0)))
    if kind_filter == "mks2013_86":
        return 0.14 * mk + 0.86 * mks13
    x2 = x * x
    return np.where(x < 0.5, 577 / 576 - 239 / 144 * x2, np.where(x < 1.5, (140 * x2 - 379 * x + 239) / 144, np.where(
        x < 2.5, -(24 * x2 - 113 * x + 130) / 144, np.where(x < 3.5, (4 * x2 - 27 * x + 45) / 144, np.where(x < 4.5, -(4 * x2 - 36 * x + 81) / 1152, 0.0)))))


def _axis_matrix(kind, n_in, n_out):
    scale = n_out / n_in
    up = scale > 1
    name = {0: "mks2013_86", 1: "catmullrom" if up else "mitchell", 2: "mks2021", 3: "catmullrom" if up else "mitchell", 4: "cubic"}[kind]
    reach = 12 if up else int(np.ceil(12 / scale))
    src = np.arange(-reach, n_in + reach)
    o = np.arange(n_out)
    if up:      # weights in input space around the output centre (stb_image_resize.d:996-1044)
        w = _kernel(name, (o[:, None] + 0.5) / scale - (src[None, :] + 0.5))
    else:       # the kernel stretched to the output grid (stb_image_resize.d:1046-1074)
        w = _kernel(name, (o[:, None] + 0.5) - (src[None, :] + 0.5) * scale)
    w = w / w.sum(axis=1, keepdims=True)
    m = np.zeros((n_out, n_in))
    np.add.at(m, (np.repeat(o, len(src)), np.tile(np.clip(src, 0, n_in - 1), n_out)), w.ravel())   # clamp edges
    return m


def _model(kind, src, ow, oh):
    f = src.astype(np.float64) / (65535.0 if src.dtype == np.uint16 else 255.0)
    mh, mv = _axis_matrix(kind, src.shape[1], ow), _axis_matrix(kind, src.shape[0], oh)
    f = np.tensordot(mv, np.tensordot(f, mh.T, axes=([1], [0])) if f.ndim == 2 else np.einsum("hwc,ow->hoc", f, mh), axes=([1], [0]))
    return np.clip(f, 0, 1) * (65535.0 if src.dtype == np.uint16 else 255.0)


def test_oracle_catmull_rom_weights_known_answer(oracle):
    """2x up-sampling of an impulse at input pixel 8 (centre 8.5): output pixel o sits at (o + 0.5) / 2, so outputs 15..18
    are at distances 0.75, 0.25, 0.25, 0.75 and outputs 13, 14, 19, 20 at 1.75 / 1.25, where Catmull-Rom is negative.
    Weights by hand from stb_image_resize.d:671-681: w(0.25) = 0.8671875, w(0.75) = 0.2265625, w(1.25) = -0.0703125,
    w(1.75) = -0.0234375 (they sum to 1, so the normalisation leaves them alone)."""
    src = np.zeros((4, 16, 4), dtype=np.uint8)
    src[:, 8] = 200
    out = oracle.resizeImage(1, src, 32, 8)[4, :, 0].astype(int)          # columns are constant: the vertical pass keeps them
    want = {15: int(200 * 0.2265625 + 0.5), 16: int(200 * 0.8671875 + 0.5), 17: int(200 * 0.8671875 + 0.5), 18: int(200 * 0.2265625 + 0.5)}
    assert want == {15: 45, 16: 173, 17: 173, 18: 45}
    for x in range(32):
        assert out[x] == want.get(x, 0), (x, out[10:24])                  # the negative lobes saturate at 0
    # the same impulse under a bright constant shows the negative lobes: 255 - 200 * w on a background of 255
    src2 = np.full((4, 16, 4), 255, dtype=np.uint8)
    src2[:, 8] = 55
    out2 = oracle.resizeImage(1, src2, 32, 8)[4, :, 0].astype(int)
    assert out2[14] == 255 and out2[19] == 255                            # 255 + 200 * 0.0703 saturates at 255
    assert out2[16] == int(255 - 200 * 0.8671875 + 0.5) and out2[15] == int(255 - 200 * 0.2265625 + 0.5)


@pytest.mark.parametrize("kind", sorted(KINDS))
def test_oracle_preserves_constants_and_copies_equal_sizes(oracle, kind):
    l16 = kind in (2, 3)
    for value in ((0, 1, 40000, 65535) if l16 else (0, 1, 200, 255)):
        src = np.full((37, 53), value, np.uint16) if l16 else np.full((37, 53, 4), value, np.uint8)
        for ow, oh in [(20, 15), (100, 80), (30, 90), (120, 20)]:
            assert (oracle.resizeImage(kind, src, ow, oh) == value).all(), (kind, value, ow, oh)   # weights sum to 1
    rng = np.random.default_rng(kind)
    src = _src(rng, kind, 40, 30)
    assert np.array_equal(oracle.resizeImage(kind, src, 40, 30), src)                              # sameSizeResize, resizer.d:453-465


@pytest.mark.parametrize("kind", sorted(KINDS))
def test_oracle_close_to_float64_model(oracle, kind):
    rng = np.random.default_rng(10 + kind)
    for (iw, ih), (ow, oh) in SIZE_PAIRS[:8]:
        if (iw, ih) == (ow, oh):
            continue
        src = _src(rng, kind, iw, ih)
        got = oracle.resizeImage(kind, src, ow, oh).astype(np.float64)
        want = _model(kind, src, ow, oh)
        tol = 0.51 + (65535 if src.dtype == np.uint16 else 255) * 2e-5    # rounding to integer + float32 accumulation
        assert np.abs(got - want).max() <= tol, (KINDS[kind], (iw, ih), (ow, oh), float(np.abs(got - want).max()))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", sorted(KINDS))
def test_resizer_parity(b2, oracle, kind):
    rng = np.random.default_rng(20 + kind)
    r = b2.ImageResizer()
    fn = [r.resizeImageDiffuse, r.resizeImageMaterial, r.resizeImageDepth, r.resizeImageGeneric, r.resizeImageSmoother][kind]
    for (iw, ih), (ow, oh) in SIZE_PAIRS:
        src = _src(rng, kind, iw, ih)
        out = np.zeros((oh, ow) + src.shape[2:], dtype=src.dtype)
        fn(src, out)
        want = oracle.resizeImage(kind, src, ow, oh)
        assert np.array_equal(out, want), (KINDS[kind], (iw, ih), (ow, oh), int((out != want).sum()))


@pytest.mark.gpu
def test_background_cache_resized_on_the_device(b2, oracle):
    """PBRBackgroundGUI (pbrbackgroundgui.d:122-199): the three full-size images are resized into the widget's cache and
    the cache is blitted into the maps' dirty rectangles — here without the cache ever leaving the device."""
    W, H = 310, 165
    rng = np.random.default_rng(7)
    d, m, z = _src(rng, 0, 1240, 660), synth.material(620, 330), _src(rng, 2, 1240, 660)
    layer = b2.Layer(W, H, withOpacity=False)
    layer.resizePlane(b2.Layer.DIFFUSE, 0, d)
    layer.resizePlane(b2.Layer.MATERIAL, 1, m)
    layer.resizePlane(b2.Layer.DEPTH, 2, z)
    c = b2.PBRCompositor(0); c.resizeBuffers(W, H)
    rects = [(0, 0, W, H)]
    c.drawLayer(layer, 0, 0, rects, mode=b2.Layer.BLIT)
    assert np.array_equal(c.diffuseMap.download(0), oracle.resizeImage(0, d, W, H))
    assert np.array_equal(c.materialMap.download(0), oracle.resizeImage(1, m, W, H))
    assert np.array_equal(c.depthMap.download(0), oracle.resizeImage(2, z, W, H))


def _axis_table(kind, n_in, n_out):
    import ctypes as C
    from dplug_b200 import _lib
    l = _lib.lib()
    start = (C.c_int * (n_out + 1))()
    n = l.b2_resize_axis_table(kind, n_in, n_out, start, None, None, 0)
    assert n > 0
    src, coef = (C.c_int * n)(), (C.c_float * n)()
    assert l.b2_resize_axis_table(kind, n_in, n_out, start, src, coef, n) == n
    return np.frombuffer(start, dtype=np.int32).copy(), np.frombuffer(src, dtype=np.int32).copy(), np.frombuffer(coef, dtype=np.float32).copy()


def _gather(x, table, n_in, axis):
    """out[o] = ((0 + x[clamp(src0)] * c0) + x[clamp(src1)] * c1) + ... in float32, one IEEE multiply and add per step:
    what k_resize_h / k_resize_v do"""
    start, src, coef = table
    x = np.moveaxis(x, axis, 0)
    out = np.zeros((len(start) - 1,) + x.shape[1:], dtype=np.float32)
    for o in range(len(start) - 1):
        acc = np.zeros(x.shape[1:], dtype=np.float32)
        for e in range(start[o], start[o + 1]):
            acc = (acc + (x[min(max(int(src[e]), 0), n_in - 1)] * np.float32(coef[e])).astype(np.float32)).astype(np.float32)
        out[o] = acc
    return np.moveaxis(out, 0, axis)


@pytest.mark.parametrize("kind", sorted(KINDS))
def test_host_tables_reproduce_the_port_bit_for_bit(b2, oracle, kind):
    """No GPU needed: the gather tables the product's host code derives (b2_resize_axis_table: what the two kernels are
    fed) applied in numpy float32 with the kernels' operation order give exactly the bytes of the restated port, which
    walks scanlines through a ring buffer and scatters when down-sampling."""
    rng = np.random.default_rng(30 + kind)
    for (iw, ih), (ow, oh) in [((62, 33), (31, 17)), ((31, 17), (62, 33)), ((50, 21), (69, 12)), ((19, 46), (18, 100)), ((124, 66), (93, 50)), ((40, 9), (7, 5)), ((40, 30), (40, 15)), ((25, 30), (60, 30))]:
        src = _src(rng, kind, iw, ih)
        maxv = np.float32(65535.0 if src.dtype == np.uint16 else 255.0)
        dec = (src.astype(np.float32) / maxv).astype(np.float32)                     # decode: IEEE division
        tmp = _gather(dec, _axis_table(kind, iw, ow), iw, axis=1)     # an axis that keeps its size is filtered too (scale 1)
        acc = _gather(tmp, _axis_table(kind, ih, oh), ih, axis=0)
        enc = (np.clip(acc, np.float32(0), np.float32(1)) * maxv + np.float32(0.5)).astype(np.float32)
        got = enc.astype(np.int64).astype(src.dtype)                                    # truncation
        want = oracle.resizeImage(kind, src, ow, oh)
        assert np.array_equal(got, want), (KINDS[kind], (iw, ih), (ow, oh), int((got != want).sum()))
