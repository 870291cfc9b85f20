"""Second, independent restatement (numpy, written from the D sources, whole-image array operations) of the widened rows of
SURVEY §8(f): the post-composite chain, the widget blend, the screenshot encoders.  tests/test_np_widening.py requires it
to agree byte for byte with the C++ restatement in oracle/ — the same role tools/np_reference.py plays for the lighting
passes.  TEST INFRASTRUCTURE: nothing in dplug_b200/ imports this."""
import numpy as np


# ---- graphics/dplug/graphics/color.d:453-482 (version(LDC) branch) ----------------------------------------------------
def blend_color_rgba(fg, bg, alpha):
    """fg, bg: (..., 4) uint8; alpha: (...) uint8.  madd of (fg, bg) with (alpha, ~alpha), * 32897, >> 23, saturating packs"""
    a = alpha.astype(np.int64)[..., None]
    inv = (~alpha).astype(np.uint8).astype(np.int64)[..., None]
    product = fg.astype(np.int64) * a + bg.astype(np.int64) * inv          # _mm_madd_epi16: fits 32 bits
    product = (product * 32897) & 0xFFFFFFFF                                # 32-bit multiply
    product >>= 23                                                          # logical shift
    return np.clip(product, 0, 255).astype(np.uint8)                       # packs_epi32 + packus_epi16


# ---- color.d:522-526: products and their sum are D `int` (32-bit, wrapping), the division is signed ----------------------
def blend_color_l16(fg, bg, alpha16):
    inv = (~alpha16.astype(np.int64)) & 0xFFFF                              # cast(ushort)(~cast(int)alpha)
    s = fg.astype(np.int64) * alpha16.astype(np.int64) + bg.astype(np.int64) * inv
    s = ((s + 2 ** 31) % 2 ** 32) - 2 ** 31                                 # wrap to int32
    q = np.where(s >= 0, s // 65535, -((-s) // 65535))                      # truncating division
    return (q & 0xFFFF).astype(np.uint16)                                   # cast(ushort)


# ---- graphics/dplug/graphics/draw.d:933-969 -----------------------------------------------------------------------------
def blend_with_alpha(src, dst, alpha):
    """in place on dst; src / dst (h, w, 4) uint8 or (h, w) uint16, alpha (h, w) uint8; alpha == 0 leaves dst alone"""
    m = alpha != 0
    if dst.dtype == np.uint16:
        a16 = (alpha.astype(np.uint16) << 8) | alpha.astype(np.uint16)
        dst[m] = blend_color_l16(src, dst, a16)[m]
    else:
        dst[m] = blend_color_rgba(src, dst, alpha)[m]
    return dst


# ---- gui/dplug/gui/graphics.d:825-868 (stages E, G, H; F = raw widgets, none here) + colorcorrection.d:157-170 -----------
def present(composited, rendered, pf, window, user_origin, rects, table=None, redraw_borders=False):
    """pf: 0 RGBA8, 1 BGRA8, 2 ARGB8 (B2_PF_*).  `rects` = _rectsToDisplayDisjointed (user coordinates).  In place."""
    ux, uy = user_origin
    H, W = composited.shape[:2]
    WH, WW = window.shape[:2]
    for (x0, y0, x1, y1) in rects:
        px = composited[y0:y1, x0:x1].copy()                                # E. blit composited -> rendered
        if table is not None:                                               # UIColorCorrection draws in the Raw pass (stage F)
            t = np.asarray(table, dtype=np.uint8).reshape(3, 256)
            px[..., 0] = t[0][px[..., 0]]; px[..., 1] = t[1][px[..., 1]]; px[..., 2] = t[2][px[..., 2]]
        r, g, b = px[..., 0].copy(), px[..., 1].copy(), px[..., 2].copy()   # G. reorderComponents, alpha forced to 255
        a = np.full_like(r, 255)
        order = {0: (r, g, b, a), 1: (b, g, r, a), 2: (a, r, g, b)}[pf]
        rendered[y0:y1, x0:x1] = np.stack(order, axis=-1)
    user = (ux, uy, min(ux + W, WW), min(uy + H, WH))                       # _userArea clipped to the window
    if redraw_borders:                                                      # H. resizeContent
        window[...] = np.array([255, 0, 0, 0] if pf == 2 else [0, 0, 0, 255], dtype=np.uint8)
        copies = [user]
    else:
        copies = []
        for (x0, y0, x1, y1) in rects:                                      # translated by _userArea.min, clipped to _userArea
            c = (max(x0 + ux, user[0]), max(y0 + uy, user[1]), min(x1 + ux, user[2]), min(y1 + uy, user[3]))
            if c[2] > c[0] and c[3] > c[1]:
                copies.append(c)
    for (x0, y0, x1, y1) in copies:
        window[y0:y1, x0:x1] = rendered[y0 - uy:y1 - uy, x0 - ux:x1 - ux]
    return window


# ---- gui/dplug/gui/screencap.d:21-84 -------------------------------------------------------------------------------------
def encode_screenshot_qb(rendered, depth):
    H, W = depth.shape
    head = bytes([1, 1, 0, 0]) + np.array([0, 0, 0, 0, 1], dtype="<u4").tobytes() + bytes([1, ord("0")]) \
        + np.array([W, 16 + 10, H], dtype="<i4").tobytes() + np.array([0, 0, 0], dtype="<i4").tobytes()
    d = 10 + (16 * depth.astype(np.int64)) // 65536                         # (h, w)
    vox = np.empty((H, 26, W, 4), dtype=np.uint8)                           # z (image row), y, x
    vox[..., :3] = rendered[:, None, :, :3]
    vox[..., 3] = np.where(d[:, None, :] >= np.arange(26)[None, :, None], 255, 0)
    return head + vox.tobytes()


# ---- screencap.d:88-131, the pixel pass ----------------------------------------------------------------------------------
def screenshot_png_pixels(rendered, pf):
    out = rendered.copy()
    if pf == 1:
        out[..., 0], out[..., 2] = rendered[..., 2], rendered[..., 0]
    elif pf == 2:
        out[..., 0], out[..., 1], out[..., 2], out[..., 3] = rendered[..., 1], rendered[..., 2], rendered[..., 3], rendered[..., 0]
    return out
