"""Second, independent restatement of the hot path in numpy float32 — TEST INFRASTRUCTURE (oracle pinning), never imported
by the product.

Written from the D sources (gui/dplug/gui/legacypbr.d, gui/dplug/gui/ransac.d, graphics/dplug/graphics/mipmap.d), not
from oracle/: whole-image array operations instead of per-pixel loops, gather-based samplers instead of pointer walks,
so a transcription mistake in one restatement is unlikely to be repeated in the other.  tests/test_np_reference.py
requires oracle/ (C++) and this file to agree BIT FOR BIT on small frames (edges included); tools/make_golden.py
writes tests/golden/ from THIS file.

Every arithmetic step is a single IEEE binary32 operation on float32 arrays, in the order the D expressions evaluate
(left to right, no contraction: LDC does not fuse without -ffast-math), e.g. `a * b * c` is `(a * b) * c`.
Assumptions shared with oracle/ because no reference-held vector exists: `_mm_pow_ps` is
sse_mathfun's exp_ps(y * log_ps(x)); `_mm_rsqrt_ss` is the Intel RSQRTSS table (oracle/rsqrtss_table.inc);
compile-time `^^` and `exp` constants are evaluated in double and narrowed once.
"""
import os
import numpy as np

f32 = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def F(x):
    return np.float32(x)


# ------------------------------------------------------------------------------------------------ RSQRTSS / pow
def _load_rsqrt_table():
    txt = open(os.path.join(ROOT, "oracle", "rsqrtss_table.inc")).read()
    import re
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)       # the header comment mentions a hex constant
    vals = [int(t, 16) for t in re.findall(r"0x[0-9a-fA-F]+", txt)]
    assert len(vals) == 2048
    return np.array(vals, dtype=np.uint32)


_RSQRT = None


def rsqrt_ss(x):
    """_mm_rsqrt_ss on positive normal floats (ransac.d:43): 2048-bin table on (exponent parity, top 10 mantissa bits),
    exact power-of-two scaling.  The table entry is the result for the bin with exponent 126 / 127."""
    global _RSQRT
    if _RSQRT is None:
        _RSQRT = _load_rsqrt_table()
    b = np.asarray(x, dtype=f32).view(np.uint32)
    e = (b >> np.uint32(23)).astype(np.int64)
    de = (e - 126) >> 1                                 # floor((e - 126) / 2)
    entry = _RSQRT[((b >> np.uint32(13)) & np.uint32(0x7ff)).astype(np.int64)].astype(np.int64)
    out = (entry - (de << 23)).astype(np.uint32).view(f32)
    return np.where(b == 0, f32(np.inf), out)


def _log_ps(x):
    """sse_mathfun log_ps (the algorithm behind intel-intrinsics' _mm_log_ps), one IEEE operation per step."""
    x = np.asarray(x, dtype=f32)
    invalid = x <= 0
    x = np.maximum(x, np.array(0x00800000, dtype=np.uint32).view(f32))
    xb = x.view(np.uint32)
    emm0 = (xb >> np.uint32(23)).astype(np.int32) - 0x7f
    x = ((xb & np.uint32(0x807fffff)) | np.uint32(0x3f000000)).view(f32)
    e = emm0.astype(f32) + F(1)
    mask = x < F(0.707106781186547524)
    tmp = np.where(mask, x, F(0))
    x = x - F(1)
    e = e - np.where(mask, F(1), F(0))
    x = x + tmp
    z = x * x
    y = np.full_like(x, F(7.0376836292E-2))
    for c in (-1.1514610310E-1, 1.1676998740E-1, -1.2420140846E-1, 1.4249322787E-1, -1.6668057665E-1, 2.0000714765E-1,
              -2.4999993993E-1, 3.3333331174E-1):
        y = y * x
        y = y + F(c)
    y = y * x
    y = y * z
    tmp = e * F(-2.12194440e-4)
    y = y + tmp
    tmp = z * F(0.5)
    y = y - tmp
    tmp = e * F(0.693359375)
    x = x + y
    x = x + tmp
    return np.where(invalid, f32(np.nan), x)


def _exp_ps(x):
    x = np.asarray(x, dtype=f32)
    x = np.minimum(x, F(88.3762626647949))
    x = np.maximum(x, F(-88.3762626647949))
    fx = x * F(1.44269504088896341)
    fx = fx + F(0.5)
    tmp = np.trunc(fx).astype(f32)
    tmp = np.where(tmp > fx, tmp - F(1), tmp)
    fx = tmp
    tmp = fx * F(0.693359375)
    z = fx * F(-2.12194440e-4)
    x = x - tmp
    x = x - z
    z = x * x
    y = np.full_like(x, F(1.9875691500E-4))
    for c in (1.3981999507E-3, 8.3334519073E-3, 4.1665795894E-2, 1.6666665459E-1, 5.0000001201E-1):
        y = y * x
        y = y + F(c)
    y = y * z
    y = y + x
    y = y + F(1)
    n = fx.astype(np.int32) + 0x7f
    pow2n = (n.astype(np.uint32) << np.uint32(23)).view(f32)
    return y * pow2n


def pow_ps(base, exponent):
    return _exp_ps(np.asarray(exponent, dtype=f32) * _log_ps(base))


def fastlog2(v):                                        # legacypbr.d:1117-1133
    x = np.asarray(v, dtype=f32).view(np.int32)
    log_2 = ((x >> 23) & 255) - 128
    x = (x & ~(255 << 23)) + (127 << 23)
    return x.astype(np.int32).view(f32) + log_2.astype(f32)


# ------------------------------------------------------------------------------------------------ mipmaps
def num_levels(max_level, w, h):                         # mipmap.d:83-94
    n = 0
    while n <= max_level and w != 0 and h != 0:
        n += 1
        w >>= 1
        h >>= 1
    return n


def box_rgba(src):                                       # generateLevelBoxRGBA, mipmap.d:676-762 == (a+b+c+d+2)>>2 (mipmap.d:563-584)
    h, w = src.shape[0] >> 1, src.shape[1] >> 1
    s = src[:2 * h, :2 * w].astype(np.uint32)
    return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint8)


def box_l16(src):                                        # generateLevelBoxL16, mipmap.d:764-794
    h, w = src.shape[0] >> 1, src.shape[1] >> 1
    s = src[:2 * h, :2 * w].astype(np.uint32)
    return ((s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2).astype(np.uint16)


def box_premul(src):                                     # generateLevelBoxAlphaCovIntoPremulRGBA, mipmap.d:823-900 (non-legacy)
    h, w = src.shape[0] >> 1, src.shape[1] >> 1
    inv = F(1.0) / F(255)

    def lin(q):                                          # convert_gammaspace_to_linear_premul
        c = q.astype(f32)
        a = c[..., 3] * inv
        out = np.empty_like(c)
        for k in range(3):
            out[..., k] = (((c[..., k] * inv) * c[..., k]) * inv) * a
        out[..., 3] = a
        return out

    s = src[:2 * h, :2 * w]
    A, B, Cq, D = lin(s[0::2, 0::2]), lin(s[0::2, 1::2]), lin(s[1::2, 0::2]), lin(s[1::2, 1::2])
    mean = ((A + B) + Cq) + D
    return (((mean * F(0.25)) * F(255.0)) + F(0.5)).astype(np.int32).astype(np.uint8)   # cast(ubyte): truncation of a value in [0.5, 255.5]


def _cubic(src, out_dtype):                              # generateLevelCubicRGBA / L16, mipmap.d:902-1109
    sh, sw = src.shape[0], src.shape[1]
    h, w = sh >> 1, sw >> 1
    ys = np.arange(h)
    xs = np.arange(w)
    rows = [np.maximum(2 * ys - 1, 0), 2 * ys, 2 * ys + 1, np.minimum(2 * ys + 2, sh - 1)]
    cols = [np.maximum(2 * xs - 1, 0), 2 * xs, 2 * xs + 1, np.minimum(2 * xs + 2, sw - 1)]
    wt = [1, 3, 3, 1]
    s = src.astype(np.int64)
    acc = 0
    for ry, wy in zip(rows, wt):
        for cx, wx in zip(cols, wt):
            acc = acc + wy * wx * s[ry][:, cx]
    return ((acc + 32) >> 6).astype(out_dtype)


def cubic_rgba(src):
    return _cubic(src, np.uint8)


def cubic_l16(src):
    return _cubic(src, np.uint16)


def diffuse_pyramid(level0):                             # Mipmap(5) + regenerateMipmaps recipe, graphics.d:1118,1378-1392
    levels = [level0]
    for l in range(1, num_levels(5, level0.shape[1], level0.shape[0])):
        levels.append(box_premul(levels[-1]) if l == 1 else cubic_rgba(levels[-1]))
    return levels


def depth_pyramid(level0):                               # Mipmap(4), graphics.d:1119,1412-1419
    levels = [level0]
    for l in range(1, num_levels(4, level0.shape[1], level0.shape[0])):
        levels.append(cubic_l16(levels[-1]) if l >= 3 else box_l16(levels[-1]))
    return levels


def skybox_pyramid(level0):                              # Mipmap(12, image) + generateMipmaps(box), legacypbr.d:868-869, mipmap.d:517-528
    levels = [level0]
    for l in range(1, num_levels(12, level0.shape[1], level0.shape[0])):
        levels.append(cubic_rgba(levels[-1]) if l >= 3 else box_rgba(levels[-1]))
    return levels


# ------------------------------------------------------------------------------------------------ samplers
def _level_scale(level):
    return F(2.0 ** -level)


def linear_sample(levels, level, x, y):                  # Mipmap.linearSample, mipmap.d:293-496
    level = min(max(level, 0), len(levels) - 1)
    img = levels[level]
    x = np.asarray(x, dtype=f32) * _level_scale(level) - F(0.5)
    y = np.asarray(y, dtype=f32) * _level_scale(level) - F(0.5)
    x = np.where(x < 0, F(0), x)
    y = np.where(y < 0, F(0), y)
    ix = x.astype(np.int32)                               # cvttps: truncation
    iy = y.astype(np.int32)
    fx = x - ix.astype(f32)
    fy = y - iy.astype(f32)
    maxx, maxy = img.shape[1] - 1, img.shape[0] - 1
    ix = np.minimum(ix, maxx)
    iy = np.minimum(iy, maxy)
    ix1 = np.minimum(ix + 1, maxx)
    iy1 = np.minimum(iy + 1, maxy)
    fxm1 = F(1) - fx
    fym1 = F(1) - fy
    A, B, Cq, D = (img[iy, ix].astype(f32), img[iy, ix1].astype(f32), img[iy1, ix].astype(f32), img[iy1, ix1].astype(f32))
    if img.ndim == 3:
        fx, fy, fxm1, fym1 = fx[..., None], fy[..., None], fxm1[..., None], fym1[..., None]
    up = A * fxm1 + B * fx
    down = Cq * fxm1 + D * fx
    return up * fym1 + down * fy


def _cubic_interp(t, x0, x1, x2, x3):                     # mipmap.d:222-229 / 263-270, D evaluation order
    a = (F(-0.5) * x0) + (F(0.5) * x2)
    b = ((x0 - (F(2.5) * x1)) + (F(2.0) * x2)) - (F(0.5) * x3)
    c = (((F(-0.5) * x0) + (F(1.5) * x1)) - (F(1.5) * x2)) + F(0.5) * x3
    return ((x1 + t * a) + (t * t) * b) + ((t * t) * t) * c


def cubic_sample(levels, level, x, y):                   # Mipmap.cubicSample, mipmap.d:167-287
    level = min(max(level, 0), len(levels) - 1)
    img = levels[level]
    x = np.asarray(x, dtype=f32) * _level_scale(level) - F(0.5)
    y = np.asarray(y, dtype=f32) * _level_scale(level) - F(0.5)
    xi = [np.clip((x + F(k)).astype(np.int32), 0, img.shape[1] - 1) for k in (-1, 0, 1, 2)]
    yi = [np.clip((y + F(k)).astype(np.int32), 0, img.shape[0] - 1) for k in (-1, 0, 1, 2)]
    a = x + F(1)
    b = y + F(1)
    a = a - a.astype(np.int32).astype(f32)
    b = b - b.astype(np.int32).astype(f32)
    if img.ndim == 3:
        a, b = a[..., None], b[..., None]
    R = [_cubic_interp(a, *[img[yy, xx].astype(f32) for xx in xi]) for yy in yi]
    r = _cubic_interp(b, *R)
    r = np.where(r < 0, F(0), r)
    return np.where(r > 65535, F(65535), r)


def linear_mipmap_sample(levels, level, x, y):           # mipmap.d:149-160 (per-pixel level)
    level = np.asarray(level, dtype=f32)
    il = level.astype(np.int32)
    fl = level - il.astype(f32)
    out = np.zeros(level.shape + (4,), dtype=f32)
    for l in np.unique(il):
        sel = il == l
        a = linear_sample(levels, int(l), x[sel], y[sel])
        b = linear_sample(levels, int(l) + 1, x[sel], y[sel])
        f = fl[sel][..., None]
        out[sel] = np.where(f == 0, a, a * (F(1) - f) + b * f)
    return out


# ------------------------------------------------------------------------------------------------ the eight passes
class Params:
    def __init__(self):
        def normalized(v):                                # vec3f.normalized, math/dplug/math/vector.d:344-376
            v = np.array(v, dtype=f32)
            s = F(0)
            for k in range(3):
                s = s + v[k] * v[k]
            inv = F(1) / np.sqrt(s, dtype=f32)
            return v * inv
        self.light1_color = np.array([0.25, 0.25, 0.25], dtype=f32) * F(1.3)     # legacypbr.d:453
        self.light2_dir = normalized([0.0, 1.0, 0.1])                             # :626
        self.light2_color = np.array([0.481] * 3, dtype=f32)                      # :629
        self.light3_dir = normalized([0.0, 1.0, 0.1])                             # :676
        self.light3_color = np.array([0.26] * 3, dtype=f32)                       # :679
        self.ambient = F(0.08125)                                                 # :391
        self.skybox_amount = F(0.52)                                              # :843
        self.tonemap_threshold = F(1.0)                                           # :1045
        self.tonemap_ratio = F(0.3)                                               # :1049


_BLUE = None


def _blue_noise():
    """BLUE_NOISE_16x16 (legacypbr.d:1012-1030), parsed from the product's copy of the constant table"""
    global _BLUE
    if _BLUE is None:
        txt = open(os.path.join(ROOT, "dplug_b200", "csrc", "pbr_math.cuh")).read()
        body = txt[txt.index("kBlueNoise16x16[256] = {"):]
        body = body[body.index("{") + 1:body.index("};")]
        import re
        body = re.sub(r"//[^\n]*", "", body)
        _BLUE = np.array([int(t) for t in re.findall(r"\d+", body)], dtype=f32)
        assert _BLUE.size == 256
    return _BLUE


def composite(diffuse0, material0, depth0, skybox0=None, p=None):
    """Full frame through PBRCompositor's eight passes (legacypbr.d:201-260) -> (H, W, 4) uint8."""
    p = p or Params()
    H, W = depth0.shape
    D = diffuse_pyramid(diffuse0)
    Z = depth_pyramid(depth0)
    jj, ii = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    div255 = F(1) / F(255.0)

    def dep(dx, dy):                                      # level-0 depth with its 1-px replicated border (graphics.d:1121-1126)
        return depth0[np.clip(jj + dy, 0, H - 1), np.clip(ii + dx, 0, W - 1)]

    # ---- PassComputeNormal, legacypbr.d:279-381 + computePlaneFittingNormal, ransac.d:53-121
    mult = F(1.0 / 4655.0)
    n = [dep(dx, dy).astype(f32) * mult for dy in (-1, 0, 1) for dx in (-1, 0, 1)]
    c, e = F(0.03660694846), F(0.118115527008)
    d03, d58 = n[0:4], n[5:9]
    w03, w58 = [c, e, c, e], [e, c, e, c]
    mean = [d03[k] * w03[k] + d58[k] * w58[k] for k in range(4)]
    filt = (((n[4] * F(0.38111009811) + mean[0]) + mean[1]) + mean[2]) + mean[3]
    d03 = [v - filt for v in d03]
    d58 = [v - filt for v in d58]
    xzw03, xzw58 = [-c, F(0), c, -e], [e, -c, F(0), c]
    yzw03, yzw58 = [-c, -e, -c, F(0)], [F(0), c, e, c]
    mxz = [d03[k] * xzw03[k] + d58[k] * xzw58[k] for k in range(4)]
    myz = [d03[k] * yzw03[k] + d58[k] * yzw58[k] for k in range(4)]
    xz = ((mxz[0] + mxz[1]) + mxz[2]) + mxz[3]
    yz = ((myz[0] + myz[1]) + myz[2]) + myz[3]
    nv = [-xz, yz, np.full_like(xz, F(0.382658847856)), np.zeros_like(xz)]

    def fast_normalize(v):                                # _mm_fast_normalize_ps, ransac.d:39-46
        sq = [a * a for a in v]
        l2 = ((sq[0] + sq[1]) + sq[2]) + sq[3]
        inv = rsqrt_ss(l2)
        return [a * inv for a in v]

    normal = fast_normalize(nv)
    # variance (old method), legacypbr.d:321-377: integer samples -> float, 12 samples of which the 4th lane is multiplied by 0
    M1 = [dep(dx, -1).astype(f32) for dx in (-1, 0, 1)]
    P0 = [dep(dx, 0).astype(f32) for dx in (-1, 0, 1)]
    P1 = [dep(dx, 1).astype(f32) for dx in (-1, 0, 1)]
    fdx = [F(1), F(-2), F(1)]
    mdx = [fdx[k] * ((M1[k] + P0[k]) + P1[k]) for k in range(3)]
    depthDX = (mdx[0] + mdx[1]) + mdx[2]
    mdy = [F(1) * (M1[k] + P1[k]) + P0[k] * F(-2) for k in range(3)]
    depthDY = (mdy[0] + mdy[1]) + mdy[2]
    depthDX = depthDX * (F(1) / F(256.0))
    depthDY = depthDY * (F(1) / F(256.0))
    variance = depthDX * depthDX + depthDY * depthDY

    accum = np.zeros((H, W, 4), dtype=f32)
    base4 = diffuse0.astype(f32) * div255                 # convertBaseColorToFloat4
    mat4 = material0.astype(f32) * div255
    base3 = base4[..., :3]
    px, py = ii.astype(f32) + F(0.5), jj.astype(f32) + F(0.5)

    # ---- PassAmbientOcclusion, legacypbr.d:400-444
    avg = (((linear_sample(Z, 1, px, py) + linear_sample(Z, 2, px, py)) + linear_sample(Z, 3, px, py)) + linear_sample(Z, 4, px, py)) * F(0.25)
    diff = depth0.astype(f32) - avg
    cavity = (diff + F(23040.0)) * (F(1.0) / F(23040))
    cavity = np.where(cavity >= 1, F(1), np.where(cavity < 0, F(0), cavity))
    accum = accum + base4 * (cavity * p.ambient)[..., None]

    # ---- PassObliqueShadowLight, legacypbr.d:460-616
    fall = 0.78
    weights = [F(1.0)] + [F(float(np.float32(fall)) ** k) for k in range(1, 11)]
    fo = float(np.float32(fall))
    total = F((1.0 - fo ** 11) / (1.0 - fo) - 1)
    inv_total = F(1 / (float(F(1.7)) * float(total)))
    light = np.zeros((H, W), dtype=f32)
    zc = depth0.astype(np.int32)
    for s in range(1, 11):
        x1 = np.minimum(ii + s, W - 1)
        x2 = np.maximum(ii - s, 0)
        yy = np.maximum(jj - s, 0)
        z = zc + s
        dif1 = z - depth0[yy, x1].astype(np.int32)
        dif2 = z - depth0[yy, x2].astype(np.int32)
        if s <= 8:                                        # the two groups of four SSE lanes
            A = F(0.00006510416)
            c1 = np.maximum(F(0), np.minimum(F(1), F(1) + dif1.astype(f32) * A))
            c2 = np.maximum(F(0), np.minimum(F(1), F(1) + dif2.astype(f32) * A))
        else:                                             # scalar tail, samples 9 and 10
            dv = F(1.0) / F(15360)
            c1 = np.where(dif1 >= 0, F(1), np.where(dif1 < -15360, F(0), (dif1 + 15360).astype(f32) * dv))
            c2 = np.where(dif2 >= 0, F(1), np.where(dif2 < -15360, F(0), (dif2 + 15360).astype(f32) * dv))
        light = light + (c1 + c2 * F(0.7)) * weights[s]
    k = light * inv_total
    accum[..., :3] = accum[..., :3] + (base3 * p.light1_color) * k[..., None]

    # ---- PassDirectionalLight, legacypbr.d:636-666
    dot = ((F(0) + normal[0] * p.light2_dir[0]) + normal[1] * p.light2_dir[1]) + normal[2] * p.light2_dir[2]
    dfac = F(0.5) + F(0.5) * dot
    rough = material0[..., 0].astype(f32) * div255
    a = F(0.24) - rough * F(0.5)
    dfac = F(0) + ((F(1.0) - F(0)) * (dfac - a)) / (F(1) - a)          # linmap, core/dplug/core/math.d:25-28
    accum[..., :3] = accum[..., :3] + (base3 * p.light2_color) * dfac[..., None]

    # ---- PassSpecularLight, legacypbr.d:720-813
    invW, invH = F(1.0) / F(W), F(1.0) / F(H)
    to_eye = fast_normalize([F(0.5) - ii.astype(f32) * invW, jj.astype(f32) * invH - F(0.5), np.ones((H, W), dtype=f32), np.zeros((H, W), dtype=f32)])
    l3 = [-p.light3_dir[0], -p.light3_dir[1], -p.light3_dir[2], F(0)]
    half = fast_normalize([to_eye[k] - l3[k] for k in range(4)])
    m = [half[k] * normal[k] for k in range(4)]
    spec = ((m[0] + m[1]) + m[2]) + m[3]
    spec = np.where(spec < F(1e-3), F(1e-3), spec)
    rb = np.arange(256)
    exp_table = (F(0.8) * np.exp(((F(1) - rb.astype(f32) / F(255.0)) * F(5.5)).astype(np.float64)).astype(f32)) * F(2.8)
    exponent = exp_table[material0[..., 0]]
    Ft = F(1.0) / (F(1.0) + (exponent * variance) * F(4e-5))
    toks = (F(1.0) + exponent * Ft) / (F(1.0) + exponent)
    spec = pow_ps(spec, exponent * Ft)
    roughF = (F(10) * (F(1.0) - mat4[..., 0])) * (F(1) - mat4[..., 1] * F(0.5))
    spec = (spec * roughF) * toks
    accum = accum + (base4 * np.array(list(p.light3_color) + [0], dtype=f32)) * (spec * mat4[..., 2])[..., None]

    # ---- PassSkyboxReflections, legacypbr.d:872-930
    if skybox0 is not None:
        S = skybox_pyramid(skybox0)
        sw, sh = skybox0.shape[1], skybox0.shape[0]
        npx = F(sw * sh)
        lvl = variance * npx
        lvl = F(0.5) * fastlog2(F(1.0) + lvl * F(0.00001)) + F(6.0 / 255.0) * material0[..., 0].astype(f32)
        dd = [to_eye[k] * normal[k] for k in range(4)]     # _mm_reflectnormal_ps(toEye, normal), ransac.d:31-36
        ssum = F(2) * (((dd[0] + dd[1]) + dd[2]) + dd[3])
        refl = [to_eye[k] - ssum * normal[k] for k in range(2)]
        skyx = F(0.5) + ((F(0.5) - refl[0] * F(0.5)) * (F(sw) - F(1.0)))
        skyy = F(0.5) + ((F(0.5) + refl[1] * F(0.5)) * (F(sh) - F(1.0)))
        sky = linear_mipmap_sample(S, lvl, skyx, skyy)
        accum = accum + (base4 * sky) * (mat4[..., 1] * (p.skybox_amount * div255))[..., None]

    # ---- PassEmissiveContribution, legacypbr.d:948-1007 (non-legacy)
    noise = (_blue_noise()[(ii & 15) * 16 + (jj & 15)] - F(127.5)) * F(0.003)
    AMT = F(0.002) * F(0.67)
    em = linear_sample(D, 1, px, py) * AMT
    em = em + linear_sample(D, 2, px, py) * AMT
    em = em + linear_sample(D, 3, px, py) * AMT
    em = em + cubic_sample(D, 4, px, py) * AMT
    em = em + (cubic_sample(D, 5, px, py) * AMT) * (F(1) + noise)[..., None]
    accum = accum + em

    # ---- PassClampAndConvertTo8bit, legacypbr.d:1057-1107 (non-legacy)
    color = accum.copy()
    color[..., 3] = F(1.0)
    exceed = np.maximum(F(0), color - p.tonemap_threshold)
    luma = (F(0.212655) * exceed[..., 0] + F(0.715158) * exceed[..., 1]) + F(0.072187) * exceed[..., 2]
    color = color + (luma * (p.tonemap_ratio / F(3)))[..., None]
    color[..., 3] = F(1.0)
    v = color * F(255.99)
    iv = np.where(np.isfinite(v) & (np.abs(v) < 2147483648.0), v, -2147483648.0).astype(np.int64)   # cvttps: out of range -> 0x80000000
    iv = np.clip(np.clip(iv, -32768, 32767), 0, 255)       # packs_epi32 then packus_epi16
    return iv.astype(np.uint8)
