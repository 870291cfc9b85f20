"""Second, independent restatement (numpy float32 scalars and arrays, written from the D source) of the reference's in-tree
stb_image_resize port as ImageResizer drives it (graphics/dplug/graphics/stb_image_resize.d, resizer.d; the
`DPLUG_USE_STB_IMAGE_RESIZE_V2 = false` branches).  tests/test_np_resizer.py requires it to agree byte for byte with the
C++ restatement oracle/ref_resizer.cpp.  TEST INFRASTRUCTURE: nothing in dplug_b200/ imports this.

Structure differs from both other statements on purpose: coefficients are computed per contributor exactly as the port
does (stb_image_resize.d:967-1190), then every output sample is accumulated source by source over whole image rows /
columns at once."""
import math

import numpy as np

f32 = np.float32
CUBICBSPLINE, CATMULLROM, MITCHELL, MKS_2013_86, MKS_2021 = 3, 4, 5, 11, 13     # stb_image_resize.d:270-286
KIND_FILTER = {0: MKS_2013_86, 1: 0, 2: MKS_2021, 3: 0, 4: CUBICBSPLINE}       # resizer.d:346-372, 39-64, 171-197, 144-169, 92-117


def kernel(flt, x):
    """stb_image_resize.d:659-774; x is np.float32; D promotes float to double where a double literal appears"""
    x = f32(abs(x))
    if flt == CUBICBSPLINE:
        if x < 1: return f32((4 + x * x * (3 * x - 6)) / 6)
        if x < 2: return f32((8 + x * (-12 + x * (6 - x))) / 6)
        return f32(0)
    if flt == CATMULLROM:
        if x < 1: return f32(1 - x * x * (f32(2.5) - f32(1.5) * x))
        if x < 2: return f32(2 - x * (4 + x * (f32(0.5) * x - f32(2.5))))
        return f32(0)
    if flt == MITCHELL:
        if x < 1: return f32((16 + x * x * (21 * x - 36)) / 18)
        if x < 2: return f32((32 + x * (-60 + x * (36 - 7 * x))) / 18)
        return f32(0)
    if flt == MKS_2013_86:
        return f32(f32(0.14) * _mk2013(x) + f32(0.86) * _mks2013(x))
    x2 = f32(x * x)                                                             # MKS 2021: float throughout
    if x < 0.5: return f32(f32(f32(577.0) / f32(576.0)) - f32(f32(239.0) / f32(144.0)) * x2)
    if x < 1.5: return f32((140 * x2 - 379 * x + 239) / f32(144.0))
    if x < 2.5: return f32(-(24 * x2 - 113 * x + 130) / f32(144.0))
    if x < 3.5: return f32((4 * x2 - 27 * x + 45) / f32(144.0))
    if x < 4.5: return f32(-(4 * x2 - 36 * x + 81) / f32(1152.0))
    return f32(0)


def _mk2013(x):                                                                 # :711-721 (double arithmetic)
    xd = float(x)
    if xd < 0.5: return f32(0.75 - float(f32(x * x)))
    if xd < 1.5: return f32(0.5 * (xd - 1.5) * (xd - 1.5))
    return f32(0)


def _mks2013(x):                                                                # :730-751 (double arithmetic)
    xd = float(x)
    if x <= f32(1.17549435e-38): return f32(f32(17.0) / f32(16.0))
    if xd < 0.5: return f32(17.0 / 16.0 - 7.0 * xd * xd / 4.0)
    if xd < 1.5:
        x2 = float(f32(x * x))
        return f32(0.25 * (4 * x2 - 11.0 * xd + 7.0))
    if xd < 2.5: return f32(-0.125 * (xd - 5.0 / 2.0) * (xd - 5.0 / 2.0))
    return f32(0)


def _support(flt):                                                              # :806-822
    return f32({CUBICBSPLINE: 2, CATMULLROM: 2, MITCHELL: 2, MKS_2013_86: 3, MKS_2021: 5}[flt])


def axis_lists(flt, n_in, n_out):
    """per output sample: list of (source index before clamping, np.float32 weight) in the order the port accumulates"""
    scale = f32(f32(n_out) / f32(n_in))                                         # :2270-2292, whole image: shift 0
    up = scale > 1
    if flt == 0:
        flt = CATMULLROM if up else MITCHELL                                   # :2294-2302, :364-366
    S = _support(flt)
    width = int(math.ceil(f32(S * 2)))                                          # :860-866
    if up:
        radius = f32(S * scale)
        out = []
        for n in range(n_out):
            centre = f32(f32(n) + f32(0.5))
            lo, hi = f32(f32(centre - radius) / scale), f32(f32(centre + radius) / scale)
            in_centre = f32(centre / scale)
            first, last = math.floor(float(lo) + 0.5), math.floor(float(hi) - 0.5)
            g, total, i = {}, f32(0), 0
            while i <= last - first:                                            # :996-1044
                g[i] = kernel(flt, f32(in_centre - f32(f32(i + first) + f32(0.5))))
                if i == 0 and g[i] == 0:
                    first += 1
                    continue
                total = f32(total + g[i]); i += 1
            norm = f32(f32(1) / total)
            ws = [f32(g[i] * norm) for i in range(last - first + 1)]
            n1 = last
            for i in range(last - first, -1, -1):
                if ws[i] != 0: break
                n1 = first + i - 1
            out.append([(k, ws[k - first]) for k in range(first, n1 + 1)])
        return out
    margin = int(math.ceil(f32(f32(S * 2) / scale))) // 2                       # :842-858
    total_c = n_in + 2 * margin
    radius = f32(S / scale)
    n0, n1, g = [0] * total_c, [0] * total_c, [None] * total_c
    for n in range(total_c):                                                    # :982-994, :1046-1074
        centre = f32(f32(n - margin) + f32(0.5))
        lo, hi = f32(f32(centre - radius) * scale), f32(f32(centre + radius) * scale)
        out_centre = f32(centre * scale)
        first, last = math.floor(float(lo) + 0.5), math.floor(float(hi) - 0.5)
        ws = [f32(kernel(flt, f32(f32(f32(i + first) + f32(0.5)) - out_centre)) * scale) for i in range(last - first + 1)]
        n0[n], n1[n] = first, last
        for i in range(last - first, -1, -1):
            if ws[i] != 0: break
            n1[n] = first + i - 1
        g[n] = ws + [f32(0)] * (width + 2)
    assert all(n1[j] - n0[j] + 1 <= width for j in range(total_c)), "a group wider than the coefficient width spills into its neighbour in the port: not modelled"
    for o in range(n_out):                                                      # :1076-1149: every output's weights sum to 1
        total = f32(0)
        for j in range(total_c):
            if n0[j] <= o <= n1[j]: total = f32(total + g[j][o - n0[j]])
            elif o < n0[j]: break
        norm = f32(f32(1) / total)
        for j in range(total_c):
            if n0[j] <= o <= n1[j]: g[j][o - n0[j]] = f32(g[j][o - n0[j]] * norm)
            elif o < n0[j]: break
    lists = [[] for _ in range(n_out)]
    for j in range(total_c):
        skip = 0
        while skip < width and g[j][skip] == 0: skip += 1
        a = n0[j] + skip
        while a < 0:
            a += 1; skip += 1
        b = min(n1[j], n_out - 1)
        for k in range(a, b + 1):
            i = (k - a) + skip
            lists[k].append((j - margin, g[j][i] if i < width else f32(0)))
    return lists


def _apply(x, lists, n_in, axis):
    x = np.moveaxis(x, axis, 0)
    out = np.zeros((len(lists),) + x.shape[1:], dtype=np.float32)
    for o, lst in enumerate(lists):
        acc = np.zeros(x.shape[1:], dtype=np.float32)
        for s, w in lst:
            acc = acc + x[min(max(s, 0), n_in - 1)] * w                         # float32 array ops: one rounding each
        out[o] = acc
    return np.moveaxis(out, 0, axis)


def resize_image(kind, src, out_w, out_h):
    ih, iw = src.shape[:2]
    if (iw, ih) == (out_w, out_h):
        return src.copy()                                                       # sameSizeResize, resizer.d:453-465
    mx = f32(65535.0 if src.dtype == np.uint16 else 255.0)
    dec = src.astype(np.float32) / mx                                           # :1236-1316
    tmp = _apply(dec, axis_lists(KIND_FILTER[kind], iw, out_w), iw, 1)
    acc = _apply(tmp, axis_lists(KIND_FILTER[kind], ih, out_h), ih, 0)
    enc = np.clip(acc, f32(0), f32(1)) * mx + f32(0.5)                          # :484-493, :1770-1785
    return enc.astype(np.int64).astype(src.dtype)
