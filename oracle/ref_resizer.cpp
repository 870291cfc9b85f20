// TEST INFRASTRUCTURE (see b2ref.h): CPU restatement of ImageResizer (graphics/dplug/graphics/resizer.d:39-200,346-372)
// over the IN-TREE stb_image_resize port (graphics/dplug/graphics/stb_image_resize.d:160-2623), i.e. the reference built
// with `DPLUG_USE_STB_IMAGE_RESIZE_V2 = false` (stb_image_resize.d:157).  The default build (`= true`) calls the
// un-vendored dub package stb_image_resize2-d instead, whose source is not under /root/reference: its filters are the
// same kernels, its float operation order is not — parity with THAT build is unpinned.
//
// The restatement keeps the port's structure on purpose — contributors and coefficient groups per axis, a decode
// buffer with margins, the ring buffer, scatter accumulation when down-sampling — so that it checks the gather
// formulation of the CUDA path independently.  Every function cites the lines it follows.  Only the paths the resizer
// reaches are restated: 1 or 4 channels, uint8 / uint16, linear colour space, no alpha weighting, clamp edges.
#include <cmath>
#include <cstring>
#include <vector>
#include "b2ref.h"

namespace {

enum { F_DEFAULT = 0, F_CUBICBSPLINE = 3, F_CATMULLROM = 4, F_MITCHELL = 5, F_MKS_2013_86 = 11, F_MKS_2021 = 13 };   // stb_image_resize.d:270-286

// ---- filter kernels, stb_image_resize.d:659-774.  D promotes like C: a double literal makes the expression double.
float filterCubic(float x) {   // :659-669
    x = std::fabs(x);
    if (x < 1.0f) return (4 + x * x * (3 * x - 6)) / 6;
    else if (x < 2.0f) return (8 + x * (-12 + x * (6 - x))) / 6;
    return 0.0f;
}
float filterCatmullRom(float x) {   // :671-681
    x = std::fabs(x);
    if (x < 1.0f) return 1 - x * x * (2.5f - 1.5f * x);
    else if (x < 2.0f) return 2 - x * (4 + x * (0.5f * x - 2.5f));
    return 0.0f;
}
float filterMitchell(float x) {   // :683-693
    x = std::fabs(x);
    if (x < 1.0f) return (16 + x * x * (21 * x - 36)) / 18;
    else if (x < 2.0f) return (32 + x * (-60 + x * (36 - 7 * x))) / 18;
    return 0.0f;
}
float filterMk2013(float x) {   // :711-721
    x = std::fabs(x);
    if (x < 0.5) return (float)(0.75 - x * x);
    if (x < 1.5) return (float)(0.5 * (x - 1.5) * (x - 1.5));
    return 0.0f;
}
float filterMks2013(float x) {   // :730-751
    x = std::fabs(x);
    if (x <= 1.17549435e-38f) return 17.0f / 16.0f;
    if (x < 0.5) return (float)(17.0 / 16.0 - 7.0 * x * x / 4.0);
    if (x < 1.5) {
        double x2 = x * x;   // float product, widened
        return (float)(0.25 * (4 * x2 - 11.0 * x + 7.0));
    }
    if (x < 2.5) return (float)(-0.125 * (x - 5.0 / 2.0) * (x - 5.0 / 2.0));
    return 0.0f;
}
float filterMks2013_86(float x) { return 0.14f * filterMk2013(x) + 0.86f * filterMks2013(x); }   // :723-728
float filterMks2021(float x) {   // :753-774
    x = std::fabs(x);
    float x2 = x * x;
    if (x < 0.5) return 577.0f / 576.0f - (239.0f / 144.0f) * x2;
    if (x < 1.5) return (140 * x2 - 379 * x + 239) / 144.0f;
    if (x < 2.5) return -(24 * x2 - 113 * x + 130) / 144.0f;
    if (x < 3.5) return (4 * x2 - 27 * x + 45) / 144.0f;
    if (x < 4.5) return -(4 * x2 - 36 * x + 81) / 1152.0f;
    return 0.0f;
}
float kernelOf(int filter, float x) {   // stbir__filter_info_table, :806-822 (the scale argument is unused by these kernels)
    switch (filter) {
        case F_CUBICBSPLINE: return filterCubic(x);
        case F_CATMULLROM: return filterCatmullRom(x);
        case F_MITCHELL: return filterMitchell(x);
        case F_MKS_2013_86: return filterMks2013_86(x);
        default: return filterMks2021(x);
    }
}
float supportOf(int filter) {   // :806-822
    switch (filter) {
        case F_CUBICBSPLINE: case F_CATMULLROM: case F_MITCHELL: return 2;
        case F_MKS_2013_86: return 3;
        default: return 5;
    }
}

bool useUpsampling(float ratio) { return ratio > 1; }   // :825-828
int filterPixelWidth(int filter, float scale) {          // :842-851
    if (useUpsampling(scale)) return (int)std::ceil(supportOf(filter) * 2);
    return (int)std::ceil(supportOf(filter) * 2 / scale);
}
int filterPixelMargin(int filter, float scale) { return filterPixelWidth(filter, scale) / 2; }   // :855-858
int coefficientWidth(int filter, float scale) { (void)scale; return (int)std::ceil(supportOf(filter) * 2); }   // :860-866 (both branches)
int numContributors(float scale, int filter, int inSize, int outSize) {   // :868-874
    return useUpsampling(scale) ? outSize : inSize + filterPixelMargin(filter, scale) * 2;
}

struct Contrib { int n0, n1; };
struct Axis {
    int filter; float scale; int inSize, outSize, margin, coefWidth;
    std::vector<Contrib> contributors;
    std::vector<float> coefficients;
    float* coef(int n, int c) { return &coefficients[(size_t)coefWidth * n + c]; }   // :895-899
};

// stbir__calculate_filters and what it calls, :967-1190 (shift = 0: the resizer maps the whole image, :196-215, :2270-2292)
void calculateFilters(Axis& A) {
    const int filter = A.filter;
    const float scale = A.scale, shift = 0;
    const int total = numContributors(scale, filter, A.inSize, A.outSize);
    A.contributors.assign(total, Contrib{0, 0});
    A.coefficients.assign((size_t)total * A.coefWidth + 16, 0.0f);   // + slack: a group may spill one entry into the next (the port's arrays are contiguous too)
    if (useUpsampling(scale)) {
        const float outRadius = supportOf(filter) * scale;   // :1160
        for (int n = 0; n < total; ++n) {
            // stbir__calculate_sample_range_upsample, :967-980
            const float outCenter = (float)n + 0.5f;
            const float inLo = ((outCenter - outRadius) + shift) / scale, inHi = ((outCenter + outRadius) + shift) / scale;
            const float inCenterOfOut = (outCenter + shift) / scale;
            int first = (int)std::floor(inLo + 0.5), last = (int)std::floor(inHi - 0.5);
            // stbir__calculate_coefficients_upsample, :996-1044
            Contrib& c = A.contributors[n];
            float* g = A.coef(n, 0);
            float totalFilter = 0;
            c.n0 = first; c.n1 = last;
            for (int i = 0; i <= last - first; ++i) {
                const float inPixelCenter = (float)(i + first) + 0.5f;
                g[i] = kernelOf(filter, inCenterOfOut - inPixelCenter);
                if (i == 0 && !g[i]) { c.n0 = ++first; --i; continue; }
                totalFilter += g[i];
            }
            const float filterScale = 1 / totalFilter;
            for (int i = 0; i <= last - first; ++i) g[i] *= filterScale;
            for (int i = last - first; i >= 0; --i) {
                if (g[i]) break;
                c.n1 = c.n0 + i - 1;
            }
        }
        return;
    }
    const float inRadius = supportOf(filter) / scale;   // :1174
    const int margin = filterPixelMargin(filter, scale);
    for (int n = 0; n < total; ++n) {
        // stbir__calculate_sample_range_downsample, :982-994
        const float inCenter = (float)(n - margin) + 0.5f;
        const float outLo = (inCenter - inRadius) * scale - shift, outHi = (inCenter + inRadius) * scale - shift;
        const float outCenterOfIn = inCenter * scale - shift;
        const int first = (int)std::floor(outLo + 0.5), last = (int)std::floor(outHi - 0.5);
        // stbir__calculate_coefficients_downsample, :1046-1074
        Contrib& c = A.contributors[n];
        float* g = A.coef(n, 0);
        c.n0 = first; c.n1 = last;
        for (int i = 0; i <= last - first; ++i) {
            const float outPixelCenter = (float)(i + first) + 0.5f;
            g[i] = kernelOf(filter, outPixelCenter - outCenterOfIn) * scale;
        }
        for (int i = last - first; i >= 0; --i) {
            if (g[i]) break;
            c.n1 = c.n0 + i - 1;
        }
    }
    // stbir__normalize_downsample_coefficients, :1076-1149
    const int numCoefficients = A.coefWidth;
    for (int i = 0; i < A.outSize; ++i) {
        float totalW = 0;
        for (int j = 0; j < total; ++j) {
            if (i >= A.contributors[j].n0 && i <= A.contributors[j].n1) totalW += *A.coef(j, i - A.contributors[j].n0);
            else if (i < A.contributors[j].n0) break;
        }
        const float s = 1 / totalW;
        for (int j = 0; j < total; ++j) {
            if (i >= A.contributors[j].n0 && i <= A.contributors[j].n1) *A.coef(j, i - A.contributors[j].n0) *= s;
            else if (i < A.contributors[j].n0) break;
        }
    }
    for (int j = 0; j < total; ++j) {
        int skip = 0;
        while (skip < A.coefWidth && *A.coef(j, skip) == 0) skip++;   // (the port has no bound: an all-zero group does not occur)
        A.contributors[j].n0 += skip;
        while (A.contributors[j].n0 < 0) { A.contributors[j].n0++; skip++; }
        const int range = A.contributors[j].n1 - A.contributors[j].n0 + 1;
        const int mx = numCoefficients < range ? numCoefficients : range;
        for (int i = 0; i < mx; ++i) {
            if (i + skip >= A.coefWidth) break;
            *A.coef(j, i) = *A.coef(j, i + skip);
        }
    }
    for (int i = 0; i < total; ++i)
        if (A.contributors[i].n1 > A.outSize - 1) A.contributors[i].n1 = A.outSize - 1;
}

int edgeClamp(int n, int mx) { return n < 0 ? 0 : (n >= mx ? mx - 1 : n); }   // stbir__edge_wrap with STBIR_EDGE_CLAMP, :901-965

struct Job {
    const uint8_t* in; size_t inPitch; uint8_t* out; size_t outPitch;
    int channels, bytes;   // bytes per channel: 1 (uint8) or 2 (uint16)
    Axis H, V;
    std::vector<float> decode;   // (inW + 2 margin) * channels, index 0 = pixel -margin
};

// stbir__decode_scanline, :1205-1316 (uint8 / uint16, linear, clamp; alpha_channel < 0 => no alpha weighting, :2371-2372)
void decodeScanline(Job& J, int n) {
    const int ch = J.channels, margin = J.H.margin, inW = J.H.inSize;
    const uint8_t* row = J.in + (size_t)edgeClamp(n, J.V.inSize) * J.inPitch;
    for (int x = -margin; x < inW + margin; ++x) {
        const int sx = edgeClamp(x, inW) * ch;
        for (int c = 0; c < ch; ++c) {
            float v;
            if (J.bytes == 1) v = (float)row[sx + c] / 255.0f;
            else v = ((const uint16_t*)row)[sx + c] / 65535.0f;
            J.decode[(size_t)(x + margin) * ch + c] = v;
        }
    }
}
// stbir__resample_horizontal_upsample / _downsample, :1457-1679: `out` (outW * channels) must be zero on entry
void resampleHorizontal(Job& J, float* out) {
    const int ch = J.channels, margin = J.H.margin;
    const float* dec = J.decode.data() + (size_t)margin * ch;   // stbir__get_decode_buffer, :1193-1198
    if (useUpsampling(J.H.scale)) {
        for (int x = 0; x < J.H.outSize; ++x) {
            const int n0 = J.H.contributors[x].n0, n1 = J.H.contributors[x].n1;
            int counter = 0;
            for (int k = n0; k <= n1; ++k) {
                const float coefficient = *J.H.coef(x, counter++);
                for (int c = 0; c < ch; ++c) out[x * ch + c] += dec[k * ch + c] * coefficient;
            }
        }
        return;
    }
    const int maxX = J.H.inSize + margin * 2;
    for (int x = 0; x < maxX; ++x) {
        const int n0 = J.H.contributors[x].n0, n1 = J.H.contributors[x].n1;
        const int inX = x - margin;
        for (int k = n0; k <= n1; ++k) {
            const float coefficient = *J.H.coef(x, k - n0);
            for (int c = 0; c < ch; ++c) out[k * ch + c] += dec[inX * ch + c] * coefficient;
        }
    }
}
// stbir__encode_scanline, :1719-1843 (linear uint8 / uint16; STBIR__ROUND_INT(saturate(f) * max))
void encodeScanline(Job& J, int row, const float* buf) {
    uint8_t* o = J.out + (size_t)row * J.outPitch;
    const int n = J.H.outSize * J.channels;
    for (int i = 0; i < n; ++i) {
        float f = buf[i];
        f = f < 0 ? 0 : (f > 1 ? 1 : f);   // stbir__saturate, :484-493
        if (J.bytes == 1) o[i] = (uint8_t)(int)(f * 255.0f + 0.5f);
        else ((uint16_t*)o)[i] = (uint16_t)(int)(f * 65535.0f + 0.5f);
    }
}

// stbir__buffer_loop_upsample, :2125-2174, with the ring buffer (:1427-1455, :1712-1717) and
// stbir__resample_vertical_upsample, :1926-2041
void loopUpsample(Job& J) {
    const int rowLen = J.H.outSize * J.channels;
    const int entries = filterPixelWidth(J.V.filter, J.V.scale) + 1;   // :2314
    std::vector<float> ring((size_t)entries * rowLen), encode(rowLen);
    int begin = -1, firstLine = 0, lastLine = 0;
    auto addEntry = [&](int n) -> float* {
        int idx;
        lastLine = n;
        if (begin < 0) { idx = begin = 0; firstLine = n; }
        else idx = (begin + (lastLine - firstLine)) % entries;
        float* e = &ring[(size_t)idx * rowLen];
        memset(e, 0, sizeof(float) * rowLen);
        return e;
    };
    auto decodeAndResample = [&](int n) { decodeScanline(J, n); resampleHorizontal(J, addEntry(n)); };   // :1681-1693
    const float outRadius = supportOf(J.V.filter) * J.V.scale;
    for (int y = 0; y < J.V.outSize; ++y) {
        const float outCenter = (float)y + 0.5f;
        const float inLo = (outCenter - outRadius) / J.V.scale, inHi = (outCenter + outRadius) / J.V.scale;
        const int inFirst = (int)std::floor(inLo + 0.5), inLast = (int)std::floor(inHi - 0.5);
        if (begin >= 0) {
            while (inFirst > firstLine) {
                if (firstLine == lastLine) { begin = -1; firstLine = 0; lastLine = 0; break; }
                firstLine++;
                begin = (begin + 1) % entries;
            }
        }
        if (begin < 0) decodeAndResample(inFirst);
        while (inLast > lastLine) decodeAndResample(lastLine + 1);
        // vertical gather of output row y
        const int n0 = J.V.contributors[y].n0, n1 = J.V.contributors[y].n1;
        memset(encode.data(), 0, sizeof(float) * rowLen);
        int counter = 0;
        for (int k = n0; k <= n1; ++k) {
            const float* e = &ring[(size_t)((begin + (k - firstLine)) % entries) * rowLen];
            const float coefficient = *J.V.coef(y, counter++);
            for (int i = 0; i < rowLen; ++i) encode[i] += e[i] * coefficient;
        }
        encodeScanline(J, y, encode.data());
    }
}

// stbir__buffer_loop_downsample, :2220-2259, stbir__empty_ring_buffer, :2176-2218, stbir__resample_vertical_downsample, :2043-2123
void loopDownsample(Job& J) {
    const int rowLen = J.H.outSize * J.channels;
    const int entries = filterPixelWidth(J.V.filter, J.V.scale) + 1;
    std::vector<float> ring((size_t)entries * rowLen), hbuf(rowLen);
    int begin = -1, firstLine = 0, lastLine = 0;
    auto addEntry = [&](int n) {
        int idx;
        lastLine = n;
        if (begin < 0) { idx = begin = 0; firstLine = n; }
        else idx = (begin + (lastLine - firstLine)) % entries;
        memset(&ring[(size_t)idx * rowLen], 0, sizeof(float) * rowLen);
    };
    auto emptyRing = [&](int firstNecessary) {
        if (begin < 0) return;
        while (firstNecessary > firstLine) {
            if (firstLine >= 0 && firstLine < J.V.outSize) encodeScanline(J, firstLine, &ring[(size_t)begin * rowLen]);
            if (firstLine == lastLine) { begin = -1; firstLine = 0; lastLine = 0; break; }
            firstLine++;
            begin = (begin + 1) % entries;
        }
    };
    const float inRadius = supportOf(J.V.filter) / J.V.scale;
    const int margin = J.V.margin, maxY = J.V.inSize + margin;
    for (int y = -margin; y < maxY; ++y) {
        const float inCenter = (float)y + 0.5f;
        const float outLo = (inCenter - inRadius) * J.V.scale, outHi = (inCenter + inRadius) * J.V.scale;
        const int outFirst = (int)std::floor(outLo + 0.5), outLast = (int)std::floor(outHi - 0.5);
        if (outLast < 0 || outFirst >= J.V.outSize) continue;
        emptyRing(outFirst);
        decodeScanline(J, y);                                  // stbir__decode_and_resample_downsample, :1695-1709
        memset(hbuf.data(), 0, sizeof(float) * rowLen);
        resampleHorizontal(J, hbuf.data());
        if (begin < 0) addEntry(outFirst);
        while (outLast > lastLine) addEntry(lastLine + 1);
        const int contributor = y + margin;
        const int n0 = J.V.contributors[contributor].n0, n1 = J.V.contributors[contributor].n1;
        for (int k = n0; k <= n1; ++k) {
            const float coefficient = *J.V.coef(contributor, k - n0);
            float* e = &ring[(size_t)((begin + (k - firstLine)) % entries) * rowLen];
            for (int i = 0; i < rowLen; ++i) e[i] += hbuf[i] * coefficient;
        }
    }
    emptyRing(J.V.outSize);
}

void setupAxis(Axis& A, int filter, int inSize, int outSize) {
    A.inSize = inSize; A.outSize = outSize;
    A.scale = (float)outSize / inSize;                         // stbir__calculate_transform, :2270-2292 (s0 = t0 = 0, s1 = t1 = 1)
    if (filter == F_DEFAULT) filter = useUpsampling(A.scale) ? F_CATMULLROM : F_MITCHELL;   // stbir__choose_filter, :2294-2302; :364-366
    A.filter = filter;
    A.margin = filterPixelMargin(filter, A.scale);
    A.coefWidth = coefficientWidth(filter, A.scale);
    calculateFilters(A);
}

}  // namespace

// kind: 0 resizeImageDiffuse (RGBA, MKS 2013 86 %; resizer.d:346-372), 1 resizeImageMaterial = resizeImageGeneric(RGBA)
// (default filters; resizer.d:39-64, :374-376), 2 resizeImageDepth (L16, MKS 2021; resizer.d:171-197), 3 resizeImageGeneric(L16)
// (resizer.d:144-169), 4 resizeImageSmoother (RGBA, cubic B-spline; resizer.d:92-117).  The `else` (v1) branches.
extern "C" int b2ref_resize_image(int kind, const void* in, int in_w, int in_h, size_t in_pitch, void* out, int out_w, int out_h, size_t out_pitch) {
    if (kind < 0 || kind > 4 || !in || !out || in_w < 1 || in_h < 1 || out_w < 1 || out_h < 1) return 1;
    const bool l16 = kind == 2 || kind == 3;
    const int channels = l16 ? 1 : 4, bytes = l16 ? 2 : 1;
    if (in_w == out_w && in_h == out_h) {                      // sameSizeResize, resizer.d:453-465
        for (int y = 0; y < in_h; ++y) memcpy((uint8_t*)out + (size_t)y * out_pitch, (const uint8_t*)in + (size_t)y * in_pitch, (size_t)in_w * channels * bytes);
        return 0;
    }
    static const int filterOfKind[5] = {F_MKS_2013_86, F_DEFAULT, F_MKS_2021, F_DEFAULT, F_CUBICBSPLINE};
    Job J;
    J.in = (const uint8_t*)in; J.inPitch = in_pitch; J.out = (uint8_t*)out; J.outPitch = out_pitch;
    J.channels = channels; J.bytes = bytes;
    setupAxis(J.H, filterOfKind[kind], in_w, out_w);
    setupAxis(J.V, filterOfKind[kind], in_h, out_h);
    J.decode.assign((size_t)(in_w + 2 * J.H.margin) * channels, 0.0f);
    if (useUpsampling(J.V.scale)) loopUpsample(J);
    else loopDownsample(J);
    return 0;
}
