// Device-side numeric building blocks of the lighting pass.
//
// The translation units that include this header are compiled with -fmad=false
// -prec-div=true -prec-sqrt=true -ftz=false, so every * + - / below is one IEEE
// binary32 operation, like the reference's SSE code (no FMA on baseline x86-64).
#pragma once
#include "common.cuh"

namespace b2 {

static __device__ const unsigned char kBlueNoise16x16[256] = {  // gui/dplug/gui/legacypbr.d:1012-1030
    127, 194, 167, 79,  64,  173, 22,  83,  167, 105, 119, 250, 201, 34,  214, 145,
    233, 56,  13,  251, 203, 124, 243, 42,  216, 34,  73,  175, 133, 64,  185, 73,
    93,  156, 109, 144, 34,  98,  153, 138, 187, 238, 155, 46,  13,  102, 247, 0,
    28,  180, 46,  218, 183, 13,  212, 69,  13,  92,  126, 228, 211, 161, 117, 197,
    134, 240, 121, 75,  234, 88,  53,  170, 109, 204, 59,  22,  86,  141, 38,  222,
    81,  205, 13,  59,  160, 198, 129, 252, 0,   147, 176, 193, 244, 71,  173, 56,
    22,  168, 104, 139, 22,  114, 38,  220, 101, 231, 77,  34,  113, 13,  189, 96,
    253, 148, 227, 190, 246, 174, 66,  155, 28,  50,  164, 131, 217, 151, 232, 128,
    115, 69,  34,  50,  93,  13,  209, 85,  192, 120, 248, 64,  90,  28,  208, 42,
    0,   200, 215, 79,  125, 148, 239, 136, 181, 22,  206, 13,  185, 108, 59,  179,
    90,  130, 159, 182, 235, 42,  106, 0,   56,  99,  226, 140, 157, 237, 77,  165,
    249, 28,  105, 13,  61,  170, 224, 75,  202, 163, 114, 81,  46,  22,  137, 223,
    189, 53,  219, 142, 196, 28,  122, 154, 254, 42,  28,  242, 196, 210, 119, 38,
    149, 86,  118, 245, 71,  96,  213, 13,  88,  178, 66,  129, 171, 0,   99,  69,
    178, 13,  207, 38,  159, 187, 50,  132, 236, 146, 191, 95,  53,  229, 163, 241,
    46,  225, 102, 135, 0,   230, 110, 199, 61,  0,   221, 22,  150, 83,  112, 22};

// x86 RSQRTSS as the reference executes it (gui/dplug/gui/ransac.d:43): a 2048-bin table indexed by
// exponent parity and the top 10 mantissa bits, exact power-of-two scaling (see tools/gen_rsqrtss_table.c).
__device__ __forceinline__ float rsqrtss_x86(const uint32_t* __restrict__ tab, float x) {
    const uint32_t b = __float_as_uint(x);
    const uint32_t e = b >> 23;  // sign bit included: negative operands give e >= 256
    // normal positive operand: table entry of the bin, exponent shifted by floor((e - 126) / 2)
    const int de = ((int)e - 126) >> 1;
    float r = __uint_as_float(__ldg(tab + ((b >> 13) & 0x7ffu)) - ((uint32_t)de << 23));
    if (e - 1u >= 254u) {  // zero / denormal / inf / NaN / negative
        if ((b << 1) > 0xff000000u) r = x + x;
        else if (b & 0x80000000u) r = __uint_as_float(((b & 0x7f800000u) == 0) ? 0xff800000u : 0xffc00000u);
        else r = __uint_as_float(e == 0 ? 0x7f800000u : 0u);
    }
    return r;
}

// log_ps / exp_ps of sse_mathfun (the algorithm behind intel-intrinsics' _mm_pow_ps, which
// gui/dplug/gui/legacypbr.d:786-790 calls); one IEEE op per statement.
__device__ __forceinline__ float cephes_logf(float x) {
    const bool invalid = x <= 0.0f;
    const float mn = __uint_as_float(0x00800000u);
    x = (x > mn) ? x : mn;
    uint32_t xb = __float_as_uint(x);
    int emm0 = (int)(xb >> 23) - 0x7f;
    x = __uint_as_float((xb & ~0x7f800000u) | 0x3f000000u);
    float e = (float)emm0;
    e = e + 1.0f;
    const bool below = x < 0.707106781186547524f;
    float tmp = below ? x : 0.0f;
    x = x - 1.0f;
    e = e - (below ? 1.0f : 0.0f);
    x = x + tmp;
    float z = x * x;
    float y = 7.0376836292E-2f;
    y = y * x; y = y + -1.1514610310E-1f;
    y = y * x; y = y + 1.1676998740E-1f;
    y = y * x; y = y + -1.2420140846E-1f;
    y = y * x; y = y + 1.4249322787E-1f;
    y = y * x; y = y + -1.6668057665E-1f;
    y = y * x; y = y + 2.0000714765E-1f;
    y = y * x; y = y + -2.4999993993E-1f;
    y = y * x; y = y + 3.3333331174E-1f;
    y = y * x;
    y = y * z;
    tmp = e * -2.12194440e-4f;
    y = y + tmp;
    tmp = z * 0.5f;
    y = y - tmp;
    tmp = e * 0.693359375f;
    x = x + y;
    x = x + tmp;
    return invalid ? __uint_as_float(0xffffffffu) : x;
}

__device__ __forceinline__ float cephes_expf(float x) {
    x = (x < 88.3762626647949f) ? x : 88.3762626647949f;
    x = (x > -88.3762626647949f) ? x : -88.3762626647949f;
    float fx = x * 1.44269504088896341f;
    fx = fx + 0.5f;
    float tmp = (float)__float2int_rz(fx);
    if (tmp > fx) tmp = tmp - 1.0f;
    fx = tmp;
    tmp = fx * 0.693359375f;
    float z = fx * -2.12194440e-4f;
    x = x - tmp;
    x = x - z;
    z = x * x;
    float y = 1.9875691500E-4f;
    y = y * x; y = y + 1.3981999507E-3f;
    y = y * x; y = y + 8.3334519073E-3f;
    y = y * x; y = y + 4.1665795894E-2f;
    y = y * x; y = y + 1.6666665459E-1f;
    y = y * x; y = y + 5.0000001201E-1f;
    y = y * z;
    y = y + x;
    y = y + 1.0f;
    int n = __float2int_rz(fx);
    return y * __uint_as_float((uint32_t)(n + 0x7f) << 23);
}

__device__ __forceinline__ float pow_ps(float base, float exponent) { return cephes_expf(exponent * cephes_logf(base)); }

// gui/dplug/gui/legacypbr.d:1117-1133
__device__ __forceinline__ float fastlog2(float v) {
    int x = __float_as_int(v);
    int log_2 = ((x >> 23) & 255) - 128;
    x = x & ~(255 << 23);
    x += 127 << 23;
    return __int_as_float(x) + (float)log_2;
}

struct F4 { float x, y, z, w; };
__device__ __forceinline__ F4 operator*(F4 a, float s) { return F4{a.x * s, a.y * s, a.z * s, a.w * s}; }
__device__ __forceinline__ F4 operator*(float s, F4 a) { return F4{s * a.x, s * a.y, s * a.z, s * a.w}; }
__device__ __forceinline__ F4 operator+(F4 a, F4 b) { return F4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
__device__ __forceinline__ F4 operator-(F4 a, F4 b) { return F4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }

__device__ __forceinline__ F4 unpackRGBA(uint32_t p) {
    return F4{(float)(p & 0xffu), (float)((p >> 8) & 0xffu), (float)((p >> 16) & 0xffu), (float)(p >> 24)};
}

// Catmull-Rom with the reference's operation order (graphics/dplug/graphics/mipmap.d:222-229, 263-270)
template <typename V>
__device__ __forceinline__ V cubicInterp(float t, V x0, V x1, V x2, V x3) {
    return x1
         + t * ((-0.5f * x0) + (0.5f * x2))
         + t * t * (x0 - (2.5f * x1) + (2.0f * x2) - (0.5f * x3))
         + t * t * t * ((-0.5f * x0) + (1.5f * x1) - (1.5f * x2) + 0.5f * x3);
}

struct Bilin { int ix, iy, ix1, iy1; float fx, fy, fxm1, fym1; };
// coordinate set-up of Mipmap.linearSample (mipmap.d:309-342); scale = 2^-level
__device__ __forceinline__ Bilin bilinSetup(float scale, float x, float y, int w, int h) {
    x = x * scale - 0.5f;
    y = y * scale - 0.5f;
    if (x < 0) x = 0;
    if (y < 0) y = 0;
    Bilin s;
    s.ix = __float2int_rz(x);
    s.iy = __float2int_rz(y);
    s.fx = x - (float)s.ix;
    s.fy = y - (float)s.iy;
    const int maxX = w - 1, maxY = h - 1;
    s.ix = min(s.ix, maxX);
    s.iy = min(s.iy, maxY);
    s.ix1 = min(s.ix + 1, maxX);
    s.iy1 = min(s.iy + 1, maxY);
    s.fxm1 = 1 - s.fx;
    s.fym1 = 1 - s.fy;
    return s;
}

__device__ __forceinline__ float levelScale(int level) { return __uint_as_float((uint32_t)(127 - level) << 23); }

// Mipmap!RGBA.linearSample — mipmap.d:293-385 (version(LDC) branch)
__device__ __forceinline__ F4 linearSampleRGBA(const DevLevel* levels, int numLevels, int level, float x, float y) {
    level = max(0, min(level, numLevels - 1));
    const DevLevel& L = levels[level];
    Bilin s = bilinSetup(levelScale(level), x, y, L.w, L.h);
    const uint32_t* r0 = rowPtr<uint32_t>(L, s.iy);
    const uint32_t* r1 = rowPtr<uint32_t>(L, s.iy1);
    F4 A = unpackRGBA(__ldg(r0 + s.ix)), B = unpackRGBA(__ldg(r0 + s.ix1));
    F4 C = unpackRGBA(__ldg(r1 + s.ix)), D = unpackRGBA(__ldg(r1 + s.ix1));
    F4 up = A * s.fxm1 + B * s.fx;
    F4 down = C * s.fxm1 + D * s.fx;
    return up * s.fym1 + down * s.fy;
}

// Mipmap!L16.linearSample — mipmap.d:478-483
__device__ __forceinline__ float linearSampleL16(const DevLevel* levels, int numLevels, int level, float x, float y) {
    level = max(0, min(level, numLevels - 1));
    const DevLevel& L = levels[level];
    Bilin s = bilinSetup(levelScale(level), x, y, L.w, L.h);
    const uint16_t* r0 = rowPtr<uint16_t>(L, s.iy);
    const uint16_t* r1 = rowPtr<uint16_t>(L, s.iy1);
    float A = (float)__ldg(r0 + s.ix), B = (float)__ldg(r0 + s.ix1);
    float C = (float)__ldg(r1 + s.ix), D = (float)__ldg(r1 + s.ix1);
    float up = A * s.fxm1 + B * s.fx;
    float down = C * s.fxm1 + D * s.fx;
    return up * s.fym1 + down * s.fy;
}

struct Bicub { int xi[4], yi[4]; float a, b; };
// coordinate set-up of Mipmap.cubicSample (mipmap.d:182-204)
__device__ __forceinline__ Bicub bicubSetup(float scale, float x, float y, int w, int h) {
    x = x * scale - 0.5f;
    y = y * scale - 0.5f;
    Bicub s;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s.xi[k] = max(0, min(__float2int_rz(x + (float)(k - 1)), w - 1));
        s.yi[k] = max(0, min(__float2int_rz(y + (float)(k - 1)), h - 1));
    }
    float a = x + 1.0f, b = y + 1.0f;
    s.a = a - (float)__float2int_rz(a);
    s.b = b - (float)__float2int_rz(b);
    return s;
}

// Mipmap!RGBA.cubicSample — mipmap.d:167-287
__device__ __forceinline__ F4 cubicSampleRGBA(const DevLevel* levels, int numLevels, int level, float x, float y) {
    level = max(0, min(level, numLevels - 1));
    const DevLevel& L = levels[level];
    Bicub s = bicubSetup(levelScale(level), x, y, L.w, L.h);
    F4 R[4];
#pragma unroll
    for (int row = 0; row < 4; ++row) {
        const uint32_t* p = rowPtr<uint32_t>(L, s.yi[row]);
        R[row] = cubicInterp<F4>(s.a, unpackRGBA(__ldg(p + s.xi[0])), unpackRGBA(__ldg(p + s.xi[1])),
                                 unpackRGBA(__ldg(p + s.xi[2])), unpackRGBA(__ldg(p + s.xi[3])));
    }
    F4 r = cubicInterp<F4>(s.b, R[0], R[1], R[2], R[3]);
    r.x = r.x < 0 ? 0.f : r.x; r.y = r.y < 0 ? 0.f : r.y; r.z = r.z < 0 ? 0.f : r.z; r.w = r.w < 0 ? 0.f : r.w;
    r.x = r.x > 65535 ? 65535.f : r.x; r.y = r.y > 65535 ? 65535.f : r.y;
    r.z = r.z > 65535 ? 65535.f : r.z; r.w = r.w > 65535 ? 65535.f : r.w;
    return r;
}

// Mipmap!L16.cubicSample — mipmap.d:214-246
__device__ __forceinline__ float cubicSampleL16(const DevLevel* levels, int numLevels, int level, float x, float y) {
    level = max(0, min(level, numLevels - 1));
    const DevLevel& L = levels[level];
    Bicub s = bicubSetup(levelScale(level), x, y, L.w, L.h);
    float R[4];
#pragma unroll
    for (int row = 0; row < 4; ++row) {
        const uint16_t* p = rowPtr<uint16_t>(L, s.yi[row]);
        R[row] = cubicInterp<float>(s.a, (float)__ldg(p + s.xi[0]), (float)__ldg(p + s.xi[1]),
                                    (float)__ldg(p + s.xi[2]), (float)__ldg(p + s.xi[3]));
    }
    float r = cubicInterp<float>(s.b, R[0], R[1], R[2], R[3]);
    r = r < 0 ? 0.f : r;
    r = r > 65535 ? 65535.f : r;
    return r;
}

// Mipmap!RGBA.linearMipmapSample — mipmap.d:149-160
__device__ __forceinline__ F4 linearMipmapSampleRGBA(const DevLevel* levels, int numLevels, float level, float x, float y) {
    int il = __float2int_rz(level);
    float fl = level - (float)il;
    F4 a = linearSampleRGBA(levels, numLevels, il, x, y);
    if (fl == 0) return a;
    F4 b = linearSampleRGBA(levels, numLevels, il + 1, x, y);
    return a * (1 - fl) + b * fl;
}

}  // namespace b2
