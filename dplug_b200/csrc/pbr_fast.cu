// Lighting kernel, throughput variant (the one compositeTile launches by default).
//
// Same eight passes as gui/dplug/gui/legacypbr.d:269-1108, fused, but organised for the SM instead of for
// a 64x64 CPU tile:
//
//   * work item = one WARP = a strip of 32 adjacent columns x up to STRIP_ROWS rows of one GUI tile; a lane
//     owns one column and walks down it, so every per-row quantity that depends only on j is warp-uniform
//     and every per-column quantity is computed once per strip;
//   * level-0 depth of the strip (+10 columns left/right, 10 rows up, 1 down: the reach of the oblique
//     shadow, legacypbr.d:525-534, and of the 3x3 normal stencil, legacypbr.d:301-314) is staged once per
//     warp in shared memory as float with clamp-to-edge addressing (= the reference's replicated border);
//   * the mip samplers (Mipmap.linearSample / cubicSample, graphics/dplug/graphics/mipmap.d:167-496) are
//     evaluated separably: the horizontal interpolation of a texel row is done once and reused for the
//     2^level rows that share it ("rolling rows" in registers), the vertical bicubic of levels 4/5 is a
//     Horner polynomial per texel interval;
//   * 8/16-bit -> float conversions use byte-permute + magic-number subtraction (FMA pipe) instead of I2F
//     (quarter-rate XU pipe), pow / reciprocal use MUFU.
//
// Numerical contract: +-1 LSB per 8-bit channel against the reference arithmetic (BASELINE.json).  The
// three RSQRTSS operands per pixel are computed with the reference's exact IEEE sequence (no FMA, same
// order) because the table-driven RSQRTSS is discontinuous and Blinn-Phong exponents up to 548 amplify
// it; everything downstream may use FMA / MUFU (relative error ~1e-6).
#include "pbr_math.cuh"

namespace b2 {

#ifndef B2_STRIP_ROWS
#define B2_STRIP_ROWS 32
#endif
constexpr int STRIP_ROWS = B2_STRIP_ROWS;
constexpr int HALO_X = 10;                     // shadow reach left/right
constexpr int HALO_UP = 10;                    // shadow reach upwards
constexpr int TILE_COLS = 32 + 2 * HALO_X + 2; // +2: the staged window starts at an even column
constexpr int TILE_ROWS = STRIP_ROWS + HALO_UP + 1;
constexpr int WARPS_PER_CTA = 4;

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return F3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return F3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return F3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ F3 fma3(float s, F3 a, F3 b) { return F3{fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)}; }

// byte k of a packed RGBA8 texel -> float, exact, without I2F: (0x4B000000 | b) is the float 2^23 + b
__device__ __forceinline__ float byteToFloat(uint32_t p, int k) {
    return __uint_as_float(__byte_perm(p, 0x4B000000u, 0x7440 + k)) - 8388608.0f;
}
__device__ __forceinline__ F3 rgbOf(uint32_t p) { return F3{byteToFloat(p, 0), byteToFloat(p, 1), byteToFloat(p, 2)}; }
__device__ __forceinline__ float u16ToFloat(uint32_t v) { return __uint_as_float(0x4B000000u | v) - 8388608.0f; }

__device__ __forceinline__ float fmaSat(float a, float b, float c) {
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float fastRcp(float x) {  // operands here are >= 0.7: no denormal handling needed
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fastPow(float x, float y) {  // x in [1e-3, ~1], y in (0, 548]: ex2(y * lg2(x))
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y * l));
    return r;
}
__device__ __forceinline__ uint32_t cvtSatU8(float f) {  // truncate, clamp to [0,255], NaN -> 0
    uint32_t r;
    asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(f));
    return r & 0xffu;
}

// horizontal set-up of a sampler for one column (lane-constant during a strip)
struct XLin { int ix, ix1; float fx; };
__device__ __forceinline__ XLin xLinear(float scale, int i, int w) {
    float x = ((float)i + 0.5f) * scale - 0.5f;  // mipmap.d:309-314 (exact: powers of two and small integers)
    x = x < 0 ? 0.f : x;
    XLin s;
    int ix = __float2int_rz(x);
    s.fx = x - (float)ix;
    s.ix = min(ix, w - 1);
    s.ix1 = min(s.ix + 1, w - 1);
    return s;
}
struct XCub { int xi[4]; float w[4]; };
__device__ __forceinline__ void catmullRomWeights(float t, float w[4]) {
    // cubicInterp (mipmap.d:263-270) regrouped per sample: p0*w0 + p1*w1 + p2*w2 + p3*w3
    float t2 = t * t, t3 = t2 * t;
    w[0] = -0.5f * t + t2 - 0.5f * t3;
    w[1] = 1.0f - 2.5f * t2 + 1.5f * t3;
    w[2] = 0.5f * t + 2.0f * t2 - 1.5f * t3;
    w[3] = -0.5f * t2 + 0.5f * t3;
}
__device__ __forceinline__ XCub xCubic(float scale, int i, int w) {
    float x = ((float)i + 0.5f) * scale - 0.5f;  // mipmap.d:182-204
    XCub s;
#pragma unroll
    for (int k = 0; k < 4; ++k) s.xi[k] = max(0, min(__float2int_rz(x + (float)(k - 1)), w - 1));
    float a = x + 1.0f;
    a = a - (float)__float2int_rz(a);
    catmullRomWeights(a, s.w);
    return s;
}

// vertical set-up of a sampler for one row.  (j + 0.5) * 2^-l - 0.5 == n / 2^(l+1) with n = 2j + 1 - 2^l, so the
// texel row and the fraction come from integer shifts on warp-uniform values (uniform datapath); the float
// results are bit-identical to the reference's float arithmetic, which is exact for these operands.
struct YLin { int iy, iy1; float fy; };
__device__ __forceinline__ YLin yLinear(int l, int j, int h) {
    int n = max(2 * j + 1 - (1 << l), 0);                      // y < 0 -> 0 (mipmap.d:315-316)
    YLin s;
    int iy = n >> (l + 1);
    s.fy = (float)(n & ((2 << l) - 1)) * levelScale(l + 1);    // fraction taken before the upper clamp
    s.iy = min(iy, h - 1);
    s.iy1 = min(s.iy + 1, h - 1);
    return s;
}

// ---- rolling rows: bilinear RGB level (diffuse levels 1..3) ------------------------------------------
struct RollRGB {
    int rowA, rowB;
    F3 a, d;  // value of row A and (row B - row A), both already interpolated horizontally for this column
};
__device__ __forceinline__ F3 hrowRGB(const DevLevel& L, const XLin& xs, int row) {
    const uint32_t* p = rowPtr<uint32_t>(L, row);
    F3 A = rgbOf(__ldg(p + xs.ix)), B = rgbOf(__ldg(p + xs.ix1));
    return fma3(xs.fx, B - A, A);
}
__device__ __forceinline__ F3 sampleRollRGB(RollRGB& r, const DevLevel& L, const XLin& xs, const YLin& ys) {
    if (ys.iy != r.rowA || ys.iy1 != r.rowB) {  // warp-uniform
        F3 na = (ys.iy == r.rowB) ? (r.a + r.d) : ((ys.iy == r.rowA) ? r.a : hrowRGB(L, xs, ys.iy));
        F3 nb = (ys.iy1 == ys.iy) ? na : hrowRGB(L, xs, ys.iy1);
        r.a = na; r.d = nb - na;
        r.rowA = ys.iy; r.rowB = ys.iy1;
    }
    return fma3(ys.fy, r.d, r.a);
}

// ---- rolling rows: bilinear L16 level (depth levels 1..4) ----------------------------------------------
struct RollZ { int rowA, rowB; float a, d; };
__device__ __forceinline__ float hrowZ(const DevLevel& L, const XLin& xs, int row) {
    const uint16_t* p = rowPtr<uint16_t>(L, row);
    float A = u16ToFloat(__ldg(p + xs.ix)), B = u16ToFloat(__ldg(p + xs.ix1));
    return fmaf(xs.fx, B - A, A);
}
__device__ __forceinline__ float sampleRollZ(RollZ& r, const DevLevel& L, const XLin& xs, const YLin& ys) {
    if (ys.iy != r.rowA || ys.iy1 != r.rowB) {
        float na = (ys.iy == r.rowB) ? (r.a + r.d) : ((ys.iy == r.rowA) ? r.a : hrowZ(L, xs, ys.iy));
        float nb = (ys.iy1 == ys.iy) ? na : hrowZ(L, xs, ys.iy1);
        r.a = na; r.d = nb - na;
        r.rowA = ys.iy; r.rowB = ys.iy1;
    }
    return fmaf(ys.fy, r.d, r.a);
}

// ---- rolling rows: bicubic RGB level (diffuse levels 4, 5) ------------------------------------------------
struct RollCub {
    int base;     // trunc(y_s) of the interval the polynomial below is valid for (INT_MIN = none)
    F3 c0, c1, c2, c3;
};
__device__ __forceinline__ F3 hrowCub(const DevLevel& L, const XCub& xs, int row) {
    const uint32_t* p = rowPtr<uint32_t>(L, row);
    F3 r = rgbOf(__ldg(p + xs.xi[0])) * xs.w[0];
    r = fma3(xs.w[1], rgbOf(__ldg(p + xs.xi[1])), r);
    r = fma3(xs.w[2], rgbOf(__ldg(p + xs.xi[2])), r);
    r = fma3(xs.w[3], rgbOf(__ldg(p + xs.xi[3])), r);
    return r;
}
__device__ __forceinline__ F3 sampleRollCub(RollCub& r, const DevLevel& L, const XCub& xs, int l, int j) {
    // y_s + 1 = (n + 2^(l+1)) / 2^(l+1), n = 2j + 1 - 2^l (mipmap.d:182-204); always positive, so the
    // reference's truncations are floors and everything below is exact integer arithmetic on uniform values
    const int n1 = 2 * j + 1 - (1 << l) + (2 << l);
    const int ib = n1 >> (l + 1);
    const float t = (float)(n1 & ((2 << l) - 1)) * levelScale(l + 1);
    if (ib != r.base) {
        const int h = L.h;
        F3 p[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int row = max(0, min(ib - 2 + k, h - 1));   // trunc(y_s + k - 1) clamped (mipmap.d:188-193)
            p[k] = hrowCub(L, xs, row);
        }
        // Catmull-Rom as a polynomial in t (mipmap.d:263-270)
        r.c0 = p[1];
        r.c1 = (p[2] - p[0]) * 0.5f;
        r.c2 = F3{p[0].x - 2.5f * p[1].x + 2.0f * p[2].x - 0.5f * p[3].x, p[0].y - 2.5f * p[1].y + 2.0f * p[2].y - 0.5f * p[3].y,
                  p[0].z - 2.5f * p[1].z + 2.0f * p[2].z - 0.5f * p[3].z};
        r.c3 = F3{-0.5f * p[0].x + 1.5f * p[1].x - 1.5f * p[2].x + 0.5f * p[3].x, -0.5f * p[0].y + 1.5f * p[1].y - 1.5f * p[2].y + 0.5f * p[3].y,
                  -0.5f * p[0].z + 1.5f * p[1].z - 1.5f * p[2].z + 0.5f * p[3].z};
        r.base = ib;
    }
    F3 v = fma3(t, r.c3, r.c2);
    v = fma3(t, v, r.c1);
    v = fma3(t, v, r.c0);
    // clamp_0_to_65535 (mipmap.d:250-261); 8-bit sources cannot reach the upper bound
    return F3{fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f)};
}

// ---- skybox: Mipmap!RGBA.linearSample on one level, RGB only (mipmap.d:293-385) ------------------------
__device__ __forceinline__ F3 skySample(const PbrArgs& P, int level, float x, float y) {
    level = max(0, min(level, P.nSky - 1));
    const DevLevel& L = P.sky[level];
    Bilin s = bilinSetup(levelScale(level), x, y, L.w, L.h);
    const uint32_t* r0 = rowPtr<uint32_t>(L, s.iy);
    const uint32_t* r1 = rowPtr<uint32_t>(L, s.iy1);
    uint32_t a = __ldg(r0 + s.ix), b = __ldg(r0 + s.ix1), c = __ldg(r1 + s.ix), d = __ldg(r1 + s.ix1);
    F3 A = rgbOf(a), B = rgbOf(b), C = rgbOf(c), D = rgbOf(d);
    F3 up = fma3(s.fx, B - A, A);
    F3 down = fma3(s.fx, D - C, C);
    return fma3(s.fy, down - up, up);
}

// exact (non-fused, reference-ordered) RSQRTSS operand helpers
__device__ __forceinline__ float rsqrtArgExact(float x, float y, float z) {
    // _mm_fast_normalize_ps: squared = v*v; x2 + y2 + z2 + w2 with w = 0 (ransac.d:41-42)
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));  // (+0 is the identity here)
}

template <bool HAS_SKY>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) k_pbr_fast(const __grid_constant__ PbrArgs P) {
    __shared__ float sDepth[WARPS_PER_CTA][TILE_ROWS][TILE_COLS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x * WARPS_PER_CTA + warp;
    if (item >= P.nBoxes) return;
    const Box box = P.boxes[item];  // one strip: at most 32 columns x STRIP_ROWS rows
    const int W = P.W, H = P.H;
    const DevLevel& D0 = P.depth[0];

    // ---------------------------------------------------------------- stage the depth window
    // columns [cx0, cx0 + TILE_COLS), rows [box.miny - 10, box.maxy + 1), clamped to the image
    const int cx0 = (box.minx - HALO_X) & ~1;
    const int ry0 = box.miny - HALO_UP;
    const int nRows = (box.maxy - box.miny) + HALO_UP + 1;
    float(*tile)[TILE_COLS] = sDepth[warp];
    for (int r = 0; r < nRows; ++r) {
        const int gy = max(0, min(ry0 + r, H - 1));
        const uint16_t* src = rowPtr<uint16_t>(D0, gy);
        const int c = 2 * lane;  // each lane converts a pair of texels
        if (c < TILE_COLS) {
            const int gx = cx0 + c;
            float f0, f1;
            if (gx >= 0 && gx + 1 < W) {
                uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src + gx));
                f0 = u16ToFloat(v & 0xffffu);
                f1 = u16ToFloat(v >> 16);
            } else {
                f0 = u16ToFloat(__ldg(src + max(0, min(gx, W - 1))));
                f1 = u16ToFloat(__ldg(src + max(0, min(gx + 1, W - 1))));
            }
            *reinterpret_cast<float2*>(&tile[r][c]) = make_float2(f0, f1);
        }
    }
    __syncwarp();

    const int i = box.minx + lane;
    const bool active = i < box.maxx;
    const int ic = min(i, box.maxx - 1);        // inactive lanes shadow the last column (no divergence, no stores)
    const int tx = ic - cx0;                    // column of this lane inside the staged window
    const uint32_t mask = P.passMask;

    // ---------------------------------------------------------------- per-column set-up
    const int nDf = P.nDiffuse, nDz = P.nDepth;
    int lvD[6], lvZ[5];
#pragma unroll
    for (int l = 1; l <= 5; ++l) lvD[l] = min(l, nDf - 1);   // level clamp of linearSample / cubicSample
#pragma unroll
    for (int l = 1; l <= 4; ++l) lvZ[l] = min(l, nDz - 1);
    XLin xD1 = xLinear(levelScale(lvD[1]), ic, P.diffuse[lvD[1]].w);
    XLin xD2 = xLinear(levelScale(lvD[2]), ic, P.diffuse[lvD[2]].w);
    XLin xD3 = xLinear(levelScale(lvD[3]), ic, P.diffuse[lvD[3]].w);
    XCub xD4 = xCubic(levelScale(lvD[4]), ic, P.diffuse[lvD[4]].w);
    XCub xD5 = xCubic(levelScale(lvD[5]), ic, P.diffuse[lvD[5]].w);
    XLin xZ1 = xLinear(levelScale(lvZ[1]), ic, P.depth[lvZ[1]].w);
    XLin xZ2 = xLinear(levelScale(lvZ[2]), ic, P.depth[lvZ[2]].w);
    XLin xZ3 = xLinear(levelScale(lvZ[3]), ic, P.depth[lvZ[3]].w);
    XLin xZ4 = xLinear(levelScale(lvZ[4]), ic, P.depth[lvZ[4]].w);
    RollRGB rD1{-1, -1, {0, 0, 0}, {0, 0, 0}}, rD2 = rD1, rD3 = rD1;
    RollCub rD4, rD5;
    rD4.base = rD5.base = (int)0x80000000;
    RollZ rZ1{-1, -1, 0.f, 0.f}, rZ2 = rZ1, rZ3 = rZ1, rZ4 = rZ1;

    // toEye.x and its square are column constants (legacypbr.d:756, 914) — exact sequence
    const float eyeX = __fsub_rn(0.5f, __fmul_rn((float)ic, P.invW));
    const float eyeX2 = __fmul_rn(eyeX, eyeX);
    const float nz2 = __fmul_rn(0.382658847856f, 0.382658847856f);
    const float multUshort = (float)(1.0 / 4655.0);
    const float kA = 0.00006510416f;  // 1/15360 as the reference's SIMD branch spells it (legacypbr.d:555)
    const int noiseCol = (ic & 15) * 16;

    // sliding 3x3 depth window (raw values as float): rows j-1, j, j+1 at columns i-1, i, i+1
    float zm[3], z0[3], zp[3];
    {
        const int r = box.miny - ry0;  // row of j = box.miny inside the window
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // the staged window is clamped per element, so tx-1 / tx+1 are valid: tx >= HALO_X
            zm[k] = tile[r - 1][tx - 1 + k];
            z0[k] = tile[r][tx - 1 + k];
        }
    }

    for (int j = box.miny; j < box.maxy; ++j) {
        const int r = j - ry0;
#pragma unroll
        for (int k = 0; k < 3; ++k) zp[k] = tile[r + 1][tx - 1 + k];

        const uint32_t diffusePx = __ldg(rowPtr<uint32_t>(P.diffuse[0], j) + ic);
        const uint32_t materialPx = __ldg(rowPtr<uint32_t>(P.material, j) + ic);
        const float div255 = 1 / 255.0f;
        const F3 base = rgbOf(diffusePx) * div255;
        const float rough = byteToFloat(materialPx, 0) * div255;
        const float metal = byteToFloat(materialPx, 1) * div255;
        const float specAmt = byteToFloat(materialPx, 2) * div255;
        const int roughByte = materialPx & 0xff;

        // ---- PassComputeNormal (legacypbr.d:279-381) + computePlaneFittingNormal (ransac.d:53-121).
        // Exact reference sequence: the result feeds RSQRTSS.
        float nx = 0, ny = 0, nzc = 0, variance = 0;
        if (mask & 1u) {
            float d[9];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                d[k] = __fmul_rn(zm[k], multUshort);
                d[3 + k] = __fmul_rn(z0[k], multUshort);
                d[6 + k] = __fmul_rn(zp[k], multUshort);
            }
            const float c = 0.03660694846f, e = 0.118115527008f;
            float m0 = __fadd_rn(__fmul_rn(d[0], c), __fmul_rn(d[5], e));
            float m1 = __fadd_rn(__fmul_rn(d[1], e), __fmul_rn(d[6], c));
            float m2 = __fadd_rn(__fmul_rn(d[2], c), __fmul_rn(d[7], e));
            float m3 = __fadd_rn(__fmul_rn(d[3], e), __fmul_rn(d[8], c));
            float filt = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[4], 0.38111009811f), m0), m1), m2), m3);
            float a0 = __fsub_rn(d[0], filt), a1 = __fsub_rn(d[1], filt), a2 = __fsub_rn(d[2], filt), a3 = __fsub_rn(d[3], filt);
            float b0 = __fsub_rn(d[5], filt), b1 = __fsub_rn(d[6], filt), b2 = __fsub_rn(d[7], filt), b3 = __fsub_rn(d[8], filt);
            // XZ weights (-c,0,c,-e | e,-c,0,c); YZ weights (-c,-e,-c,0 | 0,c,e,c).  A product with a zero
            // weight is +-0 and leaves its partner unchanged, so those adds are dropped.
            float xz0 = __fadd_rn(__fmul_rn(a0, -c), __fmul_rn(b0, e));
            float xz1 = __fmul_rn(b1, -c);
            float xz2 = __fmul_rn(a2, c);
            float xz3 = __fadd_rn(__fmul_rn(a3, -e), __fmul_rn(b3, c));
            float yz0 = __fmul_rn(a0, -c);
            float yz1 = __fadd_rn(__fmul_rn(a1, -e), __fmul_rn(b1, c));
            float yz2 = __fadd_rn(__fmul_rn(a2, -c), __fmul_rn(b2, e));
            float yz3 = __fmul_rn(b3, c);
            float xz = __fadd_rn(__fadd_rn(__fadd_rn(xz0, xz1), xz2), xz3);
            float yz = __fadd_rn(__fadd_rn(__fadd_rn(yz0, yz1), yz2), yz3);
            float vx = -xz, vy = yz;
            float len2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), nz2);  // (+ w*w = +0 dropped)
            float inv = rsqrtss_x86(P.rsqrtTab, len2);
            nx = vx * inv; ny = vy * inv; nzc = 0.382658847856f * inv;
            // variance (legacypbr.d:356-377): integer-valued floats, exact in any order
            float ddx = (zm[0] + z0[0] + zp[0]) - 2.0f * (zm[1] + z0[1] + zp[1]) + (zm[2] + z0[2] + zp[2]);
            float ddy = (zm[0] + zm[1] + zm[2]) - 2.0f * (z0[0] + z0[1] + z0[2]) + (zp[0] + zp[1] + zp[2]);
            ddx *= (1 / 256.0f);
            ddy *= (1 / 256.0f);
            variance = fmaf(ddx, ddx, ddy * ddy);
        }
        const float zc = z0[1];
        F3 accum{0, 0, 0};

        // ---- PassAmbientOcclusion (legacypbr.d:400-444)
        if (mask & 2u) {
            float s = sampleRollZ(rZ1, P.depth[lvZ[1]], xZ1, yLinear(lvZ[1], j, P.depth[lvZ[1]].h));
            s += sampleRollZ(rZ2, P.depth[lvZ[2]], xZ2, yLinear(lvZ[2], j, P.depth[lvZ[2]].h));
            s += sampleRollZ(rZ3, P.depth[lvZ[3]], xZ3, yLinear(lvZ[3], j, P.depth[lvZ[3]].h));
            s += sampleRollZ(rZ4, P.depth[lvZ[4]], xZ4, yLinear(lvZ[4], j, P.depth[lvZ[4]].h));
            float diff = fmaf(s, -0.25f, zc);
            float cavity = fmaSat(diff, 1.0f / 23040, 1.0f);
            accum = base * (cavity * P.ambient);
        }

        // ---- PassObliqueShadowLight (legacypbr.d:460-616): clamp(1 + (z + s - d)/15360, 0, 1) per tap
        if (mask & 4u) {
            float lightL = 0.f, lightR = 0.f;
#pragma unroll
            for (int s = 1; s <= 10; ++s) {
                // columns are clamped in the staged window only at the image border; inside the image the
                // window is wide enough (HALO_X), at the border the reference clamps too (legacypbr.d:532-534)
                const int x1 = min(ic + s, W - 1) - cx0, x2 = max(ic - s, 0) - cx0;
                const float ks = fmaf(zc + (float)s, kA, 1.0f);
                const float* row = tile[r - s];
                float c1 = fmaSat(row[x1], -kA, ks);
                float c2 = fmaSat(row[x2], -kA, ks);
                lightR = fmaf(c1, P.shadowW[s], lightR);
                lightL = fmaf(c2, P.shadowW[s], lightL);
            }
            float k = fmaf(lightL, 0.7f, lightR) * P.shadowInvTotal;
            accum.x = fmaf(base.x * P.light1[0], k, accum.x);
            accum.y = fmaf(base.y * P.light1[1], k, accum.y);
            accum.z = fmaf(base.z * P.light1[2], k, accum.z);
        }

        // ---- PassDirectionalLight (legacypbr.d:636-666)
        if (mask & 8u) {
            float dot = fmaf(nzc, P.light2Dir[2], fmaf(ny, P.light2Dir[1], nx * P.light2Dir[0]));
            float f = fmaf(0.5f, dot, 0.5f);
            float a = fmaf(rough, -0.5f, 0.24f);
            f = (f - a) * fastRcp(1.0f - a);  // linmap(f, a, 1, 0, 1), core/dplug/core/math.d:25-28
            accum.x = fmaf(base.x * P.light2Col[0], f, accum.x);
            accum.y = fmaf(base.y * P.light2Col[1], f, accum.y);
            accum.z = fmaf(base.z * P.light2Col[2], f, accum.z);
        }

        // toEye (legacypbr.d:756-757, 914-915): exact operand for RSQRTSS, shared by specular and skybox
        float ex = eyeX, ey = 0, ez = 1.0f;
        if (mask & 0x30u) {
            ey = __fsub_rn(__fmul_rn((float)j, P.invH), 0.5f);
            float len2 = __fadd_rn(__fadd_rn(eyeX2, __fmul_rn(ey, ey)), 1.0f);
            float inv = rsqrtss_x86(P.rsqrtTab, len2);
            ex = __fmul_rn(eyeX, inv); ey = __fmul_rn(ey, inv); ez = inv;  // 1.0f * inv
        }

        // ---- PassSpecularLight (legacypbr.d:720-813)
        if (mask & 0x10u) {
            float hx = __fsub_rn(ex, -P.light3Dir[0]), hy = __fsub_rn(ey, -P.light3Dir[1]), hz = __fsub_rn(ez, -P.light3Dir[2]);
            float inv = rsqrtss_x86(P.rsqrtTab, rsqrtArgExact(hx, hy, hz));
            float sf = fmaf(hz * inv, nzc, fmaf(hy * inv, ny, (hx * inv) * nx));
            sf = fmaxf(sf, 1e-3f);
            float exponent = __ldg(P.expTab + roughByte);
            float Ft = fastRcp(fmaf(exponent * variance, 4e-5f, 1.0f));
            float eF = exponent * Ft;
            float tok = (1.0f + eF) * fastRcp(1.0f + exponent);
            float p = fastPow(sf, eF);
            float roughFactor = 10.0f * (1.0f - rough) * fmaf(metal, -0.5f, 1.0f);
            float k = p * roughFactor * tok * specAmt;
            accum.x = fmaf(base.x * P.light3Col[0], k, accum.x);
            accum.y = fmaf(base.y * P.light3Col[1], k, accum.y);
            accum.z = fmaf(base.z * P.light3Col[2], k, accum.z);
        }

        // ---- PassSkyboxReflections (legacypbr.d:872-930)
        if (HAS_SKY && (mask & 0x20u)) {
            float lod = variance * P.skyPixels;
            lod = fmaf(0.5f, fastlog2(fmaf(lod, 0.00001f, 1.0f)), (6.0f / 255.0f) * (float)roughByte);
            float sum = 2.0f * fmaf(ez, nzc, fmaf(ey, ny, ex * nx));  // _mm_reflectnormal_ps, ransac.d:31-36
            float rx = fmaf(-sum, nx, ex), ry = fmaf(-sum, ny, ey);
            float skyx = fmaf(fmaf(rx, -0.5f, 0.5f), P.skyFX, 0.5f);
            float skyy = fmaf(fmaf(ry, 0.5f, 0.5f), P.skyFY, 0.5f);
            int il = __float2int_rz(lod);               // linearMipmapSample, mipmap.d:149-160
            float fl = lod - (float)il;
            F3 sky = skySample(P, il, skyx, skyy);
            if (fl != 0.f) {
                F3 sky1 = skySample(P, il + 1, skyx, skyy);
                sky = fma3(fl, sky1 - sky, sky);
            }
            float k = metal * (P.skyAmount * div255);
            accum.x = fmaf(base.x * sky.x, k, accum.x);
            accum.y = fmaf(base.y * sky.y, k, accum.y);
            accum.z = fmaf(base.z * sky.z, k, accum.z);
        }

        // ---- PassEmissiveContribution (legacypbr.d:948-1007, non-legacy branch)
        if (mask & 0x40u) {
            F3 s = sampleRollRGB(rD1, P.diffuse[lvD[1]], xD1, yLinear(lvD[1], j, P.diffuse[lvD[1]].h));
            s = s + sampleRollRGB(rD2, P.diffuse[lvD[2]], xD2, yLinear(lvD[2], j, P.diffuse[lvD[2]].h));
            s = s + sampleRollRGB(rD3, P.diffuse[lvD[3]], xD3, yLinear(lvD[3], j, P.diffuse[lvD[3]].h));
            s = s + sampleRollCub(rD4, P.diffuse[lvD[4]], xD4, lvD[4], j);
            F3 c5 = sampleRollCub(rD5, P.diffuse[lvD[5]], xD5, lvD[5], j);
            float noise = (u16ToFloat(kBlueNoise16x16[noiseCol + (j & 15)]) - 127.5f) * 0.003f;
            const float AMT = 0.002f * 0.67f;
            float k5 = AMT * (1.0f + noise);
            accum.x = fmaf(c5.x, k5, fmaf(s.x, AMT, accum.x));
            accum.y = fmaf(c5.y, k5, fmaf(s.y, AMT, accum.y));
            accum.z = fmaf(c5.z, k5, fmaf(s.z, AMT, accum.z));
        }

        // ---- PassClampAndConvertTo8bit (legacypbr.d:1057-1107, non-legacy branch)
        if ((mask & 0x80u) && active) {
            float ex0 = fmaxf(0.0f, accum.x - P.toneThr), ex1 = fmaxf(0.0f, accum.y - P.toneThr), ex2 = fmaxf(0.0f, accum.z - P.toneThr);
            float luma = fmaf(0.072187f, ex2, fmaf(0.715158f, ex1, 0.212655f * ex0));
            float add = luma * (P.toneRatio * (1.0f / 3.0f));
            uint32_t out = cvtSatU8((accum.x + add) * 255.99f) | (cvtSatU8((accum.y + add) * 255.99f) << 8) |
                           (cvtSatU8((accum.z + add) * 255.99f) << 16) | 0xff000000u;
            rowPtrW<uint32_t>(P.out, j)[i] = out;
        }

        // slide the depth window down
#pragma unroll
        for (int k = 0; k < 3; ++k) { zm[k] = z0[k]; z0[k] = zp[k]; }
    }
}

int pbr_fast_strip_rows() { return STRIP_ROWS; }

void launch_pbr_fast(const PbrArgs& args, cudaStream_t stream) {
    const int grid = (args.nBoxes + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    if (args.nSky > 0) k_pbr_fast<true><<<grid, WARPS_PER_CTA * 32, 0, stream>>>(args);
    else k_pbr_fast<false><<<grid, WARPS_PER_CTA * 32, 0, stream>>>(args);
}

}  // namespace b2
