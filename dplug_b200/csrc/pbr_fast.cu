// Lighting kernel, throughput variant (the one compositeTile launches by default).
//
// Same eight passes as gui/dplug/gui/legacypbr.d:269-1108, fused, but organised for the SM instead of for
// a 64x64 CPU tile:
//
//   * work item = one WARP = a strip of 32 adjacent columns x up to 32 rows of one GUI tile; a lane owns one
//     column and walks down it, so per-row quantities are warp-uniform and per-column ones are computed once;
//   * everything a strip samples besides its own pixels is staged ONCE per warp in shared memory, with all
//     global loads issued back to back (one exposed latency per strip instead of several per row):
//       - level-0 depth, +10 columns left/right, 10 rows up, 1 down (reach of the oblique shadow,
//         legacypbr.d:525-534, and of the 3x3 normal stencil, legacypbr.d:301-314), as float, with
//         clamp-to-edge addressing (= the reference's replicated border, graphics.d:1121-1126);
//       - the texel windows of diffuse levels 1..5 and depth levels 1..4 the strip's samplers touch;
//       - a per-row table of the vertical sampler set-up (texel rows, fractions, Catmull-Rom weights);
//   * the mip samplers (Mipmap.linearSample / cubicSample, graphics/dplug/graphics/mipmap.d:167-496) are
//     evaluated separably: the horizontal interpolation of a texel row is done once per lane and reused for
//     the 2^level rows that share it ("rolling rows" in registers);
//   * this row's diffuse / material pixels are prefetched while the previous row is shaded;
//   * 8/16-bit -> float conversions use byte-permute + magic-number subtraction (FMA pipe) instead of I2F
//     (quarter-rate XU pipe); pow / reciprocal use MUFU.
//
// Numerical contract: +-1 LSB per 8-bit channel against the reference arithmetic (BASELINE.json).  The
// three RSQRTSS operands per pixel are computed with the reference's exact IEEE sequence (no FMA, same
// order) because the table-driven RSQRTSS is discontinuous and Blinn-Phong exponents up to 548 amplify
// it; everything downstream may use FMA / MUFU (relative error ~1e-6).
//
// Precondition (checked by the host): full pyramids, i.e. diffuse has levels 0..5 and depth 0..4 (every canvas
// of at least 32x32 pixels).  Smaller canvases take the exact kernel.
#include "pbr_math.cuh"

namespace b2 {

constexpr int STRIP_ROWS = 32;
constexpr int HALO_X = 10;                     // shadow reach left/right
constexpr int HALO_UP = 10;                    // shadow reach upwards
constexpr int TILE_COLS = 32 + 2 * HALO_X + 2; // +2: the staged window starts at an even column
constexpr int TILE_ROWS = STRIP_ROWS + HALO_UP + 1;
constexpr int WARPS_PER_CTA = 4;
#ifndef B2_MIN_CTAS
#define B2_MIN_CTAS 4
#endif

// window extents: a bilinear sampler of level l touches at most (32 >> l) + 2 texels along 32 pixels, a
// bicubic one (32 >> l) + 4
constexpr int W1 = 18, W2 = 10, W3 = 6, WZ4 = 4, WC4 = 6, WC5 = 5;

struct RowSetup {            // vertical sampler set-up of one row; texel rows are relative to the staged windows
    int iy[4], iy1[4];       // levels 1, 2, 3 (diffuse and depth share them) and depth level 4
    float fy[4];
    int ib4, ib5;            // bicubic interval (floor(y_s) + 1) of diffuse levels 4 and 5
    float wy4[4], wy5[4];    // Catmull-Rom weights of the four rows
    float ey, ey2;           // toEye.y = j / H - 0.5 and its square (legacypbr.d:756, 914), exact
    // which rolling rows change at this row (compared with the row above; everything on a strip's first row):
    // bits 0-2 bilinear levels 1-3, bit 3 depth level 4, bits 4-5 / 6-7 bicubic levels 4 / 5 (1 = slide by one, 2 = reload)
    int refresh;
    int pad[3];
};
static_assert(sizeof(RowSetup) % 16 == 0, "RowSetup is read with 128-bit shared loads");

struct alignas(16) WarpSmem {
    RowSetup rows[STRIP_ROWS];
    uint16_t depth[TILE_ROWS][TILE_COLS];   // raw L16; converted on use (keeps 4 CTAs / 16 warps per SM resident)
    uint32_t d1[W1][W1], d2[W2][W2], d3[W3][W3], d4[WC4][WC4], d5[WC5][WC5];
    float z1[W1][W1], z2[W2][W2], z3[W3][W3], z4[WZ4][WZ4];   // depth levels 1..4, converted once when staged
    int org[12];   // window origins ox1, oy1, ox2, oy2, ox3, oy3, oxz4, oyz4, ox4, oy4, ox5, oy5 (re-read by the refresh blocks)
    int pad[4];
};
static_assert(sizeof(WarpSmem) % 16 == 0, "per-warp shared block keeps 16-byte alignment");

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return F3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return F3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return F3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ F3 fma3(float s, F3 a, F3 b) { return F3{fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)}; }

// byte k of a packed RGBA8 texel -> float, exact, without I2F: (0x4B000000 | b) is the float 2^23 + b
__device__ __forceinline__ float byteToFloat(uint32_t p, int k) {
    return __uint_as_float(__byte_perm(p, 0x4B000000u, 0x7440 + k)) - 8388608.0f;
}
__device__ __forceinline__ float byteBiased23(uint32_t p, int k) { return __uint_as_float(__byte_perm(p, 0x4B000000u, 0x7440 + k)); }   // 2^23 + byte
__device__ __forceinline__ F3 rgbOf(uint32_t p) { return F3{byteToFloat(p, 0), byteToFloat(p, 1), byteToFloat(p, 2)}; }
// byte k -> float 32768 + b with a single PRMT (b lands in mantissa bits 8..15 of 2^15).  Used where a weighted
// sum with weights adding up to 1 follows: the 32768 is subtracted once from the result (absolute error of the
// intermediate sums <= 2^-8 of a 0..255 value, i.e. < 0.005 LSB after the sky / bloom scale factors)
__device__ __forceinline__ float byteToFloatBiased(uint32_t p, int k) {
    return __uint_as_float(__byte_perm(p, 0x47000000u, 0x7604 | (k << 4)));
}
__device__ __forceinline__ F3 rgbBiased(uint32_t p) { return F3{byteToFloatBiased(p, 0), byteToFloatBiased(p, 1), byteToFloatBiased(p, 2)}; }
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

// RSQRTSS emulation for operands that are provably positive normal floats (all three uses here): the table entry of
// the bin with the exponent moved by floor((e - 126) / 2) = (bits >> 24) - 63, i.e. entry + (63 << 23) - ((bits >> 1)
// & 0x7f800000).  Same value as rsqrtss_x86() on that domain; a zero operand yields +inf like the instruction.
template <bool MAY_BE_ZERO = true>
__device__ __forceinline__ float rsqrtssPos(const uint32_t* __restrict__ tab, float x) {
    const uint32_t b = __float_as_uint(x);
    const uint32_t r = __ldg(tab + ((b >> 13) & 0x7ffu)) + 0x1F800000u - ((b >> 1) & 0x7F800000u);
    if (!MAY_BE_ZERO) return __uint_as_float(r);   // operand bounded away from zero by construction
    return b == 0u ? __uint_as_float(0x7f800000u) : __uint_as_float(r);
}

// ---- sampler coordinates.  (p + 0.5) * 2^-l - 0.5 == n / 2^(l+1) with n = 2p + 1 - 2^l: integer shifts give
// the texel index and the fraction exactly as the reference's float arithmetic does (which is exact here).
struct Lin { int i0, i1; float f; };
__device__ __forceinline__ Lin linCoord(int l, int p, int size) {  // Mipmap.linearSample, mipmap.d:309-339
    int n = max(2 * p + 1 - (1 << l), 0);  // negative coordinate -> 0
    Lin s;
    int i = n >> (l + 1);
    s.f = (float)(n & ((2 << l) - 1)) * levelScale(l + 1);  // fraction taken before the upper clamp
    s.i0 = min(i, size - 1);
    s.i1 = min(s.i0 + 1, size - 1);
    return s;
}
__device__ __forceinline__ void catmullRomWeights(float t, float w[4]) {
    // cubicInterp (mipmap.d:263-270) regrouped per sample: p0*w0 + p1*w1 + p2*w2 + p3*w3
    float t2 = t * t, t3 = t2 * t;
    w[0] = -0.5f * t + t2 - 0.5f * t3;
    w[1] = 1.0f - 2.5f * t2 + 1.5f * t3;
    w[2] = 0.5f * t + 2.0f * t2 - 1.5f * t3;
    w[3] = -0.5f * t2 + 0.5f * t3;
}
// Mipmap.cubicSample, mipmap.d:182-204: y_s + 1 = (n + 2^(l+1)) / 2^(l+1) is always positive, so the truncations
// are floors; ib = floor(y_s) + 1; the four texels are clamp(ib - 2 + k)
__device__ __forceinline__ void cubCoord(int l, int p, int& ib, float& t) {
    const int n1 = 2 * p + 1 - (1 << l) + (2 << l);
    ib = n1 >> (l + 1);
    t = (float)(n1 & ((2 << l) - 1)) * levelScale(l + 1);
}

// copies the window [y0, y0 + WH) x [x0, x0 + WW) of a level into shared memory, clamped to the level
template <typename T, int WW, int WH, typename D>
__device__ __forceinline__ void stageWindow(D (*dst)[WW], const DevLevel& L, int x0, int y0, int lane) {
#pragma unroll
    for (int idx = lane; idx < WW * WH; idx += 32) {
        const int r = idx / WW, c = idx - r * WW;
        const int gy = max(0, min(y0 + r, L.h - 1)), gx = max(0, min(x0 + c, L.w - 1));
        const T v = __ldg(rowPtr<T>(L, gy) + gx);
        if (sizeof(D) == sizeof(T)) dst[r][c] = (D)v;
        else dst[r][c] = (D)u16ToFloat((uint32_t)v);   // L16 windows are kept as floats (exact)
    }
}

// ALL_PASSES: every pass active (the reference's default; CompositorPass.setActive is a debugging aid,
// compositor.d:222-236) — the per-pass tests then fold away at compile time.
template <bool HAS_SKY, bool ALL_PASSES>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, B2_MIN_CTAS) k_pbr_fast(const __grid_constant__ PbrArgs P) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x * WARPS_PER_CTA + warp;
    if (item >= P.nBoxes) return;
    WarpSmem& S = reinterpret_cast<WarpSmem*>(smemRaw)[warp];
    const Box box = P.boxes[item];  // one strip: at most 32 columns x 32 rows
    const int W = P.W, H = P.H;
    const DevLevel& D0 = P.depth[0];
    const int xLast = box.maxx - 1, yLast = box.maxy - 1;

    // ---------------------------------------------------------------- prefetch the first row's pixels
    const int i = box.minx + lane;
    const bool active = i < box.maxx;
    const int ic = min(i, xLast);  // inactive lanes shadow the last column (no divergence, no stores)
    uint32_t diffuseNext = __ldg(rowPtr<uint32_t>(P.diffuse[0], box.miny) + ic);
    uint32_t materialNext = __ldg(rowPtr<uint32_t>(P.material, box.miny) + ic);

    // ---------------------------------------------------------------- stage the mip windows
    // window origins: first texel the first column / row of the strip touches
    const int ox1 = linCoord(1, box.minx, P.diffuse[1].w).i0, oy1 = linCoord(1, box.miny, P.diffuse[1].h).i0;
    const int ox2 = linCoord(2, box.minx, P.diffuse[2].w).i0, oy2 = linCoord(2, box.miny, P.diffuse[2].h).i0;
    const int ox3 = linCoord(3, box.minx, P.diffuse[3].w).i0, oy3 = linCoord(3, box.miny, P.diffuse[3].h).i0;
    const int oxz4 = linCoord(4, box.minx, P.depth[4].w).i0, oyz4 = linCoord(4, box.miny, P.depth[4].h).i0;
    int ibx, iby; float tt;
    cubCoord(4, box.minx, ibx, tt); cubCoord(4, box.miny, iby, tt);
    const int ox4 = max(0, min(ibx - 2, P.diffuse[4].w - 1)), oy4 = max(0, min(iby - 2, P.diffuse[4].h - 1));
    cubCoord(5, box.minx, ibx, tt); cubCoord(5, box.miny, iby, tt);
    const int ox5 = max(0, min(ibx - 2, P.diffuse[5].w - 1)), oy5 = max(0, min(iby - 2, P.diffuse[5].h - 1));

    if (lane == 0) {
        const int o[12] = {ox1, oy1, ox2, oy2, ox3, oy3, oxz4, oyz4, ox4, oy4, ox5, oy5};
#pragma unroll
        for (int k = 0; k < 12; ++k) S.org[k] = o[k];
    }
    stageWindow<uint32_t, W1, W1>(S.d1, P.diffuse[1], ox1, oy1, lane);
    stageWindow<uint16_t, W1, W1>(S.z1, P.depth[1], ox1, oy1, lane);
    stageWindow<uint32_t, W2, W2>(S.d2, P.diffuse[2], ox2, oy2, lane);
    stageWindow<uint16_t, W2, W2>(S.z2, P.depth[2], ox2, oy2, lane);
    stageWindow<uint32_t, W3, W3>(S.d3, P.diffuse[3], ox3, oy3, lane);
    stageWindow<uint16_t, W3, W3>(S.z3, P.depth[3], ox3, oy3, lane);
    stageWindow<uint32_t, WC4, WC4>(S.d4, P.diffuse[4], ox4, oy4, lane);
    stageWindow<uint16_t, WZ4, WZ4>(S.z4, P.depth[4], oxz4, oyz4, lane);
    stageWindow<uint32_t, WC5, WC5>(S.d5, P.diffuse[5], ox5, oy5, lane);

    // ---------------------------------------------------------------- stage the depth window
    // columns [cx0, cx0 + TILE_COLS), rows [box.miny - 10, box.maxy + 1), clamped to the image
    const int cx0 = (box.minx - HALO_X) & ~1;
    const int ry0 = box.miny - HALO_UP;
    const int nRows = (box.maxy - box.miny) + HALO_UP + 1;
    if (2 * lane < TILE_COLS) {
        const int c = 2 * lane;  // each lane copies a pair of texels per row
        const int gx = cx0 + c;
        const bool pairOk = gx >= 0 && gx + 1 < W;
        const int gxa = max(0, min(gx, W - 1)), gxb = max(0, min(gx + 1, W - 1));
        // two loops rather than a per-row choice (which the compiler turns into BOTH loads, predicated): the clamped
        // texel-by-texel path is taken only by the lanes hanging over the left / right image edge
        if (pairOk) {
#pragma unroll 4
            for (int r = 0; r < nRows; ++r) {
                const int gy = max(0, min(ry0 + r, H - 1));
                *reinterpret_cast<uint32_t*>(&S.depth[r][c]) = __ldg(reinterpret_cast<const uint32_t*>(rowPtr<uint16_t>(D0, gy) + gx));
            }
        } else {
            for (int r = 0; r < nRows; ++r) {
                const int gy = max(0, min(ry0 + r, H - 1));
                const uint16_t* src = rowPtr<uint16_t>(D0, gy);
                *reinterpret_cast<uint32_t*>(&S.depth[r][c]) = (uint32_t)__ldg(src + gxa) | ((uint32_t)__ldg(src + gxb) << 16);
            }
        }
    }

    // ---------------------------------------------------------------- per-row table: lane k prepares row miny + k
    if (box.miny + lane <= yLast) {
        const int j = box.miny + lane;
        RowSetup rs;
        const int oy[4] = {oy1, oy2, oy3, oyz4};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            Lin y = linCoord(k + 1, j, k < 3 ? P.diffuse[k + 1].h : P.depth[4].h);
            rs.iy[k] = y.i0 - oy[k]; rs.iy1[k] = y.i1 - oy[k]; rs.fy[k] = y.f;
        }
        float t;
        cubCoord(4, j, rs.ib4, t); catmullRomWeights(t, rs.wy4);
        cubCoord(5, j, rs.ib5, t); catmullRomWeights(t, rs.wy5);
        rs.ey = __fsub_rn(__fmul_rn((float)j, P.invH), 0.5f);
        rs.ey2 = __fmul_rn(rs.ey, rs.ey);
        int refresh = 0xaf;                        // first row of the strip: everything is (re)loaded
        if (lane > 0) {
            refresh = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                Lin y = linCoord(k + 1, j - 1, k < 3 ? P.diffuse[k + 1].h : P.depth[4].h);
                if (y.i0 - oy[k] != rs.iy[k] || y.i1 - oy[k] != rs.iy1[k]) refresh |= 1 << k;
            }
            int ibp; float tp;
            cubCoord(4, j - 1, ibp, tp);
            if (rs.ib4 != ibp) refresh |= (rs.ib4 == ibp + 1 ? 1 : 2) << 4;
            cubCoord(5, j - 1, ibp, tp);
            if (rs.ib5 != ibp) refresh |= (rs.ib5 == ibp + 1 ? 1 : 2) << 6;
        }
        rs.refresh = refresh;
        rs.pad[0] = rs.pad[1] = rs.pad[2] = 0;
        S.rows[lane] = rs;
    }
    __syncwarp();

    // ---------------------------------------------------------------- per-column sampler set-up
    // The horizontal set-up of a level (texel columns, fraction, Catmull-Rom weights) is needed only when one of its
    // rolling rows is refreshed — every 2^level rows.  It is RECOMPUTED there from the lane's column (a dozen integer
    // operations) instead of living in ~32 registers for the whole strip: the kernel is register-bound (every register
    // saved is occupancy or a spill avoided), and the refresh blocks are off the per-pixel path.  `opaque` keeps the
    // compiler from hoisting the recomputation back out of the row loop.
    auto opaque = [](int v) { asm volatile("" : "+r"(v)); return v; };
    const volatile int* org = S.org;
    // bilinear level l (1..3 diffuse + depth, 4 depth only): refresh (a, b - a) of the colour and / or the depth row pair
    auto refreshLin = [&](int l, int lw, auto& win, auto& zwin, int oIdx, int rA, int rB, bool doD, bool doZ, F3& a, F3& d, float& za, float& zd) {
        const Lin c = linCoord(l, opaque(ic), lw);
        const int ox = org[oIdx], xa = c.i0 - ox, xb = c.i1 - ox;
        if (doD) {
            // B - A is taken on the 2^23-biased values (the biases cancel exactly), only A is un-biased
            const uint32_t tA0 = win[rA][xa], tB0 = win[rA][xb], tA1 = win[rB][xa], tB1 = win[rB][xb];
            const F3 A0 = rgbOf(tA0), A1 = rgbOf(tA1);
            const F3 dB0{byteBiased23(tB0, 0) - byteBiased23(tA0, 0), byteBiased23(tB0, 1) - byteBiased23(tA0, 1), byteBiased23(tB0, 2) - byteBiased23(tA0, 2)};
            const F3 dB1{byteBiased23(tB1, 0) - byteBiased23(tA1, 0), byteBiased23(tB1, 1) - byteBiased23(tA1, 1), byteBiased23(tB1, 2) - byteBiased23(tA1, 2)};
            const F3 na = fma3(c.f, dB0, A0), nb = fma3(c.f, dB1, A1);
            a = na; d = nb - na;
        }
        if (doZ) {
            const float A0 = zwin[rA][xa], B0 = zwin[rA][xb], A1 = zwin[rB][xa], B1 = zwin[rB][xb];
            const float na = fmaf(c.f, B0 - A0, A0), nb = fmaf(c.f, B1 - A1, A1);
            za = na; zd = nb - na;
        }
    };
    // bicubic level (4 or 5): slide the four horizontally interpolated rows by one, or reload all of them
    auto refreshCub = [&](int l, const DevLevel& L, auto& win, int oIdx, int ib, bool slide, F3 (&p)[4]) {
        int ibx; float t, w[4];
        cubCoord(l, opaque(ic), ibx, t);
        catmullRomWeights(t, w);
        const int ox = org[oIdx], oy = org[oIdx + 1];
        int xc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xc[k] = max(0, min(ibx - 2 + k, L.w - 1)) - ox;
        auto row = [&](int k) {
            const int r = max(0, min(ib - 2 + k, L.h - 1)) - oy;
            F3 v = rgbOf(win[r][xc[0]]) * w[0];
            v = fma3(w[1], rgbOf(win[r][xc[1]]), v);
            v = fma3(w[2], rgbOf(win[r][xc[2]]), v);
            return fma3(w[3], rgbOf(win[r][xc[3]]), v);
        };
        if (slide) { p[0] = p[1]; p[1] = p[2]; p[2] = p[3]; p[3] = row(3); }
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k] = row(k);
        }
    };

    // rolling state: bilinear levels keep (row A, row B - row A); bicubic levels keep their four rows
    F3 a1{0, 0, 0}, d1 = a1, a2 = a1, d2 = a1, a3 = a1, d3 = a1;
    float za1 = 0, zd1 = 0, za2 = 0, zd2 = 0, za3 = 0, zd3 = 0, za4 = 0, zd4 = 0;
    F3 p4[4], p5[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) p4[k] = p5[k] = a1;

    const uint32_t mask = ALL_PASSES ? 0xffu : P.passMask;
    const bool needMipsD = (mask & 0x40u) != 0, needMipsZ = (mask & 2u) != 0;

    // toEye.x and its square are column constants (legacypbr.d:756, 914) — exact sequence
    const float eyeX = __fsub_rn(0.5f, __fmul_rn((float)ic, P.invW));
    const float eyeX2 = __fmul_rn(eyeX, eyeX);
    const float nz2 = __fmul_rn(0.382658847856f, 0.382658847856f);
    const float multUshort = (float)(1.0 / 4655.0);
    const float kA = 0.00006510416f;  // 1/15360 as the reference's SIMD branch spells it (legacypbr.d:555)
    const unsigned char* noisePtr = kBlueNoise16x16 + (ic & 15) * 16;
    const float div255 = 1 / 255.0f;

    // sliding 3x3 window of d = depth * (1/FACTOR_Z) (legacypbr.d:304-314): rows j-1, j, j+1 at columns i-1, i, i+1
    const int tx = ic - cx0;  // column of this lane inside the staged window (>= HALO_X)
    float dm[3], d0[3], dp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        dm[k] = __fmul_rn(u16ToFloat(S.depth[HALO_UP - 1][tx - 1 + k]), multUshort);
        d0[k] = __fmul_rn(u16ToFloat(S.depth[HALO_UP][tx - 1 + k]), multUshort);
    }
    // shadow taps read the raw u16 as the float 2^23 + d (one LOP3, no subtraction): the bias is folded into ks
    const float kBias = 8388608.0f * kA;

    for (int j = box.miny; j <= yLast; ++j) {
        const uint16_t* trow = &S.depth[j - ry0][tx];   // this lane's texel of row j inside the staged window
#pragma unroll
        for (int k = 0; k < 3; ++k) dp[k] = __fmul_rn(u16ToFloat(trow[TILE_COLS - 1 + k]), multUshort);
        const uint32_t diffusePx = diffuseNext, materialPx = materialNext;
        {   // prefetch the next row's pixels (clamped on the last row: harmless re-read)
            const int jn = min(j + 1, yLast);
            diffuseNext = __ldg(rowPtr<uint32_t>(P.diffuse[0], jn) + ic);
            materialNext = __ldg(rowPtr<uint32_t>(P.material, jn) + ic);
        }
        const RowSetup& rs = S.rows[j - box.miny];   // broadcast shared loads

        // rolling rows of levels 1..3: the diffuse and depth pyramids share the texel rows, one warp-uniform test
        // per level refreshes both (the rows are recomputed from the staged windows, so the result does not depend
        // on where the strip starts)
        const int refresh = rs.refresh;
        if (refresh & 1) refreshLin(1, P.diffuse[1].w, S.d1, S.z1, 0, rs.iy[0], rs.iy1[0], needMipsD, needMipsZ, a1, d1, za1, zd1);
        if (refresh & 2) refreshLin(2, P.diffuse[2].w, S.d2, S.z2, 2, rs.iy[1], rs.iy1[1], needMipsD, needMipsZ, a2, d2, za2, zd2);
        if (refresh & 4) refreshLin(3, P.diffuse[3].w, S.d3, S.z3, 4, rs.iy[2], rs.iy1[2], needMipsD, needMipsZ, a3, d3, za3, zd3);

        const F3 base = rgbOf(diffusePx) * div255;
        const float rough = byteToFloat(materialPx, 0) * div255;
        const float metal = byteToFloat(materialPx, 1) * div255;
        const float specAmt = byteToFloat(materialPx, 2) * div255;
        const int roughByte = materialPx & 0xff;

        // ---- PassComputeNormal (legacypbr.d:279-381) + computePlaneFittingNormal (ransac.d:53-121).
        // Exact reference sequence: the result feeds RSQRTSS.
        float nx = 0, ny = 0, nzc = 0, variance = 0;
        if (mask & 1u) {
            const float d[9] = {dm[0], dm[1], dm[2], d0[0], d0[1], d0[2], dp[0], dp[1], dp[2]};
            const float c = 0.03660694846f, e = 0.118115527008f;
            float m0 = __fadd_rn(__fmul_rn(d[0], c), __fmul_rn(d[5], e));
            float m1 = __fadd_rn(__fmul_rn(d[1], e), __fmul_rn(d[6], c));
            float m2 = __fadd_rn(__fmul_rn(d[2], c), __fmul_rn(d[7], e));
            float m3 = __fadd_rn(__fmul_rn(d[3], e), __fmul_rn(d[8], c));
            float filt = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[4], 0.38111009811f), m0), m1), m2), m3);
            float a0 = __fsub_rn(d[0], filt), a1_ = __fsub_rn(d[1], filt), a2_ = __fsub_rn(d[2], filt), a3_ = __fsub_rn(d[3], filt);
            float b0 = __fsub_rn(d[5], filt), b1 = __fsub_rn(d[6], filt), b2 = __fsub_rn(d[7], filt), b3 = __fsub_rn(d[8], filt);
            // XZ weights (-c,0,c,-e | e,-c,0,c); YZ weights (-c,-e,-c,0 | 0,c,e,c).  A product with a zero
            // weight is +-0 and leaves its partner unchanged, so those adds are dropped.
            float xz0 = __fadd_rn(__fmul_rn(a0, -c), __fmul_rn(b0, e));
            float xz1 = __fmul_rn(b1, -c);
            float xz2 = __fmul_rn(a2_, c);
            float xz3 = __fadd_rn(__fmul_rn(a3_, -e), __fmul_rn(b3, c));
            float yz0 = __fmul_rn(a0, -c);
            float yz1 = __fadd_rn(__fmul_rn(a1_, -e), __fmul_rn(b1, c));
            float yz2 = __fadd_rn(__fmul_rn(a2_, -c), __fmul_rn(b2, e));
            float yz3 = __fmul_rn(b3, c);
            float xz = __fadd_rn(__fadd_rn(__fadd_rn(xz0, xz1), xz2), xz3);
            float yz = __fadd_rn(__fadd_rn(__fadd_rn(yz0, yz1), yz2), yz3);
            float vx = -xz, vy = yz;
            float len2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, vx), __fmul_rn(vy, vy)), nz2);  // (+ w*w = +0 dropped)
            float inv = rsqrtssPos<false>(P.rsqrtTab, len2);   // len2 >= 0.3827^2
            nx = vx * inv; ny = vy * inv; nzc = 0.382658847856f * inv;
            // variance (legacypbr.d:356-377): second differences of the 3x3 window.  The reference sums raw depths
            // (exact); here they are rebuilt from d (relative error ~1e-7, the variance feeds no RSQRTSS)
            float ddx = (d[0] + d[3] + d[6]) - 2.0f * (d[1] + d[4] + d[7]) + (d[2] + d[5] + d[8]);
            float ddy = (d[0] + d[1] + d[2]) - 2.0f * (d[3] + d[4] + d[5]) + (d[6] + d[7] + d[8]);
            const float kVar = 4655.0f / 256.0f;
            ddx *= kVar;
            ddy *= kVar;
            variance = fmaf(ddx, ddx, ddy * ddy);
        }
        const float zc = u16ToFloat(trow[0]);
        F3 T{0, 0, 0};   // light arriving per channel; every pass but the emissive one is base colour x light

        // toEye (legacypbr.d:756-757, 914-915): exact operand for RSQRTSS, shared by specular and skybox
        float ex = eyeX, ey = 0, ez = 1.0f;
        if (mask & 0x30u) {
            ey = rs.ey;
            float len2 = __fadd_rn(__fadd_rn(eyeX2, rs.ey2), 1.0f);
            float inv = rsqrtssPos<false>(P.rsqrtTab, len2);   // len2 >= 1
            ex = __fmul_rn(eyeX, inv); ey = __fmul_rn(ey, inv); ez = inv;  // 1.0f * inv
        }

        // ---- PassSkyboxReflections (legacypbr.d:872-930), first half: the eight texel loads of the trilinear sample are the
        // only dependent global loads of a row (coordinates come from the normal), so they are issued HERE, before the
        // occlusion / shadow / directional / specular arithmetic, which then runs in their shadow; the blend follows
        // after the specular pass, where the reference adds the sky term (same order of additions into T)
        const bool doSky = HAS_SKY && (mask & 0x20u);
        uint32_t ta = 0, tb = 0, tc = 0, td = 0, ua = 0, ub = 0, uc = 0, ud = 0;
        float sfx = 0, sfy = 0, s2fx = 0, s2fy = 0, fl = 0;
        if (doSky) {
            float lod = variance * P.skyPixels;
            lod = fmaf(0.5f, fastlog2(fmaf(lod, 0.00001f, 1.0f)), (6.0f / 255.0f) * (float)roughByte);
            float sum = 2.0f * fmaf(ez, nzc, fmaf(ey, ny, ex * nx));  // _mm_reflectnormal_ps, ransac.d:31-36
            float rx = fmaf(-sum, nx, ex), ry = fmaf(-sum, ny, ey);
            float skyx = fmaf(fmaf(rx, -0.5f, 0.5f), P.skyFX, 0.5f);
            float skyy = fmaf(fmaf(ry, 0.5f, 0.5f), P.skyFY, 0.5f);
            int il = __float2int_rz(lod);               // linearMipmapSample, mipmap.d:149-160
            fl = lod - (float)il;
            const int level = max(0, min(il, P.nSky - 1));
            const DevLevel& L = P.sky[level];
            const Bilin s = bilinSetup(levelScale(level), skyx, skyy, L.w, L.h);
            const uint32_t* r0 = rowPtr<uint32_t>(L, s.iy);
            const uint32_t* r1 = rowPtr<uint32_t>(L, s.iy1);
            ta = __ldg(r0 + s.ix); tb = __ldg(r0 + s.ix1); tc = __ldg(r1 + s.ix); td = __ldg(r1 + s.ix1);
            sfx = s.fx; sfy = s.fy;
            if (fl != 0.f) {  // second level of the trilinear blend
                const int level2 = max(0, min(il + 1, P.nSky - 1));
                const DevLevel& L2 = P.sky[level2];
                const Bilin s2 = bilinSetup(levelScale(level2), skyx, skyy, L2.w, L2.h);
                const uint32_t* q0 = rowPtr<uint32_t>(L2, s2.iy);
                const uint32_t* q1 = rowPtr<uint32_t>(L2, s2.iy1);
                ua = __ldg(q0 + s2.ix); ub = __ldg(q0 + s2.ix1); uc = __ldg(q1 + s2.ix); ud = __ldg(q1 + s2.ix1);
                s2fx = s2.fx; s2fy = s2.fy;
            }
        }

        // ---- PassAmbientOcclusion (legacypbr.d:400-444): depth levels 1..4, bilinear
        if (needMipsZ) {
            if (refresh & 8) { F3 unusedA, unusedD; refreshLin(4, P.depth[4].w, S.z4, S.z4, 6, rs.iy[3], rs.iy1[3], false, true, unusedA, unusedD, za4, zd4); }
            float s = fmaf(rs.fy[0], zd1, za1) + fmaf(rs.fy[1], zd2, za2) + fmaf(rs.fy[2], zd3, za3) + fmaf(rs.fy[3], zd4, za4);
            float diff = fmaf(s, -0.25f, zc);
            float cavity = fmaSat(diff, 1.0f / 23040, 1.0f);
            const float amb = cavity * P.ambient;
            T = F3{amb, amb, amb};
        }

        // ---- PassObliqueShadowLight (legacypbr.d:460-616): clamp(1 + (z + s - d)/15360, 0, 1) per tap
        if (mask & 4u) {
            float lightL = 0.f, lightR = 0.f;
            const float k0 = fmaf(zc, kA, 1.0f) + kBias;
#pragma unroll
            for (int s = 1; s <= 10; ++s) {
                // the staged window is clamp-filled, so the taps (i +- s, j - s) (legacypbr.d:525-534, 568-576)
                // sit at compile-time offsets from this lane's texel
                const float ks = fmaf((float)s, kA, k0);
                float c1 = fmaSat(__uint_as_float(0x4B000000u | trow[-s * TILE_COLS + s]), -kA, ks);
                float c2 = fmaSat(__uint_as_float(0x4B000000u | trow[-s * TILE_COLS - s]), -kA, ks);
                lightR = fmaf(c1, P.shadowW[s], lightR);
                lightL = fmaf(c2, P.shadowW[s], lightL);
            }
            float k = fmaf(lightL, 0.7f, lightR) * P.shadowInvTotal;
            T.x = fmaf(P.light1[0], k, T.x);
            T.y = fmaf(P.light1[1], k, T.y);
            T.z = fmaf(P.light1[2], k, T.z);
        }

        // ---- PassDirectionalLight (legacypbr.d:636-666)
        if (mask & 8u) {
            float dot = fmaf(nzc, P.light2Dir[2], fmaf(ny, P.light2Dir[1], nx * P.light2Dir[0]));
            float f = fmaf(0.5f, dot, 0.5f);
            float a = fmaf(rough, -0.5f, 0.24f);
            f = (f - a) * fastRcp(1.0f - a);  // linmap(f, a, 1, 0, 1), core/dplug/core/math.d:25-28
            T.x = fmaf(P.light2Col[0], f, T.x);
            T.y = fmaf(P.light2Col[1], f, T.y);
            T.z = fmaf(P.light2Col[2], f, T.z);
        }

        // ---- PassSpecularLight (legacypbr.d:720-813)
        if (mask & 0x10u) {
            float hx = __fsub_rn(ex, -P.light3Dir[0]), hy = __fsub_rn(ey, -P.light3Dir[1]), hz = __fsub_rn(ez, -P.light3Dir[2]);
            float hl2 = __fadd_rn(__fadd_rn(__fmul_rn(hx, hx), __fmul_rn(hy, hy)), __fmul_rn(hz, hz));
            float inv = rsqrtssPos(P.rsqrtTab, hl2);
            float sf = fmaf(hz * inv, nzc, fmaf(hy * inv, ny, (hx * inv) * nx));
            sf = fmaxf(sf, 1e-3f);
            float exponent = __ldg(P.expTab + roughByte);
            float Ft = fastRcp(fmaf(exponent * variance, 4e-5f, 1.0f));
            float eF = exponent * Ft;
            float tok = (1.0f + eF) * fastRcp(1.0f + exponent);
            float p = fastPow(sf, eF);
            float roughFactor = 10.0f * (1.0f - rough) * fmaf(metal, -0.5f, 1.0f);
            float k = p * roughFactor * tok * specAmt;
            T.x = fmaf(P.light3Col[0], k, T.x);
            T.y = fmaf(P.light3Col[1], k, T.y);
            T.z = fmaf(P.light3Col[2], k, T.z);
        }

        // ---- PassSkyboxReflections, second half: blend the texels loaded above
        if (doSky) {
            F3 A = rgbBiased(ta), B = rgbBiased(tb), C = rgbBiased(tc), D = rgbBiased(td);
            F3 up = fma3(sfx, B - A, A), down = fma3(sfx, D - C, C);
            F3 sky = fma3(sfy, down - up, up);
            if (fl != 0.f) {
                F3 A2 = rgbBiased(ua), B2 = rgbBiased(ub), C2 = rgbBiased(uc), D2 = rgbBiased(ud);
                F3 up2 = fma3(s2fx, B2 - A2, A2), down2 = fma3(s2fx, D2 - C2, C2);
                F3 sky1 = fma3(s2fy, down2 - up2, up2);
                sky = fma3(fl, sky1 - sky, sky);
            }
            sky = F3{sky.x - 32768.0f, sky.y - 32768.0f, sky.z - 32768.0f};
            float k = metal * (P.skyAmount * div255);
            T.x = fmaf(sky.x, k, T.x);
            T.y = fmaf(sky.y, k, T.y);
            T.z = fmaf(sky.z, k, T.z);
        }

        F3 accum{base.x * T.x, base.y * T.y, base.z * T.z};

        // ---- PassEmissiveContribution (legacypbr.d:948-1007, non-legacy branch): diffuse levels 1..3 bilinear,
        // 4 and 5 bicubic
        if (needMipsD) {
            if (refresh & 0x30) refreshCub(4, P.diffuse[4], S.d4, 8, rs.ib4, (refresh & 0x10) != 0, p4);
            if (refresh & 0xc0) refreshCub(5, P.diffuse[5], S.d5, 10, rs.ib5, (refresh & 0x40) != 0, p5);
            F3 s = fma3(rs.fy[0], d1, a1) + fma3(rs.fy[1], d2, a2) + fma3(rs.fy[2], d3, a3);
            F3 c4 = fma3(rs.wy4[3], p4[3], fma3(rs.wy4[2], p4[2], fma3(rs.wy4[1], p4[1], p4[0] * rs.wy4[0])));
            F3 c5 = fma3(rs.wy5[3], p5[3], fma3(rs.wy5[2], p5[2], fma3(rs.wy5[1], p5[1], p5[0] * rs.wy5[0])));
            // clamp_0_to_65535 (mipmap.d:250-261); 8-bit sources cannot reach the upper bound
            s = s + F3{fmaxf(c4.x, 0.f), fmaxf(c4.y, 0.f), fmaxf(c4.z, 0.f)};
            c5 = F3{fmaxf(c5.x, 0.f), fmaxf(c5.y, 0.f), fmaxf(c5.z, 0.f)};
            float noise = (u16ToFloat(__ldg(noisePtr + (j & 15))) - 127.5f) * 0.003f;
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
        for (int k = 0; k < 3; ++k) { dm[k] = d0[k]; d0[k] = dp[k]; }
    }
}

int pbr_fast_strip_rows() { return STRIP_ROWS; }

static bool g_attrSet[64];
void launch_pbr_fast(const PbrArgs& args, cudaStream_t stream) {
    const int grid = (args.nBoxes + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const size_t smem = sizeof(WarpSmem) * WARPS_PER_CTA;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !g_attrSet[dev]) {  // > 48 KB of dynamic shared memory is an opt-in, per device
        cudaFuncSetAttribute(k_pbr_fast<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_pbr_fast<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_pbr_fast<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_pbr_fast<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        g_attrSet[dev] = true;
    }
    const bool all = (args.passMask & 0xffu) == 0xffu;
    const dim3 block(WARPS_PER_CTA * 32);
    if (args.nSky > 0) {
        if (all) k_pbr_fast<true, true><<<grid, block, smem, stream>>>(args);
        else k_pbr_fast<true, false><<<grid, block, smem, stream>>>(args);
    } else {
        if (all) k_pbr_fast<false, true><<<grid, block, smem, stream>>>(args);
        else k_pbr_fast<false, false><<<grid, block, smem, stream>>>(args);
    }
}

}  // namespace b2
