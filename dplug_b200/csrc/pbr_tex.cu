// Lighting kernel, throughput variant 2 (the one compositeTile launches by default): the texture-unit design.
//
// Same eight passes as gui/dplug/gui/legacypbr.d:269-1108, fused.  What changed against round 1's kernel (pbr_fast.cu in the history: 790 warp
// instructions per pixel row, FP32-issue-bound) is WHO evaluates the mip samplers:
//
//   * the bilinear samplers — Mipmap!L16.linearSample of depth levels 1..4 (ambient occlusion, legacypbr.d:400-444) and
//     Mipmap!RGBA.linearSample of diffuse levels 1..3 (emissive bloom, legacypbr.d:948-1007) — run on the TEXTURE UNITS:
//     every level of a device pyramid is also a pitch-2D texture object (linear filter, clamp addressing, unnormalised
//     coordinates).  The samplers are evaluated at pixel centres, (p + 0.5) * 2^-l, whose fractions are multiples of
//     2^-(l+1) >= 1/32: exactly representable in the hardware's 8-bit filter weights, and clamp addressing reproduces the
//     reference's edge rules (negative coordinate -> 0; upper index clamp after the fraction was taken, mipmap.d:309-339).
//     The rolling-row machinery of round 1 (staged windows of levels 1..3, refresh blocks, 26 registers of state)
//     disappears; the fetches for row j+1 are issued while row j is shaded.
//   * the trilinear skybox sample (linearMipmapSample, mipmap.d:149-160) fetches its 2 x 4 texels with texture GATHER
//     (tld4, one instruction per channel and level) from a CUDA array holding every skybox level (an atlas, each level
//     with a replicated one-texel border: one warp-uniform texture handle although the level varies per pixel): exact
//     texel values, filter weights computed in FP32 exactly as before — only the address arithmetic, the eight
//     dependent loads and the byte unpacking move to the texture unit.
//   * the bicubic levels 4 and 5 (16 / 32 pixels per texel) keep their four horizontally interpolated rows in registers.
//
// Work item = one CTA = a box of at most 64 x 64 pixels of one GUI tile (the host cuts areas into boxes of 64 x 2R
// rows); its four warps own the four quadrants (32 columns x R rows; a lane owns a column and walks down it).  The
// level-0 depth window of the box (+10 columns left / right, 10 rows up, 1 down: reach of the oblique shadow,
// legacypbr.d:525-534, and of the 3x3 normal stencil, legacypbr.d:301-314) is staged ONCE per CTA, already converted to
// float, with clamp-to-edge addressing (= the reference's replicated border, graphics.d:1121-1126).
//
// Numerical contract: +-1 LSB per 8-bit channel against the reference arithmetic (BASELINE.json).  The three RSQRTSS
// operands per pixel are computed with the reference's exact IEEE sequence (no FMA, same order) because the
// table-driven RSQRTSS is discontinuous and Blinn-Phong exponents up to 548 amplify it.
//
// Precondition (checked by the host): full pyramids, i.e. diffuse has levels 0..5 and depth 0..4 (every canvas of at
// least 32x32 pixels).  Smaller canvases take the exact kernel.
#include <cstdlib>
#include <cstring>
#include "pbr_math.cuh"

namespace b2 {

constexpr int T_HALO_X = 10, T_HALO_UP = 10;
constexpr int T_BOX = 64;                                  // CTA box: at most 64 x 64 pixels
constexpr int T_COLS = T_BOX + 2 * T_HALO_X + 4;           // +3: the staged window starts at a multiple of 4 columns; 88
constexpr int T_ROWS = T_BOX + T_HALO_UP + 1;              // 75
constexpr int T_WARPS = 4;
#ifndef B2_TEX_MIN_CTAS
#define B2_TEX_MIN_CTAS 4
#endif
constexpr int TC4 = 6, TC5 = 5;                            // bicubic windows: (32 >> l) + 4 texels along 32 pixels

struct RowT {                 // per-row set-up of one quadrant row (warp-uniform values), 80 bytes
    float wy4[4], wy5[4];     // Catmull-Rom weights of the four texel rows of diffuse levels 4 and 5
    float ey, ey2;            // toEye.y = j / H - 0.5 and its square (legacypbr.d:756, 914), exact
    int ib4, ib5;             // bicubic interval (floor(y_s) + 1)
    float tvD[3];             // texture v coordinate of diffuse levels 1..3: (j + 0.5) * 2^-l - first stored row
    float tvZ[4];             // ... of depth levels 1..4
    int refresh;              // bits 0-1 / 2-3: bicubic rows of level 4 / 5 (1 = slide by one, 2 = reload)
};
static_assert(sizeof(RowT) == 80, "RowT is read with 128-bit shared loads");

struct alignas(16) WarpT {
    RowT rows[32];
    uint32_t d4[TC4][TC4], d5[TC5][TC5];
    int org[4];               // window origins ox4, oy4, ox5, oy5
    int pad[3];
};
static_assert(sizeof(WarpT) % 16 == 0, "per-warp shared block keeps 16-byte alignment");

struct alignas(16) CtaT {
    float depth[T_ROWS][T_COLS];   // level-0 depth of the box + halo, as float
    WarpT w[T_WARPS];
};

struct G3 { float x, y, z; };
__device__ __forceinline__ G3 operator+(G3 a, G3 b) { return G3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ G3 operator-(G3 a, G3 b) { return G3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ G3 operator*(G3 a, float s) { return G3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ G3 fma3(float s, G3 a, G3 b) { return G3{fmaf(s, a.x, b.x), fmaf(s, a.y, b.y), fmaf(s, a.z, b.z)}; }

// byte k of a packed RGBA8 texel -> float, exact, without I2F: (0x4B000000 | b) is the float 2^23 + b
__device__ __forceinline__ float tByte(uint32_t p, int k) { return __uint_as_float(__byte_perm(p, 0x4B000000u, 0x7440 + k)) - 8388608.0f; }
__device__ __forceinline__ G3 tRgb(uint32_t p) { return G3{tByte(p, 0), tByte(p, 1), tByte(p, 2)}; }
__device__ __forceinline__ float tU16(uint32_t v) { return __uint_as_float(0x4B000000u | v) - 8388608.0f; }

__device__ __forceinline__ float tFmaSat(float a, float b, float c) {
    float r;
    asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}
__device__ __forceinline__ float tRcp(float x) {  // operands here are >= 0.7: no denormal handling needed
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float tPow(float x, float y) {  // x in [1e-3, ~1], y in (0, 548]: ex2(y * lg2(x))
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y * l));
    return r;
}
__device__ __forceinline__ uint32_t tCvtU8(float f) {  // truncate, clamp to [0,255], NaN -> 0
    uint32_t r;
    asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(f));
    return r & 0xffu;
}
// RSQRTSS emulation for positive normal operands (see pbr_math.cuh rsqrtss_x86)
template <bool MAY_BE_ZERO = true>
__device__ __forceinline__ float tRsqrtss(const uint32_t* __restrict__ tab, float x) {
    const uint32_t b = __float_as_uint(x);
    const uint32_t r = __ldg(tab + ((b >> 13) & 0x7ffu)) + 0x1F800000u - ((b >> 1) & 0x7F800000u);
    if (!MAY_BE_ZERO) return __uint_as_float(r);
    return b == 0u ? __uint_as_float(0x7f800000u) : __uint_as_float(r);
}
__device__ __forceinline__ void tCatmullRom(float t, float w[4]) {   // cubicInterp (mipmap.d:263-270) regrouped per sample
    float t2 = t * t, t3 = t2 * t;
    w[0] = -0.5f * t + t2 - 0.5f * t3;
    w[1] = 1.0f - 2.5f * t2 + 1.5f * t3;
    w[2] = 0.5f * t + 2.0f * t2 - 1.5f * t3;
    w[3] = -0.5f * t2 + 0.5f * t3;
}
// Mipmap.cubicSample, mipmap.d:182-204: y_s + 1 = (n + 2^(l+1)) / 2^(l+1) is always positive, so the truncations are
// floors; ib = floor(y_s) + 1; the four texels are clamp(ib - 2 + k)
__device__ __forceinline__ void tCubCoord(int l, int p, int& ib, float& t) {
    const int n1 = 2 * p + 1 - (1 << l) + (2 << l);
    ib = n1 >> (l + 1);
    t = (float)(n1 & ((2 << l) - 1)) * levelScale(l + 1);
}

// SKY_GATHER: the skybox footprints are fetched with texture gather and blended with FP32 weights (exact reference
// weights); otherwise the two bilinear samples use the texture unit's filter (8-bit weights; see the header comment)
template <bool HAS_SKY, bool ALL_PASSES, bool SKY_GATHER>
__global__ void __launch_bounds__(T_WARPS * 32, SKY_GATHER ? 4 : B2_TEX_MIN_CTAS) k_pbr_tex(const __grid_constant__ PbrArgs P) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    CtaT& S = *reinterpret_cast<CtaT*>(smemRaw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Box cbox = P.boxes[blockIdx.x];   // at most 64 x 64
    const int W = P.W, H = P.H;
    const DevLevel& D0 = P.depth[0];

    // ---------------------------------------------------------------- stage the depth window (whole CTA)
    // columns [cx0, cx0 + T_COLS), rows [ry0, cbox.maxy + 1), clamped to the image; four texels per thread and step
    const int cx0 = (cbox.minx - T_HALO_X) & ~3;
    const int ry0 = cbox.miny - T_HALO_UP;
    {
        const int nRows = (cbox.maxy - cbox.miny) + T_HALO_UP + 1;
        const int nQuads = (cbox.maxx + T_HALO_X - cx0 + 3) >> 2;          // <= T_COLS / 4
        const uint32_t magic = 0xffffffffu / (uint32_t)nQuads + 1u;
        for (uint32_t idx = threadIdx.x; idx < (uint32_t)(nRows * nQuads); idx += T_WARPS * 32) {
            const uint32_t r = __umulhi(idx, magic);
            const int q = (int)(idx - r * (uint32_t)nQuads);
            const int gy = max(0, min(ry0 + (int)r, H - 1)), gx = cx0 + 4 * q;
            const uint16_t* src = rowPtr<uint16_t>(D0, gy);
            float4 v;
            if (gx >= 0 && gx + 3 < W) {
                const uint2 t = __ldg(reinterpret_cast<const uint2*>(src + gx));
                v = make_float4(tU16(t.x & 0xffffu), tU16(t.x >> 16), tU16(t.y & 0xffffu), tU16(t.y >> 16));
            } else {
                v = make_float4(tU16(__ldg(src + max(0, min(gx, W - 1)))), tU16(__ldg(src + max(0, min(gx + 1, W - 1)))),
                                tU16(__ldg(src + max(0, min(gx + 2, W - 1)))), tU16(__ldg(src + max(0, min(gx + 3, W - 1)))));
            }
            *reinterpret_cast<float4*>(&S.depth[r][4 * q]) = v;
        }
    }

    // ---------------------------------------------------------------- this warp's quadrant
    const int bw = cbox.maxx - cbox.minx, bh = cbox.maxy - cbox.miny;
    const int hTop = (bh + 1) >> 1;
    Box box;
    box.minx = cbox.minx + ((warp & 1) ? 32 : 0);
    box.maxx = (warp & 1) ? cbox.maxx : min(cbox.maxx, cbox.minx + 32);
    box.miny = cbox.miny + ((warp & 2) ? hTop : 0);
    box.maxy = (warp & 2) ? cbox.maxy : cbox.miny + hTop;
    const bool quadrantActive = box.minx < box.maxx && box.miny < box.maxy && bw > 0 && bh > 0;
    WarpT& SW = S.w[warp];
    const int xLast = box.maxx - 1, yLast = box.maxy - 1;
    const int i = box.minx + lane;
    const bool active = i < box.maxx;
    const int ic = quadrantActive ? min(i, xLast) : 0;   // inactive lanes shadow the last column (no divergence, no stores)

    const uint32_t mask = ALL_PASSES ? 0xffu : P.passMask;
    const bool needMipsD = (mask & 0x40u) != 0, needMipsZ = (mask & 2u) != 0;

    if (quadrantActive) {
        // bicubic windows of diffuse levels 4 and 5
        int ibx, iby; float tt;
        tCubCoord(4, box.minx, ibx, tt); tCubCoord(4, box.miny, iby, tt);
        const int ox4 = max(0, min(ibx - 2, P.diffuse[4].w - 1)), oy4 = max(0, min(iby - 2, P.diffuse[4].h - 1));
        tCubCoord(5, box.minx, ibx, tt); tCubCoord(5, box.miny, iby, tt);
        const int ox5 = max(0, min(ibx - 2, P.diffuse[5].w - 1)), oy5 = max(0, min(iby - 2, P.diffuse[5].h - 1));
        if (lane == 0) { SW.org[0] = ox4; SW.org[1] = oy4; SW.org[2] = ox5; SW.org[3] = oy5; }
        for (int idx = lane; idx < TC4 * TC4; idx += 32) {
            const int r = idx / TC4, c = idx - r * TC4;
            const DevLevel& L = P.diffuse[4];
            SW.d4[r][c] = __ldg(rowPtr<uint32_t>(L, max(0, min(oy4 + r, L.h - 1))) + max(0, min(ox4 + c, L.w - 1)));
        }
        if (lane < TC5 * TC5) {
            const int r = lane / TC5, c = lane - r * TC5;
            const DevLevel& L = P.diffuse[5];
            SW.d5[r][c] = __ldg(rowPtr<uint32_t>(L, max(0, min(oy5 + r, L.h - 1))) + max(0, min(ox5 + c, L.w - 1)));
        }
        // per-row table: lane k prepares row miny + k
        if (box.miny + lane <= yLast) {
            const int j = box.miny + lane;
            RowT rs;
            float t;
            tCubCoord(4, j, rs.ib4, t); tCatmullRom(t, rs.wy4);
            tCubCoord(5, j, rs.ib5, t); tCatmullRom(t, rs.wy5);
            rs.ey = __fsub_rn(__fmul_rn((float)j, P.invH), 0.5f);
            rs.ey2 = __fmul_rn(rs.ey, rs.ey);
            const float jc = (float)j + 0.5f;
#pragma unroll
            for (int l = 1; l <= 3; ++l) rs.tvD[l - 1] = jc * levelScale(l) - (float)P.diffuse[l].y0;
#pragma unroll
            for (int l = 1; l <= 4; ++l) rs.tvZ[l - 1] = jc * levelScale(l) - (float)P.depth[l].y0;
            int refresh = 0xa;                         // first row of the quadrant: both levels are loaded
            if (lane > 0) {
                refresh = 0;
                int ibp; float tp;
                tCubCoord(4, j - 1, ibp, tp);
                if (rs.ib4 != ibp) refresh |= (rs.ib4 == ibp + 1 ? 1 : 2);
                tCubCoord(5, j - 1, ibp, tp);
                if (rs.ib5 != ibp) refresh |= (rs.ib5 == ibp + 1 ? 1 : 2) << 2;
            }
            rs.refresh = refresh;
            SW.rows[lane] = rs;
        }
    }
    __syncthreads();
    if (!quadrantActive) return;

    // ---------------------------------------------------------------- per-column constants
    const float ic05 = (float)ic + 0.5f;
    const float tu1 = ic05 * 0.5f, tu2 = ic05 * 0.25f, tu3 = ic05 * 0.125f, tu4 = ic05 * 0.0625f;   // texture u of levels 1..4
    // toEye.x and its square are column constants (legacypbr.d:756, 914) — exact sequence
    const float eyeX = __fsub_rn(0.5f, __fmul_rn((float)ic, P.invW));
    const float eyeX2 = __fmul_rn(eyeX, eyeX);
    const float nz2 = __fmul_rn(0.382658847856f, 0.382658847856f);
    const float multUshort = (float)(1.0 / 4655.0);
    const float kA = 0.00006510416f;  // 1/15360 as the reference's SIMD branch spells it (legacypbr.d:555)
    const unsigned char* noisePtr = kBlueNoise16x16 + (ic & 15) * 16;
    const float div255 = 1 / 255.0f;

    auto opaque = [](int v) { asm volatile("" : "+r"(v)); return v; };
    const volatile int* org = SW.org;
    // bicubic level (4 or 5): slide the four horizontally interpolated rows by one, or reload all of them.  The
    // horizontal set-up is recomputed here (every 16 / 32 rows) instead of living in registers for the whole strip.
    auto refreshCub = [&](int l, const DevLevel& L, auto& win, int oIdx, int ib, bool slide, G3 (&p)[4]) {
        int ibx; float t, w[4];
        tCubCoord(l, opaque(ic), ibx, t);
        tCatmullRom(t, w);
        const int ox = org[oIdx], oy = org[oIdx + 1];
        int xc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xc[k] = max(0, min(ibx - 2 + k, L.w - 1)) - ox;
        auto row = [&](int k) {
            const int r = max(0, min(ib - 2 + k, L.h - 1)) - oy;
            G3 v = tRgb(win[r][xc[0]]) * w[0];
            v = fma3(w[1], tRgb(win[r][xc[1]]), v);
            v = fma3(w[2], tRgb(win[r][xc[2]]), v);
            return fma3(w[3], tRgb(win[r][xc[3]]), v);
        };
        if (slide) { p[0] = p[1]; p[1] = p[2]; p[2] = p[3]; p[3] = row(3); }
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k] = row(k);
        }
    };
    G3 p4[4], p5[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) p4[k] = p5[k] = G3{0, 0, 0};

    // sliding 3x3 window of d = depth * (1/FACTOR_Z) (legacypbr.d:304-314): rows j-1, j, j+1 at columns i-1, i, i+1
    const int tx = ic - cx0;  // column of this lane inside the staged window (>= T_HALO_X)
    float dm[3], d0[3], dp[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        dm[k] = __fmul_rn(S.depth[box.miny - ry0 - 1][tx - 1 + k], multUshort);
        d0[k] = __fmul_rn(S.depth[box.miny - ry0][tx - 1 + k], multUshort);
    }

    // ---------------------------------------------------------------- fetches of the first row
    uint32_t diffuseNext = __ldg(rowPtr<uint32_t>(P.diffuse[0], box.miny) + ic);
    uint32_t materialNext = __ldg(rowPtr<uint32_t>(P.material, box.miny) + ic);
    // Raw results of the seven bilinear fetches of the NEXT row: they are summed only where the next row consumes them, a
    // whole row of arithmetic later (summing at the fetch would wait for the texture unit right there)
    float zA = 0.f, zB = 0.f, zC = 0.f, zD = 0.f;
    float4 eA = {0, 0, 0, 0}, eB = eA, eC = eA;
    auto fetchMips = [&](const RowT& rr) {
        if (needMipsZ) {
            zA = tex2D<float>(P.texZ[1], tu1, rr.tvZ[0]); zB = tex2D<float>(P.texZ[2], tu2, rr.tvZ[1]);
            zC = tex2D<float>(P.texZ[3], tu3, rr.tvZ[2]); zD = tex2D<float>(P.texZ[4], tu4, rr.tvZ[3]);
        }
        if (needMipsD) {
            eA = tex2D<float4>(P.texD[1], tu1, rr.tvD[0]); eB = tex2D<float4>(P.texD[2], tu2, rr.tvD[1]);
            eC = tex2D<float4>(P.texD[3], tu3, rr.tvD[2]);
        }
    };
    fetchMips(SW.rows[0]);

    for (int j = box.miny; j <= yLast; ++j) {
        const float* trow = &S.depth[j - ry0][tx];   // this lane's texel of row j inside the staged window
#pragma unroll
        for (int k = 0; k < 3; ++k) dp[k] = __fmul_rn(trow[T_COLS - 1 + k], multUshort);
        const uint32_t diffusePx = diffuseNext, materialPx = materialNext;
        const float zMips = (zA + zB) + (zC + zD);   // normalised (/ 65535) sum of the four bilinear depth samples
        const G3 eMips{(eA.x + eB.x) + eC.x, (eA.y + eB.y) + eC.y, (eA.z + eB.z) + eC.z};   // (/ 255) sum of the three diffuse samples
        const RowT& rs = SW.rows[j - box.miny];   // broadcast shared loads
        {   // prefetch the next row's pixels and mip samples (clamped on the last row: harmless re-read)
            const int jn = min(j + 1, yLast);
            diffuseNext = __ldg(rowPtr<uint32_t>(P.diffuse[0], jn) + ic);
            materialNext = __ldg(rowPtr<uint32_t>(P.material, jn) + ic);
            fetchMips(SW.rows[jn - box.miny]);
        }
        const int refresh = rs.refresh;

        // base colour and material stay in byte units (0..255): their 1/255 factors are folded into the light constants
        // the host passes pre-scaled (PbrArgs::*S) and into the constants below
        const G3 base = tRgb(diffusePx);
        const float roughB = tByte(materialPx, 0), metalB = tByte(materialPx, 1), specB = tByte(materialPx, 2);
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
            float inv = tRsqrtss<false>(P.rsqrtTab, len2);   // len2 >= 0.3827^2
            nx = vx * inv; ny = vy * inv; nzc = 0.382658847856f * inv;
            // variance (legacypbr.d:356-377): second differences of the 3x3 window.  The reference sums raw depths
            // (exact); here they are rebuilt from d (relative error ~1e-7, the variance feeds no RSQRTSS)
            // per row: s = left + right, X = s - 2 centre (horizontal second difference), R = s + centre (row sum)
            const float s0 = d[0] + d[2], s1 = d[3] + d[5], s2 = d[6] + d[8];
            const float ddx = (fmaf(-2.0f, d[1], s0) + fmaf(-2.0f, d[4], s1)) + fmaf(-2.0f, d[7], s2);
            const float ddy = fmaf(-2.0f, s1 + d[4], (s0 + d[1]) + (s2 + d[7]));
            const float kVar2 = (4655.0f / 256.0f) * (4655.0f / 256.0f);
            variance = fmaf(ddx, ddx, ddy * ddy) * kVar2;
        }
        const float zc = trow[0];
        G3 T{0, 0, 0};   // light arriving per channel; every pass but the emissive one is base colour x light

        // toEye (legacypbr.d:756-757, 914-915): exact operand for RSQRTSS, shared by specular and skybox
        float ex = eyeX, ey = 0, ez = 1.0f;
        if (mask & 0x30u) {
            ey = rs.ey;
            float len2 = __fadd_rn(__fadd_rn(eyeX2, rs.ey2), 1.0f);
            float inv = tRsqrtss<false>(P.rsqrtTab, len2);   // len2 >= 1
            ex = __fmul_rn(eyeX, inv); ey = __fmul_rn(ey, inv); ez = inv;  // 1.0f * inv
        }

        // ---- PassSkyboxReflections (legacypbr.d:872-930), first half: the gathers of the trilinear sample are the only
        // dependent fetches of a row (coordinates come from the normal), so they are issued HERE, before the occlusion /
        // shadow / directional / specular arithmetic, which then runs in their shadow; the blend follows after the
        // specular pass, where the reference adds the sky term (same order of additions into T)
        const bool doSky = HAS_SKY && (mask & 0x20u);
        float4 gr, gg, gb, hr, hg, hb;
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
            if (!SKY_GATHER) {
            // the two bilinear samples on the texture unit's filter.  The atlas position of the sample is clamped to the
            // level's texel centres, which reproduces the reference's edge rules (x < 0 -> 0; index clamp at w - 1)
            auto hw = [&](int level) {
                const float4 kx = P.skyCX[level], ky = P.skyCY[level];     // {o, o + 0.5, o + size - 0.5, 2^-level}
                const float cx = fminf(fmaxf(fmaf(skyx, kx.w, kx.x), kx.y), kx.z);
                const float cy = fminf(fmaxf(fmaf(skyy, ky.w, ky.x), ky.y), ky.z);
                return tex2D<float4>(P.texSkyLin, cx, cy);
            };
            gr = hw(max(0, min(il, P.nSky - 1)));
            hr = hw(max(0, min(il + 1, P.nSky - 1)));
            } else {
            // coordinate set-up of Mipmap.linearSample (mipmap.d:309-342) with the texel fetch left to the gather: the
            // footprint of a gather at (ix + 1, iy + 1) is exactly texels ix, min(ix + 1, w - 1) x iy, min(iy + 1, h - 1)
            auto setup = [&](int level, float& fx, float& fy, float& gx, float& gy) {
                const float sc = levelScale(level);
                float x = fmaxf(fmaf(skyx, sc, -0.5f), 0.f), y = fmaxf(fmaf(skyy, sc, -0.5f), 0.f);
                const float fxi = truncf(x), fyi = truncf(y);
                fx = x - fxi; fy = y - fyi;
                gx = (fminf(fxi, P.skyMaxX[level]) + 1.0f) + P.skyOx[level];   // small integers: exact
                gy = (fminf(fyi, P.skyMaxY[level]) + 1.0f) + P.skyOy[level];
            };
            const int level = max(0, min(il, P.nSky - 1));
            float gx, gy;
            setup(level, sfx, sfy, gx, gy);
            const cudaTextureObject_t t0 = P.texSky;
            gr = tex2Dgather<float4>(t0, gx, gy, 0); gg = tex2Dgather<float4>(t0, gx, gy, 1); gb = tex2Dgather<float4>(t0, gx, gy, 2);
            {   // second level of the trilinear blend: fetched also when fl == 0 (integer level), where the reference returns
                // the first level alone — the blend below then yields exactly that (sky + 0 * (sky1 - sky))
                const int level2 = max(0, min(il + 1, P.nSky - 1));
                setup(level2, s2fx, s2fy, gx, gy);
                hr = tex2Dgather<float4>(t0, gx, gy, 0); hg = tex2Dgather<float4>(t0, gx, gy, 1); hb = tex2Dgather<float4>(t0, gx, gy, 2);
            }
            }
        }

        // ---- PassAmbientOcclusion (legacypbr.d:400-444): depth levels 1..4, bilinear (texture units)
        if (needMipsZ) {
            float diff = fmaf(zMips, -0.25f * 65535.0f, zc);
            float cavity = tFmaSat(diff, 1.0f / 23040, 1.0f);
            const float amb = cavity * P.ambientS;
            T = G3{amb, amb, amb};
        }

        // ---- PassObliqueShadowLight (legacypbr.d:460-616): clamp(1 + (z + s - d)/15360, 0, 1) per tap
        if (mask & 4u) {
            float lightL = 0.f, lightR = 0.f;
            const float k0 = fmaf(zc, kA, 1.0f);
#pragma unroll
            for (int s = 1; s <= 10; ++s) {
                // the staged window is clamp-filled, so the taps (i +- s, j - s) (legacypbr.d:525-534, 568-576)
                // sit at compile-time offsets from this lane's texel
                const float ks = fmaf((float)s, kA, k0);
                float c1 = tFmaSat(trow[-s * T_COLS + s], -kA, ks);
                float c2 = tFmaSat(trow[-s * T_COLS - s], -kA, ks);
                lightR = fmaf(c1, P.shadowW[s], lightR);
                lightL = fmaf(c2, P.shadowW[s], lightL);
            }
            float k = fmaf(lightL, 0.7f, lightR) * P.shadowInvTotal;
            T.x = fmaf(P.light1S[0], k, T.x);
            T.y = fmaf(P.light1S[1], k, T.y);
            T.z = fmaf(P.light1S[2], k, T.z);
        }

        // ---- PassDirectionalLight (legacypbr.d:636-666)
        if (mask & 8u) {
            float dot = fmaf(nzc, P.light2Dir[2], fmaf(ny, P.light2Dir[1], nx * P.light2Dir[0]));
            float f = fmaf(0.5f, dot, 0.5f);
            float a = fmaf(roughB, -0.5f * div255, 0.24f);
            f = (f - a) * tRcp(1.0f - a);  // linmap(f, a, 1, 0, 1), core/dplug/core/math.d:25-28
            T.x = fmaf(P.light2ColS[0], f, T.x);
            T.y = fmaf(P.light2ColS[1], f, T.y);
            T.z = fmaf(P.light2ColS[2], f, T.z);
        }

        // ---- PassSpecularLight (legacypbr.d:720-813)
        if (mask & 0x10u) {
            float hx = __fsub_rn(ex, -P.light3Dir[0]), hy = __fsub_rn(ey, -P.light3Dir[1]), hz = __fsub_rn(ez, -P.light3Dir[2]);
            float hl2 = __fadd_rn(__fadd_rn(__fmul_rn(hx, hx), __fmul_rn(hy, hy)), __fmul_rn(hz, hz));
            float inv = tRsqrtss(P.rsqrtTab, hl2);
            float sf = fmaf(hz * inv, nzc, fmaf(hy * inv, ny, (hx * inv) * nx));
            sf = fmaxf(sf, 1e-3f);
            float exponent = __ldg(P.expTab + roughByte);
            float Ft = tRcp(fmaf(exponent * variance, 4e-5f, 1.0f));
            float eF = exponent * Ft;
            float tok = (1.0f + eF) * tRcp(1.0f + exponent);
            float p = tPow(sf, eF);
            float roughFactor = fmaf(roughB, -10.0f * div255, 10.0f) * fmaf(metalB, -0.5f * div255, 1.0f);
            float k = p * roughFactor * tok * specB;            // specular amount in byte units: light3ColS carries its 1/255
            T.x = fmaf(P.light3ColS[0], k, T.x);
            T.y = fmaf(P.light3ColS[1], k, T.y);
            T.z = fmaf(P.light3ColS[2], k, T.z);
        }

        // ---- PassEmissiveContribution (legacypbr.d:948-1007, non-legacy branch): diffuse levels 1..3 bilinear (texture
        // units), 4 and 5 bicubic.  Evaluated here, ahead of the sky blend, so that it runs in the shadow of the gathers too;
        // it is added to the accumulator last, as in the reference.
        G3 emitted{0, 0, 0};
        if (needMipsD) {
            if (refresh & 0x3) refreshCub(4, P.diffuse[4], SW.d4, 0, rs.ib4, (refresh & 0x1) != 0, p4);
            if (refresh & 0xc) refreshCub(5, P.diffuse[5], SW.d5, 2, rs.ib5, (refresh & 0x4) != 0, p5);
            G3 c4 = fma3(rs.wy4[3], p4[3], fma3(rs.wy4[2], p4[2], fma3(rs.wy4[1], p4[1], p4[0] * rs.wy4[0])));
            G3 c5 = fma3(rs.wy5[3], p5[3], fma3(rs.wy5[2], p5[2], fma3(rs.wy5[1], p5[1], p5[0] * rs.wy5[0])));
            // clamp_0_to_65535 (mipmap.d:250-261); 8-bit sources cannot reach the upper bound
            c4 = G3{fmaxf(c4.x, 0.f), fmaxf(c4.y, 0.f), fmaxf(c4.z, 0.f)};
            c5 = G3{fmaxf(c5.x, 0.f), fmaxf(c5.y, 0.f), fmaxf(c5.z, 0.f)};
            float noise = (tU16(__ldg(noisePtr + (j & 15))) - 127.5f) * 0.003f;
            const float AMT = 0.002f * 0.67f;
            float k5 = AMT * (1.0f + noise);
            emitted.x = fmaf(c5.x, k5, fmaf(c4.x, AMT, eMips.x * (AMT * 255.0f)));
            emitted.y = fmaf(c5.y, k5, fmaf(c4.y, AMT, eMips.y * (AMT * 255.0f)));
            emitted.z = fmaf(c5.z, k5, fmaf(c4.z, AMT, eMips.z * (AMT * 255.0f)));
        }

        // ---- PassSkyboxReflections, second half: blend the gathered texels (normalised floats, b / 255).
        // gather component order: x = (i0, j1), y = (i1, j1), z = (i1, j0), w = (i0, j0)
        if (doSky) {
            G3 sky;
            if (!SKY_GATHER) sky = G3{fmaf(fl, hr.x - gr.x, gr.x), fmaf(fl, hr.y - gr.y, gr.y), fmaf(fl, hr.z - gr.z, gr.z)};
            else {
                auto bil = [](const float4& g, float fx, float fy) {
                    const float up = fmaf(fx, g.z - g.w, g.w), down = fmaf(fx, g.y - g.x, g.x);
                    return fmaf(fy, down - up, up);
                };
                sky = G3{bil(gr, sfx, sfy), bil(gg, sfx, sfy), bil(gb, sfx, sfy)};
                const G3 sky1{bil(hr, s2fx, s2fy), bil(hg, s2fx, s2fy), bil(hb, s2fx, s2fy)};
                sky = fma3(fl, sky1 - sky, sky);
            }
            float k = metalB * P.skyAmountS;
            T.x = fmaf(sky.x, k, T.x);
            T.y = fmaf(sky.y, k, T.y);
            T.z = fmaf(sky.z, k, T.z);
        }

        const G3 accum{fmaf(base.x, T.x, emitted.x), fmaf(base.y, T.y, emitted.y), fmaf(base.z, T.z, emitted.z)};

        // ---- PassClampAndConvertTo8bit (legacypbr.d:1057-1107, non-legacy branch)
        if ((mask & 0x80u) && active) {
            float ex0 = fmaxf(0.0f, accum.x - P.toneThr), ex1 = fmaxf(0.0f, accum.y - P.toneThr), ex2 = fmaxf(0.0f, accum.z - P.toneThr);
            float luma = fmaf(0.072187f, ex2, fmaf(0.715158f, ex1, 0.212655f * ex0));
            float add = luma * (P.toneRatio * (1.0f / 3.0f));
            uint32_t out = tCvtU8((accum.x + add) * 255.99f) | (tCvtU8((accum.y + add) * 255.99f) << 8) |
                           (tCvtU8((accum.z + add) * 255.99f) << 16) | 0xff000000u;
            rowPtrW<uint32_t>(P.out, j)[i] = out;
        }

        // slide the depth window down
#pragma unroll
        for (int k = 0; k < 3; ++k) { dm[k] = d0[k]; d0[k] = dp[k]; }
    }
}

// ================================================================================================
// k_pbr_pair — the same passes with TWO ROWS PER LANE on the packed FP32 pipe.
//
// k_pbr_tex issues ~580 instructions per pixel, ~350 of them FP32 adds / multiplies / FMAs, and is bound by the issue
// slots (80 % busy), not by a pipe.  Blackwell's packed FP32 instructions (FADD2 / FMUL2 / FFMA2: two independent fp32
// lanes per instruction, each lane one correctly rounded operation; any operand may be a scalar register, a uniform
// register or an immediate BROADCAST to both lanes) halve the issue cost of everything that is computed the same way for
// two pixels.  Here a lane still owns a column, but walks it two rows at a time: every per-pixel quantity is a pair
// (row A = j, row B = j + 1) in a 64-bit register.  Packed: the plane fit, variance, eye / half vectors, reflection,
// occlusion, the shadow's per-tap constants and weighted sums, directional, specular set-up, the bicubic column sums
// (pair of row weights x scalar texel sum, broadcast), emissive, tone map.  Scalar per row: the table-driven RSQRTSS,
// MUFU (rcp / lg2 / ex2), saturating FMAs, min / max, byte <-> float conversions, texture fetches and the sums of
// their results.  Per-row set-up values come out of shared memory already interleaved (A, B) (PairT), the depth window
// is staged pre-multiplied by 1 / FACTOR_Z (the exact product the plane fit needs: no multiply in the loop).
//
// The exact sequences that feed RSQRTSS keep the reference's single roundings: a product that feeds an add is
// fma(a, b, -0) with the -0 pair passed as a kernel argument (mip_math.cuh: ptxas would otherwise contract FMUL2 + FADD2).
struct F2 { unsigned long long v; };
__device__ __forceinline__ F2 f2(float lo, float hi) { F2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float f2lo(F2 a) { float l, h; asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(a.v)); (void)h; return l; }
__device__ __forceinline__ float f2hi(F2 a) { float l, h; asm("mov.b64 {%0, %1}, %2;" : "=f"(l), "=f"(h) : "l"(a.v)); (void)l; return h; }
__device__ __forceinline__ F2 dup(float s) { return f2(s, s); }                                   // a broadcast operand in SASS, no instruction
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { F2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ F2 fma2(F2 a, F2 b, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ F2 operator*(F2 a, float s) { return a * dup(s); }
__device__ __forceinline__ F2 lds2(const float* p) { F2 r; r.v = *reinterpret_cast<const unsigned long long*>(p); return r; }   // 64-bit shared load of an interleaved pair

struct PairT {                 // per PAIR of rows of one quadrant (A = first row, B = second), interleaved (A, B); 160 bytes
    float wy4[4][2], wy5[4][2];   // Catmull-Rom weights of the four texel rows of diffuse levels 4 and 5
    float ey[2], ey2[2];          // toEye.y = j / H - 0.5 and its square (legacypbr.d:756, 914), exact
    float tvD[3][2];              // texture v of diffuse levels 1..3
    float tvZ[4][2];              // ... of depth levels 1..4
    int ib4[2], ib5[2];           // bicubic interval (floor(y_s) + 1)
    int refresh[2];               // bits 0-1 / 2-3: bicubic rows of level 4 / 5 (1 = slide by one, 2 = reload), for A and for B
};
static_assert(sizeof(PairT) == 160, "PairT is read with 64- and 128-bit shared loads");
struct alignas(16) WarpP {
    PairT pairs[16];
    uint32_t d4[TC4][TC4], d5[TC5][TC5];
    int org[4];
    int pad[3];
};
static_assert(sizeof(WarpP) % 16 == 0, "per-warp shared block keeps 16-byte alignment");
struct alignas(16) CtaP {
    float depth[T_ROWS][T_COLS];   // level-0 depth of the box + halo, times 1 / FACTOR_Z (the plane fit's exact operand)
    WarpP w[T_WARPS];
    float4 skyC[2][kMaxSkyLevels]; // PbrArgs::skyCX / skyCY: the level varies per lane, and a constant load replays once per distinct index
};

#ifndef B2_PAIR_MIN_CTAS
#define B2_PAIR_MIN_CTAS 4      // 128 registers, no spills.  Measured: 3 CTAs/SM (165 registers) 0.176 ms, 4: 0.157 ms, 5 (96 registers, 40 bytes spilled): 0.160 ms;
                                // 5 CTAs with the level-5 column sums in shared memory (ptxas then spills 188 bytes): 0.170 ms; RSQRTSS table staged in shared memory: 0.161 ms
#endif

template <bool HAS_SKY, bool ALL_PASSES>
__global__ void __launch_bounds__(T_WARPS * 32, B2_PAIR_MIN_CTAS) k_pbr_pair(const __grid_constant__ PbrArgs P) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    CtaP& S = *reinterpret_cast<CtaP*>(smemRaw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const Box cbox = P.boxes[blockIdx.x];   // at most 64 x 64
    const int W = P.W, H = P.H;
    const DevLevel& D0 = P.depth[0];
    const float multUshort = (float)(1.0 / 4655.0);

    // ---------------------------------------------------------------- stage the depth window (whole CTA), scaled
    const int cx0 = (cbox.minx - T_HALO_X) & ~3;
    const int ry0 = cbox.miny - T_HALO_UP;
    {
        const int nRows = (cbox.maxy - cbox.miny) + T_HALO_UP + 1;          // row B = j + 1 <= last row: its stencil ends on the row below the box
        const int nQuads = (cbox.maxx + T_HALO_X - cx0 + 3) >> 2;
        const uint32_t magic = 0xffffffffu / (uint32_t)nQuads + 1u;
        const uint32_t total = (uint32_t)(nRows * nQuads);
        auto put = [&](uint32_t r, int q, float4 v) {
            v = make_float4(__fmul_rn(v.x, multUshort), __fmul_rn(v.y, multUshort), __fmul_rn(v.z, multUshort), __fmul_rn(v.w, multUshort));
            *reinterpret_cast<float4*>(&S.depth[r][4 * q]) = v;
        };
        if (cx0 >= 0 && cx0 + 4 * nQuads <= W) {
            // window inside the image horizontally (all but the boxes on the left / right edge): four 64-bit loads per thread
            // in flight before the first conversion — the CTA starts with nothing else to hide their latency
            for (uint32_t base = threadIdx.x; base < total; base += 4 * T_WARPS * 32) {
                uint2 t[4];
                uint32_t rr[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t idx = base + u * T_WARPS * 32;
                    rr[u] = __umulhi(idx, magic);
                    if (idx < total) {
                        const int q = (int)(idx - rr[u] * (uint32_t)nQuads);
                        t[u] = __ldg(reinterpret_cast<const uint2*>(rowPtr<uint16_t>(D0, max(0, min(ry0 + (int)rr[u], H - 1))) + cx0 + 4 * q));
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t idx = base + u * T_WARPS * 32;
                    if (idx < total)
                        put(rr[u], (int)(idx - rr[u] * (uint32_t)nQuads), make_float4(tU16(t[u].x & 0xffffu), tU16(t[u].x >> 16), tU16(t[u].y & 0xffffu), tU16(t[u].y >> 16)));
                }
            }
        } else {
            for (uint32_t idx = threadIdx.x; idx < total; idx += T_WARPS * 32) {
                const uint32_t r = __umulhi(idx, magic);
                const int q = (int)(idx - r * (uint32_t)nQuads);
                const int gy = max(0, min(ry0 + (int)r, H - 1)), gx = cx0 + 4 * q;
                const uint16_t* src = rowPtr<uint16_t>(D0, gy);
                if (gx >= 0 && gx + 3 < W) {
                    const uint2 t = __ldg(reinterpret_cast<const uint2*>(src + gx));
                    put(r, q, make_float4(tU16(t.x & 0xffffu), tU16(t.x >> 16), tU16(t.y & 0xffffu), tU16(t.y >> 16)));
                } else {
                    put(r, q, make_float4(tU16(__ldg(src + max(0, min(gx, W - 1)))), tU16(__ldg(src + max(0, min(gx + 1, W - 1)))),
                                          tU16(__ldg(src + max(0, min(gx + 2, W - 1)))), tU16(__ldg(src + max(0, min(gx + 3, W - 1))))));
                }
            }
        }
    }

    if (HAS_SKY && threadIdx.x < 2 * kMaxSkyLevels) {
        const int ax = threadIdx.x >= kMaxSkyLevels ? 1 : 0, l = threadIdx.x - ax * kMaxSkyLevels;
        S.skyC[ax][l] = ax ? P.skyCY[l] : P.skyCX[l];
    }

    // ---------------------------------------------------------------- this warp's quadrant
    const int bw = cbox.maxx - cbox.minx, bh = cbox.maxy - cbox.miny;
    const int hTop = (bh + 1) >> 1;
    Box box;
    box.minx = cbox.minx + ((warp & 1) ? 32 : 0);
    box.maxx = (warp & 1) ? cbox.maxx : min(cbox.maxx, cbox.minx + 32);
    box.miny = cbox.miny + ((warp & 2) ? hTop : 0);
    box.maxy = (warp & 2) ? cbox.maxy : cbox.miny + hTop;
    const bool quadrantActive = box.minx < box.maxx && box.miny < box.maxy && bw > 0 && bh > 0;
    WarpP& SW = S.w[warp];
    const int xLast = box.maxx - 1, yLast = box.maxy - 1;
    const int i = box.minx + lane;
    const bool active = i < box.maxx;
    const int ic = quadrantActive ? min(i, xLast) : 0;

    const uint32_t mask = ALL_PASSES ? 0xffu : P.passMask;
    const bool needMipsD = (mask & 0x40u) != 0, needMipsZ = (mask & 2u) != 0;

    if (quadrantActive) {
        int ibx, iby; float tt;
        tCubCoord(4, box.minx, ibx, tt); tCubCoord(4, box.miny, iby, tt);
        const int ox4 = max(0, min(ibx - 2, P.diffuse[4].w - 1)), oy4 = max(0, min(iby - 2, P.diffuse[4].h - 1));
        tCubCoord(5, box.minx, ibx, tt); tCubCoord(5, box.miny, iby, tt);
        const int ox5 = max(0, min(ibx - 2, P.diffuse[5].w - 1)), oy5 = max(0, min(iby - 2, P.diffuse[5].h - 1));
        if (lane == 0) { SW.org[0] = ox4; SW.org[1] = oy4; SW.org[2] = ox5; SW.org[3] = oy5; }
        for (int idx = lane; idx < TC4 * TC4; idx += 32) {
            const int r = idx / TC4, c = idx - r * TC4;
            const DevLevel& L = P.diffuse[4];
            SW.d4[r][c] = __ldg(rowPtr<uint32_t>(L, max(0, min(oy4 + r, L.h - 1))) + max(0, min(ox4 + c, L.w - 1)));
        }
        if (lane < TC5 * TC5) {
            const int r = lane / TC5, c = lane - r * TC5;
            const DevLevel& L = P.diffuse[5];
            SW.d5[r][c] = __ldg(rowPtr<uint32_t>(L, max(0, min(oy5 + r, L.h - 1))) + max(0, min(ox5 + c, L.w - 1)));
        }
        // per-pair table: lane k prepares rows miny + 2k (A) and miny + 2k + 1 (B; the last row again when the quadrant has
        // an odd number of rows: computed, not stored)
        if (lane < 16 && box.miny + 2 * lane <= yLast) {
            PairT& ps = SW.pairs[lane];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = min(box.miny + 2 * lane + h, yLast);
                float t, w4[4], w5[4];
                int ib4, ib5;
                tCubCoord(4, j, ib4, t); tCatmullRom(t, w4);
                tCubCoord(5, j, ib5, t); tCatmullRom(t, w5);
#pragma unroll
                for (int k = 0; k < 4; ++k) { ps.wy4[k][h] = w4[k]; ps.wy5[k][h] = w5[k]; }
                ps.ib4[h] = ib4; ps.ib5[h] = ib5;
                const float ey = __fsub_rn(__fmul_rn((float)j, P.invH), 0.5f);
                ps.ey[h] = ey; ps.ey2[h] = __fmul_rn(ey, ey);
                const float jc = (float)j + 0.5f;
#pragma unroll
                for (int l = 1; l <= 3; ++l) ps.tvD[l - 1][h] = jc * levelScale(l) - (float)P.diffuse[l].y0;
#pragma unroll
                for (int l = 1; l <= 4; ++l) ps.tvZ[l - 1][h] = jc * levelScale(l) - (float)P.depth[l].y0;
                int refresh = 0xa;                         // first row of the quadrant: both levels are loaded
                if (lane > 0 || h > 0) {
                    refresh = 0;
                    const int jp = box.miny + 2 * lane + h - 1;   // the row processed just before (for a duplicated last row: itself - 1 is the same texel row or one above; compared below)
                    int ibp; float tp;
                    if (j != jp) {
                        tCubCoord(4, jp, ibp, tp);
                        if (ib4 != ibp) refresh |= (ib4 == ibp + 1 ? 1 : 2);
                        tCubCoord(5, jp, ibp, tp);
                        if (ib5 != ibp) refresh |= (ib5 == ibp + 1 ? 1 : 2) << 2;
                    }
                }
                ps.refresh[h] = refresh;
            }
        }
    }
    __syncthreads();
    if (!quadrantActive) return;

    // ---------------------------------------------------------------- per-column constants
    const float ic05 = (float)ic + 0.5f;
    const float tu1 = ic05 * 0.5f, tu2 = ic05 * 0.25f, tu3 = ic05 * 0.125f, tu4 = ic05 * 0.0625f;
    const float eyeX = __fsub_rn(0.5f, __fmul_rn((float)ic, P.invW));
    const float eyeX2 = __fmul_rn(eyeX, eyeX);
    const float nz2 = __fmul_rn(0.382658847856f, 0.382658847856f);
    const float kAs = 0.00006510416f * 4655.0f;   // 1/15360 (legacypbr.d:555) on the pre-scaled depth
    const unsigned char* noisePtr = kBlueNoise16x16 + (ic & 15) * 16;
    const float div255 = 1 / 255.0f;
    const F2 NZ{P.negZero2};

    auto opaque = [](int v) { asm volatile("" : "+r"(v)); return v; };
    const volatile int* org = SW.org;
    auto refreshCub = [&](int l, const DevLevel& L, auto& win, int oIdx, int ib, bool slide, G3 (&p)[4]) {
        int ibx; float t, w[4];
        tCubCoord(l, opaque(ic), ibx, t);
        tCatmullRom(t, w);
        const int ox = org[oIdx], oy = org[oIdx + 1];
        int xc[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) xc[k] = max(0, min(ibx - 2 + k, L.w - 1)) - ox;
        auto row = [&](int k) {
            const int r = max(0, min(ib - 2 + k, L.h - 1)) - oy;
            G3 v = tRgb(win[r][xc[0]]) * w[0];
            v = fma3(w[1], tRgb(win[r][xc[1]]), v);
            v = fma3(w[2], tRgb(win[r][xc[2]]), v);
            return fma3(w[3], tRgb(win[r][xc[3]]), v);
        };
        if (slide) { p[0] = p[1]; p[1] = p[2]; p[2] = p[3]; p[3] = row(3); }
        else {
#pragma unroll
            for (int k = 0; k < 4; ++k) p[k] = row(k);
        }
    };
    G3 p4[4], p5[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) p4[k] = p5[k] = G3{0, 0, 0};

    const int tx = ic - cx0;  // column of this lane inside the staged window (>= T_HALO_X)

    // ---------------------------------------------------------------- fetches of the first pair
    auto rowB = [&](int jA) { return min(jA + 1, yLast); };
    uint32_t diffuseNextA = __ldg(rowPtr<uint32_t>(P.diffuse[0], box.miny) + ic), diffuseNextB = __ldg(rowPtr<uint32_t>(P.diffuse[0], rowB(box.miny)) + ic);
    uint32_t materialNextA = __ldg(rowPtr<uint32_t>(P.material, box.miny) + ic), materialNextB = __ldg(rowPtr<uint32_t>(P.material, rowB(box.miny)) + ic);
    float zA0 = 0.f, zB0 = 0.f, zC0 = 0.f, zD0 = 0.f, zA1 = 0.f, zB1 = 0.f, zC1 = 0.f, zD1 = 0.f;   // depth mip fetches of the NEXT pair (row A: 0, row B: 1)
    auto fetchZ = [&](const PairT& pp) {
        if (needMipsZ) {
            zA0 = tex2D<float>(P.texZ[1], tu1, pp.tvZ[0][0]); zB0 = tex2D<float>(P.texZ[2], tu2, pp.tvZ[1][0]);
            zC0 = tex2D<float>(P.texZ[3], tu3, pp.tvZ[2][0]); zD0 = tex2D<float>(P.texZ[4], tu4, pp.tvZ[3][0]);
            zA1 = tex2D<float>(P.texZ[1], tu1, pp.tvZ[0][1]); zB1 = tex2D<float>(P.texZ[2], tu2, pp.tvZ[1][1]);
            zC1 = tex2D<float>(P.texZ[3], tu3, pp.tvZ[2][1]); zD1 = tex2D<float>(P.texZ[4], tu4, pp.tvZ[3][1]);
        }
    };
    fetchZ(SW.pairs[0]);

    for (int jA = box.miny; jA <= yLast; jA += 2) {
        const int pi = (jA - box.miny) >> 1;
        const PairT& ps = SW.pairs[pi];
        const bool validB = jA + 1 <= yLast;
        const int jB = validB ? jA + 1 : jA;
        const float* trow = &S.depth[jA - ry0][tx];          // row A's texel of this lane in the staged window
        const float* trowB = trow + (validB ? T_COLS : 0);   // row B's
        // 3x3 windows of d = depth / FACTOR_Z for both rows (legacypbr.d:304-314): pair = (row A's value, row B's value)
        F2 d[9];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            d[k] = f2(trow[-T_COLS - 1 + k], trowB[-T_COLS - 1 + k]);
            d[3 + k] = f2(trow[-1 + k], trowB[-1 + k]);
            d[6 + k] = f2(trow[T_COLS - 1 + k], trowB[T_COLS - 1 + k]);
        }
        const uint32_t diffuseA = diffuseNextA, diffuseB = diffuseNextB, materialA = materialNextA, materialB = materialNextB;
        const F2 zMips = (f2(zA0, zA1) + f2(zB0, zB1)) + (f2(zC0, zC1) + f2(zD0, zD1));   // normalised (/ 65535) sums of the four bilinear depth samples
        {   // prefetch the next pair's pixels and depth mips (clamped at the end of the quadrant: harmless re-read)
            const int jn = min(jA + 2, yLast), jnB = min(jA + 3, yLast);
            diffuseNextA = __ldg(rowPtr<uint32_t>(P.diffuse[0], jn) + ic); diffuseNextB = __ldg(rowPtr<uint32_t>(P.diffuse[0], jnB) + ic);
            materialNextA = __ldg(rowPtr<uint32_t>(P.material, jn) + ic); materialNextB = __ldg(rowPtr<uint32_t>(P.material, jnB) + ic);
            fetchZ(SW.pairs[(jn - box.miny) >> 1]);
        }
        const int refreshA = ps.refresh[0], refreshB = ps.refresh[1];

        // base colour and material in byte units (their 1/255 factors live in the pre-scaled light constants)
        const F2 magic = dup(-8388608.0f);
        auto byte2 = [&](uint32_t a, uint32_t b, int k) {
            return f2(__uint_as_float(__byte_perm(a, 0x4B000000u, 0x7440 + k)), __uint_as_float(__byte_perm(b, 0x4B000000u, 0x7440 + k))) + magic;
        };
        const F2 baseX = byte2(diffuseA, diffuseB, 0), baseY = byte2(diffuseA, diffuseB, 1), baseZ = byte2(diffuseA, diffuseB, 2);
        const F2 roughB = byte2(materialA, materialB, 0), metalB = byte2(materialA, materialB, 1), specB = byte2(materialA, materialB, 2);
        const int roughByteA = materialA & 0xff, roughByteB = materialB & 0xff;

        // ---- PassComputeNormal (legacypbr.d:279-381) + computePlaneFittingNormal (ransac.d:53-121): the reference's exact
        // sequence, two rows per instruction; X(a, b) = round(a * b) that may feed an add
        F2 nx = dup(0.f), ny = dup(0.f), nzc = dup(0.f), variance = dup(0.f);
        if (mask & 1u) {
            auto X = [&](F2 a, float b) { return fma2(a, dup(b), NZ); };
            auto XX = [&](F2 a, F2 b) { return fma2(a, b, NZ); };
            const float c = 0.03660694846f, e = 0.118115527008f;
            const F2 m0 = X(d[0], c) + X(d[5], e);
            const F2 m1 = X(d[1], e) + X(d[6], c);
            const F2 m2 = X(d[2], c) + X(d[7], e);
            const F2 m3 = X(d[3], e) + X(d[8], c);
            const F2 filt = (((X(d[4], 0.38111009811f) + m0) + m1) + m2) + m3;
            const F2 a0 = d[0] - filt, a1 = d[1] - filt, a2 = d[2] - filt, a3 = d[3] - filt;
            const F2 b0 = d[5] - filt, b1 = d[6] - filt, b2 = d[7] - filt, b3 = d[8] - filt;
            const F2 a0c = X(a0, -c), b3c = X(b3, c);
            const F2 xz0 = a0c + X(b0, e);
            const F2 xz1 = X(b1, -c);
            const F2 xz2 = X(a2, c);
            const F2 xz3 = X(a3, -e) + b3c;
            const F2 yz1 = X(a1, -e) + X(b1, c);
            const F2 yz2 = X(a2, -c) + X(b2, e);
            const F2 xz = ((xz0 + xz1) + xz2) + xz3;
            const F2 yz = ((a0c + yz1) + yz2) + b3c;
            const F2 vx = dup(0.f) - xz, vy = yz;
            const F2 len2 = (XX(vx, vx) + XX(vy, vy)) + dup(nz2);
            const F2 inv = f2(tRsqrtss<false>(P.rsqrtTab, f2lo(len2)), tRsqrtss<false>(P.rsqrtTab, f2hi(len2)));   // len2 >= 0.3827^2
            nx = vx * inv; ny = vy * inv; nzc = inv * 0.382658847856f;
            // variance (legacypbr.d:356-377): second differences of the 3x3 window (relative error ~1e-7, feeds no RSQRTSS)
            const F2 s0 = d[0] + d[2], s1 = d[3] + d[5], s2 = d[6] + d[8];
            const F2 ddx = (fma2(d[1], dup(-2.0f), s0) + fma2(d[4], dup(-2.0f), s1)) + fma2(d[7], dup(-2.0f), s2);
            const F2 ddy = fma2(s1 + d[4], dup(-2.0f), (s0 + d[1]) + (s2 + d[7]));
            const float kVar2 = (4655.0f / 256.0f) * (4655.0f / 256.0f);
            variance = fma2(ddx, ddx, ddy * ddy) * kVar2;
        }
        const F2 zc = d[4];       // centre depth / FACTOR_Z
        F2 Tx = dup(0.f), Ty = dup(0.f), Tz = dup(0.f);   // light arriving per channel

        // toEye (legacypbr.d:756-757, 914-915): exact operand for RSQRTSS, shared by specular and skybox
        F2 ex = dup(eyeX), ey = dup(0.f), ez = dup(1.0f);
        if (mask & 0x30u) {
            const F2 len2 = (dup(eyeX2) + lds2(ps.ey2)) + dup(1.0f);
            const F2 inv = f2(tRsqrtss<false>(P.rsqrtTab, f2lo(len2)), tRsqrtss<false>(P.rsqrtTab, f2hi(len2)));   // len2 >= 1
            ex = inv * eyeX; ey = lds2(ps.ey) * inv; ez = inv;
        }

        // ---- PassSkyboxReflections (legacypbr.d:872-930), first half: the two filtered fetches per row are issued here
        const bool doSky = HAS_SKY && (mask & 0x20u);
        float4 g0, h0, g1, h1;
        F2 fl = dup(0.f);
        if (doSky) {
            F2 lod = fma2(variance, dup(P.skyPixels * 0.00001f), dup(1.0f));
            lod = fma2(f2(fastlog2(f2lo(lod)), fastlog2(f2hi(lod))), dup(0.5f), roughB * (6.0f / 255.0f));
            const F2 nsum = fma2(ez, nzc, fma2(ey, ny, ex * nx)) * -2.0f;        // -2 (toEye . n), _mm_reflectnormal_ps (ransac.d:31-36)
            const F2 rx = fma2(nsum, nx, ex), ry = fma2(nsum, ny, ey);
            const F2 skyx = fma2(fma2(rx, dup(-0.5f), dup(0.5f)), dup(P.skyFX), dup(0.5f));
            const F2 skyy = fma2(fma2(ry, dup(0.5f), dup(0.5f)), dup(P.skyFY), dup(0.5f));
            const int il0 = __float2int_rz(f2lo(lod)), il1 = __float2int_rz(f2hi(lod));   // linearMipmapSample, mipmap.d:149-160
            fl = lod - f2((float)il0, (float)il1);
            auto hw = [&](int level, float sx, float sy) {
                const float4 kx = S.skyC[0][level], ky = S.skyC[1][level];   // {o, o + 0.5, o + size - 0.5, 2^-level}
                const float cx = fminf(fmaxf(fmaf(sx, kx.w, kx.x), kx.y), kx.z);
                const float cy = fminf(fmaxf(fmaf(sy, ky.w, ky.x), ky.y), ky.z);
                return tex2D<float4>(P.texSkyLin, cx, cy);
            };
            g0 = hw(max(0, min(il0, P.nSky - 1)), f2lo(skyx), f2lo(skyy));
            h0 = hw(max(0, min(il0 + 1, P.nSky - 1)), f2lo(skyx), f2lo(skyy));
            g1 = hw(max(0, min(il1, P.nSky - 1)), f2hi(skyx), f2hi(skyy));
            h1 = hw(max(0, min(il1 + 1, P.nSky - 1)), f2hi(skyx), f2hi(skyy));
        }

        // ---- PassAmbientOcclusion (legacypbr.d:400-444): cavity = sat(1 + (z - mean of the four mip samples) / 23040)
        if (needMipsZ) {
            const F2 t = fma2(zMips, dup(-0.25f * 65535.0f / 23040.0f), dup(1.0f));
            const F2 cavity = f2(tFmaSat(f2lo(zc), 4655.0f / 23040.0f, f2lo(t)), tFmaSat(f2hi(zc), 4655.0f / 23040.0f, f2hi(t)));
            const F2 amb = cavity * P.ambientS;
            Tx = amb; Ty = amb; Tz = amb;
        }

        // ---- PassObliqueShadowLight (legacypbr.d:460-616): clamp(1 + (z + s - d)/15360, 0, 1) per tap
        if (mask & 4u) {
            F2 lightL = dup(0.f), lightR = dup(0.f);
            const F2 k0 = fma2(zc, dup(kAs), dup(1.0f));
#pragma unroll
            for (int s = 1; s <= 10; ++s) {
                const F2 ks = k0 + dup((float)s * 0.00006510416f);
                const F2 c1 = f2(tFmaSat(trow[-s * T_COLS + s], -kAs, f2lo(ks)), tFmaSat(trowB[-s * T_COLS + s], -kAs, f2hi(ks)));
                const F2 c2 = f2(tFmaSat(trow[-s * T_COLS - s], -kAs, f2lo(ks)), tFmaSat(trowB[-s * T_COLS - s], -kAs, f2hi(ks)));
                lightR = fma2(c1, dup(P.shadowW[s]), lightR);
                lightL = fma2(c2, dup(P.shadowW[s]), lightL);
            }
            const F2 k = fma2(lightL, dup(0.7f), lightR) * P.shadowInvTotal;
            Tx = fma2(k, dup(P.light1S[0]), Tx);
            Ty = fma2(k, dup(P.light1S[1]), Ty);
            Tz = fma2(k, dup(P.light1S[2]), Tz);
        }

        // the diffuse mip fetches of this pair, consumed by the emissive pass: issued here, a directional + specular pass
        // ahead of their use (issued at the top of the row pair they hold 24 registers through the whole body)
        float4 eA0 = {0, 0, 0, 0}, eB0 = eA0, eC0 = eA0, eA1 = eA0, eB1 = eA0, eC1 = eA0;
        if (needMipsD) {
            eA0 = tex2D<float4>(P.texD[1], tu1, ps.tvD[0][0]); eB0 = tex2D<float4>(P.texD[2], tu2, ps.tvD[1][0]); eC0 = tex2D<float4>(P.texD[3], tu3, ps.tvD[2][0]);
            eA1 = tex2D<float4>(P.texD[1], tu1, ps.tvD[0][1]); eB1 = tex2D<float4>(P.texD[2], tu2, ps.tvD[1][1]); eC1 = tex2D<float4>(P.texD[3], tu3, ps.tvD[2][1]);
        }

        // ---- PassDirectionalLight (legacypbr.d:636-666)
        if (mask & 8u) {
            const F2 dot = fma2(nzc, dup(P.light2Dir[2]), fma2(ny, dup(P.light2Dir[1]), nx * P.light2Dir[0]));
            const F2 f0 = fma2(dot, dup(0.5f), dup(0.5f));
            const F2 a = fma2(roughB, dup(-0.5f * div255), dup(0.24f));
            const F2 oma = dup(1.0f) - a;
            const F2 f = (f0 - a) * f2(tRcp(f2lo(oma)), tRcp(f2hi(oma)));   // linmap(f, a, 1, 0, 1), core/dplug/core/math.d:25-28
            Tx = fma2(f, dup(P.light2ColS[0]), Tx);
            Ty = fma2(f, dup(P.light2ColS[1]), Ty);
            Tz = fma2(f, dup(P.light2ColS[2]), Tz);
        }

        // ---- PassSpecularLight (legacypbr.d:720-813)
        if (mask & 0x10u) {
            auto XX = [&](F2 a, F2 b) { return fma2(a, b, NZ); };
            const F2 hx = ex + dup(P.light3Dir[0]), hy = ey + dup(P.light3Dir[1]), hz = ez + dup(P.light3Dir[2]);   // toEye - (-light3Dir)
            const F2 hl2 = (XX(hx, hx) + XX(hy, hy)) + XX(hz, hz);
            const F2 inv = f2(tRsqrtss(P.rsqrtTab, f2lo(hl2)), tRsqrtss(P.rsqrtTab, f2hi(hl2)));
            F2 sf = fma2(hz * inv, nzc, fma2(hy * inv, ny, (hx * inv) * nx));
            sf = f2(fmaxf(f2lo(sf), 1e-3f), fmaxf(f2hi(sf), 1e-3f));
            const F2 exponent = f2(__ldg(P.expTab + roughByteA), __ldg(P.expTab + roughByteB));
            const F2 ftd = fma2(exponent * variance, dup(4e-5f), dup(1.0f));
            const F2 Ft = f2(tRcp(f2lo(ftd)), tRcp(f2hi(ftd)));
            const F2 eF = exponent * Ft;
            const F2 ope = exponent + dup(1.0f);
            const F2 tok = (eF + dup(1.0f)) * f2(tRcp(f2lo(ope)), tRcp(f2hi(ope)));
            const F2 p = f2(tPow(f2lo(sf), f2lo(eF)), tPow(f2hi(sf), f2hi(eF)));
            const F2 roughFactor = fma2(roughB, dup(-10.0f * div255), dup(10.0f)) * fma2(metalB, dup(-0.5f * div255), dup(1.0f));
            const F2 k = ((p * roughFactor) * tok) * specB;            // specular amount in byte units: light3ColS carries its 1/255
            Tx = fma2(k, dup(P.light3ColS[0]), Tx);
            Ty = fma2(k, dup(P.light3ColS[1]), Ty);
            Tz = fma2(k, dup(P.light3ColS[2]), Tz);
        }

        // ---- PassSkyboxReflections, second half: blend the two levels, add
        if (doSky) {
            const float fl0 = f2lo(fl), fl1 = f2hi(fl);
            const F2 skx = f2(fmaf(fl0, h0.x - g0.x, g0.x), fmaf(fl1, h1.x - g1.x, g1.x));
            const F2 sky_ = f2(fmaf(fl0, h0.y - g0.y, g0.y), fmaf(fl1, h1.y - g1.y, g1.y));
            const F2 skz = f2(fmaf(fl0, h0.z - g0.z, g0.z), fmaf(fl1, h1.z - g1.z, g1.z));
            const F2 k = metalB * P.skyAmountS;
            Tx = fma2(skx, k, Tx);
            Ty = fma2(sky_, k, Ty);
            Tz = fma2(skz, k, Tz);
        }

        // ---- PassEmissiveContribution (legacypbr.d:948-1007): diffuse levels 1..3 bilinear (texture units), 4 and 5 bicubic
        F2 emX = dup(0.f), emY = dup(0.f), emZ = dup(0.f);
        if (needMipsD) {
            if (refreshA & 0x3) refreshCub(4, P.diffuse[4], SW.d4, 0, ps.ib4[0], (refreshA & 0x1) != 0, p4);
            if (refreshA & 0xc) refreshCub(5, P.diffuse[5], SW.d5, 2, ps.ib5[0], (refreshA & 0x4) != 0, p5);
            F2 c4x, c4y, c4z, c5x, c5y, c5z;
            if (!(refreshB & 0xf)) {
                // both rows read the same four texel rows: pair of row weights x (broadcast) column sums
                const F2 w0 = lds2(ps.wy4[0]), w1 = lds2(ps.wy4[1]), w2 = lds2(ps.wy4[2]), w3 = lds2(ps.wy4[3]);
                c4x = fma2(w3, dup(p4[3].x), fma2(w2, dup(p4[2].x), fma2(w1, dup(p4[1].x), w0 * p4[0].x)));
                c4y = fma2(w3, dup(p4[3].y), fma2(w2, dup(p4[2].y), fma2(w1, dup(p4[1].y), w0 * p4[0].y)));
                c4z = fma2(w3, dup(p4[3].z), fma2(w2, dup(p4[2].z), fma2(w1, dup(p4[1].z), w0 * p4[0].z)));
                const F2 v0 = lds2(ps.wy5[0]), v1 = lds2(ps.wy5[1]), v2 = lds2(ps.wy5[2]), v3 = lds2(ps.wy5[3]);
                c5x = fma2(v3, dup(p5[3].x), fma2(v2, dup(p5[2].x), fma2(v1, dup(p5[1].x), v0 * p5[0].x)));
                c5y = fma2(v3, dup(p5[3].y), fma2(v2, dup(p5[2].y), fma2(v1, dup(p5[1].y), v0 * p5[0].y)));
                c5z = fma2(v3, dup(p5[3].z), fma2(v2, dup(p5[2].z), fma2(v1, dup(p5[1].z), v0 * p5[0].z)));
            } else {
                // a texel-row boundary of level 4 or 5 runs between the two rows (boxes that do not start on an even row)
                auto col = [&](const G3 (&p)[4], const float (&w)[4][2], int h) {
                    return fma3(w[3][h], p[3], fma3(w[2][h], p[2], fma3(w[1][h], p[1], p[0] * w[0][h])));
                };
                const G3 a4 = col(p4, ps.wy4, 0), a5 = col(p5, ps.wy5, 0);
                if (refreshB & 0x3) refreshCub(4, P.diffuse[4], SW.d4, 0, ps.ib4[1], (refreshB & 0x1) != 0, p4);
                if (refreshB & 0xc) refreshCub(5, P.diffuse[5], SW.d5, 2, ps.ib5[1], (refreshB & 0x4) != 0, p5);
                const G3 b4 = col(p4, ps.wy4, 1), b5 = col(p5, ps.wy5, 1);
                c4x = f2(a4.x, b4.x); c4y = f2(a4.y, b4.y); c4z = f2(a4.z, b4.z);
                c5x = f2(a5.x, b5.x); c5y = f2(a5.y, b5.y); c5z = f2(a5.z, b5.z);
            }
            // clamp_0_to_65535 (mipmap.d:250-261); 8-bit sources cannot reach the upper bound
            auto max0 = [](F2 v) { return f2(fmaxf(f2lo(v), 0.f), fmaxf(f2hi(v), 0.f)); };
            c4x = max0(c4x); c4y = max0(c4y); c4z = max0(c4z); c5x = max0(c5x); c5y = max0(c5y); c5z = max0(c5z);
            const F2 noise = (f2(tU16(__ldg(noisePtr + (jA & 15))), tU16(__ldg(noisePtr + (jB & 15)))) - dup(127.5f)) * 0.003f;
            const float AMT = 0.002f * 0.67f;
            const F2 k5 = fma2(noise, dup(AMT), dup(AMT));
            // (/ 255) sums of the three bilinear samples, per row
            const F2 eMx = f2((eA0.x + eB0.x) + eC0.x, (eA1.x + eB1.x) + eC1.x);
            const F2 eMy = f2((eA0.y + eB0.y) + eC0.y, (eA1.y + eB1.y) + eC1.y);
            const F2 eMz = f2((eA0.z + eB0.z) + eC0.z, (eA1.z + eB1.z) + eC1.z);
            emX = fma2(c5x, k5, fma2(c4x, dup(AMT), eMx * (AMT * 255.0f)));
            emY = fma2(c5y, k5, fma2(c4y, dup(AMT), eMy * (AMT * 255.0f)));
            emZ = fma2(c5z, k5, fma2(c4z, dup(AMT), eMz * (AMT * 255.0f)));
        }

        const F2 acX = fma2(baseX, Tx, emX), acY = fma2(baseY, Ty, emY), acZ = fma2(baseZ, Tz, emZ);

        // ---- PassClampAndConvertTo8bit (legacypbr.d:1057-1107, non-legacy branch)
        if ((mask & 0x80u) && active) {
            auto max0 = [](F2 v) { return f2(fmaxf(f2lo(v), 0.f), fmaxf(f2hi(v), 0.f)); };
            const F2 thr = dup(P.toneThr);
            const F2 ex0 = max0(acX - thr), ex1 = max0(acY - thr), ex2 = max0(acZ - thr);
            const F2 luma = fma2(ex2, dup(0.072187f), fma2(ex1, dup(0.715158f), ex0 * 0.212655f));
            const F2 add = luma * (P.toneRatio * (1.0f / 3.0f));
            const F2 oX = (acX + add) * 255.99f, oY = (acY + add) * 255.99f, oZ = (acZ + add) * 255.99f;
            rowPtrW<uint32_t>(P.out, jA)[i] = tCvtU8(f2lo(oX)) | (tCvtU8(f2lo(oY)) << 8) | (tCvtU8(f2lo(oZ)) << 16) | 0xff000000u;
            if (validB) rowPtrW<uint32_t>(P.out, jB)[i] = tCvtU8(f2hi(oX)) | (tCvtU8(f2hi(oY)) << 8) | (tCvtU8(f2hi(oZ)) << 16) | 0xff000000u;
        }
    }
}

// Skybox atlas (set-up time): texel (x, y) of the atlas is the clamped texel of the level whose bordered rectangle
// contains it, 0 elsewhere.  `ox / oy` = atlas position of texel (0, 0) of each level.
struct SkyAtlasArgs {
    DevLevel lv[kMaxSkyLevels];
    int ox[kMaxSkyLevels], oy[kMaxSkyLevels];
    int n, aw, ah;
    uint32_t* out;   // gapless aw x ah
};
__global__ void __launch_bounds__(256) k_sky_atlas(const __grid_constant__ SkyAtlasArgs A) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= A.aw || y >= A.ah) return;
    uint32_t v = 0;
    for (int l = 0; l < A.n; ++l) {
        const int lx = x - A.ox[l], ly = y - A.oy[l];
        if (lx >= -1 && lx <= A.lv[l].w && ly >= -1 && ly <= A.lv[l].h) {
            v = rowPtr<uint32_t>(A.lv[l], max(0, min(ly, A.lv[l].h - 1)))[max(0, min(lx, A.lv[l].w - 1))];
            break;
        }
    }
    A.out[(size_t)y * A.aw + x] = v;
}
// lays the levels out (level 0 at (1, 1), the others stacked in a column to its right) and fills `out` (aw * ah texels)
void sky_atlas_layout(const DevLevel* levels, int n, int* ox, int* oy, int* aw, int* ah) {
    ox[0] = 1; oy[0] = 1;
    int y = 1, colW = 0;
    const int x1 = levels[0].w + 3;
    for (int l = 1; l < n; ++l) {
        ox[l] = x1; oy[l] = y;
        y += levels[l].h + 2;
        if (levels[l].w > colW) colW = levels[l].w;
    }
    *aw = n > 1 ? x1 + colW + 1 : levels[0].w + 2;
    *ah = (levels[0].h + 2 > y + 1 ? levels[0].h + 2 : y + 1);
}
void launch_sky_atlas(const DevLevel* levels, int n, const int* ox, const int* oy, int aw, int ah, uint32_t* out, cudaStream_t s) {
    SkyAtlasArgs A;
    memset(&A, 0, sizeof(A));
    for (int l = 0; l < n; ++l) { A.lv[l] = levels[l]; A.ox[l] = ox[l]; A.oy[l] = oy[l]; }
    A.n = n; A.aw = aw; A.ah = ah; A.out = out;
    k_sky_atlas<<<dim3((aw + 31) / 32, (ah + 7) / 8), 256, 0, s>>>(A);
}

int pbr_tex_box_rows() { return T_BOX; }

static bool g_texAttrSet[64];
template <bool SKY, bool ALL, bool GATHER>
static void launchTexT(const PbrArgs& args, size_t smem, bool setAttr, cudaStream_t stream) {
    if (setAttr) cudaFuncSetAttribute(k_pbr_tex<SKY, ALL, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    else k_pbr_tex<SKY, ALL, GATHER><<<args.nBoxes, T_WARPS * 32, smem, stream>>>(args);
}
static void dispatchTex(const PbrArgs& args, bool gather, size_t smem, bool setAttr, cudaStream_t stream) {
    const bool all = setAttr ? true : (args.passMask & 0xffu) == 0xffu, sky = setAttr ? true : args.nSky > 0;
    for (int pass = 0; pass < (setAttr ? 6 : 1); ++pass) {     // setAttr: every instantiation once
        const bool s = setAttr ? (pass < 4) : sky, a = setAttr ? (pass & 1) : all, g = setAttr ? (pass & 2) != 0 : (gather && sky);
        if (s && a && g) launchTexT<true, true, true>(args, smem, setAttr, stream);
        else if (s && a) launchTexT<true, true, false>(args, smem, setAttr, stream);
        else if (s && g) launchTexT<true, false, true>(args, smem, setAttr, stream);
        else if (s) launchTexT<true, false, false>(args, smem, setAttr, stream);
        else if (a) launchTexT<false, true, false>(args, smem, setAttr, stream);
        else launchTexT<false, false, false>(args, smem, setAttr, stream);
    }
}
template <bool SKY, bool ALL>
static void launchPairT(const PbrArgs& args, size_t smem, bool setAttr, cudaStream_t stream) {
    (void)smem;
    if (setAttr) cudaFuncSetAttribute(k_pbr_pair<SKY, ALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtaP));
    else k_pbr_pair<SKY, ALL><<<args.nBoxes, T_WARPS * 32, sizeof(CtaP), stream>>>(args);
}
static void dispatchPair(const PbrArgs& args, size_t smem, bool setAttr, cudaStream_t stream) {
    const bool all = (args.passMask & 0xffu) == 0xffu, sky = args.nSky > 0;
    for (int pass = 0; pass < (setAttr ? 4 : 1); ++pass) {
        const bool s = setAttr ? (pass & 1) : sky, a = setAttr ? (pass & 2) != 0 : all;
        if (s && a) launchPairT<true, true>(args, smem, setAttr, stream);
        else if (s) launchPairT<true, false>(args, smem, setAttr, stream);
        else if (a) launchPairT<false, true>(args, smem, setAttr, stream);
        else launchPairT<false, false>(args, smem, setAttr, stream);
    }
}
// kernel 0 = k_pbr_pair (two rows per lane on the packed FP32 pipe), 1 = k_pbr_tex (one row per lane)
void launch_pbr_tex(const PbrArgs& args, bool skyGather, int kernel, cudaStream_t stream) {
    const size_t smem = sizeof(CtaT);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !g_texAttrSet[dev]) {  // > 48 KB of dynamic shared memory is an opt-in, per device
        dispatchTex(args, false, smem, true, stream);
        dispatchPair(args, smem, true, stream);
        g_texAttrSet[dev] = true;
    }
    if (kernel == 0 && !(skyGather && args.nSky > 0)) dispatchPair(args, smem, false, stream);
    else dispatchTex(args, skyGather, smem, false, stream);
}

}  // namespace b2
