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
                const float sc = levelScale(level);
                const float cx = fminf(fmaxf(fmaf(skyx, sc, P.skyOx[level]), P.skyOx[level] + 0.5f), P.skyOx[level] + P.skyMaxX[level] + 0.5f);
                const float cy = fminf(fmaxf(fmaf(skyy, sc, P.skyOy[level]), P.skyOy[level] + 0.5f), P.skyOy[level] + P.skyMaxY[level] + 0.5f);
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
void launch_pbr_tex(const PbrArgs& args, bool skyGather, cudaStream_t stream) {
    const size_t smem = sizeof(CtaT);
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !g_texAttrSet[dev]) {  // > 48 KB of dynamic shared memory is an opt-in, per device
        dispatchTex(args, false, smem, true, stream);
        g_texAttrSet[dev] = true;
    }
    dispatchTex(args, skyGather, smem, false, stream);
}

}  // namespace b2
