// Arithmetic shared by the mipmap kernels (mip_kernels.cu: one launch per generateLevel call; mip_fused.cu: up to
// three levels per launch).  Integer paths are bit-exact by construction; the alpha-premultiplying box is float but
// uses only single IEEE multiplies and adds (the including files are compiled with -fmad=false).
#pragma once
#include "common.cuh"

namespace b2 {

// ------------------------------------------------------------------------------------------------
// SIMD-in-register helpers: an RGBA8 texel is split into two 16-bit lanes pairs (r,b) and (g,a) so four
// channels are summed with two integer adds; sums up to 64*255 fit the 16-bit lanes.
__device__ __forceinline__ uint32_t evenBytes(uint32_t v) { return v & 0x00ff00ffu; }
__device__ __forceinline__ uint32_t oddBytes(uint32_t v) { return __byte_perm(v, 0u, 0x4341); }   // (v >> 8) & 0x00ff00ff in one PRMT

// generateLevelBoxRGBA — mipmap.d:676-762: (a+b+c+d+2)>>2 per channel
__device__ __forceinline__ uint32_t box4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t lo = evenBytes(a) + evenBytes(b) + evenBytes(c) + evenBytes(d) + 0x00020002u;
    uint32_t hi = oddBytes(a) + oddBytes(b) + oddBytes(c) + oddBytes(d) + 0x00020002u;
    return ((lo >> 2) & 0x00ff00ffu) | (((hi >> 2) & 0x00ff00ffu) << 8);
}

// generateLevelBoxAlphaCovIntoPremulRGBA — mipmap.d:861-896 (non-legacy branch).
//
// The reference's arithmetic is a chain of single IEEE multiplies and adds per channel (no FMA).  Blackwell's packed
// FP32 instructions (FMUL2 / FADD2: two independent fp32 lanes per instruction) halve the instruction count and each
// lane is still ONE correctly rounded operation.
//
// HAZARD (measured, CUDA 12.9 ptxas): a packed multiply whose result feeds a packed add is CONTRACTED into FFMA2 —
// with explicit .rn on both, with -fmad=false, and also when the two are spelled fma(a,b,-0) / fma(x,1,c) with literal
// constants — which skips the product's rounding (1 texel in ~10^4 lands on the other side of a truncation).  Scalar
// mul.rn / add.rn are never contracted.  A product that is consumed by an addition is therefore computed as
// `mulToAdd2` = fma(a, b, nz) where nz = (-0, -0) arrives as a KERNEL ARGUMENT: ptxas cannot fold an addend it does
// not know, the instruction stays one FFMA2 whose result is round(a*b + -0) = round(a*b), and there is no instruction
// that could absorb the following add.  tests/test_host_logic.py checks the SASS of the mip objects: every FFMA2
// left must have that register addend, and there is no scalar FFMA.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mulToAdd2(f32x2 a, f32x2 b, f32x2 nz) {   // lane-wise product that may feed an add2
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(nz));
    return r;
}
constexpr unsigned long long kNegZero2 = 0x8000000080000000ull;   // what the host passes for `nz`

// one 2x2 quad, scalar (ragged rectangle edges and the per-texel paths of the fused kernel)
__device__ __forceinline__ void premulLinear(uint32_t p, float o[4]) {   // convert_gammaspace_to_linear_premul, mipmap.d:869-878
    const float inv = 1.0f / 255;
    // byte -> float without I2F: (0x4B000000 | b) is the float 2^23 + b; the subtraction is exact
    float c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] = __fadd_rn(__uint_as_float(__byte_perm(p, 0x4B000000u, 0x7440 + k)), -8388608.0f);
    o[3] = __fmul_rn(c[3], inv);                                                         // res.a = col.a * inv_255
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(c[k], inv), c[k]), inv), o[3]);
}
__device__ __forceinline__ uint32_t premul4(uint32_t A, uint32_t B, uint32_t C, uint32_t D) {
    float a[4], b[4], c[4], d[4];
    premulLinear(A, a); premulLinear(B, b); premulLinear(C, c); premulLinear(D, d);
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float m = __fadd_rn(__fadd_rn(__fadd_rn(a[k], b[k]), c[k]), d[k]);        // ((A + B) + C) + D, mipmap.d:886-889
        const float v = __fadd_rn(__fmul_rn(__fmul_rn(m, 0.25f), 255.0f), 0.5f);        // mean * 0.25f * 255.0f + 0.5f, mipmap.d:891-894
        out |= ((uint32_t)__float2int_rz(v) & 0xffu) << (8 * k);
    }
    return out;
}

// Two 2x2 quads at once: lane lo = quad 0, lane hi = quad 1.  Same single rounded IEEE operations per lane and the
// same order as premul4, so the bytes are identical.
struct Lin4x2 { f32x2 r, g, b, a; };
__device__ __forceinline__ Lin4x2 premulLinear2(uint32_t p0, uint32_t p1, f32x2 nz) {
    const float invf = 1.0f / 255;
    const f32x2 inv = pack2(invf, invf);
    const f32x2 magic = pack2(-8388608.0f, -8388608.0f);
    f32x2 c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        c[k] = add2(pack2(__uint_as_float(__byte_perm(p0, 0x4B000000u, 0x7440 + k)), __uint_as_float(__byte_perm(p1, 0x4B000000u, 0x7440 + k))), magic);
    Lin4x2 o;
    o.a = mulToAdd2(c[3], inv, nz);                                             // res.a = col.a * inv_255 (summed below)
    o.r = mulToAdd2(mul2(mul2(mul2(c[0], inv), c[0]), inv), o.a, nz);           // (((c * inv) * c) * inv) * res.a
    o.g = mulToAdd2(mul2(mul2(mul2(c[1], inv), c[1]), inv), o.a, nz);
    o.b = mulToAdd2(mul2(mul2(mul2(c[2], inv), c[2]), inv), o.a, nz);
    return o;
}
__device__ __forceinline__ uint2 premul4x2(uint32_t A0, uint32_t B0, uint32_t C0, uint32_t D0,
                                           uint32_t A1, uint32_t B1, uint32_t C1, uint32_t D1, f32x2 nz) {
    // alpha (the emissive channel) is 0 on most of a plug-in UI: every texel of both quads then contributes +0 to every
    // channel, the means are +0 and ubyte(0 * 0.25f * 255.0f + 0.5f) is 0 — the result is exactly 0x00000000
    if ((((A0 | B0) | (C0 | D0) | (A1 | B1) | (C1 | D1)) & 0xff000000u) == 0u) return make_uint2(0u, 0u);
    const Lin4x2 a = premulLinear2(A0, A1, nz), b = premulLinear2(B0, B1, nz), c = premulLinear2(C0, C1, nz), d = premulLinear2(D0, D1, nz);
    const f32x2 q = pack2(0.25f, 0.25f), s255 = pack2(255.0f, 255.0f), h = pack2(0.5f, 0.5f);
    f32x2 v[4];
    v[0] = add2(add2(add2(a.r, b.r), c.r), d.r);                       // ((A + B) + C) + D, mipmap.d:886-889
    v[1] = add2(add2(add2(a.g, b.g), c.g), d.g);
    v[2] = add2(add2(add2(a.b, b.b), c.b), d.b);
    v[3] = add2(add2(add2(a.a, b.a), c.a), d.a);
    uint2 o = make_uint2(0u, 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const f32x2 t = add2(mulToAdd2(mul2(v[k], q), s255, nz), h);      // mean * 0.25f * 255.0f + 0.5f, mipmap.d:891-894
        float t0, t1;
        unpack2(t, t0, t1);
        o.x |= ((uint32_t)__float2int_rz(t0) & 0xffu) << (8 * k);
        o.y |= ((uint32_t)__float2int_rz(t1) & 0xffu) << (8 * k);
    }
    return o;
}

// generateLevelCubicRGBA, mipmap.d:902-1040: vertical [1 3 3 1] of one source column, on the two 16-bit lane pairs
__device__ __forceinline__ void vsumRGBA(uint32_t m1, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t& lo, uint32_t& hi) {
    lo = (evenBytes(m1) + evenBytes(p2)) + 3u * (evenBytes(p0) + evenBytes(p1));  // <= 8*255 per lane
    hi = (oddBytes(m1) + oddBytes(p2)) + 3u * (oddBytes(p0) + oddBytes(p1));
}

// generateLevelBoxL16, mipmap.d:764-794: two destination texels from two 32-bit words (2 x u16) per source row
__device__ __forceinline__ uint32_t boxL16pair(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    uint32_t lo = ((a0 & 0xffffu) + (a0 >> 16) + (b0 & 0xffffu) + (b0 >> 16) + 2u) >> 2;
    uint32_t hi = ((a1 & 0xffffu) + (a1 >> 16) + (b1 & 0xffffu) + (b1 >> 16) + 2u) >> 2;
    return lo | (hi << 16);
}

}  // namespace b2
