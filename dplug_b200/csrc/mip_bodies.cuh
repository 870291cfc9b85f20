// Bodies of the per-level mipmap down-samplers (graphics/dplug/graphics/mipmap.d:676-1109), shared by the
// one-launch-per-level kernels (mip_kernels.cu) and the whole-pyramid wavefront kernel (mip_wave.cu).
//
// A body does the work of ONE virtual thread block of 32 x 8 threads: `bx`, `by` are the block coordinates inside the
// update rectangle's grid, `tx` (0..31, the lane) and `ty` (0..7, the warp) the thread coordinates.  A warp owns one
// destination row segment; every lane of the warp must enter the body (the cubic bodies exchange their outer columns
// with warp shuffles).  `LD` selects the load instruction: LdNc (ld.global.nc, __ldg) when the source level is
// read-only for the whole launch, LdCg (ld.global.cg, L2 only) when another CTA of the SAME launch produced it
// (mip_wave.cu), where the non-coherent path must not be used.
//
// Integer paths are bit-exact by construction; the alpha-premultiplying box is float but uses only single IEEE
// multiplies and adds (the including files are compiled with -fmad=false).
#pragma once
#include "mip_math.cuh"

namespace b2 {

struct LdNc { template <typename T> static __device__ __forceinline__ T ld(const T* p) { return __ldg(p); } };
struct LdCg { template <typename T> static __device__ __forceinline__ T ld(const T* p) { return __ldcg(p); } };

constexpr int kBodyW = 32, kBodyH = 8;   // threads of a virtual block

// ------------------------------------------------------------------------------------------------
// Box.  Thread = 2 destination texels (x even) when the pair is inside the rectangle: one 128-bit load per source
// row, one 64-bit store; ragged rectangle edges fall back to single texels.  Virtual block = 64 x 8 texels.
template <bool PREMUL, class LD>
__device__ __forceinline__ void boxRgbaBody(const DevLevel& dst, const DevLevel& src, const Box& r, f32x2 nz, int bx, int by, int tx, int ty) {
    const int xp = (r.minx >> 1) + bx * kBodyW + tx;  // pair index in absolute coordinates
    const int y = r.miny + by * kBodyH + ty;
    if (y >= r.maxy) return;
    const int x = xp * 2;
    if (x >= r.maxx || x + 1 < r.minx) return;
    const uint32_t* s0 = rowPtr<uint32_t>(src, 2 * y);
    const uint32_t* s1 = rowPtr<uint32_t>(src, 2 * y + 1);
    uint32_t* d = rowPtrW<uint32_t>(dst, y);
    if (x >= r.minx && x + 1 < r.maxx) {
        const uint4 a = LD::ld(reinterpret_cast<const uint4*>(s0 + 2 * x));
        const uint4 b = LD::ld(reinterpret_cast<const uint4*>(s1 + 2 * x));
        uint2 o;
        if (PREMUL) o = premul4x2(a.x, a.y, b.x, b.y, a.z, a.w, b.z, b.w, nz);
        else { o.x = box4(a.x, a.y, b.x, b.y); o.y = box4(a.z, a.w, b.z, b.w); }
        *reinterpret_cast<uint2*>(d + x) = o;
    } else {
        for (int xx = max(x, r.minx); xx < min(x + 2, r.maxx); ++xx) {
            const uint2 a = LD::ld(reinterpret_cast<const uint2*>(s0 + 2 * xx));
            const uint2 b = LD::ld(reinterpret_cast<const uint2*>(s1 + 2 * xx));
            d[xx] = PREMUL ? premul4(a.x, a.y, b.x, b.y) : box4(a.x, a.y, b.x, b.y);
        }
    }
}

// generateLevelBoxL16 — mipmap.d:764-794.  Thread = 4 destination texels: 128-bit load per source row
// (8 x u16), 64-bit store.  Virtual block = 128 x 8 texels.
template <class LD>
__device__ __forceinline__ void boxL16Body(const DevLevel& dst, const DevLevel& src, const Box& r, int bx, int by, int tx, int ty) {
    const int xq = (r.minx >> 2) + bx * kBodyW + tx;  // quad index, absolute
    const int y = r.miny + by * kBodyH + ty;
    if (y >= r.maxy) return;
    const int x = xq * 4;
    if (x >= r.maxx || x + 3 < r.minx) return;
    const uint16_t* s0 = rowPtr<uint16_t>(src, 2 * y);
    const uint16_t* s1 = rowPtr<uint16_t>(src, 2 * y + 1);
    uint16_t* d = rowPtrW<uint16_t>(dst, y);
    if (x >= r.minx && x + 3 < r.maxx) {
        const uint4 a = LD::ld(reinterpret_cast<const uint4*>(s0 + 2 * x));
        const uint4 b = LD::ld(reinterpret_cast<const uint4*>(s1 + 2 * x));
        uint2 o;
        o.x = boxL16pair(a.x, a.y, b.x, b.y);
        o.y = boxL16pair(a.z, a.w, b.z, b.w);
        *reinterpret_cast<uint2*>(d + x) = o;
    } else {
        for (int xx = max(x, r.minx); xx < min(x + 4, r.maxx); ++xx) {
            uint32_t a = LD::ld(reinterpret_cast<const uint32_t*>(s0 + 2 * xx));
            uint32_t b = LD::ld(reinterpret_cast<const uint32_t*>(s1 + 2 * xx));
            d[xx] = (uint16_t)(((a & 0xffffu) + (a >> 16) + (b & 0xffffu) + (b >> 16) + 2u) >> 2);
        }
    }
}

// A run of box blocks [bx0, bx1) of one block row, for the persistent kernel: the 128-bit loads of U blocks are issued
// together before the first result is computed (the generic body's edge branches keep the compiler from overlapping
// the loads of consecutive calls), so a thread keeps 2U x 16 bytes in flight.  Threads on a ragged rectangle edge take
// the generic body.
template <bool PREMUL, class LD, int U>
__device__ __forceinline__ void boxRgbaRun(const DevLevel& dst, const DevLevel& src, const Box& r, f32x2 nz, int bx0, int bx1, int by, int tx, int ty) {
    const int y = r.miny + by * kBodyH + ty;
    if (y >= r.maxy) return;
    const uint32_t* s0 = rowPtr<uint32_t>(src, 2 * y);
    const uint32_t* s1 = rowPtr<uint32_t>(src, 2 * y + 1);
    uint32_t* d = rowPtrW<uint32_t>(dst, y);
    for (int bx = bx0; bx < bx1; bx += U) {
        uint4 a[U], b[U];
        bool vec[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int x = 2 * ((r.minx >> 1) + (bx + u) * kBodyW + tx);
            vec[u] = bx + u < bx1 && x >= r.minx && x + 1 < r.maxx;
            if (vec[u]) {
                a[u] = LD::ld(reinterpret_cast<const uint4*>(s0 + 2 * x));
                b[u] = LD::ld(reinterpret_cast<const uint4*>(s1 + 2 * x));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int x = 2 * ((r.minx >> 1) + (bx + u) * kBodyW + tx);
            if (vec[u]) {
                uint2 o;
                if (PREMUL) o = premul4x2(a[u].x, a[u].y, b[u].x, b[u].y, a[u].z, a[u].w, b[u].z, b[u].w, nz);
                else { o.x = box4(a[u].x, a[u].y, b[u].x, b[u].y); o.y = box4(a[u].z, a[u].w, b[u].z, b[u].w); }
                *reinterpret_cast<uint2*>(d + x) = o;
            } else if (bx + u < bx1) boxRgbaBody<PREMUL, LD>(dst, src, r, nz, bx + u, by, tx, ty);
        }
    }
}
template <class LD, int U>
__device__ __forceinline__ void boxL16Run(const DevLevel& dst, const DevLevel& src, const Box& r, int bx0, int bx1, int by, int tx, int ty) {
    const int y = r.miny + by * kBodyH + ty;
    if (y >= r.maxy) return;
    const uint16_t* s0 = rowPtr<uint16_t>(src, 2 * y);
    const uint16_t* s1 = rowPtr<uint16_t>(src, 2 * y + 1);
    uint16_t* d = rowPtrW<uint16_t>(dst, y);
    for (int bx = bx0; bx < bx1; bx += U) {
        uint4 a[U], b[U];
        bool vec[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int x = 4 * ((r.minx >> 2) + (bx + u) * kBodyW + tx);
            vec[u] = bx + u < bx1 && x >= r.minx && x + 3 < r.maxx;
            if (vec[u]) {
                a[u] = LD::ld(reinterpret_cast<const uint4*>(s0 + 2 * x));
                b[u] = LD::ld(reinterpret_cast<const uint4*>(s1 + 2 * x));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int x = 4 * ((r.minx >> 2) + (bx + u) * kBodyW + tx);
            if (vec[u]) {
                uint2 o;
                o.x = boxL16pair(a[u].x, a[u].y, b[u].x, b[u].y);
                o.y = boxL16pair(a[u].z, a[u].w, b[u].z, b[u].w);
                *reinterpret_cast<uint2*>(d + x) = o;
            } else if (bx + u < bx1) boxL16Body<LD>(dst, src, r, bx + u, by, tx, ty);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Cubic — generateLevelCubicRGBA / L16, mipmap.d:902-1109: [1 3 3 1] x [1 3 3 1], (sum+32)>>6, outer taps clamped
// to the image.  The filter is separable and integer, so summing rows first is exact.
//
// Thread = 2 destination texels (x even).  Source columns needed: 2x-1 .. 2x+4 (6 texels): the aligned 128-bit vector
// 2x..2x+3 is loaded by the thread itself; column 2x-1 is the last texel of the left neighbour lane's vector and column
// 2x+4 the first texel of the right neighbour's (warp shuffles); only lanes on a warp / image edge issue an extra
// scalar load.  Virtual block = 64 x 8 texels.
template <class LD>
__device__ __forceinline__ void cubicRgbaBody(const DevLevel& dst, const DevLevel& src, const Box& r, int bx, int by, int tx, int ty) {
    const int xp = (r.minx >> 1) + bx * kBodyW + tx;
    const int y = r.miny + by * kBodyH + ty;  // uniform per warp
    if (y >= r.maxy) return;
    const int x = xp * 2;
    const int sw = src.w, sh = src.h;
    const int ym = max(2 * y - 1, 0), yp = min(2 * y + 2, sh - 1);
    const uint32_t* rm = rowPtr<uint32_t>(src, ym);
    const uint32_t* r0 = rowPtr<uint32_t>(src, 2 * y);
    const uint32_t* r1 = rowPtr<uint32_t>(src, 2 * y + 1);
    const uint32_t* r2 = rowPtr<uint32_t>(src, yp);
    // vertical pass on this lane's 4 aligned source columns 2x .. 2x+3 (clamped loads past the image edge)
    uint32_t lo[6], hi[6];
    const bool active = x < r.maxx && x + 1 >= r.minx && 2 * x < sw;
    const bool vec = active && (2 * x + 3 < sw);
    if (vec) {
        uint4 a = LD::ld(reinterpret_cast<const uint4*>(rm + 2 * x));
        uint4 b = LD::ld(reinterpret_cast<const uint4*>(r0 + 2 * x));
        uint4 c = LD::ld(reinterpret_cast<const uint4*>(r1 + 2 * x));
        uint4 d = LD::ld(reinterpret_cast<const uint4*>(r2 + 2 * x));
        vsumRGBA(a.x, b.x, c.x, d.x, lo[1], hi[1]);
        vsumRGBA(a.y, b.y, c.y, d.y, lo[2], hi[2]);
        vsumRGBA(a.z, b.z, c.z, d.z, lo[3], hi[3]);
        vsumRGBA(a.w, b.w, c.w, d.w, lo[4], hi[4]);
    } else if (active) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int cx = min(2 * x + k, sw - 1);
            vsumRGBA(LD::ld(rm + cx), LD::ld(r0 + cx), LD::ld(r1 + cx), LD::ld(r2 + cx), lo[1 + k], hi[1 + k]);
        }
    } else {
#pragma unroll
        for (int k = 1; k <= 4; ++k) lo[k] = hi[k] = 0;
    }
    // neighbours: column 2x-1 = left lane's column index 4, column 2x+4 = right lane's column index 1
    const unsigned full = 0xffffffffu;
    uint32_t lLo = __shfl_up_sync(full, lo[4], 1), lHi = __shfl_up_sync(full, hi[4], 1);
    uint32_t rLo = __shfl_down_sync(full, lo[1], 1), rHi = __shfl_down_sync(full, hi[1], 1);
    const bool leftOk = __shfl_up_sync(full, (int)vec, 1) && tx > 0;
    const bool rightOk = __shfl_down_sync(full, (int)active, 1) && tx < 31;
    if (!active) return;
    if (leftOk) { lo[0] = lLo; hi[0] = lHi; }
    else {
        int cx = max(2 * x - 1, 0);
        vsumRGBA(LD::ld(rm + cx), LD::ld(r0 + cx), LD::ld(r1 + cx), LD::ld(r2 + cx), lo[0], hi[0]);
    }
    if (rightOk && 2 * x + 4 < sw) { lo[5] = rLo; hi[5] = rHi; }
    else {
        int cx = min(2 * x + 4, sw - 1);
        vsumRGBA(LD::ld(rm + cx), LD::ld(r0 + cx), LD::ld(r1 + cx), LD::ld(r2 + cx), lo[5], hi[5]);
    }
    uint32_t* d = rowPtrW<uint32_t>(dst, y);
    uint32_t o[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        // destination x+t: columns 2(x+t)-1 .. 2(x+t)+2 = local indices 2t .. 2t+3; the right-most tap is
        // clamped to the last source column (mipmap.d:934-935)
        uint32_t l3 = lo[2 * t + 3], h3 = hi[2 * t + 3];
        if (2 * (x + t) + 2 > sw - 1) { l3 = lo[2 * t + 2]; h3 = hi[2 * t + 2]; }
        uint32_t sl = lo[2 * t] + 3u * lo[2 * t + 1] + 3u * lo[2 * t + 2] + l3 + 0x00200020u;
        uint32_t sh2 = hi[2 * t] + 3u * hi[2 * t + 1] + 3u * hi[2 * t + 2] + h3 + 0x00200020u;
        o[t] = ((sl >> 6) & 0x00ff00ffu) | (((sh2 >> 6) & 0x00ff00ffu) << 8);
    }
    if (x >= r.minx && x + 1 < r.maxx) *reinterpret_cast<uint2*>(d + x) = make_uint2(o[0], o[1]);
    else {
        if (x >= r.minx && x < r.maxx) d[x] = o[0];
        if (x + 1 >= r.minx && x + 1 < r.maxx) d[x + 1] = o[1];
    }
}

// Wide variant for large rectangles: thread = 4 destination texels (x a multiple of 4), two 128-bit loads per source
// row; the addressing, the edge predicates and the two neighbour columns (warp shuffles) are paid once per four
// texels instead of once per two.  Same integer arithmetic, same bytes.  Virtual block = 128 x 8 texels.
template <class LD>
__device__ __forceinline__ void cubicRgba4Body(const DevLevel& dst, const DevLevel& src, const Box& r, int bx, int by, int tx, int ty) {
    const int xq = (r.minx >> 2) + bx * kBodyW + tx;
    const int y = r.miny + by * kBodyH + ty;  // uniform per warp
    if (y >= r.maxy) return;
    const int x = xq * 4;
    const int sw = src.w, sh = src.h;
    const uint32_t* rows[4] = {rowPtr<uint32_t>(src, max(2 * y - 1, 0)), rowPtr<uint32_t>(src, 2 * y), rowPtr<uint32_t>(src, 2 * y + 1),
                               rowPtr<uint32_t>(src, min(2 * y + 2, sh - 1))};
    uint32_t lo[10], hi[10];   // columns 2x-1 .. 2x+8
    const bool active = x < r.maxx && x + 3 >= r.minx && 2 * x < sw;
    const bool vec = active && (2 * x + 7 < sw);
    if (vec) {
        uint4 a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a[k] = LD::ld(reinterpret_cast<const uint4*>(rows[k] + 2 * x));
            b[k] = LD::ld(reinterpret_cast<const uint4*>(rows[k] + 2 * x + 4));
        }
        vsumRGBA(a[0].x, a[1].x, a[2].x, a[3].x, lo[1], hi[1]);
        vsumRGBA(a[0].y, a[1].y, a[2].y, a[3].y, lo[2], hi[2]);
        vsumRGBA(a[0].z, a[1].z, a[2].z, a[3].z, lo[3], hi[3]);
        vsumRGBA(a[0].w, a[1].w, a[2].w, a[3].w, lo[4], hi[4]);
        vsumRGBA(b[0].x, b[1].x, b[2].x, b[3].x, lo[5], hi[5]);
        vsumRGBA(b[0].y, b[1].y, b[2].y, b[3].y, lo[6], hi[6]);
        vsumRGBA(b[0].z, b[1].z, b[2].z, b[3].z, lo[7], hi[7]);
        vsumRGBA(b[0].w, b[1].w, b[2].w, b[3].w, lo[8], hi[8]);
    } else if (active) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int cx = min(2 * x + k, sw - 1);
            vsumRGBA(LD::ld(rows[0] + cx), LD::ld(rows[1] + cx), LD::ld(rows[2] + cx), LD::ld(rows[3] + cx), lo[1 + k], hi[1 + k]);
        }
    } else {
#pragma unroll
        for (int k = 1; k <= 8; ++k) lo[k] = hi[k] = 0;
    }
    // neighbours: column 2x-1 = left lane's column index 8, column 2x+8 = right lane's column index 1
    const unsigned full = 0xffffffffu;
    const uint32_t lLo = __shfl_up_sync(full, lo[8], 1), lHi = __shfl_up_sync(full, hi[8], 1);
    const uint32_t rLo = __shfl_down_sync(full, lo[1], 1), rHi = __shfl_down_sync(full, hi[1], 1);
    const bool leftOk = __shfl_up_sync(full, (int)vec, 1) && tx > 0;
    const bool rightOk = __shfl_down_sync(full, (int)active, 1) && tx < 31;
    if (!active) return;
    if (leftOk) { lo[0] = lLo; hi[0] = lHi; }
    else {
        const int cx = max(2 * x - 1, 0);
        vsumRGBA(LD::ld(rows[0] + cx), LD::ld(rows[1] + cx), LD::ld(rows[2] + cx), LD::ld(rows[3] + cx), lo[0], hi[0]);
    }
    if (rightOk && 2 * x + 8 < sw) { lo[9] = rLo; hi[9] = rHi; }
    else {
        const int cx = min(2 * x + 8, sw - 1);
        vsumRGBA(LD::ld(rows[0] + cx), LD::ld(rows[1] + cx), LD::ld(rows[2] + cx), LD::ld(rows[3] + cx), lo[9], hi[9]);
    }
    uint32_t o[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        // destination x+t: columns 2(x+t)-1 .. 2(x+t)+2 = local indices 2t .. 2t+3; the right-most tap is clamped to
        // the last source column (mipmap.d:934-935)
        uint32_t l3 = lo[2 * t + 3], h3 = hi[2 * t + 3];
        if (2 * (x + t) + 2 > sw - 1) { l3 = lo[2 * t + 2]; h3 = hi[2 * t + 2]; }
        const uint32_t sl = (lo[2 * t] + l3) + 3u * (lo[2 * t + 1] + lo[2 * t + 2]) + 0x00200020u;
        const uint32_t sh2 = (hi[2 * t] + h3) + 3u * (hi[2 * t + 1] + hi[2 * t + 2]) + 0x00200020u;
        o[t] = ((sl >> 6) & 0x00ff00ffu) | (((sh2 >> 6) & 0x00ff00ffu) << 8);
    }
    uint32_t* d = rowPtrW<uint32_t>(dst, y);
    if (x >= r.minx && x + 3 < r.maxx) *reinterpret_cast<uint4*>(d + x) = make_uint4(o[0], o[1], o[2], o[3]);
    else {
#pragma unroll
        for (int t = 0; t < 4; ++t)
            if (x + t >= r.minx && x + t < r.maxx) d[x + t] = o[t];
    }
}

// L16 variant: thread = 2 destination texels; vertical sums are 32-bit per column (<= 8*65535).  Virtual block =
// 64 x 8 texels.
template <class LD>
__device__ __forceinline__ void cubicL16Body(const DevLevel& dst, const DevLevel& src, const Box& r, int bx, int by, int tx, int ty) {
    const int xp = (r.minx >> 1) + bx * kBodyW + tx;
    const int y = r.miny + by * kBodyH + ty;
    if (y >= r.maxy) return;
    const int x = xp * 2;
    const int sw = src.w, sh = src.h;
    const int ym = max(2 * y - 1, 0), yp = min(2 * y + 2, sh - 1);
    const uint16_t* rm = rowPtr<uint16_t>(src, ym);
    const uint16_t* r0 = rowPtr<uint16_t>(src, 2 * y);
    const uint16_t* r1 = rowPtr<uint16_t>(src, 2 * y + 1);
    const uint16_t* r2 = rowPtr<uint16_t>(src, yp);
    uint32_t v[6];
    const bool active = x < r.maxx && x + 1 >= r.minx && 2 * x < sw;
    const bool vec = active && (2 * x + 3 < sw);
    if (vec) {
        uint2 a = LD::ld(reinterpret_cast<const uint2*>(rm + 2 * x));
        uint2 b = LD::ld(reinterpret_cast<const uint2*>(r0 + 2 * x));
        uint2 c = LD::ld(reinterpret_cast<const uint2*>(r1 + 2 * x));
        uint2 d = LD::ld(reinterpret_cast<const uint2*>(r2 + 2 * x));
        v[1] = (a.x & 0xffffu) + 3u * (b.x & 0xffffu) + 3u * (c.x & 0xffffu) + (d.x & 0xffffu);
        v[2] = (a.x >> 16) + 3u * (b.x >> 16) + 3u * (c.x >> 16) + (d.x >> 16);
        v[3] = (a.y & 0xffffu) + 3u * (b.y & 0xffffu) + 3u * (c.y & 0xffffu) + (d.y & 0xffffu);
        v[4] = (a.y >> 16) + 3u * (b.y >> 16) + 3u * (c.y >> 16) + (d.y >> 16);
    } else if (active) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int cx = min(2 * x + k, sw - 1);
            v[1 + k] = (uint32_t)LD::ld(rm + cx) + 3u * LD::ld(r0 + cx) + 3u * LD::ld(r1 + cx) + LD::ld(r2 + cx);
        }
    } else {
#pragma unroll
        for (int k = 1; k <= 4; ++k) v[k] = 0;
    }
    const unsigned full = 0xffffffffu;
    uint32_t lv = __shfl_up_sync(full, v[4], 1), rv = __shfl_down_sync(full, v[1], 1);
    const bool leftOk = __shfl_up_sync(full, (int)vec, 1) && tx > 0;
    const bool rightOk = __shfl_down_sync(full, (int)active, 1) && tx < 31;
    if (!active) return;
    if (leftOk) v[0] = lv;
    else {
        int cx = max(2 * x - 1, 0);
        v[0] = (uint32_t)LD::ld(rm + cx) + 3u * LD::ld(r0 + cx) + 3u * LD::ld(r1 + cx) + LD::ld(r2 + cx);
    }
    if (rightOk && 2 * x + 4 < sw) v[5] = rv;
    else {
        int cx = min(2 * x + 4, sw - 1);
        v[5] = (uint32_t)LD::ld(rm + cx) + 3u * LD::ld(r0 + cx) + 3u * LD::ld(r1 + cx) + LD::ld(r2 + cx);
    }
    uint16_t* d = rowPtrW<uint16_t>(dst, y);
    uint32_t o[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        uint32_t t3 = v[2 * t + 3];
        if (2 * (x + t) + 2 > sw - 1) t3 = v[2 * t + 2];
        o[t] = (v[2 * t] + 3u * v[2 * t + 1] + 3u * v[2 * t + 2] + t3 + 32u) >> 6;
    }
    if (x >= r.minx && x + 1 < r.maxx) *reinterpret_cast<uint32_t*>(d + x) = o[0] | (o[1] << 16);
    else {
        if (x >= r.minx && x < r.maxx) d[x] = (uint16_t)o[0];
        if (x + 1 >= r.minx && x + 1 < r.maxx) d[x + 1] = (uint16_t)o[1];
    }
}

}  // namespace b2
