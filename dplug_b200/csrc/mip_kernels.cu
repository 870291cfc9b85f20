// Mipmap down-sampling kernels (graphics/dplug/graphics/mipmap.d:676-1109), one launch per
// generateLevel call.  All are HBM-bound streaming kernels: every thread reads whole aligned vectors from
// two (box) or four (cubic) source rows and writes one vector of destination texels; warps cover
// contiguous row segments.  Integer paths are bit-exact by construction; the alpha-premultiplying box
// is float but uses only IEEE * and + (this file is compiled with -fmad=false).
#include <cstdlib>
#include "mip_math.cuh"

namespace b2 {

// ------------------------------------------------------------------------------------------------
// Box kernels.  Thread = 2 destination texels (x even) when the pair is inside the rectangle: one 128-bit
// load per source row, one 64-bit store; ragged rectangle edges fall back to single texels.
template <bool PREMUL>
__global__ void __launch_bounds__(256) k_box_rgba(DevLevel dst, DevLevel src, Box r, f32x2 nz) {
    const int xp = (r.minx >> 1) + blockIdx.x * blockDim.x + threadIdx.x;  // pair index in absolute coordinates
    const int y = r.miny + blockIdx.y * blockDim.y + threadIdx.y;
    if (y >= r.maxy) return;
    const int x = xp * 2;
    if (x >= r.maxx || x + 1 < r.minx) return;
    const uint32_t* s0 = rowPtr<uint32_t>(src, 2 * y);
    const uint32_t* s1 = rowPtr<uint32_t>(src, 2 * y + 1);
    uint32_t* d = rowPtrW<uint32_t>(dst, y);
    if (x >= r.minx && x + 1 < r.maxx) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(s0 + 2 * x));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(s1 + 2 * x));
        uint2 o;
        if (PREMUL) o = premul4x2(a.x, a.y, b.x, b.y, a.z, a.w, b.z, b.w, nz);
        else { o.x = box4(a.x, a.y, b.x, b.y); o.y = box4(a.z, a.w, b.z, b.w); }
        *reinterpret_cast<uint2*>(d + x) = o;
    } else {
        for (int xx = max(x, r.minx); xx < min(x + 2, r.maxx); ++xx) {
            const uint2 a = __ldg(reinterpret_cast<const uint2*>(s0 + 2 * xx));
            const uint2 b = __ldg(reinterpret_cast<const uint2*>(s1 + 2 * xx));
            d[xx] = PREMUL ? premul4(a.x, a.y, b.x, b.y) : box4(a.x, a.y, b.x, b.y);
        }
    }
}

// generateLevelBoxL16 — mipmap.d:764-794.  Thread = 4 destination texels: 128-bit load per source row
// (8 x u16), 64-bit store.
__global__ void __launch_bounds__(256) k_box_l16(DevLevel dst, DevLevel src, Box r) {
    const int xq = (r.minx >> 2) + blockIdx.x * blockDim.x + threadIdx.x;  // quad index, absolute
    const int y = r.miny + blockIdx.y * blockDim.y + threadIdx.y;
    if (y >= r.maxy) return;
    const int x = xq * 4;
    if (x >= r.maxx || x + 3 < r.minx) return;
    const uint16_t* s0 = rowPtr<uint16_t>(src, 2 * y);
    const uint16_t* s1 = rowPtr<uint16_t>(src, 2 * y + 1);
    uint16_t* d = rowPtrW<uint16_t>(dst, y);
    if (x >= r.minx && x + 3 < r.maxx) {
        const uint4 a = __ldg(reinterpret_cast<const uint4*>(s0 + 2 * x));
        const uint4 b = __ldg(reinterpret_cast<const uint4*>(s1 + 2 * x));
        uint2 o;
        o.x = boxL16pair(a.x, a.y, b.x, b.y);
        o.y = boxL16pair(a.z, a.w, b.z, b.w);
        *reinterpret_cast<uint2*>(d + x) = o;
    } else {
        for (int xx = max(x, r.minx); xx < min(x + 4, r.maxx); ++xx) {
            uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(s0 + 2 * xx));
            uint32_t b = __ldg(reinterpret_cast<const uint32_t*>(s1 + 2 * xx));
            d[xx] = (uint16_t)(((a & 0xffffu) + (a >> 16) + (b & 0xffffu) + (b >> 16) + 2u) >> 2);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Cubic kernels — generateLevelCubicRGBA / L16, mipmap.d:902-1109: [1 3 3 1] x [1 3 3 1], (sum+32)>>6,
// outer taps clamped to the image.  The filter is separable and integer, so summing rows first is exact.
//
// Thread = 2 destination texels (x even).  Source columns needed: 2x-1 .. 2x+4 (6 texels): the aligned
// 128-bit vector 2x..2x+3 is loaded by the thread itself; column 2x-1 is the last texel of the left
// neighbour lane's vector and column 2x+4 the first texel of the right neighbour's (warp shuffles); only
// lanes on a warp / image edge issue an extra scalar load.

__global__ void __launch_bounds__(256) k_cubic_rgba(DevLevel dst, DevLevel src, Box r) {
    const int xp = (r.minx >> 1) + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = r.miny + blockIdx.y * blockDim.y + threadIdx.y;  // uniform per warp (blockDim.x == 32)
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
        uint4 a = __ldg(reinterpret_cast<const uint4*>(rm + 2 * x));
        uint4 b = __ldg(reinterpret_cast<const uint4*>(r0 + 2 * x));
        uint4 c = __ldg(reinterpret_cast<const uint4*>(r1 + 2 * x));
        uint4 d = __ldg(reinterpret_cast<const uint4*>(r2 + 2 * x));
        vsumRGBA(a.x, b.x, c.x, d.x, lo[1], hi[1]);
        vsumRGBA(a.y, b.y, c.y, d.y, lo[2], hi[2]);
        vsumRGBA(a.z, b.z, c.z, d.z, lo[3], hi[3]);
        vsumRGBA(a.w, b.w, c.w, d.w, lo[4], hi[4]);
    } else if (active) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int cx = min(2 * x + k, sw - 1);
            vsumRGBA(__ldg(rm + cx), __ldg(r0 + cx), __ldg(r1 + cx), __ldg(r2 + cx), lo[1 + k], hi[1 + k]);
        }
    } else {
#pragma unroll
        for (int k = 1; k <= 4; ++k) lo[k] = hi[k] = 0;
    }
    // neighbours: column 2x-1 = left lane's column index 4, column 2x+4 = right lane's column index 1
    const unsigned full = 0xffffffffu;
    uint32_t lLo = __shfl_up_sync(full, lo[4], 1), lHi = __shfl_up_sync(full, hi[4], 1);
    uint32_t rLo = __shfl_down_sync(full, lo[1], 1), rHi = __shfl_down_sync(full, hi[1], 1);
    const bool leftOk = __shfl_up_sync(full, (int)vec, 1) && threadIdx.x > 0;
    const bool rightOk = __shfl_down_sync(full, (int)active, 1) && threadIdx.x < 31;
    if (!active) return;
    if (leftOk) { lo[0] = lLo; hi[0] = lHi; }
    else {
        int cx = max(2 * x - 1, 0);
        vsumRGBA(__ldg(rm + cx), __ldg(r0 + cx), __ldg(r1 + cx), __ldg(r2 + cx), lo[0], hi[0]);
    }
    if (rightOk && 2 * x + 4 < sw) { lo[5] = rLo; hi[5] = rHi; }
    else {
        int cx = min(2 * x + 4, sw - 1);
        vsumRGBA(__ldg(rm + cx), __ldg(r0 + cx), __ldg(r1 + cx), __ldg(r2 + cx), lo[5], hi[5]);
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
// texels instead of once per two.  Same integer arithmetic, same bytes.
__global__ void __launch_bounds__(256) k_cubic_rgba4(DevLevel dst, DevLevel src, Box r) {
    const int xq = (r.minx >> 2) + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = r.miny + blockIdx.y * blockDim.y + threadIdx.y;  // uniform per warp (blockDim.x == 32)
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
            a[k] = __ldg(reinterpret_cast<const uint4*>(rows[k] + 2 * x));
            b[k] = __ldg(reinterpret_cast<const uint4*>(rows[k] + 2 * x + 4));
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
            vsumRGBA(__ldg(rows[0] + cx), __ldg(rows[1] + cx), __ldg(rows[2] + cx), __ldg(rows[3] + cx), lo[1 + k], hi[1 + k]);
        }
    } else {
#pragma unroll
        for (int k = 1; k <= 8; ++k) lo[k] = hi[k] = 0;
    }
    // neighbours: column 2x-1 = left lane's column index 8, column 2x+8 = right lane's column index 1
    const unsigned full = 0xffffffffu;
    const uint32_t lLo = __shfl_up_sync(full, lo[8], 1), lHi = __shfl_up_sync(full, hi[8], 1);
    const uint32_t rLo = __shfl_down_sync(full, lo[1], 1), rHi = __shfl_down_sync(full, hi[1], 1);
    const bool leftOk = __shfl_up_sync(full, (int)vec, 1) && threadIdx.x > 0;
    const bool rightOk = __shfl_down_sync(full, (int)active, 1) && threadIdx.x < 31;
    if (!active) return;
    if (leftOk) { lo[0] = lLo; hi[0] = lHi; }
    else {
        const int cx = max(2 * x - 1, 0);
        vsumRGBA(__ldg(rows[0] + cx), __ldg(rows[1] + cx), __ldg(rows[2] + cx), __ldg(rows[3] + cx), lo[0], hi[0]);
    }
    if (rightOk && 2 * x + 8 < sw) { lo[9] = rLo; hi[9] = rHi; }
    else {
        const int cx = min(2 * x + 8, sw - 1);
        vsumRGBA(__ldg(rows[0] + cx), __ldg(rows[1] + cx), __ldg(rows[2] + cx), __ldg(rows[3] + cx), lo[9], hi[9]);
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

// TMA-fed variant for LARGE rectangles (generateLevelCubicRGBA, mipmap.d:902-1040; same integer arithmetic, same bytes).
// A CTA owns a strip of TMA_STRIP destination columns and walks down TMA_ROWS destination rows.  One elected thread of a
// producer warp streams the source rows of the strip into a shared-memory ring with 1-D bulk asynchronous copies
// (cp.async.bulk.shared.global, the TMA engine: SASS UBLKCP) that complete on mbarriers; four consumer warps read a
// stage when its barrier has flipped and hand it back through a second barrier.  No load instruction, no address
// arithmetic and no scoreboard wait for global memory is left in the consumers' instruction stream, and TMA_STAGES row
// pairs (33 KB) per CTA are in flight instead of what one round of register loads can hold.  The vertical [1 3 3 1] is
// carried as one partial sum per column: with A..D = rows 2y-1 .. 2y+2 split into even / odd byte lanes,
// V(y) = (A + 3B) + 3C + D and the carry for y + 1 is C + 3D, so every source row is read once.
// Measured on the 8192^2 pyramid (4096^2 -> 2048^2, 84 MB): 18 us against 21.2 us for k_cubic_rgba4; a register-only
// walking variant (loads by the consumers themselves, one row pair ahead) reached the same 18 us — at ~45 integer
// instructions per destination texel (two 16-bit lanes per 32-bit operation) the level is bound by the ALU pipe, not by
// HBM (13 us) — and a deeper register prefetch (163 registers) fell to 25.7 us.
constexpr int TMA_STRIP = 256;                        // destination columns per CTA (128 consumer threads x 2 texels)
#ifndef B2_TMA_ROWS
#define B2_TMA_ROWS 16
#endif
#ifndef B2_TMA_STAGES
#define B2_TMA_STAGES 4
#endif
constexpr int TMA_ROWS = B2_TMA_ROWS;                 // destination rows a CTA walks
constexpr int TMA_STAGES = B2_TMA_STAGES;             // ring depth, in source-row pairs
constexpr int TMA_ROW_TEXELS = 2 * TMA_STRIP + 8;     // source texels per staged row: columns [c0, c0 + 520), c0 = 2*x0 - 4
constexpr int TMA_ROW_BYTES = TMA_ROW_TEXELS * 4;     // 2080, a multiple of 16

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count));
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smemAddr(bar)), "r"(parity) : "memory");
        if (spins > (1u << 26)) __trap();             // a lost completion aborts the launch instead of hanging the GPU
    }
}
__device__ __forceinline__ void bulkLoad(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smemAddr(smemDst)), "l"(gmemSrc), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}

__global__ void __launch_bounds__(160) k_cubic_rgba_tma(DevLevel dst, DevLevel src, Box r) {
    extern __shared__ __align__(128) unsigned char tmaSmem[];
    uint32_t (*ring)[2][TMA_ROW_TEXELS] = reinterpret_cast<uint32_t (*)[2][TMA_ROW_TEXELS]>(tmaSmem);
    __shared__ __align__(8) uint64_t full[TMA_STAGES], empty[TMA_STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sw = src.w, sh = src.h;
    const int x0 = ((r.minx >> 2) << 2) + blockIdx.x * TMA_STRIP;     // first destination column of the strip
    const int yb = r.miny + blockIdx.y * TMA_ROWS, ye = min(yb + TMA_ROWS, r.maxy);
    const int c0 = max(2 * x0 - 4, 0);                                // first staged source column (16-byte aligned)
    const int nPairs = (ye - yb) + 1;                                 // pair 0 = rows (2yb-1, 2yb), pair k = (2(yb+k-1)+1, +2)
    if (threadIdx.x == 0) {
        for (int k = 0; k < TMA_STAGES; ++k) { mbarInit(&full[k], 1); mbarInit(&empty[k], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 4) {
        // ---------------- producer: one elected lane streams the row pairs of the strip into the ring
        if (lane == 0) {
            // bytes of a staged row: up to the end of the source row (whole 16-byte units)
            const uint32_t bytes = (uint32_t)min(TMA_ROW_BYTES, ((sw - c0) * 4 + 15) & ~15);
            for (int k = 0; k < nPairs; ++k) {
                const int st = k % TMA_STAGES;
                if (k >= TMA_STAGES) mbarWait(&empty[st], ((k / TMA_STAGES) - 1) & 1);
                const int ra = k == 0 ? 2 * yb - 1 : 2 * (yb + k - 1) + 1, rb = ra + 1;
                mbarExpectTx(&full[st], 2 * bytes);
                bulkLoad(&ring[st][0][0], rowPtr<uint32_t>(src, max(0, min(ra, sh - 1))) + c0, bytes, &full[st]);
                bulkLoad(&ring[st][1][0], rowPtr<uint32_t>(src, max(0, min(rb, sh - 1))) + c0, bytes, &full[st]);
            }
        }
        return;
    }

    // ---------------- consumers: thread = two destination texels x, x + 1 (source columns 2x-1 .. 2x+4)
    const int t = threadIdx.x;                                        // 0 .. 127
    const int x = x0 + 2 * t;
    const bool active = x < r.maxx && 2 * x < sw && x + 1 >= r.minx;
    // staged-row indices of the six source columns (clamped to the image: mipmap.d:915-935)
    int ci[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) ci[k] = max(0, min(2 * x - 1 + k, sw - 1)) - c0;
    const bool vec = active && 2 * x + 3 <= sw - 1;                   // columns 2x .. 2x+3 as one 128-bit shared load
    auto readRow = [&](const uint32_t* row, uint32_t (&e)[6], uint32_t (&o)[6]) {
        uint32_t v[6];
        if (vec) {
            const uint4 q = *reinterpret_cast<const uint4*>(row + ci[1]);
            v[0] = row[ci[0]]; v[1] = q.x; v[2] = q.y; v[3] = q.z; v[4] = q.w; v[5] = row[ci[5]];
        } else {
#pragma unroll
            for (int k = 0; k < 6; ++k) v[k] = row[ci[k]];
        }
#pragma unroll
        for (int k = 0; k < 6; ++k) { e[k] = evenBytes(v[k]); o[k] = oddBytes(v[k]); }
    };
    uint32_t pe[6], po[6];
    for (int k = 0; k < nPairs; ++k) {
        const int st = k % TMA_STAGES;
        mbarWait(&full[st], (k / TMA_STAGES) & 1);
        uint32_t ce[6], co[6], de[6], dO[6];
        if (active) { readRow(ring[st][0], ce, co); readRow(ring[st][1], de, dO); }
        __syncwarp();
        if (lane == 0) mbarArrive(&empty[st]);                        // this warp is done with the stage
        if (!active) continue;
        if (k == 0) {
#pragma unroll
            for (int c = 0; c < 6; ++c) { pe[c] = ce[c] + 3u * de[c]; po[c] = co[c] + 3u * dO[c]; }
            continue;
        }
        const int y = yb + k - 1;
        uint32_t ve[6], vo[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            ve[c] = pe[c] + 3u * ce[c] + de[c];                       // <= 8 * 255 per 16-bit lane
            vo[c] = po[c] + 3u * co[c] + dO[c];
            pe[c] = ce[c] + 3u * de[c];
            po[c] = co[c] + 3u * dO[c];
        }
        uint32_t o[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            // destination x+u: columns 2(x+u)-1 .. 2(x+u)+2 = indices 2u .. 2u+3 (already clamped through ci[])
            const uint32_t sl = (ve[2 * u] + ve[2 * u + 3]) + 3u * (ve[2 * u + 1] + ve[2 * u + 2]) + 0x00200020u;
            const uint32_t sh2 = (vo[2 * u] + vo[2 * u + 3]) + 3u * (vo[2 * u + 1] + vo[2 * u + 2]) + 0x00200020u;
            o[u] = ((sl >> 6) & 0x00ff00ffu) | (((sh2 >> 6) & 0x00ff00ffu) << 8);
        }
        uint32_t* dp = rowPtrW<uint32_t>(dst, y);
        if (x >= r.minx && x + 1 < r.maxx) *reinterpret_cast<uint2*>(dp + x) = make_uint2(o[0], o[1]);
        else {
            if (x >= r.minx && x < r.maxx) dp[x] = o[0];
            if (x + 1 >= r.minx && x + 1 < r.maxx) dp[x + 1] = o[1];
        }
    }
}

// L16 variant: thread = 2 destination texels; vertical sums are 32-bit per column (<= 8*65535).
__global__ void __launch_bounds__(256) k_cubic_l16(DevLevel dst, DevLevel src, Box r) {
    const int xp = (r.minx >> 1) + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = r.miny + blockIdx.y * blockDim.y + threadIdx.y;
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
        uint2 a = __ldg(reinterpret_cast<const uint2*>(rm + 2 * x));
        uint2 b = __ldg(reinterpret_cast<const uint2*>(r0 + 2 * x));
        uint2 c = __ldg(reinterpret_cast<const uint2*>(r1 + 2 * x));
        uint2 d = __ldg(reinterpret_cast<const uint2*>(r2 + 2 * x));
        v[1] = (a.x & 0xffffu) + 3u * (b.x & 0xffffu) + 3u * (c.x & 0xffffu) + (d.x & 0xffffu);
        v[2] = (a.x >> 16) + 3u * (b.x >> 16) + 3u * (c.x >> 16) + (d.x >> 16);
        v[3] = (a.y & 0xffffu) + 3u * (b.y & 0xffffu) + 3u * (c.y & 0xffffu) + (d.y & 0xffffu);
        v[4] = (a.y >> 16) + 3u * (b.y >> 16) + 3u * (c.y >> 16) + (d.y >> 16);
    } else if (active) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int cx = min(2 * x + k, sw - 1);
            v[1 + k] = (uint32_t)__ldg(rm + cx) + 3u * __ldg(r0 + cx) + 3u * __ldg(r1 + cx) + __ldg(r2 + cx);
        }
    } else {
#pragma unroll
        for (int k = 1; k <= 4; ++k) v[k] = 0;
    }
    const unsigned full = 0xffffffffu;
    uint32_t lv = __shfl_up_sync(full, v[4], 1), rv = __shfl_down_sync(full, v[1], 1);
    const bool leftOk = __shfl_up_sync(full, (int)vec, 1) && threadIdx.x > 0;
    const bool rightOk = __shfl_down_sync(full, (int)active, 1) && threadIdx.x < 31;
    if (!active) return;
    if (leftOk) v[0] = lv;
    else {
        int cx = max(2 * x - 1, 0);
        v[0] = (uint32_t)__ldg(rm + cx) + 3u * __ldg(r0 + cx) + 3u * __ldg(r1 + cx) + __ldg(r2 + cx);
    }
    if (rightOk && 2 * x + 4 < sw) v[5] = rv;
    else {
        int cx = min(2 * x + 4, sw - 1);
        v[5] = (uint32_t)__ldg(rm + cx) + 3u * __ldg(r0 + cx) + 3u * __ldg(r1 + cx) + __ldg(r2 + cx);
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

// ------------------------------------------------------------------------------------------------
// SimpleRawCompositor — gui/dplug/gui/compositor.d:280-296: copy diffuse RGB, alpha = 255.
__global__ void __launch_bounds__(256) k_copy_diffuse(DevLevel dst, DevLevel src, const Box* boxes, int nBoxes) {
    for (int b = blockIdx.x; b < nBoxes; b += gridDim.x) {
        const Box box = boxes[b];
        for (int j = box.miny + threadIdx.y; j < box.maxy; j += blockDim.y)
            for (int i = box.minx + threadIdx.x; i < box.maxx; i += blockDim.x)
                rowPtrW<uint32_t>(dst, j)[i] = __ldg(rowPtr<uint32_t>(src, j) + i) | 0xff000000u;
    }
}

// reorderComponents — gui/dplug/gui/graphics.d:1425-1452,1527-1654: RGBA -> RGBA/BGRA/ARGB, alpha = 255,
// written to a staging image that is then copied to the host.
__global__ void __launch_bounds__(256) k_swizzle(DevLevel dst, DevLevel src, Box box, int pf) {
    const int i = box.minx + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = box.miny + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= box.maxx || j >= box.maxy) return;
    uint32_t p = __ldg(rowPtr<uint32_t>(src, j) + i);
    uint32_t r = p & 0xff, g = (p >> 8) & 0xff, b = (p >> 16) & 0xff, o;
    if (pf == 1) o = b | (g << 8) | (r << 16) | 0xff000000u;        // BGRA8
    else if (pf == 2) o = 0xffu | (r << 8) | (g << 16) | (b << 24);  // ARGB8
    else o = p | 0xff000000u;                                       // RGBA8
    rowPtrW<uint32_t>(dst, j)[i] = o;
}

// doDraw stages E-H fused (gui/dplug/gui/graphics.d:825-868): the composited pixel goes through the colour-correction
// transfer tables (applyColorCorrection, pbrwidgets/dplug/pbrwidgets/colorcorrection.d:157-170), is re-ordered to
// the window's pixel format with alpha = 255 (reorderComponents, graphics.d:1425-1452) and lands in the window
// buffer at the user-area offset (resizeContent, graphics.d:1465-1511).
__global__ void __launch_bounds__(256) k_present(DevLevel win, DevLevel src, Box box, int dx, int dy, int pf, const uint8_t* __restrict__ lut) {
    const int i = box.minx + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = box.miny + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= box.maxx || j >= box.maxy) return;
    const uint32_t p = __ldg(rowPtr<uint32_t>(src, j) + i);
    uint32_t r = p & 0xff, g = (p >> 8) & 0xff, b = (p >> 16) & 0xff;
    if (lut) { r = __ldg(lut + r); g = __ldg(lut + 256 + g); b = __ldg(lut + 512 + b); }
    uint32_t o;
    if (pf == 1) o = b | (g << 8) | (r << 16) | 0xff000000u;        // BGRA8
    else if (pf == 2) o = 0xffu | (r << 8) | (g << 16) | (b << 24);  // ARGB8
    else o = r | (g << 8) | (b << 16) | 0xff000000u;                // RGBA8
    rowPtrW<uint32_t>(win, j + dy)[i + dx] = o;
}
__global__ void __launch_bounds__(256) k_fill32(DevLevel win, uint32_t value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < win.w && j < win.h) rowPtrW<uint32_t>(win, j)[i] = value;
}

// ------------------------------------------------------------------------------------------------
// Level-0 producers that are plain per-pixel work: a widget's cached PBR render is
// written into the three maps for its dirty rectangles, either copied (PBRBackgroundGUI.onDrawPBR,
// pbrwidgets/dplug/pbrwidgets/pbrbackgroundgui.d:147-162) or blended through per-map 8-bit opacity masks
// (UIBufferedElementPBR.onDrawPBR, gui/dplug/gui/bufferedelement.d:118-131; blendWithAlpha,
// graphics/dplug/graphics/draw.d:933-969; blendColor, graphics/dplug/graphics/color.d:453-526).
__device__ __forceinline__ uint32_t blendRGBA(uint32_t fg, uint32_t bg, uint32_t a) {
    // per channel (fg*alpha + bg*(255-alpha)) / 255 with the reference's multiply-shift division (color.d:465-475),
    // exact for operands up to 255*255
    const uint32_t ia = (~a) & 0xffu;
    uint32_t o = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t v = ((fg >> (8 * k)) & 0xffu) * a + ((bg >> (8 * k)) & 0xffu) * ia;
        o |= ((v * 32897u) >> 23) << (8 * k);
    }
    return o;
}
__device__ __forceinline__ uint32_t blendL16(uint32_t fg, uint32_t bg, uint32_t a8) {
    // color.d:522-526 with alpha16 = (alpha << 8) | alpha (draw.d:961): the products are 32-bit `int` in the
    // reference and DO wrap for bright texels with high opacity; the wrap and the signed division are kept
    const uint32_t a16 = (a8 << 8) | a8, ia16 = (~a16) & 0xffffu;
    const int32_t sum = (int32_t)(fg * a16 + bg * ia16);
    return (uint32_t)(sum / 65535) & 0xffffu;
}

template <bool BLEND>
__global__ void __launch_bounds__(256) k_layer(DevLevel dD, DevLevel dZ, DevLevel dM, DevLevel sD, DevLevel sZ, DevLevel sM,
                                              DevLevel oD, DevLevel oZ, DevLevel oM, Box box, int px, int py) {
    const int i = box.minx + blockIdx.x * blockDim.x + threadIdx.x;   // widget coordinates
    const int j = box.miny + blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= box.maxx || j >= box.maxy) return;
    const int X = i + px, Y = j + py;                                 // UI coordinates
    uint32_t* pd = rowPtrW<uint32_t>(dD, Y) + X;
    uint16_t* pz = rowPtrW<uint16_t>(dZ, Y) + X;
    uint32_t* pm = rowPtrW<uint32_t>(dM, Y) + X;
    const uint32_t fd = __ldg(rowPtr<uint32_t>(sD, j) + i);
    const uint32_t fz = __ldg(rowPtr<uint16_t>(sZ, j) + i);
    const uint32_t fm = __ldg(rowPtr<uint32_t>(sM, j) + i);
    if (!BLEND) { *pd = fd; *pz = (uint16_t)fz; *pm = fm; return; }
    const uint32_t ad = __ldg(rowPtr<uint8_t>(oD, j) + i), az = __ldg(rowPtr<uint8_t>(oZ, j) + i), am = __ldg(rowPtr<uint8_t>(oM, j) + i);
    if (ad) *pd = blendRGBA(fd, *pd, ad);          // alpha == 0: the destination is left alone (draw.d:952-954)
    if (az) *pz = (uint16_t)blendL16(fz, *pz, az);
    if (am) *pm = blendRGBA(fm, *pm, am);
}

// ------------------------------------------------------------------------------------------------
// Pitched copy executed by the SMs.  Used for SMALL transfers between page-locked host images (read / written in place
// over PCIe through their mapped addresses) and the device maps: a DMA copy has ~10 us of fixed cost and walks pitched
// images row by row, which for a 620 x 330 UI (2 MB up in three images, 0.8 MB down) is most of the host call.
__global__ void __launch_bounds__(256) k_copy2d(uint8_t* __restrict__ dst, size_t dpitch, const uint8_t* __restrict__ src, size_t spitch,
                                               int rowVecs, int rows, int vecBytes) {
    const int total = rowVecs * rows;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int r = idx / rowVecs, v = idx - r * rowVecs;
        const uint8_t* s = src + (size_t)r * spitch + (size_t)v * vecBytes;
        uint8_t* d = dst + (size_t)r * dpitch + (size_t)v * vecBytes;
        if (vecBytes == 16) *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
        else if (vecBytes == 4) *reinterpret_cast<uint32_t*>(d) = *reinterpret_cast<const uint32_t*>(s);
        else *reinterpret_cast<uint16_t*>(d) = *reinterpret_cast<const uint16_t*>(s);
    }
}
void launch_copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t rowBytes, int rows, cudaStream_t s) {
    const uintptr_t all = (uintptr_t)dst | (uintptr_t)src | dpitch | spitch | rowBytes;
    const int vec = (all % 16 == 0) ? 16 : (all % 4 == 0) ? 4 : 2;   // every pixel type here is at least 2 bytes wide
    const int rowVecs = (int)(rowBytes / vec);
    const long long total = (long long)rowVecs * rows;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) return;
    k_copy2d<<<blocks, 256, 0, s>>>((uint8_t*)dst, dpitch, (const uint8_t*)src, spitch, rowVecs, rows, vec);
}

static inline dim3 gridFor(int nx, int ny, dim3 block) { return dim3((nx + block.x - 1) / block.x, (ny + block.y - 1) / block.y); }

void launch_box_rgba(const DevLevel& dst, const DevLevel& src, Box r, bool premul, cudaStream_t s) {
    dim3 block(32, 8);
    int pairs = ((r.maxx + 1) >> 1) - (r.minx >> 1);
    dim3 grid = gridFor(pairs, r.maxy - r.miny, block);
    if (premul) k_box_rgba<true><<<grid, block, 0, s>>>(dst, src, r, kNegZero2);
    else k_box_rgba<false><<<grid, block, 0, s>>>(dst, src, r, kNegZero2);
}
void launch_box_l16(const DevLevel& dst, const DevLevel& src, Box r, cudaStream_t s) {
    dim3 block(32, 8);
    int quads = ((r.maxx + 3) >> 2) - (r.minx >> 2);
    k_box_l16<<<gridFor(quads, r.maxy - r.miny, block), block, 0, s>>>(dst, src, r);
}
void launch_cubic_rgba(const DevLevel& dst, const DevLevel& src, Box r, cudaStream_t s) {
    static const int tmaOn = getenv("B2_CUBIC_TMA") ? atoi(getenv("B2_CUBIC_TMA")) : 1;      // tuning aid (0: never, 2: always — tests)
    {   // large rectangles: the TMA-fed strip walker, when its CTAs fill the GPU
        const int x0 = (r.minx >> 2) << 2;
        const int strips = (r.maxx - x0 + TMA_STRIP - 1) / TMA_STRIP, chunks = (r.maxy - r.miny + TMA_ROWS - 1) / TMA_ROWS;
        static const long long minCtas = getenv("B2_CUBIC_TMA_MIN") ? atoll(getenv("B2_CUBIC_TMA_MIN")) : 148 * 2;
        if (tmaOn && ((long long)strips * chunks >= minCtas || tmaOn == 2) && r.maxx > r.minx && r.maxy > r.miny) {
            static bool attr = false;
            const size_t smem = sizeof(uint32_t) * TMA_STAGES * 2 * TMA_ROW_TEXELS;
            if (!attr) { cudaFuncSetAttribute(k_cubic_rgba_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
            k_cubic_rgba_tma<<<dim3(strips, chunks), 160, smem, s>>>(dst, src, r);
            return;
        }
    }
    dim3 block(32, 8);
    if (r.maxx - r.minx >= 256) {   // wide rectangles: four texels per thread
        int quads = ((r.maxx + 3) >> 2) - (r.minx >> 2);
        k_cubic_rgba4<<<gridFor(quads, r.maxy - r.miny, block), block, 0, s>>>(dst, src, r);
        return;
    }
    int pairs = ((r.maxx + 1) >> 1) - (r.minx >> 1);
    k_cubic_rgba<<<gridFor(pairs, r.maxy - r.miny, block), block, 0, s>>>(dst, src, r);
}
void launch_cubic_l16(const DevLevel& dst, const DevLevel& src, Box r, cudaStream_t s) {
    dim3 block(32, 8);
    int pairs = ((r.maxx + 1) >> 1) - (r.minx >> 1);
    k_cubic_l16<<<gridFor(pairs, r.maxy - r.miny, block), block, 0, s>>>(dst, src, r);
}
void launch_copy_diffuse(const DevLevel& dst, const DevLevel& src, const Box* boxes, int n, int grid, cudaStream_t s) {
    k_copy_diffuse<<<grid, dim3(32, 8), 0, s>>>(dst, src, boxes, n);
}
void launch_swizzle(const DevLevel& dst, const DevLevel& src, Box box, int pf, cudaStream_t s) {
    dim3 block(32, 8);
    k_swizzle<<<gridFor(box.maxx - box.minx, box.maxy - box.miny, block), block, 0, s>>>(dst, src, box, pf);
}

void launch_layer(bool blend, const DevLevel* dst3, const DevLevel* src3, const DevLevel* op3, Box box, int px, int py, cudaStream_t s) {
    dim3 block(32, 8);
    dim3 grid = gridFor(box.maxx - box.minx, box.maxy - box.miny, block);
    if (blend) k_layer<true><<<grid, block, 0, s>>>(dst3[0], dst3[1], dst3[2], src3[0], src3[1], src3[2], op3[0], op3[1], op3[2], box, px, py);
    else k_layer<false><<<grid, block, 0, s>>>(dst3[0], dst3[1], dst3[2], src3[0], src3[1], src3[2], src3[0], src3[1], src3[2], box, px, py);
}
void launch_present(const DevLevel& win, const DevLevel& src, Box box, int dx, int dy, int pf, const uint8_t* lut, cudaStream_t s) {
    dim3 block(32, 8);
    k_present<<<gridFor(box.maxx - box.minx, box.maxy - box.miny, block), block, 0, s>>>(win, src, box, dx, dy, pf, lut);
}
void launch_fill32(const DevLevel& win, uint32_t value, cudaStream_t s) {
    dim3 block(32, 8);
    k_fill32<<<gridFor(win.w, win.h, block), block, 0, s>>>(win, value);
}

}  // namespace b2
