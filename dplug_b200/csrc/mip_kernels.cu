// Mipmap down-sampling kernels (graphics/dplug/graphics/mipmap.d:676-1109), one launch per
// generateLevel call.  All are streaming kernels: every thread reads whole aligned vectors from
// two (box) or four (cubic) source rows and writes one vector of destination texels; warps cover
// contiguous row segments (bodies in mip_bodies.cuh).  Integer paths are bit-exact by construction; the
// alpha-premultiplying box is float but uses only IEEE * and + (this file is compiled with -fmad=false).
#include <cstdlib>
#include "mip_bodies.cuh"

namespace b2 {

// ------------------------------------------------------------------------------------------------
// One launch per generateLevel call: a thread block of 32 x 8 threads runs one body of mip_bodies.cuh (the source level
// is read-only for the launch: ld.global.nc).
template <bool PREMUL>
__global__ void __launch_bounds__(256) k_box_rgba(DevLevel dst, DevLevel src, Box r, f32x2 nz) {
    boxRgbaBody<PREMUL, LdNc>(dst, src, r, nz, blockIdx.x, blockIdx.y, threadIdx.x, threadIdx.y);
}
__global__ void __launch_bounds__(256) k_box_l16(DevLevel dst, DevLevel src, Box r) {
    boxL16Body<LdNc>(dst, src, r, blockIdx.x, blockIdx.y, threadIdx.x, threadIdx.y);
}
__global__ void __launch_bounds__(256) k_cubic_rgba(DevLevel dst, DevLevel src, Box r) {
    cubicRgbaBody<LdNc>(dst, src, r, blockIdx.x, blockIdx.y, threadIdx.x, threadIdx.y);
}
__global__ void __launch_bounds__(256) k_cubic_rgba4(DevLevel dst, DevLevel src, Box r) {
    cubicRgba4Body<LdNc>(dst, src, r, blockIdx.x, blockIdx.y, threadIdx.x, threadIdx.y);
}

// TMA-fed variant for LARGE rectangles (generateLevelCubicRGBA, mipmap.d:902-1040; same integer arithmetic, same bytes).
// A CTA owns a strip of TMA_STRIP destination columns and walks down TMA_ROWS destination rows.  One elected thread of a
// producer warp streams the source rows of the strip into a shared-memory ring with 1-D bulk asynchronous copies
// (cp.async.bulk.shared.global, the TMA engine: SASS UBLKCP) that complete on mbarriers; four consumer warps read a
// stage when its barrier has flipped and hand it back through a second barrier.  No load instruction, no address
// arithmetic and no scoreboard wait for global memory is left in the consumers' instruction stream, and TMA_STAGES row
// pairs per CTA are in flight instead of what one round of register loads can hold.
//
// Arithmetic per consumer thread = FOUR destination texels, horizontal filter first.  The [1 3 3 1] x [1 3 3 1] sum is
// separable and integer, so the order is free: every source row is filtered horizontally once (H = (a + d) + 3 (b + c)
// on the two 16-bit lane pairs of the RGBA texels: ten source texels give four H values), and the vertical
// [1 3 3 1] is carried as ONE partial sum per destination texel: with A..D = H of rows 2y-1 .. 2y+2,
// V(y) = (A + 3B) + 3C + D and the carry for y + 1 is C + 3D.  ~35 integer instructions per destination texel (the first
// version — two texels per thread, vertical sums per source column, then the horizontal filter — needed ~45 and was
// bound by the ALU pipe: 18 us for the 4096^2 -> 2048^2 level of the 8192^2 pyramid against 13 us of HBM time).
constexpr int TMA_TEXELS = 4;                         // destination texels per consumer thread
constexpr int TMA_STRIP = 128 * TMA_TEXELS;           // destination columns per CTA (128 consumer threads)
#ifndef B2_TMA_ROWS
#define B2_TMA_ROWS 16
#endif
#ifndef B2_TMA_STAGES
#define B2_TMA_STAGES 4
#endif
constexpr int TMA_ROWS = B2_TMA_ROWS;                 // destination rows a CTA walks
constexpr int TMA_STAGES = B2_TMA_STAGES;             // ring depth, in source-row pairs
constexpr int TMA_ROW_TEXELS = 2 * TMA_STRIP + 8;     // source texels per staged row: columns [c0, c0 + 2 * TMA_STRIP + 8), c0 = 2*x0 - 4
constexpr int TMA_ROW_BYTES = TMA_ROW_TEXELS * 4;     // a multiple of 16

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

    // ---------------- consumers: thread = four destination texels x .. x + 3 (source columns 2x-1 .. 2x+8)
    const int t = threadIdx.x;                                        // 0 .. 127
    const int x = x0 + TMA_TEXELS * t;
    const bool active = x < r.maxx && 2 * x < sw && x + TMA_TEXELS - 1 >= r.minx;
    const bool vec = active && 2 * x + 7 <= sw - 1;                   // columns 2x .. 2x+7 as two 128-bit shared loads
    const int cL = max(0, min(2 * x - 1, sw - 1)) - c0, cR = max(0, min(2 * x + 8, sw - 1)) - c0;   // outer columns, clamped to the image (mipmap.d:915-935)
    const int cM = 2 * x - c0;
    // horizontal [1 3 3 1] of one staged source row: he / ho[u] = even / odd byte lanes of destination texel x + u
    auto hrow = [&](const uint32_t* row, uint32_t (&he)[TMA_TEXELS], uint32_t (&ho)[TMA_TEXELS]) {
        uint32_t v[2 * TMA_TEXELS + 2];
        if (vec) {
            const uint4 q0 = *reinterpret_cast<const uint4*>(row + cM), q1 = *reinterpret_cast<const uint4*>(row + cM + 4);
            v[0] = row[cL]; v[1] = q0.x; v[2] = q0.y; v[3] = q0.z; v[4] = q0.w; v[5] = q1.x; v[6] = q1.y; v[7] = q1.z; v[8] = q1.w; v[9] = row[cR];
        } else {
#pragma unroll
            for (int k = 0; k < 2 * TMA_TEXELS + 2; ++k) v[k] = row[max(0, min(2 * x - 1 + k, sw - 1)) - c0];
        }
        uint32_t e[2 * TMA_TEXELS + 2], o[2 * TMA_TEXELS + 2];
#pragma unroll
        for (int k = 0; k < 2 * TMA_TEXELS + 2; ++k) { e[k] = evenBytes(v[k]); o[k] = oddBytes(v[k]); }
#pragma unroll
        for (int u = 0; u < TMA_TEXELS; ++u) {                        // columns 2(x+u)-1 .. 2(x+u)+2 = indices 2u .. 2u+3; <= 8 * 255 per lane
            he[u] = (e[2 * u] + e[2 * u + 3]) + 3u * (e[2 * u + 1] + e[2 * u + 2]);
            ho[u] = (o[2 * u] + o[2 * u + 3]) + 3u * (o[2 * u + 1] + o[2 * u + 2]);
        }
    };
    uint32_t pe[TMA_TEXELS], po[TMA_TEXELS];                         // carried vertical partial sums
    for (int k = 0; k < nPairs; ++k) {
        const int st = k % TMA_STAGES;
        mbarWait(&full[st], (k / TMA_STAGES) & 1);
        uint32_t ce[TMA_TEXELS], co[TMA_TEXELS], de[TMA_TEXELS], dO[TMA_TEXELS];
        if (active) { hrow(ring[st][0], ce, co); hrow(ring[st][1], de, dO); }
        __syncwarp();
        if (lane == 0) mbarArrive(&empty[st]);                        // this warp is done with the stage
        if (!active) continue;
        if (k == 0) {
#pragma unroll
            for (int u = 0; u < TMA_TEXELS; ++u) { pe[u] = ce[u] + 3u * de[u]; po[u] = co[u] + 3u * dO[u]; }
            continue;
        }
        const int y = yb + k - 1;
        uint32_t out[TMA_TEXELS];
#pragma unroll
        for (int u = 0; u < TMA_TEXELS; ++u) {
            const uint32_t sl = (pe[u] + 3u * ce[u]) + (de[u] + 0x00200020u);          // <= 64 * 255 + 32 per 16-bit lane
            const uint32_t sh2 = (po[u] + 3u * co[u]) + (dO[u] + 0x00200020u);
            pe[u] = ce[u] + 3u * de[u];
            po[u] = co[u] + 3u * dO[u];
            out[u] = ((sl >> 6) & 0x00ff00ffu) | ((sh2 << 2) & 0xff00ff00u);
        }
        uint32_t* dp = rowPtrW<uint32_t>(dst, y);
        if (x >= r.minx && x + TMA_TEXELS - 1 < r.maxx) *reinterpret_cast<uint4*>(dp + x) = make_uint4(out[0], out[1], out[2], out[3]);
        else {
#pragma unroll
            for (int u = 0; u < TMA_TEXELS; ++u)
                if (x + u >= r.minx && x + u < r.maxx) dp[x + u] = out[u];
        }
    }
}

__global__ void __launch_bounds__(256) k_cubic_l16(DevLevel dst, DevLevel src, Box r) {
    cubicL16Body<LdNc>(dst, src, r, blockIdx.x, blockIdx.y, threadIdx.x, threadIdx.y);
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
// encodeScreenshotAsQB — gui/dplug/gui/screencap.d:21-84: one column of 26 voxels per pixel of the rendered buffer (the
// composited pixel after colour correction and reorderComponents(pf): graphics.d:843-857), visible up to
// 10 + 16 * depth / 65536.  Voxel (x, y, z) is at ((z * 26) + y) * W + x; its colour bytes are bytes 0, 1, 2 of the
// re-ordered pixel, whatever the window's pixel format (the reference reads them through RGBA's fields).
__global__ void __launch_bounds__(256) k_screenshot_qb(uint32_t* __restrict__ vox, DevLevel src, DevLevel depth, int W, int H, int pf,
                                                      const uint8_t* __restrict__ lut) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, z = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= W || z >= H) return;
    const uint32_t p = __ldg(rowPtr<uint32_t>(src, z) + x);
    uint32_t r = p & 0xff, g = (p >> 8) & 0xff, b = (p >> 16) & 0xff;
    if (lut) { r = __ldg(lut + r); g = __ldg(lut + 256 + g); b = __ldg(lut + 512 + b); }
    uint32_t rgb;
    if (pf == 1) rgb = b | (g << 8) | (r << 16);          // BGRA8: bytes b, g, r
    else if (pf == 2) rgb = 0xffu | (r << 8) | (g << 16);  // ARGB8: bytes a (255), r, g
    else rgb = r | (g << 8) | (b << 16);
    const int d = 10 + ((16 * (int)__ldg(rowPtr<uint16_t>(depth, z) + x)) >> 16);
    uint32_t* o = vox + (size_t)z * 26 * W + x;
#pragma unroll
    for (int y = 0; y < 26; ++y) o[(size_t)y * W] = rgb | (d >= y ? 0xff000000u : 0u);
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
    {   // large rectangles: the TMA-fed strip walker, when there are enough strips of 512 x 16 texels to fill the GPU
        const int x0 = (r.minx >> 2) << 2;
        const int strips = (r.maxx - x0 + TMA_STRIP - 1) / TMA_STRIP, chunks = (r.maxy - r.miny + TMA_ROWS - 1) / TMA_ROWS;
        // measured (us, TMA / four-texel register kernel): 4096^2 source 18.9 / 23.1; 8192 x 1024 (a band of the 16K canvas, 256 CTAs)
        // 12.4 / 14.4; 2048^2 (128 CTAs) 8.6 / 8.4
        static const long long minCtas = getenv("B2_CUBIC_TMA_MIN") ? atoll(getenv("B2_CUBIC_TMA_MIN")) : 256;
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
void launch_screenshot_qb(uint32_t* vox, const DevLevel& src, const DevLevel& depth, int W, int H, int pf, const uint8_t* lut, cudaStream_t s) {
    dim3 block(32, 8);
    k_screenshot_qb<<<gridFor(W, H, block), block, 0, s>>>(vox, src, depth, W, H, pf, lut);
}
void launch_fill32(const DevLevel& win, uint32_t value, cudaStream_t s) {
    dim3 block(32, 8);
    k_fill32<<<gridFor(win.w, win.h, block), block, 0, s>>>(win, value);
}

}  // namespace b2
