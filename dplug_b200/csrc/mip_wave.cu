// Whole-pyramid wavefront kernel: every level of a Mipmap.generateMipmaps / GUIGraphics.regenerateMipmaps chain
// (graphics/dplug/graphics/mipmap.d:515-618, gui/dplug/gui/graphics.d:1378-1419) in ONE persistent launch.
//
// Why: on a large image the chain is a sequence of kernels with opposite bottlenecks — the level-1 box streams 5/6 of
// all bytes and is HBM-bound while the SM pipes idle; the cubic levels are ALU-bound while HBM idles; the small levels
// are pure launch-to-launch dependency latency — and every level is written to HBM and read back by the next launch.
// Here the levels advance together as a wavefront down the image:
//
//   * the work is a list of ITEMS, one item = a run of virtual thread blocks (mip_bodies.cuh: 32 x 8 threads, 8
//     destination rows) of one level; the host orders the list so that an item comes after everything it reads
//     (key = the last level-0 row its value depends on: level l + 1 follows level l by a few rows);
//   * a persistent grid takes items from an atomic queue head IN LIST ORDER.  Row chunks (8 destination rows of a
//     level) carry a completion counter; an item waits (ld.acquire.gpu) until the source chunks it reads are complete,
//     runs its body with L2-coherent loads (ld.global.cg — the texels were written by another SM in this launch), then
//     releases its own chunk (fence + atomic add).  Because items are claimed in order and an item only waits for
//     earlier items, the smallest unfinished item can always run: no co-residency requirement, no deadlock;
//   * level l + 1 reads level l a few microseconds after it was written: from the 126 MB L2, not from HBM.  DRAM
//     traffic drops to the algorithmic minimum (source once, every level written once) and the box / cubic / small
//     levels overlap inside one launch;
//   * the last CTA to leave zeroes the queue head and the counters, so the same plan can be launched again (and replayed
//     from a CUDA graph) without a memset in front.
//
// Bit-exactness: the bodies are the ones the per-level kernels run; only the schedule differs.
#include <algorithm>
#include <vector>
#include "mip_bodies.cuh"
#include "mip_wave.hpp"

namespace b2 {

__device__ __forceinline__ int ldAcquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

#ifndef B2_WAVE_MIN_CTAS
#define B2_WAVE_MIN_CTAS 4
#endif
#ifndef B2_WAVE_UNROLL
#define B2_WAVE_UNROLL 2      // box blocks whose loads a thread issues together
#endif
#ifndef B2_WAVE_ITEM_TEXELS
#define B2_WAVE_ITEM_TEXELS 512
#endif
constexpr int kWaveItemTexels = B2_WAVE_ITEM_TEXELS;   // destination columns of an item 

// (Measured alternative: WARPS as the workers — no block-wide barrier, a warp walks the 8 rows of its item alone.  An
// item then takes 8-16 dependent load rounds and every level of the chain adds that latency to the tail: 250 us for
// the 8192^2 pyramid against 140-155 us with CTA-wide items.)
__global__ void __launch_bounds__(256, B2_WAVE_MIN_CTAS) k_mip_wave(const __grid_constant__ WaveArgs A) {
    __shared__ int sIdx, sLast;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for (;;) {
        if (tid == 0) sIdx = atomicAdd(A.ctl, 1);
        __syncthreads();
        const int idx = sIdx;
        if (idx >= A.nItems) break;
        const int4 i0 = __ldg(reinterpret_cast<const int4*>(A.items + idx));
        const int4 i1 = __ldg(reinterpret_cast<const int4*>(A.items + idx) + 1);
        const int step = i0.x & 0xffff, kind = i0.x >> 16, by = i0.y, bx0 = i0.z, bx1 = i0.w;
        const int depLo = i1.x, depHi = i1.y, depTarget = i1.z, done = i1.w;
        // wait for the source row chunks (at most four: rows 2y-1 .. 2y+16 of the level above)
        if (tid <= depHi - depLo) {
            const int* c = A.ctl + depLo + tid;
            unsigned spins = 0;
            while (ldAcquire(c) < depTarget) {
                __nanosleep(40);
                if (++spins > (1u << 24)) __trap();   // a lost release aborts the launch instead of hanging the GPU
            }
        }
        __syncthreads();
        const DevLevel& src = A.lv[step];
        const DevLevel& dst = A.lv[step + 1];
        const Box& r = A.upd[step];
        switch (kind) {
            case WAVE_BOX_RGBA: boxRgbaRun<false, LdCg, B2_WAVE_UNROLL>(dst, src, r, A.nz, bx0, bx1, by, tx, ty); break;
            case WAVE_BOX_PREMUL: boxRgbaRun<true, LdCg, B2_WAVE_UNROLL>(dst, src, r, A.nz, bx0, bx1, by, tx, ty); break;
            case WAVE_CUBIC_RGBA:
                for (int bx = bx0; bx < bx1; ++bx) cubicRgba4Body<LdCg>(dst, src, r, bx, by, tx, ty);
                break;
            case WAVE_BOX_L16: boxL16Run<LdCg, B2_WAVE_UNROLL>(dst, src, r, bx0, bx1, by, tx, ty); break;
            default:
                for (int bx = bx0; bx < bx1; ++bx) cubicL16Body<LdCg>(dst, src, r, bx, by, tx, ty);
                break;
        }
        __syncthreads();                                  // every thread's stores happen-before the release below
        if (tid == 0) { __threadfence(); atomicAdd(A.ctl + done, 1); }
    }
    if (tid == 0) { __threadfence(); sLast = atomicAdd(A.ctl + 1, 1) == (int)gridDim.x - 1; }
    __syncthreads();
    if (sLast)                                            // every other CTA has left its loop: nobody reads the counters
        for (int i = tid; i < A.nCtl; i += blockDim.x) A.ctl[i] = 0;
}

// ------------------------------------------------------------------------------------------------ host: the plan
// Virtual-block geometry of a step (must match the bodies): texels per block along x, and the first block's origin.
static int kindOf(bool rgba, int q) {
    if (rgba) return q == 1 ? WAVE_CUBIC_RGBA : (q == 2 ? WAVE_BOX_PREMUL : WAVE_BOX_RGBA);   // B2_QUALITY_*: 0 box, 1 cubic, 2 premul box
    return q == 1 ? WAVE_CUBIC_L16 : WAVE_BOX_L16;
}
static int blockTexels(int kind) { return (kind == WAVE_CUBIC_RGBA || kind == WAVE_BOX_L16) ? 128 : 64; }
static int blocksX(int kind, const Box& r) {
    const int sh = blockTexels(kind) == 128 ? 2 : 1;      // thread = 4 / 2 texels, origin rounded down to that
    const int units = ((r.maxx + (1 << sh) - 1) >> sh) - (r.minx >> sh);
    return (units + kBodyW - 1) / kBodyW;
}

int wave_build_plan(const DevLevel* lv, const Box* upd, const int* q, int n, bool rgba, std::vector<WaveItem>& items, int* nCtl) {
    items.clear();
    struct StepInfo { int kind, nby, nItemsX, cntBase; bool empty; };
    std::vector<StepInfo> st(n);
    std::vector<std::vector<long long>> key(n);
    int cnt = 2;   // ctl[0] = queue head, ctl[1] = CTAs that have left
    struct Keyed { long long key; int step; WaveItem it; };
    std::vector<Keyed> all;
    for (int k = 0; k < n; ++k) {
        const Box& r = upd[k];
        StepInfo& s = st[k];
        s.kind = kindOf(rgba, q[k]);
        s.empty = r.maxx <= r.minx || r.maxy <= r.miny;
        s.nby = s.empty ? 0 : (r.maxy - r.miny + kBodyH - 1) / kBodyH;
        const int nbx = s.empty ? 0 : blocksX(s.kind, r);
        const int perItem = kWaveItemTexels / blockTexels(s.kind);   // an item = 8 rows x kWaveItemTexels destination columns
        s.nItemsX = (nbx + perItem - 1) / perItem;
        s.cntBase = cnt;
        cnt += s.nby;
        key[k].assign(s.nby, 0);
        const bool cubic = s.kind == WAVE_CUBIC_RGBA || s.kind == WAVE_CUBIC_L16;
        for (int by = 0; by < s.nby; ++by) {
            const int y0 = r.miny + by * kBodyH, y1 = std::min(y0 + kBodyH, r.maxy);      // destination rows [y0, y1)
            // source rows read (clamped to the image: mipmap.d:915-935)
            int lo = cubic ? 2 * y0 - 1 : 2 * y0, hi = cubic ? 2 * (y1 - 1) + 2 : 2 * (y1 - 1) + 1;
            lo = std::max(lo, 0); hi = std::min(hi, lv[k].h - 1);
            int depLo = 0, depHi = -1, target = 0;
            long long ky = (long long)hi;                                                  // first step: the last level-0 row read
            if (k > 0) {
                ky = ((long long)(hi + 1) << k) - 1;
                if (!st[k - 1].empty) {
                    const Box& rs = upd[k - 1];
                    const int a = std::max(lo, rs.miny), b = std::min(hi, rs.maxy - 1);   // rows written in this launch
                    if (a <= b) {
                        const int c0 = (a - rs.miny) / kBodyH, c1 = (b - rs.miny) / kBodyH;
                        depLo = st[k - 1].cntBase + c0; depHi = st[k - 1].cntBase + c1; target = st[k - 1].nItemsX;
                        ky = 0;
                        for (int c = c0; c <= c1; ++c) ky = std::max(ky, key[k - 1][c]);
                    }
                }
            }
            if (depHi - depLo + 1 > 4) return -1;                                          // cannot happen with 8-row chunks
            key[k][by] = ky;
            for (int ix = 0; ix < s.nItemsX; ++ix) {
                WaveItem it;
                it.stepKind = k | (s.kind << 16);
                it.by = by; it.bx0 = ix * perItem; it.bx1 = std::min(nbx, (ix + 1) * perItem);
                it.depLo = depLo; it.depHi = depHi; it.depTarget = target; it.done = s.cntBase + by;
                all.push_back(Keyed{ky, k, it});
            }
        }
    }
    // list order = (key, step): an item's sources have a key that is not larger and a smaller step
    std::stable_sort(all.begin(), all.end(), [](const Keyed& a, const Keyed& b) { return a.key != b.key ? a.key < b.key : a.step < b.step; });
    items.reserve(all.size());
    for (const Keyed& e : all) items.push_back(e.it);
    *nCtl = cnt;
    return 0;
}

int wave_max_ctas() {
    static int ctas = 0;
    if (!ctas) {
        int perSm = 0, dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_mip_wave, 256, 0);
        ctas = std::max(1, perSm) * std::max(1, sms);
    }
    return ctas;
}

void launch_mip_wave(const WaveArgs& a, cudaStream_t s) {
    const int grid = std::min(a.nItems, wave_max_ctas());
    if (grid <= 0) return;
    k_mip_wave<<<grid, 256, 0, s>>>(a);
}

}  // namespace b2
