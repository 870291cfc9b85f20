// Fused mipmap chain: up to THREE consecutive generateLevel calls (graphics/dplug/graphics/mipmap.d:534-618) in one
// launch.  The level-by-level chain is bandwidth-wasteful on the GPU (every intermediate level is written to HBM and
// read back by the next launch, and the small top levels are pure launch latency); here a CTA owns one tile of the
// LAST level of the group, derives the regions of the intermediate levels that tile depends on (the [1 3 3 1] cubic
// reaches one texel beyond its 2x2 children on each side), produces them in shared memory and writes to HBM only the
// texels it owns.  The source level is read once (+ the halo a neighbouring CTA re-reads through L2).
//
// Semantics are EXACTLY those of the sequential chain on the same update rectangles: a texel of an intermediate
// level is recomputed when it lies inside that level's update rectangle and is otherwise taken from the stored level
// (what the next generateLevel call of the chain would read), so the result does not depend on the pyramid having
// been consistent outside the rectangles.  All arithmetic is the per-level kernels' (mip_math.cuh): integer paths
// bit-exact, the alpha-premultiplying box in single IEEE operations (this file is compiled with -fmad=false).
#include <cstdio>
#include "mip_math.cuh"

namespace b2 {

constexpr int Q_CUBIC = 1, Q_PREMUL = 2;   // B2_QUALITY_CUBIC, B2_QUALITY_BOX_ALPHACOV_PREMUL of include/b2pbr.h (0 = box)
constexpr int FUSED_THREADS = 256;

struct Iv { int a, b; };   // half-open interval of texel indices
__device__ __forceinline__ bool inIv(Iv r, int v) { return v >= r.a && v < r.b; }
__device__ __forceinline__ Iv ivUnion(Iv p, Iv q) { return Iv{min(p.a, q.a), max(p.b, q.b)}; }
__device__ __forceinline__ Iv ivClip(Iv r, int lo, int hi) { const int a = max(r.a, lo), b = min(r.b, hi); return Iv{a, max(a, b)}; }
// texels of the finer level owned by the CTA that owns `up` of the coarser one: the 2x2 children, plus the odd last
// column / row of the finer level (it has no parent, mipmap.d:119-125) for the tile touching the edge
__device__ __forceinline__ Iv ivChildren(Iv up, int dimUp, int dimDown) { return Iv{2 * up.a, up.b == dimUp ? dimDown : 2 * up.b}; }
// source texels read to produce `r` (mipmap.d:683-700 box: 2x, 2x+1; mipmap.d:906-935 cubic: 2x-1 .. 2x+2, clamped)
__device__ __forceinline__ Iv ivNeed(int q, Iv r, int dimSrc) {
    if (q == Q_CUBIC) return Iv{max(2 * r.a - 1, 0), min(2 * r.b + 1, dimSrc)};
    return Iv{2 * r.a, 2 * r.b};
}
__device__ __forceinline__ bool inBox(const Box& b, int x, int y) { return x >= b.minx && x < b.maxx && y >= b.miny && y < b.maxy; }

// a rectangle of one level held in shared memory: texel (x, y) of the level is p[(y - y0) * stride + (x - x0)]
template <typename T>
struct Region {
    T* p; int x0, y0, stride;
#ifdef B2_MIP_DEBUG   // bounds-checked build (python dplug_b200/build.py with B2_MIP_DEBUG=1): traps on any access outside the buffer
    int rowsCap;
    __device__ __forceinline__ T& at(int x, int y) const {
        if (x < x0 || x - x0 >= stride || y < y0 || y - y0 >= rowsCap) { printf("mip_fused: (%d,%d) outside region x0=%d y0=%d %dx%d\n", x, y, x0, y0, stride, rowsCap); __trap(); }
        return p[(y - y0) * stride + (x - x0)];
    }
#else
    __device__ __forceinline__ T& at(int x, int y) const { return p[(y - y0) * stride + (x - x0)]; }
#endif
};
#ifdef B2_MIP_DEBUG
#define B2_REGION(T, buf, x0, y0, stride, rows) Region<T>{buf, x0, y0, stride, rows}
#else
#define B2_REGION(T, buf, x0, y0, stride, rows) Region<T>{buf, x0, y0, stride}
#endif
template <typename T>
struct GlobalSrc {
    const DevLevel& L;
    __device__ __forceinline__ uint32_t operator()(int x, int y) const { return rowPtr<T>(L, y)[x]; }
};
template <typename T>
struct SharedSrc {
    const Region<T>& R;
    __device__ __forceinline__ uint32_t operator()(int x, int y) const { return R.at(x, y); }
};

// one destination texel of any quality from any source; (sw, sh) = source level size
template <typename T, typename F>
__device__ __forceinline__ uint32_t downTexel(int q, int x, int y, int sw, int sh, F fetch) {
    if (q == Q_CUBIC) {
        const int xs[4] = {max(2 * x - 1, 0), 2 * x, 2 * x + 1, min(2 * x + 2, sw - 1)};
        const int ym = max(2 * y - 1, 0), yp = min(2 * y + 2, sh - 1);
        if (sizeof(T) == 4) {
            uint32_t lo[4], hi[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) vsumRGBA(fetch(xs[k], ym), fetch(xs[k], 2 * y), fetch(xs[k], 2 * y + 1), fetch(xs[k], yp), lo[k], hi[k]);
            const uint32_t sl = lo[0] + 3u * lo[1] + 3u * lo[2] + lo[3] + 0x00200020u;
            const uint32_t sh2 = hi[0] + 3u * hi[1] + 3u * hi[2] + hi[3] + 0x00200020u;
            return ((sl >> 6) & 0x00ff00ffu) | (((sh2 >> 6) & 0x00ff00ffu) << 8);
        } else {
            uint32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = fetch(xs[k], ym) + 3u * fetch(xs[k], 2 * y) + 3u * fetch(xs[k], 2 * y + 1) + fetch(xs[k], yp);
            return (v[0] + 3u * v[1] + 3u * v[2] + v[3] + 32u) >> 6;
        }
    }
    const uint32_t a = fetch(2 * x, 2 * y), b = fetch(2 * x + 1, 2 * y), c = fetch(2 * x, 2 * y + 1), d = fetch(2 * x + 1, 2 * y + 1);
    if (sizeof(T) == 4) return q == Q_PREMUL ? premul4(a, b, c, d) : box4(a, b, c, d);
    return (a + b + c + d + 2u) >> 2;   // generateLevelBoxL16, mipmap.d:764-794
}

__device__ __forceinline__ uint32_t fastDiv(uint32_t v, uint32_t d, uint32_t magic) { return d == 1u ? v : __umulhi(v, magic); }

// ---- step 1: global source level -> level s+1 (shared region `out` when another step follows, HBM for owned texels)
// One group = G adjacent destination texels = one 128-bit load per source row.
template <typename T>
__device__ __forceinline__ uint2 downVec(int q, const uint4& a, const uint4& b, f32x2 nz) {
    uint2 o;
    if (sizeof(T) == 4) {
        if (q == Q_PREMUL) o = premul4x2(a.x, a.y, b.x, b.y, a.z, a.w, b.z, b.w, nz);
        else { o.x = box4(a.x, a.y, b.x, b.y); o.y = box4(a.z, a.w, b.z, b.w); }
    } else {
        o.x = boxL16pair(a.x, a.y, b.x, b.y);
        o.y = boxL16pair(a.z, a.w, b.z, b.w);
    }
    return o;
}
template <typename T>
__device__ __forceinline__ uint2 downGroup(int q, const T* s0, const T* s1, int x, f32x2 nz) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(s0 + 2 * x));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(s1 + 2 * x));
    return downVec<T>(q, a, b, nz);
}
template <typename T>
__device__ __forceinline__ void storeGroup(uint2 o, T* d, int x, Iv ownX) {   // d = destination row; only owned texels are written
    constexpr int G = sizeof(T) == 4 ? 2 : 4;
    if (x >= ownX.a && x + G <= ownX.b) *reinterpret_cast<uint2*>(d + x) = o;
    else {
#pragma unroll
        for (int t = 0; t < G; ++t) {
            const uint32_t v = sizeof(T) == 4 ? (t ? o.y : o.x) : (((t < 2 ? o.x : o.y) >> (16 * (t & 1))) & 0xffffu);
            if (inIv(ownX, x + t)) d[x + t] = (T)v;
        }
    }
}

template <typename T>
__device__ __forceinline__ void stepFromGlobal(int q, const DevLevel& src, const DevLevel& dst, const Box& upd, Iv ownX, Iv ownY,
                                               Iv compX, Iv compY, bool keep, const Region<T>& out, f32x2 nz) {
    constexpr int G = sizeof(T) == 4 ? 2 : 4;   // destination texels per 128-bit source vector pair
    const int gx0 = out.x0;                     // == compX.a rounded down to a multiple of G
    const int ngx = (compX.b - gx0 + G - 1) / G;
    const int rows = compY.b - compY.a;
    if (q != Q_CUBIC && gx0 >= upd.minx && gx0 + ngx * G <= upd.maxx && compY.a >= upd.miny && compY.b <= upd.maxy) {
        // every group is regenerated (always the case on a full-frame update, except on the ragged right edge).
        // Main block: the 32 groups starting at the CTA's own first column, one warp per row, fully coalesced;
        // the few halo groups left and right of it are swept afterwards, a thread per (row, group).
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int p0 = (ownX.a - gx0) / G;
        const int nMain = min(32, ngx - p0);
        // RB rows per warp and pass, all loads issued before the first use: the source level streams from HBM and a
        // CTA has only ~9 rows per warp, so the loads in flight per SM decide the bandwidth
        constexpr int RB = 3, WARPS = FUSED_THREADS / 32;
        const int xm = gx0 + (p0 + lane) * G;
        for (int r0 = warp; r0 < rows; r0 += WARPS * RB) {
            uint4 va[RB], vb[RB];
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                const int y = compY.a + r0 + k * WARPS;
                if (r0 + k * WARPS < rows && lane < nMain) {
                    va[k] = __ldg(reinterpret_cast<const uint4*>(rowPtr<T>(src, 2 * y) + 2 * xm));
                    vb[k] = __ldg(reinterpret_cast<const uint4*>(rowPtr<T>(src, 2 * y + 1) + 2 * xm));
                }
            }
#pragma unroll
            for (int k = 0; k < RB; ++k) {
                const int y = compY.a + r0 + k * WARPS;
                if (r0 + k * WARPS < rows && lane < nMain) {
                    const uint2 o = downVec<T>(q, va[k], vb[k], nz);
                    if (keep) *reinterpret_cast<uint2*>(&out.at(xm, y)) = o;
                    if (inIv(ownY, y)) storeGroup<T>(o, rowPtrW<T>(dst, y), xm, ownX);
                }
            }
        }
        const int nSide = ngx - nMain;
        if (nSide > 0) {
            const uint32_t magic = 0xffffffffu / (uint32_t)nSide + 1u;
            for (uint32_t idx = threadIdx.x; idx < (uint32_t)(nSide * rows); idx += FUSED_THREADS) {
                const uint32_t r = fastDiv(idx, (uint32_t)nSide, magic);
                const int sIdx = (int)(idx - r * (uint32_t)nSide);
                const int y = compY.a + (int)r, x = gx0 + (sIdx < p0 ? sIdx : sIdx + nMain) * G;
                const uint2 o = downGroup<T>(q, rowPtr<T>(src, 2 * y), rowPtr<T>(src, 2 * y + 1), x, nz);
                if (keep) *reinterpret_cast<uint2*>(&out.at(x, y)) = o;
                if (inIv(ownY, y)) storeGroup<T>(o, rowPtrW<T>(dst, y), x, ownX);
            }
        }
        return;
    }
    // general case: dirty rectangles, ragged edges, cubic from global memory
    const uint32_t magic = 0xffffffffu / (uint32_t)ngx + 1u;
    const GlobalSrc<T> fetch{src};
    for (uint32_t idx = threadIdx.x; idx < (uint32_t)(ngx * rows); idx += FUSED_THREADS) {
        const uint32_t r = fastDiv(idx, (uint32_t)ngx, magic);
        const int y = compY.a + (int)r, x = gx0 + (int)(idx - r * (uint32_t)ngx) * G;
        const bool rowUpd = y >= upd.miny && y < upd.maxy;
        if (q != Q_CUBIC && rowUpd && x >= upd.minx && x + G <= upd.maxx) {
            // the whole group is recomputed (2x + 2G - 1 <= src.w - 1 because x + G <= dst.w)
            const uint2 o = downGroup<T>(q, rowPtr<T>(src, 2 * y), rowPtr<T>(src, 2 * y + 1), x, nz);
            if (keep) *reinterpret_cast<uint2*>(&out.at(x, y)) = o;
            if (inIv(ownY, y)) storeGroup<T>(o, rowPtrW<T>(dst, y), x, ownX);
        } else {
            for (int t = 0; t < G; ++t) {
                const int xx = x + t;
                if (xx >= dst.w) break;
                const bool isUpd = rowUpd && xx >= upd.minx && xx < upd.maxx;
                uint32_t v;
                if (isUpd) v = downTexel<T>(q, xx, y, src.w, src.h, fetch);
                else v = rowPtr<T>(dst, y)[xx];                      // not regenerated: the stored texel is what the chain reads
                if (keep) out.at(xx, y) = (T)v;
                if (isUpd && inIv(ownX, xx) && inIv(ownY, y)) rowPtrW<T>(dst, y)[xx] = (T)v;
            }
        }
    }
}

// copies the rectangle [rx) x [ry) of a level into a shared region (the source of a cubic first step: the 4x4
// footprints of neighbouring texels overlap, so they are fetched once, coalesced, instead of 16 loads per texel)
template <typename T>
__device__ __forceinline__ void stageRegion(const DevLevel& src, Iv rx, Iv ry, const Region<T>& out) {
    // out.x0 is rx.a rounded down to an even texel: rows are copied as aligned 64-bit words (the word straddling the
    // right edge of the level is completed texel by texel), one warp per row
    constexpr int E = 8 / sizeof(T);   // texels per 64-bit word
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = out.x0 & ~(E - 1);  // (for L16, x0 is only even: fall back to 32-bit words there)
    if (sizeof(T) == 4 || x0 == out.x0) {
        const int nw = (rx.b - x0 + E - 1) / E;
        // UR rows per warp and pass, loads first: a small-level launch is a chain of short phases, and a row-by-row
        // copy would expose one L2 round trip per row
        constexpr int UR = 4, WARPS = FUSED_THREADS / 32;
        for (int yb = ry.a + warp; yb < ry.b; yb += WARPS * UR) {
            for (int wI = lane; wI < nw; wI += 32) {
                const int x = x0 + wI * E;
                const bool whole = x + E <= src.w;
                uint2 v[UR];
#pragma unroll
                for (int k = 0; k < UR; ++k) {
                    const int y = yb + k * WARPS;
                    if (y < ry.b && whole) v[k] = __ldg(reinterpret_cast<const uint2*>(rowPtr<T>(src, y) + x));
                }
#pragma unroll
                for (int k = 0; k < UR; ++k) {
                    const int y = yb + k * WARPS;
                    if (y >= ry.b) continue;
                    T* drow = &out.at(x0, y);
                    if (whole) *reinterpret_cast<uint2*>(drow + wI * E) = v[k];
                    else for (int t = 0; x + t < src.w && t < E; ++t) drow[wI * E + t] = __ldg(rowPtr<T>(src, y) + x + t);
                }
            }
        }
    } else {
        const int nw = (rx.b - out.x0 + 1) / 2;
        for (int y = ry.a + warp; y < ry.b; y += FUSED_THREADS / 32) {
            const T* srow = rowPtr<T>(src, y);
            T* drow = &out.at(out.x0, y);
            for (int wI = lane; wI < nw; wI += 32) {
                const int x = out.x0 + wI * 2;
                if (x + 2 <= src.w) *reinterpret_cast<uint32_t*>(drow + wI * 2) = __ldg(reinterpret_cast<const uint32_t*>(srow + x));
                else drow[wI * 2] = __ldg(srow + x);
            }
        }
    }
}

// ---- steps from a shared region of level k-1 -> level k
template <typename T>
__device__ __forceinline__ void stepFromShared(int q, const Region<T>& in, int sw, int sh, const DevLevel& dst, const Box& upd, Iv ownX, Iv ownY,
                                               Iv compX, Iv compY, bool keep, const Region<T>& out) {
    const uint32_t ngx = (uint32_t)(compX.b - compX.a + 1) / 2;   // pairs of adjacent texels
    const uint32_t rows = (uint32_t)(compY.b - compY.a);
    const uint32_t magic = 0xffffffffu / ngx + 1u;
    const SharedSrc<T> fetch{in};
    // pairs that can take the two-texel cubic path: both texels regenerated and no horizontal clamp inside 2x-1 .. 2x+4
    const int fastLo = max(upd.minx, 1), fastHi = min(min(upd.maxx, compX.b) - 2, (sw - 5) >> 1);
    for (uint32_t idx = threadIdx.x; idx < ngx * rows; idx += FUSED_THREADS) {
        const uint32_t r = fastDiv(idx, ngx, magic);
        const int y = compY.a + (int)r, x = compX.a + (int)(idx - r * ngx) * 2;
        const bool rowUpd = y >= upd.miny && y < upd.maxy;
        if (sizeof(T) == 4 && q == Q_CUBIC && rowUpd && x >= fastLo && x <= fastHi) {
            // two adjacent texels share four of their six source columns 2x-1 .. 2x+4
            const T* col = in.p + (2 * x - in.x0);                   // (2x - in.x0) is even: 64-bit aligned rows
#ifdef B2_MIP_DEBUG
            (void)in.at(2 * x - 1, max(2 * y - 1, 0)); (void)in.at(2 * x + 4, min(2 * y + 2, sh - 1));
#endif
            const int ys[4] = {max(2 * y - 1, 0), 2 * y, 2 * y + 1, min(2 * y + 2, sh - 1)};
            uint32_t c[6][4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const T* row = col + (ys[k] - in.y0) * in.stride;
                const uint2 m = *reinterpret_cast<const uint2*>(row), p = *reinterpret_cast<const uint2*>(row + 2);
                c[0][k] = row[-1]; c[1][k] = m.x; c[2][k] = m.y; c[3][k] = p.x; c[4][k] = p.y; c[5][k] = row[4];
            }
            uint32_t lo[6], hi[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) vsumRGBA(c[k][0], c[k][1], c[k][2], c[k][3], lo[k], hi[k]);
            uint32_t o[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const uint32_t sl = (lo[2 * t] + lo[2 * t + 3]) + 3u * (lo[2 * t + 1] + lo[2 * t + 2]) + 0x00200020u;
                const uint32_t sh2 = (hi[2 * t] + hi[2 * t + 3]) + 3u * (hi[2 * t + 1] + hi[2 * t + 2]) + 0x00200020u;
                o[t] = ((sl >> 6) & 0x00ff00ffu) | (((sh2 >> 6) & 0x00ff00ffu) << 8);
            }
            if (keep) { T* w = &out.at(x, y); w[0] = (T)o[0]; w[1] = (T)o[1]; }
            if (inIv(ownY, y)) {
                T* d = rowPtrW<T>(dst, y);
                if (inIv(ownX, x)) d[x] = (T)o[0];
                if (inIv(ownX, x + 1)) d[x + 1] = (T)o[1];
            }
        } else {
            for (int t = 0; t < 2; ++t) {
                const int xx = x + t;
                if (xx >= compX.b) break;
                const bool isUpd = rowUpd && xx >= upd.minx && xx < upd.maxx;
                uint32_t v;
                if (isUpd) v = downTexel<T>(q, xx, y, sw, sh, fetch);
                else v = rowPtr<T>(dst, y)[xx];
                if (keep) out.at(xx, y) = (T)v;
                if (isUpd && inIv(ownX, xx) && inIv(ownY, y)) rowPtrW<T>(dst, y)[xx] = (T)v;
            }
        }
    }
}

// TW x TW = texels of level s+1 a CTA owns (before the odd-edge extension); the tile of the last level of an N-step
// launch is (TW >> (N-1))^2.  STAGE0: the source region of the first step goes through shared memory (cubic first
// step; small tiles only, the region of a 64 x 64 tile would not fit).
template <typename T, int N, int TW, bool STAGE0>
__global__ void __launch_bounds__(FUSED_THREADS, 4) k_mip_fused(const __grid_constant__ FusedArgs A) {
    constexpr int B1W = TW + 8, B1H = TW + 8, B2W = TW / 2 + 4, B2H = TW / 2 + 4, B0W = 2 * B1W + 4, B0H = 2 * B1H + 2;
    __shared__ __align__(16) T buf0[STAGE0 ? B0H * B0W : 8];
    __shared__ __align__(16) T buf1[N > 1 ? B1H * B1W : 8];
    __shared__ __align__(16) T buf2[N > 2 ? B2H * B2W : 8];
    const int tx = A.tile0x + (int)(blockIdx.x % (unsigned)A.tilesX), ty = A.tile0y + (int)(blockIdx.x / (unsigned)A.tilesX);
    constexpr int tw = TW >> (N - 1);

    // regions, from the last level down: own = texels this CTA writes, comp = texels it has to hold
    Iv ownX[4], ownY[4], compX[4], compY[4];
    ownX[N] = Iv{tx * tw, min((tx + 1) * tw, A.lv[N].w)};
    ownY[N] = Iv{ty * tw, min((ty + 1) * tw, A.lv[N].h)};
    compX[N] = ownX[N]; compY[N] = ownY[N];
#pragma unroll
    for (int k = N; k >= 2; --k) {
        ownX[k - 1] = ivChildren(ownX[k], A.lv[k].w, A.lv[k - 1].w);
        ownY[k - 1] = ivChildren(ownY[k], A.lv[k].h, A.lv[k - 1].h);
        compX[k - 1] = ivUnion(ownX[k - 1], ivNeed(A.q[k - 1], compX[k], A.lv[k - 1].w));
        compY[k - 1] = ivUnion(ownY[k - 1], ivNeed(A.q[k - 1], compY[k], A.lv[k - 1].h));
    }
    // Row bands (multi-GPU) store only a window of rows of every level: nothing outside it is read
    // or held.  Every texel a band regenerates has its sources inside the band's windows (b2_band_need_rows), and the
    // texels of a tile outside the update rectangles are only ever copied from the stored level, so clipping the
    // regions to the stored rows changes nothing on an unbanded pyramid and keeps a band inside its allocation.
#pragma unroll
    for (int k = 1; k <= N; ++k) {
        compY[k] = ivClip(compY[k], A.lv[k].y0, A.lv[k].y0 + A.lv[k].rows);
        ownY[k] = ivClip(ownY[k], A.lv[k].y0, A.lv[k].y0 + A.lv[k].rows);
    }
    constexpr int G = sizeof(T) == 4 ? 2 : 4;
    const Region<T> r1 = B2_REGION(T, buf1, compX[1].a & ~(G - 1), compY[1].a, B1W, B1H);
    if (STAGE0) {
        const Iv sx = ivNeed(A.q[0], compX[1], A.lv[0].w);
        const Iv sy = ivClip(ivNeed(A.q[0], compY[1], A.lv[0].h), A.lv[0].y0, A.lv[0].y0 + A.lv[0].rows);
        const Region<T> r0 = B2_REGION(T, buf0, sx.a & ~1, sy.a, B0W, B0H);
        stageRegion<T>(A.lv[0], sx, sy, r0);
        __syncthreads();
        stepFromShared<T>(A.q[0], r0, A.lv[0].w, A.lv[0].h, A.lv[1], A.upd[0], ownX[1], ownY[1], compX[1], compY[1], N > 1, r1);
    } else {
        stepFromGlobal<T>(A.q[0], A.lv[0], A.lv[1], A.upd[0], ownX[1], ownY[1], compX[1], compY[1], N > 1, r1, A.negZero2);
    }
    if (N >= 2) {
        const Region<T> r2 = B2_REGION(T, buf2, compX[2].a & ~1, compY[2].a, B2W, B2H);
        __syncthreads();
        stepFromShared<T>(A.q[1], r1, A.lv[1].w, A.lv[1].h, A.lv[2], A.upd[1], ownX[2], ownY[2], compX[2], compY[2], N > 2, r2);
        if (N >= 3) {
            __syncthreads();
            stepFromShared<T>(A.q[2], r2, A.lv[2].w, A.lv[2].h, A.lv[3], A.upd[2], ownX[3], ownY[3], compX[3], compY[3], false, r2);
        }
    }
}

// Tile variants (texels of level s+1 per CTA): 64 x 64 streaming from global memory for large levels whose first step
// is a box; 32 x 32 or 16 x 16 with the source region staged in shared memory when the first step is cubic; 16 x 16
// for small levels, where the number of CTAs (parallelism, latency) matters and the halo overhead does not.
enum { FUSED_V64 = 0, FUSED_V32S = 1, FUSED_V16 = 2, FUSED_V16S = 3 };
int mip_fused_tile(int n, int variant) { return (variant == FUSED_V64 ? 64 : variant == FUSED_V32S ? 32 : 16) >> (n - 1); }
// candidate variants for a chain starting with quality q0, largest tile first
void mip_fused_variants(int q0, int out[2]) {
    if (q0 == Q_CUBIC) { out[0] = FUSED_V32S; out[1] = FUSED_V16S; }
    else { out[0] = FUSED_V64; out[1] = FUSED_V16; }
}

template <typename T, int TW, bool STAGE0>
static void launchFusedT(const FusedArgs& a, int grid, cudaStream_t s) {
    if (a.n == 1) k_mip_fused<T, 1, TW, STAGE0><<<grid, FUSED_THREADS, 0, s>>>(a);
    else if (a.n == 2) k_mip_fused<T, 2, TW, STAGE0><<<grid, FUSED_THREADS, 0, s>>>(a);
    else k_mip_fused<T, 3, TW, STAGE0><<<grid, FUSED_THREADS, 0, s>>>(a);
}
template <typename T>
static void launchFusedV(int variant, const FusedArgs& a, int grid, cudaStream_t s) {
    if (variant == FUSED_V64) launchFusedT<T, 64, false>(a, grid, s);
    else if (variant == FUSED_V32S) launchFusedT<T, 32, true>(a, grid, s);
    else if (variant == FUSED_V16) launchFusedT<T, 16, false>(a, grid, s);
    else launchFusedT<T, 16, true>(a, grid, s);
}
void launch_mip_fused(bool rgba, int variant, const FusedArgs& a0, int tilesY, cudaStream_t s) {
    FusedArgs a = a0;
    a.negZero2 = kNegZero2;
    const int grid = a.tilesX * tilesY;
    if (rgba) launchFusedV<uint32_t>(variant, a, grid, s);
    else launchFusedV<uint16_t>(variant, a, grid, s);
}

}  // namespace b2
