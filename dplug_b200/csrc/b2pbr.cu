// libb2pbr — host runtime and C ABI (include/b2pbr.h) of the B200 PBR compositor.
//
// Owns the device twins of the reference's OwnedImage / Mipmap buffers (graphics/dplug/graphics/image.d,
// mipmap.d), the dirty-rectangle tile scheduler (gui/dplug/gui/boxlist.d, graphics.d:1338-1423) and the
// kernel launches.  There is no CPU implementation of any pixel work in this library: if CUDA is not
// usable every entry point that would touch pixels fails with B2_ERR_NO_DEVICE / B2_ERR_CUDA.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/b2pbr.h"
#include "common.cuh"
#include "host_rects.hpp"
#include "mip_wave.hpp"

namespace b2 {
// kernels (mip_kernels.cu, mip_fused.cu, pbr_exact.cu, pbr_tex.cu)
void launch_box_rgba(const DevLevel& dst, const DevLevel& src, Box r, bool premul, cudaStream_t s);
void launch_box_l16(const DevLevel& dst, const DevLevel& src, Box r, cudaStream_t s);
void launch_cubic_rgba(const DevLevel& dst, const DevLevel& src, Box r, cudaStream_t s);
void launch_cubic_l16(const DevLevel& dst, const DevLevel& src, Box r, cudaStream_t s);
void launch_copy2d(void* dst, size_t dpitch, const void* src, size_t spitch, size_t rowBytes, int rows, cudaStream_t s);
void launch_mip_fused(bool rgba, int variant, const FusedArgs& a, int tilesY, cudaStream_t s);
int mip_fused_tile(int n, int variant);
void mip_fused_variants(int q0, int out[2]);
void launch_copy_diffuse(const DevLevel& dst, const DevLevel& src, const Box* boxes, int n, int grid, cudaStream_t s);
void launch_swizzle(const DevLevel& dst, const DevLevel& src, Box box, int pf, cudaStream_t s);
void launch_layer(bool blend, const DevLevel* dst3, const DevLevel* src3, const DevLevel* op3, Box box, int px, int py, cudaStream_t s);
void launch_present(const DevLevel& win, const DevLevel& src, Box box, int dx, int dy, int pf, const uint8_t* lut, cudaStream_t s);
int resize_axis_table(int kind, int inSize, int outSize, int* start, int* src, float* coef, int cap);
int resize_host_to_device(int kind, const void* in, int inW, int inH, size_t inPitch, uint8_t* dOut, size_t outPitch, int outW, int outH, cudaStream_t s);
void launch_screenshot_qb(uint32_t* vox, const DevLevel& src, const DevLevel& depth, int W, int H, int pf, const uint8_t* lut, cudaStream_t s);
void launch_fill32(const DevLevel& win, uint32_t value, cudaStream_t s);
void launch_pbr_exact(const PbrArgs& args, int gridBlocks, cudaStream_t stream);
constexpr int kQuadrantRows = 32;   // rows of a full-height warp quadrant of k_pbr_tex (CTA box = 64 x 64)
void launch_pbr_tex(const PbrArgs& args, bool skyGather, int kernel, cudaStream_t stream);
void sky_atlas_layout(const DevLevel* levels, int n, int* ox, int* oy, int* aw, int* ah);
void launch_sky_atlas(const DevLevel* levels, int n, const int* ox, const int* oy, int aw, int ah, uint32_t* out, cudaStream_t s);
void launch_sample(const DevLevel* levels, int nLevels, int color, int kind, int n, const float* lvl, const float* x,
                   const float* y, float* out4, cudaStream_t stream);
}  // namespace b2

using namespace b2;

static const uint32_t kRsqrtssTable[2048] = {
#include "rsqrtss_table.inc"
};

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(B2_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                \
    } while (0)

static inline Box toBox(const b2_box2i& b) { return Box{b.min_x, b.min_y, b.max_x, b.max_y}; }
static inline size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ================================================================================================ b2mip
struct b2mip {
    int device = 0, color = 0, elem = 4;
    int W = 0, H = 0;
    int bandY0 = 0, bandY1 = 0, haloUp = 0, haloDown = 0;  // level-0 rows owned / stored around them
    std::vector<DevLevel> levels;
    uint8_t* alloc = nullptr;
    size_t allocBytes = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    uint64_t launches = 0;
    uint8_t* staging = nullptr;     // linear device buffer for uploads whose host pitch differs from the level's
    size_t stagingBytes = 0;
    // every level as a pitch-2D texture object over the same memory (linear filter, clamp addressing, unnormalised
    // coordinates, normalised-float reads): the lighting kernel's bilinear samplers run on the texture units
    std::vector<cudaTextureObject_t> tex;
    // plans of the whole-pyramid wavefront kernel (mip_wave.cu): item list + counters on the device, re-used while the
    // caller regenerates the same chain (a full redraw never changes it)
    struct WavePlan {
        std::vector<int> key;
        WaveItem* dItems = nullptr; WaveItem* hItems = nullptr; int cap = 0, nItems = 0;
        int* dCtl = nullptr; int ctlCap = 0, nCtl = 0;
        cudaStream_t lastStream = nullptr; bool usedOnce = false, pinned = false;
        uint64_t used = 0;
    };
    std::vector<WavePlan> wavePlans;
    uint64_t waveClock = 0;
};

// A pitched host<->device copy is executed row by row by the copy engine (measured: 2 MB of 2480-byte rows take 160 us
// = 13 GB/s, against 37 us as one linear copy).  Whole-row rectangles of a gapless host image therefore cross PCIe as
// ONE linear copy into / out of a linear device buffer and are re-pitched on the device (a device-to-device 2-D copy
// is a kernel and runs at HBM speed).
static int stagingReserve(uint8_t** buf, size_t* have, size_t need, cudaStream_t s) {
    if (need <= *have) return B2_OK;
    if (*buf) { CU(cudaStreamSynchronize(s)); cudaFree(*buf); *buf = nullptr; *have = 0; }
    need = alignUp(need + need / 4, 1 << 16);
    CU(cudaMalloc((void**)buf, need));
    *have = need;
    return B2_OK;
}

// Device-visible address of a page-locked host buffer (cudaMallocHost / cudaHostRegister), or nullptr for pageable
// memory.  Small transfers then bypass the copy engine (k_copy2d).
constexpr size_t kSmCopyMaxBytes = 4u << 20;
static void* mappedHostPointer(const void* host) {
    void* dev = nullptr;
    if (cudaHostGetDevicePointer(&dev, const_cast<void*>(host), 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return dev;
}

struct b2mip;
static int mipUploadRows(b2mip* m, int level, int y0, int y1, const void* hostPixels, size_t hostPitch, cudaStream_t st, bool allowSmCopy = true);

static int countLevels(int maxLevel, int w, int h) {  // mipmap.d:83-94
    int needed = 0;
    for (; needed <= maxLevel; ++needed) {
        if (w == 0 || h == 0) break;
        w >>= 1; h >>= 1;
    }
    return needed;
}

static int mipAllocate(b2mip* m, int maxLevel) {
    const int n = countLevels(maxLevel, m->W, m->H);
    m->levels.resize(n);
    const int sy0 = m->bandY0 - m->haloUp < 0 ? 0 : m->bandY0 - m->haloUp;
    const int sy1 = m->bandY1 + m->haloDown > m->H ? m->H : m->bandY1 + m->haloDown;
    const bool whole = (sy0 == 0 && sy1 == m->H);
    size_t total = 0;
    std::vector<size_t> offs(n);
    int w = m->W, h = m->H;
    for (int l = 0; l < n; ++l) {
        DevLevel& L = m->levels[l];
        L.w = w; L.h = h;
        L.pitch = (int)alignUp((size_t)w * m->elem, 128);
        if (whole) { L.y0 = 0; L.rows = h; }
        else {
            int lo = (sy0 >> l) - 2, hi = ((sy1 + (1 << l) - 1) >> l) + 2;
            if (lo < 0) lo = 0;
            if (hi > h) hi = h;
            if (hi < lo) hi = lo;
            L.y0 = lo; L.rows = hi - lo;
        }
        offs[l] = total;
        total += alignUp((size_t)L.pitch * (size_t)(L.rows > 0 ? L.rows : 1), 512);   // 512 B: texture base alignment
        w >>= 1; h >>= 1;
    }
    if (total == 0) total = 256;
    CU(cudaMalloc((void**)&m->alloc, total));
    m->allocBytes = total;
    CU(cudaMemsetAsync(m->alloc, 0, total, m->stream));
    for (int l = 0; l < n; ++l) m->levels[l].base = m->alloc + offs[l];
    m->tex.assign(n, 0);
    for (int l = 1; l < n; ++l) {      // level 0 is never sampled through a texture (and may exceed the 65000-row limit of pitch-2D textures)
        const DevLevel& L = m->levels[l];
        if (L.w <= 0 || L.rows <= 0) continue;
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypePitch2D;
        rd.res.pitch2D.devPtr = L.base;
        rd.res.pitch2D.desc = m->color == B2_COLOR_RGBA8 ? cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned)
                                                         : cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
        rd.res.pitch2D.width = (size_t)L.w;
        rd.res.pitch2D.height = (size_t)L.rows;
        rd.res.pitch2D.pitchInBytes = (size_t)L.pitch;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 0;
        CU(cudaCreateTextureObject(&m->tex[l], &rd, &td, nullptr));
    }
    return B2_OK;
}

static int mipCreate(b2mip** out, int device, int color, int maxLevel, int w, int h, int y0, int y1, int up, int down) {
    if (!out || w < 0 || h < 0 || maxLevel < 0 || (color != B2_COLOR_RGBA8 && color != B2_COLOR_L16))
        return fail(B2_ERR_INVALID, "b2mip_create: bad arguments");
    if (y0 < 0 || y1 > h || y0 > y1) return fail(B2_ERR_INVALID, "b2mip_create: bad band");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(B2_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)");
    if (device < 0 || device >= ndev) return fail(B2_ERR_INVALID, "bad device index");
    CU(cudaSetDevice(device));
    b2mip* m = new b2mip();
    m->device = device; m->color = color; m->elem = (color == B2_COLOR_RGBA8) ? 4 : 2;
    m->W = w; m->H = h; m->bandY0 = y0; m->bandY1 = y1; m->haloUp = up; m->haloDown = down;
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) { delete m; return fail(B2_ERR_CUDA, "cudaStreamCreate failed"); }
    m->ownStream = true;
    int rc = mipAllocate(m, maxLevel);
    if (rc != B2_OK) {
        for (cudaTextureObject_t t : m->tex) if (t) cudaDestroyTextureObject(t);
        if (m->alloc) cudaFree(m->alloc);
        cudaStreamDestroy(m->stream); delete m; return rc;
    }
    *out = m;
    return B2_OK;
}

// whole rows [y0, y1) of a level from a host image (pixel (0,0) at hostPixels), on stream `st`
static int mipUploadRows(b2mip* m, int level, int y0, int y1, const void* hostPixels, size_t hostPitch, cudaStream_t st, bool allowSmCopy) {
    if (y1 <= y0) return B2_OK;
    const DevLevel& L = m->levels[level];
    uint8_t* dst = L.base + (size_t)(y0 - L.y0) * L.pitch;
    const uint8_t* src = (const uint8_t*)hostPixels + (size_t)y0 * hostPitch;
    const size_t rowBytes = (size_t)L.w * m->elem, rows = (size_t)(y1 - y0);
    if (allowSmCopy && rowBytes * rows <= kSmCopyMaxBytes)
        if (void* mapped = mappedHostPointer(src)) {                          // small + page-locked: the SMs read the host image in place
            launch_copy2d(dst, (size_t)L.pitch, mapped, hostPitch, rowBytes, (int)rows, st);
            m->launches++;
            CU(cudaGetLastError());
            return B2_OK;
        }
    if (rowBytes == hostPitch && hostPitch == (size_t)L.pitch)               // gapless on both sides: one linear copy
        CU(cudaMemcpyAsync(dst, src, rowBytes * rows, cudaMemcpyHostToDevice, st));
    else if (rowBytes == hostPitch && rows > 1) {                             // gapless host rows: linear copy, re-pitch on the device
        int rc = stagingReserve(&m->staging, &m->stagingBytes, rowBytes * rows, st);
        if (rc) return rc;
        CU(cudaMemcpyAsync(m->staging, src, rowBytes * rows, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpy2DAsync(dst, (size_t)L.pitch, m->staging, rowBytes, rowBytes, rows, cudaMemcpyDeviceToDevice, st));
    } else
        CU(cudaMemcpy2DAsync(dst, (size_t)L.pitch, src, hostPitch, rowBytes, rows, cudaMemcpyHostToDevice, st));
    return B2_OK;
}

static int mipCheckRect(const b2mip* m, int level, const b2_box2i& r, const char* who) {
    if (level < 0 || level >= (int)m->levels.size()) return fail(B2_ERR_INVALID, std::string(who) + ": level out of range");
    const DevLevel& L = m->levels[level];
    if (!boxSorted(r) || r.min_x < 0 || r.min_y < 0 || r.max_x > L.w || r.max_y > L.h)
        return fail(B2_ERR_INVALID, std::string(who) + ": rectangle outside the level");
    if (!boxEmpty(r) && (r.min_y < L.y0 || r.max_y > L.y0 + L.rows))
        return fail(B2_ERR_INVALID, std::string(who) + ": rectangle outside the rows stored by this band");
    return B2_OK;
}

// the source rows a destination rectangle needs must be stored too (only relevant for row bands)
static int mipCheckSourceRows(const b2mip* m, int level, int quality, const b2_box2i& r) {
    const DevLevel& src = m->levels[level - 1];
    int lo = quality == B2_QUALITY_CUBIC ? 2 * r.min_y - 1 : 2 * r.min_y;
    int hi = quality == B2_QUALITY_CUBIC ? 2 * (r.max_y - 1) + 2 : 2 * (r.max_y - 1) + 1;
    if (lo < 0) lo = 0;
    if (hi > src.h - 1) hi = src.h - 1;
    if (lo < src.y0 || hi >= src.y0 + src.rows) return fail(B2_ERR_INVALID, "generateLevel: source rows outside the band's halo");
    return B2_OK;
}

static int mipGenerateLevel(b2mip* m, int level, int quality, b2_box2i r) {
    if (level < 1 || level >= (int)m->levels.size()) return fail(B2_ERR_INVALID, "generateLevel: level out of range");
    // a degenerate rectangle makes every reference loop a no-op (mipmap.d:683,771,830,906,1046); Box.intersection
    // hands such boxes back un-clipped (box.d:303-308), so they may even lie outside the level
    if (boxW(r) <= 0 || boxH(r) <= 0) return B2_OK;
    int rc = mipCheckRect(m, level, r, "generateLevel");
    if (rc) return rc;
    const DevLevel& dst = m->levels[level];
    const DevLevel& src = m->levels[level - 1];
    if ((rc = mipCheckSourceRows(m, level, quality, r))) return rc;
    CU(cudaSetDevice(m->device));
    if (m->color == B2_COLOR_RGBA8) {
        if (quality == B2_QUALITY_BOX) launch_box_rgba(dst, src, toBox(r), false, m->stream);
        else if (quality == B2_QUALITY_BOX_ALPHACOV_PREMUL) launch_box_rgba(dst, src, toBox(r), true, m->stream);
        else if (quality == B2_QUALITY_CUBIC) launch_cubic_rgba(dst, src, toBox(r), m->stream);
        else return fail(B2_ERR_INVALID, "generateLevel: unknown quality");
    } else {
        if (quality == B2_QUALITY_BOX) launch_box_l16(dst, src, toBox(r), m->stream);
        else if (quality == B2_QUALITY_CUBIC) launch_cubic_l16(dst, src, toBox(r), m->stream);
        else return fail(B2_ERR_INVALID, "generateLevel: boxAlphaCovIntoPremul is RGBA-only (mipmap.d:594-595)");
    }
    m->launches++;
    CU(cudaGetLastError());
    return B2_OK;
}

// ---- fused chain (mip_fused.cu).  `rects[k]` is the update rectangle of level first + k, `q[k]` the quality of the
// step producing it; exactly the result of n successive mipGenerateLevel calls.
static bool g_mipFusion = true;
static long long g_mipFuseBelow = getenv("B2_MIP_FUSE_BELOW") ? atoll(getenv("B2_MIP_FUSE_BELOW")) : (1ll << 20);   // source levels up to this many texels start a fused group (env: tuning aid)

static int mipGenerateGroup(b2mip* m, int first, int n, const int* q, const b2_box2i* rects) {
    FusedArgs a{};
    a.n = n;
    a.lv[0] = m->levels[first - 1];
    const int last = first + n - 1;
    const DevLevel& LL = m->levels[last];
    for (int k = 0; k < n; ++k) {
        a.lv[k + 1] = m->levels[first + k];
        a.q[k] = q[k];
        if (m->color != B2_COLOR_RGBA8 && q[k] == B2_QUALITY_BOX_ALPHACOV_PREMUL)
            return fail(B2_ERR_INVALID, "generateLevel: boxAlphaCovIntoPremul is RGBA-only (mipmap.d:594-595)");
        if (q[k] < 0 || q[k] > 2) return fail(B2_ERR_INVALID, "generateLevel: unknown quality");
        const b2_box2i& r = rects[k];
        if (boxW(r) <= 0 || boxH(r) <= 0) { a.upd[k] = Box{0, 0, 0, 0}; continue; }   // a no-op in the reference
        if (int rc = mipCheckRect(m, first + k, r, "generateLevel")) return rc;
        if (int rc = mipCheckSourceRows(m, first + k, q[k], r)) return rc;
        a.upd[k] = toBox(r);
    }
    // tiles of the last level owning a texel of some rectangle: the rectangles' ancestors (the odd last column / row
    // of a level belongs to the tile on the edge).  Few large tiles -> small tiles (more CTAs).
    int tile = 0, tx0 = 0, ty0 = 0, tx1 = -1, ty1 = -1, variant = 0, variants[2];
    mip_fused_variants(q[0], variants);
    for (int pass = 0; pass < 2; ++pass) {
        variant = variants[pass];
        tile = mip_fused_tile(n, variant);
        tx0 = ty0 = 1 << 30; tx1 = ty1 = -1;
        for (int k = 0; k < n; ++k) {
            const b2_box2i& r = rects[k];
            if (boxW(r) <= 0 || boxH(r) <= 0) continue;
            const int sh = n - 1 - k;
            auto anc = [&](int v, int dim) { v >>= sh; return v < dim ? v : dim - 1; };
            const int ax0 = anc(r.min_x, LL.w) / tile, ax1 = anc(r.max_x - 1, LL.w) / tile;
            const int ay0 = anc(r.min_y, LL.h) / tile, ay1 = anc(r.max_y - 1, LL.h) / tile;
            if (ax0 < tx0) tx0 = ax0;
            if (ay0 < ty0) ty0 = ay0;
            if (ax1 > tx1) tx1 = ax1;
            if (ay1 > ty1) ty1 = ay1;
        }
        if (tx1 < 0 || (long long)(tx1 - tx0 + 1) * (ty1 - ty0 + 1) >= 2 * 148) break;
    }
    if (tx1 < 0) return B2_OK;   // nothing to regenerate
    a.tile0x = tx0; a.tile0y = ty0; a.tilesX = tx1 - tx0 + 1;
    CU(cudaSetDevice(m->device));
    launch_mip_fused(m->color == B2_COLOR_RGBA8, variant, a, ty1 - ty0 + 1, m->stream);
    m->launches++;
    CU(cudaGetLastError());
    return B2_OK;
}

// ---- whole chain in one persistent launch (mip_wave.cu), for chains that start on a large level (the regime where the
// levels otherwise stream one launch each).  OFF by default: measured on the B200 it moves the minimum of bytes (8192^2
// pyramid: 269 MB read + 71 MB written per chain, level l + 1 reads level l from L2) but is slower than the launches it
// replaces (131 - 152 us against 98 us; 3840x2160 diffuse chain 38 - 48 us against 25 us): every level adds a
// claim / wait / compute / release hop to the tail and the software queue costs more per item than the hardware block
// scheduler (profiles/r2_mip_wave.md).  b2_set_mip_wave(1) / B2_MIP_WAVE=1 switch it on.  `*launched` tells whether the
// chain was enqueued; when it was not (knob off, row bands, plan not yet on the device while a graph is being recorded)
// the caller falls back to the per-level / fused launches, which write the same bytes.
// (Also measured and removed: the per-level kernels cut into 2 - 8 row chunks on one stream per level with events, so
// that the ALU-bound cubic levels run beside the HBM-bound level-1 box: 106 - 130 us against 98 us.)
static int g_mipWave = getenv("B2_MIP_WAVE") ? atoi(getenv("B2_MIP_WAVE")) : 0;
static int g_mipWaveSteps = getenv("B2_MIP_WAVE_STEPS") ? atoi(getenv("B2_MIP_WAVE_STEPS")) : kWaveMaxSteps;                   // tuning aid: at most this many levels per wavefront launch
static long long g_mipWaveTail = getenv("B2_MIP_WAVE_TAIL") ? atoll(getenv("B2_MIP_WAVE_TAIL")) : 0;                           // levels whose source has at most this many texels are left to the fused launches
static long long g_mipWaveAbove = getenv("B2_MIP_WAVE_ABOVE") ? atoll(getenv("B2_MIP_WAVE_ABOVE")) : (1ll << 20);

static int mipGenerateWave(b2mip* m, int first, int n, const int* q, const b2_box2i* rects, bool* launched) {
    *launched = false;
    if (!g_mipWave || n < 2 || n > kWaveMaxSteps) return B2_OK;
    if (m->levels[0].rows != m->H) return B2_OK;                                   // row bands keep their own schedule
    // only where the levels would otherwise stream one launch each (not below the fused kernel's regime, b2_set_mip_fusion)
    const long long above = g_mipWaveAbove > g_mipFuseBelow ? g_mipWaveAbove : g_mipFuseBelow;
    if (boxW(rects[0]) <= 0 || boxH(rects[0]) <= 0 || 4ll * boxW(rects[0]) * boxH(rects[0]) <= above) return B2_OK;
    const bool rgba = m->color == B2_COLOR_RGBA8;
    Box upd[kWaveMaxSteps];
    for (int k = 0; k < n; ++k) {
        if (!rgba && q[k] == B2_QUALITY_BOX_ALPHACOV_PREMUL) return fail(B2_ERR_INVALID, "generateLevel: boxAlphaCovIntoPremul is RGBA-only (mipmap.d:594-595)");
        if (q[k] < 0 || q[k] > 2) return fail(B2_ERR_INVALID, "generateLevel: unknown quality");
        const b2_box2i& r = rects[k];
        if (boxW(r) <= 0 || boxH(r) <= 0) { upd[k] = Box{0, 0, 0, 0}; continue; }   // a no-op in the reference
        if (int rc = mipCheckRect(m, first + k, r, "generateLevel")) return rc;
        upd[k] = toBox(r);
    }
    CU(cudaSetDevice(m->device));
    std::vector<int> key;
    key.reserve(2 + 5 * n);
    key.push_back(first); key.push_back(n);
    for (int k = 0; k < n; ++k) { key.push_back(q[k]); key.push_back(upd[k].minx); key.push_back(upd[k].miny); key.push_back(upd[k].maxx); key.push_back(upd[k].maxy); }
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CU(cudaStreamIsCapturing(m->stream, &cap));
    const bool capturing = cap != cudaStreamCaptureStatusNone;
    b2mip::WavePlan* plan = nullptr;
    for (auto& wp : m->wavePlans)
        if (wp.key == key) { plan = &wp; break; }
    if (!plan) {
        if (capturing) return B2_OK;   // the plan has to be uploaded outside a capture: this recording takes the per-level launches
        std::vector<WaveItem> items;
        int nCtl = 0;
        DevLevel lv[kWaveMaxSteps + 1];
        for (int k = 0; k <= n; ++k) lv[k] = m->levels[first - 1 + k];
        if (wave_build_plan(lv, upd, q, n, rgba, items, &nCtl) != 0 || items.empty()) return B2_OK;
        // slot: a free one (up to 8), else the least recently used that no graph refers to
        if (m->wavePlans.size() < 8) { m->wavePlans.emplace_back(); plan = &m->wavePlans.back(); }
        else {
            for (auto& wp : m->wavePlans)
                if (!wp.pinned && (!plan || wp.used < plan->used)) plan = &wp;
            if (!plan) { m->wavePlans.emplace_back(); plan = &m->wavePlans.back(); }
        }
        if (plan->usedOnce) CU(cudaStreamSynchronize(plan->lastStream));            // its last launch and upload are over
        const int nItems = (int)items.size();
        if (nItems > plan->cap) {
            if (plan->dItems) cudaFree(plan->dItems);
            if (plan->hItems) cudaFreeHost(plan->hItems);
            plan->dItems = nullptr; plan->hItems = nullptr;
            plan->cap = nItems + nItems / 2 + 256;
            CU(cudaMalloc((void**)&plan->dItems, sizeof(WaveItem) * plan->cap));
            CU(cudaMallocHost((void**)&plan->hItems, sizeof(WaveItem) * plan->cap));
        }
        if (nCtl > plan->ctlCap) {
            if (plan->dCtl) cudaFree(plan->dCtl);
            plan->dCtl = nullptr;
            plan->ctlCap = nCtl + nCtl / 2 + 64;
            CU(cudaMalloc((void**)&plan->dCtl, sizeof(int) * plan->ctlCap));
            CU(cudaMemsetAsync(plan->dCtl, 0, sizeof(int) * plan->ctlCap, m->stream));   // the kernel leaves them zero
        }
        memcpy(plan->hItems, items.data(), sizeof(WaveItem) * nItems);
        CU(cudaMemcpyAsync(plan->dItems, plan->hItems, sizeof(WaveItem) * nItems, cudaMemcpyHostToDevice, m->stream));
        plan->key = key; plan->nItems = nItems; plan->nCtl = nCtl;
    }
    if (capturing) plan->pinned = true;                                             // a recorded graph replays this plan
    plan->used = ++m->waveClock; plan->usedOnce = true; plan->lastStream = m->stream;
    WaveArgs a{};
    for (int k = 0; k <= n; ++k) a.lv[k] = m->levels[first - 1 + k];
    for (int k = 0; k < n; ++k) a.upd[k] = upd[k];
    a.items = plan->dItems; a.nItems = plan->nItems; a.ctl = plan->dCtl; a.nCtl = plan->nCtl; a.nz = 0x8000000080000000ull;
    launch_mip_wave(a, m->stream);
    m->launches++;
    CU(cudaGetLastError());
    *launched = true;
    return B2_OK;
}

// levels first .. first + n - 1, in launches of up to three levels (or level by level with fusion off); row bands
// take the same launches (the fused kernel clips its regions to the rows a band stores)
static int mipGenerateLevels(b2mip* m, int first, int n, const int* q, const b2_box2i* rects, bool fuseAll = false) {
    if (first < 1 || n < 0 || first + n > (int)m->levels.size()) return fail(B2_ERR_INVALID, "generateLevel: level out of range");
    if (!g_mipFusion) {
        for (int k = 0; k < n; ++k) {
            int rc = mipGenerateLevel(m, first + k, q[k], rects[k]);
            if (rc) return rc;
        }
        return B2_OK;
    }
    int k0 = 0;
    if (!fuseAll) {
        // the wavefront launch takes the large levels (those that would stream one launch each); the small levels that
        // follow are latency: they stay with the fused launches
        int nw = 0;
        while (nw < n && nw < g_mipWaveSteps && boxW(rects[nw]) > 0 && boxH(rects[nw]) > 0 && 4ll * boxW(rects[nw]) * boxH(rects[nw]) > g_mipWaveTail) ++nw;
        bool launched = false;
        if (int rc = mipGenerateWave(m, first, nw, q, rects, &launched)) return rc;
        if (launched) k0 = nw;
    }
    for (int k = k0; k < n;) {
        // Large updates are bandwidth work and the lean one-launch-per-level kernels stream them best (measured: the
        // fused kernel's region bookkeeping costs more issue slots than the saved HBM round trip is worth once a step
        // reads ~1 M texels).  Small updates — small levels, or dirty rectangles of large ones — are launch latency:
        // they go through the fused kernel, up to three levels per launch.
        const long long srcTexels = (boxW(rects[k]) > 0 && boxH(rects[k]) > 0) ? 4ll * boxW(rects[k]) * boxH(rects[k]) : 0;
        if (!fuseAll && srcTexels > g_mipFuseBelow) {
            int rc = mipGenerateLevel(m, first + k, q[k], rects[k]);
            if (rc) return rc;
            k += 1;
            continue;
        }
        const int g = n - k < 3 ? n - k : 3;
        int rc = mipGenerateGroup(m, first + k, g, q + k, rects + k);
        if (rc) return rc;
        k += g;
    }
    return B2_OK;
}

extern "C" {

const char* b2_last_error(void) { return g_err.c_str(); }
int b2_abi_version(void) { return B2PBR_ABI_VERSION; }
int b2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int b2mip_create(b2mip** out, int device, int color, int max_level, int w, int h) {
    return mipCreate(out, device, color, max_level, w, h, 0, h, 0, 0);
}
int b2mip_create_band(b2mip** out, int device, int color, int max_level, int w, int h, int y0, int y1, int halo_up, int halo_down) {
    return mipCreate(out, device, color, max_level, w, h, y0, y1, halo_up, halo_down);
}
void b2mip_destroy(b2mip* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->stream) cudaStreamSynchronize(m->stream);
    for (cudaTextureObject_t t : m->tex) if (t) cudaDestroyTextureObject(t);
    if (m->alloc) cudaFree(m->alloc);
    if (m->staging) cudaFree(m->staging);
    for (auto& wp : m->wavePlans) {
        if (wp.usedOnce && wp.lastStream != m->stream) cudaStreamSynchronize(wp.lastStream);
        if (wp.dItems) cudaFree(wp.dItems);
        if (wp.hItems) cudaFreeHost(wp.hItems);
        if (wp.dCtl) cudaFree(wp.dCtl);
    }
    if (m->ownStream && m->stream) cudaStreamDestroy(m->stream);
    delete m;
}
int b2mip_set_stream(b2mip* m, void* s) {
    if (!m) return fail(B2_ERR_INVALID, "null mip");
    if (m->ownStream && m->stream) { cudaStreamSynchronize(m->stream); cudaStreamDestroy(m->stream); }
    m->stream = (cudaStream_t)s;
    m->ownStream = false;
    return B2_OK;
}
int b2mip_num_levels(const b2mip* m) { return m ? (int)m->levels.size() : 0; }
int b2mip_width(const b2mip* m) { return m ? m->W : 0; }
int b2mip_height(const b2mip* m) { return m ? m->H : 0; }
int b2mip_level(const b2mip* m, int level, b2_imageref* out, int* first_row) {
    if (!m || !out || level < 0 || level >= (int)m->levels.size()) return fail(B2_ERR_INVALID, "b2mip_level: bad arguments");
    const DevLevel& L = m->levels[level];
    out->w = L.w; out->h = L.h; out->pitch = (size_t)L.pitch; out->pixels = L.base;
    if (first_row) *first_row = L.y0;
    return B2_OK;
}

int b2mip_upload(b2mip* m, int level, const void* host, size_t host_pitch, const b2_box2i* rects, int n) {
    if (!m || !host || (n > 0 && !rects)) return fail(B2_ERR_INVALID, "b2mip_upload: bad arguments");
    CU(cudaSetDevice(m->device));
    for (int k = 0; k < n; ++k) {
        int rc = mipCheckRect(m, level, rects[k], "b2mip_upload");
        if (rc) return rc;
    }
    std::vector<b2_box2i> merged;
    mergeRects(rects, n, merged);
    for (const b2_box2i& r : merged) {
        const DevLevel& L = m->levels[level];
        if (r.min_x == 0 && r.max_x == L.w) {    // whole rows
            int rc = mipUploadRows(m, level, r.min_y, r.max_y, host, host_pitch, m->stream);
            if (rc) return rc;
            continue;
        }
        uint8_t* dst = L.base + (size_t)(r.min_y - L.y0) * L.pitch + (size_t)r.min_x * m->elem;
        const uint8_t* src = (const uint8_t*)host + (size_t)r.min_y * host_pitch + (size_t)r.min_x * m->elem;
        if ((size_t)boxW(r) * m->elem * (size_t)boxH(r) <= kSmCopyMaxBytes)
            if (void* mapped = mappedHostPointer(src)) {
                launch_copy2d(dst, (size_t)L.pitch, mapped, host_pitch, (size_t)boxW(r) * m->elem, boxH(r), m->stream);
                m->launches++;
                CU(cudaGetLastError());
                continue;
            }
        CU(cudaMemcpy2DAsync(dst, (size_t)L.pitch, src, host_pitch, (size_t)boxW(r) * m->elem, (size_t)boxH(r), cudaMemcpyHostToDevice, m->stream));
    }
    return B2_OK;
}
int b2mip_download(b2mip* m, int level, void* host, size_t host_pitch, const b2_box2i* rects, int n) {
    if (!m || !host || (n > 0 && !rects)) return fail(B2_ERR_INVALID, "b2mip_download: bad arguments");
    CU(cudaSetDevice(m->device));
    for (int k = 0; k < n; ++k) {
        int rc = mipCheckRect(m, level, rects[k], "b2mip_download");
        if (rc) return rc;
    }
    std::vector<b2_box2i> merged;
    mergeRects(rects, n, merged);
    for (const b2_box2i& r : merged) {
        const DevLevel& L = m->levels[level];
        CU(cudaMemcpy2DAsync((uint8_t*)host + (size_t)r.min_y * host_pitch + (size_t)r.min_x * m->elem, host_pitch,
                             L.base + (size_t)(r.min_y - L.y0) * L.pitch + (size_t)r.min_x * m->elem, (size_t)L.pitch,
                             (size_t)boxW(r) * m->elem, (size_t)boxH(r), cudaMemcpyDeviceToHost, m->stream));
    }
    CU(cudaStreamSynchronize(m->stream));
    return B2_OK;
}

int b2mip_generate_level(b2mip* m, int level, int quality, b2_box2i r) {
    if (!m) return fail(B2_ERR_INVALID, "null mip");
    return mipGenerateLevel(m, level, quality, r);
}
int b2mip_generate_next_level(b2mip* m, int quality, b2_box2i prev, int level, b2_box2i* out_rect) {
    if (!m || level < 1 || level >= (int)m->levels.size()) return fail(B2_ERR_INVALID, "generateNextLevel: level out of range");
    const DevLevel& P = m->levels[level - 1];
    b2_box2i r = impactOnNextLevel(quality, prev, P.w, P.h);
    if (out_rect) *out_rect = r;
    return mipGenerateLevel(m, level, quality, r);
}
int b2mip_generate_levels(b2mip* m, int first_level, int n, const int* qualities, b2_box2i prev, b2_box2i* out_rects) {
    if (!m || n < 0 || (n > 0 && !qualities) || n > 32) return fail(B2_ERR_INVALID, "generate_levels: bad arguments");
    if (first_level < 1 || first_level + n > (int)m->levels.size()) return fail(B2_ERR_INVALID, "generate_levels: level out of range");
    b2_box2i rects[32];
    for (int k = 0; k < n; ++k) {
        const DevLevel& P = m->levels[first_level + k - 1];
        prev = impactOnNextLevel(qualities[k], prev, P.w, P.h);
        rects[k] = prev;
        if (out_rects) out_rects[k] = prev;
    }
    return mipGenerateLevels(m, first_level, n, qualities, rects);
}
int b2mip_generate_mipmaps(b2mip* m, int quality) {
    if (!m) return fail(B2_ERR_INVALID, "null mip");
    int q[32];
    const int n = (int)m->levels.size() - 1;
    for (int level = 1; level <= n && level <= 32; ++level) {
        if (level >= 3 && quality == B2_QUALITY_BOX) quality = B2_QUALITY_CUBIC;  // mipmap.d:522-524
        q[level - 1] = quality;
    }
    return b2mip_generate_levels(m, 1, n, q, b2_box2i{0, 0, m->W, m->H}, nullptr);
}
int b2mip_generate_chain(b2mip* m, int first_quality, int cubic_from_level) {
    if (!m) return fail(B2_ERR_INVALID, "null mip");
    int q[32];
    const int n = (int)m->levels.size() - 1;
    for (int level = 1; level <= n && level <= 32; ++level)
        q[level - 1] = level >= cubic_from_level ? B2_QUALITY_CUBIC : (level == 1 ? first_quality : B2_QUALITY_BOX);
    return b2mip_generate_levels(m, 1, n, q, b2_box2i{0, 0, m->W, m->H}, nullptr);
}
int b2_set_mip_fusion(int enabled) {
    const int was = g_mipFusion ? (g_mipFuseBelow >= (1ll << 40) ? 2 : 1) : 0;
    g_mipFusion = enabled != 0;
    g_mipFuseBelow = enabled == 2 ? (1ll << 40) : (getenv("B2_MIP_FUSE_BELOW") ? atoll(getenv("B2_MIP_FUSE_BELOW")) : (1ll << 20));
    return was;
}
int b2_mip_wave_plan(int color, int w, int h, int first_level, int n, const int* qualities, b2_box2i prev, int* items8, int cap_items, int* n_ctl) {
    if (n < 1 || n > kWaveMaxSteps || !qualities || first_level < 1 || w < 1 || h < 1) return fail(B2_ERR_INVALID, "b2_mip_wave_plan: bad arguments");
    DevLevel lv[kWaveMaxSteps + 1];
    Box upd[kWaveMaxSteps];
    for (int k = 0; k <= n; ++k) {
        const int l = first_level - 1 + k;
        lv[k] = DevLevel{nullptr, 0, w >> l, h >> l, 0, h >> l};
        if (k > 0 && (lv[k].w < 1 || lv[k].h < 1)) return fail(B2_ERR_INVALID, "b2_mip_wave_plan: level out of range");
    }
    for (int k = 0; k < n; ++k) {
        prev = impactOnNextLevel(qualities[k], prev, lv[k].w, lv[k].h);
        upd[k] = (boxW(prev) > 0 && boxH(prev) > 0) ? toBox(prev) : Box{0, 0, 0, 0};
    }
    std::vector<WaveItem> items;
    int nCtl = 0;
    if (wave_build_plan(lv, upd, qualities, n, color == B2_COLOR_RGBA8, items, &nCtl) != 0) return fail(B2_ERR_STATE, "b2_mip_wave_plan: chain cannot be planned");
    if (n_ctl) *n_ctl = nCtl;
    if (items8)
        for (size_t i = 0; i < items.size() && (int)i < cap_items; ++i) memcpy(items8 + 8 * i, &items[i], sizeof(WaveItem));
    return (int)items.size();
}
int b2_set_mip_wave(int enabled) {
    const int was = g_mipWave;
    g_mipWave = enabled != 0;
    return was;
}
int b2mip_sample(b2mip* m, int kind, int n, const float* level, const float* x, const float* y, float* out4) {
    if (!m || n < 0 || kind < 0 || kind > 2) return fail(B2_ERR_INVALID, "b2mip_sample: bad arguments");
    if (n == 0) return B2_OK;
    CU(cudaSetDevice(m->device));
    float *dl = nullptr, *dx = nullptr, *dy = nullptr, *dout = nullptr;
    DevLevel* dlev = nullptr;
    CU(cudaMalloc((void**)&dl, sizeof(float) * n));
    CU(cudaMalloc((void**)&dx, sizeof(float) * n));
    CU(cudaMalloc((void**)&dy, sizeof(float) * n));
    CU(cudaMalloc((void**)&dout, sizeof(float) * 4 * n));
    CU(cudaMalloc((void**)&dlev, sizeof(DevLevel) * m->levels.size()));
    CU(cudaMemcpyAsync(dl, level, sizeof(float) * n, cudaMemcpyHostToDevice, m->stream));
    CU(cudaMemcpyAsync(dx, x, sizeof(float) * n, cudaMemcpyHostToDevice, m->stream));
    CU(cudaMemcpyAsync(dy, y, sizeof(float) * n, cudaMemcpyHostToDevice, m->stream));
    CU(cudaMemcpyAsync(dlev, m->levels.data(), sizeof(DevLevel) * m->levels.size(), cudaMemcpyHostToDevice, m->stream));
    launch_sample(dlev, (int)m->levels.size(), m->color, kind, n, dl, dx, dy, dout, m->stream);
    m->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out4, dout, sizeof(float) * 4 * n, cudaMemcpyDeviceToHost, m->stream));
    CU(cudaStreamSynchronize(m->stream));
    cudaFree(dl); cudaFree(dx); cudaFree(dy); cudaFree(dout); cudaFree(dlev);
    return B2_OK;
}
int b2mip_sync(b2mip* m) {
    if (!m) return fail(B2_ERR_INVALID, "null mip");
    CU(cudaStreamSynchronize(m->stream));
    return B2_OK;
}
uint64_t b2mip_kernel_launches(const b2mip* m) { return m ? m->launches : 0; }

b2_box2i b2_impact_on_next_level(int quality, b2_box2i area, int w, int h) { return impactOnNextLevel(quality, area, w, h); }
int b2_tile_areas(const b2_box2i* areas, int n, int max_w, int max_h, b2_box2i* out, int cap) {
    if (max_w <= 0 || max_h <= 0) return 0;
    std::vector<b2_box2i> v;
    tileAreas(areas, n, max_w, max_h, v);
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}
int b2_remove_overlapping_areas(b2_box2i* boxes, int n, b2_box2i* out, int cap) {
    std::vector<b2_box2i> in(boxes, boxes + n), f;
    removeOverlappingAreas(in, f);
    for (int i = 0; i < (int)f.size() && i < cap; ++i) out[i] = f[i];
    return (int)f.size();
}
b2_box2i b2_grow_rect(b2_box2i rect, int margin, int width, int height) { return growRect(rect, margin, width, height); }

struct b2_dirty_rects { DirtyRectList list; std::vector<b2_box2i> pulled; };
int b2_dirty_rects_create(b2_dirty_rects** out) {
    if (!out) return fail(B2_ERR_INVALID, "dirty_rects_create: null");
    *out = new b2_dirty_rects();
    return B2_OK;
}
void b2_dirty_rects_destroy(b2_dirty_rects* d) { delete d; }
int b2_dirty_rects_add(b2_dirty_rects* d, b2_box2i rect) {
    if (!d || !boxSorted(rect)) return fail(B2_ERR_INVALID, "addRect: the rectangle must be sorted (boxlist.d:239)");
    d->list.addRect(rect);
    return B2_OK;
}
int b2_dirty_rects_is_empty(b2_dirty_rects* d) { return (!d || (d->pulled.empty() && d->list.isEmpty())) ? 1 : 0; }
int b2_dirty_rects_pull_all(b2_dirty_rects* d, b2_box2i* out, int cap) {
    if (!d || cap < 0 || (cap > 0 && !out)) return 0;
    d->list.pullAllRectangles(d->pulled);      // rectangles a too-small buffer could not take stay queued here
    const int n = (int)d->pulled.size();
    if (n <= cap) {
        for (int i = 0; i < n; ++i) out[i] = d->pulled[i];
        d->pulled.clear();
    }
    return n;
}
int b2_have_no_overlap(const b2_box2i* boxes, int n) { return (n <= 0 || (boxes && haveNoOverlap(boxes, n))) ? 1 : 0; }

int b2_host_register(void* ptr, size_t bytes) {
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable));
    return B2_OK;
}
int b2_host_unregister(void* ptr) {
    CU(cudaHostUnregister(ptr));
    return B2_OK;
}

}  // extern "C"

// ================================================================================================ b2layer
struct b2layer {
    int device = 0, w = 0, h = 0;
    bool opacity = false;
    DevLevel plane[6];   // diffuse RGBA, depth L16, material RGBA, then the three L8 opacity masks
    uint8_t* alloc = nullptr;
};
static const int kLayerElem[6] = {4, 2, 4, 1, 1, 1};

extern "C" {
int b2layer_create(b2layer** out, int device, int w, int h, int with_opacity) {
    if (!out || w <= 0 || h <= 0) return fail(B2_ERR_INVALID, "b2layer_create: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(B2_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)"); }
    if (device < 0 || device >= ndev) return fail(B2_ERR_INVALID, "bad device index");
    CU(cudaSetDevice(device));
    b2layer* l = new b2layer();
    l->device = device; l->w = w; l->h = h; l->opacity = with_opacity != 0;
    const int n = with_opacity ? 6 : 3;
    size_t total = 0, offs[6] = {0};
    for (int k = 0; k < n; ++k) {
        DevLevel& L = l->plane[k];
        L.w = w; L.h = h; L.y0 = 0; L.rows = h;
        L.pitch = (int)alignUp((size_t)w * kLayerElem[k], 128);
        offs[k] = total;
        total += alignUp((size_t)L.pitch * h, 256);
    }
    if (cudaMalloc((void**)&l->alloc, total) != cudaSuccess) { delete l; return fail(B2_ERR_CUDA, "b2layer_create: cudaMalloc failed"); }
    cudaMemset(l->alloc, 0, total);
    cudaStreamSynchronize(cudaStreamLegacy);
    for (int k = 0; k < n; ++k) l->plane[k].base = l->alloc + offs[k];
    *out = l;
    return B2_OK;
}
void b2layer_destroy(b2layer* l) {
    if (!l) return;
    cudaSetDevice(l->device);
    cudaDeviceSynchronize();
    if (l->alloc) cudaFree(l->alloc);
    delete l;
}
int b2layer_upload(b2layer* l, int plane, const void* host, size_t pitch, const b2_box2i* rects, int n) {
    if (!l || !host || plane < 0 || plane >= (l->opacity ? 6 : 3) || (n > 0 && !rects)) return fail(B2_ERR_INVALID, "b2layer_upload: bad arguments");
    CU(cudaSetDevice(l->device));
    const DevLevel& L = l->plane[plane];
    const int e = kLayerElem[plane];
    for (int k = 0; k < n; ++k) {
        const b2_box2i& r = rects[k];
        if (!boxSorted(r) || r.min_x < 0 || r.min_y < 0 || r.max_x > l->w || r.max_y > l->h) return fail(B2_ERR_INVALID, "b2layer_upload: rectangle outside the layer");
        if (boxEmpty(r)) continue;
        CU(cudaMemcpy2D(L.base + (size_t)r.min_y * L.pitch + (size_t)r.min_x * e, (size_t)L.pitch,
                        (const uint8_t*)host + (size_t)r.min_y * pitch + (size_t)r.min_x * e, pitch, (size_t)boxW(r) * e, (size_t)boxH(r), cudaMemcpyHostToDevice));
    }
    // a synchronous copy from pageable memory may return while the DMA from the staging buffer is still in flight;
    // the consumers (b2pbr_draw_layer) run on non-blocking streams that the legacy stream does not order
    CU(cudaStreamSynchronize(cudaStreamLegacy));
    return B2_OK;
}

// ---- ImageResizer (resizer.cu)
static int resizeCheck(int kind, const void* in, int in_w, int in_h, size_t in_pitch, int out_w, int out_h) {
    if (kind < 0 || kind > 4 || !in || in_w < 1 || in_h < 1 || out_w < 1 || out_h < 1) return fail(B2_ERR_INVALID, "resize: bad arguments");
    const size_t texel = (kind == B2_RESIZE_DEPTH || kind == B2_RESIZE_GENERIC_L16) ? 2 : 4;
    if (in_pitch < (size_t)in_w * texel) return fail(B2_ERR_INVALID, "resize: input pitch smaller than a row");
    if (in_h > 65535 || out_h > 65535) return fail(B2_ERR_INVALID, "resize: more than 65535 rows");
    return B2_OK;
}
int b2_resize_axis_table(int kind, int in_size, int out_size, int* start, int* src, float* coef, int cap) {
    if (kind < 0 || kind > 4 || in_size < 1 || out_size < 1 || cap < 0) return fail(B2_ERR_INVALID, "b2_resize_axis_table: bad arguments") * -1;
    return resize_axis_table(kind, in_size, out_size, start, src, coef, cap);
}
int b2_resize_image(int device, int kind, const void* in, int in_w, int in_h, size_t in_pitch, void* out, int out_w, int out_h, size_t out_pitch) {
    if (int rc = resizeCheck(kind, in, in_w, in_h, in_pitch, out_w, out_h)) return rc;
    const size_t texel = (kind == B2_RESIZE_DEPTH || kind == B2_RESIZE_GENERIC_L16) ? 2 : 4;
    if (!out || out_pitch < (size_t)out_w * texel) return fail(B2_ERR_INVALID, "resize: bad output image");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(B2_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)"); }
    if (device < 0 || device >= ndev) return fail(B2_ERR_INVALID, "bad device index");
    CU(cudaSetDevice(device));
    uint8_t* dOut = nullptr;
    const size_t dPitch = alignUp((size_t)out_w * texel, 128);
    CU(cudaMalloc((void**)&dOut, dPitch * out_h));
    int rc = resize_host_to_device(kind, in, in_w, in_h, in_pitch, dOut, dPitch, out_w, out_h, cudaStreamPerThread);
    cudaError_t e = rc ? cudaSuccess : cudaMemcpy2DAsync(out, out_pitch, dOut, dPitch, (size_t)out_w * texel, (size_t)out_h, cudaMemcpyDeviceToHost, cudaStreamPerThread);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);
    cudaFree(dOut);
    if (rc) return fail(B2_ERR_CUDA, cudaGetErrorString((cudaError_t)rc));
    CU(e);
    return B2_OK;
}
int b2layer_resize_plane(b2layer* l, int plane, int kind, const void* in, int in_w, int in_h, size_t in_pitch) {
    if (!l || plane < 0 || plane > 2) return fail(B2_ERR_INVALID, "b2layer_resize_plane: bad arguments");
    if (int rc = resizeCheck(kind, in, in_w, in_h, in_pitch, l->w, l->h)) return rc;
    const bool l16 = kind == B2_RESIZE_DEPTH || kind == B2_RESIZE_GENERIC_L16;
    if (l16 != (plane == 1)) return fail(B2_ERR_INVALID, "b2layer_resize_plane: the depth plane takes the L16 kinds, the others the RGBA kinds");
    CU(cudaSetDevice(l->device));
    const DevLevel& L = l->plane[plane];
    if (int rc = resize_host_to_device(kind, in, in_w, in_h, in_pitch, L.base, (size_t)L.pitch, l->w, l->h, cudaStreamPerThread))
        return fail(B2_ERR_CUDA, cudaGetErrorString((cudaError_t)rc));
    return B2_OK;
}
}  // extern "C"

// ================================================================================================ b2pbr
struct TraceEvent { std::string name; cudaEvent_t beg, end; };

struct b2pbr {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool ownStream = false;
    int W = 0, H = 0, tileW = 64, tileH = 64;
    bool banded = false;
    int bandY0 = 0, bandY1 = 0;
    // rows [lo, hi) of every pyramid level this band has to hold valid data for (banded objects only)
    int needD[kMaxDiffuseLevels][2], needZ[kMaxDepthLevels][2];
    b2mip *diffuse = nullptr, *material = nullptr, *depth = nullptr, *sky = nullptr;
    cudaArray_t skyArr = nullptr;                 // every skybox level once more in one CUDA array (texture gather needs one)
    cudaTextureObject_t skyTex = 0, skyTexLin = 0;
    int skyOx[kMaxSkyLevels], skyOy[kMaxSkyLevels];   // atlas position of texel (0, 0) of each level
    DevLevel out{}; uint8_t* outAlloc = nullptr;
    DevLevel swz{}; uint8_t* swzAlloc = nullptr;
    DevLevel win{}; uint8_t* winAlloc = nullptr;     // device twin of the window buffer (_resizedBuffer, graphics.d:1146)
    uint8_t* dLut = nullptr; bool lutOn = false;      // UIColorCorrection transfer tables (3 x 256)
    b2pbr_params params;
    uint32_t* dRsqrt = nullptr;
    float* dExpTab = nullptr;
    // device work lists, cached per (area list, kind, strip cut): a caller that alternates between a few lists (the
    // interior and the boundary tiles of a row band; a recorded graph and live calls) never re-uploads them
    struct WorkList { std::vector<Box> key; int kind = -1, tail1 = 0, tail2 = 0; Box* d = nullptr; int cap = 0, n = 0; uint64_t used = 0; bool pinned = false; };
    std::vector<WorkList> workLists;
    uint64_t workClock = 0;
    Box* dBoxes = nullptr;           // the current list (points into workLists)
    Box* hBoxes = nullptr; int hBoxCap = 0;   // pinned staging
    cudaEvent_t boxCopied = nullptr;
    cudaStream_t sIn = nullptr, sOut = nullptr;   // copy streams of the pipelined host path
    cudaEvent_t evMatFork = nullptr, evMatDone = nullptr;   // small-frame host call: the material upload runs beside the mips
    std::vector<cudaGraphExec_t> graphs;          // frames recorded with b2pbr_graph_begin / _end
    std::vector<uint64_t> graphLaunches;          // kernels each recorded graph launches
    uint64_t launchesAtCapture = 0;
    bool capturing = false;
    cudaStream_t sAux = nullptr;                  // the depth pyramid is regenerated next to the diffuse one
    cudaStream_t sAuxHi = nullptr;                // ... and on this high-priority twin when the caller's stream is a priority stream (b2band's halo pass)
    bool useAuxHi = false;
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    std::vector<cudaEvent_t> pipeEvents;
    uint8_t* dlStaging = nullptr; size_t dlStagingBytes = 0;   // linear device buffer of b2pbr_download
    int tailRow = 1 << 30;           // areas ending below this row are cut into 8-row strips (pipelined host call)
    int tailRow2 = 1 << 30;          // ... and below this one into 4-row strips
    int nDeviceBoxes = 0;
    int mode = B2_MODE_FAST;
    uint64_t launches = 0;
    bool profile = false;
    std::vector<TraceEvent> trace;
    float shadowW[11]; float shadowInvTotal;
};

// Recorded graphs bake device addresses (maps, skybox, output) and the pass parameters into their kernel arguments:
// whatever frees or changes those destroys the recordings; b2pbr_graph_launch then fails with B2_ERR_STATE.
static void pbrInvalidateGraphs(b2pbr* c) {
    for (auto& g : c->graphs)
        if (g) { cudaGraphExecDestroy(g); g = nullptr; }
    for (auto& wl : c->workLists) wl.pinned = false;
}

static void pbrFreeSkyTextures(b2pbr* c) {
    if (c->skyTex) cudaDestroyTextureObject(c->skyTex);
    if (c->skyTexLin) cudaDestroyTextureObject(c->skyTexLin);
    c->skyTexLin = 0;
    if (c->skyArr) cudaFreeArray(c->skyArr);
    c->skyTex = 0; c->skyArr = nullptr;
}

static void pbrConstants(b2pbr* c) {
    // PassObliqueShadowLight's compile-time tables (gui/dplug/gui/legacypbr.d:469-490): fallOff ^^ k and
    // the normalisation, evaluated in double and narrowed once.
    const float fallOff = 0.78f;
    for (int s = 0; s < 11; ++s) c->shadowW[s] = (float)std::pow((double)fallOff, s);
    const float total = (float)((1.0 - std::pow((double)fallOff, 11)) / (1.0 - (double)fallOff) - 1);
    c->shadowInvTotal = (float)(1 / ((double)1.7f * (double)total));
}

static void normalized3(float x, float y, float z, float out[3]) {  // vec3f.normalized, math/dplug/math/vector.d:344-376
    float s = 0;
    s += x * x; s += y * y; s += z * z;
    float inv = 1 / std::sqrt(s);
    out[0] = x * inv; out[1] = y * inv; out[2] = z * inv;
}

// ---- which rows of each level does a band [y0, y1) read?  (host replica of the samplers' row arithmetic)
static void linRows(int l, int j, int h, int& r0, int& r1) {   // Mipmap.linearSample, mipmap.d:309-339
    int n = 2 * j + 1 - (1 << l);
    if (n < 0) n = 0;
    r0 = n >> (l + 1);
    if (r0 > h - 1) r0 = h - 1;
    r1 = r0 + 1 > h - 1 ? h - 1 : r0 + 1;
}
static void cubRows(int l, int j, int h, int& r0, int& r1) {   // Mipmap.cubicSample, mipmap.d:182-193
    int ib = (2 * j + 1 - (1 << l) + (2 << l)) >> (l + 1);
    r0 = ib - 2 < 0 ? 0 : (ib - 2 > h - 1 ? h - 1 : ib - 2);
    r1 = ib + 1 < 0 ? 0 : (ib + 1 > h - 1 ? h - 1 : ib + 1);
}
static void unite(int* need, int lo, int hi) {  // half-open [lo, hi)
    if (need[0] == need[1]) { need[0] = lo; need[1] = hi; return; }
    if (lo < need[0]) need[0] = lo;
    if (hi > need[1]) need[1] = hi;
}
// rows of the previous level a generateLevel over dst rows [a, b) reads (mipmap.d:683-686, 906-919)
static void genSourceRows(bool cubic, int a, int b, int hs, int& lo, int& hi) {
    if (cubic) { lo = 2 * a - 1 < 0 ? 0 : 2 * a - 1; hi = (2 * (b - 1) + 2 > hs - 1 ? hs - 1 : 2 * (b - 1) + 2) + 1; }
    else { lo = 2 * a; hi = 2 * (b - 1) + 2; }
}
// Pure host arithmetic (no device): rows [lo, hi) of every level a band [y0, y1) of a W x H canvas must hold.
static void computeBandNeeds(int W, int H, int y0, int y1, int needD[kMaxDiffuseLevels][2], int needZ[kMaxDepthLevels][2]) {
    const int nD = countLevels(5, W, H), nZ = countLevels(4, W, H);   // graphics.d:1118-1119
    for (int l = 0; l < kMaxDiffuseLevels; ++l) needD[l][0] = needD[l][1] = 0;
    for (int l = 0; l < kMaxDepthLevels; ++l) needZ[l][0] = needZ[l][1] = 0;
    if (y1 <= y0 || nD == 0) return;
    int a, b, e, f;
    // what the passes sample (level clamp: levels beyond the pyramid read the last one)
    for (int l = 1; l <= 5; ++l) {
        const int lv = l < nD ? l : nD - 1;
        if (lv == 0) continue;
        const int h = H >> lv;
        if (l <= 3) { linRows(lv, y0, h, a, b); linRows(lv, y1 - 1, h, e, f); }
        else { cubRows(lv, y0, h, a, b); cubRows(lv, y1 - 1, h, e, f); }
        unite(needD[lv], a, f + 1);
    }
    for (int l = 1; l <= 4; ++l) {
        const int lv = l < nZ ? l : nZ - 1;
        if (lv == 0) continue;
        const int h = H >> lv;
        linRows(lv, y0, h, a, b); linRows(lv, y1 - 1, h, e, f);
        unite(needZ[lv], a, f + 1);
    }
    // level 0: own rows, the shadow reach (10 rows up) and the normal stencil (1 row each way)
    unite(needD[0], y0, y1);
    unite(needZ[0], y0 - 10 < 0 ? 0 : y0 - 10, y1 + 1 > H ? H : y1 + 1);
    // what generating those rows reads, top level first
    for (int l = nD - 1; l >= 1; --l) {
        if (needD[l][0] == needD[l][1]) continue;
        int lo, hi;
        genSourceRows(l >= 2, needD[l][0], needD[l][1], H >> (l - 1), lo, hi);  // graphics.d:1380-1384
        unite(needD[l - 1], lo, hi);
    }
    for (int l = nZ - 1; l >= 1; --l) {
        if (needZ[l][0] == needZ[l][1]) continue;
        int lo, hi;
        genSourceRows(l >= 3, needZ[l][0], needZ[l][1], H >> (l - 1), lo, hi);   // graphics.d:1414
        unite(needZ[l - 1], lo, hi);
    }
}

static int pbrBandNeeds(b2pbr* c) {
    computeBandNeeds(c->W, c->H, c->bandY0, c->bandY1, c->needD, c->needZ);
    // the band objects must store all of it
    for (int l = 0; l < (int)c->diffuse->levels.size(); ++l) {
        const DevLevel& L = c->diffuse->levels[l];
        if (c->needD[l][0] < L.y0 || c->needD[l][1] > L.y0 + L.rows) return fail(B2_ERR_STATE, "band: diffuse halo too small");
    }
    for (int l = 0; l < (int)c->depth->levels.size(); ++l) {
        const DevLevel& L = c->depth->levels[l];
        if (c->needZ[l][0] < L.y0 || c->needZ[l][1] > L.y0 + L.rows) return fail(B2_ERR_STATE, "band: depth halo too small");
    }
    return B2_OK;
}

static void pbrFreeMaps(b2pbr* c) {
    b2mip_destroy(c->diffuse); b2mip_destroy(c->material); b2mip_destroy(c->depth);
    c->diffuse = c->material = c->depth = nullptr;
    if (c->outAlloc) cudaFree(c->outAlloc);
    if (c->swzAlloc) cudaFree(c->swzAlloc);
    c->outAlloc = c->swzAlloc = nullptr;
}

static void traceBegin(b2pbr* c, const char* name) {
    if (!c->profile) return;
    TraceEvent t; t.name = name;
    cudaEventCreate(&t.beg); cudaEventCreate(&t.end);
    cudaEventRecord(t.beg, c->stream);
    c->trace.push_back(t);
}
static void traceEnd(b2pbr* c) {
    if (!c->profile || c->trace.empty()) return;
    cudaEventRecord(c->trace.back().end, c->stream);
}

// Rows per warp strip.  A warp walks down its strip row by row, so the strip height is the length of the kernel's
// critical path: full-height strips (32 rows) amortise the per-strip staging best and are used whenever they already
// fill the GPU (148 SMs x 16 resident warps); small canvases (the 620 x 330 default UI has 240 such strips) get
// shorter strips so that the same pixels are spread over more warps.
static int stripRowsOfArea(const b2_box2i& a, int R, int tailRow, int tailRow2) {
    if (a.max_y > tailRow2 && R > 4) return 4;
    return (a.max_y > tailRow && R > 8) ? 8 : R;
}
static int stripRowsFor(const b2_box2i* areas, int n);
struct b2pbr;
static int stripRowsOfAreaAt(const b2pbr* c, int k, int n, const b2_box2i& a, int R0);
static int stripRowsFor(const b2_box2i* areas, int n) {
    static const int forced = getenv("B2_STRIP_ROWS") ? atoi(getenv("B2_STRIP_ROWS")) : 0;   // tuning aid
    if (forced >= 1 && forced <= kQuadrantRows) return forced;
    const int full = kQuadrantRows;
    // Wave model: a launch runs ceil(strips / resident warps) waves, each as long as one strip (R rows + a staging cost
    // worth ~6 rows).  Long lists take full-height strips (least staging per pixel); short ones — the 620 x 330 default
    // UI, the boundary tiles of a row band — are cut finer when that fills the last wave better.
    static const long long slots = 148ll * (getenv("B2_PBR_PAIR") && atoi(getenv("B2_PBR_PAIR")) == 0 ? 20 : 16);   // resident warps: k_pbr_pair 4 CTAs x 4 warps per SM (k_pbr_tex: 5 x 4)
    int best = full; long long bestCost = -1;
    for (int R = full; R >= 4; R >>= 1) {
        long long strips = 0;
        for (int k = 0; k < n; ++k) strips += (long long)((boxH(areas[k]) + R - 1) / R) * ((boxW(areas[k]) + 31) / 32);
        const long long cost = ((strips + slots - 1) / slots) * (R + 6);
        if (bestCost < 0 || cost < bestCost) { bestCost = cost; best = R; }
    }
    return best;
}

// Strip height of area k of an n-area list.  Tail balancing: CTAs are dispatched in list order and a frame is only a
// few waves of them (3.4 at 4K: the last wave would run at 45 % occupancy), so the last 8 % of a long list are cut into
// 8-row strips that fill the slots the full-height strips leave empty (4K lighting 0.239 -> 0.234 ms; 15 % and 25 %
// measured slower, as did 16-row strips everywhere: 0.250 ms).
static int stripRowsOfAreaAt(const b2pbr* c, int k, int n, const b2_box2i& a, int R0) {
    static const int tailPct = getenv("B2_TAIL_PCT") ? atoi(getenv("B2_TAIL_PCT")) : 8;   // tuning aid
    int R = stripRowsOfArea(a, R0, c->tailRow, c->tailRow2);
    if (R0 == kQuadrantRows && k >= n - n * tailPct / 100 && R > 8) R = 8;
    return R;
}

// Uploads the work list of a launch.  kind 0: the caller's areas as they are (one CTA each, k_pbr_exact); kind 2: the
// areas cut into CTA boxes of 64 columns x 2R rows for k_pbr_tex.  The list is kept on the device and
// re-used while the caller passes the same areas (a full-frame tile list never changes between frames).
static int pbrUploadBoxes(b2pbr* c, const b2_box2i* areas, int n, int kind) {
    for (auto& wl : c->workLists)
        if (wl.kind == kind && wl.tail1 == c->tailRow && wl.tail2 == c->tailRow2 && (int)wl.key.size() == n && n > 0 &&
            memcmp(wl.key.data(), areas, sizeof(Box) * n) == 0) {
            wl.used = ++c->workClock;
            if (c->capturing) wl.pinned = true;   // a recorded graph reads this list at every replay
            c->dBoxes = wl.d; c->nDeviceBoxes = wl.n;
            return B2_OK;
        }
    if (c->capturing) return fail(B2_ERR_STATE, "graph capture: composite this area list once before recording it (its work list is uploaded then)");
    std::vector<Box> work;
    if (kind == 0) work.assign((const Box*)areas, (const Box*)areas + n);
    else {
        // CTA boxes of 64 columns x 2R rows, whose four warps take the four 32 x R quadrants
        const int R0 = stripRowsFor(areas, n);
        const int bw = 64;
        for (int k = 0; k < n; ++k) {
            const b2_box2i& a = areas[k];
            const int R = stripRowsOfAreaAt(c, k, n, a, R0) * 2;
            for (int y = a.min_y; y < a.max_y; y += R)
                for (int x = a.min_x; x < a.max_x; x += bw)
                    work.push_back(Box{x, y, x + bw < a.max_x ? x + bw : a.max_x, y + R < a.max_y ? y + R : a.max_y});
        }
    }
    const int m = (int)work.size();
    // slot: a free one (up to 4), else the least recently used that no graph refers to, else a new one
    b2pbr::WorkList* slot = nullptr;
    if (c->workLists.size() < 4) { c->workLists.emplace_back(); slot = &c->workLists.back(); }
    else {
        for (auto& wl : c->workLists)
            if (!wl.pinned && (!slot || wl.used < slot->used)) slot = &wl;
        if (!slot) { c->workLists.emplace_back(); slot = &c->workLists.back(); }
    }
    if (m > slot->cap) {
        const int cap = m < 16384 ? 16384 : m * 2;
        if (slot->d) { CU(cudaStreamSynchronize(c->stream)); cudaFree(slot->d); slot->d = nullptr; }
        CU(cudaMalloc((void**)&slot->d, sizeof(Box) * cap));
        slot->cap = cap;
    }
    if (m > c->hBoxCap) {
        CU(cudaEventSynchronize(c->boxCopied));
        if (c->hBoxes) cudaFreeHost(c->hBoxes);
        c->hBoxCap = m < 16384 ? 16384 : m * 2;
        CU(cudaMallocHost((void**)&c->hBoxes, sizeof(Box) * c->hBoxCap));
    }
    if (m > 0) {
        CU(cudaEventSynchronize(c->boxCopied));  // staging buffer free again?
        memcpy(c->hBoxes, work.data(), sizeof(Box) * m);
        CU(cudaMemcpyAsync(slot->d, c->hBoxes, sizeof(Box) * m, cudaMemcpyHostToDevice, c->stream));
        CU(cudaEventRecord(c->boxCopied, c->stream));
    }
    slot->key.assign((const Box*)areas, (const Box*)areas + n);
    slot->kind = kind; slot->tail1 = c->tailRow; slot->tail2 = c->tailRow2;
    slot->n = m; slot->used = ++c->workClock;
    c->dBoxes = slot->d; c->nDeviceBoxes = m;
    return B2_OK;
}

// `workFirst`/`workCount` (optional) restrict the launch to a sub-range of the device work list built from `areas`
// (used by the pipelined host path, which uploads the list of the whole frame once and launches it in slices).
static int pbrComposite(b2pbr* c, const b2_box2i* areas, int n, b2mip* diffuse, b2mip* material, b2mip* depth,
                        int workFirst = 0, int workCount = -1, bool* usedFast = nullptr) {
    if (!c->outAlloc) return fail(B2_ERR_STATE, "compositeTile before resizeBuffers (compositor.d:41)");
    if (!diffuse) diffuse = c->diffuse;
    if (!material) material = c->material;
    if (!depth) depth = c->depth;
    if (diffuse->color != B2_COLOR_RGBA8 || material->color != B2_COLOR_RGBA8 || depth->color != B2_COLOR_L16)
        return fail(B2_ERR_INVALID, "compositeTile: wrong map pixel types");
    if (diffuse->W != c->W || diffuse->H != c->H || material->W != c->W || material->H != c->H || depth->W != c->W || depth->H != c->H)
        return fail(B2_ERR_INVALID, "compositeTile: map size differs from resizeBuffers size");
    if (n <= 0) return B2_OK;
    for (int k = 0; k < n; ++k) {
        const b2_box2i& a = areas[k];
        if (!boxSorted(a) || a.min_x < 0 || a.min_y < 0 || a.max_x > c->W || a.max_y > c->H)
            return fail(B2_ERR_INVALID, "compositeTile: area outside the frame");
        if (a.min_y < c->out.y0 || a.max_y > c->out.y0 + c->out.rows)
            if (!boxEmpty(a)) return fail(B2_ERR_INVALID, "compositeTile: area outside this band");
    }
    CU(cudaSetDevice(c->device));
    // the throughput kernel assumes full pyramids (diffuse levels 0..5, depth 0..4: any canvas >= 32x32);
    // smaller canvases are shaded by the exact kernel
    const bool fast = c->mode != B2_MODE_EXACT && (int)diffuse->levels.size() >= kMaxDiffuseLevels && (int)depth->levels.size() >= kMaxDepthLevels;
    int rc = pbrUploadBoxes(c, areas, n, fast ? 2 : 0);
    if (rc) return rc;
    if (usedFast) *usedFast = fast;
    if (c->nDeviceBoxes == 0) return B2_OK;
    if (workCount < 0) workCount = c->nDeviceBoxes - workFirst;
    if (workCount == 0) return B2_OK;
    PbrArgs A;
    memset(&A, 0, sizeof(A));
    A.nDiffuse = (int)diffuse->levels.size();
    A.nDepth = (int)depth->levels.size();
    for (int l = 0; l < A.nDiffuse && l < kMaxDiffuseLevels; ++l) A.diffuse[l] = diffuse->levels[l];
    for (int l = 0; l < A.nDepth && l < kMaxDepthLevels; ++l) A.depth[l] = depth->levels[l];
    if (A.nDiffuse > kMaxDiffuseLevels) A.nDiffuse = kMaxDiffuseLevels;
    if (A.nDepth > kMaxDepthLevels) A.nDepth = kMaxDepthLevels;
    A.material = material->levels[0];
    A.nSky = c->sky ? (int)c->sky->levels.size() : 0;
    for (int l = 0; l < A.nSky; ++l) A.sky[l] = c->sky->levels[l];
    A.out = c->out;
    A.W = c->W; A.H = c->H;
    A.boxes = c->dBoxes + workFirst; A.nBoxes = workCount;
    A.rsqrtTab = c->dRsqrt; A.expTab = c->dExpTab;
    const b2pbr_params& p = c->params;
    memcpy(A.light1, p.light1_color, 12); memcpy(A.light2Dir, p.light2_dir, 12); memcpy(A.light2Col, p.light2_color, 12);
    memcpy(A.light3Dir, p.light3_dir, 12); memcpy(A.light3Col, p.light3_color, 12);
    A.ambient = p.ambient_light; A.skyAmount = p.skybox_amount; A.toneThr = p.tonemap_threshold; A.toneRatio = p.tonemap_ratio;
    A.passMask = p.pass_active_mask;
    {
        const float d255 = 1.0f / 255.0f;
        for (int k = 0; k < 3; ++k) { A.light1S[k] = p.light1_color[k] * d255; A.light2ColS[k] = p.light2_color[k] * d255; A.light3ColS[k] = p.light3_color[k] * d255 * d255; }
        A.ambientS = p.ambient_light * d255;
        A.skyAmountS = p.skybox_amount * d255 * d255;
    }
    memcpy(A.shadowW, c->shadowW, sizeof(A.shadowW));
    A.shadowInvTotal = c->shadowInvTotal;
    A.invW = 1.0f / (float)c->W; A.invH = 1.0f / (float)c->H;
    if (c->sky) {
        A.skyPixels = (float)(c->sky->W * c->sky->H);
        A.skyFX = (float)c->sky->W - 1.0f;
        A.skyFY = (float)c->sky->H - 1.0f;
        A.texSky = c->skyTex; A.texSkyLin = c->skyTexLin;
        for (int l = 0; l < A.nSky; ++l) {
            A.skyMaxX[l] = (float)(c->sky->levels[l].w - 1);
            A.skyMaxY[l] = (float)(c->sky->levels[l].h - 1);
            A.skyOx[l] = (float)c->skyOx[l];
            A.skyOy[l] = (float)c->skyOy[l];
            const float sc = ldexpf(1.0f, -l);
            A.skyCX[l] = make_float4(A.skyOx[l], A.skyOx[l] + 0.5f, A.skyOx[l] + A.skyMaxX[l] + 0.5f, sc);
            A.skyCY[l] = make_float4(A.skyOy[l], A.skyOy[l] + 0.5f, A.skyOy[l] + A.skyMaxY[l] + 0.5f, sc);
        }
    }
    for (int l = 0; l < A.nDiffuse; ++l) A.texD[l] = diffuse->tex[l];
    for (int l = 0; l < A.nDepth; ++l) A.texZ[l] = depth->tex[l];
    traceBegin(c, "PBR compositeTile");
    static const int pairKernel = getenv("B2_PBR_PAIR") ? atoi(getenv("B2_PBR_PAIR")) : 1;   // tuning aid: 0 = one row per lane (k_pbr_tex)
    A.negZero2 = 0x8000000080000000ull;
    if (fast) launch_pbr_tex(A, c->mode == B2_MODE_FAST_SKY_GATHER, pairKernel ? 0 : 1, c->stream);
    else launch_pbr_exact(A, A.nBoxes, c->stream);
    c->launches++;
    traceEnd(c);
    CU(cudaGetLastError());
    return B2_OK;
}

extern "C" {

void b2pbr_default_params(b2pbr_params* p) {
    if (!p) return;
    for (int k = 0; k < 3; ++k) p->light1_color[k] = 0.25f * 1.3f;   // legacypbr.d:453
    normalized3(0.0f, 1.0f, 0.1f, p->light2_dir);                    // legacypbr.d:626
    for (int k = 0; k < 3; ++k) p->light2_color[k] = 0.481f;         // legacypbr.d:629
    normalized3(0.0f, 1.0f, 0.1f, p->light3_dir);                    // legacypbr.d:676
    for (int k = 0; k < 3; ++k) p->light3_color[k] = 0.26f;          // legacypbr.d:679
    p->ambient_light = 0.08125f;                                     // legacypbr.d:391
    p->skybox_amount = 0.52f;                                        // legacypbr.d:843
    p->tonemap_threshold = 1.0f;                                     // legacypbr.d:1045
    p->tonemap_ratio = 0.3f;                                         // legacypbr.d:1049
    p->pass_active_mask = 0xffu;
}

static int pbrCreate(b2pbr** out, int device, bool banded, int y0, int y1) {
    if (!out) return fail(B2_ERR_INVALID, "b2pbr_create: null out");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(B2_ERR_NO_DEVICE, "no CUDA device (this library has no CPU path)"); }
    if (device < 0 || device >= ndev) return fail(B2_ERR_INVALID, "bad device index");
    CU(cudaSetDevice(device));
    b2pbr* c = new b2pbr();
    c->device = device; c->banded = banded; c->bandY0 = y0; c->bandY1 = y1;
    b2pbr_default_params(&c->params);
    pbrConstants(c);
    CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->ownStream = true;
    CU(cudaEventCreateWithFlags(&c->boxCopied, cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&c->sAux, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
    CU(cudaMalloc((void**)&c->dRsqrt, sizeof(kRsqrtssTable)));
    CU(cudaMemcpy(c->dRsqrt, kRsqrtssTable, sizeof(kRsqrtssTable), cudaMemcpyHostToDevice));
    float expTab[256];
    for (int r = 0; r < 256; ++r) {  // PassSpecularLight._exponentTable, legacypbr.d:696-702
        float arg = (1 - r / 255.0f) * 5.5f;
        float t = 0.8f * (float)std::exp((double)arg);
        t *= 2.8f;
        expTab[r] = t;
    }
    CU(cudaMalloc((void**)&c->dExpTab, sizeof(expTab)));
    CU(cudaMemcpy(c->dExpTab, expTab, sizeof(expTab), cudaMemcpyHostToDevice));
    CU(cudaStreamSynchronize(cudaStreamLegacy));   // pageable sources: see b2layer_upload
    *out = c;
    return B2_OK;
}
int b2pbr_create(b2pbr** out, int device) { return pbrCreate(out, device, false, 0, 0); }
int b2pbr_create_band(b2pbr** out, int device, int y0, int y1) {
    if (y0 < 0 || y1 < y0) return fail(B2_ERR_INVALID, "b2pbr_create_band: bad rows");
    return pbrCreate(out, device, true, y0, y1);
}
void b2pbr_destroy(b2pbr* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    pbrFreeMaps(c);
    pbrFreeSkyTextures(c);
    b2mip_destroy(c->sky);
    if (c->winAlloc) cudaFree(c->winAlloc);
    if (c->dLut) cudaFree(c->dLut);
    if (c->dRsqrt) cudaFree(c->dRsqrt);
    if (c->dExpTab) cudaFree(c->dExpTab);
    for (auto& wl : c->workLists)
        if (wl.d) cudaFree(wl.d);
    if (c->dlStaging) cudaFree(c->dlStaging);
    if (c->hBoxes) cudaFreeHost(c->hBoxes);
    if (c->boxCopied) cudaEventDestroy(c->boxCopied);
    for (auto g : c->graphs) if (g) cudaGraphExecDestroy(g);
    if (c->sAux) cudaStreamDestroy(c->sAux);
    if (c->sAuxHi) cudaStreamDestroy(c->sAuxHi);
    if (c->evFork) cudaEventDestroy(c->evFork);
    if (c->evJoin) cudaEventDestroy(c->evJoin);
    if (c->evMatFork) cudaEventDestroy(c->evMatFork);
    if (c->evMatDone) cudaEventDestroy(c->evMatDone);
    if (c->sIn) cudaStreamDestroy(c->sIn);
    if (c->sOut) cudaStreamDestroy(c->sOut);
    for (auto e : c->pipeEvents) cudaEventDestroy(e);
    for (auto& t : c->trace) { cudaEventDestroy(t.beg); cudaEventDestroy(t.end); }
    if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}
int b2pbr_set_stream(b2pbr* c, void* s) {
    if (!c) return fail(B2_ERR_INVALID, "null compositor");
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->ownStream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)s;
    c->ownStream = false;
    for (b2mip* m : {c->diffuse, c->material, c->depth, c->sky})
        if (m) b2mip_set_stream(m, s);
    return B2_OK;
}

int b2pbr_resize(b2pbr* c, int width, int height, int tw, int th) {
    if (!c || width <= 0 || height <= 0 || tw <= 0 || th <= 0) return fail(B2_ERR_INVALID, "resizeBuffers: bad arguments");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    pbrInvalidateGraphs(c);
    pbrFreeMaps(c);
    for (auto& wl : c->workLists) { wl.key.clear(); wl.kind = -1; wl.pinned = false; }   // strip cuts depend on the size
    c->W = width; c->H = height; c->tileW = tw; c->tileH = th;
    int y0 = 0, y1 = height, dUp = 0, dDown = 0, zUp = 0, zDown = 0;
    if (c->banded) {
        if (c->bandY1 > height) return fail(B2_ERR_INVALID, "resizeBuffers: band outside the canvas");
        y0 = c->bandY0; y1 = c->bandY1;
        // rows of level 0 a band needs beyond its own for the level-5 bicubic bloom
        // (94) and the level-4 AO pyramid / 10-row shadow reach (28)
        dUp = dDown = 96; zUp = zDown = 32;
    }
    int rc;
    // GUIGraphics.doResize, graphics.d:1118-1128
    if ((rc = mipCreate(&c->diffuse, c->device, B2_COLOR_RGBA8, 5, width, height, y0, y1, dUp, dDown))) return rc;
    if ((rc = mipCreate(&c->depth, c->device, B2_COLOR_L16, 4, width, height, y0, y1, zUp, zDown))) return rc;
    if ((rc = mipCreate(&c->material, c->device, B2_COLOR_RGBA8, 0, width, height, y0, y1, 0, 0))) return rc;
    for (b2mip* m : {c->diffuse, c->material, c->depth}) {
        cudaStreamSynchronize(m->stream);
        b2mip_set_stream(m, c->stream);
    }
    // _compositedBuffer (graphics.d:1136)
    c->out.w = width; c->out.h = height; c->out.pitch = (int)alignUp((size_t)width * 4, 128);
    c->out.y0 = y0; c->out.rows = y1 - y0;
    CU(cudaMalloc((void**)&c->outAlloc, (size_t)c->out.pitch * (size_t)(c->out.rows > 0 ? c->out.rows : 1)));
    CU(cudaMemsetAsync(c->outAlloc, 0, (size_t)c->out.pitch * (size_t)(c->out.rows > 0 ? c->out.rows : 1), c->stream));
    c->out.base = c->outAlloc;
    c->swz = c->out;
    CU(cudaMalloc((void**)&c->swzAlloc, (size_t)c->out.pitch * (size_t)(c->out.rows > 0 ? c->out.rows : 1)));
    c->swz.base = c->swzAlloc;
    if (c->banded) return pbrBandNeeds(c);
    return B2_OK;
}

int b2pbr_set_params(b2pbr* c, const b2pbr_params* p) {
    if (!c || !p) return fail(B2_ERR_INVALID, "set_params: null");
    if (memcmp(&c->params, p, sizeof(*p)) != 0) pbrInvalidateGraphs(c);
    c->params = *p;
    return B2_OK;
}
int b2pbr_set_mode(b2pbr* c, int mode) {
    if (!c || (mode != B2_MODE_FAST && mode != B2_MODE_EXACT && mode != B2_MODE_FAST_SKY_GATHER)) return fail(B2_ERR_INVALID, "set_mode: bad arguments");
    if (mode != c->mode) pbrInvalidateGraphs(c);
    c->mode = mode;
    return B2_OK;
}
int b2pbr_get_params(const b2pbr* c, b2pbr_params* p) {
    if (!c || !p) return fail(B2_ERR_INVALID, "get_params: null");
    *p = c->params;
    return B2_OK;
}

int b2pbr_set_skybox(b2pbr* c, const uint8_t* rgba, int w, int h, size_t pitch) {
    if (!c) return fail(B2_ERR_INVALID, "null compositor");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    pbrInvalidateGraphs(c);
    pbrFreeSkyTextures(c);
    b2mip_destroy(c->sky);
    c->sky = nullptr;
    if (!rgba) return B2_OK;
    if (w <= 0 || h <= 0) return fail(B2_ERR_INVALID, "setSkybox: empty image");
    int rc = b2mip_create(&c->sky, c->device, B2_COLOR_RGBA8, 12, w, h);  // legacypbr.d:868
    if (rc) return rc;
    cudaStreamSynchronize(c->sky->stream);
    b2mip_set_stream(c->sky, c->stream);
    b2_box2i whole{0, 0, w, h};
    if ((rc = b2mip_upload(c->sky, 0, rgba, pitch, &whole, 1))) return rc;
    CU(cudaStreamSynchronize(c->stream));  // the caller may free its image right after this call
    if ((rc = b2mip_generate_mipmaps(c->sky, B2_QUALITY_BOX))) return rc;  // legacypbr.d:869
    c->launches += c->sky->launches;
    // every level once more in ONE CUDA array with the gather flag (an atlas, each level with a replicated one-texel
    // border): the lighting kernel reads the 2 x 2 footprints of the trilinear sample with texture gather (exact texels;
    // the filter weights stay in FP32 in the kernel) through a single, warp-uniform texture handle
    {
        const int nl = (int)c->sky->levels.size();
        int aw = 0, ah = 0;
        sky_atlas_layout(c->sky->levels.data(), nl, c->skyOx, c->skyOy, &aw, &ah);
        uint32_t* lin = nullptr;
        CU(cudaMalloc((void**)&lin, (size_t)aw * ah * 4));
        launch_sky_atlas(c->sky->levels.data(), nl, c->skyOx, c->skyOy, aw, ah, lin, c->stream);
        c->launches++;
        CU(cudaGetLastError());
        const cudaChannelFormatDesc fd = cudaCreateChannelDesc(8, 8, 8, 8, cudaChannelFormatKindUnsigned);
        CU(cudaMallocArray(&c->skyArr, &fd, (size_t)aw, (size_t)ah, cudaArrayTextureGather));
        CU(cudaMemcpy2DToArrayAsync(c->skyArr, 0, 0, lin, (size_t)aw * 4, (size_t)aw * 4, (size_t)ah, cudaMemcpyDeviceToDevice, c->stream));
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeArray;
        rd.res.array.array = c->skyArr;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModePoint;
        td.readMode = cudaReadModeNormalizedFloat;
        td.normalizedCoords = 0;
        CU(cudaCreateTextureObject(&c->skyTex, &rd, &td, nullptr));
        td.filterMode = cudaFilterModeLinear;
        CU(cudaCreateTextureObject(&c->skyTexLin, &rd, &td, nullptr));
        CU(cudaStreamSynchronize(c->stream));
        cudaFree(lin);
    }
    CU(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

b2mip* b2pbr_map(b2pbr* c, int which) {
    if (!c) return nullptr;
    switch (which) {
        case B2_MAP_DIFFUSE: return c->diffuse;
        case B2_MAP_MATERIAL: return c->material;
        case B2_MAP_DEPTH: return c->depth;
    }
    return nullptr;
}
b2mip* b2pbr_skybox(b2pbr* c) { return c ? c->sky : nullptr; }
int b2pbr_output(b2pbr* c, b2_imageref* out) {
    if (!c || !out || !c->outAlloc) return fail(B2_ERR_STATE, "b2pbr_output: no buffers");
    out->w = c->out.w; out->h = c->out.h; out->pitch = (size_t)c->out.pitch; out->pixels = c->out.base;
    return B2_OK;
}

// One pyramid of regenerateMipmaps (graphics.d:1378-1419): `q[l-1]` = quality of level l, `need` = rows a band has to
// regenerate per level (nullptr: unbanded).
//
// The reference walks level-major: every rectangle of level l before level l + 1.  A rectangle's whole chain may run
// before the next rectangle's (rect-major: fused launches, up to three levels each) exactly when no rectangle's
// level-l source footprint can see another rectangle's level-l update: the chains, grown by the cubic reach, are
// pairwise disjoint on every level.  One rectangle (a full redraw, most interactive redraws, a band's own rows) and
// far-apart widgets (or a band's two halo rectangles) qualify; overlapping footprints keep the reference order,
// one launch per level and rectangle.
static int pbrRegeneratePyramid(b2pbr* c, b2mip* m, const int* q, const int (*need)[2], const b2_box2i* rects, int n) {
    const int nl = (int)m->levels.size() - 1;
    if (nl <= 0 || n <= 0) return B2_OK;
    std::vector<b2_box2i> chain((size_t)n * nl);      // chain[r * nl + l - 1] = update rectangle of level l for rectangle r
    for (int r = 0; r < n; ++r) {
        b2_box2i area = rects[r];
        for (int l = 1; l <= nl; ++l) {
            const DevLevel& Pv = m->levels[l - 1];
            area = impactOnNextLevel(q[l - 1], area, Pv.w, Pv.h);
            chain[(size_t)r * nl + l - 1] = area;
        }
    }
    auto clipped = [&](int r, int l) {   // a band only regenerates the rows it reads (its neighbours own the others)
        b2_box2i b = chain[(size_t)r * nl + l - 1];
        if (need) {
            if (b.min_y < need[l][0]) b.min_y = need[l][0];
            if (b.max_y > need[l][1]) b.max_y = need[l][1];
            if (b.max_y <= b.min_y) b = b2_box2i{0, 0, 0, 0};
        }
        return b;
    };
    bool independent = n <= 16;           // the pair test is quadratic; long lists keep the reference order
    for (int i = 0; i < n && independent; ++i)
        for (int j = 0; j < n && independent; ++j) {
            if (i == j) continue;
            for (int l = 1; l <= nl; ++l) {
                const b2_box2i a = chain[(size_t)i * nl + l - 1], b = chain[(size_t)j * nl + l - 1];
                if (boxW(a) <= 0 || boxH(a) <= 0 || boxW(b) <= 0 || boxH(b) <= 0) continue;
                const int reach = 4;      // [1 3 3 1] reads 2x-1 .. 2x+2: within 3 texels of the rectangle one level down
                if (a.min_x - reach < b.max_x && b.min_x < a.max_x + reach && a.min_y - reach < b.max_y && b.min_y < a.max_y + reach) { independent = false; break; }
            }
        }
    if (independent) {
        std::vector<b2_box2i> lv(nl);
        for (int r = 0; r < n; ++r) {
            for (int l = 1; l <= nl; ++l) lv[l - 1] = clipped(r, l);
            if (int rc = mipGenerateLevels(m, 1, nl, q, lv.data())) return rc;
        }
        return B2_OK;
    }
    for (int l = 1; l <= nl; ++l)
        for (int r = 0; r < n; ++r)
            if (int rc = mipGenerateLevel(m, l, q[l - 1], clipped(r, l))) return rc;
    return B2_OK;
}

int b2pbr_regenerate_mipmaps(b2pbr* c, const b2_box2i* rects, int n) {
    if (!c || !c->diffuse) return fail(B2_ERR_STATE, "regenerateMipmaps before resize");
    if (n <= 0) return B2_OK;  // graphics.d:1358-1360
    if (!rects) return fail(B2_ERR_INVALID, "regenerateMipmaps: null rectangle list");
    uint64_t before = c->diffuse->launches + c->depth->launches;
    CU(cudaSetDevice(c->device));
    // The two pyramids are independent (the reference builds them one after the other on one thread,
    // graphics.d:1369-1370): the depth chain runs on a second stream beside the diffuse chain and joins before
    // anything else is enqueued on the compositor's stream.
    const bool fork = !c->profile;
    cudaStream_t aux = c->sAux;
    if (fork && c->useAuxHi) {
        if (!c->sAuxHi) {
            int least = 0, greatest = 0;
            CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
            CU(cudaStreamCreateWithPriority(&c->sAuxHi, cudaStreamNonBlocking, greatest));
        }
        aux = c->sAuxHi;
    }
    if (fork) {
        CU(cudaEventRecord(c->evFork, c->stream));
        CU(cudaStreamWaitEvent(aux, c->evFork, 0));
        c->depth->stream = aux;
    }
    int qD[kMaxDiffuseLevels], qZ[kMaxDepthLevels];
    for (int level = 1; level < kMaxDiffuseLevels; ++level) qD[level - 1] = level == 1 ? B2_QUALITY_BOX_ALPHACOV_PREMUL : B2_QUALITY_CUBIC;  // graphics.d:1380-1384
    for (int level = 1; level < kMaxDepthLevels; ++level) qZ[level - 1] = level >= 3 ? B2_QUALITY_CUBIC : B2_QUALITY_BOX;                   // graphics.d:1414
    traceBegin(c, "diffuse mipmap");
    int err = pbrRegeneratePyramid(c, c->diffuse, qD, c->banded ? c->needD : nullptr, rects, n);
    traceEnd(c);
    // depth: replicateBordersTouching(area) (graphics.d:1406-1408) is clamp-to-edge addressing on the device
    traceBegin(c, "depth mipmap");
    if (!err) err = pbrRegeneratePyramid(c, c->depth, qZ, c->banded ? c->needZ : nullptr, rects, n);
    traceEnd(c);
    if (fork) {
        c->depth->stream = c->stream;
        CU(cudaEventRecord(c->evJoin, aux));
        CU(cudaStreamWaitEvent(c->stream, c->evJoin, 0));
    }
    if (err) return err;
    c->launches += c->diffuse->launches + c->depth->launches - before;
    return B2_OK;
}

int b2pbr_regenerate_mipmaps_on(b2pbr* c, const b2_box2i* rects, int n, void* cuda_stream) {
    if (!c || !c->diffuse) return fail(B2_ERR_STATE, "regenerateMipmaps before resize");
    // same work, enqueued on the caller's stream (no synchronisation): the caller orders it against the compositor's
    // own stream with events (multi-GPU: the halo rows' mips run beside the lighting of the tiles that do not read them)
    cudaStream_t keep = c->stream, keepD = c->diffuse->stream, keepZ = c->depth->stream;
    c->stream = c->diffuse->stream = c->depth->stream = (cudaStream_t)cuda_stream;
    const int rc = b2pbr_regenerate_mipmaps(c, rects, n);
    c->stream = keep; c->diffuse->stream = keepD; c->depth->stream = keepZ;
    return rc;
}

int b2pbr_composite_tile(b2pbr* c, const b2_box2i* areas, int n, b2mip* diffuse, b2mip* material, b2mip* depth) {
    if (!c || (n > 0 && !areas)) return fail(B2_ERR_INVALID, "compositeTile: bad arguments");
    return pbrComposite(c, areas, n, diffuse, material, depth);
}

// One frame = regenerateMipmaps + compositeTile (GUIGraphics.doDraw stages C and D, graphics.d:812-821) as one call.
// (Measured and NOT done: running the mip chain of the lower part of the frame beside the lighting of the upper part on a
// second, high-priority stream.  The lighting kernel fills every SM's register file (5 CTAs x 128 threads x 96 registers),
// so the mip CTAs cannot co-reside with it and only displace lighting CTAs: 0.209 - 0.222 ms per 4K frame against
// 0.196 ms for the plain sequence.)
int b2pbr_redraw(b2pbr* c, const b2_box2i* update_rects, int n_update, const b2_box2i* areas, int n_areas) {
    if (!c || !c->diffuse || (n_update > 0 && !update_rects) || (n_areas > 0 && !areas)) return fail(B2_ERR_INVALID, "redraw: bad arguments");
    if (int rc = b2pbr_regenerate_mipmaps(c, update_rects, n_update)) return rc;
    return pbrComposite(c, areas, n_areas, nullptr, nullptr, nullptr);
}

int b2pbr_composite_raw(b2pbr* c, const b2_box2i* areas, int n) {
    if (!c || !c->outAlloc || (n > 0 && !areas)) return fail(B2_ERR_STATE, "composite_raw: bad state");
    if (n <= 0) return B2_OK;
    for (int k = 0; k < n; ++k) {   // same validation as compositeTile
        const b2_box2i& a = areas[k];
        if (!boxSorted(a) || a.min_x < 0 || a.min_y < 0 || a.max_x > c->W || a.max_y > c->H)
            return fail(B2_ERR_INVALID, "composite_raw: area outside the frame");
        if (!boxEmpty(a) && (a.min_y < c->out.y0 || a.max_y > c->out.y0 + c->out.rows))
            return fail(B2_ERR_INVALID, "composite_raw: area outside this band");
    }
    CU(cudaSetDevice(c->device));
    int rc = pbrUploadBoxes(c, areas, n, 0);
    if (rc) return rc;
    launch_copy_diffuse(c->out, c->diffuse->levels[0], c->dBoxes, n, n, c->stream);
    c->launches++;
    CU(cudaGetLastError());
    return B2_OK;
}

static int pbrDownloadEnqueue(b2pbr* c, void* host, size_t host_pitch, const b2_box2i* rects, int n) {
    if (!c || !c->outAlloc || !host) return fail(B2_ERR_STATE, "download: bad state");
    CU(cudaSetDevice(c->device));
    for (int k = 0; k < n; ++k) {
        const b2_box2i& r = rects[k];
        if (!boxSorted(r) || r.min_x < 0 || r.min_y < c->out.y0 || r.max_x > c->W || r.max_y > c->out.y0 + c->out.rows)
            return fail(B2_ERR_INVALID, "download: rectangle outside the stored frame");
    }
    std::vector<b2_box2i> merged;
    mergeRects(rects, n, merged);
    size_t stagedAt = 0;
    for (const b2_box2i& r : merged) {
        uint8_t* hp = (uint8_t*)host + (size_t)r.min_y * host_pitch + (size_t)r.min_x * 4;
        const uint8_t* dp = c->out.base + (size_t)(r.min_y - c->out.y0) * c->out.pitch + (size_t)r.min_x * 4;
        const size_t rowBytes = (size_t)boxW(r) * 4, bytes = rowBytes * (size_t)boxH(r);
        if (bytes <= kSmCopyMaxBytes)
            if (void* mapped = mappedHostPointer(hp)) {                            // small + page-locked: the SMs write the host image in place
                launch_copy2d(mapped, host_pitch, dp, (size_t)c->out.pitch, rowBytes, boxH(r), c->stream);
                c->launches++;
                CU(cudaGetLastError());
                continue;
            }
        if (rowBytes == host_pitch && host_pitch == (size_t)c->out.pitch)          // gapless on both sides
            CU(cudaMemcpyAsync(hp, dp, bytes, cudaMemcpyDeviceToHost, c->stream));
        else if (rowBytes == host_pitch && boxH(r) > 1 && stagedAt + bytes <= ((size_t)c->W * 4) * (size_t)c->H) {
            // whole rows of a gapless host image: gathered on the device, one linear copy over PCIe (see stagingReserve)
            int rc = stagingReserve(&c->dlStaging, &c->dlStagingBytes, ((size_t)c->W * 4) * (size_t)c->H, c->stream);
            if (rc) return rc;
            CU(cudaMemcpy2DAsync(c->dlStaging + stagedAt, rowBytes, dp, (size_t)c->out.pitch, rowBytes, (size_t)boxH(r), cudaMemcpyDeviceToDevice, c->stream));
            CU(cudaMemcpyAsync(hp, c->dlStaging + stagedAt, bytes, cudaMemcpyDeviceToHost, c->stream));
            stagedAt += alignUp(bytes, 256);
        } else
            CU(cudaMemcpy2DAsync(hp, host_pitch, dp, (size_t)c->out.pitch, rowBytes, (size_t)boxH(r), cudaMemcpyDeviceToHost, c->stream));
    }
    return B2_OK;
}
int b2pbr_download(b2pbr* c, void* host, size_t host_pitch, const b2_box2i* rects, int n) {
    int rc = pbrDownloadEnqueue(c, host, host_pitch, rects, n);
    if (rc) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

int b2pbr_download_swizzled(b2pbr* c, int pf, void* host, size_t host_pitch, const b2_box2i* rects, int n) {
    if (!c || !c->outAlloc || !host || pf < 0 || pf > 2) return fail(B2_ERR_STATE, "download_swizzled: bad arguments");
    CU(cudaSetDevice(c->device));
    for (int k = 0; k < n; ++k) {
        const b2_box2i& r = rects[k];
        if (!boxSorted(r) || r.min_x < 0 || r.min_y < c->out.y0 || r.max_x > c->W || r.max_y > c->out.y0 + c->out.rows)
            return fail(B2_ERR_INVALID, "download_swizzled: rectangle outside the stored frame");
    }
    std::vector<b2_box2i> merged;
    mergeRects(rects, n, merged);
    for (const b2_box2i& r : merged) {
        launch_swizzle(c->swz, c->out, toBox(r), pf, c->stream);
        c->launches++;
        CU(cudaMemcpy2DAsync((uint8_t*)host + (size_t)r.min_y * host_pitch + (size_t)r.min_x * 4, host_pitch,
                             c->swz.base + (size_t)(r.min_y - c->swz.y0) * c->swz.pitch + (size_t)r.min_x * 4, (size_t)c->swz.pitch,
                             (size_t)boxW(r) * 4, (size_t)boxH(r), cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

// Full-frame host call as a 3-stage pipeline over row chunks: while chunk k+1 of the three level-0 maps crosses
// PCIe, the mip rows that became computable are generated and the tiles whose whole sampling footprint (bloom:
// 94 rows) has arrived are composited, and finished tile rows travel back on a third stream.
// Same results as the sequential path: the work is only re-ordered (every mip row is generated once, from
// complete sources; every area is composited once).
static int pbrHostPipelined(b2pbr* c, b2_imageref wfb, const b2_box2i* areas, int n_areas,
                            b2_imageref dI, b2_imageref mI, b2_imageref zI, bool wait) {
    const int W = c->W, H = c->H;
    CU(cudaSetDevice(c->device));
    if (!c->sIn) {
        CU(cudaStreamCreateWithFlags(&c->sIn, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&c->sOut, cudaStreamNonBlocking));
    }
    // Chunk schedule.  Few chunks: with the finished rows travelling back at the same time every additional chunk
    // costs PCIe efficiency (measured on the B200 box, tools/probes/pcie_probe.py, 83 + 33 MB of a 4K frame: 1 chunk 1.63 ms,
    // 4 chunks 1.72 ms, 6 chunks 1.79 ms).  A small last chunk: whatever is composited after the final bytes have
    // arrived is the exposed tail of the call.  The diffuse map runs `ahead` rows in front of the other two, because
    // it is the bloom (94 rows of diffuse below a tile) that makes tiles wait; depth needs 28 rows.
    const int ahead = 128;
    std::vector<int> bounds{0};
    {
        // Sizes from a fluid model of the measured rates (uploads 1.27 rows/us, kernels 5.3 rows/us plus ~100 us of
        // dependent launches per stage): a stage must finish before the next chunk has arrived,
        // s[k+1] >= 127 + 0.24 s[k], and the last chunk — whose stage and copy-back are exposed — is as small as that allows.
        auto r64 = [](double v) { return (int)((v + 32.0) / 64.0) * 64; };
        const int cuts[3] = {r64(0.44 * H), r64(0.77 * H), r64(0.91 * H)};
        for (int cut : cuts)
            if (cut >= bounds.back() + 128 && cut < H) bounds.push_back(cut);
        bounds.push_back(H);
    }
    const int nChunks = (int)bounds.size() - 1;
    while ((int)c->pipeEvents.size() < 2 * nChunks + 1) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->pipeEvents.push_back(e);
    }
    static const bool dbg = getenv("B2_PIPE_DEBUG") != nullptr;
    cudaEvent_t d0 = nullptr, d1 = nullptr, d2 = nullptr, d3 = nullptr;
    std::vector<cudaEvent_t> dIn, dK0, dK1, dK2;   // per chunk: upload done, stage start, mips done, tiles done
    if (dbg) { cudaEventCreate(&d0); cudaEventCreate(&d1); cudaEventCreate(&d2); cudaEventCreate(&d3); }
    auto mark = [&](std::vector<cudaEvent_t>& v, cudaStream_t st) { if (dbg) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); v.push_back(e); } };
    // uploads must not overtake kernels of a previous call that still read the maps
    CU(cudaEventRecord(c->pipeEvents[2 * nChunks], c->stream));
    CU(cudaStreamWaitEvent(c->sIn, c->pipeEvents[2 * nChunks], 0));
    if (dbg) cudaEventRecord(d0, c->sIn);
    // areas in the order their rows complete; the device work list is uploaded once for the whole frame (and kept
    // across frames while the caller passes the same areas), every stage launches a slice of it
    std::vector<int> order(n_areas);
    for (int k = 0; k < n_areas; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return areas[a].max_y < areas[b].max_y; });
    std::vector<b2_box2i> sorted(n_areas);
    for (int k = 0; k < n_areas; ++k) sorted[k] = areas[order[k]];
    // the tiles composited after the last bytes have arrived are the exposed tail of the call: short strips spread them
    // over the whole GPU (results do not depend on the strip cut)
    auto diffuseUpTo = [&](int k) { int v = bounds[k] + (k > 0 ? ahead : 0); return (k == nChunks || v > H) ? H : v; };   // rows of diffuse present after chunk k-1
    struct TailGuard { b2pbr* c; ~TailGuard() { c->tailRow = c->tailRow2 = 1 << 30; } } tailGuard{c};
    // rows [0, cEnd[k]) can be composited once chunk k has arrived: every level-0 row their passes (and the mip rows
    // they sample) read is present — diffuse up to diffuseUpTo(k + 1), depth and material up to bounds[k + 1]
    std::vector<int> cEnd(nChunks, 0);
    {
        int nd[kMaxDiffuseLevels][2], nz[kMaxDepthLevels][2], prev = 0;
        for (int k = 0; k < nChunks; ++k) {
            const int y1 = bounds[k + 1], yD = diffuseUpTo(k + 1);
            int cNew = prev;
            if (y1 == H) cNew = H;
            else
                for (int t = prev + 64; t <= y1; t += 64) {
                    computeBandNeeds(W, H, prev, t, nd, nz);
                    if (nd[0][1] > yD || nz[0][1] > y1) break;
                    cNew = t;
                }
            cEnd[k] = prev = cNew;
        }
    }
    // the last two stages run next to (and after) the end of the transfer: shorter strips, every mip level fused
    c->tailRow = nChunks >= 3 ? cEnd[nChunks - 3] : H - 256;
    c->tailRow2 = nChunks >= 2 ? cEnd[nChunks - 2] : H - 128;
    const int R0 = stripRowsFor(areas, n_areas);
    auto workItems = [&](int k, const b2_box2i& a, bool fast) {   // k = index in the sorted list that is uploaded
        if (!fast) return 1;
        const int R = stripRowsOfAreaAt(c, k, n_areas, a, R0) * 2, bw = 64;   // CTA boxes of 64 x 2R (k_pbr_tex)
        return ((boxH(a) + R - 1) / R) * ((boxW(a) + bw - 1) / bw);
    };
    // stage 1: every upload is enqueued up front, so the copy engine never waits for this thread (the host-side cost
    // of enqueueing a stage's kernels is comparable to the transfer time of the small last chunks)
    for (int k = 0; k < nChunks; ++k) {
        struct { b2mip* m; b2_imageref im; int y0, y1; } ups[3] = {{c->diffuse, dI, diffuseUpTo(k), diffuseUpTo(k + 1)},
                                                                     {c->material, mI, bounds[k], bounds[k + 1]},
                                                                     {c->depth, zI, bounds[k], bounds[k + 1]}};
        for (auto& u : ups) {
            int rc = mipUploadRows(u.m, 0, u.y0, u.y1, u.im.pixels, u.im.pitch, c->sIn, false);   // large frames: the copy engine, next to the kernels
            if (rc) return rc;
        }
        CU(cudaEventRecord(c->pipeEvents[2 * k], c->sIn));
        mark(dIn, c->sIn);
    }
    if (dbg) cudaEventRecord(d1, c->sIn);
    size_t pos = 0, dlAt = 0;
    int workPos = 0;
    int cPrev = 0;
    if (wfb.pitch != (size_t)c->out.pitch) {   // the frame goes back through the linear device buffer (see stagingReserve)
        int rc = stagingReserve(&c->dlStaging, &c->dlStagingBytes, ((size_t)W * 4 + 256) * (size_t)H, c->stream);
        if (rc) return rc;
    }
    int genD[kMaxDiffuseLevels] = {0}, genZ[kMaxDepthLevels] = {0};
    const int nD = (int)c->diffuse->levels.size(), nZ = (int)c->depth->levels.size();
    std::vector<b2_box2i> merged;
    const uint64_t mipLaunches0 = c->diffuse->launches + c->depth->launches;
    for (int k = 0; k < nChunks; ++k) {
        CU(cudaStreamWaitEvent(c->stream, c->pipeEvents[2 * k], 0));
        const int cNew = cEnd[k];
        int nd[kMaxDiffuseLevels][2], nz[kMaxDepthLevels][2];
        if (cNew == cPrev) continue;
        computeBandNeeds(W, H, cPrev, cNew, nd, nz);
        mark(dK0, c->stream);
        {   // the mip rows that became computable, as one chain per pyramid (small levels share launches)
            int q[kMaxDiffuseLevels];
            b2_box2i rects[kMaxDiffuseLevels];
            for (int l = 1; l < nD; ++l) {
                q[l - 1] = l == 1 ? B2_QUALITY_BOX_ALPHACOV_PREMUL : B2_QUALITY_CUBIC;   // graphics.d:1380-1384
                const int hi = nd[l][1] > genD[l] ? nd[l][1] : genD[l];
                rects[l - 1] = b2_box2i{0, genD[l], c->diffuse->levels[l].w, hi};
                genD[l] = hi;
            }
            // the two pyramids are independent: the depth chain runs beside the diffuse chain (as in regenerateMipmaps)
            CU(cudaEventRecord(c->evFork, c->stream));
            CU(cudaStreamWaitEvent(c->sAux, c->evFork, 0));
            const bool late = k >= nChunks - 2;
            int rc = mipGenerateLevels(c->diffuse, 1, nD - 1, q, rects, late);
            if (rc) return rc;
            for (int l = 1; l < nZ; ++l) {
                q[l - 1] = l >= 3 ? B2_QUALITY_CUBIC : B2_QUALITY_BOX;                    // graphics.d:1414
                const int hi = nz[l][1] > genZ[l] ? nz[l][1] : genZ[l];
                rects[l - 1] = b2_box2i{0, genZ[l], c->depth->levels[l].w, hi};
                genZ[l] = hi;
            }
            c->depth->stream = c->sAux;
            rc = mipGenerateLevels(c->depth, 1, nZ - 1, q, rects, late);
            c->depth->stream = c->stream;
            if (rc) return rc;
            CU(cudaEventRecord(c->evJoin, c->sAux));
            CU(cudaStreamWaitEvent(c->stream, c->evJoin, 0));
        }
        mark(dK1, c->stream);
        const size_t first = pos;
        while (pos < sorted.size() && sorted[pos].max_y <= cNew) ++pos;
        if (pos > first) {
            bool fast = false;
            // first call uploads (or re-uses) the list of the whole frame and tells which kernel will run
            int rc = pbrComposite(c, sorted.data(), n_areas, nullptr, nullptr, nullptr, 0, 0, &fast);
            if (rc) return rc;
            int cnt = 0;
            for (size_t q = first; q < pos; ++q) cnt += workItems((int)q, sorted[q], fast);
            rc = pbrComposite(c, sorted.data(), n_areas, nullptr, nullptr, nullptr, workPos, cnt, nullptr);
            if (rc) return rc;
            workPos += cnt;
            mark(dK2, c->stream);
            CU(cudaEventRecord(c->pipeEvents[2 * k + 1], c->stream));
            CU(cudaStreamWaitEvent(c->sOut, c->pipeEvents[2 * k + 1], 0));
            mergeRects(sorted.data() + first, (int)(pos - first), merged);
            for (const b2_box2i& r : merged) {
                uint8_t* hp = (uint8_t*)wfb.pixels + (size_t)r.min_y * wfb.pitch + (size_t)r.min_x * 4;
                const uint8_t* dp = c->out.base + (size_t)(r.min_y - c->out.y0) * c->out.pitch + (size_t)r.min_x * 4;
                const size_t rowBytes = (size_t)boxW(r) * 4, bytes = rowBytes * (size_t)boxH(r);
                if (rowBytes == wfb.pitch && wfb.pitch == (size_t)c->out.pitch)                      // whole rows, gapless on both sides
                    CU(cudaMemcpyAsync(hp, dp, bytes, cudaMemcpyDeviceToHost, c->sOut));
                else if (rowBytes == wfb.pitch && boxH(r) > 1 && c->dlStaging && dlAt + bytes <= c->dlStagingBytes) {
                    CU(cudaMemcpy2DAsync(c->dlStaging + dlAt, rowBytes, dp, (size_t)c->out.pitch, rowBytes, (size_t)boxH(r), cudaMemcpyDeviceToDevice, c->sOut));
                    CU(cudaMemcpyAsync(hp, c->dlStaging + dlAt, bytes, cudaMemcpyDeviceToHost, c->sOut));
                    dlAt += alignUp(bytes, 256);
                } else
                    CU(cudaMemcpy2DAsync(hp, wfb.pitch, dp, (size_t)c->out.pitch, rowBytes, (size_t)boxH(r), cudaMemcpyDeviceToHost, c->sOut));
            }
        }
        cPrev = cNew;
    }
    c->launches += c->diffuse->launches + c->depth->launches - mipLaunches0;
    if (dbg) { cudaEventRecord(d2, c->stream); cudaEventRecord(d3, c->sOut); }
    // the copy-back stream joins the compositor's stream, so one wait on c->stream (b2pbr_composite_tile_host_end, or
    // anything the caller enqueues there) covers the whole call
    CU(cudaEventRecord(c->pipeEvents[2 * nChunks], c->sOut));
    CU(cudaStreamWaitEvent(c->stream, c->pipeEvents[2 * nChunks], 0));
    if (!wait && !dbg) return B2_OK;
    CU(cudaStreamSynchronize(c->stream));
    if (dbg) {
        float a = 0, b = 0, e = 0;
        cudaEventElapsedTime(&a, d0, d1); cudaEventElapsedTime(&b, d0, d2); cudaEventElapsedTime(&e, d0, d3);
        fprintf(stderr, "[b2pbr pipe] chunks=%d  H2D done %.3f ms, kernels done %.3f ms, D2H done %.3f ms\n", nChunks, a, b, e);
        auto dump = [&](const char* name, std::vector<cudaEvent_t>& v) {
            fprintf(stderr, "[b2pbr pipe]   %s:", name);
            for (cudaEvent_t ev : v) { float t = 0; cudaEventElapsedTime(&t, d0, ev); fprintf(stderr, " %.3f", t); cudaEventDestroy(ev); }
            fprintf(stderr, "\n");
        };
        dump("upload done", dIn); dump("stage start", dK0); dump("mips done", dK1); dump("tiles done", dK2);
        cudaEventDestroy(d0); cudaEventDestroy(d1); cudaEventDestroy(d2); cudaEventDestroy(d3);
    }
    return B2_OK;
}

static int pbrCompositeTileHost(b2pbr* c, b2_imageref wfb, const b2_box2i* areas, int n_areas, const b2_box2i* update_rects,
                                int n_update, b2_imageref diffuse_l0, b2_imageref material_l0, b2_imageref depth_l0, bool wait) {
    if (!c || !c->diffuse) return fail(B2_ERR_STATE, "compositeTile(host) before resizeBuffers");
    if (diffuse_l0.w != c->W || diffuse_l0.h != c->H || material_l0.w != c->W || material_l0.h != c->H ||
        depth_l0.w != c->W || depth_l0.h != c->H || wfb.w != c->W || wfb.h != c->H)
        return fail(B2_ERR_INVALID, "compositeTile(host): image size differs from resizeBuffers size");
    // The reference hands compositeTile only the tile list (compositor.d:63-68): with update_rects == NULL the areas
    // are what changed.  Either way the list is coalesced first (a full-frame 64x64 tile list becomes ONE rectangle):
    // regenerating the union of adjacent rectangles recomputes exactly the texels the separate rectangles would
    // (impactOnNextLevel of a union of adjacent boxes is the union of their impacts, and every regenerated texel is a
    // pure function of the level below), so the uploads, the mip chain and the whole-frame test see few large boxes.
    std::vector<b2_box2i> coalesced;
    if (!update_rects) mergeRects(areas, n_areas, coalesced);
    else {
        if (n_update < 0) return fail(B2_ERR_INVALID, "compositeTile(host): negative update rectangle count");
        mergeRects(update_rects, n_update, coalesced);
    }
    update_rects = coalesced.data(); n_update = (int)coalesced.size();
    int rc;
    // whole frame dirty (first frame, resize, animated UI): overlap PCIe with the kernels
    if (!c->banded && !c->profile && n_update == 1 && update_rects[0].min_x == 0 && update_rects[0].min_y == 0 &&
        update_rects[0].max_x == c->W && update_rects[0].max_y == c->H && c->H >= 512)
        return pbrHostPipelined(c, wfb, areas, n_areas, diffuse_l0, material_l0, depth_l0, wait);
    // Small frames and dirty rectangles: the fixed cost of every transfer is what matters (a 620 x 330 call is 0.12 ms of
    // which the lighting is 14 us).  The three maps cross PCIe on three streams at once: diffuse on the compositor's
    // stream, depth on the stream its pyramid is regenerated on, material — which the mip chain does not read — on a third
    // one, beside the mips; the lighting waits for all of them.
    const bool split = !c->profile;
    if (split) {
        CU(cudaSetDevice(c->device));
        if (!c->sIn) {
            CU(cudaStreamCreateWithFlags(&c->sIn, cudaStreamNonBlocking));
            CU(cudaStreamCreateWithFlags(&c->sOut, cudaStreamNonBlocking));
        }
        if (!c->evMatFork) {
            CU(cudaEventCreateWithFlags(&c->evMatFork, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&c->evMatDone, cudaEventDisableTiming));
        }
        CU(cudaEventRecord(c->evMatFork, c->stream));          // not before the previous call's kernels have read the map
        CU(cudaStreamWaitEvent(c->sIn, c->evMatFork, 0));
    }
    traceBegin(c, "upload dirty rects");
    if ((rc = b2mip_upload(c->diffuse, 0, diffuse_l0.pixels, diffuse_l0.pitch, update_rects, n_update))) return rc;
    {   // depth: on the stream its pyramid is regenerated on (b2pbr_regenerate_mipmaps forks the depth chain to sAux)
        cudaStream_t keep = c->depth->stream;
        if (split) { CU(cudaStreamWaitEvent(c->sAux, c->evMatFork, 0)); c->depth->stream = c->sAux; }
        rc = b2mip_upload(c->depth, 0, depth_l0.pixels, depth_l0.pitch, update_rects, n_update);
        c->depth->stream = keep;
        if (rc) return rc;
    }
    {
        cudaStream_t keep = c->material->stream;
        if (split) c->material->stream = c->sIn;
        rc = b2mip_upload(c->material, 0, material_l0.pixels, material_l0.pitch, update_rects, n_update);
        c->material->stream = keep;
        if (rc) return rc;
        if (split) CU(cudaEventRecord(c->evMatDone, c->sIn));
    }
    traceEnd(c);
    if ((rc = b2pbr_regenerate_mipmaps(c, update_rects, n_update))) return rc;
    if (split) CU(cudaStreamWaitEvent(c->stream, c->evMatDone, 0));
    if ((rc = pbrComposite(c, areas, n_areas, nullptr, nullptr, nullptr))) return rc;
    traceBegin(c, "download composited");
    rc = pbrDownloadEnqueue(c, wfb.pixels, wfb.pitch, areas, n_areas);
    traceEnd(c);
    if (rc || !wait) return rc;
    CU(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

int b2pbr_composite_tile_host(b2pbr* c, b2_imageref wfb, const b2_box2i* areas, int n_areas, const b2_box2i* update_rects,
                              int n_update, b2_imageref diffuse_l0, b2_imageref material_l0, b2_imageref depth_l0) {
    return pbrCompositeTileHost(c, wfb, areas, n_areas, update_rects, n_update, diffuse_l0, material_l0, depth_l0, true);
}
int b2pbr_composite_tile_host_begin(b2pbr* c, b2_imageref wfb, const b2_box2i* areas, int n_areas, const b2_box2i* update_rects,
                                    int n_update, b2_imageref diffuse_l0, b2_imageref material_l0, b2_imageref depth_l0) {
    return pbrCompositeTileHost(c, wfb, areas, n_areas, update_rects, n_update, diffuse_l0, material_l0, depth_l0, false);
}
int b2pbr_composite_tile_host_end(b2pbr* c) {
    if (!c) return fail(B2_ERR_INVALID, "null compositor");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

int b2_band_need_rows(int W, int H, int y0, int y1, int map, int level, int* need_lo, int* need_hi) {
    if (W <= 0 || H <= 0 || y0 < 0 || y1 > H || y0 > y1 || level < 0) return fail(B2_ERR_INVALID, "band_need_rows: bad arguments");
    int nd[kMaxDiffuseLevels][2], nz[kMaxDepthLevels][2];
    computeBandNeeds(W, H, y0, y1, nd, nz);
    int lo, hi;
    if (map == B2_MAP_DIFFUSE && level < kMaxDiffuseLevels) { lo = nd[level][0]; hi = nd[level][1]; }
    else if (map == B2_MAP_DEPTH && level < kMaxDepthLevels) { lo = nz[level][0]; hi = nz[level][1]; }
    else if (map == B2_MAP_MATERIAL && level == 0) { lo = y0; hi = y1; }
    else return fail(B2_ERR_INVALID, "band_need_rows: bad map / level");
    if (need_lo) *need_lo = lo;
    if (need_hi) *need_hi = hi;
    return B2_OK;
}

int b2pbr_band_rows(const b2pbr* c, int map, int level, int* stored_lo, int* stored_hi, int* need_lo, int* need_hi) {
    if (!c || !c->diffuse) return fail(B2_ERR_STATE, "band_rows before resize");
    const b2mip* m = map == B2_MAP_DIFFUSE ? c->diffuse : (map == B2_MAP_DEPTH ? c->depth : (map == B2_MAP_MATERIAL ? c->material : nullptr));
    if (!m || level < 0 || level >= (int)m->levels.size()) return fail(B2_ERR_INVALID, "band_rows: bad map / level");
    const DevLevel& L = m->levels[level];
    if (stored_lo) *stored_lo = L.y0;
    if (stored_hi) *stored_hi = L.y0 + L.rows;
    int lo = L.y0, hi = L.y0 + L.rows;
    if (c->banded) {
        if (map == B2_MAP_DIFFUSE) { lo = c->needD[level][0]; hi = c->needD[level][1]; }
        else if (map == B2_MAP_DEPTH) { lo = c->needZ[level][0]; hi = c->needZ[level][1]; }
        else { lo = c->bandY0; hi = c->bandY1; }
    }
    if (need_lo) *need_lo = lo;
    if (need_hi) *need_hi = hi;
    return B2_OK;
}

// ---- CUDA graphs: record the device work of one frame (regenerateMipmaps + compositeTile on device-resident maps)
// once and replay it; the launch-bound chain of small mip kernels then costs one graph launch per frame.
int b2pbr_graph_begin(b2pbr* c) {
    if (!c || c->capturing) return fail(B2_ERR_STATE, "graph_begin: bad state");
    CU(cudaSetDevice(c->device));
    CU(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
    c->capturing = true;
    c->launchesAtCapture = c->launches;
    return B2_OK;
}
int b2pbr_graph_end(b2pbr* c, int* graph_id) {
    if (!c || !c->capturing || !graph_id) return fail(B2_ERR_STATE, "graph_end: not recording");
    c->capturing = false;
    cudaGraph_t g = nullptr;
    CU(cudaStreamEndCapture(c->stream, &g));
    cudaGraphExec_t e = nullptr;
    cudaError_t err = cudaGraphInstantiate(&e, g, 0);
    cudaGraphDestroy(g);
    if (err != cudaSuccess) return fail(B2_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(err));
    c->graphs.push_back(e);
    c->graphLaunches.push_back(c->launches - c->launchesAtCapture);   // recorded, not executed
    c->launches = c->launchesAtCapture;
    *graph_id = (int)c->graphs.size() - 1;
    return B2_OK;
}
int b2pbr_graph_launch(b2pbr* c, int graph_id) {
    if (!c || graph_id < 0 || graph_id >= (int)c->graphs.size()) return fail(B2_ERR_INVALID, "graph_launch: bad id");
    if (!c->graphs[graph_id]) return fail(B2_ERR_STATE, "graph_launch: the recording was invalidated by resizeBuffers / setSkybox / set_params / set_mode; record it again");
    CU(cudaGraphLaunch(c->graphs[graph_id], c->stream));
    c->launches += c->graphLaunches[graph_id];
    return B2_OK;
}

int b2pbr_draw_layer(b2pbr* c, b2layer* l, int x, int y, const b2_box2i* rects, int n, int mode) {
    if (!c || !c->diffuse || !l || (n > 0 && !rects) || (mode != B2_LAYER_BLIT && mode != B2_LAYER_BLEND)) return fail(B2_ERR_INVALID, "draw_layer: bad arguments");
    if (mode == B2_LAYER_BLEND && !l->opacity) return fail(B2_ERR_INVALID, "draw_layer: this layer has no opacity masks");
    if (l->device != c->device) return fail(B2_ERR_INVALID, "draw_layer: layer lives on another device");
    if (x < 0 || y < 0 || x + l->w > c->W || y + l->h > c->H) return fail(B2_ERR_INVALID, "draw_layer: widget outside the UI");
    CU(cudaSetDevice(c->device));
    const DevLevel dst3[3] = {c->diffuse->levels[0], c->depth->levels[0], c->material->levels[0]};
    for (int k = 0; k < n; ++k) {
        const b2_box2i& r = rects[k];
        if (!boxSorted(r) || r.min_x < 0 || r.min_y < 0 || r.max_x > l->w || r.max_y > l->h) return fail(B2_ERR_INVALID, "draw_layer: dirty rectangle outside the widget");
        if (boxEmpty(r)) continue;
        if (r.min_y + y < dst3[0].y0 || r.max_y + y > dst3[0].y0 + dst3[0].rows) return fail(B2_ERR_INVALID, "draw_layer: rectangle outside this band");
        launch_layer(mode == B2_LAYER_BLEND, dst3, l->plane, l->plane + 3, toBox(r), x, y, c->stream);
        c->launches++;
    }
    CU(cudaGetLastError());
    return B2_OK;
}

int b2pbr_set_color_correction(b2pbr* c, const uint8_t* table768) {
    if (!c) return fail(B2_ERR_INVALID, "null compositor");
    CU(cudaSetDevice(c->device));
    if (!table768) { c->lutOn = false; return B2_OK; }
    if (!c->dLut) CU(cudaMalloc((void**)&c->dLut, 768));
    CU(cudaMemcpyAsync(c->dLut, table768, 768, cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    c->lutOn = true;
    return B2_OK;
}

void b2_lift_gamma_gain_contrast_table(const float v[12], uint8_t out[768]) {
    // UIColorCorrection.setLiftGammaGainContrastRGB (colorcorrection.d:69-116); v = {lift, gamma, gain, contrast} x {r, g, b}
    for (int ch = 0; ch < 3; ++ch) {
        const float lift = v[4 * ch], gamma = v[4 * ch + 1], gain = v[4 * ch + 2], contrast = v[4 * ch + 3];
        for (int b = 0; b < 256; ++b) {
            float inp = b / 255.0f;
            float o = gain * (inp + lift * (1 - inp));
            float a = o < 0 ? 0.f : (o > 1 ? 1.f : o);                 // safePow clamps its base
            o = (float)std::pow((double)a, (double)(1.0f / gamma));
            if (o < 0) o = 0;
            if (o > 1) o = 1;
            float x = o;                                               // smoothStep(0, 1, o), core/dplug/core/math.d:212-222
            float ss = x <= 0 ? 0.f : (x >= 1 ? 1.f : x * x * (3 - 2 * x));
            o = contrast * ss + (1 - contrast) * o;                    // lerp(o, ss, contrast)
            out[256 * ch + b] = (uint8_t)(0.5f + o * 255.0f);
        }
    }
}

int b2pbr_present(b2pbr* c, int pf, void* window, size_t window_pitch, int window_w, int window_h, int user_x, int user_y,
                  const b2_box2i* rects, int n, int redraw_borders) {
    if (!c || !c->outAlloc || !window || pf < 0 || pf > 2 || window_w <= 0 || window_h <= 0 || (n > 0 && !rects))
        return fail(B2_ERR_INVALID, "present: bad arguments");
    if (c->banded) return fail(B2_ERR_STATE, "present: not available on a row band");
    CU(cudaSetDevice(c->device));
    if (!c->winAlloc || c->win.w != window_w || c->win.h != window_h) {
        CU(cudaStreamSynchronize(c->stream));
        if (c->winAlloc) cudaFree(c->winAlloc);
        c->win.w = window_w; c->win.h = window_h; c->win.y0 = 0; c->win.rows = window_h;
        c->win.pitch = (int)alignUp((size_t)window_w * 4, 128);
        CU(cudaMalloc((void**)&c->winAlloc, (size_t)c->win.pitch * window_h));
        c->win.base = c->winAlloc;
        redraw_borders = 1;   // a new window buffer has no valid content yet (graphics.d:1091)
    }
    const b2_box2i userArea{user_x, user_y, user_x + c->W < window_w ? user_x + c->W : window_w, user_y + c->H < window_h ? user_y + c->H : window_h};
    if (user_x < 0 || user_y < 0) return fail(B2_ERR_INVALID, "present: negative user-area origin");
    std::vector<b2_box2i> work;
    if (redraw_borders) {
        // resizeContent: fill with black (RGBA/BGRA: 0,0,0,255; ARGB: 255,0,0,0), then copy the whole user area
        launch_fill32(c->win, pf == B2_PF_ARGB8 ? 0x000000ffu : 0xff000000u, c->stream);
        c->launches++;
        work.push_back(b2_box2i{0, 0, userArea.max_x - user_x, userArea.max_y - user_y});
    } else {
        for (int k = 0; k < n; ++k) {
            b2_box2i r = rects[k];   // user coordinates; clip to what is visible in the window
            if (r.min_x < 0) r.min_x = 0;
            if (r.min_y < 0) r.min_y = 0;
            if (r.max_x > userArea.max_x - user_x) r.max_x = userArea.max_x - user_x;
            if (r.max_y > userArea.max_y - user_y) r.max_y = userArea.max_y - user_y;
            if (r.max_x > c->W) r.max_x = c->W;
            if (r.max_y > c->H) r.max_y = c->H;
            if (boxW(r) > 0 && boxH(r) > 0) work.push_back(r);
        }
    }
    std::vector<b2_box2i> merged;
    mergeRects(work.data(), (int)work.size(), merged);
    for (const b2_box2i& r : merged) {
        launch_present(c->win, c->out, toBox(r), user_x, user_y, pf, c->lutOn ? c->dLut : nullptr, c->stream);
        c->launches++;
    }
    CU(cudaGetLastError());
    if (redraw_borders) {
        CU(cudaMemcpy2DAsync(window, window_pitch, c->win.base, (size_t)c->win.pitch, (size_t)window_w * 4, (size_t)window_h, cudaMemcpyDeviceToHost, c->stream));
    } else {
        for (const b2_box2i& r : merged)
            CU(cudaMemcpy2DAsync((uint8_t*)window + (size_t)(r.min_y + user_y) * window_pitch + (size_t)(r.min_x + user_x) * 4, window_pitch,
                                 c->win.base + (size_t)(r.min_y + user_y) * c->win.pitch + (size_t)(r.min_x + user_x) * 4, (size_t)c->win.pitch,
                                 (size_t)boxW(r) * 4, (size_t)boxH(r), cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

// ---- screenshot encoders (gui/dplug/gui/screencap.d:21-136), taken where doDraw takes them: on the
// rendered buffer of stage G.bis (graphics.d:848-857), i.e. the composited frame after the colour-correction tables and
// reorderComponents(pf).
static const size_t kQbHeader = 50;
size_t b2pbr_screenshot_qb_size(const b2pbr* c) {
    return (c && c->outAlloc) ? kQbHeader + (size_t)c->W * 26 * (size_t)c->H * 4 : 0;
}
int b2pbr_screenshot_qb(b2pbr* c, int pf, void* out, size_t cap) {
    if (!c || !c->outAlloc || !out || pf < 0 || pf > 2) return fail(B2_ERR_INVALID, "screenshot_qb: bad arguments");
    if (c->banded) return fail(B2_ERR_STATE, "screenshot: not available on a row band");
    const size_t voxBytes = (size_t)c->W * 26 * (size_t)c->H * 4;
    if (cap < kQbHeader + voxBytes) return fail(B2_ERR_INVALID, "screenshot_qb: buffer smaller than b2pbr_screenshot_qb_size()");
    CU(cudaSetDevice(c->device));
    uint8_t* o = (uint8_t*)out;
    size_t n = 0;
    auto put8 = [&](uint8_t v) { o[n++] = v; };
    auto put32 = [&](uint32_t v) { for (int k = 0; k < 4; ++k) put8((uint8_t)(v >> (8 * k))); };
    put8(1); put8(1); put8(0); put8(0);                    // .qb version 1.1.0.0 (screencap.d:33-36)
    put32(0); put32(0); put32(0); put32(0); put32(1);      // RGBA, left handed, uncompressed, alpha = visibility, one matrix
    put8(1); put8('0');                                    // matrix name "0"
    put32((uint32_t)c->W); put32(26u); put32((uint32_t)c->H);
    put32(0); put32(0); put32(0);                          // position
    uint32_t* dVox = nullptr;
    CU(cudaMalloc((void**)&dVox, voxBytes));
    launch_screenshot_qb(dVox, c->out, c->depth->levels[0], c->W, c->H, pf, c->lutOn ? c->dLut : nullptr, c->stream);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(o + kQbHeader, dVox, voxBytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(dVox);
    CU(e);
    return B2_OK;
}
// the pixel pass of encodeScreenshotAsPNG (screencap.d:88-131): the rendered buffer brought back from the window's pixel
// format to RGBA.  Re-ordering to `pf` and back is the identity on the colour channels and the alpha written by
// reorderComponents is 255 in every format, so the result does not depend on `pf`: colour-corrected r, g, b, 255.
int b2pbr_screenshot_rgba(b2pbr* c, int pf, void* out_rgba, size_t pitch) {
    if (!c || !c->outAlloc || !out_rgba || pf < 0 || pf > 2 || pitch < (size_t)c->W * 4) return fail(B2_ERR_INVALID, "screenshot_rgba: bad arguments");
    if (c->banded) return fail(B2_ERR_STATE, "screenshot: not available on a row band");
    CU(cudaSetDevice(c->device));
    DevLevel tmp = c->out;
    uint8_t* d = nullptr;
    CU(cudaMalloc((void**)&d, (size_t)tmp.pitch * c->H));
    tmp.base = d;
    launch_present(tmp, c->out, Box{0, 0, c->W, c->H}, 0, 0, B2_PF_RGBA8, c->lutOn ? c->dLut : nullptr, c->stream);
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(out_rgba, pitch, d, (size_t)tmp.pitch, (size_t)c->W * 4, (size_t)c->H, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d);
    CU(e);
    return B2_OK;
}
// encodeScreenshotAsPNG: the pixels above in a PNG container (8-bit RGBA, filter 0, deflate "stored" blocks — gamut's
// encoder is an un-vendored dependency and PNG bytes are not canonical; any decoder returns the reference's pixels).
static uint32_t crc32Update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t v = i;
            for (int k = 0; k < 8; ++k) v = (v & 1) ? 0xEDB88320u ^ (v >> 1) : v >> 1;
            table[i] = v;
        }
        init = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return crc;
}
size_t b2pbr_screenshot_png_size(const b2pbr* c) {
    if (!c || !c->outAlloc) return 0;
    const size_t raw = ((size_t)c->W * 4 + 1) * (size_t)c->H, blocks = (raw + 65534) / 65535;
    return 8 + 25 + (12 + 2 + raw + 5 * blocks + 4) + 12;
}
int b2pbr_screenshot_png(b2pbr* c, int pf, void* out, size_t cap, size_t* written) {
    if (!c || !c->outAlloc || !out) return fail(B2_ERR_INVALID, "screenshot_png: bad arguments");
    const size_t need = b2pbr_screenshot_png_size(c);
    if (cap < need) return fail(B2_ERR_INVALID, "screenshot_png: buffer smaller than b2pbr_screenshot_png_size()");
    const size_t rowBytes = (size_t)c->W * 4 + 1, raw = rowBytes * (size_t)c->H;
    std::vector<uint8_t> px(raw);
    if (int rc = b2pbr_screenshot_rgba(c, pf, px.data() + 1, rowBytes)) return rc;   // byte 0 of every scanline: filter type 0
    for (int y = 0; y < c->H; ++y) px[(size_t)y * rowBytes] = 0;
    uint8_t* o = (uint8_t*)out;
    size_t n = 0;
    auto be32 = [&](uint32_t v) { o[n++] = (uint8_t)(v >> 24); o[n++] = (uint8_t)(v >> 16); o[n++] = (uint8_t)(v >> 8); o[n++] = (uint8_t)v; };
    auto chunk = [&](const char* type, const uint8_t* data, size_t len) {
        be32((uint32_t)len);
        const size_t start = n;
        memcpy(o + n, type, 4); n += 4;
        if (len) { memcpy(o + n, data, len); n += len; }
        be32(crc32Update(0xffffffffu, o + start, n - start) ^ 0xffffffffu);
    };
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    memcpy(o, sig, 8); n = 8;
    uint8_t ihdr[13] = {(uint8_t)(c->W >> 24), (uint8_t)(c->W >> 16), (uint8_t)(c->W >> 8), (uint8_t)c->W,
                        (uint8_t)(c->H >> 24), (uint8_t)(c->H >> 16), (uint8_t)(c->H >> 8), (uint8_t)c->H, 8, 6, 0, 0, 0};
    chunk("IHDR", ihdr, 13);
    std::vector<uint8_t> z;
    z.reserve(raw + raw / 65535 * 5 + 16);
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;                                  // Adler-32 of the raw scanlines
    for (size_t off = 0; off < raw || off == 0; off += 65535) {
        const size_t len = raw - off < 65535 ? raw - off : 65535;
        z.push_back(off + len >= raw ? 1 : 0);            // BFINAL, BTYPE = 00 (stored)
        z.push_back((uint8_t)len); z.push_back((uint8_t)(len >> 8));
        z.push_back((uint8_t)~len); z.push_back((uint8_t)(~len >> 8));
        z.insert(z.end(), px.begin() + off, px.begin() + off + len);
        for (size_t i = 0; i < len; ++i) { a += px[off + i]; if (a >= 65521) a -= 65521; b += a; if (b >= 65521) b -= 65521; }
        if (raw == 0) break;
    }
    const uint32_t adler = (b << 16) | a;
    z.push_back((uint8_t)(adler >> 24)); z.push_back((uint8_t)(adler >> 16)); z.push_back((uint8_t)(adler >> 8)); z.push_back((uint8_t)adler);
    chunk("IDAT", z.data(), z.size());
    chunk("IEND", nullptr, 0);
    if (written) *written = n;
    return B2_OK;
}

int b2pbr_sync(b2pbr* c) {
    if (!c) return fail(B2_ERR_INVALID, "null compositor");
    CU(cudaStreamSynchronize(c->stream));
    return B2_OK;
}

int b2pbr_profile_enable(b2pbr* c, int on) {
    if (!c) return fail(B2_ERR_INVALID, "null compositor");
    c->profile = on != 0;
    return B2_OK;
}

int b2pbr_trace_json(b2pbr* c, char* buf, size_t cap, size_t* needed) {
    if (!c) return fail(B2_ERR_INVALID, "null compositor");
    CU(cudaStreamSynchronize(c->stream));
    // Chrome Trace Event Format, like TraceProfiler (core/dplug/core/profiler.d:177-207): timestamps in
    // microseconds relative to the first recorded event.
    std::string s = "[";
    char line[256];
    for (size_t i = 0; i < c->trace.size(); ++i) {
        float t0 = 0, dt = 0;
        cudaEventElapsedTime(&t0, c->trace[0].beg, c->trace[i].beg);
        cudaEventElapsedTime(&dt, c->trace[i].beg, c->trace[i].end);
        snprintf(line, sizeof(line), "%s{\"name\":\"%s\",\"cat\":\"gpu\",\"ph\":\"X\",\"ts\":%.3f,\"dur\":%.3f,\"pid\":0,\"tid\":%d}",
                 i ? "," : "", c->trace[i].name.c_str(), t0 * 1000.0, dt * 1000.0, c->device);
        s += line;
    }
    s += "]";
    if (needed) *needed = s.size() + 1;
    if (buf && cap > 0) {
        size_t nb = s.size() + 1 < cap ? s.size() + 1 : cap;
        memcpy(buf, s.c_str(), nb);
        buf[nb - 1] = 0;
    }
    return B2_OK;
}

uint64_t b2pbr_kernel_launches(const b2pbr* c) { return c ? c->launches : 0; }

}  // extern "C"

#include "band_exchange.inl"
