// Multi-GPU row bands: halo exchange and frame schedule behind the C ABI (b2band_*, include/b2pbr.h).
// Textually included at the end of b2pbr.cu (it uses that file's b2pbr / b2mip internals).
//
// BASELINE config 5 / SURVEY section 8e: one large canvas is cut into horizontal bands of whole 64-row GUI tiles
// (gui/dplug/gui/graphics.d:59-60), one band per GPU.  A band object (b2pbr_create_band) stores only its own rows plus
// the rows its samplers and down-samplers reach beyond them (computeBandNeeds: 94 diffuse / 28 depth rows per side,
// SURVEY Appendix C).  The ONE exchange step per frame moves the level-0 halo rows (diffuse RGBA8 + depth L16)
// between neighbouring bands; material needs no halo.  Two transports:
//   * NCCL  — one process per GPU (torchrun, or a D host per GPU): ncclSend / ncclRecv over NVLink, grouped, on the
//             band's exchange stream.  libnccl.so.2 is loaded at run time (dlopen): the library has no link-time
//             dependency on it, and a caller-owned ncclComm_t can be passed instead of letting the library create one.
//   * PEERS — all bands in one process (one host thread drives several GPUs): every band pushes its boundary rows into
//             its neighbours' halo rows with cudaMemcpyPeerAsync (copy engines over NVLink, no SM is used), ordered by
//             events.  With every band on one device this degenerates to device-to-device copies (the 1-GPU test).
// Frame schedule (b2band_composite_frame), bit-identical to the unbanded frame: while the halo rows travel, the band
// regenerates the mip rows that depend only on its own rows and composites its interior tiles (those that read neither
// a halo row nor a mip texel the halo pass writes); on a second stream the halo pass — the impactOnNextLevel chains of
// the halo rectangles, the same argument that makes the reference's dirty-rectangle updates correct
// (graphics.d:1378-1419) — runs as soon as the rows have arrived; the boundary tiles follow.
#include <dlfcn.h>

namespace {

// ---- NCCL through dlopen ---------------------------------------------------------------------------------------
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Send)(const void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::string error;
};
constexpr int kNcclUint8 = 1;   // ncclUint8 (nccl.h: ncclInt8 = 0, ncclUint8 = 1)

NcclApi* ncclApi() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    const char* names[] = {getenv("B2_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n || !*n) continue;
        api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.handle) break;
    }
    if (!api.handle) { api.error = std::string("cannot load libnccl.so.2 (set B2_NCCL_LIB): ") + (dlerror() ? dlerror() : ""); return nullptr; }
    bool ok = true;
    auto sym = [&](const char* name) { void* p = dlsym(api.handle, name); if (!p) { ok = false; api.error = std::string("libnccl lacks ") + name; } return p; };
    api.GetUniqueId = (int (*)(NcclUniqueId*))sym("ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))sym("ncclCommInitRank");
    api.CommDestroy = (int (*)(NcclComm))sym("ncclCommDestroy");
    api.Send = (int (*)(const void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclSend");
    api.Recv = (int (*)(void*, size_t, int, int, NcclComm, cudaStream_t))sym("ncclRecv");
    api.GroupStart = (int (*)())sym("ncclGroupStart");
    api.GroupEnd = (int (*)())sym("ncclGroupEnd");
    api.GetErrorString = (const char* (*)(int))sym("ncclGetErrorString");
    if (!ok) { dlclose(api.handle); api.handle = nullptr; return nullptr; }
    return &api;
}

#define NC(call)                                                                                         \
    do {                                                                                                 \
        int r_ = (call);                                                                                 \
        if (r_ != 0) return fail(B2_ERR_CUDA, std::string(#call) + ": " + ncclApi()->GetErrorString(r_)); \
    } while (0)

// ---- planning (pure host arithmetic) -----------------------------------------------------------------------------
// rows [y0, y1) of band `rank`: equal shares of the `align`-row GUI tiles
void bandRowsOf(int H, int rank, int nranks, int align, int& y0, int& y1) {
    const int tiles = (H + align - 1) / align, base = tiles / nranks, extra = tiles % nranks;
    const int t0 = rank * base + (rank < extra ? rank : extra), t1 = t0 + base + (rank < extra ? 1 : 0);
    y0 = t0 * align < H ? t0 * align : H;
    y1 = t1 * align < H ? t1 * align : H;
}

struct BandPlan {
    b2_box2i own;                    // the band's own rows
    b2_box2i halo[2]; int nHalo;     // level-0 halo rectangles (above, below): what the halo pass regenerates from
    int interiorLo, interiorHi;      // rows (whole tiles) that can be composited before the halos exist
};

// rows of levels 1.. written by regenerateMipmaps({rect}) for one pyramid
void rowsWritten(int W, int H, b2_box2i area, const int* q, int nl, int (*out)[2]) {
    int w = W, h = H;
    for (int l = 1; l <= nl; ++l) {
        area = impactOnNextLevel(q[l - 1], area, w, h);
        out[l][0] = area.min_y; out[l][1] = area.max_y;
        w >>= 1; h >>= 1;
    }
}

void planBand(int W, int H, int y0, int y1, int align, BandPlan& P) {
    int nd[kMaxDiffuseLevels][2], nz[kMaxDepthLevels][2];
    computeBandNeeds(W, H, y0, y1, nd, nz);
    P.own = b2_box2i{0, y0, W, y1};
    P.nHalo = 0;
    // the widest halo is the diffuse one (level-5 bicubic bloom); the depth rows beyond its own need are never read
    if (nd[0][0] < y0) P.halo[P.nHalo++] = b2_box2i{0, nd[0][0], W, y0};
    if (nd[0][1] > y1) P.halo[P.nHalo++] = b2_box2i{0, y1, W, nd[0][1]};
    const int nD = countLevels(5, W, H) - 1, nZ = countLevels(4, W, H) - 1;
    const int qD[5] = {B2_QUALITY_BOX_ALPHACOV_PREMUL, B2_QUALITY_CUBIC, B2_QUALITY_CUBIC, B2_QUALITY_CUBIC, B2_QUALITY_CUBIC};
    const int qZ[4] = {B2_QUALITY_BOX, B2_QUALITY_BOX, B2_QUALITY_CUBIC, B2_QUALITY_CUBIC};
    int a = y0, b = y1;
    for (int k = 0; k < P.nHalo; ++k) {
        const bool top = P.halo[k].min_y < y0;
        int wD[kMaxDiffuseLevels][2] = {}, wZ[kMaxDepthLevels][2] = {};
        rowsWritten(W, H, P.halo[k], qD, nD, wD);
        rowsWritten(W, H, P.halo[k], qZ, nZ, wZ);
        while (a < b) {
            const int c0 = top ? a : (b - align > a ? b - align : a), c1 = top ? (a + align < b ? a + align : b) : b;
            int cd[kMaxDiffuseLevels][2], cz[kMaxDepthLevels][2];
            computeBandNeeds(W, H, c0, c1, cd, cz);
            bool clash = cd[0][0] < y0 || cd[0][1] > y1 || cz[0][0] < y0 || cz[0][1] > y1;
            for (int l = 1; l <= nD && !clash; ++l) clash = cd[l][0] < wD[l][1] && wD[l][0] < cd[l][1];
            for (int l = 1; l <= nZ && !clash; ++l) clash = cz[l][0] < wZ[l][1] && wZ[l][0] < cz[l][1];
            if (!clash) break;
            if (top) a = c1; else b = c0;
        }
    }
    P.interiorLo = a; P.interiorHi = b;
}

struct HaloCopy { int map; int peer; int lo, hi; };   // rows [lo, hi) of level 0 of `map` (B2_MAP_*), to / from band `peer`

}  // namespace

struct b2band {
    int transport = 0, nranks = 0;
    std::vector<int> bandY;            // nranks + 1 boundaries
    struct Local {
        b2pbr* c = nullptr; int rank = 0;
        cudaStream_t xs = nullptr, side = nullptr;
        cudaEvent_t evReady = nullptr, evX = nullptr, evOwn = nullptr, evHalo = nullptr, evDone = nullptr;
        BandPlan plan;
        std::vector<HaloCopy> sends, recvs;
        std::vector<b2_box2i> interior, boundary, all;
        bool begun = false;
        // B2_BAND_DEBUG: per-phase device times (events with timing), accumulated over the frames and printed by b2band_destroy
        cudaEvent_t t[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // frame start, own mips, interior, arrival, halo mips, boundary
        double acc[5] = {0, 0, 0, 0, 0}; int frames = 0; bool timed = false;
    };
    std::vector<Local> locals;         // NCCL: the one band of this process; PEERS: every band, index = rank
    NcclComm comm = nullptr; bool ownComm = false;
    uint64_t haloBytesPerFrame = 0;    // received by the local bands
};

namespace {

uint8_t* level0Rows(b2mip* m, int row) {
    const DevLevel& L = m->levels[0];
    return L.base + (size_t)(row - L.y0) * (size_t)L.pitch;
}
b2mip* mapOf(b2pbr* c, int which) { return which == B2_MAP_DIFFUSE ? c->diffuse : (which == B2_MAP_DEPTH ? c->depth : c->material); }

int bandSetupLocal(b2band* g, b2band::Local& L) {
    b2pbr* c = L.c;
    if (!c || !c->banded || !c->diffuse) return fail(B2_ERR_STATE, "b2band: every band needs b2pbr_create_band + b2pbr_resize first");
    const int W = c->W, H = c->H, r = L.rank, n = g->nranks;
    if (c->bandY0 != g->bandY[r] || c->bandY1 != g->bandY[r + 1]) return fail(B2_ERR_INVALID, "b2band: band rows differ from the band boundaries given");
    if (g->bandY[0] != 0 || g->bandY[n] != H) return fail(B2_ERR_INVALID, "b2band: the bands must cover the canvas");
    planBand(W, H, c->bandY0, c->bandY1, c->tileH, L.plan);
    // what this band receives, and what its neighbours need from it (the same arithmetic, from their side)
    for (int side = -1; side <= 1; side += 2) {
        const int p = r + side;
        if (p < 0 || p >= n) continue;
        int nd[kMaxDiffuseLevels][2], nz[kMaxDepthLevels][2];
        computeBandNeeds(W, H, c->bandY0, c->bandY1, nd, nz);
        const int (*mine[2])[2] = {nd, nz};
        int pd[kMaxDiffuseLevels][2], pz[kMaxDepthLevels][2];
        computeBandNeeds(W, H, g->bandY[p], g->bandY[p + 1], pd, pz);
        const int (*theirs[2])[2] = {pd, pz};
        const int maps[2] = {B2_MAP_DIFFUSE, B2_MAP_DEPTH};
        for (int k = 0; k < 2; ++k) {
            if (side < 0) {
                if (mine[k][0][0] < c->bandY0) {
                    if (mine[k][0][0] < g->bandY[p]) return fail(B2_ERR_INVALID, "b2band: a halo reaches past the adjacent band (bands thinner than the bloom reach)");
                    L.recvs.push_back(HaloCopy{maps[k], p, mine[k][0][0], c->bandY0});
                }
                if (theirs[k][0][1] > g->bandY[p + 1]) {
                    if (theirs[k][0][1] > c->bandY1) return fail(B2_ERR_INVALID, "b2band: a halo reaches past the adjacent band (bands thinner than the bloom reach)");
                    L.sends.push_back(HaloCopy{maps[k], p, c->bandY0, theirs[k][0][1]});
                }
            } else {
                if (mine[k][0][1] > c->bandY1) {
                    if (mine[k][0][1] > g->bandY[p + 1]) return fail(B2_ERR_INVALID, "b2band: a halo reaches past the adjacent band (bands thinner than the bloom reach)");
                    L.recvs.push_back(HaloCopy{maps[k], p, c->bandY1, mine[k][0][1]});
                }
                if (theirs[k][0][0] < g->bandY[p]) {
                    if (theirs[k][0][0] < c->bandY0) return fail(B2_ERR_INVALID, "b2band: a halo reaches past the adjacent band (bands thinner than the bloom reach)");
                    L.sends.push_back(HaloCopy{maps[k], p, theirs[k][0][0], c->bandY1});
                }
            }
        }
    }
    for (const HaloCopy& h : L.recvs) g->haloBytesPerFrame += (uint64_t)(h.hi - h.lo) * (uint64_t)mapOf(c, h.map)->levels[0].pitch;
    const b2_box2i whole{0, c->bandY0, W, c->bandY1};
    tileAreas(&whole, 1, c->tileW, c->tileH, L.all);
    if (L.plan.interiorHi > L.plan.interiorLo) {
        const b2_box2i in{0, L.plan.interiorLo, W, L.plan.interiorHi};
        tileAreas(&in, 1, c->tileW, c->tileH, L.interior);
    }
    const b2_box2i edges[2] = {{0, c->bandY0, W, L.plan.interiorLo}, {0, L.plan.interiorHi, W, c->bandY1}};
    for (const b2_box2i& e : edges)
        if (e.max_y > e.min_y) tileAreas(&e, 1, c->tileW, c->tileH, L.boundary);
    CU(cudaSetDevice(c->device));
    // The transfers and the halo pass are short and everything after them waits for them, while the band's own stream
    // is busy with a lighting launch of thousands of CTAs: on high-priority streams their CTAs are dispatched as soon as
    // an SM slot frees up instead of after the last CTA of that launch (measured at 2 GPUs: the halo mips otherwise
    // start 3.2 ms after the rows arrived, i.e. when the interior tiles are done).
    int least = 0, greatest = 0;
    CU(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    CU(cudaStreamCreateWithPriority(&L.xs, cudaStreamNonBlocking, greatest));
    CU(cudaStreamCreateWithPriority(&L.side, cudaStreamNonBlocking, greatest));
    for (cudaEvent_t* e : {&L.evReady, &L.evX, &L.evOwn, &L.evHalo, &L.evDone}) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    if (getenv("B2_BAND_DEBUG"))
        for (cudaEvent_t& e : L.t) CU(cudaEventCreate(&e));
    return B2_OK;
}

// phase A of a frame: start the transfers of this band, then the work that does not depend on them
int bandBegin(b2band* g, b2band::Local& L, bool composite) {
    b2pbr* c = L.c;
    CU(cudaSetDevice(c->device));
    if (L.t[0] && composite) {
        if (L.timed && cudaEventQuery(L.t[5]) == cudaSuccess && cudaEventQuery(L.t[4]) == cudaSuccess) {
            const int a[5] = {0, 1, 0, 3, 0}, b[5] = {1, 2, 3, 4, 5};   // own mips, interior tiles, arrival since start, halo mips, whole frame
            for (int k = 0; k < 5; ++k) { float ms = 0; cudaEventElapsedTime(&ms, L.t[a[k]], L.t[b[k]]); L.acc[k] += ms; }
            L.frames++;
        }
        cudaGetLastError();
        CU(cudaEventRecord(L.t[0], c->stream));
    }
    // the transfers read this band's own rows: after everything the caller enqueued on the band's stream (its uploads)
    CU(cudaEventRecord(L.evReady, c->stream));
    CU(cudaStreamWaitEvent(L.xs, L.evReady, 0));
    if (g->transport == B2_BAND_TRANSPORT_NCCL) {
        // (halo rows are overwritten: not before this band's previous frame has finished reading them — evReady, recorded
        // on the same stream after that frame, covers it)
        if (!L.sends.empty() || !L.recvs.empty()) {
            NcclApi* api = ncclApi();
            NC(api->GroupStart());
            for (const HaloCopy& h : L.sends) {
                b2mip* m = mapOf(c, h.map);
                NC(api->Send(level0Rows(m, h.lo), (size_t)(h.hi - h.lo) * (size_t)m->levels[0].pitch, kNcclUint8, h.peer, g->comm, L.xs));
            }
            for (const HaloCopy& h : L.recvs) {
                b2mip* m = mapOf(c, h.map);
                NC(api->Recv(level0Rows(m, h.lo), (size_t)(h.hi - h.lo) * (size_t)m->levels[0].pitch, kNcclUint8, h.peer, g->comm, L.xs));
            }
            NC(api->GroupEnd());
        }
    } else {
        // push: this band's boundary rows into the neighbours' halo rows, once the neighbour's previous frame no longer
        // reads them (evDone of that band; an event that was never recorded does not block)
        for (const HaloCopy& h : L.sends) {
            b2band::Local& P = g->locals[h.peer];
            CU(cudaStreamWaitEvent(L.xs, P.evDone, 0));
        }
        for (const HaloCopy& h : L.sends) {
            b2band::Local& P = g->locals[h.peer];
            b2mip *src = mapOf(c, h.map), *dst = mapOf(P.c, h.map);
            const size_t bytes = (size_t)(h.hi - h.lo) * (size_t)src->levels[0].pitch;
            if (P.c->device == c->device) CU(cudaMemcpyAsync(level0Rows(dst, h.lo), level0Rows(src, h.lo), bytes, cudaMemcpyDeviceToDevice, L.xs));
            else CU(cudaMemcpyPeerAsync(level0Rows(dst, h.lo), P.c->device, level0Rows(src, h.lo), c->device, bytes, L.xs));
        }
    }
    CU(cudaEventRecord(L.evX, L.xs));
    if (L.t[0] && composite) CU(cudaEventRecord(L.t[3], L.xs));
    L.begun = true;
    if (!composite) return B2_OK;
    if (L.plan.nHalo == 0) return B2_OK;   // a single band: everything happens in phase B
    if (int rc = b2pbr_regenerate_mipmaps(c, &L.plan.own, 1)) return rc;
    CU(cudaEventRecord(L.evOwn, c->stream));
    if (L.t[0]) CU(cudaEventRecord(L.t[1], c->stream));
    if (!L.interior.empty())
        if (int rc = pbrComposite(c, L.interior.data(), (int)L.interior.size(), nullptr, nullptr, nullptr)) return rc;
    if (L.t[0]) CU(cudaEventRecord(L.t[2], c->stream));
    return B2_OK;
}

// `st` waits until this band's halo rows of the current frame have arrived
int bandWaitArrival(b2band* g, b2band::Local& L, cudaStream_t st) {
    if (g->transport == B2_BAND_TRANSPORT_NCCL) { CU(cudaStreamWaitEvent(st, L.evX, 0)); return B2_OK; }
    for (const HaloCopy& h : L.recvs) CU(cudaStreamWaitEvent(st, g->locals[h.peer].evX, 0));
    return B2_OK;
}

// phase B: halo pass + boundary tiles
int bandFinish(b2band* g, b2band::Local& L, bool composite) {
    b2pbr* c = L.c;
    CU(cudaSetDevice(c->device));
    if (!L.begun) return fail(B2_ERR_STATE, "b2band: exchange_end / frame without exchange_begin");
    L.begun = false;
    if (!composite) {
        if (int rc = bandWaitArrival(g, L, c->stream)) return rc;
        CU(cudaStreamWaitEvent(c->stream, L.evX, 0));   // and this band's own rows are free to change again
        return B2_OK;
    }
    if (L.plan.nHalo == 0) {
        if (int rc = b2pbr_regenerate_mipmaps(c, &L.plan.own, 1)) return rc;
        if (int rc = pbrComposite(c, L.all.data(), (int)L.all.size(), nullptr, nullptr, nullptr)) return rc;
    } else {
        CU(cudaStreamWaitEvent(L.side, L.evOwn, 0));
        if (int rc = bandWaitArrival(g, L, L.side)) return rc;
        c->useAuxHi = true;
        const int rcHalo = b2pbr_regenerate_mipmaps_on(c, L.plan.halo, L.plan.nHalo, L.side);
        c->useAuxHi = false;
        if (rcHalo) return rcHalo;
        CU(cudaEventRecord(L.evHalo, L.side));
        if (L.t[0]) CU(cudaEventRecord(L.t[4], L.side));
        CU(cudaStreamWaitEvent(c->stream, L.evHalo, 0));
        if (!L.boundary.empty())
            if (int rc = pbrComposite(c, L.boundary.data(), (int)L.boundary.size(), nullptr, nullptr, nullptr)) return rc;
        if (L.t[0]) { CU(cudaEventRecord(L.t[5], c->stream)); L.timed = true; }
    }
    CU(cudaStreamWaitEvent(c->stream, L.evX, 0));       // this band's sends have left: its rows may change again
    CU(cudaEventRecord(L.evDone, c->stream));           // its halo rows may be overwritten by the next frame's transfers
    return B2_OK;
}

}  // namespace

extern "C" {

int b2_band_rows(int H, int nranks, int rank, int align, int* y0, int* y1) {
    if (H <= 0 || nranks <= 0 || rank < 0 || rank >= nranks || align <= 0 || !y0 || !y1) return fail(B2_ERR_INVALID, "b2_band_rows: bad arguments");
    bandRowsOf(H, rank, nranks, align, *y0, *y1);
    return B2_OK;
}

int b2_band_plan(int W, int H, int y0, int y1, int align, b2_box2i halo_rects[2], int* n_halo, int* interior_lo, int* interior_hi) {
    if (W <= 0 || H <= 0 || y0 < 0 || y1 > H || y0 >= y1 || align <= 0) return fail(B2_ERR_INVALID, "b2_band_plan: bad arguments");
    BandPlan P;
    planBand(W, H, y0, y1, align, P);
    if (halo_rects) for (int k = 0; k < P.nHalo; ++k) halo_rects[k] = P.halo[k];
    if (n_halo) *n_halo = P.nHalo;
    if (interior_lo) *interior_lo = P.interiorLo;
    if (interior_hi) *interior_hi = P.interiorHi;
    return B2_OK;
}

int b2_band_nccl_unique_id(void* id128) {
    if (!id128) return fail(B2_ERR_INVALID, "b2_band_nccl_unique_id: null");
    NcclApi* api = ncclApi();
    if (!api) return fail(B2_ERR_STATE, "NCCL unavailable: cannot load libnccl.so.2 (set B2_NCCL_LIB to its path)");
    NcclUniqueId id;
    NC(api->GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return B2_OK;
}

int b2band_create_nccl(b2band** out, b2pbr* band, int rank, int nranks, const int* band_y, const void* id128, void* nccl_comm) {
    if (!out || !band || nranks <= 0 || rank < 0 || rank >= nranks || !band_y || (!id128 && !nccl_comm && nranks > 1))
        return fail(B2_ERR_INVALID, "b2band_create_nccl: bad arguments");
    b2band* g = new b2band();
    g->transport = B2_BAND_TRANSPORT_NCCL; g->nranks = nranks;
    g->bandY.assign(band_y, band_y + nranks + 1);
    g->locals.resize(1);
    g->locals[0].c = band; g->locals[0].rank = rank;
    int rc = bandSetupLocal(g, g->locals[0]);
    if (rc) { delete g; return rc; }
    if (nranks > 1) {
        NcclApi* api = ncclApi();
        if (!api) { delete g; return fail(B2_ERR_STATE, "b2band_create_nccl: cannot load libnccl.so.2 (set B2_NCCL_LIB to its path)"); }
        if (nccl_comm) g->comm = (NcclComm)nccl_comm;
        else {
            NcclUniqueId id;
            memcpy(id.internal, id128, 128);
            cudaSetDevice(band->device);
            int r = api->CommInitRank(&g->comm, nranks, id, rank);   // collective: every rank calls this
            if (r != 0) { delete g; return fail(B2_ERR_CUDA, std::string("ncclCommInitRank: ") + api->GetErrorString(r)); }
            g->ownComm = true;
        }
    }
    *out = g;
    return B2_OK;
}

int b2band_create_peers(b2band** out, b2pbr* const* bands, int nranks) {
    if (!out || !bands || nranks <= 0) return fail(B2_ERR_INVALID, "b2band_create_peers: bad arguments");
    b2band* g = new b2band();
    g->transport = B2_BAND_TRANSPORT_PEERS; g->nranks = nranks;
    g->bandY.resize(nranks + 1);
    for (int r = 0; r < nranks; ++r) {
        if (!bands[r]) { delete g; return fail(B2_ERR_INVALID, "b2band_create_peers: null band"); }
        g->bandY[r] = bands[r]->bandY0;
        g->bandY[r + 1] = bands[r]->bandY1;
    }
    g->locals.resize(nranks);
    for (int r = 0; r < nranks; ++r) {
        g->locals[r].c = bands[r]; g->locals[r].rank = r;
        if (r > 0 && bands[r]->bandY0 != bands[r - 1]->bandY1) { delete g; return fail(B2_ERR_INVALID, "b2band_create_peers: bands are not adjacent in rank order"); }
    }
    for (int r = 0; r < nranks; ++r) {
        int rc = bandSetupLocal(g, g->locals[r]);
        if (rc) { delete g; return rc; }
    }
    // neighbouring bands on different devices: direct peer access, so the copies go GPU to GPU over NVLink
    for (int r = 0; r + 1 < nranks; ++r) {
        const int a = bands[r]->device, b = bands[r + 1]->device;
        if (a == b) continue;
        for (int k = 0; k < 2; ++k) {
            const int from = k ? b : a, to = k ? a : b;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, from, to);
            if (!can) continue;                       // cudaMemcpyPeerAsync then stages through the host: still correct
            cudaSetDevice(from);
            cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { delete g; return fail(B2_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
            cudaGetLastError();
        }
    }
    *out = g;
    return B2_OK;
}

void b2band_destroy(b2band* g) {
    if (!g) return;
    for (auto& L : g->locals) {
        if (!L.c) continue;
        cudaSetDevice(L.c->device);
        if (L.xs) { cudaStreamSynchronize(L.xs); cudaStreamDestroy(L.xs); }
        if (L.side) { cudaStreamSynchronize(L.side); cudaStreamDestroy(L.side); }
        for (cudaEvent_t e : {L.evReady, L.evX, L.evOwn, L.evHalo, L.evDone})
            if (e) cudaEventDestroy(e);
        if (L.t[0] && L.frames > 0)
            fprintf(stderr, "[b2band] rank %d rows [%d,%d) interior [%d,%d): per frame over %d frames: own mips %.3f ms, interior tiles %.3f ms, halo rows arrived at +%.3f ms, "
                            "halo mips %.3f ms after that, frame %.3f ms\n", L.rank, L.c->bandY0, L.c->bandY1, L.plan.interiorLo, L.plan.interiorHi, L.frames,
                    L.acc[0] / L.frames, L.acc[1] / L.frames, L.acc[2] / L.frames, L.acc[3] / L.frames, L.acc[4] / L.frames);
        for (cudaEvent_t e : L.t)
            if (e) cudaEventDestroy(e);
    }
    if (g->ownComm && g->comm && ncclApi()) ncclApi()->CommDestroy(g->comm);
    delete g;
}

int b2band_exchange_begin(b2band* g) {
    if (!g) return fail(B2_ERR_INVALID, "null band group");
    for (auto& L : g->locals)
        if (int rc = bandBegin(g, L, false)) return rc;
    return B2_OK;
}
int b2band_exchange_end(b2band* g) {
    if (!g) return fail(B2_ERR_INVALID, "null band group");
    for (auto& L : g->locals)
        if (int rc = bandFinish(g, L, false)) return rc;
    return B2_OK;
}
int b2band_composite_frame(b2band* g) {
    if (!g) return fail(B2_ERR_INVALID, "null band group");
    for (auto& L : g->locals)
        if (int rc = bandBegin(g, L, true)) return rc;
    for (auto& L : g->locals)
        if (int rc = bandFinish(g, L, true)) return rc;
    return B2_OK;
}
int b2band_sync(b2band* g) {
    if (!g) return fail(B2_ERR_INVALID, "null band group");
    for (auto& L : g->locals) {
        CU(cudaSetDevice(L.c->device));
        CU(cudaStreamSynchronize(L.c->stream));
        CU(cudaStreamSynchronize(L.xs));
        CU(cudaStreamSynchronize(L.side));
    }
    return B2_OK;
}
uint64_t b2band_halo_bytes_per_frame(const b2band* g) { return g ? g->haloBytesPerFrame : 0; }
int b2band_info(const b2band* g, int local_index, int* rank, int* interior_lo, int* interior_hi, int* n_sends, int* n_recvs) {
    if (!g || local_index < 0 || local_index >= (int)g->locals.size()) return fail(B2_ERR_INVALID, "b2band_info: bad arguments");
    const auto& L = g->locals[local_index];
    if (rank) *rank = L.rank;
    if (interior_lo) *interior_lo = L.plan.interiorLo;
    if (interior_hi) *interior_hi = L.plan.interiorHi;
    if (n_sends) *n_sends = (int)L.sends.size();
    if (n_recvs) *n_recvs = (int)L.recvs.size();
    return B2_OK;
}

}  // extern "C"
