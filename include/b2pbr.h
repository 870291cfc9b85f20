/* b2pbr — C ABI of the B200-native (sm_100a) PBR compositor + mipmap pyramid.
 *
 * This is the drop-in boundary for ONE path of AuburnSounds/Dplug: the dplug:gui
 * PBR compositor and the dplug:graphics Mipmap pyramid.  Every entry point names
 * the reference interface it replaces (paths relative to the Dplug tree).  All
 * arguments are plain pointers, sizes and PODs; no torch / C++ types.  The D-side
 * binding a maintainer would add is shown in INTEGRATION.md and d/.
 *
 * Conventions
 *   - every function returns 0 on success, a B2_ERR_* code otherwise; the
 *     reference's methods are `nothrow @nogc` and return void, so the D shim
 *     asserts on non-zero.  b2_last_error() gives a message for the calling thread.
 *   - there is NO CPU fallback: if no CUDA device is usable, create() fails.
 *   - work is enqueued on the object's CUDA stream (b2*_set_stream, default: a
 *     private non-blocking stream).  Functions that read results back to host
 *     memory synchronise that stream; everything else is asynchronous.
 *   - rectangles are half-open [min,max) in pixels, as box2i.
 */
#ifndef B2PBR_H
#define B2PBR_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B2PBR_ABI_VERSION 2

/* box2i — math/dplug/math/box.d:29-30,51-55 ({min.x,min.y,max.x,max.y}, 4 x int32) */
typedef struct { int min_x, min_y, max_x, max_y; } b2_box2i;
/* ImageRef!T — graphics/dplug/graphics/image.d:127-131 (pitch in bytes) */
typedef struct { int w, h; size_t pitch; void* pixels; } b2_imageref;

enum { B2_OK = 0, B2_ERR_INVALID = 1, B2_ERR_CUDA = 2, B2_ERR_NO_DEVICE = 3, B2_ERR_STATE = 4, B2_ERR_NOMEM = 5 };

/* pixel formats — graphics/dplug/graphics/color.d:307-321 (RGBA: r in the low byte) */
enum { B2_COLOR_RGBA8 = 0, B2_COLOR_L16 = 1 };
/* Mipmap!T.Quality — graphics/dplug/graphics/mipmap.d:36-46 (same order) */
enum { B2_QUALITY_BOX = 0, B2_QUALITY_CUBIC = 1, B2_QUALITY_BOX_ALPHACOV_PREMUL = 2 };
/* pass indices — gui/dplug/gui/legacypbr.d:129-139 */
enum {
    B2_PASS_NORMAL = 0, B2_PASS_AO = 1, B2_PASS_OBLIQUE_SHADOW = 2, B2_PASS_DIRECTIONAL = 3,
    B2_PASS_SPECULAR = 4, B2_PASS_SKYBOX = 5, B2_PASS_EMISSIVE = 6, B2_PASS_CLAMP = 7
};
/* the maps GUIGraphics owns and lends to the compositor — gui/dplug/gui/graphics.d:117-122 */
enum { B2_MAP_DIFFUSE = 0, B2_MAP_MATERIAL = 1, B2_MAP_DEPTH = 2 };
/* lighting kernels (all but EXACT under the +-1 LSB contract of BASELINE.json):
 *   FAST            = throughput kernel: the bilinear mip samplers and the skybox samples run on the texture units
 *                     (pixel-centre samples of the pyramids have weights the 8-bit hardware filter represents exactly;
 *                     the skybox samples do not: their weight error is at most 2^-8 per axis, i.e. at most
 *                     skyboxAmount * metalness * (|dAB| + |dCD|) / 256 LSB of a channel — measured 0.2 % of channels
 *                     flipping by 1 LSB on the synthetic frames);
 *   FAST_SKY_GATHER = the same kernel with the skybox footprints fetched by texture gather and blended with the
 *                     reference's FP32 weights (15 % slower);
 *   EXACT           = the reference's IEEE operation order, bit-identical to the CPU restatement (several times slower;
 *                     the device-side yardstick). */
enum { B2_MODE_FAST = 0, B2_MODE_EXACT = 1, B2_MODE_FAST_SKY_GATHER = 3 };
/* window pixel orders for the fused post-composite swizzle — window/dplug/window/window.d WindowPixelFormat */
enum { B2_PF_RGBA8 = 0, B2_PF_BGRA8 = 1, B2_PF_ARGB8 = 2 };

const char* b2_last_error(void);
int b2_abi_version(void);
/* number of usable CUDA devices (0 if none) */
int b2_device_count(void);

/* ------------------------------------------------------------------------------------------
 * b2mip — device-resident twin of Mipmap!RGBA / Mipmap!L16 (graphics/dplug/graphics/mipmap.d:30-646)
 * ------------------------------------------------------------------------------------------ */
typedef struct b2mip b2mip;

/* Mipmap.this(maxLevel, w, h) + size() — mipmap.d:60-64, 80-126.  Level l is (w>>l) x (h>>l);
 * the number of levels is min(maxLevel+1, halvings until a dimension is 0). */
int b2mip_create(b2mip** out, int device, int color, int max_level, int w, int h);
/* row-banded variant: the object stores only rows [y0 - halo_up, y1 + halo_down) of the
 * global w x h level 0 (and the matching rows of every level); coordinates stay global. */
int b2mip_create_band(b2mip** out, int device, int color, int max_level, int w, int h, int y0, int y1,
                      int halo_up, int halo_down);
/* ~this — mipmap.d:128-132 */
void b2mip_destroy(b2mip*);
int b2mip_set_stream(b2mip*, void* cuda_stream);
/* numLevels / width / height — mipmap.d:499-514 */
int b2mip_num_levels(const b2mip*);
int b2mip_width(const b2mip*);
int b2mip_height(const b2mip*);
/* levels[level] as a *device* ImageRef (pixels = device address of the first stored row; for banded objects
 * `first_row` receives the global row index that address corresponds to) */
int b2mip_level(const b2mip*, int level, b2_imageref* out, int* first_row);
/* host -> device copy of `n` rectangles of one level (host image addressed like ImageRef: `host` = pixel (0,0)) */
int b2mip_upload(b2mip*, int level, const void* host, size_t host_pitch, const b2_box2i* rects, int n);
/* device -> host, synchronises */
int b2mip_download(b2mip*, int level, void* host, size_t host_pitch, const b2_box2i* rects, int n);
/* Mipmap.generateNextLevel(quality, updateRectPreviousLevel, level) — mipmap.d:534-540; returns the updated
 * rectangle of `level` (impactOnNextLevel, mipmap.d:623-645) in *out_rect */
int b2mip_generate_next_level(b2mip*, int quality, b2_box2i update_rect_previous_level, int level, b2_box2i* out_rect);
/* Mipmap.generateLevel(level, quality, updateRect) — mipmap.d:544-618 */
int b2mip_generate_level(b2mip*, int level, int quality, b2_box2i update_rect);
/* Mipmap.generateMipmaps(quality) — mipmap.d:517-528 (box silently becomes cubic from level 3) */
int b2mip_generate_mipmaps(b2mip*, int quality);
/* Whole pyramid from level 0 in fused launches (diffuse recipe: level 1 boxAlphaCovIntoPremul, levels >= 2
 * cubic — gui/dplug/gui/graphics.d:1378-1392; depth recipe: levels 1,2 box, >= 3 cubic — graphics.d:1412-1419).
 * Same results as the level-by-level chain over the full rectangle. */
int b2mip_generate_chain(b2mip*, int first_quality, int cubic_from_level);
/* `n` successive generateNextLevel calls — levels first_level .. first_level + n - 1 with qualities[k], the update
 * rectangle propagated by impactOnNextLevel from `update_rect_previous_level` (what GUIGraphics.regenerateMipmaps does
 * per dirty rectangle, graphics.d:1378-1419) — executed in launches of up to three levels that keep the intermediate
 * levels in shared memory.  Bit-identical to the level-by-level chain, including when the stored levels are stale
 * outside the rectangles.  out_rects (optional) receives the n updated rectangles. */
int b2mip_generate_levels(b2mip*, int first_level, int n, const int* qualities, b2_box2i update_rect_previous_level,
                          b2_box2i* out_rects);
/* tuning / debugging knob (process-wide): 0 = every generateLevel is its own launch; 1 (default) = levels whose source
 * has at most 2^20 texels are fused, larger ones stream level by level; 2 = fuse everything.  Returns the previous setting. */
int b2_set_mip_fusion(int enabled);
/* experiment knob (process-wide): 1 = a chain of two or more levels that starts on a large level (more than 2^20 source
 * texels; unbanded objects) runs as ONE persistent launch in which the levels advance together down the image
 * (csrc/mip_wave.cu: DRAM traffic at the algorithmic minimum, but measured slower than the per-level launches on the
 * B200); 0 (default) = never.  Same bytes either way.  Returns the previous setting. */
int b2_set_mip_wave(int enabled);
/* host only (no GPU needed; test hook): the work list of that launch for the chain b2mip_generate_levels(first_level, n,
 * qualities, update_rect_previous_level) would run on a w x h pyramid.  Writes up to cap_items items of 8 ints each
 * {step | kind << 16, block_row, block_col_begin, block_col_end, dep_lo, dep_hi, dep_target, done_counter} in launch
 * order, the number of control words to *n_ctl, and returns the number of items (negative: B2_ERR_*). */
int b2_mip_wave_plan(int color, int w, int h, int first_level, int n, const int* qualities, b2_box2i update_rect_previous_level,
                     int* items8, int cap_items, int* n_ctl);
/* the three samplers, evaluated on the device for `n` query points (test hook for mipmap.d:149-496) */
int b2mip_sample(b2mip*, int kind /*0 linear,1 cubic,2 linearMipmap*/, int n, const float* level, const float* x,
                 const float* y, float* out4);
int b2mip_sync(b2mip*);

/* ---- dirty-rectangle accumulator — DirtyRectList, gui/dplug/gui/boxlist.d:200-303 (host-side, thread-safe like the
 * reference's: addRect may be called from any thread, the draw thread pulls).  haveNoOverlap: boxlist.d:171-186. */
typedef struct b2_dirty_rects b2_dirty_rects;
int b2_dirty_rects_create(b2_dirty_rects** out);                                   /* makeDirtyRectList */
void b2_dirty_rects_destroy(b2_dirty_rects*);
int b2_dirty_rects_add(b2_dirty_rects*, b2_box2i rect);                            /* addRect (rect must be sorted) */
int b2_dirty_rects_is_empty(b2_dirty_rects*);                                      /* isEmpty */
int b2_dirty_rects_pull_all(b2_dirty_rects*, b2_box2i* out, int cap);              /* pullAllRectangles: returns the count;
                                                                                    * the list is emptied only if cap suffices */
int b2_have_no_overlap(const b2_box2i* boxes, int n);

/* rectangle propagation — Mipmap.impactOnNextLevel, mipmap.d:623-645 (host-side, no device work) */
b2_box2i b2_impact_on_next_level(int quality, b2_box2i area, int current_level_w, int current_level_h);
/* tileAreas — gui/dplug/gui/boxlist.d:139-167; returns the number of tiles (writes at most `cap`) */
int b2_tile_areas(const b2_box2i* areas, int n, int max_w, int max_h, b2_box2i* out, int cap);
/* removeOverlappingAreas — boxlist.d:54-118; may reorder `boxes`; returns the number of disjoint boxes */
int b2_remove_overlapping_areas(b2_box2i* boxes, int n, b2_box2i* out, int cap);
/* convertPBRLayerRectToRawLayerRect — gui/dplug/gui/graphics.d:981-1002 */
b2_box2i b2_grow_rect(b2_box2i rect, int margin, int width, int height);

/* ------------------------------------------------------------------------------------------
 * b2pbr — device PBRCompositor : MultipassCompositor : ICompositor
 *         (gui/dplug/gui/compositor.d:29-69,81-194; gui/dplug/gui/legacypbr.d:67-266)
 * ------------------------------------------------------------------------------------------ */
typedef struct b2pbr b2pbr;

/* run-time settings of the 8 passes: the legacy setters (legacypbr.d:75-123), pass members
 * (legacypbr.d:391,453,626-629,676-679,843,1045-1049) and CompositorPass.setActive (compositor.d:222-225) */
typedef struct {
    float light1_color[3];   /* PassObliqueShadowLight.color   */
    float light2_dir[3];     /* PassDirectionalLight.direction */
    float light2_color[3];   /* PassDirectionalLight.color     */
    float light3_dir[3];     /* PassSpecularLight.direction    */
    float light3_color[3];   /* PassSpecularLight.color        */
    float ambient_light;     /* PassAmbientOcclusion.amount    */
    float skybox_amount;     /* PassSkyboxReflections.amount   */
    float tonemap_threshold; /* PassClampAndConvertTo8bit.tonemapThreshold */
    float tonemap_ratio;     /* PassClampAndConvertTo8bit.tonemapRatio     */
    uint32_t pass_active_mask; /* bit p = pass p active */
} b2pbr_params;

/* PBRCompositor defaults (legacypbr.d:391,453,626-629,676-679,843,1045-1049) */
void b2pbr_default_params(b2pbr_params*);

/* PBRCompositor.this(CompositorCreationContext*) — legacypbr.d:141-165.  The thread pool of the
 * creation context has no device counterpart: tiles are scheduled over the GPU's SMs. */
int b2pbr_create(b2pbr** out, int device);
/* band variant: this object composites rows [y0,y1) of a global_w x global_h canvas; see b2mip_create_band */
int b2pbr_create_band(b2pbr** out, int device, int y0, int y1);
void b2pbr_destroy(b2pbr*);
int b2pbr_set_stream(b2pbr*, void* cuda_stream);
/* ICompositor.resizeBuffers(width, height, areaMaxWidth, areaMaxHeight) — compositor.d:43-46; legacypbr.d:180-198.
 * Also (re)allocates the device twins GUIGraphics.doResize sizes next (graphics.d:1115-1137): diffuse
 * Mipmap(5), depth Mipmap(4), material Mipmap(0) and the composited RGBA8 buffer. */
int b2pbr_resize(b2pbr*, int width, int height, int area_max_width, int area_max_height);
int b2pbr_set_params(b2pbr*, const b2pbr_params*);
int b2pbr_get_params(const b2pbr*, b2pbr_params*);
/* selects the lighting kernel compositeTile launches (B2_MODE_*; default B2_MODE_FAST) */
int b2pbr_set_mode(b2pbr*, int mode);
/* PassSkyboxReflections.setSkybox(OwnedImage!RGBA) — legacypbr.d:861-870: copies the host image to the
 * device, builds Mipmap!RGBA(12, image) and generateMipmaps(Quality.box).  rgba == NULL removes the skybox.
 * (The reference takes ownership of the image; here the caller keeps it.) */
int b2pbr_set_skybox(b2pbr*, const uint8_t* rgba, int w, int h, size_t pitch);
/* device twins owned by the compositor object */
b2mip* b2pbr_map(b2pbr*, int which /* B2_MAP_* */);
b2mip* b2pbr_skybox(b2pbr*);
/* device ImageRef of the composited buffer (wfb twin, gapless RGBA8, r in the low byte) */
int b2pbr_output(b2pbr*, b2_imageref* out);

/* GUIGraphics.regenerateMipmaps() — gui/dplug/gui/graphics.d:1354-1423, for the given
 * _rectsToUpdateDisjointedPBR: diffuse level 1 boxAlphaCovIntoPremul, levels 2..5 cubic; depth border
 * (replicateBordersTouching == clamp-to-edge addressing on the device), depth levels 1,2 box, 3,4 cubic. */
int b2pbr_regenerate_mipmaps(b2pbr*, const b2_box2i* rects, int n);
/* the same, enqueued on `cuda_stream` instead of the compositor's stream (no synchronisation; order it with events) */
int b2pbr_regenerate_mipmaps_on(b2pbr*, const b2_box2i* rects, int n, void* cuda_stream);

/* ICompositor.compositeTile(wfb, areas, diffuseMap, materialMap, depthMap, profiler) —
 * compositor.d:63-68, legacypbr.d:201-260 — on device-resident maps; writes the composited twin.
 * `areas` are disjoint rectangles no larger than the area_max_* given to resize (gui tiles, boxlist.d:139-167).
 * Passing NULL maps uses the compositor's own twins. */
int b2pbr_composite_tile(b2pbr*, const b2_box2i* areas, int n, b2mip* diffuse, b2mip* material, b2mip* depth);

/* One frame on device-resident maps = GUIGraphics.regenerateMipmaps() followed by compositeGUI() (what doDraw does back
 * to back, gui/dplug/gui/graphics.d:812-821): b2pbr_regenerate_mipmaps(update_rects) + b2pbr_composite_tile(areas) as ONE
 * call.  Same pyramid levels and the same frame as the two calls. */
int b2pbr_redraw(b2pbr*, const b2_box2i* update_rects, int n_update, const b2_box2i* areas, int n_areas);

/* Same call with HOST images, exactly the reference's arguments: uploads `areas` (grown by the sampling
 * reach where needed: see DESIGN.md) of the three level-0 maps, regenerates the device mip levels the
 * passes sample, composites, and copies `areas` of the result into wfb.  update_rects/n_update are
 * GUIGraphics._rectsToUpdateDisjointedPBR (what changed in level 0); pass NULL/0 to treat `areas` as changed. */
int b2pbr_composite_tile_host(b2pbr*, b2_imageref wfb, const b2_box2i* areas, int n_areas,
                              const b2_box2i* update_rects, int n_update,
                              b2_imageref diffuse_l0, b2_imageref material_l0, b2_imageref depth_l0);
/* The same call split in two for callers that drive SEVERAL compositors (a batch of plug-in instance UIs): _begin
 * enqueues the whole call (uploads, device mips, tiles, copy-back) and returns; _end waits until `wfb` holds the frame.
 * The images must stay valid and unmodified until _end; asynchronous only for page-locked images (pageable memory makes
 * the copies synchronous).  One call in flight per compositor. */
int b2pbr_composite_tile_host_begin(b2pbr*, b2_imageref wfb, const b2_box2i* areas, int n_areas, const b2_box2i* update_rects,
                                    int n_update, b2_imageref diffuse_l0, b2_imageref material_l0, b2_imageref depth_l0);
int b2pbr_composite_tile_host_end(b2pbr*);

/* device -> host copy of rectangles of the composited buffer; synchronises */
int b2pbr_download(b2pbr*, void* host, size_t host_pitch, const b2_box2i* rects, int n);
/* fused post-composite step (graphics.d:1425-1452,1527-1654 reorderComponents): same download but
 * swizzled to the window's pixel order with alpha forced to 255 */
int b2pbr_download_swizzled(b2pbr*, int pixel_format, void* host, size_t host_pitch, const b2_box2i* rects, int n);
/* Row ranges [lo, hi) of one pyramid level: `stored_*` = rows this object holds, `need_*` = rows it must hold
 * VALID data for before compositing its band (for level 0 of a banded object these are the rows to fill by
 * upload + halo exchange with the neighbouring bands).  Unbanded: the whole level. */
/* the same "need" rows as pure host arithmetic (no device, no object): used to plan the halo exchange */
int b2_band_need_rows(int W, int H, int y0, int y1, int map, int level, int* need_lo, int* need_hi);
int b2pbr_band_rows(const b2pbr*, int map, int level, int* stored_lo, int* stored_hi, int* need_lo, int* need_hi);
/* ------------------------------------------------------------------------------------------
 * b2band — multi-GPU row bands of one canvas (BASELINE config 5): the per-frame exchange of the
 * level-0 halo rows between neighbouring bands and the overlapped frame schedule.  The caller of
 * GUIGraphics.compositeGUI (gui/dplug/gui/graphics.d:1338-1349) drives one band object (b2pbr_create_band) per GPU.
 * ------------------------------------------------------------------------------------------ */
typedef struct b2band b2band;
enum { B2_BAND_TRANSPORT_NCCL = 0, B2_BAND_TRANSPORT_PEERS = 1 };
/* rows [y0, y1) of band `rank` of `nranks`: equal shares of the `align`-row GUI tiles (graphics.d:59-60: 64) */
int b2_band_rows(int H, int nranks, int rank, int align, int* y0, int* y1);
/* planning (host arithmetic only): the level-0 halo rectangles of band [y0, y1) (0..2: above / below) and the rows
 * [interior_lo, interior_hi) of whole tiles that can be composited before the halo rows have arrived */
int b2_band_plan(int W, int H, int y0, int y1, int align, b2_box2i halo_rects[2], int* n_halo, int* interior_lo, int* interior_hi);
/* NCCL transport, one process per GPU.  libnccl.so.2 is loaded at run time (dlopen; B2_NCCL_LIB overrides the name).
 * b2_band_nccl_unique_id = ncclGetUniqueId: rank 0 calls it and hands the 128 bytes to the other ranks over any host
 * channel.  b2band_create_nccl is collective (ncclCommInitRank) unless `nccl_comm` — an ncclComm_t the caller owns,
 * whose ranks are the band indices — is given.  band_y[nranks + 1] = the band boundaries every rank agrees on. */
int b2_band_nccl_unique_id(void* id128);
int b2band_create_nccl(b2band** out, b2pbr* band, int rank, int nranks, const int* band_y, const void* id128, void* nccl_comm);
/* PEERS transport, all bands in this process (bands[r] = band r, adjacent in rank order, any mix of devices): halo rows
 * are pushed GPU to GPU with cudaMemcpyPeerAsync on the copy engines; peer access is enabled where the devices differ. */
int b2band_create_peers(b2band** out, b2pbr* const* bands, int nranks);
void b2band_destroy(b2band*);
/* Starts this frame's halo exchange on the bands' exchange streams, ordered after everything enqueued so far on each
 * band's stream (its level-0 uploads / widget draws); returns without waiting.  Only halo rows are written. */
int b2band_exchange_begin(b2band*);
/* Makes each band's stream wait for the exchange (no host synchronisation).  Level 0 of a band may be modified again
 * only after this (the neighbours read the band's own rows). */
int b2band_exchange_end(b2band*);
/* One frame of every local band = GUIGraphics.regenerateMipmaps + compositeGUI over the band's rows, with the exchange
 * hidden behind the work that does not depend on it: exchange_begin; mips of the own rows; interior tiles; on a second
 * stream, once the rows are there, the mips that depend on halo rows; boundary tiles.  Bit-identical to the unbanded
 * frame.  Asynchronous: enqueue-only, results are in b2pbr_output / b2pbr_download of each band. */
int b2band_composite_frame(b2band*);
int b2band_sync(b2band*);
uint64_t b2band_halo_bytes_per_frame(const b2band*);     /* level-0 bytes the local bands receive per frame */
int b2band_info(const b2band*, int local_index, int* rank, int* interior_lo, int* interior_hi, int* n_sends, int* n_recvs);

/* CUDA graph of one frame's device work: between _begin and _end, calls on this object (regenerate_mipmaps,
 * composite_tile on device maps, composite_raw) are recorded instead of executed; _launch replays them on the
 * object's stream.  Valid while sizes, rectangle lists and parameters stay the same (pixel contents may change).
 * The area list must have been composited once before it is recorded. */
int b2pbr_graph_begin(b2pbr*);
int b2pbr_graph_end(b2pbr*, int* graph_id);
int b2pbr_graph_launch(b2pbr*, int graph_id);
/* SURVEY §8(f) rows 2 and 3 — data-parallel level-0 producers.  b2layer = device-resident cache of one widget's PBR
 * render: the buffers UIBufferedElementPBR keeps (gui/dplug/gui/bufferedelement.d:69-116: diffuse RGBA, depth L16,
 * material RGBA + one L8 opacity mask each) or the resized background maps of PBRBackgroundGUI
 * (pbrwidgets/dplug/pbrwidgets/pbrbackgroundgui.d:122-145, no masks; the stb_image_resize2 resampling that fills
 * them stays on the host). */
typedef struct b2layer b2layer;
enum { B2_LAYER_DIFFUSE = 0, B2_LAYER_DEPTH = 1, B2_LAYER_MATERIAL = 2,
       B2_LAYER_DIFFUSE_OPACITY = 3, B2_LAYER_DEPTH_OPACITY = 4, B2_LAYER_MATERIAL_OPACITY = 5 };
enum { B2_LAYER_BLIT = 0, B2_LAYER_BLEND = 1 };
int b2layer_create(b2layer** out, int device, int w, int h, int with_opacity);
void b2layer_destroy(b2layer*);
/* host -> device for rectangles of one plane (B2_LAYER_*), when the widget re-rendered its cache */
int b2layer_upload(b2layer*, int plane, const void* host, size_t host_pitch, const b2_box2i* rects, int n);
/* the widget's onDrawPBR for `dirty_rects` (widget coordinates), widget placed at (x, y) of the UI: writes level 0 of
 * the compositor's three maps.  B2_LAYER_BLIT = PBRBackgroundGUI.onDrawPBR (pbrbackgroundgui.d:147-162);
 * B2_LAYER_BLEND = UIBufferedElementPBR.onDrawPBR (bufferedelement.d:118-131; blendWithAlpha draw.d:933-969;
 * blendColor color.d:453-526, bit-exact incl. the 32-bit wrap of the L16 products). */
int b2pbr_draw_layer(b2pbr*, b2layer*, int x, int y, const b2_box2i* dirty_rects, int n, int mode);

/* SURVEY §8(f) row 2, second half — ImageResizer (graphics/dplug/graphics/resizer.d): the resampling that fills a
 * background widget's cache from its full-size images (PBRBackgroundGUI.resizeFun, pbrbackgroundgui.d:172-199).
 * Semantics = resizer.d over the IN-TREE stb_image_resize port (graphics/dplug/graphics/stb_image_resize.d, i.e. the
 * `DPLUG_USE_STB_IMAGE_RESIZE_V2 = false` branches), bit-exact; the default build's un-vendored stb_image_resize2 rounds
 * differently.  kind selects pixel type and filter:
 *   B2_RESIZE_DIFFUSE      resizeImageDiffuse       RGBA8, Magic Kernel Sharp 2013 at 86 %   (resizer.d:346-372)
 *   B2_RESIZE_MATERIAL     resizeImageMaterial = resizeImageGeneric(RGBA): Catmull-Rom up / Mitchell down (resizer.d:39-64,374-376)
 *   B2_RESIZE_DEPTH        resizeImageDepth(L16)    Magic Kernel Sharp 2021                  (resizer.d:171-197)
 *   B2_RESIZE_GENERIC_L16  resizeImageGeneric(L16)  default filters                          (resizer.d:144-169)
 *   B2_RESIZE_SMOOTHER     resizeImageSmoother      RGBA8, cubic B-spline                    (resizer.d:92-117)
 * Equal sizes copy (sameSizeResize, resizer.d:453-465).  Both calls synchronise.
 *  b2_resize_image: host image -> host image through the device `device`.
 *  b2layer_resize_plane: host image -> plane 0 (diffuse), 1 (depth) or 2 (material) of a device-resident layer at the
 *      layer's size — the widget cache never exists on the host. */
enum { B2_RESIZE_DIFFUSE = 0, B2_RESIZE_MATERIAL = 1, B2_RESIZE_DEPTH = 2, B2_RESIZE_GENERIC_L16 = 3, B2_RESIZE_SMOOTHER = 4 };
int b2_resize_image(int device, int kind, const void* in, int in_w, int in_h, size_t in_pitch, void* out, int out_w, int out_h,
                    size_t out_pitch);
int b2layer_resize_plane(b2layer*, int plane, int kind, const void* in, int in_w, int in_h, size_t in_pitch);
/* host only (no GPU needed; test hook): the gather table the two resampling kernels run for one axis of a resize of
 * `kind` from in_size to out_size samples.  Output sample o is the sum, in this order and from 0,
 * of source[clamp(src[e])] * coef[e] for e in [start[o], start[o + 1]).  Writes out_size + 1 values to `start` and up to
 * `cap` entries to `src` / `coef` (any may be NULL); returns the number of entries (negative: -B2_ERR_*). */
int b2_resize_axis_table(int kind, int in_size, int out_size, int* start, int* src, float* coef, int cap);

/* SURVEY §8(f) row 1 — the post-composite per-pixel chain of GUIGraphics.doDraw, stages E-H (graphics.d:825-868),
 * fused into one pass over the composited twin: blit composited -> rendered, UIColorCorrection transfer tables
 * (pbrwidgets/dplug/pbrwidgets/colorcorrection.d:128-170), reorderComponents(pf) with alpha = 255
 * (graphics.d:1425-1452) and resizeContent into the window buffer (graphics.d:1465-1511).
 * `rects` = _rectsToDisplayDisjointed in user coordinates; the user area sits at (user_x, user_y) in the window;
 * redraw_borders = _redrawBlackBordersAndResizedArea (fill black, copy the whole user area).  Valid when no other
 * Raw-layer widget draws between the blit and the re-ordering.  Synchronises. */
int b2pbr_present(b2pbr*, int pixel_format, void* window, size_t window_pitch, int window_w, int window_h,
                  int user_x, int user_y, const b2_box2i* rects, int n, int redraw_borders);
/* SURVEY §8(f) row 4 — the screenshot encoders (gui/dplug/gui/screencap.d:21-136), taken where GUIGraphics.doDraw takes
 * them (stage G.bis, graphics.d:848-857): on the rendered buffer, i.e. the composited frame after the colour-correction
 * tables and reorderComponents(pixel_format).
 *  _qb: encodeScreenshotAsQB (screencap.d:21-84) — a .qb voxel file of W x 26 x H voxels built on the device from the
 *       rendered pixels and level 0 of the depth map; byte-identical to the reference's file.  `cap` must be at least
 *       b2pbr_screenshot_qb_size().
 *  _rgba: the pixel pass of encodeScreenshotAsPNG (screencap.d:88-131): the rendered buffer brought back to RGBA.
 *  _png: those pixels as a PNG file (8-bit RGBA; the reference hands them to the un-vendored `gamut` encoder — PNG bytes
 *       are not canonical, the decoded image is the reference's).  `cap` >= b2pbr_screenshot_png_size().
 * All three synchronise. */
size_t b2pbr_screenshot_qb_size(const b2pbr*);
int b2pbr_screenshot_qb(b2pbr*, int pixel_format, void* out, size_t cap);
int b2pbr_screenshot_rgba(b2pbr*, int pixel_format, void* out_rgba, size_t pitch);
size_t b2pbr_screenshot_png_size(const b2pbr*);
int b2pbr_screenshot_png(b2pbr*, int pixel_format, void* out, size_t cap, size_t* written);
/* 3 x 256 transfer tables (red, green, blue) of UIColorCorrection (colorcorrection.d:26-41); NULL = none */
int b2pbr_set_color_correction(b2pbr*, const uint8_t* rgb_table_768);
/* setLiftGammaGainContrastRGB (colorcorrection.d:69-116) as a host helper: v = {lift, gamma, gain, contrast} for
 * r, g, b.  (Table arithmetic uses pow; a caller that needs the reference's exact bytes passes its own table.) */
void b2_lift_gamma_gain_contrast_table(const float v[12], uint8_t out768[768]);
int b2pbr_sync(b2pbr*);

/* SimpleRawCompositor (compositor.d:260-298): copy diffuse RGB, alpha = 255, into the composited twin */
int b2pbr_composite_raw(b2pbr*, const b2_box2i* areas, int n);

/* IProfiler stand-in (core/dplug/core/profiler.d:29-64): when enabled, each stage is bracketed with CUDA
 * events; b2pbr_trace_json writes Chrome Trace Event JSON like TraceProfiler (profiler.d:177-207). */
int b2pbr_profile_enable(b2pbr*, int on);
int b2pbr_trace_json(b2pbr*, char* buf, size_t cap, size_t* needed);
/* number of kernels this object launched since creation (bench.py's gpu_launches) */
uint64_t b2pbr_kernel_launches(const b2pbr*);
uint64_t b2mip_kernel_launches(const b2mip*);

/* pin / unpin caller memory so the host-image entry points copy at full PCIe rate */
int b2_host_register(void* ptr, size_t bytes);
int b2_host_unregister(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* B2PBR_H */
