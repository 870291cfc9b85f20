/* ORACLE — TEST INFRASTRUCTURE ONLY.  C ABI of the CPU restatement (libb2ref.so).
 *
 * Mirrors include/b2pbr.h symbol for symbol with the prefix b2ref_ so the parity
 * tests can drive both back ends with the same calls.  Host memory only.
 * Only tests/, bench.py (cpu_baseline / --impl reference) and
 * __graft_entry__.smoke() may load this library.
 */
#ifndef B2REF_H
#define B2REF_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int min_x, min_y, max_x, max_y; } b2ref_box2i;       /* math/dplug/math/box.d:29-30 */
typedef struct { int w, h; size_t pitch; void* pixels; } b2ref_imageref; /* graphics/dplug/graphics/image.d:127-131 */

enum { B2REF_COLOR_RGBA8 = 0, B2REF_COLOR_L16 = 1 };
enum { B2REF_QUALITY_BOX = 0, B2REF_QUALITY_CUBIC = 1, B2REF_QUALITY_BOX_ALPHACOV_PREMUL = 2 }; /* mipmap.d:36-46 */

typedef struct {
    float light1_color[3], light2_dir[3], light2_color[3], light3_dir[3], light3_color[3];
    float ambient_light, skybox_amount, tonemap_threshold, tonemap_ratio;
    uint32_t pass_active_mask;
} b2ref_params;

typedef struct b2ref_mip b2ref_mip;
typedef struct b2ref_pbr b2ref_pbr;
typedef struct b2ref_gfx b2ref_gfx;

/* ---- Mipmap!T (graphics/dplug/graphics/mipmap.d) */
b2ref_mip* b2ref_mip_create(int color, int max_level, int w, int h);
void b2ref_mip_destroy(b2ref_mip*);
int b2ref_mip_num_levels(b2ref_mip*);
int b2ref_mip_level(b2ref_mip*, int level, b2ref_imageref* out);
/* OwnedImage.size on one level (graphics.d:1121-1126 re-sizes depth level 0 with a border) */
int b2ref_mip_size_level(b2ref_mip*, int level, int w, int h, int border, int row_align, int x_mult, int trailing);
int b2ref_mip_replicate_borders_touching(b2ref_mip*, int level, b2ref_box2i rect);
int b2ref_mip_generate_next_level(b2ref_mip*, int quality, b2ref_box2i prev_rect, int level, b2ref_box2i* out);
int b2ref_mip_generate_mipmaps(b2ref_mip*, int quality);
int b2ref_mip_linear_sample(b2ref_mip*, int level, float x, float y, float out[4]);
int b2ref_mip_cubic_sample(b2ref_mip*, int level, float x, float y, float out[4]);
int b2ref_mip_linear_mipmap_sample(b2ref_mip*, float level, float x, float y, float out[4]);

/* ---- PBRCompositor (gui/dplug/gui/legacypbr.d) */
void b2ref_default_params(b2ref_params*);
b2ref_pbr* b2ref_pbr_create(int num_threads);
void b2ref_pbr_destroy(b2ref_pbr*);
int b2ref_pbr_resize(b2ref_pbr*, int w, int h, int tile_w, int tile_h);
int b2ref_pbr_set_params(b2ref_pbr*, const b2ref_params*);
int b2ref_pbr_set_skybox(b2ref_pbr*, const uint8_t* rgba, int w, int h, size_t pitch);
int b2ref_pbr_composite_tile(b2ref_pbr*, b2ref_imageref wfb, const b2ref_box2i* areas, int n,
                             b2ref_mip* diffuse, b2ref_mip* material, b2ref_mip* depth);

/* ---- GUIGraphics slice (gui/dplug/gui/graphics.d:1115-1137,1338-1423) */
b2ref_gfx* b2ref_gfx_create(int num_threads);
void b2ref_gfx_destroy(b2ref_gfx*);
int b2ref_gfx_resize(b2ref_gfx*, int w, int h);
b2ref_pbr* b2ref_gfx_compositor(b2ref_gfx*);
b2ref_mip* b2ref_gfx_map(b2ref_gfx*, int which /*0 diffuse, 1 material, 2 depth*/);
int b2ref_gfx_output(b2ref_gfx*, b2ref_imageref* out);
int b2ref_gfx_regenerate_mipmaps(b2ref_gfx*, const b2ref_box2i* rects, int n);
int b2ref_gfx_composite_gui(b2ref_gfx*, const b2ref_box2i* rects, int n);

/* ---- post-composite chain, GUIGraphics.doDraw stages E-H (graphics.d:825-868); `rendered` plays _renderedBuffer */
int b2ref_present(b2ref_imageref composited, b2ref_imageref rendered, int pixel_format, b2ref_imageref window, int user_x, int user_y,
                  const b2ref_box2i* rects, int n, const uint8_t* table768, int redraw_borders);

/* ---- screenshot encoders (gui/dplug/gui/screencap.d:21-136) on the rendered buffer of stage G.bis (graphics.d:848-857) */
size_t b2ref_screenshot_qb(b2ref_imageref rendered, int pixel_format, b2ref_imageref depth, uint8_t* out, size_t cap);
int b2ref_screenshot_png_pixels(b2ref_imageref rendered, int pixel_format, uint8_t* out_rgba, size_t out_pitch);

/* ---- ImageResizer over the in-tree stb_image_resize port (graphics/dplug/graphics/resizer.d, stb_image_resize.d; the
 * `DPLUG_USE_STB_IMAGE_RESIZE_V2 = false` branches).  kind: 0 diffuse, 1 material / generic RGBA, 2 depth, 3 generic L16,
 * 4 smoother RGBA */
int b2ref_resize_image(int kind, const void* in, int in_w, int in_h, size_t in_pitch, void* out, int out_w, int out_h, size_t out_pitch);

/* ---- widget cache -> level-0 maps: blit (pbrbackgroundgui.d:147-162) or blendWithAlpha (bufferedelement.d:118-131) */
int b2ref_draw_layer(b2ref_imageref dstD, b2ref_imageref dstZ, b2ref_imageref dstM, b2ref_imageref srcD, b2ref_imageref srcZ,
                     b2ref_imageref srcM, b2ref_imageref opD, b2ref_imageref opZ, b2ref_imageref opM, int x, int y,
                     const b2ref_box2i* rects, int n, int mode);

/* ---- rectangle lists (gui/dplug/gui/boxlist.d) */
int b2ref_tile_areas(const b2ref_box2i* areas, int n, int max_w, int max_h, b2ref_box2i* out, int cap);
int b2ref_remove_overlapping_areas(b2ref_box2i* boxes, int n, b2ref_box2i* out, int cap);
b2ref_box2i b2ref_bounding_box(const b2ref_box2i* boxes, int n);
int b2ref_have_no_overlap(const b2ref_box2i* boxes, int n);
/* DirtyRectList.addRect (boxlist.d:236-292) on a caller-held list of *n rectangles (capacity cap); returns the new count */
int b2ref_dirty_rect_list_add(b2ref_box2i* list, int n, int cap, b2ref_box2i rect);
b2ref_box2i b2ref_impact_on_next_level(int quality, b2ref_box2i area, int w, int h);

/* ---- scalar helpers, exposed for unit tests */
double b2ref_linmap(double v, double a, double b, double c, double d);
float b2ref_rsqrtss(float x, int mode /*0 table, 1 host instruction, 2 exact*/);
void b2ref_set_rsqrt_mode(int mode);
long b2ref_rsqrtss_table_vs_host(void);   /* exhaustive sweep; number of mismatching operands */
float b2ref_pow_ps(float base, float exponent);
float b2ref_fastlog2(float x);
void b2ref_plane_fitting_normal(const float depth9[9], float out3[3]);
/* OwnedImage.size layout rule incl. the stale-width quirk (image.d:354-407) */
void b2ref_image_layout(int elem_size, int stale_w, int width, int height, int border, int row_align, int x_mult,
                        int trailing, int* pitch, int* border_right, int* rows);
/* the reference's replicateBorders known-answer test (image.d:673-713): fills out[6*7] */
int b2ref_kat_replicate_borders(uint8_t out[42], int* border_left, int* border_right, int* pitch);
int b2ref_compute_right_padding(int width, int border, int x_mult);
size_t b2ref_next_multiple_of(size_t base, size_t multiple);

#ifdef __cplusplus
}
#endif
#endif
