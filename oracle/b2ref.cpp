// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.hpp header).  C ABI glue for libb2ref.so.
#include "b2ref.h"
#include "ref_pbr.hpp"

using namespace b2ref;

struct b2ref_mip {
    int color;
    Mipmap<RGBA>* rgba = nullptr;
    Mipmap<L16>* l16 = nullptr;
    bool owned = true;
};
struct b2ref_pbr {
    PBRCompositor* c;
    bool owned;
};
struct b2ref_gfx {
    RefGraphics* g;
    b2ref_pbr pbr;
    b2ref_mip maps[3];
};

static inline box2i toBox(b2ref_box2i b) { return box2i{b.min_x, b.min_y, b.max_x, b.max_y}; }
static inline b2ref_box2i fromBox(box2i b) { return b2ref_box2i{b.minx, b.miny, b.maxx, b.maxy}; }

extern "C" {

b2ref_mip* b2ref_mip_create(int color, int max_level, int w, int h) {
    b2ref_mip* m = new b2ref_mip();
    m->color = color;
    if (color == B2REF_COLOR_RGBA8) m->rgba = new Mipmap<RGBA>(max_level, w, h);
    else if (color == B2REF_COLOR_L16) m->l16 = new Mipmap<L16>(max_level, w, h);
    else { delete m; return nullptr; }
    return m;
}
void b2ref_mip_destroy(b2ref_mip* m) {
    if (!m) return;
    if (m->owned) { delete m->rgba; delete m->l16; }
    delete m;
}
int b2ref_mip_num_levels(b2ref_mip* m) { return m->rgba ? m->rgba->numLevels() : m->l16->numLevels(); }
int b2ref_mip_level(b2ref_mip* m, int level, b2ref_imageref* out) {
    if (level < 0 || level >= b2ref_mip_num_levels(m)) return 1;
    if (m->rgba) {
        auto* im = m->rgba->levels[level];
        *out = b2ref_imageref{im->w, im->h, (size_t)im->pitchInBytes(), im->scanlinePtr(0)};
    } else {
        auto* im = m->l16->levels[level];
        *out = b2ref_imageref{im->w, im->h, (size_t)im->pitchInBytes(), im->scanlinePtr(0)};
    }
    return 0;
}
int b2ref_mip_size_level(b2ref_mip* m, int level, int w, int h, int border, int row_align, int x_mult, int trailing) {
    if (level < 0 || level >= b2ref_mip_num_levels(m)) return 1;
    if (m->rgba) m->rgba->levels[level]->size(w, h, border, row_align, x_mult, trailing);
    else m->l16->levels[level]->size(w, h, border, row_align, x_mult, trailing);
    return 0;
}
int b2ref_mip_replicate_borders_touching(b2ref_mip* m, int level, b2ref_box2i rect) {
    if (level < 0 || level >= b2ref_mip_num_levels(m)) return 1;
    if (m->rgba) m->rgba->levels[level]->replicateBordersTouching(toBox(rect));
    else m->l16->levels[level]->replicateBordersTouching(toBox(rect));
    return 0;
}
int b2ref_mip_generate_next_level(b2ref_mip* m, int quality, b2ref_box2i prev_rect, int level, b2ref_box2i* out) {
    if (level < 1 || level >= b2ref_mip_num_levels(m)) return 1;
    box2i r = m->rgba ? m->rgba->generateNextLevel(quality, toBox(prev_rect), level)
                      : m->l16->generateNextLevel(quality, toBox(prev_rect), level);
    if (out) *out = fromBox(r);
    return 0;
}
int b2ref_mip_generate_mipmaps(b2ref_mip* m, int quality) {
    if (m->rgba) m->rgba->generateMipmaps(quality); else m->l16->generateMipmaps(quality);
    return 0;
}
int b2ref_mip_linear_sample(b2ref_mip* m, int level, float x, float y, float out[4]) {
    if (m->rgba) { vec4f v = m->rgba->linearSample(level, x, y); out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; }
    else { out[0] = m->l16->linearSample(level, x, y); out[1] = out[2] = out[3] = 0; }
    return 0;
}
int b2ref_mip_cubic_sample(b2ref_mip* m, int level, float x, float y, float out[4]) {
    if (m->rgba) { vec4f v = m->rgba->cubicSample(level, x, y); out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; }
    else { out[0] = m->l16->cubicSample(level, x, y); out[1] = out[2] = out[3] = 0; }
    return 0;
}
int b2ref_mip_linear_mipmap_sample(b2ref_mip* m, float level, float x, float y, float out[4]) {
    if (m->rgba) { vec4f v = m->rgba->linearMipmapSample(level, x, y); out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; }
    else { out[0] = m->l16->linearMipmapSample(level, x, y); out[1] = out[2] = out[3] = 0; }
    return 0;
}

static void toParams(const b2ref_params* in, PBRParams* p) {
    memcpy(p->light1Color, in->light1_color, 12);
    memcpy(p->light2Dir, in->light2_dir, 12);
    memcpy(p->light2Color, in->light2_color, 12);
    memcpy(p->light3Dir, in->light3_dir, 12);
    memcpy(p->light3Color, in->light3_color, 12);
    p->ambientLight = in->ambient_light;
    p->skyboxAmount = in->skybox_amount;
    p->tonemapThreshold = in->tonemap_threshold;
    p->tonemapRatio = in->tonemap_ratio;
    p->passActiveMask = in->pass_active_mask;
}
void b2ref_default_params(b2ref_params* out) {
    PBRParams p;
    defaultParams(&p);
    memcpy(out->light1_color, p.light1Color, 12);
    memcpy(out->light2_dir, p.light2Dir, 12);
    memcpy(out->light2_color, p.light2Color, 12);
    memcpy(out->light3_dir, p.light3Dir, 12);
    memcpy(out->light3_color, p.light3Color, 12);
    out->ambient_light = p.ambientLight;
    out->skybox_amount = p.skyboxAmount;
    out->tonemap_threshold = p.tonemapThreshold;
    out->tonemap_ratio = p.tonemapRatio;
    out->pass_active_mask = p.passActiveMask;
}
b2ref_pbr* b2ref_pbr_create(int num_threads) { return new b2ref_pbr{new PBRCompositor(num_threads), true}; }
void b2ref_pbr_destroy(b2ref_pbr* p) {
    if (!p) return;
    if (p->owned) { delete p->c; delete p; }
}
int b2ref_pbr_resize(b2ref_pbr* p, int w, int h, int tw, int th) { p->c->resizeBuffers(w, h, tw, th); return 0; }
int b2ref_pbr_set_params(b2ref_pbr* p, const b2ref_params* in) { toParams(in, &p->c->params); return 0; }
int b2ref_pbr_set_skybox(b2ref_pbr* p, const uint8_t* rgba, int w, int h, size_t pitch) {
    p->c->setSkybox((const RGBA*)rgba, w, h, pitch);
    return 0;
}
int b2ref_pbr_composite_tile(b2ref_pbr* p, b2ref_imageref wfb, const b2ref_box2i* areas, int n, b2ref_mip* diffuse,
                             b2ref_mip* material, b2ref_mip* depth) {
    if (!diffuse->rgba || !material->rgba || !depth->l16) return 1;
    ImageRef<RGBA> ref{wfb.w, wfb.h, wfb.pitch, (RGBA*)wfb.pixels};
    p->c->compositeTile(ref, (const box2i*)areas, n, *diffuse->rgba, *material->rgba, *depth->l16);
    return 0;
}

b2ref_gfx* b2ref_gfx_create(int num_threads) {
    b2ref_gfx* g = new b2ref_gfx();
    g->g = new RefGraphics(num_threads);
    g->pbr = b2ref_pbr{&g->g->compositor, false};
    g->maps[0].color = B2REF_COLOR_RGBA8; g->maps[0].rgba = &g->g->diffuseMap; g->maps[0].owned = false;
    g->maps[1].color = B2REF_COLOR_RGBA8; g->maps[1].rgba = &g->g->materialMap; g->maps[1].owned = false;
    g->maps[2].color = B2REF_COLOR_L16; g->maps[2].l16 = &g->g->depthMap; g->maps[2].owned = false;
    return g;
}
void b2ref_gfx_destroy(b2ref_gfx* g) { if (g) { delete g->g; delete g; } }
int b2ref_gfx_resize(b2ref_gfx* g, int w, int h) { g->g->resize(w, h); return 0; }
b2ref_pbr* b2ref_gfx_compositor(b2ref_gfx* g) { return &g->pbr; }
b2ref_mip* b2ref_gfx_map(b2ref_gfx* g, int which) { return (which >= 0 && which < 3) ? &g->maps[which] : nullptr; }
int b2ref_gfx_output(b2ref_gfx* g, b2ref_imageref* out) {
    auto& b = g->g->compositedBuffer;
    *out = b2ref_imageref{b.w, b.h, (size_t)b.pitchInBytes(), b.scanlinePtr(0)};
    return 0;
}
int b2ref_gfx_regenerate_mipmaps(b2ref_gfx* g, const b2ref_box2i* rects, int n) {
    g->g->regenerateMipmaps((const box2i*)rects, n);
    return 0;
}
int b2ref_gfx_composite_gui(b2ref_gfx* g, const b2ref_box2i* rects, int n) {
    g->g->compositeGUI((const box2i*)rects, n);
    return 0;
}

int b2ref_tile_areas(const b2ref_box2i* areas, int n, int max_w, int max_h, b2ref_box2i* out, int cap) {
    std::vector<box2i> v;
    tileAreas((const box2i*)areas, n, max_w, max_h, v);
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = fromBox(v[i]);
    return (int)v.size();
}
int b2ref_remove_overlapping_areas(b2ref_box2i* boxes, int n, b2ref_box2i* out, int cap) {
    std::vector<box2i> in((const box2i*)boxes, (const box2i*)boxes + n), f;
    removeOverlappingAreas(in, f);
    for (int i = 0; i < (int)f.size() && i < cap; ++i) out[i] = fromBox(f[i]);
    return (int)f.size();
}
b2ref_box2i b2ref_bounding_box(const b2ref_box2i* boxes, int n) {
    std::vector<box2i> in((const box2i*)boxes, (const box2i*)boxes + n);
    return fromBox(boundingBox(in));
}
int b2ref_dirty_rect_list_add(b2ref_box2i* list, int n, int cap, b2ref_box2i rect) {
    std::vector<box2i> v((const box2i*)list, (const box2i*)list + n);
    dirtyRectListAdd(v, toBox(rect));
    for (int i = 0; i < (int)v.size() && i < cap; ++i) list[i] = fromBox(v[i]);
    return (int)v.size();
}
int b2ref_have_no_overlap(const b2ref_box2i* boxes, int n) { return haveNoOverlap((const box2i*)boxes, n) ? 1 : 0; }
b2ref_box2i b2ref_impact_on_next_level(int quality, b2ref_box2i area, int w, int h) {
    return fromBox(Mipmap<RGBA>::impactOnNextLevel(quality, toBox(area), w, h));
}

double b2ref_linmap(double v, double a, double b, double c, double d) { return linmap<double>(v, a, b, c, d); }
float b2ref_rsqrtss(float x, int mode) {
    int saved = g_rsqrt_mode;
    g_rsqrt_mode = mode;
    float r = rsqrtss(x);
    g_rsqrt_mode = saved;
    return r;
}
void b2ref_set_rsqrt_mode(int mode) { g_rsqrt_mode = mode; }
long b2ref_rsqrtss_table_vs_host(void) {
    // Every (sign, exponent, 10-bit mantissa bin) with four low-bit patterns: 2 x 256 x 1024 x 4 operands.
    static const uint32_t low[4] = {0u, 0x1fffu, 0x0a5au, 0x1000u};
    long bad = 0;
    for (uint32_t hi = 0; hi < (1u << 19); ++hi)
        for (int k = 0; k < 4; ++k) {
            uint32_t bits = (hi << 13) | low[k];
            float x; memcpy(&x, &bits, 4);
            float t = rsqrtss_table(x);
            float h = _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x)));
            uint32_t tb, hb; memcpy(&tb, &t, 4); memcpy(&hb, &h, 4);
            bool tn = (tb << 1) > 0xff000000u, hn = (hb << 1) > 0xff000000u;
            if (tn && hn) continue;
            if (tb != hb) bad++;
        }
    return bad;
}
float b2ref_pow_ps(float base, float exponent) { return pow_ps(base, exponent); }
float b2ref_fastlog2(float x) { return fastlog2(x); }
void b2ref_plane_fitting_normal(const float d[9], float out[3]) { b2ref::planeFittingNormalPublic(d, out); }

void b2ref_image_layout(int elem_size, int stale_w, int width, int height, int border, int row_align, int x_mult,
                        int trailing, int* pitch, int* border_right, int* rows) {
    int rightPadding = OwnedImage<L8>::computeRightPadding(stale_w, border, x_mult);
    int br = border + rightPadding;
    if (br < trailing) br = trailing;
    int p = elem_size * (border + width + br);
    p = (int)OwnedImage<L8>::nextMultipleOf((size_t)p, (size_t)row_align);
    if (pitch) *pitch = p;
    if (border_right) *border_right = br;
    if (rows) *rows = border + height + border;
}

int b2ref_kat_replicate_borders(uint8_t out[42], int* border_left, int* border_right, int* pitch) {
    // image.d:673-713.  The reference test passes (width, height, border, xMultiplicity, trailing)
    // positionally into size(width, height, border, rowAlignment, xMultiplicity, ...): rowAlignment=1,
    // xMultiplicity=3, trailingSamples=0 on a fresh image.
    OwnedImage<L8> img;
    img.size(2, 2, 2, 1, 3);
    if (img.w != 2 || img.h != 2) return 1;
    *border_left = img.borderLeft();
    *border_right = img.borderRight();
    *pitch = img.pitchInBytes();
    img.fillWith(L8{5});
    img.scanlinePtr(0)[0].l = 1;
    img.scanlinePtr(0)[1].l = 2;
    img.scanlinePtr(1)[0].l = 3;
    img.scanlinePtr(1)[1].l = 4;
    img.replicateBorders();
    for (int y = -2; y < 4; ++y)
        for (int x = -2; x < 5; ++x) out[(y + 2) * 7 + (x + 2)] = img.unsafeScanlinePtr(y)[x].l;
    return 0;
}
int b2ref_compute_right_padding(int width, int border, int x_mult) { return OwnedImage<L8>::computeRightPadding(width, border, x_mult); }
size_t b2ref_next_multiple_of(size_t base, size_t multiple) { return OwnedImage<L8>::nextMultipleOf(base, multiple); }

}  // extern "C"
