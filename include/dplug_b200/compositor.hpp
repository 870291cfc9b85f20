// Header-only C++ mirror of the reference interface for this path, above the C ABI (include/b2pbr.h).
//
// Same names and argument meaning as Dplug's D classes: Mipmap!T (graphics/dplug/graphics/mipmap.d:30-646),
// ICompositor / PBRCompositor (gui/dplug/gui/compositor.d:29-69, gui/dplug/gui/legacypbr.d:67-266), tileAreas
// (gui/dplug/gui/boxlist.d:139-167).  The reference's methods are nothrow and return void; here a failing call
// throws b2::Error (there is no CPU fallback to continue with).
#pragma once
#include <stdexcept>
#include <string>
#include <vector>
#include "../b2pbr.h"

namespace b2 {

struct Error : std::runtime_error {
    int code;
    Error(int c) : std::runtime_error(std::string("b2pbr: ") + b2_last_error()), code(c) {}
};
inline void check(int rc) { if (rc != B2_OK) throw Error(rc); }

using box2i = b2_box2i;
template <typename T> struct ImageRef { int w, h; size_t pitch; T* pixels; };  // image.d:127-131
struct RGBA { unsigned char r, g, b, a; };
struct L16 { unsigned short l; };
template <typename T> struct ColorOf;
template <> struct ColorOf<RGBA> { static constexpr int value = B2_COLOR_RGBA8; };
template <> struct ColorOf<L16> { static constexpr int value = B2_COLOR_L16; };

inline std::vector<box2i> tileAreas(const std::vector<box2i>& areas, int maxWidth, int maxHeight) {
    int n = b2_tile_areas(areas.data(), (int)areas.size(), maxWidth, maxHeight, nullptr, 0);
    std::vector<box2i> out((size_t)n);
    b2_tile_areas(areas.data(), (int)areas.size(), maxWidth, maxHeight, out.data(), n);
    return out;
}

inline bool haveNoOverlap(const std::vector<box2i>& areas) { return b2_have_no_overlap(areas.data(), (int)areas.size()) != 0; }  // boxlist.d:171-186

// DirtyRectList, gui/dplug/gui/boxlist.d:200-303
class DirtyRectList {
public:
    DirtyRectList() { check(b2_dirty_rects_create(&_h)); }
    ~DirtyRectList() { b2_dirty_rects_destroy(_h); }
    DirtyRectList(const DirtyRectList&) = delete;
    DirtyRectList& operator=(const DirtyRectList&) = delete;
    bool isEmpty() { return b2_dirty_rects_is_empty(_h) != 0; }
    void addRect(box2i rect) { check(b2_dirty_rects_add(_h, rect)); }
    void pullAllRectangles(std::vector<box2i>& result) {   // appends, like the reference (boxlist.d:224-234)
        int n = b2_dirty_rects_pull_all(_h, nullptr, 0);
        size_t at = result.size();
        result.resize(at + (size_t)n);
        b2_dirty_rects_pull_all(_h, result.data() + at, n);
    }

private:
    b2_dirty_rects* _h = nullptr;
};

template <typename COLOR>
class Mipmap {
public:
    enum class Quality { box = B2_QUALITY_BOX, cubic = B2_QUALITY_CUBIC, boxAlphaCovIntoPremul = B2_QUALITY_BOX_ALPHACOV_PREMUL };

    Mipmap(int maxLevel, int w, int h, int device = 0) { check(b2mip_create(&_h, device, ColorOf<COLOR>::value, maxLevel, w, h)); }
    explicit Mipmap(b2mip* borrowed) : _h(borrowed), _owned(false) {}
    ~Mipmap() { if (_owned) b2mip_destroy(_h); }
    Mipmap(const Mipmap&) = delete;
    Mipmap& operator=(const Mipmap&) = delete;

    int numLevels() const { return b2mip_num_levels(_h); }
    int width() const { return b2mip_width(_h); }
    int height() const { return b2mip_height(_h); }
    b2mip* handle() const { return _h; }

    box2i generateNextLevel(Quality q, box2i updateRectPreviousLevel, int level) {
        box2i out;
        check(b2mip_generate_next_level(_h, (int)q, updateRectPreviousLevel, level, &out));
        return out;
    }
    void generateLevel(int level, Quality q, box2i updateRect) { check(b2mip_generate_level(_h, level, (int)q, updateRect)); }
    void generateMipmaps(Quality q) { check(b2mip_generate_mipmaps(_h, (int)q)); }
    // levels firstLevel .. firstLevel + qualities.size() - 1, the rectangle chained through impactOnNextLevel: what
    // regenerateMipmaps does per dirty rectangle (graphics.d:1378-1419), in fused launches
    std::vector<box2i> generateLevels(int firstLevel, const std::vector<Quality>& qualities, box2i updateRectPreviousLevel) {
        std::vector<int> q(qualities.size());
        for (size_t i = 0; i < q.size(); ++i) q[i] = (int)qualities[i];
        std::vector<box2i> out(q.size());
        check(b2mip_generate_levels(_h, firstLevel, (int)q.size(), q.data(), updateRectPreviousLevel, out.data()));
        return out;
    }

    void upload(int level, ImageRef<const COLOR> host, const std::vector<box2i>& rects) {
        check(b2mip_upload(_h, level, host.pixels, host.pitch, rects.data(), (int)rects.size()));
    }
    void download(int level, ImageRef<COLOR> host, const std::vector<box2i>& rects) {
        check(b2mip_download(_h, level, host.pixels, host.pitch, rects.data(), (int)rects.size()));
    }

private:
    b2mip* _h = nullptr;
    bool _owned = true;
};

struct ICompositor {
    virtual ~ICompositor() {}
    virtual void resizeBuffers(int width, int height, int areaMaxWidth, int areaMaxHeight) = 0;
    virtual void compositeTile(ImageRef<RGBA> wfb, const std::vector<box2i>& areas, ImageRef<RGBA> diffuseLevel0,
                               ImageRef<RGBA> materialLevel0, ImageRef<L16> depthLevel0) = 0;
};

class PBRCompositor : public ICompositor {
public:
    explicit PBRCompositor(int device = 0) { check(b2pbr_create(&_h, device)); b2pbr_default_params(&_p); }
    ~PBRCompositor() override { b2pbr_destroy(_h); }
    PBRCompositor(const PBRCompositor&) = delete;
    PBRCompositor& operator=(const PBRCompositor&) = delete;

    void resizeBuffers(int width, int height, int areaMaxWidth, int areaMaxHeight) override {
        check(b2pbr_resize(_h, width, height, areaMaxWidth, areaMaxHeight));
    }
    void compositeTile(ImageRef<RGBA> wfb, const std::vector<box2i>& areas, ImageRef<RGBA> d, ImageRef<RGBA> m, ImageRef<L16> z) override {
        check(b2pbr_composite_tile_host(_h, b2_imageref{wfb.w, wfb.h, wfb.pitch, wfb.pixels}, areas.data(), (int)areas.size(), nullptr, 0,
                                        b2_imageref{d.w, d.h, d.pitch, d.pixels}, b2_imageref{m.w, m.h, m.pitch, m.pixels},
                                        b2_imageref{z.w, z.h, z.pitch, z.pixels}));
    }
    // device-resident variant: maps already on the device (b2pbr_map)
    void compositeTile(const std::vector<box2i>& areas) { check(b2pbr_composite_tile(_h, areas.data(), (int)areas.size(), nullptr, nullptr, nullptr)); }
    void regenerateMipmaps(const std::vector<box2i>& rects) { check(b2pbr_regenerate_mipmaps(_h, rects.data(), (int)rects.size())); }

    // legacy setters, legacypbr.d:75-123
    void light1Color(float r, float g, float b) { set3(_p.light1_color, r, g, b); }
    void light2Dir(float x, float y, float z) { set3(_p.light2_dir, x, y, z); }
    void light2Color(float r, float g, float b) { set3(_p.light2_color, r, g, b); }
    void light3Dir(float x, float y, float z) { set3(_p.light3_dir, x, y, z); }
    void light3Color(float r, float g, float b) { set3(_p.light3_color, r, g, b); }
    void skyboxAmount(float a) { _p.skybox_amount = a; push(); }
    void ambientLight(float a) { _p.ambient_light = a; push(); }
    void tonemapThreshold(float v) { _p.tonemap_threshold = v; push(); }
    void tonemapRatio(float v) { _p.tonemap_ratio = v; push(); }
    void setPassActive(int pass, bool active) {
        if (active) _p.pass_active_mask |= (1u << pass); else _p.pass_active_mask &= ~(1u << pass);
        push();
    }
    void setSkybox(ImageRef<const RGBA> image) { check(b2pbr_set_skybox(_h, (const uint8_t*)image.pixels, image.w, image.h, image.pitch)); }
    b2pbr* handle() const { return _h; }

    // screenshot encoders, gui/dplug/gui/screencap.d:21-136, on the rendered buffer of doDraw stage G.bis (graphics.d:848-857)
    std::vector<unsigned char> encodeScreenshotAsQB(int pixelFormat) {
        std::vector<unsigned char> out(b2pbr_screenshot_qb_size(_h));
        check(b2pbr_screenshot_qb(_h, pixelFormat, out.data(), out.size()));
        return out;
    }
    std::vector<unsigned char> encodeScreenshotAsPNG(int pixelFormat) {
        std::vector<unsigned char> out(b2pbr_screenshot_png_size(_h));
        size_t written = 0;
        check(b2pbr_screenshot_png(_h, pixelFormat, out.data(), out.size(), &written));
        out.resize(written);
        return out;
    }

private:
    b2pbr* _h = nullptr;
    b2pbr_params _p;
    void push() { check(b2pbr_set_params(_h, &_p)); }
    void set3(float* dst, float a, float b, float c) { dst[0] = a; dst[1] = b; dst[2] = c; push(); }
};

// ImageResizer, graphics/dplug/graphics/resizer.d (in-tree stb_image_resize port semantics): resampling on the device
class ImageResizer {
public:
    explicit ImageResizer(int device = 0) : _device(device) {}
    void resizeImageDiffuse(ImageRef<const RGBA> in, ImageRef<RGBA> out) { run(B2_RESIZE_DIFFUSE, in.pixels, in.w, in.h, in.pitch, out.pixels, out.w, out.h, out.pitch); }
    void resizeImageMaterial(ImageRef<const RGBA> in, ImageRef<RGBA> out) { run(B2_RESIZE_MATERIAL, in.pixels, in.w, in.h, in.pitch, out.pixels, out.w, out.h, out.pitch); }
    void resizeImageGeneric(ImageRef<const RGBA> in, ImageRef<RGBA> out) { resizeImageMaterial(in, out); }
    void resizeImageSmoother(ImageRef<const RGBA> in, ImageRef<RGBA> out) { run(B2_RESIZE_SMOOTHER, in.pixels, in.w, in.h, in.pitch, out.pixels, out.w, out.h, out.pitch); }
    void resizeImageDepth(ImageRef<const L16> in, ImageRef<L16> out) { run(B2_RESIZE_DEPTH, in.pixels, in.w, in.h, in.pitch, out.pixels, out.w, out.h, out.pitch); }
    void resizeImageGeneric(ImageRef<const L16> in, ImageRef<L16> out) { run(B2_RESIZE_GENERIC_L16, in.pixels, in.w, in.h, in.pitch, out.pixels, out.w, out.h, out.pitch); }

private:
    int _device;
    void run(int kind, const void* in, int iw, int ih, size_t ip, void* out, int ow, int oh, size_t op) {
        check(b2_resize_image(_device, kind, in, iw, ih, ip, out, ow, oh, op));
    }
};

}  // namespace b2
