// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.hpp header).
//
// CPU restatement of the post-composite per-pixel chain of GUIGraphics.doDraw, stages E-H
// (gui/dplug/gui/graphics.d:825-868): blit composited -> rendered, UIColorCorrection.onDrawRaw
// (pbrwidgets/dplug/pbrwidgets/colorcorrection.d:128-170), reorderComponents (graphics.d:1425-1452, the
// shuffleComponents* helpers graphics.d:1527-1654) and resizeContent (graphics.d:1465-1511).
#include "b2ref.h"
#include "ref_core.hpp"

using namespace b2ref;

// colorcorrection.d:157-170
static void applyColorCorrection(ImageRef<RGBA> image, const uint8_t* rgbTable) {
    for (int j = 0; j < image.h; ++j) {
        uint8_t* scan = (uint8_t*)image.scanline(j);
        for (int i = 0; i < image.w; ++i) {
            uint8_t r = scan[4 * i], g = scan[4 * i + 1], b = scan[4 * i + 2];
            scan[4 * i] = rgbTable[r];
            scan[4 * i + 1] = rgbTable[g + 256];
            scan[4 * i + 2] = rgbTable[b + 512];
        }
    }
}

// graphics.d:1527-1654: alpha forced to 255, then RGBA -> RGBA / BGRA / ARGB byte order
static void shuffleComponents(ImageRef<RGBA> image, int pf) {
    for (int j = 0; j < image.h; ++j) {
        uint8_t* scan = (uint8_t*)image.scanline(j);
        for (int i = 0; i < image.w; ++i) {
            uint8_t r = scan[4 * i], g = scan[4 * i + 1], b = scan[4 * i + 2];
            if (pf == 1) { scan[4 * i] = b; scan[4 * i + 1] = g; scan[4 * i + 2] = r; scan[4 * i + 3] = 255; }
            else if (pf == 2) { scan[4 * i] = 255; scan[4 * i + 1] = r; scan[4 * i + 2] = g; scan[4 * i + 3] = b; }
            else scan[4 * i + 3] = 255;
        }
    }
}

static ImageRef<RGBA> crop(ImageRef<RGBA> src, box2i r) {  // cropImageRef, image.d:217-228
    return ImageRef<RGBA>{r.width(), r.height(), src.pitch, src.scanline(r.miny) + r.minx};
}

extern "C" int b2ref_present(b2ref_imageref composited, b2ref_imageref rendered, int pf, b2ref_imageref window, int user_x, int user_y,
                             const b2ref_box2i* rects, int n, const uint8_t* table768, int redraw_borders) {
    ImageRef<RGBA> comp{composited.w, composited.h, composited.pitch, (RGBA*)composited.pixels};
    ImageRef<RGBA> rend{rendered.w, rendered.h, rendered.pitch, (RGBA*)rendered.pixels};
    ImageRef<RGBA> win{window.w, window.h, window.pitch, (RGBA*)window.pixels};
    const box2i* R = (const box2i*)rects;
    // E. copy composited -> rendered for every rect that will be displayed (graphics.d:825-835)
    for (int k = 0; k < n; ++k)
        for (int j = R[k].miny; j < R[k].maxy; ++j)
            memcpy(rend.scanline(j) + R[k].minx, comp.scanline(j) + R[k].minx, (size_t)R[k].width() * 4);
    // F. Raw layer: UIColorCorrection.onDrawRaw on the dirty rects (colorcorrection.d:128-134)
    if (table768)
        for (int k = 0; k < n; ++k) applyColorCorrection(crop(rend, R[k]), table768);
    // G. reorderComponents (graphics.d:1425-1452)
    for (int k = 0; k < n; ++k) shuffleComponents(crop(rend, R[k]), pf);
    // H. resizeContent (graphics.d:1465-1511)
    const box2i userArea{user_x, user_y, user_x + rend.w < win.w ? user_x + rend.w : win.w, user_y + rend.h < win.h ? user_y + rend.h : win.h};
    std::vector<box2i> toCopy;
    if (redraw_borders) {
        RGBA black = (pf == 2) ? RGBA{255, 0, 0, 0} : RGBA{0, 0, 0, 255};
        for (int j = 0; j < win.h; ++j)
            for (int i = 0; i < win.w; ++i) win.scanline(j)[i] = black;
        toCopy.push_back(userArea);
    } else {
        for (int k = 0; k < n; ++k) {   // _rectsToResizeDisjointed: translated by _userArea.min, clipped to _userArea (graphics.d:959-964)
            box2i t = R[k].translate(user_x, user_y).intersection(userArea);
            if (!t.empty()) toCopy.push_back(t);
        }
    }
    for (const box2i& r : toCopy)
        for (int j = r.miny; j < r.maxy; ++j)
            memcpy(win.scanline(j) + r.minx, rend.scanline(j - user_y) + (r.minx - user_x), (size_t)r.width() * 4);
    return 0;
}

// ---------------------------------------------------------------- level-0 producers
// blendColor(RGBA) — graphics/dplug/graphics/color.d:453-482 (version(LDC) branch: madd, *32897, >>23)
static RGBA blendColorRGBA(RGBA fg, RGBA bg, uint8_t alpha) {
    uint8_t invAlpha = (uint8_t)(~(int)alpha);
    int32_t fgp, bgp;
    memcpy(&fgp, &fg, 4); memcpy(&bgp, &bg, 4);
    __m128i alphaMask = _mm_set1_epi32((invAlpha << 16) | alpha);
    __m128i zero = _mm_setzero_si128();
    __m128i colorMask = _mm_unpacklo_epi8(_mm_cvtsi32_si128(fgp), _mm_cvtsi32_si128(bgp));
    colorMask = _mm_unpacklo_epi8(colorMask, zero);
    __m128i product = _mm_madd_epi16(colorMask, alphaMask);
    product = _mm_mullo_epi32(product, _mm_set1_epi32(32897));
    product = _mm_srli_epi32(product, 23);
    __m128i c = _mm_packus_epi16(_mm_packs_epi32(product, zero), zero);
    int32_t r = _mm_cvtsi128_si32(c);
    RGBA out;
    memcpy(&out, &r, 4);
    return out;
}
// blendColor(L16) — color.d:522-526: `int` products (they wrap for bright texels), signed division
static L16 blendColorL16(L16 fg, L16 bg, uint16_t alpha) {
    uint16_t inv = (uint16_t)(~(int)alpha);
    int32_t sum = (int32_t)((uint32_t)fg.l * (uint32_t)alpha + (uint32_t)bg.l * (uint32_t)inv);
    return L16{(uint16_t)(sum / 65535)};
}

// UIBufferedElementPBR.onDrawPBR blend loop (bufferedelement.d:118-131) / PBRBackgroundGUI.onDrawPBR blit
// (pbrbackgroundgui.d:147-162).  dst* are the three level-0 maps, src* / op* the widget's cached buffers, the widget
// sits at (x, y); rects are in widget coordinates.  mode 0 = blit, 1 = blendWithAlpha (draw.d:933-969).
extern "C" int b2ref_draw_layer(b2ref_imageref dstD, b2ref_imageref dstZ, b2ref_imageref dstM, b2ref_imageref srcD, b2ref_imageref srcZ,
                                b2ref_imageref srcM, b2ref_imageref opD, b2ref_imageref opZ, b2ref_imageref opM, int x, int y,
                                const b2ref_box2i* rects, int n, int mode) {
    auto rowp = [](const b2ref_imageref& im, int j) { return (uint8_t*)im.pixels + (size_t)j * im.pitch; };
    for (int k = 0; k < n; ++k) {
        const b2ref_box2i r = rects[k];
        for (int j = r.min_y; j < r.max_y; ++j) {
            RGBA* dd = (RGBA*)rowp(dstD, j + y) + x; L16* dz = (L16*)rowp(dstZ, j + y) + x; RGBA* dm = (RGBA*)rowp(dstM, j + y) + x;
            const RGBA* sd = (const RGBA*)rowp(srcD, j); const L16* sz = (const L16*)rowp(srcZ, j); const RGBA* sm = (const RGBA*)rowp(srcM, j);
            if (mode == 0) {
                for (int i = r.min_x; i < r.max_x; ++i) { dd[i] = sd[i]; dz[i] = sz[i]; dm[i] = sm[i]; }
                continue;
            }
            const uint8_t* ad = rowp(opD, j); const uint8_t* az = rowp(opZ, j); const uint8_t* am = rowp(opM, j);
            for (int i = r.min_x; i < r.max_x; ++i) {
                if (ad[i]) dd[i] = blendColorRGBA(sd[i], dd[i], ad[i]);
                if (az[i]) dz[i] = blendColorL16(sz[i], dz[i], (uint16_t)((az[i] << 8) | az[i]));
                if (am[i]) dm[i] = blendColorRGBA(sm[i], dm[i], am[i]);
            }
        }
    }
    return 0;
}

// ---------------------------------------------------------------- screenshot encoders
// encodeScreenshotAsQB — gui/dplug/gui/screencap.d:21-84.  `color` is the rendered buffer (after reorderComponents(pf),
// graphics.d:843-857), read through RGBA's fields: for BGRA8 / ARGB8 windows the "r, g, b" written are simply bytes 0, 1, 2
// of the re-ordered pixel (the reference does not look at `pf`).  Returns the size of the file; writes at most `cap` bytes.
extern "C" size_t b2ref_screenshot_qb(b2ref_imageref color, int /*pf*/, b2ref_imageref depth, uint8_t* out, size_t cap) {
    const int DEPTH = 16, ADD_DEPTH = 10;                  // screencap.d:25-26
    const int W = color.w, H = color.h;
    size_t n = 0;
    auto put8 = [&](uint8_t v) { if (n < cap) out[n] = v; ++n; };
    auto put32 = [&](uint32_t v) { for (int k = 0; k < 4; ++k) put8((uint8_t)(v >> (8 * k))); };   // writeLE
    put8(1); put8(1); put8(0); put8(0);                    // .qb version (screencap.d:33-36)
    put32(0); put32(0); put32(0); put32(0); put32(1);      // RGBA, left handed, uncompressed, visibility mask, one matrix
    put8(1); put8('0');                                    // matrix name
    put32((uint32_t)W); put32((uint32_t)(DEPTH + ADD_DEPTH)); put32((uint32_t)H);
    put32(0); put32(0); put32(0);                          // position
    for (int z = 0; z < H; ++z) {
        const uint16_t* depthScan = (const uint16_t*)((const uint8_t*)depth.pixels + (size_t)z * depth.pitch);
        const uint8_t* colorScan = (const uint8_t*)color.pixels + (size_t)z * color.pitch;
        for (int y = 0; y < DEPTH + ADD_DEPTH; ++y)
            for (int x = 0; x < W; ++x) {
                const int d = ADD_DEPTH + (DEPTH * (int)depthScan[x]) / 65536;
                put8(colorScan[4 * x + 0]); put8(colorScan[4 * x + 1]); put8(colorScan[4 * x + 2]);
                put8(d >= y ? 255 : 0);
            }
    }
    return n;
}

// encodeScreenshotAsPNG, the pixel pass (screencap.d:88-131): a clone of the rendered buffer brought back from the
// window's pixel format to RGBA; the clone is what the PNG encoder (gamut, un-vendored) receives.
extern "C" int b2ref_screenshot_png_pixels(b2ref_imageref color, int pf, uint8_t* out_rgba, size_t out_pitch) {
    for (int y = 0; y < color.h; ++y) {
        uint8_t* scan = out_rgba + (size_t)y * out_pitch;
        memcpy(scan, (const uint8_t*)color.pixels + (size_t)y * color.pitch, (size_t)color.w * 4);
        for (int x = 0; x < color.w; ++x) {
            if (pf == 1) { uint8_t t = scan[4 * x + 0]; scan[4 * x + 0] = scan[4 * x + 2]; scan[4 * x + 2] = t; }   // BGRA8
            else if (pf == 2) {                                                                                   // ARGB8
                const uint8_t a = scan[4 * x + 0], r = scan[4 * x + 1], g = scan[4 * x + 2], b = scan[4 * x + 3];
                scan[4 * x + 0] = r; scan[4 * x + 1] = g; scan[4 * x + 2] = b; scan[4 * x + 3] = a;
            }
        }
    }
    return 0;
}
