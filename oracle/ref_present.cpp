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
