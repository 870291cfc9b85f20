// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
//
// CPU restatement of the reference's pixel types, rectangle algebra, image
// layout and the numeric helpers of the PBR path.  Only tests/, bench.py's
// cpu_baseline / --impl reference leg and __graft_entry__.smoke() may use it.
//
// PARITY STATUS: the reference (pure D) cannot be built in this environment.
// The known-answer tests the reference holds for this path (image.d:644-713,
// boxlist.d:120-135,188-192, core/math.d:29-33) are restated in
// tests/test_oracle_kat.py and pin the layout / border / rectangle code below.
// No reference test pins the lighting passes, samplers or generateLevel*:
// for those functions this oracle is "parity unpinned" (see DESIGN.md).
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <emmintrin.h>
#include <smmintrin.h>

namespace b2ref {

// ---------------------------------------------------------------- pixel types
// graphics/dplug/graphics/color.d:307-321
struct L8   { uint8_t l; };
struct L16  { uint16_t l; };
struct RGBA { uint8_t r, g, b, a; };
struct RGBf  { float r, g, b; };
struct RGBAf { float r, g, b, a; };
struct vec4f { float x, y, z, w; };

// ---------------------------------------------------------------- box2i
// math/dplug/math/box.d:21-55 (POD layout {min.x,min.y,max.x,max.y})
struct box2i {
    int minx, miny, maxx, maxy;
    int width() const { return maxx - minx; }
    int height() const { return maxy - miny; }
    // box.d:188-192: empty if ANY dimension has min == max
    bool empty() const { return minx == maxx || miny == maxy; }
    // box.d:476-484
    bool isSorted() const { return minx <= maxx && miny <= maxy; }
    // box.d:237-245
    bool contains(const box2i& o) const {
        if (o.minx < minx || o.maxx > maxx) return false;
        if (o.miny < miny || o.maxy > maxy) return false;
        return true;
    }
    // box.d:298-319 — returns *this / o unchanged when one of them is empty
    box2i intersection(const box2i& o) const {
        if (empty()) return *this;
        if (o.empty()) return o;
        box2i r;
        int lo = minx > o.minx ? minx : o.minx, hi = maxx < o.maxx ? maxx : o.maxx;
        r.minx = lo; r.maxx = hi >= lo ? hi : lo;
        lo = miny > o.miny ? miny : o.miny; hi = maxy < o.maxy ? maxy : o.maxy;
        r.miny = lo; r.maxy = hi >= lo ? hi : lo;
        return r;
    }
    // box.d:324-328
    bool intersects(const box2i& o) const {
        box2i i = intersection(o);
        return i.isSorted() && !i.empty();
    }
    // box.d:453-471
    box2i expand(const box2i& o) const {
        if (empty()) return o;
        if (o.empty()) return *this;
        box2i r;
        r.minx = minx < o.minx ? minx : o.minx;
        r.miny = miny < o.miny ? miny : o.miny;
        r.maxx = maxx > o.maxx ? maxx : o.maxx;
        r.maxy = maxy > o.maxy ? maxy : o.maxy;
        return r;
    }
    box2i translate(int dx, int dy) const { return box2i{minx + dx, miny + dy, maxx + dx, maxy + dy}; }
    bool operator==(const box2i& o) const { return minx == o.minx && miny == o.miny && maxx == o.maxx && maxy == o.maxy; }
};

// ---------------------------------------------------------------- ImageRef
// graphics/dplug/graphics/image.d:127-145
template <typename T>
struct ImageRef {
    int w, h;
    size_t pitch;  // bytes
    T* pixels;
    T* scanline(int y) const { return (T*)((uint8_t*)pixels + (size_t)(uint32_t)y * pitch); }
};

// ---------------------------------------------------------------- OwnedImage
// graphics/dplug/graphics/image.d:236-642
template <typename T>
class OwnedImage {
public:
    int w = 0, h = 0;

    OwnedImage() {}
    OwnedImage(int w_, int h_, int border = 0, int rowAlign = 1, int xMult = 1, int trailing = 0) {
        size(w_, h_, border, rowAlign, xMult, trailing);
    }
    ~OwnedImage() { freeImageData(); }
    OwnedImage(const OwnedImage&) = delete;
    OwnedImage& operator=(const OwnedImage&) = delete;

    // image.d:301-315
    void freeImageData() {
        if (_buffer) {
            free(_buffer);
            w = h = 0; _pixels = nullptr; _bytePitch = 0; _buffer = nullptr; _border = 0; _borderRight = 0;
        }
    }

    // image.d:318-326.  The reference computes the byte offset in 32 bits; the
    // oracle uses size_t (SURVEY Appendix B "32-bit offsets").
    T* scanlinePtr(int y) const { return (T*)((uint8_t*)_pixels + (size_t)_bytePitch * (size_t)y); }
    // image.d:616-622 (negative y reaches the top border)
    T* unsafeScanlinePtr(int y) const { return (T*)((uint8_t*)_pixels + (ptrdiff_t)_bytePitch * (ptrdiff_t)y); }

    // image.d:354-407.  NOTE the reference quirk kept here on purpose: the right
    // padding is computed from the STALE member `w`, not from `width`
    // (image.d:369); it is inert when xMultiplicity == 1.
    void size(int width, int height, int border = 0, int rowAlign = 1, int xMult = 1, int trailing = 0) {
        int rightPadding = computeRightPadding(w, border, xMult);
        int borderRight = border + rightPadding;
        if (borderRight < trailing) borderRight = trailing;
        int actualW = border + width + borderRight;
        int actualH = border + height + border;
        int bytePitch = (int)sizeof(T) * actualW;
        bytePitch = (int)nextMultipleOf((size_t)bytePitch, (size_t)rowAlign);
        w = width; h = height; _border = border; _borderRight = borderRight; _bytePitch = bytePitch;
        size_t alloc = (size_t)bytePitch * (size_t)actualH + (size_t)(rowAlign - 1);
        free(_buffer);
        _buffer = (uint8_t*)malloc(alloc ? alloc : 1);
        size_t off = (size_t)bytePitch * (size_t)border + sizeof(T) * (size_t)border;
        _pixels = (T*)nextMultipleOf((size_t)(_buffer + off), (size_t)rowAlign);
    }

    int pitchInBytes() const { return _bytePitch; }   // image.d:418-421
    int borderLeft() const { return _border; }        // image.d:424-445
    int borderRight() const { return _borderRight; }
    int borderTop() const { return _border; }
    int borderBottom() const { return _border; }
    bool isGapless() const { return (size_t)w * sizeof(T) == (size_t)_bytePitch; }

    ImageRef<T> toRef() const { return ImageRef<T>{w, h, (size_t)_bytePitch, scanlinePtr(0)}; }  // image.d:169-190

    // image.d:451-459
    void fillWith(T fill) {
        for (int y = -borderTop(); y < h + borderBottom(); ++y) {
            int n = borderLeft() + w + borderRight();
            T* scan = unsafeScanlinePtr(y) - borderLeft();
            for (int x = 0; x < n; ++x) scan[x] = fill;
        }
    }

    void replicateBorders() { replicateBordersTouching(box2i{0, 0, w, h}); }  // image.d:487-490

    // image.d:493-596
    void replicateBordersTouching(box2i r) {
        if (w < 1 || h < 1) return;
        const int W = w, H = h;
        const bool touchL = r.minx == 0, touchT = r.miny == 0, touchR = r.maxx == W, touchB = r.maxy == H;
        if (touchT) {
            if (touchL) {
                T c = scanlinePtr(0)[0];
                for (int y = -borderTop(); y < 0; ++y)
                    for (int x = -borderLeft(); x < 0; ++x) unsafeScanlinePtr(y)[x] = c;
            }
            for (int y = -borderTop(); y < 0; ++y)
                for (int x = r.minx; x < r.maxx; ++x) unsafeScanlinePtr(y)[x] = scanlinePtr(0)[x];
            if (touchR) {
                T c = scanlinePtr(0)[W - 1];
                for (int y = -borderTop(); y < 0; ++y)
                    for (int x = 0; x < borderRight(); ++x) unsafeScanlinePtr(y)[W + x] = c;
            }
        }
        if (touchL)
            for (int y = r.miny; y < r.maxy; ++y) {
                T c = scanlinePtr(y)[0];
                for (int x = -borderLeft(); x < 0; ++x) unsafeScanlinePtr(y)[x] = c;
            }
        if (touchR)
            for (int y = r.miny; y < r.maxy; ++y) {
                T c = scanlinePtr(y)[W - 1];
                for (int x = 0; x < borderRight(); ++x) unsafeScanlinePtr(y)[W + x] = c;
            }
        if (touchB) {
            if (touchL) {
                T c = scanlinePtr(H - 1)[0];
                for (int y = H; y < H + borderTop(); ++y)
                    for (int x = -borderLeft(); x < 0; ++x) unsafeScanlinePtr(y)[x] = c;
            }
            for (int y = H; y < H + borderBottom(); ++y)
                for (int x = r.minx; x < r.maxx; ++x) unsafeScanlinePtr(y)[x] = scanlinePtr(H - 1)[x];
            if (touchR) {
                T c = scanlinePtr(H - 1)[W - 1];
                for (int y = H; y < H + borderTop(); ++y)
                    for (int x = 0; x < borderRight(); ++x) unsafeScanlinePtr(y)[W + x] = c;
            }
        }
    }

    // image.d:624-635
    static int computeRightPadding(int width, int border, int xMult) {
        int next = (int)nextMultipleOf((size_t)(width + border), (size_t)xMult);
        return next - (width + border);
    }
    static size_t nextMultipleOf(size_t base, size_t multiple) {
        size_t n = (base + multiple - 1) / multiple;
        return multiple * n;
    }

private:
    T* _pixels = nullptr;
    int _bytePitch = 0;
    uint8_t* _buffer = nullptr;
    int _border = 0;
    int _borderRight = 0;
};

// ---------------------------------------------------------------- numeric helpers

// core/dplug/core/math.d:25-28
template <typename T>
inline T linmap(T value, T a, T b, T c, T d) { return c + (d - c) * (value - a) / (b - a); }

// gui/dplug/gui/legacypbr.d:1117-1133 (Laurent de Soras' log2 approximation)
inline float fastlog2(float val) {
    int32_t x; memcpy(&x, &val, 4);
    int log_2 = ((x >> 23) & 255) - 128;
    x = x & ~(255 << 23);
    x += 127 << 23;
    float f; memcpy(&f, &x, 4);
    return f + (float)log_2;
}

// RSQRTSS (gui/dplug/gui/ransac.d:43).  mode 0: emulate the Intel instruction
// with the 2048-entry table (deterministic on any host); mode 1: execute the
// host's own instruction; mode 2: exact 1/sqrt (intel-intrinsics' non-x86
// fallback).
extern const uint32_t kRsqrtssTable[2048];
extern int g_rsqrt_mode;
inline float rsqrtss_table(float x) {
    uint32_t b; memcpy(&b, &x, 4);
    if ((b << 1) > 0xff000000u) return x + x;                 // NaN in, NaN out
    if (b & 0x80000000u) {                                    // negative: -0 / -denormal -> -inf, else NaN
        uint32_t r = ((b & 0x7f800000u) == 0) ? 0xff800000u : 0xffc00000u;
        float f; memcpy(&f, &r, 4); return f;
    }
    uint32_t e = b >> 23;
    uint32_t r;
    if (e == 0) r = 0x7f800000u;                              // +0 and denormals -> +inf
    else if (e == 255) r = 0;                                 // +inf -> +0
    else {
        int de = ((int)e - ((e & 1) ? 127 : 126)) / 2;
        r = kRsqrtssTable[(b >> 13) & 0x7ff] - ((uint32_t)de << 23);
    }
    float f; memcpy(&f, &r, 4); return f;
}
inline float rsqrtss(float x) {
    if (g_rsqrt_mode == 1) return _mm_cvtss_f32(_mm_rsqrt_ss(_mm_set_ss(x)));
    if (g_rsqrt_mode == 2) return 1.0f / __builtin_sqrtf(x);
    return rsqrtss_table(x);
}

// _mm_pow_ps (legacypbr.d:786-790) lives in the un-vendored dub dependency
// `intel-intrinsics` (dub.json: "~>1.0", unpinned).  Its published algorithm is
// exp_ps(y * log_ps(x)), both being ports of Julien Pommier's Cephes-derived
// sse_mathfun; restated here in scalar form, one IEEE single op per step.
float cephes_logf(float x);
float cephes_expf(float x);
inline float pow_ps(float base, float exponent) { return cephes_expf(exponent * cephes_logf(base)); }

// ---------------------------------------------------------------- rectangle lists (gui/dplug/gui/boxlist.d)
box2i boundingBox(const std::vector<box2i>& boxes);                                        // boxlist.d:16-28
void boxSubtraction(const box2i& A, const box2i& C, box2i& D, box2i& E, box2i& F, box2i& G);  // boxlist.d:42-48
void removeOverlappingAreas(std::vector<box2i>& boxes, std::vector<box2i>& filtered);      // boxlist.d:54-118
// DirtyRectList.addRect, boxlist.d:236-292: the list stays free of overlaps (the mutex of the reference is left out:
// the oracle is single-threaded)
void dirtyRectListAdd(std::vector<box2i>& list, box2i rect);
void tileAreas(const box2i* areas, int n, int maxW, int maxH, std::vector<box2i>& out);    // boxlist.d:139-167
bool haveNoOverlap(const box2i* areas, int n);                                             // boxlist.d:171-186

}  // namespace b2ref
