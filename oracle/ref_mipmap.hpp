// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.hpp header).
//
// CPU restatement of graphics/dplug/graphics/mipmap.d: pyramid sizing, dirty
// rectangle propagation, the three down-sampling qualities and the three
// samplers, for RGBA8 and L16 (RGBA16 is "untested and unused" in the
// reference, mipmap.d:800,1115, and is not restated).  Follows the
// version(LDC) / non-legacyPBREmissive branches.
#pragma once
#include "ref_core.hpp"

namespace b2ref {

enum Quality { Q_BOX = 0, Q_CUBIC = 1, Q_BOX_ALPHACOV_PREMUL = 2 };  // mipmap.d:36-46

template <typename COLOR> struct SampleOf;
template <> struct SampleOf<RGBA> { typedef vec4f type; };
template <> struct SampleOf<L16>  { typedef float type; };

// mipmap.d:177-180, 304-307
static const float kLevelFactors[14] = {1.0f, 0.5f, 0.25f, 0.125f, 0.0625f, 0.03125f, 0.015625f, 0.0078125f,
                                        0.00390625f, 0.001953125f, 0.0009765625f, 0.00048828125f,
                                        0.000244140625f, 0.0001220703125f};

inline vec4f operator*(vec4f a, float s) { return vec4f{a.x * s, a.y * s, a.z * s, a.w * s}; }
inline vec4f operator*(float s, vec4f a) { return vec4f{s * a.x, s * a.y, s * a.z, s * a.w}; }
inline vec4f operator+(vec4f a, vec4f b) { return vec4f{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline vec4f operator-(vec4f a, vec4f b) { return vec4f{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }

// mipmap.d:222-229 / 263-270 — Catmull-Rom; D precedence kept: t*t*t == (t*t)*t.
template <typename V>
inline V cubicInterp(float t, V x0, V x1, V x2, V x3) {
    return x1
         + t * ((-0.5f * x0) + (0.5f * x2))
         + t * t * (x0 - (2.5f * x1) + (2.0f * x2) - (0.5f * x3))
         + t * t * t * ((-0.5f * x0) + (1.5f * x1) - (1.5f * x2) + 0.5f * x3);
}

void generateLevelBoxRGBA(OwnedImage<RGBA>& dst, OwnedImage<RGBA>& src, box2i r);       // mipmap.d:676-762
void generateLevelBoxL16(OwnedImage<L16>& dst, OwnedImage<L16>& src, box2i r);          // mipmap.d:764-794
void generateLevelBoxAlphaCovIntoPremulRGBA(OwnedImage<RGBA>& dst, OwnedImage<RGBA>& src, box2i r);  // mipmap.d:823-900
void generateLevelCubicRGBA(OwnedImage<RGBA>& dst, OwnedImage<RGBA>& src, box2i r);     // mipmap.d:902-1040
void generateLevelCubicL16(OwnedImage<L16>& dst, OwnedImage<L16>& src, box2i r);        // mipmap.d:1042-1109

template <typename COLOR>
class Mipmap {
public:
    typedef typename SampleOf<COLOR>::type Sample;
    std::vector<OwnedImage<COLOR>*> levels;  // mipmap.d:48

    Mipmap() {}
    Mipmap(int maxLevel, int w, int h) { size(maxLevel, w, h); }
    ~Mipmap() { for (auto* l : levels) delete l; }

    // mipmap.d:80-126
    void size(int maxLevel, int w, int h) {
        int needed = 0;
        {
            int wr = w, hr = h;
            for (; needed <= maxLevel; ++needed) {
                if (wr == 0 || hr == 0) break;
                wr >>= 1; hr >>= 1;
            }
        }
        while ((int)levels.size() < needed) levels.push_back(new OwnedImage<COLOR>());
        for (int level = 0; level < needed; ++level) {
            levels[level]->size(w, h);
            w >>= 1; h >>= 1;
        }
    }

    int width() const { return levels[0]->w; }
    int height() const { return levels[0]->h; }
    int numLevels() const { return (int)levels.size(); }

    // mipmap.d:623-645
    static box2i impactOnNextLevel(int quality, box2i area, int curW, int curH) {
        box2i maxArea{0, 0, curW / 2, curH / 2};
        if (quality == Q_CUBIC) {
            box2i b{(area.minx - 1) / 2, (area.miny - 1) / 2, (area.maxx + 2) / 2, (area.maxy + 2) / 2};
            return b.intersection(maxArea);
        }
        box2i b{area.minx / 2, area.miny / 2, (area.maxx + 1) / 2, (area.maxy + 1) / 2};
        return b.intersection(maxArea);
    }

    // mipmap.d:534-540
    box2i generateNextLevel(int quality, box2i updateRectPrev, int level) {
        OwnedImage<COLOR>* prev = levels[level - 1];
        box2i r = impactOnNextLevel(quality, updateRectPrev, prev->w, prev->h);
        generateLevel(level, quality, r);
        return r;
    }

    // mipmap.d:544-618
    void generateLevel(int level, int quality, box2i r);

    // mipmap.d:517-528
    void generateMipmaps(int quality) {
        box2i r{0, 0, width(), height()};
        for (int level = 1; level < numLevels(); ++level) {
            if (level >= 3 && quality == Q_BOX) quality = Q_CUBIC;
            r = generateNextLevel(quality, r, level);
        }
    }

    // mipmap.d:293-496
    Sample linearSample(int level, float x, float y) const;
    // mipmap.d:167-287
    Sample cubicSample(int level, float x, float y) const;

    // mipmap.d:149-160
    Sample linearMipmapSample(float level, float x, float y) const {
        int ilevel = (int)level;
        float flevel = level - (float)ilevel;
        Sample a = linearSample(ilevel, x, y);
        if (flevel == 0) return a;
        Sample b = linearSample(ilevel + 1, x, y);
        return a * (1 - flevel) + b * flevel;
    }
};

template <> inline void Mipmap<RGBA>::generateLevel(int level, int quality, box2i r) {
    OwnedImage<RGBA>& dst = *levels[level];
    OwnedImage<RGBA>& src = *levels[level - 1];
    if (quality == Q_BOX) generateLevelBoxRGBA(dst, src, r);
    else if (quality == Q_BOX_ALPHACOV_PREMUL) generateLevelBoxAlphaCovIntoPremulRGBA(dst, src, r);
    else generateLevelCubicRGBA(dst, src, r);
}
template <> inline void Mipmap<L16>::generateLevel(int level, int quality, box2i r) {
    OwnedImage<L16>& dst = *levels[level];
    OwnedImage<L16>& src = *levels[level - 1];
    if (quality == Q_BOX) generateLevelBoxL16(dst, src, r);
    else if (quality == Q_CUBIC) generateLevelCubicL16(dst, src, r);
    else abort();  // mipmap.d:594-595: assert(false) for non-RGBA
}

// ---- linearSample ----------------------------------------------------------
namespace detail {
struct BilinearSetup { int ix, iy, ixp1, iyp1; float fx, fy, fxm1, fym1; };
inline BilinearSetup bilinearSetup(int level, float x, float y, int w, int h) {
    float divider = kLevelFactors[level];
    x = x * divider - 0.5f;
    y = y * divider - 0.5f;
    if (x < 0) x = 0;
    if (y < 0) y = 0;
    __m128i t = _mm_cvttps_epi32(_mm_setr_ps(x, y, 0, 0));  // mipmap.d:318-321
    BilinearSetup s;
    s.ix = _mm_extract_epi32(t, 0);
    s.iy = _mm_extract_epi32(t, 1);
    s.fx = x - (float)s.ix;   // fractional part taken BEFORE the upper clamp (mipmap.d:324-332)
    s.fy = y - (float)s.iy;
    const int maxX = w - 1, maxY = h - 1;
    if (s.ix > maxX) s.ix = maxX;
    if (s.iy > maxY) s.iy = maxY;
    s.ixp1 = s.ix + 1; s.iyp1 = s.iy + 1;
    if (s.ixp1 > maxX) s.ixp1 = maxX;
    if (s.iyp1 > maxY) s.iyp1 = maxY;
    s.fxm1 = 1 - s.fx;
    s.fym1 = 1 - s.fy;
    return s;
}
inline int clampLevel(int level, int numLevels) {
    if (level < 0) level = 0;
    if (level >= numLevels) level = numLevels - 1;
    return level;
}
}  // namespace detail

template <> inline vec4f Mipmap<RGBA>::linearSample(int level, float x, float y) const {
    level = detail::clampLevel(level, numLevels());
    const OwnedImage<RGBA>& img = *levels[level];
    detail::BilinearSetup s = detail::bilinearSetup(level, x, y, img.w, img.h);
    const RGBA* L0 = img.scanlinePtr(s.iy);
    const RGBA* L1 = img.scanlinePtr(s.iyp1);
    int32_t Ai, Bi, Ci, Di;
    memcpy(&Ai, &L0[s.ix], 4); memcpy(&Bi, &L0[s.ixp1], 4);
    memcpy(&Ci, &L1[s.ix], 4); memcpy(&Di, &L1[s.ixp1], 4);
    // version(LDC) branch, mipmap.d:355-385
    __m128i z = _mm_setzero_si128();
    __m128i abcd = _mm_setr_epi32(Ai, Bi, Ci, Di);
    __m128i ab = _mm_unpacklo_epi8(abcd, z), cd = _mm_unpackhi_epi8(abcd, z);
    __m128 vA = _mm_cvtepi32_ps(_mm_unpacklo_epi16(ab, z));
    __m128 vB = _mm_cvtepi32_ps(_mm_unpackhi_epi16(ab, z));
    __m128 vC = _mm_cvtepi32_ps(_mm_unpacklo_epi16(cd, z));
    __m128 vD = _mm_cvtepi32_ps(_mm_unpackhi_epi16(cd, z));
    __m128 vfx = _mm_set1_ps(s.fx), vfxm1 = _mm_set1_ps(s.fxm1);
    __m128 up = _mm_add_ps(_mm_mul_ps(vA, vfxm1), _mm_mul_ps(vB, vfx));
    __m128 down = _mm_add_ps(_mm_mul_ps(vC, vfxm1), _mm_mul_ps(vD, vfx));
    __m128 res = _mm_add_ps(_mm_mul_ps(up, _mm_set1_ps(s.fym1)), _mm_mul_ps(down, _mm_set1_ps(s.fy)));
    vec4f out;
    _mm_storeu_ps(&out.x, res);
    return out;
}

template <> inline float Mipmap<L16>::linearSample(int level, float x, float y) const {
    level = detail::clampLevel(level, numLevels());
    const OwnedImage<L16>& img = *levels[level];
    detail::BilinearSetup s = detail::bilinearSetup(level, x, y, img.w, img.h);
    const L16* L0 = img.scanlinePtr(s.iy);
    const L16* L1 = img.scanlinePtr(s.iyp1);
    // mipmap.d:478-483
    float up = (float)L0[s.ix].l * s.fxm1 + (float)L0[s.ixp1].l * s.fx;
    float down = (float)L1[s.ix].l * s.fxm1 + (float)L1[s.ixp1].l * s.fx;
    return up * s.fym1 + down * s.fy;
}

// ---- cubicSample -----------------------------------------------------------
namespace detail {
struct BicubicSetup { int xi[4], yi[4]; float a, b; };
inline BicubicSetup bicubicSetup(int level, float x, float y, int w, int h) {
    float divider = kLevelFactors[level];
    x = x * divider - 0.5f;
    y = y * divider - 0.5f;
    __m128 off = _mm_setr_ps(-1, 0, 1, 2);                       // mipmap.d:186-193
    __m128i xi = _mm_cvttps_epi32(_mm_add_ps(_mm_set1_ps(x), off));
    __m128i yi = _mm_cvttps_epi32(_mm_add_ps(_mm_set1_ps(y), off));
    __m128i zero = _mm_setzero_si128();
    xi = _mm_min_epi32(_mm_max_epi32(xi, zero), _mm_set1_epi32(w - 1));
    yi = _mm_min_epi32(_mm_max_epi32(yi, zero), _mm_set1_epi32(h - 1));
    BicubicSetup s;
    _mm_storeu_si128((__m128i*)s.xi, xi);
    _mm_storeu_si128((__m128i*)s.yi, yi);
    float a = x + 1.0f, b = y + 1.0f;                            // mipmap.d:201-204
    s.a = a - (float)(int)a;
    s.b = b - (float)(int)b;
    return s;
}
}  // namespace detail

template <> inline float Mipmap<L16>::cubicSample(int level, float x, float y) const {
    level = detail::clampLevel(level, numLevels());
    const OwnedImage<L16>& img = *levels[level];
    detail::BicubicSetup s = detail::bicubicSetup(level, x, y, img.w, img.h);
    float R[4];
    for (int row = 0; row < 4; ++row) {
        const L16* p = img.scanlinePtr(s.yi[row]);
        R[row] = cubicInterp<float>(s.a, (float)p[s.xi[0]].l, (float)p[s.xi[1]].l, (float)p[s.xi[2]].l, (float)p[s.xi[3]].l);
    }
    float r = cubicInterp<float>(s.b, R[0], R[1], R[2], R[3]);
    if (r < 0) r = 0;                                            // mipmap.d:216-221
    if (r > 65535) r = 65535;
    return r;
}

template <> inline vec4f Mipmap<RGBA>::cubicSample(int level, float x, float y) const {
    level = detail::clampLevel(level, numLevels());
    const OwnedImage<RGBA>& img = *levels[level];
    detail::BicubicSetup s = detail::bicubicSetup(level, x, y, img.w, img.h);
    vec4f R[4];
    for (int row = 0; row < 4; ++row) {
        const RGBA* p = img.scanlinePtr(s.yi[row]);
        vec4f P[4];
        for (int k = 0; k < 4; ++k) {
            RGBA c = p[s.xi[k]];
            P[k] = vec4f{(float)c.r, (float)c.g, (float)c.b, (float)c.a};
        }
        R[row] = cubicInterp<vec4f>(s.a, P[0], P[1], P[2], P[3]);
    }
    vec4f r = cubicInterp<vec4f>(s.b, R[0], R[1], R[2], R[3]);
    float* c = &r.x;                                             // mipmap.d:250-261
    for (int k = 0; k < 4; ++k) if (c[k] < 0) c[k] = 0;
    for (int k = 0; k < 4; ++k) if (c[k] > 65535) c[k] = 65535;
    return r;
}

}  // namespace b2ref
