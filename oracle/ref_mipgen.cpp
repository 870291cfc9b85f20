// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.hpp header).
//
// The five generateLevel* down-samplers of graphics/dplug/graphics/mipmap.d,
// the rectangle-list helpers of gui/dplug/gui/boxlist.d, the Cephes log/exp
// behind _mm_pow_ps and the RSQRTSS table.
#include "ref_mipmap.hpp"

namespace b2ref {

const uint32_t kRsqrtssTable[2048] = {
#include "rsqrtss_table.inc"
};
int g_rsqrt_mode = 0;

// ---------------------------------------------------------------- Cephes log / exp
// Algorithm of sse_mathfun's log_ps / exp_ps (what intel-intrinsics' _mm_log_ps /
// _mm_exp_ps port); each statement is one IEEE binary32 operation.
float cephes_logf(float x) {
    const bool invalid = x <= 0.0f;                 // cmple(x, 0): false for NaN
    uint32_t xb;
    {
        const uint32_t minNormPos = 0x00800000u;
        float mn; memcpy(&mn, &minNormPos, 4);
        // maxps(x, min_norm_pos): returns the second operand if either is NaN
        x = (x > mn) ? x : mn;
    }
    memcpy(&xb, &x, 4);
    int32_t emm0 = (int32_t)(xb >> 23);
    xb = (xb & ~0x7f800000u) | 0x3f000000u;         // keep mantissa, set exponent of 0.5
    memcpy(&x, &xb, 4);
    emm0 -= 0x7f;
    float e = (float)emm0;
    e = e + 1.0f;
    const bool below = x < 0.707106781186547524f;
    float tmp = below ? x : 0.0f;
    x = x - 1.0f;
    e = e - (below ? 1.0f : 0.0f);
    x = x + tmp;
    float z = x * x;
    float y = 7.0376836292E-2f;
    y = y * x; y = y + -1.1514610310E-1f;
    y = y * x; y = y + 1.1676998740E-1f;
    y = y * x; y = y + -1.2420140846E-1f;
    y = y * x; y = y + 1.4249322787E-1f;
    y = y * x; y = y + -1.6668057665E-1f;
    y = y * x; y = y + 2.0000714765E-1f;
    y = y * x; y = y + -2.4999993993E-1f;
    y = y * x; y = y + 3.3333331174E-1f;
    y = y * x;
    y = y * z;
    tmp = e * -2.12194440e-4f;
    y = y + tmp;
    tmp = z * 0.5f;
    y = y - tmp;
    tmp = e * 0.693359375f;
    x = x + y;
    x = x + tmp;
    if (invalid) { uint32_t nanb = 0xffffffffu; memcpy(&x, &nanb, 4); }
    return x;
}

float cephes_expf(float x) {
    // minps / maxps return the second operand when the first is NaN
    x = (x < 88.3762626647949f) ? x : 88.3762626647949f;
    x = (x > -88.3762626647949f) ? x : -88.3762626647949f;
    float fx = x * 1.44269504088896341f;
    fx = fx + 0.5f;
    // floor via truncate-and-correct
    float tmp = (float)(int32_t)fx;
    if (tmp > fx) tmp = tmp - 1.0f;
    fx = tmp;
    tmp = fx * 0.693359375f;
    float z = fx * -2.12194440e-4f;
    x = x - tmp;
    x = x - z;
    z = x * x;
    float y = 1.9875691500E-4f;
    y = y * x; y = y + 1.3981999507E-3f;
    y = y * x; y = y + 8.3334519073E-3f;
    y = y * x; y = y + 4.1665795894E-2f;
    y = y * x; y = y + 1.6666665459E-1f;
    y = y * x; y = y + 5.0000001201E-1f;
    y = y * z;
    y = y + x;
    y = y + 1.0f;
    int32_t n = (int32_t)fx;
    uint32_t pb = (uint32_t)(n + 0x7f) << 23;
    float pow2n; memcpy(&pow2n, &pb, 4);
    return y * pow2n;
}

// ---------------------------------------------------------------- box down-samplers
// mipmap.d:676-762: (a+b+c+d+2)>>2 per channel.
void generateLevelBoxRGBA(OwnedImage<RGBA>& dst, OwnedImage<RGBA>& src, box2i r) {
    const int width = r.width(), height = r.height();
    const __m128i zero = _mm_setzero_si128(), two = _mm_set1_epi16(2);
    for (int y = 0; y < height; ++y) {
        const RGBA* L0 = src.scanlinePtr((r.miny + y) * 2) + r.minx * 2;
        const RGBA* L1 = src.scanlinePtr((r.miny + y) * 2 + 1) + r.minx * 2;
        RGBA* d = dst.scanlinePtr(r.miny + y) + r.minx;
        int x = 0;
        for (; x + 1 < width; x += 2) {
            __m128i abef = _mm_loadu_si128((const __m128i*)&L0[2 * x]);
            __m128i cdgh = _mm_loadu_si128((const __m128i*)&L1[2 * x]);
            __m128i ab = _mm_add_epi16(_mm_unpacklo_epi8(abef, zero), _mm_unpacklo_epi8(cdgh, zero));
            __m128i ef = _mm_add_epi16(_mm_unpackhi_epi8(abef, zero), _mm_unpackhi_epi8(cdgh, zero));
            __m128i sum = _mm_add_epi16(_mm_unpacklo_epi64(ab, ef), _mm_unpackhi_epi64(ab, ef));
            sum = _mm_srai_epi16(_mm_add_epi16(sum, two), 2);
            _mm_storel_epi64((__m128i*)&d[x], _mm_packus_epi16(sum, zero));
        }
        for (; x < width; ++x) {
            RGBA A = L0[2 * x], B = L0[2 * x + 1], C = L1[2 * x], D = L1[2 * x + 1];
            d[x] = RGBA{(uint8_t)((A.r + B.r + C.r + D.r + 2) >> 2), (uint8_t)((A.g + B.g + C.g + D.g + 2) >> 2),
                        (uint8_t)((A.b + B.b + C.b + D.b + 2) >> 2), (uint8_t)((A.a + B.a + C.a + D.a + 2) >> 2)};
        }
    }
}

// mipmap.d:764-794
void generateLevelBoxL16(OwnedImage<L16>& dst, OwnedImage<L16>& src, box2i r) {
    const int width = r.width(), height = r.height();
    for (int y = 0; y < height; ++y) {
        const L16* L0 = src.scanlinePtr((r.miny + y) * 2) + r.minx * 2;
        const L16* L1 = src.scanlinePtr((r.miny + y) * 2 + 1) + r.minx * 2;
        L16* d = dst.scanlinePtr(r.miny + y) + r.minx;
        for (int x = 0; x < width; ++x)
            d[x].l = (uint16_t)(((int)L0[2 * x].l + L0[2 * x + 1].l + L1[2 * x].l + L1[2 * x + 1].l + 2) >> 2);
    }
}

// mipmap.d:823-900, non-legacy branch: gamma->linear (square), alpha-premultiply,
// mean of 4, truncating conversion after +0.5.
static inline RGBAf gammaToLinearPremul(RGBA c) {  // mipmap.d:869-878
    const float inv_255 = 1.0f / 255;
    RGBAf o;
    o.a = (float)c.a * inv_255;
    o.r = (float)c.r * inv_255 * (float)c.r * inv_255 * o.a;
    o.g = (float)c.g * inv_255 * (float)c.g * inv_255 * o.a;
    o.b = (float)c.b * inv_255 * (float)c.b * inv_255 * o.a;
    return o;
}
void generateLevelBoxAlphaCovIntoPremulRGBA(OwnedImage<RGBA>& dst, OwnedImage<RGBA>& src, box2i r) {
    const int width = r.width(), height = r.height();
    for (int y = 0; y < height; ++y) {
        const RGBA* L0 = src.scanlinePtr((r.miny + y) * 2) + r.minx * 2;
        const RGBA* L1 = src.scanlinePtr((r.miny + y) * 2 + 1) + r.minx * 2;
        RGBA* d = dst.scanlinePtr(r.miny + y) + r.minx;
        for (int x = 0; x < width; ++x) {
            RGBAf A = gammaToLinearPremul(L0[2 * x]), B = gammaToLinearPremul(L0[2 * x + 1]);
            RGBAf C = gammaToLinearPremul(L1[2 * x]), D = gammaToLinearPremul(L1[2 * x + 1]);
            float mR = A.r + B.r + C.r + D.r;
            float mG = A.g + B.g + C.g + D.g;
            float mB = A.b + B.b + C.b + D.b;
            float mA = A.a + B.a + C.a + D.a;
            d[x] = RGBA{(uint8_t)(mR * 0.25f * 255.0f + 0.5f), (uint8_t)(mG * 0.25f * 255.0f + 0.5f),
                        (uint8_t)(mB * 0.25f * 255.0f + 0.5f), (uint8_t)(mA * 0.25f * 255.0f + 0.5f)};
        }
    }
}

// ---------------------------------------------------------------- cubic down-samplers
// mipmap.d:902-1040: [1 3 3 1]x[1 3 3 1]/64 with clamped outer taps, 16-bit SSE.
void generateLevelCubicRGBA(OwnedImage<RGBA>& dst, OwnedImage<RGBA>& src, box2i r) {
    const __m128i zero = _mm_setzero_si128();
    const __m128i k1133 = _mm_setr_epi16(1, 1, 1, 1, 3, 3, 3, 3);
    const __m128i k3399 = _mm_setr_epi16(3, 3, 3, 3, 9, 9, 9, 9);
    for (int y = r.miny; y < r.maxy; ++y) {
        int ym = 2 * y - 1; if (ym < 0) ym = 0;
        int yp = 2 * y + 2; if (yp > src.h - 1) yp = src.h - 1;
        const RGBA* LM1 = src.scanlinePtr(ym);
        const RGBA* L0 = src.scanlinePtr(2 * y);
        const RGBA* L1 = src.scanlinePtr(2 * y + 1);
        const RGBA* L2 = src.scanlinePtr(yp);
        RGBA* d = dst.scanlinePtr(y);
        for (int x = r.minx; x < r.maxx; ++x) {
            int xm = 2 * x - 1; if (xm < 0) xm = 0;
            int x0 = 2 * x;
            int xp = 2 * x + 2; if (xp > src.w - 1) xp = src.w - 1;
            alignas(16) RGBA buf[16] = {LM1[xm], LM1[x0], LM1[x0 + 1], LM1[xp], L0[xm], L0[x0], L0[x0 + 1], L0[xp],
                                        L1[xm],  L1[x0],  L1[x0 + 1],  L1[xp],  L2[xm], L2[x0], L2[x0 + 1], L2[xp]};
            __m128i abcd = _mm_load_si128((const __m128i*)&buf[0]);
            __m128i efgh = _mm_load_si128((const __m128i*)&buf[4]);
            __m128i ijkl = _mm_load_si128((const __m128i*)&buf[8]);
            __m128i mnop = _mm_load_si128((const __m128i*)&buf[12]);
            __m128i ab = _mm_add_epi16(_mm_unpacklo_epi8(abcd, zero), _mm_unpacklo_epi8(mnop, zero));
            __m128i cd = _mm_add_epi16(_mm_unpackhi_epi8(abcd, zero), _mm_unpackhi_epi8(mnop, zero));
            __m128i ef = _mm_add_epi16(_mm_unpacklo_epi8(efgh, zero), _mm_unpacklo_epi8(ijkl, zero));
            __m128i gh = _mm_add_epi16(_mm_unpackhi_epi8(efgh, zero), _mm_unpackhi_epi8(ijkl, zero));
            ab = _mm_add_epi16(ab, _mm_shuffle_epi32(cd, 0x4e));
            ef = _mm_add_epi16(ef, _mm_shuffle_epi32(gh, 0x4e));
            __m128i s = _mm_add_epi16(_mm_mullo_epi16(ab, k1133), _mm_mullo_epi16(ef, k3399));
            s = _mm_add_epi16(s, _mm_srli_si128(s, 8));
            s = _mm_srli_epi16(_mm_add_epi16(s, _mm_set1_epi16(32)), 6);
            int32_t px = _mm_cvtsi128_si32(_mm_packus_epi16(s, zero));
            memcpy(&d[x], &px, 4);
        }
    }
}

// mipmap.d:1042-1109
void generateLevelCubicL16(OwnedImage<L16>& dst, OwnedImage<L16>& src, box2i r) {
    for (int y = r.miny; y < r.maxy; ++y) {
        int ym = 2 * y - 1; if (ym < 0) ym = 0;
        int yp = 2 * y + 2; if (yp > src.h - 1) yp = src.h - 1;
        const L16* LM1 = src.scanlinePtr(ym);
        const L16* L0 = src.scanlinePtr(2 * y);
        const L16* L1 = src.scanlinePtr(2 * y + 1);
        const L16* L2 = src.scanlinePtr(yp);
        L16* d = dst.scanlinePtr(y);
        for (int x = r.minx; x < r.maxx; ++x) {
            int xm = 2 * x - 1; if (xm < 0) xm = 0;
            int x0 = 2 * x;
            int xp = 2 * x + 2; if (xp > src.w - 1) xp = src.w - 1;
            int corners = LM1[xm].l + LM1[xp].l + L2[xm].l + L2[xp].l;
            int edges = LM1[x0].l + LM1[x0 + 1].l + L0[xm].l + L0[xp].l + L1[xm].l + L1[xp].l + L2[x0].l + L2[x0 + 1].l;
            int centre = L0[x0].l + L0[x0 + 1].l + L1[x0].l + L1[x0 + 1].l;
            int sum = corners + 3 * edges + 9 * centre;
            d[x].l = (uint16_t)((sum + 32) >> 6);
        }
    }
}

// ---------------------------------------------------------------- boxlist.d
box2i boundingBox(const std::vector<box2i>& boxes) {
    if (boxes.empty()) return box2i{0, 0, 0, 0};
    box2i u = boxes[0];
    for (size_t i = 1; i < boxes.size(); ++i) u = u.expand(boxes[i]);
    return u;
}

void boxSubtraction(const box2i& A, const box2i& C, box2i& D, box2i& E, box2i& F, box2i& G) {
    D = box2i{A.minx, A.miny, A.maxx, C.miny};
    E = box2i{A.minx, C.miny, C.minx, C.maxy};
    F = box2i{C.maxx, C.miny, A.maxx, C.maxy};
    G = box2i{A.minx, C.maxy, A.maxx, A.maxy};
}

void removeOverlappingAreas(std::vector<box2i>& boxes, std::vector<box2i>& filtered) {
    for (int i = 0; i < (int)boxes.size(); ++i) {
        box2i A = boxes[i];
        if (A.width() * A.height() <= 0) continue;
        bool found = false;
        for (int j = i + 1; j < (int)boxes.size(); ++j) {
            box2i B = boxes[j];
            box2i C = A.intersection(B);
            if (C.isSorted() && !C.empty()) {
                if (A.contains(B)) {  // B swallowed: remove-and-replace-by-last (boxlist.d:79-85)
                    boxes[j] = boxes.back();
                    boxes.pop_back();
                    j = j - 1;
                    continue;
                }
                found = true;
                if (B.contains(A)) break;
                box2i D, E, F, G;
                boxSubtraction(A, C, D, E, F, G);
                if (!D.empty()) boxes.push_back(D);
                if (!E.empty()) boxes.push_back(E);
                if (!F.empty()) boxes.push_back(F);
                if (!G.empty()) boxes.push_back(G);
                break;
            }
        }
        if (!found) filtered.push_back(A);
    }
}

void dirtyRectListAdd(std::vector<box2i>& list, box2i rect) {  // boxlist.d:236-292
    if (rect.empty()) return;
    bool processed = false;
    for (int i = 0; i < (int)list.size(); ++i) {
        box2i other = list[i];
        if (other.contains(rect)) {  // candidate inside an element: discard it (boxlist.d:249-254)
            processed = true;
            break;
        } else if (rect.contains(other)) {  // element inside the candidate: it goes (boxlist.d:255-260)
            box2i last = list.back();
            list.pop_back();
            if (i < (int)list.size()) list[i] = last;
            i--;
        } else {
            box2i common = other.intersection(rect);
            if (!common.empty()) {  // keep `other` minus the common part (boxlist.d:263-280)
                box2i D, E, F, G;
                boxSubtraction(other, common, D, E, F, G);
                box2i last = list.back();
                list.pop_back();
                if (i < (int)list.size()) list[i] = last;
                i--;
                if (!D.empty()) list.push_back(D);
                if (!E.empty()) list.push_back(E);
                if (!F.empty()) list.push_back(F);
                if (!G.empty()) list.push_back(G);
            }
        }
    }
    if (!processed) list.push_back(rect);
}

void tileAreas(const box2i* areas, int n, int maxW, int maxH, std::vector<box2i>& out) {
    for (int k = 0; k < n; ++k) {
        const box2i& a = areas[k];
        int nW = (a.width() + maxW - 1) / maxW;
        int nH = (a.height() + maxH - 1) / maxH;
        for (int j = 0; j < nH; ++j) {
            int y0 = maxH * j, y1 = y0 + maxH;
            if (y1 > a.height()) y1 = a.height();
            for (int i = 0; i < nW; ++i) {
                int x0 = maxW * i, x1 = x0 + maxW;
                if (x1 > a.width()) x1 = a.width();
                out.push_back(box2i{x0, y0, x1, y1}.translate(a.minx, a.miny));
            }
        }
    }
}

bool haveNoOverlap(const box2i* areas, int n) {
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (areas[i].intersects(areas[j])) return false;
    return true;
}

}  // namespace b2ref
