// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.hpp header).
//
// The 8 PBR passes (gui/dplug/gui/legacypbr.d:269-1108), the plane-fit normal
// (gui/dplug/gui/ransac.d:53-121), the per-tile driver (legacypbr.d:201-260) and
// the GUIGraphics slice that feeds it (graphics.d:1115-1137, 1338-1423).
// Float arithmetic: IEEE binary32, one operation per statement fragment in the
// reference's order; build with -ffp-contract=off.
#include "ref_pbr.hpp"
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

namespace b2ref {

static const float div255 = 1 / 255.0f;            // legacypbr.d:1175
static const float FACTOR_Z = 4655.0f;             // legacypbr.d:63

// legacypbr.d:1012-1030
static const uint8_t BLUE_NOISE_16x16[256] = {
    127, 194, 167, 79,  64,  173, 22,  83,  167, 105, 119, 250, 201, 34,  214, 145,
    233, 56,  13,  251, 203, 124, 243, 42,  216, 34,  73,  175, 133, 64,  185, 73,
    93,  156, 109, 144, 34,  98,  153, 138, 187, 238, 155, 46,  13,  102, 247, 0,
    28,  180, 46,  218, 183, 13,  212, 69,  13,  92,  126, 228, 211, 161, 117, 197,
    134, 240, 121, 75,  234, 88,  53,  170, 109, 204, 59,  22,  86,  141, 38,  222,
    81,  205, 13,  59,  160, 198, 129, 252, 0,   147, 176, 193, 244, 71,  173, 56,
    22,  168, 104, 139, 22,  114, 38,  220, 101, 231, 77,  34,  113, 13,  189, 96,
    253, 148, 227, 190, 246, 174, 66,  155, 28,  50,  164, 131, 217, 151, 232, 128,
    115, 69,  34,  50,  93,  13,  209, 85,  192, 120, 248, 64,  90,  28,  208, 42,
    0,   200, 215, 79,  125, 148, 239, 136, 181, 22,  206, 13,  185, 108, 59,  179,
    90,  130, 159, 182, 235, 42,  106, 0,   56,  99,  226, 140, 157, 237, 77,  165,
    249, 28,  105, 13,  61,  170, 224, 75,  202, 163, 114, 81,  46,  22,  137, 223,
    189, 53,  219, 142, 196, 28,  122, 154, 254, 42,  28,  242, 196, 210, 119, 38,
    149, 86,  118, 245, 71,  96,  213, 13,  88,  178, 66,  129, 171, 0,   99,  69,
    178, 13,  207, 38,  159, 187, 50,  132, 236, 146, 191, 95,  53,  229, 163, 241,
    46,  225, 102, 135, 0,   230, 110, 199, 61,  0,   221, 22,  150, 83,  112, 22};

// ---------------------------------------------------------------- small SSE-shaped helpers
struct f4 { float v[4]; };
static inline f4 baseColor4(RGBA c) {  // convertBaseColorToFloat4, legacypbr.d:1155-1163
    return f4{{(float)c.r * div255, (float)c.g * div255, (float)c.b * div255, (float)c.a * div255}};
}
static inline float dot_ps(const f4& a, const f4& b) {  // ransac.d:14-18
    float m0 = a.v[0] * b.v[0], m1 = a.v[1] * b.v[1], m2 = a.v[2] * b.v[2], m3 = a.v[3] * b.v[3];
    return m0 + m1 + m2 + m3;
}
static inline f4 fastNormalize(f4 v) {  // ransac.d:39-46
    float s0 = v.v[0] * v.v[0], s1 = v.v[1] * v.v[1], s2 = v.v[2] * v.v[2], s3 = v.v[3] * v.v[3];
    float len2 = s0 + s1 + s2 + s3;
    float inv = rsqrtss(len2);
    return f4{{v.v[0] * inv, v.v[1] * inv, v.v[2] * inv, v.v[3] * inv}};
}
static inline f4 reflectNormal(const f4& a, const f4& n) {  // ransac.d:31-36
    float sum = 2 * dot_ps(a, n);
    return f4{{a.v[0] - sum * n.v[0], a.v[1] - sum * n.v[1], a.v[2] - sum * n.v[2], a.v[3] - sum * n.v[3]}};
}

// ransac.d:53-121
static inline void planeFittingNormal(const float d[9], float out[3]) {
    const float c = 0.03660694846, e = 0.118115527008;
    const float wA[4] = {c, e, c, e}, wB[4] = {e, c, e, c};
    const float xzA[4] = {-c, 0.0f, c, -e}, xzB[4] = {e, -c, 0.0f, c};
    const float yzA[4] = {-c, -e, -c, 0.0f}, yzB[4] = {0.0f, c, e, c};
    float A[4] = {d[0], d[1], d[2], d[3]}, B[4] = {d[5], d[6], d[7], d[8]};
    float m[4];
    for (int k = 0; k < 4; ++k) m[k] = A[k] * wA[k] + B[k] * wB[k];
    float filt = d[4] * 0.38111009811f + m[0] + m[1] + m[2] + m[3];
    for (int k = 0; k < 4; ++k) { A[k] = A[k] - filt; B[k] = B[k] - filt; }
    float XZ[4], YZ[4];
    for (int k = 0; k < 4; ++k) {
        XZ[k] = A[k] * xzA[k] + B[k] * xzB[k];
        YZ[k] = A[k] * yzA[k] + B[k] * yzB[k];
    }
    float xz = XZ[0] + XZ[1] + XZ[2] + XZ[3];
    float yz = YZ[0] + YZ[1] + YZ[2] + YZ[3];
    f4 n = fastNormalize(f4{{-xz, yz, 0.382658847856f, 0.0f}});
    out[0] = n.v[0]; out[1] = n.v[1]; out[2] = n.v[2];
}

void planeFittingNormalPublic(const float d[9], float out[3]) { planeFittingNormal(d, out); }

// ---------------------------------------------------------------- parameters
static void normalized3(float x, float y, float z, float out[3]) {  // math/dplug/math/vector.d:319-324,344-347,365-369
    float s = 0;
    s += x * x; s += y * y; s += z * z;
    float inv = 1 / std::sqrt(s);
    out[0] = x * inv; out[1] = y * inv; out[2] = z * inv;
}

void defaultParams(PBRParams* p) {
    for (int k = 0; k < 3; ++k) p->light1Color[k] = 0.25f * 1.3f;
    normalized3(0.0f, 1.0f, 0.1f, p->light2Dir);
    for (int k = 0; k < 3; ++k) p->light2Color[k] = 0.481f;
    normalized3(0.0f, 1.0f, 0.1f, p->light3Dir);
    for (int k = 0; k < 3; ++k) p->light3Color[k] = 0.26f;
    p->ambientLight = 0.08125f;
    p->skyboxAmount = 0.52f;
    p->tonemapThreshold = 1.0f;
    p->tonemapRatio = 0.3f;
    p->passActiveMask = 0xffu;
}

// ---------------------------------------------------------------- compositor
PBRCompositor::PBRCompositor(int numThreads) : _numThreads(numThreads < 1 ? 1 : numThreads) {
    defaultParams(&params);
    for (int t = 0; t < _numThreads; ++t) {
        _normalBuffers.push_back(new OwnedImage<RGBf>());
        _accumBuffers.push_back(new OwnedImage<RGBAf>());
        _varianceBuffers.push_back(new OwnedImage<float>());
    }
    _specularFactor.resize(_numThreads);
    _exponentFactor.resize(_numThreads);
    _toksvigScaleFactor.resize(_numThreads);
    for (int r = 0; r < 256; ++r) {  // legacypbr.d:696-702
        float arg = (1 - r / 255.0f) * 5.5f;
        float t = 0.8f * (float)std::exp((double)arg);
        t *= 2.8f;
        _exponentTable[r] = t;
    }
}

PBRCompositor::~PBRCompositor() {
    for (auto* p : _normalBuffers) delete p;
    for (auto* p : _accumBuffers) delete p;
    for (auto* p : _varianceBuffers) delete p;
    delete _skybox;
}

void PBRCompositor::resizeBuffers(int width, int, int areaMaxWidth, int areaMaxHeight) {
    for (int t = 0; t < _numThreads; ++t) {
        _normalBuffers[t]->size(areaMaxWidth, areaMaxHeight, 0, 1);
        _accumBuffers[t]->size(areaMaxWidth, areaMaxHeight, 0, 16);
        _varianceBuffers[t]->size(areaMaxWidth, areaMaxHeight, 0, 1);
        _specularFactor[t].assign(width, 0.f);
        _exponentFactor[t].assign(width, 0.f);
        _toksvigScaleFactor[t].assign(width, 0.f);
    }
}

void PBRCompositor::setSkybox(const RGBA* pixels, int w, int h, size_t pitch) {
    delete _skybox;
    _skybox = new Mipmap<RGBA>(12, w, h);
    for (int y = 0; y < h; ++y)
        memcpy(_skybox->levels[0]->scanlinePtr(y), (const uint8_t*)pixels + (size_t)y * pitch, (size_t)w * 4);
    _skybox->generateMipmaps(Q_BOX);
}

// legacypbr.d:279-381
void PBRCompositor::passNormal(int t, box2i area, Mipmap<L16>& depthMap) {
    OwnedImage<RGBf>& normalBuffer = *_normalBuffers[t];
    OwnedImage<float>& varianceBuffer = *_varianceBuffers[t];
    OwnedImage<L16>& depth0 = *depthMap.levels[0];
    const int pitch = depth0.pitchInBytes();
    const float multUshort = 1.0 / FACTOR_Z;  // enum float = double division narrowed, legacypbr.d:304
    for (int j = area.miny; j < area.maxy; ++j) {
        RGBf* normalScan = normalBuffer.scanlinePtr(j - area.miny);
        float* varianceScan = varianceBuffer.scanlinePtr(j - area.miny);
        const L16* depthScan = depth0.scanlinePtr(j);
        for (int i = area.minx; i < area.maxx; ++i) {
            const L16* here = depthScan + i;
            const L16* up = (const L16*)((const uint8_t*)here - pitch);
            const L16* dn = (const L16*)((const uint8_t*)here + pitch);
            float d[9] = {up[-1].l * multUshort, up[0].l * multUshort, up[1].l * multUshort,
                          here[-1].l * multUshort, here[0].l * multUshort, here[1].l * multUshort,
                          dn[-1].l * multUshort, dn[0].l * multUshort, dn[1].l * multUshort};
            float n[3];
            planeFittingNormal(d, n);
            normalScan[i - area.minx] = RGBf{n[0], n[1], n[2]};

            // "old method" variance, legacypbr.d:356-377; lane 3 is multiplied by 0 and never summed
            float M1[3] = {(float)up[-1].l, (float)up[0].l, (float)up[1].l};
            float P0[3] = {(float)here[-1].l, (float)here[0].l, (float)here[1].l};
            float P1[3] = {(float)dn[-1].l, (float)dn[0].l, (float)dn[1].l};
            const float fx[3] = {1.0f, -2.0f, 1.0f};
            float ddx[3], ddy[3];
            for (int k = 0; k < 3; ++k) {
                ddx[k] = fx[k] * (M1[k] + P0[k] + P1[k]);
                ddy[k] = 1.0f * (M1[k] + P1[k]) + P0[k] * -2.0f;
            }
            float depthDX = ddx[0] + ddx[1] + ddx[2];
            float depthDY = ddy[0] + ddy[1] + ddy[2];
            depthDX *= (1 / 256.0f);
            depthDY *= (1 / 256.0f);
            varianceScan[i - area.minx] = depthDX * depthDX + depthDY * depthDY;
        }
    }
}

// legacypbr.d:400-444
void PBRCompositor::passAO(int t, box2i area, Mipmap<RGBA>& diffuseMap, Mipmap<L16>& depthMap) {
    OwnedImage<RGBA>& diffuse0 = *diffuseMap.levels[0];
    OwnedImage<L16>& depth0 = *depthMap.levels[0];
    OwnedImage<RGBAf>& accumBuffer = *_accumBuffers[t];
    const float amount = params.ambientLight;
    const float divider23040 = 1.0f / 23040;
    for (int j = area.miny; j < area.maxy; ++j) {
        const RGBA* diffuseScan = diffuse0.scanlinePtr(j);
        const L16* depthScan = depth0.scanlinePtr(j);
        RGBAf* accumScan = accumBuffer.scanlinePtr(j - area.miny);
        for (int i = area.minx; i < area.maxx; ++i) {
            f4 base = baseColor4(diffuseScan[i]);
            float px = i + 0.5f, py = j + 0.5f;
            float avg = (depthMap.linearSample(1, px, py) + depthMap.linearSample(2, px, py)
                         + depthMap.linearSample(3, px, py) + depthMap.linearSample(4, px, py)) * 0.25f;
            float diff = depthScan[i].l - avg;
            float cavity = (diff + 23040.0f) * divider23040;
            if (cavity >= 1) cavity = 1;
            else if (cavity < 0) cavity = 0;
            float k = cavity * amount;
            RGBAf& a = accumScan[i - area.minx];
            a.r = a.r + base.v[0] * k; a.g = a.g + base.v[1] * k; a.b = a.b + base.v[2] * k; a.a = a.a + base.v[3] * k;
        }
    }
}

// legacypbr.d:460-616
void PBRCompositor::passObliqueShadow(int t, box2i area, Mipmap<RGBA>& diffuseMap, Mipmap<L16>& depthMap) {
    OwnedImage<L16>& depth0 = *depthMap.levels[0];
    OwnedImage<RGBA>& diffuse0 = *diffuseMap.levels[0];
    OwnedImage<RGBAf>& accumBuffer = *_accumBuffers[t];
    const float fallOff = 0.78f;
    const int samples = 11;
    float weights[11];
    for (int s = 0; s < 11; ++s) weights[s] = (float)std::pow((double)fallOff, s);  // CTFE `^^`, legacypbr.d:474-487
    const float totalWeights = (float)((1.0 - std::pow((double)fallOff, 11)) / (1.0 - (double)fallOff) - 1);
    const float invTotalWeights = (float)(1 / ((double)1.7f * (double)totalWeights));
    const int wholeWidth = depth0.w;
    const float* color = params.light1Color;

    for (int j = area.miny; j < area.maxy; ++j) {
        const RGBA* diffuseScan = diffuse0.scanlinePtr(j);
        const L16* depthScan = depth0.scanlinePtr(j);
        RGBAf* accumScan = accumBuffer.scanlinePtr(j - area.miny);
        for (int i = area.minx; i < area.maxx; ++i) {
            RGBA ib = diffuseScan[i];
            float base[3] = {(float)ib.r * div255, (float)ib.g * div255, (float)ib.b * div255};
            float lightPassed = 0.0f;
            const int depthCenter = depthScan[i].l;
            int sample = 1;
            for (; sample + 3 < samples; sample += 4) {  // SSE groups 1-4 and 5-8, legacypbr.d:519-564
                float contrib[4];
                for (int lane = 0; lane < 4; ++lane) {
                    int s = sample + lane;
                    int x1 = i + s, x2 = i - s, y = j - s;
                    if (x1 > wholeWidth - 1) x1 = wholeWidth - 1;
                    if (x2 < 0) x2 = 0;
                    if (y < 0) y = 0;
                    int z = depthCenter + s;
                    const L16* scan = depth0.scanlinePtr(y);
                    float diff1 = (float)(z - (int)scan[x1].l);
                    float diff2 = (float)(z - (int)scan[x2].l);
                    float v1 = 1.0f + diff1 * 0.00006510416f;
                    float v2 = 1.0f + diff2 * 0.00006510416f;
                    float c1 = (1.0f < v1) ? 1.0f : v1;  // minps(ones, v)
                    float c2 = (1.0f < v2) ? 1.0f : v2;
                    c1 = (0.0f > c1) ? 0.0f : c1;        // maxps(zero, c)
                    c2 = (0.0f > c2) ? 0.0f : c2;
                    contrib[lane] = (c1 + c2 * 0.7f) * weights[s];
                }
                lightPassed += contrib[0];
                lightPassed += contrib[1];
                lightPassed += contrib[2];
                lightPassed += contrib[3];
            }
            for (; sample < samples; ++sample) {  // scalar tail s=9,10, legacypbr.d:566-609
                int x1 = i + sample; if (x1 >= wholeWidth) x1 = wholeWidth - 1;
                int x2 = i - sample; if (x2 < 0) x2 = 0;
                int y = j - sample;  if (y < 0) y = 0;
                int z = depthCenter + sample;
                const L16* scan = depth0.scanlinePtr(y);
                int diff1 = z - scan[x1].l, diff2 = z - scan[x2].l;
                const float divider15360 = 1.0f / 15360;
                float c1, c2;
                if (diff1 >= 0) c1 = 1; else if (diff1 < -15360) c1 = 0; else c1 = (diff1 + 15360) * divider15360;
                if (diff2 >= 0) c2 = 1; else if (diff2 < -15360) c2 = 0; else c2 = (diff2 + 15360) * divider15360;
                lightPassed += (c1 + c2 * 0.7f) * weights[sample];
            }
            float k = lightPassed * invTotalWeights;
            RGBAf& a = accumScan[i - area.minx];
            a.r = a.r + base[0] * color[0] * k;
            a.g = a.g + base[1] * color[1] * k;
            a.b = a.b + base[2] * color[2] * k;
            a.a = a.a + 0.0f;
        }
    }
}

// legacypbr.d:636-666
void PBRCompositor::passDirectional(int t, box2i area, Mipmap<RGBA>& diffuseMap, Mipmap<RGBA>& materialMap) {
    OwnedImage<RGBA>& diffuse0 = *diffuseMap.levels[0];
    OwnedImage<RGBA>& material0 = *materialMap.levels[0];
    OwnedImage<RGBf>& normalBuffer = *_normalBuffers[t];
    OwnedImage<RGBAf>& accumBuffer = *_accumBuffers[t];
    const float* dir = params.light2Dir;
    const float* color = params.light2Color;
    for (int j = area.miny; j < area.maxy; ++j) {
        const RGBA* materialScan = material0.scanlinePtr(j);
        const RGBA* diffuseScan = diffuse0.scanlinePtr(j);
        const RGBf* normalScan = normalBuffer.scanlinePtr(j - area.miny);
        RGBAf* accumScan = accumBuffer.scanlinePtr(j - area.miny);
        for (int i = area.minx; i < area.maxx; ++i) {
            RGBf n = normalScan[i - area.minx];
            float roughness = materialScan[i].r * div255;
            RGBA ib = diffuseScan[i];
            float base[3] = {(float)ib.r * div255, (float)ib.g * div255, (float)ib.b * div255};
            float dot = 0;  // math/dplug/math/vector.d:587-592
            dot += n.r * dir[0]; dot += n.g * dir[1]; dot += n.b * dir[2];
            float f = 0.5f + 0.5f * dot;
            f = linmap<float>(f, 0.24f - roughness * 0.5f, 1, 0, 1.0f);
            RGBAf& a = accumScan[i - area.minx];
            a.r += base[0] * color[0] * f;
            a.g += base[1] * color[1] * f;
            a.b += base[2] * color[2] * f;
            a.a += 0.0f;
        }
    }
}

// legacypbr.d:720-813
void PBRCompositor::passSpecular(int t, box2i area, Mipmap<RGBA>& diffuseMap, Mipmap<RGBA>& materialMap) {
    OwnedImage<RGBA>& diffuse0 = *diffuseMap.levels[0];
    OwnedImage<RGBA>& material0 = *materialMap.levels[0];
    OwnedImage<RGBf>& normalBuffer = *_normalBuffers[t];
    OwnedImage<RGBAf>& accumBuffer = *_accumBuffers[t];
    OwnedImage<float>& varianceBuffer = *_varianceBuffers[t];
    const int w = diffuse0.w, h = diffuse0.h;
    const float invW = 1.0f / w, invH = 1.0f / h;
    const f4 mmlight3Dir{{-params.light3Dir[0], -params.light3Dir[1], -params.light3Dir[2], 0.0f}};
    const float* color = params.light3Color;
    float* pSpecular = _specularFactor[t].data();
    float* pExponent = _exponentFactor[t].data();
    float* pToksvig = _toksvigScaleFactor[t].data();
    for (int j = area.miny; j < area.maxy; ++j) {
        const RGBA* materialScan = material0.scanlinePtr(j);
        const RGBA* diffuseScan = diffuse0.scanlinePtr(j);
        const RGBf* normalScan = normalBuffer.scanlinePtr(j - area.miny);
        RGBAf* accumScan = accumBuffer.scanlinePtr(j - area.miny);
        const float* varianceScan = varianceBuffer.scanlinePtr(j - area.miny);
        for (int i = area.minx; i < area.maxx; ++i) {
            RGBf nb = normalScan[i - area.minx];
            f4 normal{{nb.r, nb.g, nb.b, 0.0f}};
            f4 toEye = fastNormalize(f4{{0.5f - i * invW, j * invH - 0.5f, 1.0f, 0.0f}});
            f4 half{{toEye.v[0] - mmlight3Dir.v[0], toEye.v[1] - mmlight3Dir.v[1], toEye.v[2] - mmlight3Dir.v[2],
                     toEye.v[3] - mmlight3Dir.v[3]}};
            half = fastNormalize(half);
            float specularFactor = dot_ps(half, normal);
            if (specularFactor < 1e-3f) specularFactor = 1e-3f;
            float exponent = _exponentTable[materialScan[i].r];
            const float VARIANCE_FACTOR = 4e-5f;
            float variance = varianceScan[i - area.minx];
            float Ft = 1.0f / (1.0f + exponent * variance * VARIANCE_FACTOR);
            float tok = (1.0f + exponent * Ft) / (1.0f + exponent);
            pToksvig[i] = tok;
            pSpecular[i] = specularFactor;
            pExponent[i] = exponent * Ft;
        }
        // _mm_pow_ps on groups of 4, _mm_pow_ss on the tail: same per-lane function (legacypbr.d:782-792)
        for (int i = area.minx; i < area.maxx; ++i) pSpecular[i] = pow_ps(pSpecular[i], pExponent[i]);
        for (int i = area.minx; i < area.maxx; ++i) {
            float specularFactor = pSpecular[i];
            f4 material = baseColor4(materialScan[i]);
            float roughness = material.v[0], metalness = material.v[1], specular = material.v[2];
            f4 base = baseColor4(diffuseScan[i]);
            float roughFactor = 10 * (1.0f - roughness) * (1 - metalness * 0.5f);
            specularFactor = specularFactor * roughFactor * pToksvig[i];
            float k = specularFactor * specular;
            RGBAf& a = accumScan[i - area.minx];
            a.r = a.r + base.v[0] * color[0] * k;
            a.g = a.g + base.v[1] * color[1] * k;
            a.b = a.b + base.v[2] * color[2] * k;
            a.a = a.a + base.v[3] * 0.0f * k;
        }
    }
}

// legacypbr.d:872-930
void PBRCompositor::passSkybox(int t, box2i area, Mipmap<RGBA>& diffuseMap, Mipmap<RGBA>& materialMap) {
    if (!_skybox) return;
    OwnedImage<RGBA>& diffuse0 = *diffuseMap.levels[0];
    OwnedImage<RGBA>& material0 = *materialMap.levels[0];
    OwnedImage<RGBf>& normalBuffer = *_normalBuffers[t];
    OwnedImage<RGBAf>& accumBuffer = *_accumBuffers[t];
    OwnedImage<float>& varianceBuffer = *_varianceBuffers[t];
    const int w = diffuse0.w, h = diffuse0.h;
    const float invW = 1.0f / w, invH = 1.0f / h;
    for (int j = area.miny; j < area.maxy; ++j) {
        const RGBA* materialScan = material0.scanlinePtr(j);
        const RGBA* diffuseScan = diffuse0.scanlinePtr(j);
        const RGBf* normalScan = normalBuffer.scanlinePtr(j - area.miny);
        RGBAf* accumScan = accumBuffer.scanlinePtr(j - area.miny);
        const float* varianceScan = varianceBuffer.scanlinePtr(j - area.miny);
        const float amountOfSkyboxPixels = (float)(_skybox->width() * _skybox->height());
        for (int i = area.minx; i < area.maxx; ++i) {
            float mipmapLevel = varianceScan[i - area.minx] * amountOfSkyboxPixels;
            const float ROUGH_FACT = 6.0f / 255.0f;
            float roughness = materialScan[i].r;
            mipmapLevel = 0.5f * fastlog2(1.0f + mipmapLevel * 0.00001f) + ROUGH_FACT * roughness;
            const float fskyX = _skybox->width() - 1.0f;
            const float fskyY = _skybox->height() - 1.0f;
            const float amountFactor = params.skyboxAmount * div255;
            f4 toEye = fastNormalize(f4{{0.5f - i * invW, j * invH - 0.5f, 1.0f, 0.0f}});
            RGBf nb = normalScan[i - area.minx];
            f4 normal{{nb.r, nb.g, nb.b, 0.0f}};
            f4 refl = reflectNormal(toEye, normal);
            f4 material = baseColor4(materialScan[i]);
            float metalness = material.v[1];
            f4 base = baseColor4(diffuseScan[i]);
            float skyx = 0.5f + ((0.5f - refl.v[0] * 0.5f) * fskyX);
            float skyy = 0.5f + ((0.5f + refl.v[1] * 0.5f) * fskyY);
            vec4f sky = _skybox->linearMipmapSample(mipmapLevel, skyx, skyy);
            float k = metalness * amountFactor;
            RGBAf& a = accumScan[i - area.minx];
            a.r = a.r + base.v[0] * sky.x * k;
            a.g = a.g + base.v[1] * sky.y * k;
            a.b = a.b + base.v[2] * sky.z * k;
            a.a = a.a + base.v[3] * sky.w * k;
        }
    }
}

// legacypbr.d:948-1007, non-legacy branch
void PBRCompositor::passEmissive(int t, box2i area, Mipmap<RGBA>& diffuseMap) {
    OwnedImage<RGBAf>& accumBuffer = *_accumBuffers[t];
    for (int j = area.miny; j < area.maxy; ++j) {
        RGBAf* accumScan = accumBuffer.scanlinePtr(j - area.miny);
        for (int i = area.minx; i < area.maxx; ++i) {
            float ic = i + 0.5f, jc = j + 0.5f;
            vec4f c1 = diffuseMap.linearSample(1, ic, jc);
            vec4f c2 = diffuseMap.linearSample(2, ic, jc);
            vec4f c3 = diffuseMap.linearSample(3, ic, jc);
            vec4f c4 = diffuseMap.cubicSample(4, ic, jc);
            vec4f c5 = diffuseMap.cubicSample(5, ic, jc);
            float noise = (BLUE_NOISE_16x16[(i & 15) * 16 + (j & 15)] - 127.5f) * 0.003f;
            const float AMT = 0.002f * 0.67f;
            vec4f em = c1 * AMT;
            em = em + c2 * AMT;
            em = em + c3 * AMT;
            em = em + c4 * AMT;
            em = em + c5 * AMT * (1 + noise);
            RGBAf& a = accumScan[i - area.minx];
            a.r += em.x; a.g += em.y; a.b += em.z; a.a += em.w;
        }
    }
}

// legacypbr.d:1057-1107, non-legacy branch
void PBRCompositor::passClamp(int t, box2i area, ImageRef<RGBA>& wfb) {
    OwnedImage<RGBAf>& accumBuffer = *_accumBuffers[t];
    const __m128 mm255_99 = _mm_set1_ps(255.99f);
    const __m128i zero = _mm_setzero_si128();
    const float toneRatio = params.tonemapRatio / 3;
    for (int j = area.miny; j < area.maxy; ++j) {
        int32_t* wfbScan = (int32_t*)wfb.scanline(j);
        const RGBAf* accumScan = accumBuffer.scanlinePtr(j - area.miny);
        for (int i = area.minx; i < area.maxx; ++i) {
            RGBAf acc = accumScan[i - area.minx];
            __m128 color = _mm_setr_ps(acc.r, acc.g, acc.b, 1.0f);
            __m128 exceed = _mm_max_ps(_mm_setzero_ps(), _mm_sub_ps(color, _mm_set1_ps(params.tonemapThreshold)));
            alignas(16) float ex[4];
            _mm_store_ps(ex, exceed);
            float exceedLuma = 0.212655f * ex[0] + 0.715158f * ex[1] + 0.072187f * ex[2];
            color = _mm_add_ps(color, _mm_set1_ps(exceedLuma * toneRatio));
            alignas(16) float cv[4];
            _mm_store_ps(cv, color);
            cv[3] = 1.0f;
            color = _mm_load_ps(cv);
            __m128i d = _mm_cvttps_epi32(_mm_mul_ps(color, mm255_99));
            __m128i wv = _mm_packs_epi32(d, zero);
            __m128i bv = _mm_packus_epi16(wv, zero);
            wfbScan[i] = _mm_cvtsi128_si32(bv);
        }
    }
}

void PBRCompositor::compositeOneTile(int t, box2i area, ImageRef<RGBA>& wfb, Mipmap<RGBA>& diffuseMap,
                                     Mipmap<RGBA>& materialMap, Mipmap<L16>& depthMap) {
    OwnedImage<RGBAf>& accum = *_accumBuffers[t];
    for (int j = 0; j < area.height(); ++j) {  // legacypbr.d:229-237
        RGBAf* s = accum.scanlinePtr(j);
        for (int i = 0; i < area.width(); ++i) s[i] = RGBAf{0, 0, 0, 0};
    }
    const uint32_t m = params.passActiveMask;
    if (m & (1u << PASS_NORMAL)) passNormal(t, area, depthMap);
    if (m & (1u << PASS_AO)) passAO(t, area, diffuseMap, depthMap);
    if (m & (1u << PASS_OBLIQUE_SHADOW)) passObliqueShadow(t, area, diffuseMap, depthMap);
    if (m & (1u << PASS_DIRECTIONAL)) passDirectional(t, area, diffuseMap, materialMap);
    if (m & (1u << PASS_SPECULAR)) passSpecular(t, area, diffuseMap, materialMap);
    if (m & (1u << PASS_SKYBOX)) passSkybox(t, area, diffuseMap, materialMap);
    if (m & (1u << PASS_EMISSIVE)) passEmissive(t, area, diffuseMap);
    if (m & (1u << PASS_CLAMP)) passClamp(t, area, wfb);
}

// ThreadPool.parallelFor (core/dplug/core/thread.d:431-520): a PERSISTENT pool — the reference creates its workers once
// (thread.d:392-405) and wakes them per call; work items are handed out through a shared atomic counter
// (thread.d:586-601).  One pool per thread count, created on first use and kept for the life of the process.
namespace {
class Pool {
public:
    explicit Pool(int n) : n_(n) {
        for (int t = 0; t < n; ++t) workers_.emplace_back([this, t] { run(t); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; ++gen_; }
        wake_.notify_all();
        for (auto& w : workers_) w.join();
    }
    void parallelFor(int count, void (*fn)(int, int, void*), void* user) {
        std::unique_lock<std::mutex> g(m_);
        fn_ = fn; user_ = user; count_ = count; next_.store(0); pending_ = n_; ++gen_;
        wake_.notify_all();
        done_.wait(g, [this] { return pending_ == 0; });
    }
private:
    void run(int t) {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> g(m_);
                wake_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
            }
            for (;;) {
                const int i = next_.fetch_add(1);
                if (i >= count_) break;
                fn_(i, t, user_);
            }
            std::lock_guard<std::mutex> g(m_);
            if (--pending_ == 0) done_.notify_one();
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable wake_, done_;
    uint64_t gen_ = 0;
    bool stop_ = false;
    void (*fn_)(int, int, void*) = nullptr;
    void* user_ = nullptr;
    int count_ = 0, pending_ = 0;
    std::atomic<int> next_{0};
};
std::mutex g_poolsLock;
std::vector<std::pair<int, Pool*>> g_pools;   // leaked on purpose: workers must outlive static destruction order
Pool* poolFor(int n) {
    std::lock_guard<std::mutex> g(g_poolsLock);
    for (auto& p : g_pools)
        if (p.first == n) return p.second;
    g_pools.emplace_back(n, new Pool(n));
    return g_pools.back().second;
}
std::mutex g_callLock;   // one parallelFor at a time per process (the reference's pool is not re-entrant either)
}  // namespace

void parallelFor(int numThreads, int count, void (*fn)(int, int, void*), void* user) {
    if (count <= 0) return;
    if (count == 1 || numThreads <= 1) {  // thread.d:484-489: inline on threadIndex 0
        for (int i = 0; i < count; ++i) fn(i, 0, user);
        return;
    }
    std::lock_guard<std::mutex> g(g_callLock);
    poolFor(numThreads)->parallelFor(count, fn, user);
}

void PBRCompositor::compositeTile(ImageRef<RGBA> wfb, const box2i* areas, int numAreas, Mipmap<RGBA>& diffuseMap,
                                  Mipmap<RGBA>& materialMap, Mipmap<L16>& depthMap) {
    struct Ctx { PBRCompositor* self; ImageRef<RGBA>* wfb; const box2i* areas; Mipmap<RGBA>* d; Mipmap<RGBA>* m; Mipmap<L16>* z; };
    Ctx c{this, &wfb, areas, &diffuseMap, &materialMap, &depthMap};
    parallelFor(_numThreads, numAreas,
                [](int i, int t, void* u) {
                    Ctx* c = (Ctx*)u;
                    c->self->compositeOneTile(t, c->areas[i], *c->wfb, *c->d, *c->m, *c->z);
                },
                &c);
}

// ---------------------------------------------------------------- GUIGraphics slice
RefGraphics::RefGraphics(int numThreads) : compositor(numThreads) {}
RefGraphics::~RefGraphics() {}

// graphics.d:1115-1137
void RefGraphics::resize(int W_, int H_) {
    W = W_; H = H_;
    compositor.resizeBuffers(W, H, 64, 64);
    diffuseMap.size(5, W, H);
    depthMap.size(4, W, H);
    depthMap.levels[0]->size(W, H, 1, 1, 1, 2);
    materialMap.size(0, W, H);
    compositedBuffer.size(W, H, 0, 16, 1, 0);
}

// graphics.d:1354-1423
void RefGraphics::regenerateMipmaps(const box2i* rects, int n) {
    if (n == 0) return;
    std::vector<box2i> scratch0(rects, rects + n), scratch1(rects, rects + n);
    for (int level = 1; level < diffuseMap.numLevels(); ++level) {
        int q = (level == 1) ? Q_BOX_ALPHACOV_PREMUL : Q_CUBIC;
        for (auto& area : scratch0) area = diffuseMap.generateNextLevel(q, area, level);
    }
    for (auto& area : scratch1) depthMap.levels[0]->replicateBordersTouching(area);
    for (int level = 1; level < depthMap.numLevels(); ++level) {
        int q = level >= 3 ? Q_CUBIC : Q_BOX;
        for (auto& area : scratch1) area = depthMap.generateNextLevel(q, area, level);
    }
}

// graphics.d:1338-1349
void RefGraphics::compositeGUI(const box2i* rects, int n) {
    std::vector<box2i> tiles;
    tileAreas(rects, n, 64, 64, tiles);
    compositor.compositeTile(compositedBuffer.toRef(), tiles.data(), (int)tiles.size(), diffuseMap, materialMap, depthMap);
}

void RefGraphics::fullFrame() {
    box2i whole{0, 0, W, H};
    regenerateMipmaps(&whole, 1);
    compositeGUI(&whole, 1);
}

}  // namespace b2ref
