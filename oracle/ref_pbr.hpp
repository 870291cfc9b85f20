// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_core.hpp header).
//
// CPU restatement of gui/dplug/gui/legacypbr.d (PBRCompositor + its 8 passes),
// gui/dplug/gui/ransac.d (plane-fit normal) and the parts of
// gui/dplug/gui/compositor.d / graphics.d that drive them (tile loop,
// regenerateMipmaps, compositeGUI, buffer sizing).
#pragma once
#include "ref_mipmap.hpp"

namespace b2ref {

enum {
    PASS_NORMAL = 0, PASS_AO = 1, PASS_OBLIQUE_SHADOW = 2, PASS_DIRECTIONAL = 3,
    PASS_SPECULAR = 4, PASS_SKYBOX = 5, PASS_EMISSIVE = 6, PASS_CLAMP = 7, NUM_PASSES = 8
};  // legacypbr.d:129-139

struct PBRParams {
    float light1Color[3];  // PassObliqueShadowLight.color      legacypbr.d:453
    float light2Dir[3];    // PassDirectionalLight.direction    legacypbr.d:626
    float light2Color[3];  //                                   legacypbr.d:629
    float light3Dir[3];    // PassSpecularLight.direction       legacypbr.d:676
    float light3Color[3];  //                                   legacypbr.d:679
    float ambientLight;    // PassAmbientOcclusion.amount       legacypbr.d:391
    float skyboxAmount;    // PassSkyboxReflections.amount      legacypbr.d:843
    float tonemapThreshold;  // legacypbr.d:1045
    float tonemapRatio;      // legacypbr.d:1049
    uint32_t passActiveMask; // CompositorPass._active, compositor.d:222-236
};
void defaultParams(PBRParams* p);
void planeFittingNormalPublic(const float depth9[9], float out3[3]);  // ransac.d:53-121

class PBRCompositor {
public:
    explicit PBRCompositor(int numThreads);
    ~PBRCompositor();

    PBRParams params;

    // legacypbr.d:180-198, 706-718
    void resizeBuffers(int width, int height, int areaMaxWidth, int areaMaxHeight);
    // legacypbr.d:861-870 — copies the pixels (the reference takes ownership of the image)
    void setSkybox(const RGBA* pixels, int w, int h, size_t pitch);
    // legacypbr.d:201-260
    void compositeTile(ImageRef<RGBA> wfb, const box2i* areas, int numAreas, Mipmap<RGBA>& diffuseMap,
                       Mipmap<RGBA>& materialMap, Mipmap<L16>& depthMap);
    Mipmap<RGBA>* skybox() { return _skybox; }

private:
    int _numThreads;
    std::vector<OwnedImage<RGBf>*> _normalBuffers;     // legacypbr.d:48
    std::vector<OwnedImage<RGBAf>*> _accumBuffers;     // legacypbr.d:51
    std::vector<OwnedImage<float>*> _varianceBuffers;  // legacypbr.d:54 (L32f)
    std::vector<std::vector<float>> _specularFactor, _exponentFactor, _toksvigScaleFactor;
    float _exponentTable[256];
    Mipmap<RGBA>* _skybox = nullptr;

    void compositeOneTile(int threadIndex, box2i area, ImageRef<RGBA>& wfb, Mipmap<RGBA>& diffuseMap,
                          Mipmap<RGBA>& materialMap, Mipmap<L16>& depthMap);
    void passNormal(int t, box2i area, Mipmap<L16>& depthMap);
    void passAO(int t, box2i area, Mipmap<RGBA>& diffuseMap, Mipmap<L16>& depthMap);
    void passObliqueShadow(int t, box2i area, Mipmap<RGBA>& diffuseMap, Mipmap<L16>& depthMap);
    void passDirectional(int t, box2i area, Mipmap<RGBA>& diffuseMap, Mipmap<RGBA>& materialMap);
    void passSpecular(int t, box2i area, Mipmap<RGBA>& diffuseMap, Mipmap<RGBA>& materialMap);
    void passSkybox(int t, box2i area, Mipmap<RGBA>& diffuseMap, Mipmap<RGBA>& materialMap);
    void passEmissive(int t, box2i area, Mipmap<RGBA>& diffuseMap);
    void passClamp(int t, box2i area, ImageRef<RGBA>& wfb);
};

// The slice of GUIGraphics that feeds the compositor: buffer sizing (graphics.d:1115-1137),
// regenerateMipmaps (graphics.d:1354-1423) and compositeGUI (graphics.d:1338-1349).
class RefGraphics {
public:
    explicit RefGraphics(int numThreads);
    ~RefGraphics();
    void resize(int W, int H);
    void regenerateMipmaps(const box2i* rects, int n);
    void compositeGUI(const box2i* rectsToCompositeDisjointed, int n);
    // full-frame convenience used by the CPU baseline: one dirty rect = whole UI
    void fullFrame();

    int W = 0, H = 0;
    PBRCompositor compositor;
    Mipmap<RGBA> diffuseMap, materialMap;
    Mipmap<L16> depthMap;
    OwnedImage<RGBA> compositedBuffer;
};

// threadPool.parallelFor stand-in (core/dplug/core/thread.d:478-496): blocks until all
// items ran; count==1 runs inline on threadIndex 0.
void parallelFor(int numThreads, int count, void (*fn)(int i, int threadIndex, void* user), void* user);

}  // namespace b2ref
