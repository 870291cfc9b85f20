// Shared device/host structures of libb2pbr (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace b2 {

// One level of a device-resident pyramid.  Coordinates handed to kernels are GLOBAL
// (whole-canvas) level coordinates; `y0` is the global row stored at `base`, so a
// row-band object simply has y0 != 0 and rows < h.  Pitch is a multiple of 128 B and
// `base` is 256-B aligned, so absolute-x-aligned 128-bit accesses are legal.
struct DevLevel {
    uint8_t* base;
    int pitch;  // bytes
    int w, h;   // global level size
    int y0;     // global row index of stored row 0
    int rows;   // stored rows
};

template <typename T>
__device__ __forceinline__ const T* rowPtr(const DevLevel& L, int y) {
    return reinterpret_cast<const T*>(L.base + (size_t)(y - L.y0) * (size_t)L.pitch);
}
template <typename T>
__device__ __forceinline__ T* rowPtrW(const DevLevel& L, int y) {
    return reinterpret_cast<T*>(L.base + (size_t)(y - L.y0) * (size_t)L.pitch);
}

struct Box { int minx, miny, maxx, maxy; };

// One launch of the fused mipmap chain (mip_fused.cu): up to three consecutive generateLevel calls.
struct FusedArgs {
    DevLevel lv[4];   // source level s, then s+1 .. s+n
    Box upd[3];       // update rectangle of level s+1+k (generateLevel's `updateRect`), possibly empty
    int q[3];         // quality of the step producing level s+1+k (B2_QUALITY_*)
    int n;            // steps in this launch, 1..3
    int tile0x, tile0y, tilesX;   // first tile (in tiles of the last level) and tiles per grid row
    unsigned long long negZero2;  // (-0.0f, -0.0f): see mip_math.cuh, mulToAdd2
};

constexpr int kMaxDiffuseLevels = 6;   // Mipmap(5): gui/dplug/gui/graphics.d:1118
constexpr int kMaxDepthLevels = 5;     // Mipmap(4): graphics.d:1119
constexpr int kMaxSkyLevels = 13;      // Mipmap(12): gui/dplug/gui/legacypbr.d:868

// Everything the lighting kernel needs, passed by value (__grid_constant__).
struct PbrArgs {
    DevLevel diffuse[kMaxDiffuseLevels];
    DevLevel depth[kMaxDepthLevels];
    DevLevel material;
    DevLevel sky[kMaxSkyLevels];
    DevLevel out;
    int nDiffuse, nDepth, nSky;
    int W, H;                 // level-0 size (global)
    const Box* boxes;         // work list
    int nBoxes;
    const uint32_t* rsqrtTab; // 2048 entries
    const float* expTab;      // 256 entries, PassSpecularLight._exponentTable
    // parameters
    float light1[3], light2Dir[3], light2Col[3], light3Dir[3], light3Col[3];
    float ambient, skyAmount, toneThr, toneRatio;
    uint32_t passMask;
    // precomputed constants (host side, see pbr_constants())
    float shadowW[11];
    float shadowInvTotal;
    float invW, invH;
    float skyPixels, skyFX, skyFY;
    // k_pbr_tex keeps base colour and material in byte units; these copies carry the folded 1/255 factors:
    // light1S = light1 / 255, light2ColS = light2Col / 255, light3ColS = light3Col / 255^2 (base colour and specular
    // amount), ambientS = ambient / 255, skyAmountS = skyAmount / 255^2 (base colour and metalness)
    float light1S[3], light2ColS[3], light3ColS[3], ambientS, skyAmountS;
    // texture-unit path (pbr_tex.cu): every pyramid level is also a pitch-2D texture (linear filter, clamp, unnormalised
    // coordinates, normalised-float reads); the skybox levels are CUDA arrays read with texture gather
    cudaTextureObject_t texD[kMaxDiffuseLevels];   // diffuse levels (1..3 are sampled)
    cudaTextureObject_t texZ[kMaxDepthLevels];     // depth levels (1..4 are sampled)
    // all skybox levels in ONE CUDA array (an atlas: every level with a one-texel replicated border around it), so that
    // the texture handle of the gathers is warp-uniform while the level varies per pixel
    cudaTextureObject_t texSky;
    cudaTextureObject_t texSkyLin;   // the same array with the linear filter (B2_SKY_HW experiment)
    float skyMaxX[kMaxSkyLevels], skyMaxY[kMaxSkyLevels];   // w - 1, h - 1 of every skybox level
    float skyOx[kMaxSkyLevels], skyOy[kMaxSkyLevels];       // atlas position of texel (0, 0) of every level
    // the filtered skybox fetch of level l samples the atlas at clamp(fma(sky, 2^-l, o), lo, hi) per axis, with
    // lo = o + 0.5 and hi = o + (size - 1) + 0.5 (texel centres: the reference's edge rules): {o, lo, hi, 2^-l}, one
    // 128-bit constant load per level and axis
    float4 skyCX[kMaxSkyLevels], skyCY[kMaxSkyLevels];
    unsigned long long negZero2;   // (-0.0f, -0.0f): the addend of exact packed products (mip_math.cuh, mulToAdd2)
};

}  // namespace b2
