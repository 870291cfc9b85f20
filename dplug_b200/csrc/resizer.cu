// ImageResizer on the device — SURVEY §8(f) row 2, second half: the resampling that fills a background widget's cache
// (PBRBackgroundGUI.resizeFun, pbrwidgets/dplug/pbrwidgets/pbrbackgroundgui.d:172-199) before it is blitted into the
// level-0 maps.  Semantics: graphics/dplug/graphics/resizer.d over the IN-TREE stb_image_resize port
// (graphics/dplug/graphics/stb_image_resize.d, the `DPLUG_USE_STB_IMAGE_RESIZE_V2 = false` branches): separable
// resampling in float, filter per call (Magic Kernel Sharp 2013 at 86 % for diffuse, Catmull-Rom up / Mitchell down for
// material, Magic Kernel Sharp 2021 for depth), clamp edges, linear colour space, no alpha weighting.
//
// The port walks scanlines through a ring buffer and, when down-sampling, SCATTERS every input sample into the outputs
// it touches.  Every output sample is nevertheless a plain sum `((0 + s0*c0) + s1*c1) + ...` over its sources in
// ascending source order, with single IEEE multiplies and adds — so the device runs two GATHER passes over tables the
// host derives from the port's contributor / coefficient arithmetic (set-up work, a few thousand floats):
//   k_resize_h: decode (v / 255, v / 65535: IEEE division) + horizontal sums, one thread per (output column, input row);
//   k_resize_v: vertical sums + encode (saturate, * max, + 0.5, truncate), one thread per output pixel.
// Bit-exact against a CPU restatement of the port that keeps its ring buffer and scatter (tests/test_resizer.py).  The
// default build of the reference calls the un-vendored stb_image_resize2 package instead: same kernels, different float
// operation order — parity with that build is unpinned.
#include <cmath>
#include <cstring>
#include <vector>
#include "common.cuh"
#include "../../include/b2pbr.h"

namespace b2 {

// ------------------------------------------------------------------------------------------------ host: filter tables
namespace {

enum Filter { F_DEFAULT = 0, F_CUBICBSPLINE = 3, F_CATMULLROM = 4, F_MITCHELL = 5, F_MKS_2013_86 = 11, F_MKS_2021 = 13 };   // stb_image_resize.d:270-286

// filter kernels (stb_image_resize.d:659-774); mixed float / double exactly as the D expressions promote
float kCubic(float x) { x = fabsf(x); return x < 1.0f ? (4 + x * x * (3 * x - 6)) / 6 : (x < 2.0f ? (8 + x * (-12 + x * (6 - x))) / 6 : 0.0f); }
float kCatmullRom(float x) { x = fabsf(x); return x < 1.0f ? 1 - x * x * (2.5f - 1.5f * x) : (x < 2.0f ? 2 - x * (4 + x * (0.5f * x - 2.5f)) : 0.0f); }
float kMitchell(float x) { x = fabsf(x); return x < 1.0f ? (16 + x * x * (21 * x - 36)) / 18 : (x < 2.0f ? (32 + x * (-60 + x * (36 - 7 * x))) / 18 : 0.0f); }
float kMagic2013(float x) {
    x = fabsf(x);
    if (x < 0.5) return (float)(0.75 - x * x);
    if (x < 1.5) return (float)(0.5 * (x - 1.5) * (x - 1.5));
    return 0.0f;
}
float kMagicSharp2013(float x) {
    x = fabsf(x);
    if (x <= 1.17549435e-38f) return 17.0f / 16.0f;
    if (x < 0.5) return (float)(17.0 / 16.0 - 7.0 * x * x / 4.0);
    if (x < 1.5) { const double x2 = x * x; return (float)(0.25 * (4 * x2 - 11.0 * x + 7.0)); }
    if (x < 2.5) return (float)(-0.125 * (x - 5.0 / 2.0) * (x - 5.0 / 2.0));
    return 0.0f;
}
float kMagicSharp2021(float x) {
    x = fabsf(x);
    const float x2 = x * x;
    if (x < 0.5) return 577.0f / 576.0f - (239.0f / 144.0f) * x2;
    if (x < 1.5) return (140 * x2 - 379 * x + 239) / 144.0f;
    if (x < 2.5) return -(24 * x2 - 113 * x + 130) / 144.0f;
    if (x < 3.5) return (4 * x2 - 27 * x + 45) / 144.0f;
    if (x < 4.5) return -(4 * x2 - 36 * x + 81) / 1152.0f;
    return 0.0f;
}
float evalKernel(int f, float x) {
    switch (f) {
        case F_CUBICBSPLINE: return kCubic(x);
        case F_CATMULLROM: return kCatmullRom(x);
        case F_MITCHELL: return kMitchell(x);
        case F_MKS_2013_86: return 0.14f * kMagic2013(x) + 0.86f * kMagicSharp2013(x);   // stb_image_resize.d:723-728
        default: return kMagicSharp2021(x);
    }
}
float support(int f) { return (f == F_CUBICBSPLINE || f == F_CATMULLROM || f == F_MITCHELL) ? 2.0f : (f == F_MKS_2013_86 ? 3.0f : 5.0f); }   // :806-822

// One axis as a gather table: output sample o = sum over e in [start[o], start[o+1]) of source[src[e]] * coef[e], added
// in that order starting from 0.  `src` may lie outside the image (margins): the kernels clamp it (STBIR_EDGE_CLAMP).
struct AxisTable {
    std::vector<int> start, src;
    std::vector<float> coef;
};

// stbir__calculate_filters (stb_image_resize.d:1153-1190) with the sample ranges (:967-994), the coefficient groups
// (:996-1074) and the down-sampling normalisation (:1076-1149), then turned inside out into the gather table.
void buildAxis(int filter, int inSize, int outSize, AxisTable& T) {
    const float scale = (float)outSize / inSize;              // stbir__calculate_transform, :2270-2292 (whole image: no shift)
    const bool up = scale > 1;                                // stbir__use_upsampling, :825-828
    if (filter == F_DEFAULT) filter = up ? F_CATMULLROM : F_MITCHELL;   // stbir__choose_filter, :2294-2302
    const float S = support(filter);
    const int width = (int)ceilf(S * 2);                      // stbir__get_coefficient_width, :860-866
    const int margin = up ? (int)ceilf(S * 2) / 2 : (int)ceilf(S * 2 / scale) / 2;   // stbir__get_filter_pixel_margin, :842-858
    const int total = up ? outSize : inSize + margin * 2;     // stbir__get_contributors, :868-874
    std::vector<int> n0(total), n1(total);
    std::vector<float> g((size_t)total * width + 16, 0.0f);  // groups are contiguous, as in the port: a group of width + 1 runs into the next one
    T.start.assign(outSize + 1, 0);
    T.src.clear(); T.coef.clear();
    if (up) {
        const float radius = S * scale;
        for (int n = 0; n < total; ++n) {
            const float centre = (float)n + 0.5f;
            const float lo = (centre - radius) / scale, hi = (centre + radius) / scale, inCentre = centre / scale;
            int first = (int)floor(lo + 0.5), last = (int)floor(hi - 0.5);
            float* c = &g[(size_t)n * width];
            float sum = 0;
            n0[n] = first; n1[n] = last;
            for (int i = 0; i <= last - first; ++i) {
                c[i] = evalKernel(filter, inCentre - ((float)(i + first) + 0.5f));
                if (i == 0 && !c[i]) { n0[n] = ++first; --i; continue; }   // a leading zero weight moves the window
                sum += c[i];
            }
            const float norm = 1 / sum;
            for (int i = 0; i <= last - first; ++i) c[i] *= norm;
            for (int i = last - first; i >= 0 && !c[i]; --i) n1[n] = n0[n] + i - 1;
        }
        for (int o = 0; o < outSize; ++o) {                   // an output's group is its gather list
            T.start[o] = (int)T.src.size();
            for (int k = n0[o]; k <= n1[o]; ++k) { T.src.push_back(k); T.coef.push_back(g[(size_t)o * width + (k - n0[o])]); }
        }
        T.start[outSize] = (int)T.src.size();
        return;
    }
    const float radius = S / scale;
    for (int n = 0; n < total; ++n) {
        const float centre = (float)(n - margin) + 0.5f;
        const float lo = (centre - radius) * scale, hi = (centre + radius) * scale, outCentre = centre * scale;
        const int first = (int)floor(lo + 0.5), last = (int)floor(hi - 0.5);
        float* c = &g[(size_t)n * width];
        n0[n] = first; n1[n] = last;
        for (int i = 0; i <= last - first; ++i) c[i] = evalKernel(filter, ((float)(i + first) + 0.5f) - outCentre) * scale;
        for (int i = last - first; i >= 0 && !c[i]; --i) n1[n] = n0[n] + i - 1;
    }
    for (int o = 0; o < outSize; ++o) {                       // every output's weights sum to 1
        float sum = 0;
        for (int j = 0; j < total; ++j) {
            if (o >= n0[j] && o <= n1[j]) sum += g[(size_t)j * width + (o - n0[j])];
            else if (o < n0[j]) break;
        }
        const float norm = 1 / sum;
        for (int j = 0; j < total; ++j) {
            if (o >= n0[j] && o <= n1[j]) g[(size_t)j * width + (o - n0[j])] *= norm;
            else if (o < n0[j]) break;
        }
    }
    for (int j = 0; j < total; ++j) {                         // drop leading zero weights and outputs left of the image
        float* c = &g[(size_t)j * width];
        int skip = 0;
        while (skip < width && c[skip] == 0) skip++;
        n0[j] += skip;
        while (n0[j] < 0) { n0[j]++; skip++; }
        const int range = n1[j] - n0[j] + 1, mx = width < range ? width : range;
        for (int i = 0; i < mx && i + skip < width; ++i) c[i] = c[i + skip];
    }
    for (int j = 0; j < total; ++j)
        if (n1[j] > outSize - 1) n1[j] = outSize - 1;
    // inside out: input sample j - margin contributes to outputs n0[j] .. n1[j]; ascending j = ascending source order
    std::vector<int> count(outSize + 1, 0);
    for (int j = 0; j < total; ++j)
        for (int k = n0[j]; k <= n1[j]; ++k) count[k + 1]++;
    for (int o = 0; o < outSize; ++o) count[o + 1] += count[o];
    T.start = count;
    T.src.assign(count[outSize], 0); T.coef.assign(count[outSize], 0.0f);
    std::vector<int> fill(count.begin(), count.end() - 1);
    for (int j = 0; j < total; ++j)
        for (int k = n0[j]; k <= n1[j]; ++k) { const int e = fill[k]++; T.src[e] = j - margin; T.coef[e] = g[(size_t)j * width + (k - n0[j])]; }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ device
template <typename T, int CH>
__global__ void __launch_bounds__(256) k_resize_h(float* __restrict__ tmp, const uint8_t* __restrict__ in, size_t inPitch, int inW, int inH, int outW,
                                                  const int* __restrict__ start, const int* __restrict__ src, const float* __restrict__ coef) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= outW || y >= inH) return;
    const T* row = reinterpret_cast<const T*>(in + (size_t)y * inPitch);
    const float maxv = sizeof(T) == 1 ? 255.0f : 65535.0f;
    float acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = 0.0f;
    for (int e = __ldg(start + x), e1 = __ldg(start + x + 1); e < e1; ++e) {
        const int s = max(0, min(__ldg(src + e), inW - 1));   // stbir__edge_wrap, STBIR_EDGE_CLAMP
        const float w = __ldg(coef + e);
#pragma unroll
        for (int c = 0; c < CH; ++c)
            acc[c] = __fadd_rn(acc[c], __fmul_rn(__fdiv_rn((float)row[s * CH + c], maxv), w));   // decode (:1236-1316), then out += in * coefficient (:1457-1679)
    }
    float* o = tmp + ((size_t)y * outW + x) * CH;
#pragma unroll
    for (int c = 0; c < CH; ++c) o[c] = acc[c];
}

template <typename T, int CH>
__global__ void __launch_bounds__(256) k_resize_v(uint8_t* __restrict__ out, size_t outPitch, const float* __restrict__ tmp, int inH, int outW, int outH,
                                                  const int* __restrict__ start, const int* __restrict__ src, const float* __restrict__ coef) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= outW || y >= outH) return;
    const float maxv = sizeof(T) == 1 ? 255.0f : 65535.0f;
    float acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = 0.0f;
    for (int e = __ldg(start + y), e1 = __ldg(start + y + 1); e < e1; ++e) {
        const int r = max(0, min(__ldg(src + e), inH - 1));
        const float w = __ldg(coef + e);
        const float* t = tmp + ((size_t)r * outW + x) * CH;
#pragma unroll
        for (int c = 0; c < CH; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(t[c], w));   // :1926-2123
    }
    T* o = reinterpret_cast<T*>(out + (size_t)y * outPitch) + x * CH;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        float f = acc[c];
        f = f < 0 ? 0 : (f > 1 ? 1 : f);                                               // stbir__saturate, :484-493
        o[c] = (T)(int)__fadd_rn(__fmul_rn(f, maxv), 0.5f);                            // STBIR__ENCODE_LINEAR8 / 16, :1770-1785
    }
}

// dIn / dOut are device images; enqueued on `s`; the temporaries are released after the stream has run them
static cudaError_t resizeOnDevice(int kind, const uint8_t* dIn, size_t inPitch, int inW, int inH, uint8_t* dOut, size_t outPitch, int outW, int outH, cudaStream_t s) {
    const bool l16 = kind == B2_RESIZE_DEPTH || kind == B2_RESIZE_GENERIC_L16;
    const size_t texel = l16 ? 2 : 4;
    if (inW == outW && inH == outH)                           // sameSizeResize, resizer.d:453-465
        return cudaMemcpy2DAsync(dOut, outPitch, dIn, inPitch, (size_t)inW * texel, (size_t)inH, cudaMemcpyDeviceToDevice, s);
    static const int filterOf[5] = {F_MKS_2013_86, F_DEFAULT, F_MKS_2021, F_DEFAULT, F_CUBICBSPLINE};   // resizer.d:346-372, 39-64, 171-197, 144-169, 92-117 (v1 branches)
    AxisTable H, V;
    buildAxis(filterOf[kind], inW, outW, H);
    buildAxis(filterOf[kind], inH, outH, V);
    const int ch = l16 ? 1 : 4;
    const size_t nH = H.src.size(), nV = V.src.size();
    const size_t ints = (size_t)(outW + 1) + nH + (size_t)(outH + 1) + nV;
    const size_t tabBytes = ints * 4 + (nH + nV) * 4, tmpBytes = (size_t)inH * outW * ch * sizeof(float);
    std::vector<uint8_t> host(tabBytes);
    int* hi = (int*)host.data();
    memcpy(hi, H.start.data(), (size_t)(outW + 1) * 4); hi += outW + 1;
    memcpy(hi, H.src.data(), nH * 4); hi += nH;
    memcpy(hi, V.start.data(), (size_t)(outH + 1) * 4); hi += outH + 1;
    memcpy(hi, V.src.data(), nV * 4); hi += nV;
    memcpy(hi, H.coef.data(), nH * 4); hi += nH;
    memcpy(hi, V.coef.data(), nV * 4);
    uint8_t* dTab = nullptr;
    float* dTmp = nullptr;
    cudaError_t e = cudaMalloc((void**)&dTab, tabBytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&dTmp, tmpBytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dTab, host.data(), tabBytes, cudaMemcpyHostToDevice, s);   // pageable: staged before the call returns
    if (e == cudaSuccess) {
        const int* dHs = (const int*)dTab; const int* dHsrc = dHs + outW + 1;
        const int* dVs = dHsrc + nH; const int* dVsrc = dVs + outH + 1;
        const float* dHc = (const float*)(dVsrc + nV); const float* dVc = dHc + nH;
        const dim3 block(256), gh((outW + 255) / 256, inH), gv((outW + 255) / 256, outH);
        if (l16) {
            k_resize_h<uint16_t, 1><<<gh, block, 0, s>>>(dTmp, dIn, inPitch, inW, inH, outW, dHs, dHsrc, dHc);
            k_resize_v<uint16_t, 1><<<gv, block, 0, s>>>(dOut, outPitch, dTmp, inH, outW, outH, dVs, dVsrc, dVc);
        } else {
            k_resize_h<uint8_t, 4><<<gh, block, 0, s>>>(dTmp, dIn, inPitch, inW, inH, outW, dHs, dHsrc, dHc);
            k_resize_v<uint8_t, 4><<<gv, block, 0, s>>>(dOut, outPitch, dTmp, inH, outW, outH, dVs, dVsrc, dVc);
        }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (dTab) cudaFree(dTab);
    if (dTmp) cudaFree(dTmp);
    return e;
}

// host-only test hook: the gather table of one axis (b2_resize_axis_table)
int resize_axis_table(int kind, int inSize, int outSize, int* start, int* src, float* coef, int cap) {
    static const int filterOf[5] = {F_MKS_2013_86, F_DEFAULT, F_MKS_2021, F_DEFAULT, F_CUBICBSPLINE};
    AxisTable T;
    buildAxis(filterOf[kind], inSize, outSize, T);
    const int n = (int)T.src.size();
    if (start) memcpy(start, T.start.data(), sizeof(int) * (size_t)(outSize + 1));
    for (int e = 0; e < n && e < cap; ++e) {
        if (src) src[e] = T.src[e];
        if (coef) coef[e] = T.coef[e];
    }
    return n;
}

// host image -> device image `dOut` (outPitch bytes per row) of outW x outH; synchronises
int resize_host_to_device(int kind, const void* in, int inW, int inH, size_t inPitch, uint8_t* dOut, size_t outPitch, int outW, int outH, cudaStream_t s) {
    const bool l16 = kind == B2_RESIZE_DEPTH || kind == B2_RESIZE_GENERIC_L16;
    const size_t rowBytes = (size_t)inW * (l16 ? 2 : 4);
    uint8_t* dIn = nullptr;
    size_t dPitch = (rowBytes + 127) & ~(size_t)127;
    cudaError_t e = cudaMalloc((void**)&dIn, dPitch * inH);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(dIn, dPitch, in, inPitch, rowBytes, (size_t)inH, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = resizeOnDevice(kind, dIn, dPitch, inW, inH, dOut, outPitch, outW, outH, s);
    if (dIn) cudaFree(dIn);
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace b2
