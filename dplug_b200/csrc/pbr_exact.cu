// Lighting kernel, exact-arithmetic variant: one thread per pixel evaluates the eight passes of
// gui/dplug/gui/legacypbr.d:269-1108 in registers with the reference's IEEE operation order (compiled
// with -fmad=false), so its output equals the CPU restatement bit for bit.  The per-tile scratch images
// of the reference (normal RGBf, variance L32f, accum RGBAf — legacypbr.d:41-55) never exist on the
// device.  It is the device-side yardstick for the throughput kernel in pbr_tex.cu.
#include "pbr_math.cuh"

namespace b2 {

__device__ __forceinline__ int depthAt(const DevLevel& L, int x, int y) {
    // level-0 depth has a 1-px replicated border in the reference (gui/dplug/gui/graphics.d:1121-1126,
    // graphics/dplug/graphics/image.d:493-596); on the device that is clamp-to-edge addressing.
    x = max(0, min(x, L.w - 1));
    y = max(0, min(y, L.h - 1));
    return (int)__ldg(rowPtr<uint16_t>(L, y) + x);
}

__device__ __forceinline__ float dot4(F4 a, F4 b) {  // gui/dplug/gui/ransac.d:14-18
    float m0 = a.x * b.x, m1 = a.y * b.y, m2 = a.z * b.z, m3 = a.w * b.w;
    return m0 + m1 + m2 + m3;
}
__device__ __forceinline__ F4 fastNormalize(const uint32_t* tab, F4 v) {  // ransac.d:39-46
    float s0 = v.x * v.x, s1 = v.y * v.y, s2 = v.z * v.z, s3 = v.w * v.w;
    float inv = rsqrtss_x86(tab, s0 + s1 + s2 + s3);
    return F4{v.x * inv, v.y * inv, v.z * inv, v.w * inv};
}

__device__ __forceinline__ uint32_t packOutput(float r, float g, float b, float a) {
    // _mm_cvttps_epi32 -> _mm_packs_epi32 -> _mm_packus_epi16 (legacypbr.d:1101-1103)
    float v[4] = {r, g, b, a};
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float f = v[k] * 255.99f;
        int i = (f >= -2147483648.0f && f < 2147483648.0f) ? __float2int_rz(f) : (int)0x80000000;
        i = max(-32768, min(32767, i));
        i = max(0, min(255, i));
        out |= (uint32_t)i << (8 * k);
    }
    return out;
}

__device__ __forceinline__ uint32_t shadePixelExact(const PbrArgs& P, int i, int j) {
    const DevLevel& D0 = P.depth[0];
    const uint32_t diffusePx = __ldg(rowPtr<uint32_t>(P.diffuse[0], j) + i);
    const uint32_t materialPx = __ldg(rowPtr<uint32_t>(P.material, j) + i);
    const float div255 = 1 / 255.0f;
    const F4 base = unpackRGBA(diffusePx) * div255;       // convertBaseColorToFloat4, legacypbr.d:1155-1163
    const F4 mat = unpackRGBA(materialPx) * div255;
    const int roughByte = materialPx & 0xff;

    // ---- PassComputeNormal (legacypbr.d:279-381) + computePlaneFittingNormal (ransac.d:53-121)
    int dz[9];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) dz[(dy + 1) * 3 + dx + 1] = depthAt(D0, i + dx, j + dy);
    F4 normal = F4{0, 0, 0, 0};
    float variance = 0;
    if (P.passMask & (1u << 0)) {
        const float multUshort = (float)(1.0 / 4655.0);
        float d[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) d[k] = (float)dz[k] * multUshort;
        const float c = 0.03660694846f, e = 0.118115527008f;
        const float wA[4] = {c, e, c, e}, wB[4] = {e, c, e, c};
        const float xzA[4] = {-c, 0.0f, c, -e}, xzB[4] = {e, -c, 0.0f, c};
        const float yzA[4] = {-c, -e, -c, 0.0f}, yzB[4] = {0.0f, c, e, c};
        float A[4] = {d[0], d[1], d[2], d[3]}, B[4] = {d[5], d[6], d[7], d[8]};
        float m[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) m[k] = A[k] * wA[k] + B[k] * wB[k];
        float filt = d[4] * 0.38111009811f + m[0] + m[1] + m[2] + m[3];
        float xz4[4], yz4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float a = A[k] - filt, b = B[k] - filt;
            xz4[k] = a * xzA[k] + b * xzB[k];
            yz4[k] = a * yzA[k] + b * yzB[k];
        }
        float xz = xz4[0] + xz4[1] + xz4[2] + xz4[3];
        float yz = yz4[0] + yz4[1] + yz4[2] + yz4[3];
        normal = fastNormalize(P.rsqrtTab, F4{-xz, yz, 0.382658847856f, 0.0f});

        float ddx[3], ddy[3];
        const float fx3[3] = {1.0f, -2.0f, 1.0f};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float M1 = (float)dz[k], P0 = (float)dz[3 + k], P1 = (float)dz[6 + k];
            ddx[k] = fx3[k] * (M1 + P0 + P1);
            ddy[k] = 1.0f * (M1 + P1) + P0 * -2.0f;
        }
        float depthDX = ddx[0] + ddx[1] + ddx[2];
        float depthDY = ddy[0] + ddy[1] + ddy[2];
        depthDX *= (1 / 256.0f);
        depthDY *= (1 / 256.0f);
        variance = depthDX * depthDX + depthDY * depthDY;
    }

    F4 accum = F4{0, 0, 0, 0};

    // ---- PassAmbientOcclusion (legacypbr.d:400-444)
    if (P.passMask & (1u << 1)) {
        float px = (float)i + 0.5f, py = (float)j + 0.5f;
        float avg = (linearSampleL16(P.depth, P.nDepth, 1, px, py) + linearSampleL16(P.depth, P.nDepth, 2, px, py)
                     + linearSampleL16(P.depth, P.nDepth, 3, px, py) + linearSampleL16(P.depth, P.nDepth, 4, px, py)) * 0.25f;
        float diff = (float)dz[4] - avg;
        float cavity = (diff + 23040.0f) * (1.0f / 23040);
        if (cavity >= 1) cavity = 1;
        else if (cavity < 0) cavity = 0;
        accum = accum + base * (cavity * P.ambient);
    }

    // ---- PassObliqueShadowLight (legacypbr.d:460-616)
    if (P.passMask & (1u << 2)) {
        float lightPassed = 0.0f;
        const int depthCenter = dz[4];
#pragma unroll
        for (int s = 1; s <= 8; ++s) {  // the two SSE groups (legacypbr.d:519-564)
            int x1 = min(i + s, P.W - 1), x2 = max(i - s, 0), y = max(j - s, 0);
            const uint16_t* scan = rowPtr<uint16_t>(D0, y);
            int z = depthCenter + s;
            float diff1 = (float)(z - (int)__ldg(scan + x1));
            float diff2 = (float)(z - (int)__ldg(scan + x2));
            float c1 = fmaxf(0.0f, fminf(1.0f, 1.0f + diff1 * 0.00006510416f));
            float c2 = fmaxf(0.0f, fminf(1.0f, 1.0f + diff2 * 0.00006510416f));
            lightPassed += (c1 + c2 * 0.7f) * P.shadowW[s];
        }
#pragma unroll
        for (int s = 9; s <= 10; ++s) {  // scalar tail (legacypbr.d:566-609)
            int x1 = min(i + s, P.W - 1), x2 = max(i - s, 0), y = max(j - s, 0);
            const uint16_t* scan = rowPtr<uint16_t>(D0, y);
            int z = depthCenter + s;
            int diff1 = z - (int)__ldg(scan + x1), diff2 = z - (int)__ldg(scan + x2);
            float c1 = diff1 >= 0 ? 1.f : (diff1 < -15360 ? 0.f : (float)(diff1 + 15360) * (1.0f / 15360));
            float c2 = diff2 >= 0 ? 1.f : (diff2 < -15360 ? 0.f : (float)(diff2 + 15360) * (1.0f / 15360));
            lightPassed += (c1 + c2 * 0.7f) * P.shadowW[s];
        }
        float k = lightPassed * P.shadowInvTotal;
        accum.x = accum.x + base.x * P.light1[0] * k;
        accum.y = accum.y + base.y * P.light1[1] * k;
        accum.z = accum.z + base.z * P.light1[2] * k;
    }

    // ---- PassDirectionalLight (legacypbr.d:636-666)
    if (P.passMask & (1u << 3)) {
        float dot = 0;
        dot += normal.x * P.light2Dir[0];
        dot += normal.y * P.light2Dir[1];
        dot += normal.z * P.light2Dir[2];
        float f = 0.5f + 0.5f * dot;
        float a = 0.24f - mat.x * 0.5f;
        f = 0.0f + (1.0f - 0.0f) * (f - a) / (1.0f - a);  // linmap, core/dplug/core/math.d:25-28
        accum.x += base.x * P.light2Col[0] * f;
        accum.y += base.y * P.light2Col[1] * f;
        accum.z += base.z * P.light2Col[2] * f;
    }

    // toEye is shared by the specular and skybox passes (legacypbr.d:756-757, 914-915)
    const F4 toEye = fastNormalize(P.rsqrtTab, F4{0.5f - (float)i * P.invW, (float)j * P.invH - 0.5f, 1.0f, 0.0f});

    // ---- PassSpecularLight (legacypbr.d:720-813)
    if (P.passMask & (1u << 4)) {
        F4 half = fastNormalize(P.rsqrtTab, F4{toEye.x - (-P.light3Dir[0]), toEye.y - (-P.light3Dir[1]),
                                               toEye.z - (-P.light3Dir[2]), toEye.w - 0.0f});
        float sf = dot4(half, normal);
        if (sf < 1e-3f) sf = 1e-3f;
        float exponent = __ldg(P.expTab + roughByte);
        float Ft = 1.0f / (1.0f + exponent * variance * 4e-5f);
        float tok = (1.0f + exponent * Ft) / (1.0f + exponent);
        sf = pow_ps(sf, exponent * Ft);
        float roughFactor = 10 * (1.0f - mat.x) * (1 - mat.y * 0.5f);
        sf = sf * roughFactor * tok;
        float k = sf * mat.z;
        accum.x = accum.x + base.x * P.light3Col[0] * k;
        accum.y = accum.y + base.y * P.light3Col[1] * k;
        accum.z = accum.z + base.z * P.light3Col[2] * k;
        accum.w = accum.w + base.w * 0.0f * k;
    }

    // ---- PassSkyboxReflections (legacypbr.d:872-930)
    if ((P.passMask & (1u << 5)) && P.nSky > 0) {
        float lod = variance * P.skyPixels;
        lod = 0.5f * fastlog2(1.0f + lod * 0.00001f) + (6.0f / 255.0f) * (float)roughByte;
        float sum = 2 * dot4(toEye, normal);  // _mm_reflectnormal_ps, ransac.d:31-36
        float rx = toEye.x - sum * normal.x, ry = toEye.y - sum * normal.y;
        float skyx = 0.5f + ((0.5f - rx * 0.5f) * P.skyFX);
        float skyy = 0.5f + ((0.5f + ry * 0.5f) * P.skyFY);
        F4 sky = linearMipmapSampleRGBA(P.sky, P.nSky, lod, skyx, skyy);
        float k = mat.y * (P.skyAmount * div255);
        accum.x = accum.x + base.x * sky.x * k;
        accum.y = accum.y + base.y * sky.y * k;
        accum.z = accum.z + base.z * sky.z * k;
        accum.w = accum.w + base.w * sky.w * k;
    }

    // ---- PassEmissiveContribution (legacypbr.d:948-1007, non-legacy branch)
    if (P.passMask & (1u << 6)) {
        float ic = (float)i + 0.5f, jc = (float)j + 0.5f;
        F4 c1 = linearSampleRGBA(P.diffuse, P.nDiffuse, 1, ic, jc);
        F4 c2 = linearSampleRGBA(P.diffuse, P.nDiffuse, 2, ic, jc);
        F4 c3 = linearSampleRGBA(P.diffuse, P.nDiffuse, 3, ic, jc);
        F4 c4 = cubicSampleRGBA(P.diffuse, P.nDiffuse, 4, ic, jc);
        F4 c5 = cubicSampleRGBA(P.diffuse, P.nDiffuse, 5, ic, jc);
        float noise = ((float)kBlueNoise16x16[(i & 15) * 16 + (j & 15)] - 127.5f) * 0.003f;
        const float AMT = 0.002f * 0.67f;
        F4 em = c1 * AMT;
        em = em + c2 * AMT;
        em = em + c3 * AMT;
        em = em + c4 * AMT;
        em = em + c5 * AMT * (1 + noise);
        accum = accum + em;
    }

    // ---- PassClampAndConvertTo8bit (legacypbr.d:1057-1107, non-legacy branch)
    float toneRatio = P.toneRatio / 3;
    float ex0 = fmaxf(0.0f, accum.x - P.toneThr), ex1 = fmaxf(0.0f, accum.y - P.toneThr), ex2 = fmaxf(0.0f, accum.z - P.toneThr);
    float luma = 0.212655f * ex0 + 0.715158f * ex1 + 0.072187f * ex2;
    float add = luma * toneRatio;
    return packOutput(accum.x + add, accum.y + add, accum.z + add, 1.0f);
}

// One CTA per work box (a GUI tile, at most 64x64 — gui/dplug/gui/graphics.d:59-60); 32x8 threads
// sweep it in 32x8 steps so each warp reads and writes whole 128-byte row segments.
__global__ void __launch_bounds__(256) k_pbr_exact(const __grid_constant__ PbrArgs P) {
    for (int b = blockIdx.x; b < P.nBoxes; b += gridDim.x) {
        const Box box = P.boxes[b];
        for (int j = box.miny + threadIdx.y; j < box.maxy; j += blockDim.y)
            for (int i = box.minx + threadIdx.x; i < box.maxx; i += blockDim.x) {
                if (!(P.passMask & (1u << 7))) continue;  // no clamp pass: the frame buffer is left untouched
                rowPtrW<uint32_t>(P.out, j)[i] = shadePixelExact(P, i, j);
            }
    }
}

// Test hook behind b2mip_sample: evaluates Mipmap.linearSample / cubicSample / linearMipmapSample
// (graphics/dplug/graphics/mipmap.d:149-496) for a list of query points.
__global__ void k_sample(const DevLevel* levels, int nLevels, int color, int kind, int n, const float* lvl,
                         const float* x, const float* y, float* out4) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    F4 r{0, 0, 0, 0};
    if (color == 0) {
        if (kind == 0) r = linearSampleRGBA(levels, nLevels, (int)lvl[t], x[t], y[t]);
        else if (kind == 1) r = cubicSampleRGBA(levels, nLevels, (int)lvl[t], x[t], y[t]);
        else r = linearMipmapSampleRGBA(levels, nLevels, lvl[t], x[t], y[t]);
    } else {
        if (kind == 0) r.x = linearSampleL16(levels, nLevels, (int)lvl[t], x[t], y[t]);
        else if (kind == 1) r.x = cubicSampleL16(levels, nLevels, (int)lvl[t], x[t], y[t]);
        else {
            int il = __float2int_rz(lvl[t]);
            float fl = lvl[t] - (float)il;
            float a = linearSampleL16(levels, nLevels, il, x[t], y[t]);
            r.x = (fl == 0) ? a : a * (1 - fl) + linearSampleL16(levels, nLevels, il + 1, x[t], y[t]) * fl;
        }
    }
    out4[4 * t + 0] = r.x; out4[4 * t + 1] = r.y; out4[4 * t + 2] = r.z; out4[4 * t + 3] = r.w;
}

void launch_sample(const DevLevel* levels, int nLevels, int color, int kind, int n, const float* lvl, const float* x,
                   const float* y, float* out4, cudaStream_t stream) {
    k_sample<<<(n + 127) / 128, 128, 0, stream>>>(levels, nLevels, color, kind, n, lvl, x, y, out4);
}

void launch_pbr_exact(const PbrArgs& args, int gridBlocks, cudaStream_t stream) {
    k_pbr_exact<<<gridBlocks, dim3(32, 8), 0, stream>>>(args);
}

}  // namespace b2
