// Probe for the ptxas FMUL2+FADD2 -> FFMA2 contraction (dplug_b200/csrc/mip_math.cuh): evaluates one 2x2 quad with the
// packed premultiplying box, the scalar one and step-by-step intrinsics, and prints the intermediate values.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o premul_rounding_probe premul_rounding_probe.cu
#include <cstdio>
#include "../../dplug_b200/csrc/mip_math.cuh"
using namespace b2;
__global__ void k(const uint32_t* q, uint32_t* out, float* f, f32x2 nz) {
    // quad 0 = q[0..3], quad 1 = q[4..7]
    uint2 o = premul4x2(q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], nz);
    out[0] = o.x; out[1] = o.y;
    out[2] = premul4(q[0], q[1], q[2], q[3]);
    out[3] = premul4(q[4], q[5], q[6], q[7]);
    // scalar reference with intrinsics
    const float inv = 1.0f / 255;
    for (int quad = 0; quad < 2; ++quad) {
        float s[4] = {0, 0, 0, 0};
        for (int t = 0; t < 4; ++t) {
            uint32_t p = q[quad * 4 + t];
            float a = __fmul_rn((float)(p >> 24), inv);
            for (int ch = 0; ch < 3; ++ch) {
                float c = (float)((p >> (8 * ch)) & 0xff);
                float v = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(c, inv), c), inv), a);
                s[ch] = t == 0 ? v : __fadd_rn(s[ch], v);
                f[quad * 32 + t * 4 + ch] = v;
            }
            s[3] = t == 0 ? a : __fadd_rn(s[3], a);
        }
        uint32_t r = 0;
        for (int ch = 0; ch < 4; ++ch) {
            float v = __fadd_rn(__fmul_rn(__fmul_rn(s[ch], 0.25f), 255.0f), 0.5f);
            f[quad * 32 + 16 + ch] = v;
            r |= ((uint32_t)__float2int_rz(v) & 0xff) << (8 * ch);
        }
        out[4 + quad] = r;
    }
    // packed intermediates for quad pair
    Lin4x2 A = premulLinear2(q[0], q[4], nz), B = premulLinear2(q[1], q[5], nz), C = premulLinear2(q[2], q[6], nz), D = premulLinear2(q[3], q[7], nz);
    float lo, hi;
    unpack2(A.g, lo, hi); f[64] = lo; unpack2(B.g, lo, hi); f[65] = lo; unpack2(C.g, lo, hi); f[66] = lo; unpack2(D.g, lo, hi); f[67] = lo;
    f32x2 g = add2(add2(add2(A.g, B.g), C.g), D.g);
    unpack2(g, lo, hi); f[68] = lo;
    const f32x2 qq = pack2(0.25f, 0.25f), s255 = pack2(255.0f, 255.0f), h = pack2(0.5f, 0.5f);
    f32x2 g1 = mul2(g, qq); unpack2(g1, lo, hi); f[69] = lo;
    f32x2 g2 = mul2(g1, s255); unpack2(g2, lo, hi); f[70] = lo;
    f32x2 g3 = add2(g2, h); unpack2(g3, lo, hi); f[71] = lo;
}
int main() {
    uint32_t hq[8];
    int quad[4][4] = {{117,203,90,151},{252,120,196,176},{233,153,28,71},{115,55,0,210}};
    for (int t = 0; t < 4; ++t) { hq[t] = quad[t][0] | (quad[t][1] << 8) | (quad[t][2] << 16) | ((uint32_t)quad[t][3] << 24); hq[4 + t] = hq[t] ^ 0x01010101u; }
    uint32_t *dq, *dout; float* df;
    cudaMalloc(&dq, 32); cudaMalloc(&dout, 32); cudaMalloc(&df, 4 * 128);
    cudaMemcpy(dq, hq, 32, cudaMemcpyHostToDevice);
    k<<<1, 1>>>(dq, dout, df, kNegZero2);
    uint32_t ho[8]; float hf[128];
    cudaMemcpy(ho, dout, 32, cudaMemcpyDeviceToHost); cudaMemcpy(hf, df, 512, cudaMemcpyDeviceToHost);
    printf("x2: %08x %08x  old: %08x %08x  scalar: %08x %08x  (%s)\n", ho[0], ho[1], ho[2], ho[3], ho[4], ho[5], cudaGetErrorString(cudaGetLastError()));
    for (int t = 0; t < 4; ++t) printf("scalar g lin[%d] = %a   packed = %a\n", t, hf[t * 4 + 1], hf[64 + t]);
    printf("scalar final g = %a\n", hf[16 + 1]);
    printf("packed sum %a  *0.25 %a  *255 %a  +0.5 %a\n", hf[68], hf[69], hf[70], hf[71]);
    return 0;
}
