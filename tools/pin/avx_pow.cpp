// Oracle pinning aid (test infrastructure): _mm_pow_ps(x, y) = exp_ps(y * log_ps(x)) evaluated by a THIRD-PARTY port of
// Julien Pommier's sse_mathfun that ships with PyTorch (ATen/native/cpu/avx_mathfun.h, the AVX2 twin of the SSE code
// intel-intrinsics' _mm_log_ps / _mm_exp_ps are ported from).  tests/test_np_reference.py compares it with the
// restatements over 10^6 operands.  Built by the test with: g++ -O2 -mavx2 -ffp-contract=off -DCPU_CAPABILITY_AVX2
// -I<torch include> -shared -fPIC.
#include <ATen/native/cpu/avx_mathfun.h>

extern "C" void avx_pow(const float* x, const float* y, float* out, int n) {
    int i = 0;
    for (; i + 8 <= n; i += 8) {
        v8sf vx = _mm256_loadu_ps(x + i), vy = _mm256_loadu_ps(y + i);
        _mm256_storeu_ps(out + i, exp256_ps(_mm256_mul_ps(vy, log256_ps(vx))));
    }
    if (i < n) {
        float bx[8] = {1, 1, 1, 1, 1, 1, 1, 1}, by[8] = {0}, bo[8];
        for (int k = 0; i + k < n; ++k) { bx[k] = x[i + k]; by[k] = y[i + k]; }
        _mm256_storeu_ps(bo, exp256_ps(_mm256_mul_ps(_mm256_loadu_ps(by), log256_ps(_mm256_loadu_ps(bx)))));
        for (int k = 0; i + k < n; ++k) out[i + k] = bo[k];
    }
}
