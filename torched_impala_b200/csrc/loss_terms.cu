// Stand-alone pieces of the loss API (reference learner.py:298-321) for callers that want the
// helper functions rather than the fused kernel: per-row log-prob of the taken action and
// negative entropy (+ their backward), and float64 scalar reductions.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kMaxA = 64;

// one thread per row; A is small (<= 64), logits row-contiguous
__global__ void policy_terms_kernel(const float* __restrict__ logits, const int32_t* __restrict__ actions,
                                    float* __restrict__ lp_out, float* __restrict__ negent_out, int M, int A) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float* z = logits + (size_t)m * A;
    float mx = z[0];
    for (int k = 1; k < A; ++k) mx = fmaxf(mx, z[k]);
    float se = 0.f;
    for (int k = 0; k < A; ++k) se += expf(z[k] - mx);
    const float lse = mx + logf(se);
    float ne = 0.f;
    for (int k = 0; k < A; ++k) {
        const float l = z[k] - lse;
        ne += expf(l) * l;  // sum p log p  (learner.py:310-314)
    }
    const int a = actions[m];
    lp_out[m] = z[(a >= 0 && a < A) ? a : 0] - lse;  // learner.py:298-303
    negent_out[m] = ne;
}

// dlogits[m,k] = g_lp[m] (1[k==a] - p_k) + g_ne[m] p_k (log p_k - sum_j p_j log p_j)
__global__ void policy_terms_bwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ actions,
                                        const float* __restrict__ g_lp, const float* __restrict__ g_ne,
                                        float* __restrict__ dlogits, int M, int A) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float* z = logits + (size_t)m * A;
    float mx = z[0];
    for (int k = 1; k < A; ++k) mx = fmaxf(mx, z[k]);
    float se = 0.f;
    for (int k = 0; k < A; ++k) se += expf(z[k] - mx);
    const float lse = mx + logf(se);
    float ne = 0.f;
    for (int k = 0; k < A; ++k) {
        const float l = z[k] - lse;
        ne += expf(l) * l;
    }
    const float gl = g_lp ? g_lp[m] : 0.f, gn = g_ne ? g_ne[m] : 0.f;
    const int a = actions[m];
    for (int k = 0; k < A; ++k) {
        const float l = z[k] - lse, p = expf(l);
        dlogits[(size_t)m * A + k] = gl * ((k == a ? 1.f : 0.f) - p) + gn * p * (l - ne);
    }
}

// mode 0: sum a ; 1: 0.5 sum a^2 ; 2: sum a*b   -> atomically added to *out (float64)
__global__ void reduce_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, int mode,
                              double* __restrict__ out) {
    __shared__ double s_w[32];
    double s = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = a[i];
        s += mode == 0 ? x : (mode == 1 ? 0.5 * x * x : x * (double)b[i]);
    }
    s = warp_sum_f64(s);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_w[w];
        atomicAdd(out, t);
    }
}

}  // namespace

extern "C" int impala_policy_terms(const float* logits, const int32_t* actions, float* log_prob,
                                   float* neg_entropy, int M, int A, void* stream) {
    if (!logits || !actions || !log_prob || !neg_entropy || M < 1) return IMPALA_ERR_BAD_ARG;
    if (A < 1 || A > kMaxA) return IMPALA_ERR_UNSUPPORTED_SHAPE;
    policy_terms_kernel<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(logits, actions, log_prob,
                                                                          neg_entropy, M, A);
    return impala_launch_status();
}

extern "C" int impala_policy_terms_backward(const float* logits, const int32_t* actions,
                                            const float* grad_log_prob, const float* grad_neg_entropy,
                                            float* dlogits, int M, int A, void* stream) {
    if (!logits || !actions || !dlogits || M < 1) return IMPALA_ERR_BAD_ARG;
    if (A < 1 || A > kMaxA) return IMPALA_ERR_UNSUPPORTED_SHAPE;
    policy_terms_bwd_kernel<<<(M + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        logits, actions, grad_log_prob, grad_neg_entropy, dlogits, M, A);
    return impala_launch_status();
}

extern "C" int impala_reduce(const float* a, const float* b, int64_t n, int mode, double* out,
                             void* stream) {
    if (!a || !out || n < 1 || mode < 0 || mode > 2 || (mode == 2 && !b)) return IMPALA_ERR_BAD_ARG;
    cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    int grid = (int)((n + 255) / 256);
    if (grid > 592) grid = 592;
    reduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, n, mode, out);
    return impala_launch_status();
}
