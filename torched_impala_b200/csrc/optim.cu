// Per-group gradient clipping + Adam + step counter in ONE launch
// (reference learner.py:176-183: two clip_grad_norm_ calls, Adam.step, LambdaLR.step).
//
// The parameter vector is tiny (14 085 floats at H=256, 69 125 at H=512), so a single
// 1024-thread CTA does both phases: float64 sum of squares per group -> clip
// coefficients -> Adam update.  The gradient arrives as float64 (sum of per-CTA float32
// partials, possibly all-reduced over ranks); optimizer state stays float32 in HBM and
// the arithmetic of one step is carried out in float64.
#include <math.h>

#include "common.cuh"

namespace {

constexpr int kAdamThreads = 1024;

__global__ void __launch_bounds__(kAdamThreads)
clip_adam_kernel(float* __restrict__ params, const double* __restrict__ grad, float* __restrict__ m,
                 float* __restrict__ v, int64_t* __restrict__ step, int64_t n_policy,
                 int64_t n_total, float max_norm, float lr, float beta1, float beta2, float eps,
                 double* __restrict__ norms_out) {
    __shared__ double s_part[2][kAdamThreads / 32];
    __shared__ double s_coef[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double ss0 = 0.0, ss1 = 0.0;
    for (int64_t i = tid; i < n_total; i += kAdamThreads) {
        const double g = grad[i];
        if (i < n_policy) ss0 += g * g;
        else ss1 += g * g;
    }
    ss0 = warp_sum_f64(ss0);
    ss1 = warp_sum_f64(ss1);
    if (lane == 0) s_part[0][warp] = ss0, s_part[1][warp] = ss1;
    __syncthreads();
    if (tid < 2) {
        double s = 0.0;
        for (int i = 0; i < kAdamThreads / 32; ++i) s += s_part[tid][i];
        const double norm = sqrt(s);
        // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
        s_coef[tid] = fmin(1.0, (double)max_norm / (norm + 1e-6));
        if (norms_out) norms_out[tid] = norm;
    }
    __syncthreads();
    const int64_t t = *step + 1;
    const double b1 = beta1, b2 = beta2;
    const double bc1 = 1.0 - pow(b1, (double)t);
    const double bc2_sqrt = sqrt(1.0 - pow(b2, (double)t));
    const double step_size = (double)lr / bc1;
    const double c0 = s_coef[0], c1 = s_coef[1];
    for (int64_t i = tid; i < n_total; i += kAdamThreads) {
        const double g = grad[i] * (i < n_policy ? c0 : c1);
        const double mi = b1 * (double)m[i] + (1.0 - b1) * g;
        const double vi = b2 * (double)v[i] + (1.0 - b2) * g * g;
        const double denom = sqrt(vi) / bc2_sqrt + (double)eps;
        params[i] = (float)((double)params[i] - step_size * mi / denom);
        m[i] = (float)mi;
        v[i] = (float)vi;
    }
    __syncthreads();
    if (tid == 0) *step = t;
}

}  // namespace

extern "C" int impala_clip_adam(float* params, const double* grad, float* m, float* v,
                                int64_t* step, int64_t n_policy, int64_t n_total, float max_norm,
                                float lr, float beta1, float beta2, float eps, double* norms_out,
                                void* stream) {
    if (!params || !grad || !m || !v || !step) return IMPALA_ERR_BAD_ARG;
    if (n_total < 1 || n_policy < 0 || n_policy > n_total) return IMPALA_ERR_BAD_ARG;
    clip_adam_kernel<<<1, kAdamThreads, 0, (cudaStream_t)stream>>>(
        params, grad, m, v, step, n_policy, n_total, max_norm, lr, beta1, beta2, eps, norms_out);
    return impala_launch_status();
}
