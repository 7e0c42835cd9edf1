// Per-group gradient clipping + Adam + step counter in ONE launch
// (reference learner.py:176-183: two clip_grad_norm_ calls, Adam.step, LambdaLR.step).
//
// The parameter vector is tiny (14 144 floats at H=256, 69 312 at H=512) but the two clip
// norms need every gradient entry before any parameter can move.  One thread-block
// cluster of 8 CTAs (8 x 1024 threads, co-scheduled by hardware) does both phases in a
// single launch: each CTA reduces the squares of its slice in float64, the 8 partial pairs
// are exchanged through distributed shared memory, one cluster barrier later every CTA
// holds the same two norms and applies Adam to its slice.  The gradient arrives as float64
// (sum of per-CTA float32 partials, possibly all-reduced over ranks); optimizer state stays
// float32 in HBM and the arithmetic of one step is carried out in float64.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kAdamThreads = 1024;
constexpr int kAdamCluster = 8;

__global__ void __cluster_dims__(kAdamCluster, 1, 1) __launch_bounds__(kAdamThreads)
clip_adam_kernel(float* __restrict__ params, const double* __restrict__ grad, float* __restrict__ m,
                 float* __restrict__ v, int64_t* __restrict__ step, int64_t n_policy,
                 int64_t n_total, float max_norm, float lr, float beta1, float beta2, float eps,
                 double* __restrict__ norms_out) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double s_warp[2][kAdamThreads / 32];
    __shared__ double s_cta[2];   // this CTA's partial sums of squares (read by the peers)
    __shared__ double s_coef[2];
    __shared__ double s_bias[2];  // step_size, sqrt(bias_correction2)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t first = (int64_t)cluster.block_rank() * kAdamThreads + tid;
    const int64_t stride = (int64_t)kAdamCluster * kAdamThreads;
    const int64_t t = *step + 1;  // nobody writes *step before the final cluster barrier

    double ss0 = 0.0, ss1 = 0.0;
    for (int64_t i = first; i < n_total; i += stride) {
        const double g = grad[i];
        if (i < n_policy) ss0 += g * g;
        else ss1 += g * g;
    }
    ss0 = warp_sum_f64(ss0);
    ss1 = warp_sum_f64(ss1);
    if (lane == 0) s_warp[0][warp] = ss0, s_warp[1][warp] = ss1;
    __syncthreads();
    if (tid < 2) {
        double s = 0.0;
        for (int i = 0; i < kAdamThreads / 32; ++i) s += s_warp[tid][i];
        s_cta[tid] = s;
    }
    if (tid == 2) {
        const double bc1 = 1.0 - pow((double)beta1, (double)t);
        s_bias[0] = (double)lr / bc1;
        s_bias[1] = sqrt(1.0 - pow((double)beta2, (double)t));
    }
    cluster.sync();  // all 8 partial pairs are in place
    if (tid < 2) {
        double s = 0.0;
        for (int r = 0; r < kAdamCluster; ++r) s += *cluster.map_shared_rank(&s_cta[tid], r);
        const double norm = sqrt(s);
        // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
        s_coef[tid] = fmin(1.0, (double)max_norm / (norm + 1e-6));
        if (norms_out && cluster.block_rank() == 0) norms_out[tid] = norm;
    }
    __syncthreads();
    const double b1 = beta1, b2 = beta2;
    const double step_size = s_bias[0], bc2_sqrt = s_bias[1];
    const double c0 = s_coef[0], c1 = s_coef[1];
    for (int64_t i = first; i < n_total; i += stride) {
        const double g = grad[i] * (i < n_policy ? c0 : c1);
        const double mi = b1 * (double)m[i] + (1.0 - b1) * g;
        const double vi = b2 * (double)v[i] + (1.0 - b2) * g * g;
        const double denom = sqrt(vi) / bc2_sqrt + (double)eps;
        params[i] = (float)((double)params[i] - step_size * mi / denom);
        m[i] = (float)mi;
        v[i] = (float)vi;
    }
    cluster.sync();  // peers finished reading this CTA's shared memory; every CTA has read *step
    if (cluster.block_rank() == 0 && tid == 0) *step = t;
}

}  // namespace

extern "C" int impala_clip_adam(float* params, const double* grad, float* m, float* v,
                                int64_t* step, int64_t n_policy, int64_t n_total, float max_norm,
                                float lr, float beta1, float beta2, float eps, double* norms_out,
                                void* stream) {
    if (!params || !grad || !m || !v || !step) return IMPALA_ERR_BAD_ARG;
    if (n_total < 1 || n_policy < 0 || n_policy > n_total) return IMPALA_ERR_BAD_ARG;
    clip_adam_kernel<<<kAdamCluster, kAdamThreads, 0, (cudaStream_t)stream>>>(
        params, grad, m, v, step, n_policy, n_total, max_norm, lr, beta1, beta2, eps, norms_out);
    return impala_launch_status();
}
