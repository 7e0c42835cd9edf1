// Per-group gradient clipping + Adam + step counter in ONE launch
// (reference learner.py:176-183: two clip_grad_norm_ calls, Adam.step, LambdaLR.step).
//
// The parameter vector is tiny (14 144 floats at H=256, 69 312 at H=512) but the two clip
// norms need every gradient entry before any parameter can move.  One thread-block
// cluster of 8 CTAs (8 x 1024 threads, co-scheduled by hardware) does both phases in a
// single launch: each CTA reduces the squares of its slice in float64, the 8 partial pairs
// are exchanged through distributed shared memory, one cluster barrier later every CTA
// holds the same two norms and applies Adam to its slice.  The gradient arrives as float64
// (sum of per-CTA float32 partials, possibly all-reduced over ranks) and the norms are reduced
// in float64; optimizer state stays float32 in HBM and the per-element update runs in float32.
#include <cooperative_groups.h>
#include <math.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

constexpr int kAdamThreads = 1024;
constexpr int kAdamCluster = 8;

// state[0] = step count (int64), state[1] / state[2] = beta1^t / beta2^t as float64 bit patterns
// (all-zero state = fresh optimizer): running powers replace two float64 pow() calls per step.
__global__ void __cluster_dims__(kAdamCluster, 1, 1) __launch_bounds__(kAdamThreads)
clip_adam_kernel(float* __restrict__ params, const double* __restrict__ grad, float* __restrict__ m,
                 float* __restrict__ v, int64_t* __restrict__ state, int64_t n_policy,
                 int64_t n_total, float max_norm, float lr, float beta1, float beta2, float eps,
                 double* __restrict__ norms_out) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double s_warp[2][kAdamThreads / 32];
    __shared__ double s_cta[2];   // this CTA's partial sums of squares (read by the peers)
    __shared__ float s_coef[2];
    __shared__ float s_bias[2];   // step_size = lr / (1 - beta1^t), 1 / sqrt(1 - beta2^t)
    __shared__ double s_pow[2];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t first = (int64_t)cluster.block_rank() * kAdamThreads + tid;
    const int64_t stride = (int64_t)kAdamCluster * kAdamThreads;

    // phase 1: this thread's gradient entries (kept in registers when there are at most two, the
    // benchmark sizes) and the float32 state they will update - all loads issued up front
    constexpr int kKeep = 2;
    double gk[kKeep];
    float pk[kKeep], mk[kKeep], vk[kKeep];
    double ss0 = 0.0, ss1 = 0.0;
#pragma unroll
    for (int k = 0; k < kKeep; ++k) {
        const int64_t i = first + k * stride;
        gk[k] = i < n_total ? grad[i] : 0.0;
        pk[k] = i < n_total ? params[i] : 0.f;
        mk[k] = i < n_total ? m[i] : 0.f;
        vk[k] = i < n_total ? v[i] : 0.f;
    }
#pragma unroll
    for (int k = 0; k < kKeep; ++k) {
        const int64_t i = first + k * stride;
        if (i < n_policy) ss0 += gk[k] * gk[k];
        else ss1 += gk[k] * gk[k];
    }
    for (int64_t i = first + kKeep * stride; i < n_total; i += stride) {
        const double g = grad[i];
        if (i < n_policy) ss0 += g * g;
        else ss1 += g * g;
    }
    ss0 = warp_sum_f64(ss0);
    ss1 = warp_sum_f64(ss1);
    if (lane == 0) s_warp[0][warp] = ss0, s_warp[1][warp] = ss1;
    if (tid == 64) {  // bias corrections from the running powers (nobody writes state before the end)
        const double p1 = state[0] == 0 ? 1.0 : __longlong_as_double(state[1]);
        const double p2 = state[0] == 0 ? 1.0 : __longlong_as_double(state[2]);
        s_pow[0] = p1 * (double)beta1, s_pow[1] = p2 * (double)beta2;
        s_bias[0] = (float)((double)lr / (1.0 - s_pow[0]));
        s_bias[1] = (float)(1.0 / sqrt(1.0 - s_pow[1]));
    }
    __syncthreads();
    if (tid < 2) {
        double s = 0.0;
        for (int i = 0; i < kAdamThreads / 32; ++i) s += s_warp[tid][i];
        s_cta[tid] = s;
    }
    cluster.sync();  // all 8 partial pairs are in place
    if (tid < 2) {
        double s = 0.0;
        for (int r = 0; r < kAdamCluster; ++r) s += *cluster.map_shared_rank(&s_cta[tid], r);
        const double norm = sqrt(s);
        // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
        s_coef[tid] = (float)fmin(1.0, (double)max_norm / (norm + 1e-6));
        if (norms_out && cluster.block_rank() == 0) norms_out[tid] = norm;
    }
    __syncthreads();
    // phase 2: Adam in float32 arithmetic (the state is float32; one step's rounding is ~1e-7)
    const float b1 = beta1, b2 = beta2, step_size = s_bias[0], inv_bc2_sqrt = s_bias[1];
    const float c0 = s_coef[0], c1 = s_coef[1];
    auto update = [&](int64_t i, float g, float p, float mi, float vi) {
        g *= (i < n_policy ? c0 : c1);
        mi = fmaf(b1, mi, (1.f - b1) * g);
        vi = fmaf(b2, vi, (1.f - b2) * g * g);
        const float denom = fmaf(sqrtf(vi), inv_bc2_sqrt, eps);
        params[i] = p - step_size * mi / denom;
        m[i] = mi;
        v[i] = vi;
    };
#pragma unroll
    for (int k = 0; k < kKeep; ++k) {
        const int64_t i = first + k * stride;
        if (i < n_total) update(i, (float)gk[k], pk[k], mk[k], vk[k]);
    }
    for (int64_t i = first + kKeep * stride; i < n_total; i += stride)
        update(i, (float)grad[i], params[i], m[i], v[i]);
    cluster.sync();  // peers finished reading this CTA's shared memory; every CTA has read state
    if (cluster.block_rank() == 0 && tid == 64) {
        state[0] += 1;
        state[1] = __double_as_longlong(s_pow[0]);
        state[2] = __double_as_longlong(s_pow[1]);
    }
}

}  // namespace

extern "C" int impala_clip_adam(float* params, const double* grad, float* m, float* v,
                                int64_t* state, int64_t n_policy, int64_t n_total, float max_norm,
                                float lr, float beta1, float beta2, float eps, double* norms_out,
                                void* stream) {
    if (!params || !grad || !m || !v || !state) return IMPALA_ERR_BAD_ARG;
    if (n_total < 1 || n_policy < 0 || n_policy > n_total) return IMPALA_ERR_BAD_ARG;
    clip_adam_kernel<<<kAdamCluster, kAdamThreads, 0, (cudaStream_t)stream>>>(
        params, grad, m, v, state, n_policy, n_total, max_norm, lr, beta1, beta2, eps, norms_out);
    return impala_launch_status();
}
