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


// ---------------------------------------------------------------------------------------------
// Data-parallel learners: one-shot all-reduce over NVLink peer memory fused with clip + Adam.
//
// Every rank leaves its float64 [gradient | extra scalars] contribution in a buffer that the other
// ranks of the node have mapped (CUDA IPC).  Instead of an NCCL all-reduce between the backward
// and the optimizer, the optimizer kernel of each rank
//   1. posts "my contribution for step s is complete" into every peer's flag block,
//   2. waits until all ranks' flags for step s have arrived in its own block,
//   3. reads element i from all ranks (peer loads over NVLink / NVSwitch) and adds them in rank
//      order - every rank forms bit-identical sums, so the replicas cannot drift,
//   4. runs the clip norms and Adam on the sum.
// The contribution buffers are double-buffered by step parity (buffer s & 1 for step s), which
// makes a second "I have read your buffer" round trip unnecessary: a rank overwrites buffer s & 1
// in the backward of step s + 2, i.e. after its optimizer kernel of step s + 1 saw every peer's
// ready flag for s + 1 - and a peer posts that flag only after its kernel of step s (the one that
// read the buffer) has finished.
// Flags are monotonically increasing step numbers (int64, never reset); spins are bounded by a
// clock timeout (~35 s) that traps (a lost rank becomes a launch failure on the others, not a hang).
// Flag block of a rank: int64[world], entry r written by rank r.
struct PeerArgs {
    const double* const* contrib;  // device array [world]: every rank's 2 x [n_total + n_pad] doubles
    int64_t buf_stride;            // doubles between the two parity buffers
    long long* const* flags;       // device array [world]: every rank's flag block
    long long* seq;                // this rank's step counter (device, 1 word)
    int rank, world, n_extra;
};

__device__ __forceinline__ void st_relaxed_sys(long long* p, long long v) {
    asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_acquire_sys(const long long* p) {
    long long v;
    asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_peer_f64(const double* p) {  // not served from a local cache
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// lane r of the calling warp polls entry r of the block (all ranks in parallel), then the warp meets
__device__ __forceinline__ void wait_flags(const long long* block, int n, long long seq, int lane) {
    if (lane < n) {
        const long long t0 = clock64();
        while (ld_acquire_sys(block + lane) < seq)
            if (clock64() - t0 > (1ll << 36)) __trap();  // ~35 s: a peer is gone
    }
    __syncwarp();
}

__global__ void __cluster_dims__(kAdamCluster, 1, 1) __launch_bounds__(kAdamThreads)
allreduce_clip_adam_kernel(float* __restrict__ params, double* __restrict__ reduced, PeerArgs peer,
                           float* __restrict__ m, float* __restrict__ v, int64_t* __restrict__ state,
                           int64_t n_policy, int64_t n_total, float max_norm, float lr, float beta1,
                           float beta2, float eps, double* __restrict__ norms_out) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double s_warp[2][kAdamThreads / 32];
    __shared__ double s_cta[2];
    __shared__ float s_coef[2];
    __shared__ float s_bias[2];
    __shared__ double s_pow[2];
    __shared__ long long s_seq;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int crank = (int)cluster.block_rank();
    const int64_t first = (int64_t)crank * kAdamThreads + tid;
    const int64_t stride = (int64_t)kAdamCluster * kAdamThreads;
    long long* my_flags = peer.flags[peer.rank];

    // 1. + 2.: post "ready" to everyone (one CTA, one lane per peer), then every CTA waits for all
    // ranks' flags.  This rank's contribution was written by earlier kernels of the stream, i.e. it
    // is already performed in its L2 (where peer reads are served); one system fence orders the
    // flag stores behind it.
    if (warp == 0) {
        const long long seq = *peer.seq + 1;
        if (lane == 0) s_seq = seq;
        if (crank == 0) {
            __threadfence_system();
            if (lane < peer.world) st_relaxed_sys(peer.flags[lane] + peer.rank, seq);
        }
        wait_flags(my_flags, peer.world, seq, lane);
    }
    if (tid == 64) {  // bias corrections from the running powers (nobody writes state before the end)
        const double p1 = state[0] == 0 ? 1.0 : __longlong_as_double(state[1]);
        const double p2 = state[0] == 0 ? 1.0 : __longlong_as_double(state[2]);
        s_pow[0] = p1 * (double)beta1, s_pow[1] = p2 * (double)beta2;
        s_bias[0] = (float)((double)lr / (1.0 - s_pow[0]));
        s_bias[1] = (float)(1.0 / sqrt(1.0 - s_pow[1]));
    }
    __syncthreads();

    // 3. rank-ordered sums of this thread's entries (all loads of an entry are independent)
    constexpr int kKeep = 2, kMaxWorld = 8;  // one NVLink node
    const int64_t boff = (s_seq & 1) * peer.buf_stride;  // this step's parity buffer
    auto gather = [&](int64_t i) {
        double c[kMaxWorld];
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if (r < peer.world) c[r] = ld_peer_f64(peer.contrib[r] + boff + i);
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r)
            if (r < peer.world) s += c[r];
        return s;
    };
    double gk[kKeep];
    float pk[kKeep], mk[kKeep], vk[kKeep];
    double ss0 = 0.0, ss1 = 0.0;
#pragma unroll
    for (int k = 0; k < kKeep; ++k) {
        const int64_t i = first + k * stride;
        gk[k] = i < n_total ? gather(i) : 0.0;
        pk[k] = i < n_total ? params[i] : 0.f;
        mk[k] = i < n_total ? m[i] : 0.f;
        vk[k] = i < n_total ? v[i] : 0.f;
        if (i < n_total) reduced[i] = gk[k];
        if (i < n_policy) ss0 += gk[k] * gk[k];
        else ss1 += gk[k] * gk[k];
    }
    for (int64_t i = first + kKeep * stride; i < n_total; i += stride) {
        const double g = gather(i);
        reduced[i] = g;
        if (i < n_policy) ss0 += g * g;
        else ss1 += g * g;
    }
    if (crank == 0 && tid < peer.n_extra) reduced[n_total + tid] = gather(n_total + tid);  // logged scalars
    ss0 = warp_sum_f64(ss0);
    ss1 = warp_sum_f64(ss1);
    if (lane == 0) s_warp[0][warp] = ss0, s_warp[1][warp] = ss1;
    __syncthreads();
    if (tid < 2) {
        double s = 0.0;
        for (int i = 0; i < kAdamThreads / 32; ++i) s += s_warp[tid][i];
        s_cta[tid] = s;
    }
    cluster.sync();  // all 8 partial pairs are in place; every CTA has finished its peer reads
    if (tid < 2) {
        double s = 0.0;
        for (int r = 0; r < kAdamCluster; ++r) s += *cluster.map_shared_rank(&s_cta[tid], r);
        const double norm = sqrt(s);
        s_coef[tid] = (float)fmin(1.0, (double)max_norm / (norm + 1e-6));
        if (norms_out && crank == 0) norms_out[tid] = norm;
    }
    __syncthreads();
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
        update(i, (float)reduced[i], params[i], m[i], v[i]);
    cluster.sync();  // peers finished reading this CTA's shared memory; every CTA has read state
    if (crank == 0 && tid == 64) {
        state[0] += 1;
        state[1] = __double_as_longlong(s_pow[0]);
        state[2] = __double_as_longlong(s_pow[1]);
    }
    if (crank == 0 && tid == 0) *peer.seq = s_seq;
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

extern "C" int impala_allreduce_clip_adam(float* params, double* reduced, const double* const* peer_contrib,
                                          int64_t buf_stride, long long* const* peer_flags, long long* seq,
                                          int rank, int world, int n_extra, float* m, float* v, int64_t* state, int64_t n_policy,
                                          int64_t n_total, float max_norm, float lr, float beta1, float beta2,
                                          float eps, double* norms_out, void* stream) {
    if (!params || !reduced || !peer_contrib || !peer_flags || !seq || !m || !v || !state)
        return IMPALA_ERR_BAD_ARG;
    if (n_total < 1 || n_policy < 0 || n_policy > n_total) return IMPALA_ERR_BAD_ARG;
    if (world < 1 || world > 8 || rank < 0 || rank >= world || n_extra < 0 || n_extra > kAdamThreads)
        return IMPALA_ERR_BAD_ARG;
    if (buf_stride < n_total + n_extra) return IMPALA_ERR_BAD_ARG;
    PeerArgs peer{peer_contrib, buf_stride, peer_flags, seq, rank, world, n_extra};
    allreduce_clip_adam_kernel<<<kAdamCluster, kAdamThreads, 0, (cudaStream_t)stream>>>(
        params, reduced, peer, m, v, state, n_policy, n_total, max_norm, lr, beta1, beta2, eps, norms_out);
    return impala_launch_status();
}
