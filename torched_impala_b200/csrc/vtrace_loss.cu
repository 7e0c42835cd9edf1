// V-trace targets, the three losses and their closed-form backward
// (reference learner.py:116-162 + helpers :298-321 + the non-MLP part of :175).
//
// One warp per trajectory, lanes across time.  A CTA owns kTraj = 8 consecutive
// trajectories (one per warp) and walks the unroll backwards in chunks of TC (<=128)
// steps: all 256 threads stage the chunk's (TC, 8[, A]) slices of the time-major
// tensors into shared memory with row-contiguous global loads (8 floats = one 32-byte
// sector per row of the scalar tensors, 8*A floats per row of the logits), each warp
// then reads its own column (row stride 9 / 8*AP+1 floats -> bank-conflict free), runs
// the backward recurrence as an affine-map suffix scan over the 32 lanes of each
// 32-step pass (carry between passes and chunks in a register), computes loss terms and
// gradients in registers, and the results go back through shared memory so the global
// stores are row-contiguous as well.
//
// Transcendentals use the hardware approximations (ex2/lg2.approx.ftz, relative error ~2^-22) in
// the base-2 domain: the arguments are differences from the row maximum (<= 0) and sums in
// [1, A], so the absolute error stays ~1e-7, far inside the 1e-5 parity budget, at a fraction of
// the instruction count of expf/logf.
//
// Reference quirks reproduced in IMPALA_MODE_REFERENCE:
//   delta_t = rho_t (r_t + gamma v_{t+1} - v_0)                  learner.py:126  (v[:1])
//   acc_i   = delta_i + disc_i c_i (acc_{i+1} - v_{i+1})          learner.py:130
//   vs = acc + v (:131);  pg_t = rho_t (r_t + disc_t vs_{t+1} - v_t)   (:135)
// i.e. the affine map F_i(x) = (delta_i - g_i v_{i+1}) + g_i x with g_i = disc_i c_i.
#include <math.h>

#include <cstdlib>

#include "common.cuh"

namespace {

// ex2 / lg2 hardware approximations with flush-to-zero (no denormal fix-up code): relative error
// ~2^-22.  Softmax is evaluated in the base-2 domain: zs = z * log2(e), p_k = 2^(zs_k - lse2).
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;

constexpr int kTraj = 8;             // trajectories (= warps) per CTA
constexpr int kThreads = kTraj * 32;
constexpr int kColStride = kTraj + 1;  // padded row stride of the scalar tiles

struct VtArgs {
    const float* cur_logits;
    const float* beh_logits;
    const int32_t* actions;
    const float* rewards;
    const uint8_t* done;
    const int32_t* lens;
    const float* v;
    float* vs;
    float* pg_adv;
    float* dlogits;
    float* dv;
    double* scalars;
    double* partials;        // [gridDim.x][4] per-CTA loss sums (workspace)
    unsigned int* counter;   // CTA arrival counter (workspace; zero on entry, zero on exit)
    int T, B, A, TC, mode;
    float gamma, rho_bar, c_bar, v_loss_c, policy_loss_c, entropy_c, inv_batch;
};

// AEXACT: the action count equals the padded count AP (compile-time divisions in the staging loops)
template <int AP, bool WITH_LOSS, bool VEC, bool AEXACT>
__global__ void __launch_bounds__(kThreads, 3) vtrace_kernel(VtArgs a) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int T = a.T, B = a.B, A = AEXACT ? AP : a.A, TC = a.TC;
    const int b0 = blockIdx.x * kTraj;
    const int LS = kTraj * AP + 1;  // logits tile row stride (odd -> conflict free)

    float* s_v = smem;                           // [TC+1][9]
    float* s_vs = s_v + (TC + 1) * kColStride;   // [TC+1][9]
    float* s_dv = s_vs + (TC + 1) * kColStride;  // [TC+1][9]
    float* s_r = s_dv + (TC + 1) * kColStride;   // [TC][9]   rewards in, pg_adv out
    int* s_act = reinterpret_cast<int*>(s_r + TC * kColStride);  // [TC][9] action | done<<30
    float* s_cur = reinterpret_cast<float*>(s_act + TC * kColStride);  // [TC][LS] logits in, dlogits out
    float* s_beh = s_cur + TC * LS;                                    // [TC][LS]
    __shared__ double s_red[kTraj][4];

    const int b = b0 + w;
    const bool live = b < B;
    int L = 0;
    float v0 = 0.f;
    if (live) {
        L = min(max(__ldg(a.lens + b), 0), T);
        v0 = __ldg(a.v + b);  // V(x_0): the reference's v[:1]
    }
    const int nb = min(kTraj, B - b0);  // live columns of this CTA

    double sum_vl = 0.0, sum_pl = 0.0, sum_ent = 0.0, sum_rw = 0.0;
    float carry = 0.f;  // acc at the first index after the current pass
    const int nchunks = (T + TC - 1) / TC;

    for (int c = nchunks - 1; c >= 0; --c) {
        const int t0 = c * TC;
        __syncthreads();  // previous chunk's stores have drained the tiles
        // ---- stage: row-contiguous global reads (128-bit when B % 8 == 0) ----
        if constexpr (VEC) {
            for (int idx = tid; idx < (TC + 1) * 2; idx += kThreads) {
                const int tt = idx >> 1, h = idx & 1, t = t0 + tt;
                float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
                if (t <= T) x = __ldg(reinterpret_cast<const float4*>(a.v + (size_t)t * B + b0) + h);
                float* d = s_v + tt * kColStride + 4 * h;
                d[0] = x.x, d[1] = x.y, d[2] = x.z, d[3] = x.w;
            }
            for (int idx = tid; idx < TC * 2; idx += kThreads) {
                const int tt = idx >> 1, h = idx & 1, t = t0 + tt;
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                int4 ac = make_int4(0, 0, 0, 0);
                uchar4 dn = make_uchar4(0, 0, 0, 0);
                if (t < T) {
                    const size_t g = (size_t)t * B + b0;
                    r = __ldg(reinterpret_cast<const float4*>(a.rewards + g) + h);
                    ac = __ldg(reinterpret_cast<const int4*>(a.actions + g) + h);
                    dn = __ldg(reinterpret_cast<const uchar4*>(a.done + g) + h);
                }
                float* dr = s_r + tt * kColStride + 4 * h;
                int* da = s_act + tt * kColStride + 4 * h;
                dr[0] = r.x, dr[1] = r.y, dr[2] = r.z, dr[3] = r.w;
                da[0] = (ac.x & 0x3fffffff) | (dn.x ? (1 << 30) : 0);
                da[1] = (ac.y & 0x3fffffff) | (dn.y ? (1 << 30) : 0);
                da[2] = (ac.z & 0x3fffffff) | (dn.z ? (1 << 30) : 0);
                da[3] = (ac.w & 0x3fffffff) | (dn.w ? (1 << 30) : 0);
            }
            const int q4 = 2 * A;  // float4 per time step: 8 trajectories x A logits
            for (int idx = tid; idx < TC * q4; idx += kThreads) {
                const int tt = idx / q4, q = idx - tt * q4, t = t0 + tt;
                float4 zc = make_float4(0.f, 0.f, 0.f, 0.f), zb = zc;
                if (t < T) {
                    const size_t g = ((size_t)t * B + b0) * A;
                    zc = __ldg(reinterpret_cast<const float4*>(a.cur_logits + g) + q);
                    zb = __ldg(reinterpret_cast<const float4*>(a.beh_logits + g) + q);
                }
                const float c4[4] = {zc.x, zc.y, zc.z, zc.w}, b4[4] = {zb.x, zb.y, zb.z, zb.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int rem = 4 * q + e, col = rem / A, k = rem - col * A;
                    s_cur[tt * LS + col * AP + k] = c4[e];
                    s_beh[tt * LS + col * AP + k] = b4[e];
                }
            }
        } else {
            for (int idx = tid; idx < (TC + 1) * kTraj; idx += kThreads) {
                const int tt = idx / kTraj, col = idx - tt * kTraj;
                const int t = t0 + tt;
                s_v[tt * kColStride + col] =
                    (t <= T && col < nb) ? __ldg(a.v + (size_t)t * B + b0 + col) : 0.f;
            }
            for (int idx = tid; idx < TC * kTraj; idx += kThreads) {
                const int tt = idx / kTraj, col = idx - tt * kTraj;
                const int t = t0 + tt;
                const bool ok = t < T && col < nb;
                const size_t g = (size_t)t * B + b0 + col;
                s_r[tt * kColStride + col] = ok ? __ldg(a.rewards + g) : 0.f;
                int packed = 0;
                if (ok) packed = (__ldg(a.actions + g) & 0x3fffffff) | (__ldg(a.done + g) ? (1 << 30) : 0);
                s_act[tt * kColStride + col] = packed;
            }
            const int rowlen = nb * A;  // contiguous floats per time step for this CTA
            for (int idx = tid; idx < TC * rowlen; idx += kThreads) {
                const int tt = idx / rowlen, rem = idx - tt * rowlen;
                const int t = t0 + tt;
                const int col = rem / A, k = rem - col * A;
                float zc = 0.f, zb = 0.f;
                if (t < T) {
                    const size_t g = ((size_t)t * B + b0) * A + rem;
                    zc = __ldg(a.cur_logits + g);
                    zb = __ldg(a.beh_logits + g);
                }
                s_cur[tt * LS + col * AP + k] = zc;
                s_beh[tt * LS + col * AP + k] = zb;
            }
        }
        __syncthreads();

        // ---- compute: this warp's trajectory, 32 steps per pass, last pass first ----
        if (live) {
            for (int p = TC / 32 - 1; p >= 0; --p) {
                const int tt = p * 32 + lane;
                const int t = t0 + tt;
                const bool valid = t < L;
                const float r = s_r[tt * kColStride + w];
                const int packed = s_act[tt * kColStride + w];
                const int act = packed & 0x3fffffff;
                const bool dn = (packed >> 30) & 1;
                const float v_t = s_v[tt * kColStride + w];
                const float v_n = s_v[(tt + 1) * kColStride + w];
                float z[AP], zb[AP];
#pragma unroll
                for (int k = 0; k < AP; ++k) {
                    z[k] = s_cur[tt * LS + w * AP + k];
                    zb[k] = s_beh[tt * LS + w * AP + k];
                }
                // log-softmax of both logit vectors (learner.py:298-303)
#pragma unroll
                for (int k = 0; k < AP; ++k) z[k] *= kLog2e, zb[k] *= kLog2e;  // base-2 domain
                float mx = z[0], mxb = zb[0];
#pragma unroll
                for (int k = 1; k < AP; ++k)
                    if (k < A) mx = fmaxf(mx, z[k]), mxb = fmaxf(mxb, zb[k]);
                float se = 0.f, seb = 0.f;
#pragma unroll
                for (int k = 0; k < AP; ++k)
                    if (k < A) se += ex2f(z[k] - mx), seb += ex2f(zb[k] - mxb);
                const float lse = mx + lg2f(se), lseb = mxb + lg2f(seb);
                float z_a = z[0], zb_a = zb[0];
#pragma unroll
                for (int k = 1; k < AP; ++k)
                    if (k == act) z_a = z[k], zb_a = zb[k];
                const float lp2_cur = z_a - lse, lp2_beh = zb_a - lseb;  // log2 pi(a)
                const float lp_cur = lp2_cur * kLn2;
                const float ratio = ex2f(lp2_cur - lp2_beh);                   // :121-123
                const float rho = valid ? fminf(ratio, a.rho_bar) : 0.f;       // :124
                const float cc = valid ? fminf(ratio, a.c_bar) : 0.f;          // :125
                const float disc = (valid && !dn) ? a.gamma : 0.f;             // :109
                const float g = disc * cc;
                float fa;  // affine map F(x) = fa + g x
                if (a.mode == IMPALA_MODE_REFERENCE) {
                    const float delta = rho * (r + a.gamma * v_n - v0);        // :126
                    fa = delta - g * v_n;                                      // :130
                } else {
                    fa = rho * (r + disc * v_n - v_t);
                }
                // inclusive suffix composition over the lanes: (A,G) <- F_lane o ... o F_31
                float sa = fa, sg = g;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const float a2 = __shfl_down_sync(IMPALA_FULL_MASK, sa, off);
                    const float g2 = __shfl_down_sync(IMPALA_FULL_MASK, sg, off);
                    if (lane + off < 32) {
                        sa = fmaf(sg, a2, sa);
                        sg = sg * g2;
                    }
                }
                const float acc_t = fmaf(sg, carry, sa);
                float acc_n = __shfl_down_sync(IMPALA_FULL_MASK, acc_t, 1);
                if (lane == 31) acc_n = carry;
                carry = __shfl_sync(IMPALA_FULL_MASK, acc_t, 0);
                const float vs_n = acc_n + v_n;                                // :131
                const float pg = rho * (r + disc * vs_n - v_t);                // :135
                s_vs[tt * kColStride + w] = (t <= L) ? acc_t + v_t : 0.f;
                s_r[tt * kColStride + w] = pg;  // rho == 0 on padding
                if (WITH_LOSS) {
                    // d total / d v = v_loss_c (v - vs) / B = -v_loss_c acc / B  (:149, :306-307)
                    s_dv[tt * kColStride + w] = valid ? -a.v_loss_c * a.inv_batch * acc_t : 0.f;
                    float ent = 0.f;
                    float lz[AP], pk[AP];
#pragma unroll
                    for (int k = 0; k < AP; ++k) {
                        lz[k] = (z[k] - lse) * kLn2;
                        pk[k] = (k < A) ? ex2f(z[k] - lse) : 0.f;
                        if (k < A) ent -= pk[k] * lz[k];                       // :310-314, :153
                    }
#pragma unroll
                    for (int k = 0; k < AP; ++k) {
                        const float onehot = (k == act) ? 1.f : 0.f;
                        const float dz = a.inv_batch * (a.policy_loss_c * pg * (pk[k] - onehot) +
                                                        a.entropy_c * pk[k] * (lz[k] + ent));
                        s_cur[tt * LS + w * AP + k] = (valid && k < A) ? dz : 0.f;
                    }
                    if (valid) {
                        sum_vl += 0.5 * (double)acc_t * (double)acc_t;
                        sum_pl += (double)(-lp_cur * pg);                      // :317-321
                        sum_ent += (double)ent;
                        sum_rw += (double)r;                                   // :108
                    }
                }
            }
            // the row one past this chunk (index t0+TC) belongs to the next chunk, except the
            // bootstrap row T when it is exactly the last chunk's extra row
            if (c == nchunks - 1 && t0 + TC == T && lane == 0) {
                s_vs[TC * kColStride + w] = (L == T) ? s_v[TC * kColStride + w] : 0.f;
                if (WITH_LOSS) s_dv[TC * kColStride + w] = 0.f;
            }
        }
        __syncthreads();

        // ---- store: row-contiguous global writes ----
        const int rows_v = (c == nchunks - 1 && t0 + TC == T) ? TC + 1 : TC;
        if constexpr (VEC) {
            for (int idx = tid; idx < rows_v * 2; idx += kThreads) {
                const int tt = idx >> 1, h = idx & 1, t = t0 + tt;
                if (t <= T) {
                    const size_t g = (size_t)t * B + b0;
                    const float* sv = s_vs + tt * kColStride + 4 * h;
                    if (a.vs) reinterpret_cast<float4*>(a.vs + g)[h] = make_float4(sv[0], sv[1], sv[2], sv[3]);
                    if (WITH_LOSS) {
                        const float* sd = s_dv + tt * kColStride + 4 * h;
                        reinterpret_cast<float4*>(a.dv + g)[h] = make_float4(sd[0], sd[1], sd[2], sd[3]);
                    }
                }
            }
            if (a.pg_adv) {
                for (int idx = tid; idx < TC * 2; idx += kThreads) {
                    const int tt = idx >> 1, h = idx & 1, t = t0 + tt;
                    if (t < T) {
                        const float* sp = s_r + tt * kColStride + 4 * h;
                        reinterpret_cast<float4*>(a.pg_adv + (size_t)t * B + b0)[h] =
                            make_float4(sp[0], sp[1], sp[2], sp[3]);
                    }
                }
            }
            if (WITH_LOSS) {
                const int q4 = 2 * A;
                for (int idx = tid; idx < TC * q4; idx += kThreads) {
                    const int tt = idx / q4, q = idx - tt * q4, t = t0 + tt;
                    if (t < T) {
                        float o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int rem = 4 * q + e, col = rem / A, k = rem - col * A;
                            o[e] = s_cur[tt * LS + col * AP + k];
                        }
                        reinterpret_cast<float4*>(a.dlogits + ((size_t)t * B + b0) * A)[q] =
                            make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
            }
        } else {
            for (int idx = tid; idx < rows_v * kTraj; idx += kThreads) {
                const int tt = idx / kTraj, col = idx - tt * kTraj;
                const int t = t0 + tt;
                if (t <= T && col < nb) {
                    const size_t g = (size_t)t * B + b0 + col;
                    if (a.vs) a.vs[g] = s_vs[tt * kColStride + col];
                    if (WITH_LOSS) a.dv[g] = s_dv[tt * kColStride + col];
                }
            }
            if (a.pg_adv) {
                for (int idx = tid; idx < TC * kTraj; idx += kThreads) {
                    const int tt = idx / kTraj, col = idx - tt * kTraj;
                    const int t = t0 + tt;
                    if (t < T && col < nb) a.pg_adv[(size_t)t * B + b0 + col] = s_r[tt * kColStride + col];
                }
            }
            if (WITH_LOSS) {
                const int rowlen = nb * A;
                for (int idx = tid; idx < TC * rowlen; idx += kThreads) {
                    const int tt = idx / rowlen, rem = idx - tt * rowlen;
                    const int t = t0 + tt;
                    const int col = rem / A, k = rem - col * A;
                    if (t < T) a.dlogits[((size_t)t * B + b0) * A + rem] = s_cur[tt * LS + col * AP + k];
                }
            }
        }
    }

    if (WITH_LOSS) {
        // per-CTA sums -> workspace; the last CTA to arrive adds them up in a fixed order
        // (bitwise reproducible, no float64 atomics, no memset node) and re-arms the counter.
        __shared__ bool s_last;
        __shared__ double s_fin[64][4];
        sum_vl = warp_sum_f64(sum_vl);
        sum_pl = warp_sum_f64(sum_pl);
        sum_ent = warp_sum_f64(sum_ent);
        sum_rw = warp_sum_f64(sum_rw);
        if (lane == 0) {
            s_red[w][0] = sum_vl, s_red[w][1] = sum_pl, s_red[w][2] = sum_ent, s_red[w][3] = sum_rw;
        }
        __syncthreads();
        if (tid < 4) {
            double s = 0.0;
            for (int i = 0; i < kTraj; ++i) s += s_red[i][tid];
            a.partials[(size_t)blockIdx.x * 4 + tid] = s;
            __threadfence();
        }
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            const int which = tid & 3, stripe = tid >> 2;  // 64 stripes x 4 scalars
            double s = 0.0;
            for (unsigned cta = stripe; cta < gridDim.x; cta += 64)
                s += __ldcg(a.partials + (size_t)cta * 4 + which);
            s_fin[stripe][which] = s;
            __syncthreads();
            if (tid < 4) {
                double tot = 0.0;
                for (int i = 0; i < 64; ++i) tot += s_fin[i][tid];
                a.scalars[tid] = tot * (double)a.inv_batch;
            }
            if (tid == 0) *a.counter = 0u;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Fast path for A == 4 and B % 8 == 0 (the benchmark shapes): same algorithm, but the unroll is
// walked in 32-step chunks through a 3-stage cp.async pipeline, so every CTA always has the
// next two chunks of its five input tensors in flight while one warp-pass of math runs -
// the stand-alone scan is HBM-latency bound otherwise (one CTA = 8 trajectories = ~36 KB of
// loads per 100 steps).  Logits live in shared memory as float4 with an XOR swizzle
// ([tt][w ^ (tt & 7)]): 16-byte cp.async in, conflict-free 128-bit column reads out.
// ------------------------------------------------------------------------------------------------
constexpr int kFStages = 3;

struct __align__(16) FastStage {
    float4 cur[32 * kTraj];
    float4 beh[32 * kTraj];
    float v[33 * kColStride + 3];
    float r[32 * kColStride];
    int act[32 * kColStride];
    uint2 done[32];
};
struct __align__(16) FastOut {
    float4 dl[32 * kTraj];
    float vs[33 * kColStride + 3];
    float dv[33 * kColStride + 3];
    float pg[32 * kColStride];
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
    const unsigned n = valid ? 16u : 0u;  // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_addr(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src, bool valid) {
    const unsigned n = valid ? 8u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_addr(dst)), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src, bool valid) {
    const unsigned n = valid ? 4u : 0u;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_addr(dst)), "l"(src), "r"(n) : "memory");
}

template <bool WITH_LOSS>
__global__ void __launch_bounds__(kThreads, 4) vtrace_fast_kernel(VtArgs a) {
    extern __shared__ __align__(16) unsigned char fsmem[];
    FastStage* stages = reinterpret_cast<FastStage*>(fsmem);
    FastOut* out = reinterpret_cast<FastOut*>(fsmem + kFStages * sizeof(FastStage));
    __shared__ double s_red[kTraj][4];

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int T = a.T, B = a.B;
    const int b0 = blockIdx.x * kTraj, b = b0 + w;
    const int L = min(max(__ldg(a.lens + b), 0), T);
    const float v0 = __ldg(a.v + b);  // V(x_0): the reference's v[:1]
    const int nch = (T + 31) >> 5;

    // Staging role of this thread (fixed for the kernel): time row ltt, column lcol of every chunk.
    // Global pointers are set to the LAST chunk and bumped backwards by 32 rows per chunk, so the
    // steady state issues its cp.asyncs without any 64-bit index arithmetic.
    const int ltt = tid >> 3, lcol = tid & 7;
    const int lslot = ltt * kTraj + (lcol ^ (ltt & 7));
    const int lrow9 = ltt * kColStride + lcol;
    const size_t lrow = (size_t)((nch - 1) * 32 + ltt) * B + b0 + lcol;  // element (t, b0 + lcol)
    const float4* p_cur = reinterpret_cast<const float4*>(a.cur_logits) + lrow;
    const float4* p_beh = reinterpret_cast<const float4*>(a.beh_logits) + lrow;
    const float* p_r = a.rewards + lrow;
    const int32_t* p_act = a.actions + lrow;
    const float* p_v = a.v + lrow;
    const float* p_v33 = a.v + (size_t)(nch * 32) * B + b0 + (tid & 7);       // threads 0..7
    const uint8_t* p_done = a.done + (size_t)((nch - 1) * 32 + (tid & 31)) * B + b0;  // threads 32..63
    const ptrdiff_t step = (ptrdiff_t)32 * B;  // elements per 32 time steps

    auto issue_chunk = [&](int c, int stg) {
        FastStage& st = stages[stg];
        // only the last chunk can reach past the end of the unroll
        const int t = c * 32 + ltt;
        const bool ok = t < T, okv = t <= T;
        cp_async16(&st.cur[lslot], ok ? p_cur : reinterpret_cast<const float4*>(a.cur_logits), ok);
        cp_async16(&st.beh[lslot], ok ? p_beh : reinterpret_cast<const float4*>(a.beh_logits), ok);
        cp_async4(&st.r[lrow9], ok ? p_r : a.rewards, ok);
        cp_async4(&st.act[lrow9], ok ? p_act : a.actions, ok);
        cp_async4(&st.v[lrow9], okv ? p_v : a.v, okv);
        if (tid < kTraj) {  // 33rd row of v (first row of the chunk processed before this one)
            const bool ok33 = c * 32 + 32 <= T;
            cp_async4(&st.v[32 * kColStride + tid], ok33 ? p_v33 : a.v, ok33);
        } else if (tid >= 32 && tid < 64) {
            const bool okd = c * 32 + (tid - 32) < T;
            cp_async8(&st.done[tid - 32], okd ? p_done : a.done, okd);
        }
        p_cur -= step, p_beh -= step, p_r -= step, p_act -= step, p_v -= step, p_v33 -= step, p_done -= step;
    };

    for (int k = 0; k < kFStages - 1; ++k) {
        if (nch - 1 - k >= 0) issue_chunk(nch - 1 - k, k);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    // compute role: lane = time step inside the chunk, warp = trajectory
    const int crow9 = lane * kColStride + w, cslot = lane * kTraj + (w ^ (lane & 7));
    int stg_c = 0, stg_n = kFStages - 1;  // stage being consumed / stage being refilled
    // output pointers of this thread's store roles, bumped backwards like the inputs
    const int s_tt = tid >> 1, s_h = tid & 1;  // vs / dv: threads 0..65
    float* q_vs = a.vs + (size_t)((nch - 1) * 32 + s_tt) * B + b0 + 4 * s_h;
    float* q_dv = WITH_LOSS ? a.dv + (size_t)((nch - 1) * 32 + s_tt) * B + b0 + 4 * s_h : nullptr;
    const int p_tt = (tid - 128) >> 1, p_h = tid & 1;  // pg: threads 128..191
    float* q_pg = a.pg_adv ? a.pg_adv + (size_t)((nch - 1) * 32 + p_tt) * B + b0 + 4 * p_h : nullptr;
    float4* q_dl = WITH_LOSS ? reinterpret_cast<float4*>(a.dlogits) + lrow : nullptr;

    double sum_vl = 0.0, sum_pl = 0.0, sum_ent = 0.0, sum_rw = 0.0;
    float carry = 0.f;  // acc at the first index after the current chunk
    for (int ci = 0; ci < nch; ++ci) {
        const int c = nch - 1 - ci, t0 = c * 32;
        asm volatile("cp.async.wait_group %0;" ::"n"(kFStages - 2) : "memory");
        __syncthreads();  // chunk c is in shared memory; the out tile and stage (ci-1)%S are free
        {
            const int cn = c - (kFStages - 1);
            if (cn >= 0) issue_chunk(cn, stg_n);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const FastStage& st = stages[stg_c];
        stg_n = stg_c;
        stg_c = stg_c + 1 == kFStages ? 0 : stg_c + 1;
        {
            const int tt = lane, t = t0 + tt;
            const bool valid = t < L;
            const float r = st.r[crow9];
            const int act = st.act[crow9];
            const bool dn = (reinterpret_cast<const unsigned char*>(&st.done[tt]))[w] != 0;
            const float v_t = st.v[crow9];
            const float v_n = st.v[crow9 + kColStride];
            const int slot = cslot;
            const float4 zc = st.cur[slot], zbv = st.beh[slot];
            // base-2 domain: zs = z log2(e); log-softmax_k = (zs_k - lse2) ln 2   (learner.py:298-303)
            const float z[4] = {zc.x * kLog2e, zc.y * kLog2e, zc.z * kLog2e, zc.w * kLog2e};
            const float zb[4] = {zbv.x * kLog2e, zbv.y * kLog2e, zbv.z * kLog2e, zbv.w * kLog2e};
            const float mx = fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3]));
            const float mxb = fmaxf(fmaxf(zb[0], zb[1]), fmaxf(zb[2], zb[3]));
            const float se = (ex2f(z[0] - mx) + ex2f(z[1] - mx)) + (ex2f(z[2] - mx) + ex2f(z[3] - mx));
            const float seb = (ex2f(zb[0] - mxb) + ex2f(zb[1] - mxb)) + (ex2f(zb[2] - mxb) + ex2f(zb[3] - mxb));
            const float lse = mx + lg2f(se), lseb = mxb + lg2f(seb);
            const float z_a = act == 1 ? z[1] : (act == 2 ? z[2] : (act == 3 ? z[3] : z[0]));
            const float zb_a = act == 1 ? zb[1] : (act == 2 ? zb[2] : (act == 3 ? zb[3] : zb[0]));
            const float lp2_cur = z_a - lse, lp2_beh = zb_a - lseb;      // log2 pi(a)
            const float lp_cur = lp2_cur * kLn2;
            const float ratio = ex2f(lp2_cur - lp2_beh);                   // :121-123
            const float rho = valid ? fminf(ratio, a.rho_bar) : 0.f;       // :124
            const float cc = valid ? fminf(ratio, a.c_bar) : 0.f;          // :125
            const float disc = (valid && !dn) ? a.gamma : 0.f;             // :109
            const float g = disc * cc;
            float fa;  // affine map F(x) = fa + g x
            if (a.mode == IMPALA_MODE_REFERENCE) {
                const float delta = rho * (r + a.gamma * v_n - v0);        // :126
                fa = delta - g * v_n;                                      // :130
            } else {
                fa = rho * (r + disc * v_n - v_t);
            }
            float sa = fa, sg = g;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const float a2 = __shfl_down_sync(IMPALA_FULL_MASK, sa, off);
                const float g2 = __shfl_down_sync(IMPALA_FULL_MASK, sg, off);
                if (lane + off < 32) {
                    sa = fmaf(sg, a2, sa);
                    sg = sg * g2;
                }
            }
            const float acc_t = fmaf(sg, carry, sa);
            float acc_n = __shfl_down_sync(IMPALA_FULL_MASK, acc_t, 1);
            if (lane == 31) acc_n = carry;
            carry = __shfl_sync(IMPALA_FULL_MASK, acc_t, 0);
            const float vs_n = acc_n + v_n;                                // :131
            const float pg = rho * (r + disc * vs_n - v_t);                // :135
            out->vs[crow9] = (t <= L) ? acc_t + v_t : 0.f;
            out->pg[crow9] = pg;  // rho == 0 on padding
            if (WITH_LOSS) {
                out->dv[crow9] = valid ? -a.v_loss_c * a.inv_batch * acc_t : 0.f;
                float lz[4], pk[4], ent = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    pk[k] = ex2f(z[k] - lse);
                    lz[k] = (z[k] - lse) * kLn2;
                    ent -= pk[k] * lz[k];                                  // :310-314, :153
                }
                float dz[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float onehot = (k == act) ? 1.f : 0.f;
                    const float d = a.inv_batch * (a.policy_loss_c * pg * (pk[k] - onehot) +
                                                   a.entropy_c * pk[k] * (lz[k] + ent));
                    dz[k] = valid ? d : 0.f;
                }
                out->dl[slot] = make_float4(dz[0], dz[1], dz[2], dz[3]);
                if (valid) {
                    sum_vl += 0.5 * (double)acc_t * (double)acc_t;
                    sum_pl += (double)(-lp_cur * pg);                      // :317-321
                    sum_ent += (double)ent;
                    sum_rw += (double)r;                                   // :108
                }
            }
            // bootstrap row T when it is exactly the extra (33rd) row of the last chunk
            if (ci == 0 && t0 + 32 == T && lane == 0) {
                out->vs[32 * kColStride + w] = (L == T) ? st.v[32 * kColStride + w] : 0.f;
                if (WITH_LOSS) out->dv[32 * kColStride + w] = 0.f;
            }
        }
        __syncthreads();
        // ---- row-contiguous 128-bit stores of this chunk's outputs
        const int rows_v = (ci == 0 && t0 + 32 == T) ? 33 : 32;
        if (tid < rows_v * 2) {
            if (t0 + s_tt <= T) {
                const float* sv = out->vs + s_tt * kColStride + 4 * s_h;
                *reinterpret_cast<float4*>(q_vs) = make_float4(sv[0], sv[1], sv[2], sv[3]);
                if (WITH_LOSS) {
                    const float* sd = out->dv + s_tt * kColStride + 4 * s_h;
                    *reinterpret_cast<float4*>(q_dv) = make_float4(sd[0], sd[1], sd[2], sd[3]);
                }
            }
        } else if (tid >= 128 && tid < 192 && q_pg) {
            if (t0 + p_tt < T) {
                const float* sp = out->pg + p_tt * kColStride + 4 * p_h;
                *reinterpret_cast<float4*>(q_pg) = make_float4(sp[0], sp[1], sp[2], sp[3]);
            }
        }
        if (WITH_LOSS) {
            if (t0 + ltt < T) *q_dl = out->dl[lslot];
            q_dl -= step;
            q_dv -= step;
        }
        q_vs -= step;
        if (q_pg) q_pg -= step;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    if (WITH_LOSS) {
        __shared__ bool s_last;
        __shared__ double s_fin[64][4];
        sum_vl = warp_sum_f64(sum_vl);
        sum_pl = warp_sum_f64(sum_pl);
        sum_ent = warp_sum_f64(sum_ent);
        sum_rw = warp_sum_f64(sum_rw);
        if (lane == 0) {
            s_red[w][0] = sum_vl, s_red[w][1] = sum_pl, s_red[w][2] = sum_ent, s_red[w][3] = sum_rw;
        }
        __syncthreads();
        if (tid < 4) {
            double s = 0.0;
            for (int i = 0; i < kTraj; ++i) s += s_red[i][tid];
            a.partials[(size_t)blockIdx.x * 4 + tid] = s;
            __threadfence();
        }
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(a.counter, 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            const int which = tid & 3, stripe = tid >> 2;
            double s = 0.0;
            for (unsigned cta = stripe; cta < gridDim.x; cta += 64)
                s += __ldcg(a.partials + (size_t)cta * 4 + which);
            s_fin[stripe][which] = s;
            __syncthreads();
            if (tid < 4) {
                double tot = 0.0;
                for (int i = 0; i < 64; ++i) tot += s_fin[i][tid];
                a.scalars[tid] = tot * (double)a.inv_batch;
            }
            if (tid == 0) *a.counter = 0u;
        }
    }
}

int pick_ap(int A) {
    if (A <= 2) return 2;
    if (A <= 4) return 4;
    if (A <= 8) return 8;
    if (A <= 16) return 16;
    return 0;
}

size_t smem_bytes(int TC, int AP) {
    const size_t LS = kTraj * AP + 1;
    return ((size_t)3 * (TC + 1) * kColStride + (size_t)2 * TC * kColStride + (size_t)2 * TC * LS) *
           sizeof(float);
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <bool WITH_LOSS>
int launch(VtArgs& a, cudaStream_t st) {
    if (a.T < 1 || a.B < 1 || a.A < 1) return IMPALA_ERR_BAD_ARG;
    const int AP = pick_ap(a.A);
    if (!AP) return IMPALA_ERR_UNSUPPORTED_SHAPE;
    int tc_max = AP <= 4 ? 128 : (AP == 8 ? 64 : 32);
    int tc = (int)impala_round_up(a.T, 32);
    a.TC = tc < tc_max ? tc : tc_max;
    const size_t smem = smem_bytes(a.TC, AP);
    const unsigned grid = (unsigned)((a.B + kTraj - 1) / kTraj);
    const bool vec = a.B % kTraj == 0 && aligned16(a.cur_logits) && aligned16(a.beh_logits) &&
                     aligned16(a.actions) && aligned16(a.rewards) && aligned16(a.v) &&
                     aligned16(a.vs) && aligned16(a.pg_adv) && aligned16(a.dlogits) &&
                     aligned16(a.dv) && (reinterpret_cast<uintptr_t>(a.done) & 3) == 0;
    {
        const char* fenv = std::getenv("IMPALA_VTRACE_FAST");
        if (vec && a.A == 4 && (reinterpret_cast<uintptr_t>(a.done) & 7) == 0 && !(fenv && fenv[0] == '0')) {
            const size_t fsmem = kFStages * sizeof(FastStage) + sizeof(FastOut);
            vtrace_fast_kernel<WITH_LOSS><<<grid, kThreads, fsmem, st>>>(a);
            return impala_launch_status();
        }
    }
#define VT_LAUNCH(APV)                                   \
    if (vec && a.A == APV) VT_LAUNCH_V(APV, true, true)  \
    else if (vec) VT_LAUNCH_V(APV, true, false)          \
    else VT_LAUNCH_V(APV, false, false)
#define VT_LAUNCH_V(APV, VECV, AEX)                                                                   \
    {                                                                                            \
        auto k = vtrace_kernel<APV, WITH_LOSS, VECV, AEX>;                                            \
        static size_t opted_in = 48 * 1024; /* per instantiation; avoids API calls in capture */ \
        if (smem > opted_in) {                                                                   \
            cudaError_t e =                                                                      \
                cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return (int)e;                                                 \
            opted_in = smem;                                                                     \
        }                                                                                        \
        k<<<grid, kThreads, smem, st>>>(a);                                                      \
    }
    switch (AP) {
        case 2: VT_LAUNCH(2) break;
        case 4: VT_LAUNCH(4) break;
        case 8: VT_LAUNCH(8) break;
        default: VT_LAUNCH(16) break;
    }
#undef VT_LAUNCH
#undef VT_LAUNCH_V
    return impala_launch_status();
}

}  // namespace

extern "C" int impala_vtrace(const float* cur_logits, const float* beh_logits,
                             const int32_t* actions, const float* rewards, const uint8_t* done,
                             const int32_t* lens, const float* v, float* vs, float* pg_adv, int T,
                             int B, int A, float gamma, float rho_bar, float c_bar, int mode,
                             void* stream) {
    if (!cur_logits || !beh_logits || !actions || !rewards || !done || !lens || !v || !vs || !pg_adv)
        return IMPALA_ERR_BAD_ARG;
    if (mode != IMPALA_MODE_REFERENCE && mode != IMPALA_MODE_PAPER) return IMPALA_ERR_BAD_ARG;
    VtArgs a{};
    a.cur_logits = cur_logits, a.beh_logits = beh_logits, a.actions = actions, a.rewards = rewards;
    a.done = done, a.lens = lens, a.v = v, a.vs = vs, a.pg_adv = pg_adv;
    a.T = T, a.B = B, a.A = A, a.mode = mode;
    a.gamma = gamma, a.rho_bar = rho_bar, a.c_bar = c_bar;
    return launch<false>(a, (cudaStream_t)stream);
}

extern "C" int64_t impala_vtrace_loss_workspace(int T, int B, int A) {
    if (T < 1 || B < 1 || A < 1) return IMPALA_ERR_BAD_ARG;
    const int64_t grid = (B + kTraj - 1) / kTraj;
    return grid * 4 * (int64_t)sizeof(double) + 16;  // per-CTA sums + arrival counter
}

extern "C" int impala_vtrace_loss(const float* cur_logits, const float* beh_logits,
                                  const int32_t* actions, const float* rewards,
                                  const uint8_t* done, const int32_t* lens, const float* v,
                                  float* vs, float* pg_adv, float* dlogits, float* dv,
                                  double* scalars, void* workspace, int64_t workspace_bytes, int T,
                                  int B, int A, float gamma, float rho_bar, float c_bar,
                                  float v_loss_c, float policy_loss_c, float entropy_c,
                                  float inv_batch, int mode, void* stream) {
    if (!cur_logits || !beh_logits || !actions || !rewards || !done || !lens || !v || !dlogits ||
        !dv || !scalars || !workspace)
        return IMPALA_ERR_BAD_ARG;
    if (mode != IMPALA_MODE_REFERENCE && mode != IMPALA_MODE_PAPER) return IMPALA_ERR_BAD_ARG;
    const int64_t need = impala_vtrace_loss_workspace(T, B, A);
    if (need < 0) return (int)need;
    if (workspace_bytes < need) return IMPALA_ERR_WORKSPACE_TOO_SMALL;
    if (reinterpret_cast<uintptr_t>(workspace) & 15) return IMPALA_ERR_BAD_ARG;
    VtArgs a{};
    a.cur_logits = cur_logits, a.beh_logits = beh_logits, a.actions = actions, a.rewards = rewards;
    a.done = done, a.lens = lens, a.v = v, a.vs = vs, a.pg_adv = pg_adv, a.dlogits = dlogits;
    a.dv = dv, a.scalars = scalars;
    // workspace = [counter (16 bytes) | per-CTA sums]
    a.counter = reinterpret_cast<unsigned int*>(workspace);
    a.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 16);
    a.T = T, a.B = B, a.A = A, a.mode = mode;
    a.gamma = gamma, a.rho_bar = rho_bar, a.c_bar = c_bar;
    a.v_loss_c = v_loss_c, a.policy_loss_c = policy_loss_c, a.entropy_c = entropy_c;
    a.inv_batch = inv_batch;
    return launch<true>(a, (cudaStream_t)stream);
}
