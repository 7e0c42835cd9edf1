// "Wide" MLP forward / backward on the 5th-gen tensor cores (tcgen05 + TMEM, 3xTF32) for the shapes the
// kernels of mlp_fwd_tc.cu / mlp_bwd_tc.cu do not take: observation widths up to 64 (two 128-byte
// swizzle atoms of K) and hidden layers that are a multiple of 128 units of any size - BASELINE
// config c5 (obs = 64, hidden = 512), where the first layer is a real GEMM (models.py:13-18,41-46 at
// learner.py:112-113,175) and the step is bound by tensor-core operand traffic, not by the epilogues.
//
// W1' hi + lo of the WHOLE hidden layer no longer fits shared memory beside the operand stages
// (512 x 64 x 4 B x 2 = 256 KiB), so a persistent CTA walks the hidden layer in BLOCKS of 128 units:
// for every block it stages that block's W1 rows once and then streams all of its row tiles past it.
// x is re-read once per block (4 x 212 MB at c5; one thread issues cp.async.bulk.prefetch.L2 three
// tiles ahead so the producers' register loads hit L2), in exchange nothing is exchanged between CTAs:
//   forward   out[row] = (b2 + z_0) + z_1 + ...: the epilogue thread that owns a row adds the block's
//             partial second-layer sum to what the SAME thread wrote in the previous pass (fixed
//             order, bitwise reproducible, no workspace);
//   backward  every gradient entry of W1 / b1 / W2 belongs to exactly one hidden block, so a pass
//             produces final per-CTA partial sums for its block (dW1' accumulates in TMEM over all
//             row tiles of the pass and is read out once per pass).
// The bias is applied by the epilogues (an extra K atom would cost 48 KiB of shared memory).
// Operand split and UMMA descriptors are those of the narrow kernels (tc_common.cuh).
//
// What bounds these kernels (measured on B200, c5, ncu + CUDA events): a tf32 UMMA
// covers only K = 8, so per instruction it moves M x 32 B of A and N x 32 B of B, and both operand
// sources - the shared-memory descriptor fetch and tensor-memory reads (UMMA A operand AND the
// epilogues' tcgen05.ld share that port) - deliver about 64 B/clk per SM.  Forward with both operands in
// shared memory: 24 UMMAs x 8 KiB = 192 KiB -> 3 072 cycles per (128-row tile, block) against 1 536 of
// math; measured 3 200 (value net) / 3 680 (policy).  Putting X' into tensor memory (TS form, producer
// thread = row writing tcgen05.st) was measured SLOWER (3 850 / 4 960 cycles), and so was the split
// "x_hi from tensor memory, x_lo from shared memory" that halves the traffic of either port
// (4 250 / 4 950): in the forward the A reads queue behind the epilogue's 64 KiB of accumulator reads
// per tile on the same port, and the row-per-thread producer loses the coalesced loads.  Backward: UMMA1 with W1' in shared memory fetched 6 KiB per
// 32 cycles of math (4 110 cycles per (64-row tile, block)); with W1' in TENSOR memory (TS, like DP for
// UMMA2) all A operands and the epilogue share the TMEM port: 192 KiB -> 3 072, measured 3 100 / 3 730.
// (Taking W1'_lo from shared memory again for the lo*hi product - 160 / 128 KiB - with two x stages: 3 330 / 3 850.)
//
// Forward, per 128-row tile and hidden block:   D[128, 128] = X'[128, 64] * W1'_blk[128, 64]^T   (SS)
//   warps 0-15  epilogue (TMEM lane = row; four warps per lane quarter take 32 hidden units each)
//   warps 16-23 producer: coalesced 128-bit global loads of the raw x tile (next tile requested into
//               registers before the wait for the stage) -> hi/lo split -> swizzled tiles, 2 stages
//   warp  24    TMEM allocator + UMMA issuer (24 UMMAs M128 N128 K8 per tile at O = 64)
// Backward, per 64-row tile and hidden block (thread owns a hidden unit, see mlp_bwd_tc.cu):
//   UMMA1  PRE[128, 64]  = W1'_blk[128, 64] * X'[64, 64]^T       (TS: W1' hi / lo in TMEM, recompute)
//   CUDA   h = relu(PRE + b1); dh = W2^T dz; dW2 += dz h; DP = PRE + b1 > 0 ? dh : 0; db1 += DP
//   UMMA2  dW1'_blk[128, 64 | 64] += DP[128, 64] * [X'^T_hi ; X'^T_lo]    (TS: DP hi / lo from TMEM)
//   The row-major x tile (UMMA1) and the transposed one (UMMA2) have separate full / empty
//   barriers: the former is free again as soon as UMMA1 has retired.  DP_lo is double buffered like
//   PRE / DP_hi (UMMAs retire in issue order: when PRE of tile i + 2 has arrived, UMMA2 of tile i is
//   done with both).  TMEM: [0,128) PRE / DP_hi x 2, [128,256) DP_lo x 2, [256,384) dW1', [384,512) W1'.
//   warps 0-7 epilogue (thread = hidden unit x half of the tile's 64 rows; 13 warps leave each thread
//   128 registers), warps 8-11 producer (thread = (row, K atom): row-major stores, then the transposed
//   tile from what it has just written; dz; db2), warp 12 UMMA issuer.
// Accumulation order: the tensor core truncates when it adds into the fp32 accumulator, so the small
// lo*hi / hi*lo correction terms are accumulated first (or into their own columns) and hi*hi last.
#include <cstdlib>

#include "mlp_kernels.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kHB = 128;                         // hidden units per pass (UMMA M / N)
constexpr int kKA = 2;                           // K atoms of 32 floats: K' = 64
constexpr int kWAtomBytes = kHB * 128;           // 16 KiB
constexpr int kWBytes = 2 * kKA * kWAtomBytes;   // [hi a0][hi a1][lo a0][lo a1] = 64 KiB

// ---------------------------------------------------------------------------------------- helpers
__device__ __forceinline__ float4 ldg_v4(const float* p) {  // volatile: stays where it is written
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_f(const float* p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void split4(const float4& v, float4& hi, float4& lo) {
    tc::split_tf32(v.x, hi.x, lo.x);
    tc::split_tf32(v.y, hi.y, lo.y);
    tc::split_tf32(v.z, hi.z, lo.z);
    tc::split_tf32(v.w, hi.w, lo.w);
}
// The same split for the streamed operand in 3 instead of 5 SASS instructions per element (the
// producers share issue slots with the epilogue): hi = x rounded to tf32, half away from zero (add half
// a tf32 ulp to the magnitude bits, clear the low 13), lo = x - hi exactly.  (|x| within 2^-11 of FLT_MAX
// rounds up to inf like any round-to-nearest conversion would; observations are nowhere near.)
__device__ __forceinline__ void split_fast(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = x - hi;
}
__device__ __forceinline__ void split4_fast(const float4& v, float4& hi, float4& lo) {
    split_fast(v.x, hi.x, lo.x);
    split_fast(v.y, hi.y, lo.y);
    split_fast(v.z, hi.z, lo.z);
    split_fast(v.w, hi.w, lo.w);
}
// `bytes` contiguous bytes (multiple of 16, 16-byte aligned) towards L2, no destination: issued by one
// thread a few tiles ahead, so that the producers' register loads are L2 hits (x is re-read once per
// hidden block and does not stay in L2 between passes: 212 MB at c5)
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Rows [128 hb, 128 hb + 128) of W1 (H, O) -> hi / lo K-major SWIZZLE_128B tiles (columns >= O zero).
// All global loads of a thread are issued before the first conversion (the rows are L2-resident
// parameters; a load -> convert -> store loop would serialise its round trips at every pass).
template <int NTHREADS>
__device__ __forceinline__ void stage_w_block(uint8_t* wt, const float* __restrict__ W1, int hb, int O, int tid) {
    constexpr int kIt = (kHB * 16 + NTHREADS - 1) / NTHREADS;
    const int ochunks = O >> 2;
    float4 v[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int idx = tid + it * NTHREADS, r = idx >> 4, c = idx & 15;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < kHB * 16 && c < ochunks) v[it] = __ldg(reinterpret_cast<const float4*>(W1 + (size_t)(hb * kHB + r) * O) + c);
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
        const int idx = tid + it * NTHREADS, r = idx >> 4, c = idx & 15;
        if (idx < kHB * 16) {
            float4 hi, lo;
            split4(v[it], hi, lo);
            const uint32_t off = (c >> 3) * kWAtomBytes + tc::sw128_offset(r, c & 7);
            *reinterpret_cast<float4*>(wt + off) = hi;
            *reinterpret_cast<float4*>(wt + 2 * kWAtomBytes + off) = lo;
        }
    }
}

// ======================================================================================= forward
constexpr int kFTileM = 128;
constexpr int kFAtomBytes = kFTileM * 128;             // 16 KiB
constexpr int kFStageBytes = 2 * kKA * kFAtomBytes;    // [hi a0][hi a1][lo a0][lo a1] = 64 KiB
constexpr int kFStages = 2;
constexpr int kFProdWarps = 8;
constexpr int kFThreads = (16 + kFProdWarps + 1) * 32;  // 800
constexpr int kFIssuer = 16 + kFProdWarps;
constexpr int kFAccCols = 128;

struct FwdWArgs {
    const float* x;
    const float* params;
    float* out;
    int M, O, H, N2, num_tiles;
    MlpLayout lay;
};

struct __align__(8) FBarriers {
    uint64_t full[kFStages], empty[kFStages], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

template <int NP>
__global__ void __launch_bounds__(kFThreads, 1) mlp_fwd_tcw_kernel(const __grid_constant__ FwdWArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* wt = smem;                                                     // 64 KiB
    uint8_t* xs = wt + kWBytes;                                             // kFStages x 64 KiB
    float* w2s = reinterpret_cast<float*>(xs + kFStages * kFStageBytes);    // [kHB][NP]
    float* b1s = w2s + kHB * NP;                                            // [kHB]
    float* part = b1s + kHB;                                                // [2][3][128 rows][NP]
    FBarriers* bars = reinterpret_cast<FBarriers*>(part + 2 * 3 * kFTileM * NP);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int nblk = a.H / kHB;
    const int n_my = (a.num_tiles - cta + ncta - 1) / ncta;

    if (warp == kFIssuer && lane == 0) {
        for (int s = 0; s < kFStages; ++s) {
            tc::mbar_init(&bars->full[s], kFProdWarps * 32);  // every producer thread arrives
            tc::mbar_init(&bars->empty[s], 1);                // tcgen05.commit
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&bars->acc_full[s], 1);         // tcgen05.commit
            tc::mbar_init(&bars->acc_empty[s], 16 * 32);  // every epilogue thread arrives
        }
        tc::mbar_fence_init();
    }
    if (warp == kFIssuer) tc::tmem_alloc(&bars->tmem_base, 256);

    // Every role runs the same pass loop (one pass per block of 128 hidden units): begin_pass stages the
    // block's weights with all threads and ends in a CTA barrier; end_pass is the CTA barrier after
    // which the next block may overwrite them - the epilogue has seen the last accumulator of the
    // pass by then, so every UMMA has retired.  The tile counter `it` keeps running across passes,
    // so the mbarrier phases simply continue.
    auto begin_pass = [&](int hb) {
        stage_w_block<kFThreads>(wt, a.params + a.lay.oW1, hb, a.O, tid);
        const float* __restrict__ W2 = a.params + a.lay.oW2;
        const float* __restrict__ b1 = a.params + a.lay.ob1;
        for (int idx = tid; idx < kHB * NP; idx += kFThreads) {
            const int j = idx / NP, n = idx - j * NP;
            w2s[idx] = n < a.N2 ? __ldg(W2 + (size_t)n * a.H + hb * kHB + j) : 0.f;
        }
        for (int idx = tid; idx < kHB; idx += kFThreads) b1s[idx] = __ldg(b1 + hb * kHB + idx);
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    };
    auto end_pass = [&]() {
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    };

    if (warp < 16) {
        // =============================== epilogue ===============================
        const int q = warp & 3, grp = warp >> 2;  // TMEM lane quarter, 32-column group
        const int c0 = 32 * grp, rl = 32 * q + lane;  // first hidden unit of the group, row of the tile
        const float* __restrict__ b2 = a.params + a.lay.ob2;
        int it = 0;
        for (int hb = 0; hb < nblk; ++hb) {
            begin_pass(hb);
            const uint32_t taddr0 = bars->tmem_base + (static_cast<uint32_t>(32 * q) << 16) + c0;
            for (int i = 0; i < n_my; ++i, ++it) {
                const int as = it & 1, aph = (it >> 1) & 1;
                // the sum of the previous passes for this row - written by this very thread (plain loads,
                // program order) - requested before the wait so that its latency is off the tile's chain
                const int row = (cta + i * ncta) * kFTileM + rl;
                float prev[NP];
                if (grp == 0 && row < a.M) {
#pragma unroll
                    for (int n = 0; n < NP; ++n) prev[n] = 0.f;
                    if (hb == 0) {
#pragma unroll
                        for (int n = 0; n < NP; ++n)
                            if (n < a.N2) prev[n] = __ldg(b2 + n);
                    } else {
                        bool vec = false;
                        if constexpr (NP == 4) {
                            if (a.N2 == 4) {
                                const float4 t = *reinterpret_cast<const float4*>(a.out + (size_t)row * 4);
                                prev[0] = t.x, prev[1] = t.y, prev[2] = t.z, prev[3] = t.w;
                                vec = true;
                            }
                        }
                        if (!vec) {
#pragma unroll
                            for (int n = 0; n < NP; ++n)
                                if (n < a.N2) prev[n] = a.out[(size_t)row * a.N2 + n];
                        }
                    }
                }
                tc::mbar_wait(&bars->acc_full[as], aph);
                tc::tc_fence_after();
                float raw[32];
                tc::tmem_ld32(taddr0 + as * kFAccCols, raw);
                tc::tc_fence_before();
                tc::mbar_arrive(&bars->acc_empty[as]);  // this thread's TMEM reads are complete
                float2 acc[NP == 4 ? 2 : 1];
#pragma unroll
                for (int k = 0; k < (NP == 4 ? 2 : 1); ++k) acc[k] = make_float2(0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 32; k += 2) {
                    const float2 bb = *reinterpret_cast<const float2*>(b1s + c0 + k);
                    const float h0 = fmaxf(raw[k] + bb.x, 0.f);
                    const float h1 = fmaxf(raw[k + 1] + bb.y, 0.f);
                    if constexpr (NP == 4) {
                        const float4 wa = *reinterpret_cast<const float4*>(w2s + (c0 + k) * 4);
                        const float4 wb = *reinterpret_cast<const float4*>(w2s + (c0 + k + 1) * 4);
                        const float2 h0p = make_float2(h0, h0), h1p = make_float2(h1, h1);
                        acc[0] = tc::ffma2(h0p, make_float2(wa.x, wa.y), acc[0]);
                        acc[1] = tc::ffma2(h0p, make_float2(wa.z, wa.w), acc[1]);
                        acc[0] = tc::ffma2(h1p, make_float2(wb.x, wb.y), acc[0]);
                        acc[1] = tc::ffma2(h1p, make_float2(wb.z, wb.w), acc[1]);
                    } else {
                        const float2 w = *reinterpret_cast<const float2*>(w2s + c0 + k);
                        acc[0] = tc::ffma2(make_float2(h0, h1), w, acc[0]);
                    }
                }
                float* pbuf = part + (it & 1) * 3 * kFTileM * NP;
                if (grp > 0) {
                    float* pb = pbuf + ((grp - 1) * kFTileM + rl) * NP;
                    if constexpr (NP == 4) *reinterpret_cast<float4*>(pb) = make_float4(acc[0].x, acc[0].y, acc[1].x, acc[1].y);
                    else pb[0] = acc[0].x + acc[0].y;
                }
                asm volatile("bar.sync 2, 512;" ::: "memory");  // the four column groups meet
                if (grp == 0 && row < a.M) {
                    if constexpr (NP == 4) {
                        const float4 p1 = *reinterpret_cast<const float4*>(pbuf + (0 * kFTileM + rl) * 4);
                        const float4 p2 = *reinterpret_cast<const float4*>(pbuf + (1 * kFTileM + rl) * 4);
                        const float4 p3 = *reinterpret_cast<const float4*>(pbuf + (2 * kFTileM + rl) * 4);
                        const float z[4] = {((acc[0].x + p1.x) + p2.x) + p3.x, ((acc[0].y + p1.y) + p2.y) + p3.y,
                                            ((acc[1].x + p1.z) + p2.z) + p3.z, ((acc[1].y + p1.w) + p2.w) + p3.w};
                        if (a.N2 == 4) {
                            *reinterpret_cast<float4*>(a.out + (size_t)row * 4) =
                                make_float4(prev[0] + z[0], prev[1] + z[1], prev[2] + z[2], prev[3] + z[3]);
                        } else {
#pragma unroll
                            for (int n = 0; n < 4; ++n)
                                if (n < a.N2) a.out[(size_t)row * a.N2 + n] = prev[n] + z[n];
                        }
                    } else {
                        const float z = (((acc[0].x + acc[0].y) + pbuf[rl]) + pbuf[kFTileM + rl]) + pbuf[2 * kFTileM + rl];
                        a.out[row] = prev[0] + z;
                    }
                }
            }
            end_pass();
        }
    } else if (warp < kFIssuer) {
        // =============================== producer ===============================
        // 256 threads x 8 chunks of 16 bytes: consecutive threads take consecutive chunks of the tile's
        // 128 x 16 chunk grid, so a warp reads two whole rows (512 contiguous bytes at O = 64)
        const int ptid = tid - 16 * 32, ochunks = a.O >> 2;
        float4 v[8];
        auto load = [&](int i) {
            const int tile = cta + i * ncta;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int idx = ptid + 256 * k, r = idx >> 4, c = idx & 15;
                const int row = tile * kFTileM + r;
                v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < a.M && c < ochunks) v[k] = ldg_v4(a.x + (size_t)row * a.O + 4 * c);
            }
        };
        auto prefetch = [&](int i) {  // one thread: the tile's rows are contiguous
            const int64_t row0 = (int64_t)(cta + i * ncta) * kFTileM;
            const int64_t rows = a.M - row0 < kFTileM ? a.M - row0 : kFTileM;
            if (i < n_my && rows > 0) prefetch_l2(a.x + row0 * a.O, (uint32_t)(rows * a.O * 4));
        };
        int it = 0;
        for (int hb = 0; hb < nblk; ++hb) {
            begin_pass(hb);
            if (hb == 0) {  // later passes: requested before the previous pass's closing barrier (below)
                if (ptid == 0) prefetch(1), prefetch(2);
                load(0);
            }
            for (int i = 0; i < n_my; ++i, ++it) {
                const int s = it % kFStages, ph = (it / kFStages) & 1;
                if (ptid == 0) prefetch(i + 3);
                tc::mbar_wait(&bars->empty[s], ph ^ 1);  // UMMAs that read this stage have retired
                uint8_t* th = xs + s * kFStageBytes;
                uint8_t* tl = th + 2 * kFAtomBytes;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int idx = ptid + 256 * k, r = idx >> 4, c = idx & 15;
                    float4 hi, lo;
                    split4_fast(v[k], hi, lo);
                    const uint32_t off = (c >> 3) * kFAtomBytes + tc::sw128_offset(r, c & 7);
                    *reinterpret_cast<float4*>(th + off) = hi;
                    *reinterpret_cast<float4*>(tl + off) = lo;
                }
                tc::fence_proxy_async();
                tc::mbar_arrive(&bars->full[s]);
                if (i + 1 < n_my) load(i + 1);  // in flight while this thread waits for the next stage
            }
            if (hb + 1 < nblk) {  // the next pass starts with the same tiles: its first rows travel during the
                if (ptid == 0) prefetch(1), prefetch(2);  // drain of this pass and the staging of the next block
                load(0);
            }
            end_pass();
        }
    } else {
        // =============================== UMMA issuer ===============================
        const uint32_t idesc = tc::instr_desc_tf32_m128(kHB);
        const int ksteps = (a.O + 7) >> 3;
        const uint64_t dw = tc::smem_desc_k_sw128(wt, 0), dx = tc::smem_desc_k_sw128(xs, 0);
        int it = 0;
        for (int hb = 0; hb < nblk; ++hb) {
            begin_pass(hb);
            const uint32_t tmem_base = bars->tmem_base;
            for (int i = 0; i < n_my; ++i, ++it) {
                const int s = it % kFStages, ph = (it / kFStages) & 1;
                const int as = it & 1, aph = (it >> 1) & 1;
                tc::mbar_wait(&bars->full[s], ph);
                tc::mbar_wait(&bars->acc_empty[as], aph ^ 1);
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint32_t d = tmem_base + as * kFAccCols;
                    const uint64_t xh = dx + static_cast<uint64_t>((s * kFStageBytes) >> 4);
                    const uint64_t xl = xh + static_cast<uint64_t>((2 * kFAtomBytes) >> 4);
                    const uint64_t wh = dw, wl = dw + static_cast<uint64_t>((2 * kWAtomBytes) >> 4);
                    // The tensor core truncates (rounds toward zero) when it adds into the fp32 accumulator:
                    // every accumulation at full magnitude costs ~half an ulp of bias (measured: gradient
                    // norms 1e-6 low with 24 of them).  So the small correction terms lo*hi + hi*lo of
                    // ALL K steps go first, the hi*hi terms last: 8 full-magnitude accumulations, not 24.
#pragma unroll
                    for (int kk = 0; kk < 4 * kKA; ++kk) {
                        if (kk < ksteps) {
                            // atom kk / 4 (16 KiB apart in both operands), 32 bytes per K = 8 step inside it
                            const uint64_t ko = static_cast<uint64_t>((kk >> 2) * (kFAtomBytes >> 4) + (kk & 3) * 2);
                            tc::umma_tf32(d, xl + ko, wh + ko, idesc, kk > 0);
                            tc::umma_tf32(d, xh + ko, wl + ko, idesc, true);
                        }
                    }
#pragma unroll
                    for (int kk = 0; kk < 4 * kKA; ++kk) {
                        if (kk < ksteps) {
                            const uint64_t ko = static_cast<uint64_t>((kk >> 2) * (kFAtomBytes >> 4) + (kk & 3) * 2);
                            tc::umma_tf32(d, xh + ko, wh + ko, idesc, true);
                        }
                    }
                    tc::umma_commit(&bars->empty[s]);
                    tc::umma_commit(&bars->acc_full[as]);
                }
                __syncwarp();
            }
            end_pass();
        }
    }
    if (warp == kFIssuer) tc::tmem_dealloc(bars->tmem_base, 256);
}

static_assert(kFAtomBytes == kWAtomBytes, "the forward issuer uses one K offset for both operands");

constexpr size_t fwd_smem_bytes(int np) {
    return 1024 + kWBytes + kFStages * kFStageBytes + (size_t)(kHB * np + kHB + 2 * 3 * kFTileM * np) * sizeof(float) +
           sizeof(FBarriers);
}

// ======================================================================================= backward
constexpr int kBRows = 64;                            // batch rows per tile: N of UMMA1, K of UMMA2
constexpr int kBXAtomBytes = kBRows * 128;            // 8 KiB
constexpr int kBXaBytes = 2 * kKA * kBXAtomBytes;     // row-major [hi a0][hi a1][lo a0][lo a1] = 32 KiB
constexpr int kBXtChunkBytes = 128 * 128;             // 32 batch rows of K: rows = 64 hi features | 64 lo features
constexpr int kBXtBytes = 2 * kBXtChunkBytes;         // 32 KiB
constexpr int kBStages = 3;
constexpr int kBProdWarps = 4;
constexpr int kBEpiWarps = 8;
constexpr int kBThreads = (kBEpiWarps + kBProdWarps + 1) * 32;  // 416: 13 warps leave 128 registers per thread
constexpr int kBIssuer = kBEpiWarps + kBProdWarps;
constexpr int kBColLo = 128;   // TMEM columns: [0,128) PRE / DP_hi x 2, [128,256) DP_lo x 2, [256,384) dW1',
constexpr int kBColAcc = 256;  // [384,512) this pass's W1 block: 64 hi + 64 lo columns (lane = hidden unit)
constexpr int kBColW = 384;

struct BwdWArgs {
    const float* x;
    const float* params;
    const float* dout;
    float* ws;
    int M, O, H, N2, num_tiles;
    MlpLayout lay;
};

struct __align__(8) BBarriers {
    uint64_t xa_full[kBStages], xa_empty[kBStages], xt_full[kBStages], xt_empty[kBStages];
    uint64_t d1_full[2], dp_full[2], done;
    float gb2_part[4];
    uint32_t tmem_base;
};

template <int NP>
__global__ void __launch_bounds__(kBThreads, 1) mlp_bwd_tcw_kernel(const __grid_constant__ BwdWArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* xa = smem;                                                      // kBStages x 32 KiB
    uint8_t* xtb = xa + kBStages * kBXaBytes;                                // kBStages x 32 KiB
    float* dzs = reinterpret_cast<float*>(xtb + kBStages * kBXtBytes);       // [kBStages][32 row pairs][NP][2]
    float* exch = dzs + kBStages * kBRows * NP;                              // [NP + 1][128]: sums of the odd half
    BBarriers* bars = reinterpret_cast<BBarriers*>(exch + (NP + 1) * kHB);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cta = blockIdx.x, ncta = gridDim.x;
    const int nblk = a.H / kHB;
    const int n_my = (a.num_tiles - cta + ncta - 1) / ncta;

    if (warp == kBIssuer && lane == 0) {
        for (int s = 0; s < kBStages; ++s) {
            tc::mbar_init(&bars->xa_full[s], kBProdWarps * 32);
            tc::mbar_init(&bars->xa_empty[s], 1);  // tcgen05.commit after UMMA1
            tc::mbar_init(&bars->xt_full[s], kBProdWarps * 32);
            tc::mbar_init(&bars->xt_empty[s], 1);  // tcgen05.commit after UMMA2
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&bars->d1_full[s], 1);        // tcgen05.commit after UMMA1
            tc::mbar_init(&bars->dp_full[s], kBEpiWarps * 32);  // every epilogue thread
        }
        tc::mbar_init(&bars->done, 1);
        tc::mbar_fence_init();
    }
    if (warp == kBIssuer) tc::tmem_alloc(&bars->tmem_base, 512);
    tc::tc_fence_before();
    __syncthreads();  // barriers initialised, TMEM base address published
    tc::tc_fence_after();

    // One pass per block of 128 hidden units, the same loop in every role (see the forward kernel);
    // the tile counter `it` keeps running across passes, so the mbarrier phases simply continue.
    // W1' of the pass is the A operand of UMMA1 and lives in TENSOR MEMORY (TS form, like DP for UMMA2):
    // with both operands in shared memory an M128 N64 K8 tf32 UMMA fetches 6 KiB at ~64 B/clk = 96 cycles
    // against 32 of math.  The epilogue warps stage it (thread = hidden unit = TMEM lane, the two row
    // halves take one K atom each): W1 row -> hi / lo -> tcgen05.st.
    auto begin_pass = [&](int hb) {
        if (warp < kBEpiWarps) {
            const int q = warp & 3, atom = warp >> 2, ochunks = a.O >> 2;
            const float* wrow = a.params + a.lay.oW1 + (size_t)(hb * kHB + 32 * q + lane) * a.O + 32 * atom;
            const uint32_t waddr = bars->tmem_base + (static_cast<uint32_t>(32 * q) << 16) + kBColW + 32 * atom;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float4 w = make_float4(0.f, 0.f, 0.f, 0.f), fh, fl;
                    if (8 * atom + 4 * h + c < ochunks) w = __ldg(reinterpret_cast<const float4*>(wrow) + 4 * h + c);
                    split4(w, fh, fl);
                    hi[4 * c] = __float_as_uint(fh.x), hi[4 * c + 1] = __float_as_uint(fh.y);
                    hi[4 * c + 2] = __float_as_uint(fh.z), hi[4 * c + 3] = __float_as_uint(fh.w);
                    lo[4 * c] = __float_as_uint(fl.x), lo[4 * c + 1] = __float_as_uint(fl.y);
                    lo[4 * c + 2] = __float_as_uint(fl.z), lo[4 * c + 3] = __float_as_uint(fl.w);
                }
                tc::tmem_st16(waddr + 16 * h, hi);
                tc::tmem_st16(waddr + 64 + 16 * h, lo);
            }
            tc::tmem_wait_st();
        }
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    };
    auto end_pass = [&]() {  // dW1' of the pass has been read out, every UMMA retired
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    };

    if (warp < kBEpiWarps) {
        // =============================== epilogue ===============================
        const int q = warp & 3, hh = warp >> 2;  // TMEM lane quarter, half of the tile's 64 rows
        const int jl = 32 * q + lane, O = a.O, H = a.H, ochunks = O >> 2;
        float* wsb = a.ws + (size_t)cta * a.lay.total;  // this CTA's partial gradient row
        int it = 0;
        for (int hb = 0; hb < nblk; ++hb) {
            begin_pass(hb);
            const int j = hb * kHB + jl;
            const float b1j = __ldg(a.params + a.lay.ob1 + j);
            const float2 b1p = make_float2(b1j, b1j);
            float2 w2p[NP], gw2p[NP], gb1p = make_float2(0.f, 0.f);
#pragma unroll
            for (int n = 0; n < NP; ++n) {
                const float w = n < a.N2 ? __ldg(a.params + a.lay.oW2 + (size_t)n * H + j) : 0.f;
                w2p[n] = make_float2(w, w);
                gw2p[n] = make_float2(0.f, 0.f);
            }
            const uint32_t lane_addr = bars->tmem_base + (static_cast<uint32_t>(32 * q) << 16);
            for (int i = 0; i < n_my; ++i, ++it) {
                const int s = it % kBStages, ph = (it / kBStages) & 1, d1 = it & 1, dph = (it >> 1) & 1;
                const uint32_t c_hi = lane_addr + d1 * 64 + 32 * hh;            // PRE in, DP_hi out
                const uint32_t c_lo = lane_addr + kBColLo + d1 * 64 + 32 * hh;  // DP_lo out
                tc::mbar_wait(&bars->xt_full[s], ph);   // dz rows of this tile are visible
                tc::mbar_wait(&bars->d1_full[d1], dph);  // PRE of this tile is in TMEM
                tc::tc_fence_after();
                const float* dz_h = dzs + (s * (kBRows / 2) + 16 * hh) * 2 * NP;  // [row pair][n][2]
                // the thread's 32 pre-activations as two halves of 16 columns: the second half is in flight
                // while the first is consumed, DP of the first is on its way back while the second is computed
                uint32_t va[16], vb[16], la[16], lb[16];
                auto half = [&](uint32_t (&v)[16], uint32_t (&lo)[16], const int pr0) {
#pragma unroll
                    for (int q2 = 0; q2 < 8; ++q2) {
                        const float* zp = dz_h + (pr0 + q2) * 2 * NP;
                        float2 dz[NP];
                        if constexpr (NP == 4) {
                            const float4 t0 = *reinterpret_cast<const float4*>(zp);
                            const float4 t1 = *reinterpret_cast<const float4*>(zp + 4);
                            dz[0] = make_float2(t0.x, t0.y), dz[1] = make_float2(t0.z, t0.w);
                            dz[2] = make_float2(t1.x, t1.y), dz[3] = make_float2(t1.z, t1.w);
                        } else {
                            dz[0] = *reinterpret_cast<const float2*>(zp);
                        }
                        const float2 pre = tc::fadd2(make_float2(__uint_as_float(v[2 * q2]), __uint_as_float(v[2 * q2 + 1])), b1p);
                        const float2 h = make_float2(fmaxf(pre.x, 0.f), fmaxf(pre.y, 0.f));
                        float2 dh = tc::fmul2(dz[0], w2p[0]);
#pragma unroll
                        for (int n = 1; n < NP; ++n) dh = tc::ffma2(dz[n], w2p[n], dh);
#pragma unroll
                        for (int n = 0; n < NP; ++n) gw2p[n] = tc::ffma2(dz[n], h, gw2p[n]);
                        const float2 dp = make_float2(pre.x > 0.f ? dh.x : 0.f, pre.y > 0.f ? dh.y : 0.f);  // relu'(0) = 0
                        gb1p = tc::fadd2(gb1p, dp);
                        float2 hi;  // dp truncated to tf32 + the exact remainder
                        hi.x = __uint_as_float(__float_as_uint(dp.x) & 0xffffe000u);
                        hi.y = __uint_as_float(__float_as_uint(dp.y) & 0xffffe000u);
                        const float2 l = tc::fsub2(dp, hi);
                        v[2 * q2] = __float_as_uint(hi.x), v[2 * q2 + 1] = __float_as_uint(hi.y);
                        lo[2 * q2] = __float_as_uint(l.x), lo[2 * q2 + 1] = __float_as_uint(l.y);
                    }
                };
                tc::tmem_ld16_nowait(c_hi, va);
                tc::tmem_wait_ld16(va);
                tc::tmem_ld16_nowait(c_hi + 16, vb);
                half(va, la, 0);
                tc::tmem_st16(c_hi, va);  // DP_hi replaces PRE in place
                tc::tmem_st16(c_lo, la);
                tc::tmem_wait_ld16(vb);
                half(vb, lb, 8);
                tc::tmem_st16(c_hi + 16, vb);
                tc::tmem_st16(c_lo + 16, lb);
                tc::tmem_wait_st();
                tc::tc_fence_before();
                tc::mbar_arrive(&bars->dp_full[d1]);
            }
            // ---- end of the pass: the two row halves of a hidden unit meet, dW1' leaves TMEM
            if (hh == 1) {
#pragma unroll
                for (int n = 0; n < NP; ++n) exch[n * kHB + jl] = gw2p[n].x + gw2p[n].y;
                exch[NP * kHB + jl] = gb1p.x + gb1p.y;
            }
            asm volatile("bar.sync 3, 256;" ::: "memory");
            tc::mbar_wait(&bars->done, hb & 1);  // every UMMA2 of this pass has retired
            tc::tc_fence_after();
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {  // this thread's 32 features, 16 at a time
                uint32_t g[16], g2[16];
                const int f0 = 32 * hh + 16 * hf;
                tc::tmem_ld16_nowait(lane_addr + kBColAcc + f0, g);        // dp_hi*x_hi
                tc::tmem_ld16_nowait(lane_addr + kBColAcc + 64 + f0, g2);  // dp_hi*x_lo + dp_lo*x_hi
                tc::tmem_wait_ld16(g);
                tc::tmem_wait_ld16(g2);
                float4* wrow = reinterpret_cast<float4*>(wsb + a.lay.oW1 + (size_t)j * O + f0);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if ((f0 >> 2) + c < ochunks)
                        wrow[c] = make_float4(__uint_as_float(g[4 * c]) + __uint_as_float(g2[4 * c]),
                                              __uint_as_float(g[4 * c + 1]) + __uint_as_float(g2[4 * c + 1]),
                                              __uint_as_float(g[4 * c + 2]) + __uint_as_float(g2[4 * c + 2]),
                                              __uint_as_float(g[4 * c + 3]) + __uint_as_float(g2[4 * c + 3]));
            }
            if (hh == 0) {
                wsb[a.lay.ob1 + j] = (gb1p.x + gb1p.y) + exch[NP * kHB + jl];
#pragma unroll
                for (int n = 0; n < NP; ++n)
                    if (n < a.N2) wsb[a.lay.oW2 + (size_t)n * H + j] = (gw2p[n].x + gw2p[n].y) + exch[n * kHB + jl];
            }
            if (hb == 0) {  // pads of the partial row
                const int64_t lo4[4] = {a.lay.oW1 + (int64_t)H * O, a.lay.ob1 + H, a.lay.oW2 + (int64_t)a.N2 * H,
                                        a.lay.ob2 + a.N2};
                const int64_t hi4[4] = {a.lay.ob1, a.lay.oW2, a.lay.ob2, a.lay.total};
                for (int sgm = 0; sgm < 4; ++sgm)
                    for (int64_t p = lo4[sgm] + tid; p < hi4[sgm]; p += kBEpiWarps * 32) wsb[p] = 0.f;
            }
            end_pass();
        }
    } else if (warp < kBIssuer) {
        // ============ producer: thread = (row of the tile, K atom); 4 warps ============
        const int pw = warp - kBEpiWarps, r = 32 * (pw & 1) + lane, atom = pw >> 1, ochunks = a.O >> 2;
        uint32_t xk[8];  // 16-byte chunk (lane >> 2) of a transposed row, swizzled by the row's phase k
#pragma unroll
        for (int k = 0; k < 8; ++k) xk[k] = (static_cast<uint32_t>(lane >> 2) ^ k) << 4;
        float4 v[8];
        float z[NP], zn[NP];
        auto load = [&](int i) {
            const int row = (cta + i * ncta) * kBRows + r;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < a.M && 8 * atom + c < ochunks) v[c] = ldg_v4(a.x + (size_t)row * a.O + 32 * atom + 4 * c);
            }
#pragma unroll
            for (int n = 0; n < NP; ++n) {
                zn[n] = 0.f;
                if (atom == 0 && row < a.M && n < a.N2) zn[n] = ldg_f(a.dout + (size_t)row * a.N2 + n);
            }
        };
        auto prefetch = [&](int i) {  // one thread: the tile's x rows (and dout rows) are contiguous
            const int64_t row0 = (int64_t)(cta + i * ncta) * kBRows;
            const int64_t rows = a.M - row0 < kBRows ? a.M - row0 : kBRows;
            if (i < n_my && rows > 0) {
                prefetch_l2(a.x + row0 * a.O, (uint32_t)(rows * a.O * 4));
                const float* z0 = a.dout + row0 * a.N2;
                const int64_t zb4 = rows * a.N2 * 4;
                if ((reinterpret_cast<uintptr_t>(z0) & 15) == 0 && (zb4 & 15) == 0) prefetch_l2(z0, (uint32_t)zb4);
            }
        };
        int it = 0;
        for (int hb = 0; hb < nblk; ++hb) {
            begin_pass(hb);
            float gb2[NP];
#pragma unroll
            for (int n = 0; n < NP; ++n) gb2[n] = 0.f;
            if (hb == 0) {  // later passes: requested before the previous pass's closing barrier (below)
                if (pw == 0 && lane == 0) prefetch(1), prefetch(2);
                load(0);
            }
            for (int i = 0; i < n_my; ++i, ++it) {
                const int s = it % kBStages, ph = (it / kBStages) & 1;
                if (pw == 0 && lane == 0) prefetch(i + 3);
                // row-major tile (B of UMMA1): free once UMMA1 of the tile two back has retired
                tc::mbar_wait(&bars->xa_empty[s], ph ^ 1);
                uint8_t* th = xa + s * kBXaBytes + atom * kBXAtomBytes;
                uint8_t* tl = th + 2 * kBXAtomBytes;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 hi, lo;
                    split4_fast(v[c], hi, lo);
                    const uint32_t off = tc::sw128_offset(r, c);
                    *reinterpret_cast<float4*>(th + off) = hi;
                    *reinterpret_cast<float4*>(tl + off) = lo;
                }
                tc::fence_proxy_async();
                tc::mbar_arrive(&bars->xa_full[s]);
#pragma unroll
                for (int n = 0; n < NP; ++n) z[n] = zn[n];
                // the registers are free again: the next tile's rows travel while this thread waits for
                // the transposed stage and fills it FROM THE ROW-MAJOR TILE it has just written (its own
                // row; the stage is not rewritten before this thread does so kBStages tiles later)
                if (i + 1 < n_my) load(i + 1);
                // transposed tile (B of UMMA2: row = feature, +64 for lo; K = this warp's 32 batch rows)
                // and dz: free once UMMA2 of the tile two back has retired
                tc::mbar_wait(&bars->xt_empty[s], ph ^ 1);
                // feature f = 32 atom + 4 c + e -> row f of the chunk: (f >> 3) * 1024 + k * 128 + xk[k] with
                // k = f & 7 = 4 (c & 1) + e known at compile time: eight swizzle terms, constant offsets
                uint8_t* tth = xtb + s * kBXtBytes + (pw & 1) * kBXtChunkBytes + (lane & 3) * 4 + atom * 4096;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const uint32_t off = tc::sw128_offset(r, c);
                    const float4 hi = *reinterpret_cast<const float4*>(th + off);
                    const float4 lo = *reinterpret_cast<const float4*>(tl + off);
                    const float hv[4] = {hi.x, hi.y, hi.z, hi.w}, lv[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int k = 4 * (c & 1) + e;
                        uint8_t* dst = tth + xk[k] + ((c >> 1) * 1024 + k * 128);
                        *reinterpret_cast<float*>(dst) = hv[e];
                        *reinterpret_cast<float*>(dst + 64 * 128) = lv[e];  // 64 is a multiple of the 8-row swizzle period
                    }
                }
                if (atom == 0) {
#pragma unroll
                    for (int n = 0; n < NP; ++n) {  // [row pair][n][2]: the epilogue reads pairs of rows
                        dzs[((s * (kBRows / 2) + (r >> 1)) * NP + n) * 2 + (r & 1)] = z[n];
                        gb2[n] += z[n];
                    }
                }
                tc::fence_proxy_async();
                tc::mbar_arrive(&bars->xt_full[s]);
            }
            if (hb + 1 < nblk) {  // the next pass starts with the same tiles: its first rows travel during the
                if (pw == 0 && lane == 0) prefetch(1), prefetch(2);  // read-out of this pass and the staging of the next block
                load(0);
            }
            if (hb == 0) {  // db2 = column sums of dout over this CTA's rows: fixed-order tree over 64 threads
#pragma unroll
                for (int n = 0; n < NP; ++n) {
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) gb2[n] += __shfl_xor_sync(0xffffffffu, gb2[n], off);
                }
                if (pw == 1 && lane == 0) {
#pragma unroll
                    for (int n = 0; n < NP; ++n) bars->gb2_part[n] = gb2[n];
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                if (pw == 0 && lane == 0) {
                    float* wsb = a.ws + (size_t)cta * a.lay.total;
#pragma unroll
                    for (int n = 0; n < NP; ++n)
                        if (n < a.N2) wsb[a.lay.ob2 + n] = gb2[n] + bars->gb2_part[n];
                }
            }
            end_pass();
        }
    } else {
        // =============================== UMMA issuer ===============================
        // UMMAs execute in issue order: UMMA1(i + 2) overwriting the PRE / DP_hi buffer that UMMA2(i)
        // reads needs no barrier, and when PRE(i + 2) has arrived UMMA2(i) is done with DP_lo too.
        const uint32_t idesc1 = tc::instr_desc_tf32_m128(kBRows);  // N = 64 batch rows
        const uint32_t idesc2w = tc::instr_desc_tf32_m128(128);    // N = 128: [hi | lo] features
        const uint32_t idesc2 = tc::instr_desc_tf32_m128(64);      // N = 64: hi features
        const int ksteps = (a.O + 7) >> 3;
        const uint64_t dxa = tc::smem_desc_k_sw128(xa, 0), dxt = tc::smem_desc_k_sw128(xtb, 0);
        int it = 0;  // tiles whose UMMA2 has been issued
        for (int hb = 0; hb < nblk; ++hb) {
            begin_pass(hb);
            const uint32_t tmem_base = bars->tmem_base;
            auto issue_umma1 = [&](int t) {  // t = running index of the tile
                const int s = t % kBStages, ph = (t / kBStages) & 1, d1 = t & 1;
                tc::mbar_wait(&bars->xa_full[s], ph);
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint64_t xh = dxa + static_cast<uint64_t>((s * kBXaBytes) >> 4);
                    const uint64_t xl = xh + static_cast<uint64_t>((2 * kBXAtomBytes) >> 4);
                    const uint32_t d0 = tmem_base + d1 * 64;
                    // A = W1' block from tensor memory (8 columns per K step), B = the x tile; correction terms
                    // first, hi*hi last (accumulation truncates, see the forward kernel)
                    const uint32_t w_hi = tmem_base + kBColW, w_lo = w_hi + 64;
#pragma unroll
                    for (int kk = 0; kk < 4 * kKA; ++kk) {
                        if (kk < ksteps) {
                            const uint64_t kx = static_cast<uint64_t>((kk >> 2) * (kBXAtomBytes >> 4) + (kk & 3) * 2);
                            tc::umma_tf32_ts(d0, w_lo + 8 * kk, xh + kx, idesc1, kk > 0);
                            tc::umma_tf32_ts(d0, w_hi + 8 * kk, xl + kx, idesc1, true);
                        }
                    }
#pragma unroll
                    for (int kk = 0; kk < 4 * kKA; ++kk) {
                        if (kk < ksteps) {
                            const uint64_t kx = static_cast<uint64_t>((kk >> 2) * (kBXAtomBytes >> 4) + (kk & 3) * 2);
                            tc::umma_tf32_ts(d0, w_hi + 8 * kk, xh + kx, idesc1, true);
                        }
                    }
                    tc::umma_commit(&bars->d1_full[d1]);
                    tc::umma_commit(&bars->xa_empty[s]);
                }
                __syncwarp();
            };
            if (n_my > 0) issue_umma1(it);
            if (n_my > 1) issue_umma1(it + 1);
            for (int i = 0; i < n_my; ++i, ++it) {
                const int s = it % kBStages, ph = (it / kBStages) & 1, d1 = it & 1;
                tc::mbar_wait(&bars->dp_full[d1], (it >> 1) & 1);  // DP hi / lo of this tile are in TMEM
                tc::mbar_wait(&bars->xt_full[s], ph);   // its transposed x tile is in shared memory
                tc::tc_fence_after();
                if (tc::elect_one()) {
                    const uint64_t xts = dxt + static_cast<uint64_t>((s * kBXtBytes) >> 4);
                    const uint32_t a_hi0 = tmem_base + d1 * 64, a_lo0 = tmem_base + kBColLo + d1 * 64;
                    const uint32_t acc0 = tmem_base + kBColAcc;
                    const uint32_t first = i > 0;
#pragma unroll
                    for (int kk = 0; kk < kBRows / 8; ++kk) {  // K = 64 batch rows, 8 per step
                        const uint64_t ko = static_cast<uint64_t>((kk >> 2) * (kBXtChunkBytes >> 4) + (kk & 3) * 2);
                        // columns [0,64) collect dp_hi*x_hi only; both correction terms (dp_hi*x_lo from the wide
                        // UMMA, dp_lo*x_hi from the second) share columns [64,128), summed at read-out
                        tc::umma_tf32_ts(acc0, a_hi0 + 8 * kk, xts + ko, idesc2w, first | (kk > 0));
                        tc::umma_tf32_ts(acc0 + 64, a_lo0 + 8 * kk, xts + ko, idesc2, true);
                    }
                    tc::umma_commit(&bars->xt_empty[s]);
                }
                __syncwarp();
                if (i + 2 < n_my) issue_umma1(it + 2);
            }
            if (tc::elect_one()) tc::umma_commit(&bars->done);
            __syncwarp();
            end_pass();
        }
    }
    if (warp == kBIssuer) tc::tmem_dealloc(bars->tmem_base, 512);
}

constexpr size_t bwd_smem_bytes(int np) {
    return 1024 + kBStages * (kBXaBytes + kBXtBytes) +
           (size_t)(kBStages * kBRows * np + (np + 1) * kHB) * sizeof(float) + sizeof(BBarriers);
}

template <typename K>
cudaError_t opt_in(K kernel, size_t smem, bool (&done)[64]) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !done[dev]) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) done[dev] = true;
    }
    return cudaSuccess;
}

}  // namespace

// Shapes the wide tensor-core kernels cover (the narrow kernels are preferred where they apply).
bool impala_mlp_tcw_eligible(const float* x, int M, int O, int H, int N2) {
    return M >= 1 && O >= 4 && O <= 64 && (O & 3) == 0 && H >= kHB && H % kHB == 0 && H <= 4096 && N2 >= 1 &&
           N2 <= 4 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && impala_env_int("IMPALA_MLP_TCW", 1) != 0;
}

int impala_mlp_fwd_tcw(const float* x, const float* params, float* out, int M, int O, int H, int N2,
                       cudaStream_t st) {
    FwdWArgs a{};
    a.x = x, a.params = params, a.out = out;
    a.M = M, a.O = O, a.H = H, a.N2 = N2;
    a.num_tiles = (M + kFTileM - 1) / kFTileM;
    a.lay = impala_make_layout(O, H, N2);
    int sms = 0;
    cudaError_t e;
    if ((e = impala_sm_count(&sms)) != cudaSuccess) return (int)e;
    const int grid = a.num_tiles < sms ? a.num_tiles : sms;
    static bool opted1[64] = {}, opted4[64] = {};
    if (N2 == 1) {
        if ((e = opt_in(mlp_fwd_tcw_kernel<1>, fwd_smem_bytes(1), opted1)) != cudaSuccess) return (int)e;
        mlp_fwd_tcw_kernel<1><<<grid, kFThreads, fwd_smem_bytes(1), st>>>(a);
    } else {
        if ((e = opt_in(mlp_fwd_tcw_kernel<4>, fwd_smem_bytes(4), opted4)) != cudaSuccess) return (int)e;
        mlp_fwd_tcw_kernel<4><<<grid, kFThreads, fwd_smem_bytes(4), st>>>(a);
    }
    return impala_launch_status();
}

// Per-CTA float32 partial gradient rows into ws (row stride = layout total); *nparts = rows written.
int impala_mlp_bwd_tcw(const float* x, const float* params, const float* dout, float* ws, int M, int O, int H,
                       int N2, cudaStream_t st, int* nparts) {
    BwdWArgs a{};
    a.x = x, a.params = params, a.dout = dout, a.ws = ws;
    a.M = M, a.O = O, a.H = H, a.N2 = N2;
    a.num_tiles = (M + kBRows - 1) / kBRows;
    a.lay = impala_make_layout(O, H, N2);
    int sms = 0;
    cudaError_t e;
    if ((e = impala_sm_count(&sms)) != cudaSuccess) return (int)e;
    int grid = a.num_tiles < sms ? a.num_tiles : sms;
    if (grid > kMaxParts) grid = kMaxParts;
    static bool opted1[64] = {}, opted4[64] = {};
    if (N2 == 1) {
        if ((e = opt_in(mlp_bwd_tcw_kernel<1>, bwd_smem_bytes(1), opted1)) != cudaSuccess) return (int)e;
        mlp_bwd_tcw_kernel<1><<<grid, kBThreads, bwd_smem_bytes(1), st>>>(a);
    } else {
        if ((e = opt_in(mlp_bwd_tcw_kernel<4>, bwd_smem_bytes(4), opted4)) != cudaSuccess) return (int)e;
        mlp_bwd_tcw_kernel<4><<<grid, kBThreads, bwd_smem_bytes(4), st>>>(a);
    }
    *nparts = grid;
    return impala_launch_status();
}
