// Debug only (not part of the public ABI): checks the thread <-> (lane, column) mapping of the
// fragment-shaped TMEM accesses that tc_common.cuh documents.  One CTA of 128 threads fills a
// 128-lane x 32-column tile through the plain 32x32b store (lane L, column c <- 100 L + c), reads
// it back through 16x256b.x4 / .x2 loads, and round-trips a second tile through the .x2 store.
#include "tc_common.cuh"

namespace {
__global__ void __launch_bounds__(128, 1) tmem_fragment_kernel(float* out) {
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc(&s_base, 64);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t base = s_base, q_addr = base + (static_cast<uint32_t>(32 * warp) << 16);
    float v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = 100.f * tid + c;
    tc::tmem_st32(q_addr, v);
    tc::tmem_wait_st();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    float* o = out + tid * 96;
    // [0, 32): x4 loads of lane halves 0 and 1 (16 registers each)
    float f[16];
    tc::tmem_ld_16x256b_x4(q_addr, f);
    tc::tmem_wait_ld();
    for (int i = 0; i < 16; ++i) o[i] = f[i];
    tc::tmem_ld_16x256b_x4(q_addr + (16u << 16), f);
    tc::tmem_wait_ld();
    for (int i = 0; i < 16; ++i) o[16 + i] = f[i];
    // [32, 48): x2 loads at column 16 of both lane halves
    float g[8];
    tc::tmem_ld_16x256b_x2(q_addr + 16, g);
    tc::tmem_wait_ld();
    for (int i = 0; i < 8; ++i) o[32 + i] = g[i];
    tc::tmem_ld_16x256b_x2(q_addr + (16u << 16) + 16, g);
    tc::tmem_wait_ld();
    for (int i = 0; i < 8; ++i) o[40 + i] = g[i];
    // [48, 80): store fragments (value = 1000 * tid + register index) at columns 32..47 with the
    // .x2 store, read the tile back with the plain per-lane load
    for (int i = 0; i < 8; ++i) g[i] = 1000.f * tid + i;
    tc::tmem_st_16x256b_x2(q_addr + 32, g);
    for (int i = 0; i < 8; ++i) g[i] = 1000.f * tid + 8 + i;
    tc::tmem_st_16x256b_x2(q_addr + (16u << 16) + 32, g);
    tc::tmem_wait_st();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    tc::tmem_ld32(q_addr + 32, v);
    for (int c = 0; c < 32; ++c) o[48 + c] = v[c];
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc::tc_fence_after();
        tc::tmem_dealloc(base, 64);
    }
}
}  // namespace

extern "C" int impala_debug_tmem_fragment(float* out_128x96) {
    tmem_fragment_kernel<<<1, 128>>>(out_128x96);
    cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? 0 : (int)e;
}
