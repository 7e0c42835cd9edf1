// Shared helpers for the sm_100a learner kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/impala_b200.h"

#define IMPALA_FULL_MASK 0xffffffffu

static inline int64_t impala_round_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct MlpLayout {
    int64_t oW1, ob1, oW2, ob2, total;
};

static inline MlpLayout impala_make_layout(int O, int H, int N2) {
    MlpLayout l;
    const int64_t al = IMPALA_PARAM_ALIGN;
    l.oW1 = 0;
    l.ob1 = impala_round_up(l.oW1 + (int64_t)H * O, al);
    l.oW2 = impala_round_up(l.ob1 + H, al);
    l.ob2 = impala_round_up(l.oW2 + (int64_t)N2 * H, al);
    l.total = impala_round_up(l.ob2 + N2, al);
    return l;
}

static inline int impala_launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? IMPALA_OK : (int)e;
}

__device__ __forceinline__ double warp_sum_f64(double x) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(IMPALA_FULL_MASK, x, off);
    return x;
}
