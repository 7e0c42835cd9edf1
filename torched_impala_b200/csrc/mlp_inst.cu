// Instantiations of the MLP kernels for one padded observation width (IMPALA_OP) and one
// direction (IMPALA_BWD).  Built once per (width, direction) by torched_impala_b200/build.py.
#include "mlp_kernels.cuh"

#ifndef IMPALA_OP
#error "compile with -DIMPALA_OP=<8|24|32|64> -DIMPALA_BWD=<0|1>"
#endif

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#if IMPALA_BWD
#define KERNEL impala_mlp::mlp_bwd_kernel
#define ENTRY CAT(impala_mlp_bwd_op, IMPALA_OP)
#else
#define KERNEL impala_mlp::mlp_fwd_kernel
#define ENTRY CAT(impala_mlp_fwd_op, IMPALA_OP)
#endif

#define BY_NP(JPT, MAXT)                                                                      \
    switch (c.np) {                                                                           \
        case 1: return impala_mlp_launch(KERNEL<JPT, IMPALA_OP, 1, MAXT>, a, c, smem, st, grid);  \
        case 4: return impala_mlp_launch(KERNEL<JPT, IMPALA_OP, 4, MAXT>, a, c, smem, st, grid);  \
        default: return impala_mlp_launch(KERNEL<JPT, IMPALA_OP, 16, MAXT>, a, c, smem, st, grid); \
    }

#define BY_NP_KS2(MAXT)                                                                          \
    switch (c.np) {                                                                              \
        case 1: return impala_mlp_launch(KERNEL<1, IMPALA_OP, 1, MAXT, 2>, a, c, smem, st, grid);    \
        case 4: return impala_mlp_launch(KERNEL<1, IMPALA_OP, 4, MAXT, 2>, a, c, smem, st, grid);    \
        default: return impala_mlp_launch(KERNEL<1, IMPALA_OP, 16, MAXT, 2>, a, c, smem, st, grid);  \
    }

int ENTRY(const MlpArgs& a, const MlpConfig& c, size_t smem, cudaStream_t st, int* grid) {
#if IMPALA_BWD && IMPALA_OP == 64
    // wide observations: a lane pair per hidden unit (KS = 2), see mlp_kernels.cuh
    if (c.maxt == 128) { BY_NP_KS2(128) }
    BY_NP_KS2(256)
#else
    if (c.jpt == 1) { BY_NP(1, 128) }
    BY_NP(2, 256)
#endif
}
