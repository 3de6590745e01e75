// GEMM argument block and epilogues shared by the FFMA kernel (mq_gemm.cu) and the tcgen05 kernel (mq_gemm_tc.cu).
#pragma once
#include "mq_common.cuh"

#define GEMM_BK 16
#define EPI_SCALE 0
#define EPI_BIAS_ACT 1
#define EPI_ACT_GRAD 2

struct GemmArgs {
    const float *a; int64_t a_rs, a_cs;
    const float *b; int64_t b_rs, b_cs;
    const int32_t *idx;      // gather on the batch axis of A (or NULL)
    int idx_on_k;            // 0: index applies to A rows (m), 1: to A columns (k)
    float *c;                // [M,N] row-major (or split-K partials [S,M,N])
    int64_t m, n, k;
    int64_t k_per_split;
    float alpha;
    int epi, act; float leak;
    const float *bias;       // EPI_BIAS_ACT
    const float *y;          // EPI_ACT_GRAD, [M,N] row-major
    int64_t a_ones_row;      // > 0: rows m >= a_ones_row of A read as 1.0 (dW = x^T g with the bias gradient 1^T g as
                             // one more output row: W and b are adjacent in the flat parameter vector).
};

__device__ __forceinline__ float gemm_epilogue(const GemmArgs &g, float acc, int64_t m, int64_t n) {
    if (g.epi == EPI_BIAS_ACT) return mq_act_apply(acc + g.bias[n], g.act, g.leak);
    if (g.epi == EPI_ACT_GRAD) return mq_act_grad(g.y[m * g.n + n], acc, g.act, g.leak);
    return g.alpha * acc;
}

