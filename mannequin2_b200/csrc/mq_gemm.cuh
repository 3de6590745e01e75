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
    // Operand planes of A that already exist (tensor-core path only; every other path reads `a`): the image
    // [plane][a_planes_rows_pad][a_planes_pitch] of the SOURCE matrix of A in the mode of this call, read K-major
    // (a_planes_mn = 0: rows = m, columns = k) or MN-major (a_planes_mn = 1: rows = k, columns = m).  The layered MLP packs its
    // gathered minibatch once per step (with a column of ones behind the last feature) and feeds it to the forward product
    // K-major and to dW = x^T g MN-major.
    const void *a_planes; int64_t a_planes_rows_pad, a_planes_pitch; int a_planes_mn;
};

// Optional second outputs of the tcgen05 GEMM (mq_gemm_tma.cu): the epilogue's values written ALSO as operand planes,
// in the kernel's own operand format, so that the next product of a fused plan (csrc/mq_vae.cu) reads them with TMA
// instead of a pack launch over an fp32 copy.  The destination buffers are zero outside what is written here.
struct GmEmit {
    void *out_k; int64_t k_rows_pad, k_kpad;     // planes of C as a K-major operand:  [plane][k_rows_pad][k_kpad], element (m, n)
    void *out_t; int64_t t_rows_pad, t_kpad;     // planes of C^T (operand of dW = x^T g): [plane][t_rows_pad][t_kpad], element (n, m)
    int64_t ldc;                                 // row stride of g.c (0 = g.n); g.c may be NULL when only planes are wanted
};

// One operand-packing job of mq_gemm_tma_pack_multi: fp32 source -> planes, written at an offset of a larger plane image
// (two sources side by side make the [W_mean | W_logstd] operands of the VAE heads).
struct GmPackJob {
    const float *base; int64_t rs, cs;           // src(r, k) = base[row(r) * rs + col(k) * cs]
    const int32_t *idx; int idx_on_k;            // optional gather on the rows (0) or on k (1)
    int64_t rows, k;                             // extents of the source
    int64_t ones_row;                            // > 0: rows ones_row <= r < rows read as 1.0 (bias-gradient row of dW)
    int64_t ones_col;                            // > 0: columns ones_col <= k < k-extent read as 1.0, the source has ones_col columns
                                                 // (the same row of ones when x^T is fed as an MN-major operand: plain form only)
    int64_t rows_cover, k_cover;                 // extents WRITTEN (zero outside the source); k_cover % 8 == 0
    int64_t rows_pad, k_pad;                     // geometry of a destination plane
    int64_t row_off, k_off;                      // offsets inside the destination plane (k_off % 8 == 0)
    void *out;
    unsigned block0, blocks;                     // filled in by the launcher
};
#define GM_MAX_PACK_JOBS 20

int64_t mq_gemm_tma_pack_bytes(int64_t m, int64_t n, int64_t k, int planes);
int mq_gemm_tma_pack_multi(GmPackJob *jobs, int n_jobs, int mode, void *stream);
int mq_gemm_tma_packed(const void *pa, int64_t a_rows_pad, const void *pb, int64_t b_rows_pad, int64_t k_pad,
                       const GemmArgs &g, const GmEmit &em, int mode, int64_t splits, void *stream);
// the same with either operand MN-major (a_mn / b_mn != 0): planes [plane][k rows pad][m or n pad] -- the operand's M (N) axis
// is the unit-stride axis of its image; *_rows_pad is then the K-row count of a plane and *_mn_pad its row length in elements
int mq_gemm_tma_packed_mn(const void *pa, int64_t a_rows_pad, int a_mn, int64_t a_mn_pad, const void *pb, int64_t b_rows_pad,
                          int b_mn, int64_t b_mn_pad, int64_t k_pad, const GemmArgs &g, const GmEmit &em, int mode,
                          int64_t splits, void *stream);

__device__ __forceinline__ float gemm_epilogue(const GemmArgs &g, float acc, int64_t m, int64_t n) {
    if (g.epi == EPI_BIAS_ACT) return mq_act_apply(acc + g.bias[n], g.act, g.leak);
    if (g.epi == EPI_ACT_GRAD) return mq_act_grad(g.y[m * g.n + n], acc, g.act, g.leak);
    return g.alpha * acc;
}

#if !defined(MQ_EMU) && defined(__CUDACC__)
#include <cuda_bf16.h>
#include <cuda_fp16.h>
// ---- the operand planes of the tcgen05 GEMM, shared with the kernels that write planes themselves (csrc/mq_vae.cu) ----------
#define GM_LOW_SCALE 2048.0f
#define GM_LOW_UNSCALE 4.8828125e-4f

// P words of packed 16-bit pairs from two values: bf16 planes (F = 0) or fp16 with the scaled low plane (F = 1, P = 2)
template <int P, int F>
__device__ __forceinline__ void gm_split_pair(float a, float b, uint32_t *w) {
    if (F == 1) {
        const __half2 h0 = __floats2half2_rn(a, b);
        const float2 f0 = __half22float2(h0);
        const __half2 h1 = __floats2half2_rn((a - f0.x) * GM_LOW_SCALE, (b - f0.y) * GM_LOW_SCALE);
        w[0] = *reinterpret_cast<const uint32_t *>(&h0);
        w[1] = *reinterpret_cast<const uint32_t *>(&h1);
        return;
    }
#pragma unroll
    for (int pl = 0; pl < P; ++pl) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        w[pl] = *reinterpret_cast<const uint32_t *>(&h);
        a -= __uint_as_float(w[pl] << 16);
        b -= __uint_as_float(w[pl] & 0xFFFF0000u);
    }
}
#endif
