// Running mean/variance normaliser kernels.
//
// Replace the data-dependent parts of mannequin/_normalize.py:40-71
// (RunningNormalize.update / apply).  The scalar bookkeeping (total weights,
// sample counts, which branch is taken) depends only on batch sizes, so it
// stays on the host (mannequin2_b200/_normalize.py); the passes over the data
// run here with fp64 accumulators like the reference:
//   mq_col_mean : mean over the batch axis                     (4|8 B read / element)
//   mq_norm_cov : mean of (x-mean_old)*(x-mean_new)            (4|8 B read / element)
//   mq_norm_apply: (x-mean)/max(1e-8,sqrt(var)) [+tanh]        (4|8 B read + 4|8 B write)
// HBM-bound; two-stage fixed-order reductions (deterministic).
#include "mq_common.cuh"

#define NORM_MAX_BLOCKS 1024

static int64_t norm_blocks(int64_t rows) {
    int64_t nb = mq_ceil_div(rows, (int64_t)MQ_NT * 8);
    const int64_t cap = (int64_t)mq_sm_count() * 4;
    if (nb > cap) nb = cap;
    if (nb > NORM_MAX_BLOCKS) nb = NORM_MAX_BLOCKS;
    return nb < 1 ? 1 : nb;
}

extern "C" int64_t mq_norm_ws_bytes(int64_t rows, int64_t cols) {
    return norm_blocks(rows) * (cols < 1 ? 1 : cols) * (int64_t)sizeof(double);
}

// cols == 1: all threads cooperate on one column.
template <typename XT, bool COV>
__global__ void __launch_bounds__(MQ_NT)
norm_reduce1_kernel(const XT *__restrict__ x, int64_t rows, const double *__restrict__ mo,
                    const double *__restrict__ mn, double *__restrict__ partial) {
    __shared__ double red[MQ_NT / 32];
    double a = 0.0, b = 0.0;
    if (COV) { a = mo[0]; b = mn[0]; }
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * MQ_NT + threadIdx.x; i < rows; i += (int64_t)gridDim.x * MQ_NT) {
        const double v = (double)x[i];
        acc += COV ? (v - a) * (v - b) : v;
    }
    acc = mq_block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}

// cols > 1: one thread per column, blockIdx.y splits the rows.
template <typename XT, bool COV>
__global__ void __launch_bounds__(MQ_NT)
norm_reduceN_kernel(const XT *__restrict__ x, int64_t rows, int64_t cols,
                    const double *__restrict__ mo, const double *__restrict__ mn,
                    double *__restrict__ partial) {
    const int64_t c = (int64_t)blockIdx.x * MQ_NT + threadIdx.x;
    if (c >= cols) return;
    const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = (int64_t)blockIdx.y * per;
    const int64_t r1 = r0 + per < rows ? r0 + per : rows;
    double a = 0.0, b = 0.0;
    if (COV) { a = mo[c]; b = mn[c]; }
    double acc = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
        const double v = (double)x[r * cols + c];
        acc += COV ? (v - a) * (v - b) : v;
    }
    partial[(int64_t)blockIdx.y * cols + c] = acc;
}

__global__ void __launch_bounds__(MQ_NT)
norm_finish_kernel(const double *__restrict__ partial, int nb, int64_t cols, double scale,
                   double *__restrict__ out) {
    for (int64_t c = (int64_t)blockIdx.x * MQ_NT + threadIdx.x; c < cols; c += (int64_t)gridDim.x * MQ_NT) {
        double s = 0.0;
        for (int b = 0; b < nb; ++b) s += partial[(int64_t)b * cols + c];
        out[c] = s * scale;
    }
}

template <typename XT, bool COV>
static int norm_reduce(const void *x, int64_t rows, int64_t cols, const double *mo, const double *mn,
                       double scale, double *out, void *ws, int64_t ws_bytes, void *stream,
                       const char *what) {
    MQ_REQUIRE(x && out && ws, MQ_ERR_INVALID, "%s: null pointer", what);
    MQ_REQUIRE(rows >= 1 && cols >= 1, MQ_ERR_SHAPE, "%s: empty input", what);
    MQ_REQUIRE(ws_bytes >= mq_norm_ws_bytes(rows, cols), MQ_ERR_SHAPE, "%s: workspace too small", what);
    const int nb = (int)norm_blocks(rows);
    double *partial = (double *)ws;
    if (cols == 1) {
        auto kfn = norm_reduce1_kernel<XT, COV>;
        MQ_LAUNCH(kfn, (unsigned)nb, MQ_NT, 0, stream, (const XT *)x, rows, mo, mn, partial);
    } else {
        auto kfn = norm_reduceN_kernel<XT, COV>;
        MQ_LAUNCH(kfn, dim3((unsigned)mq_ceil_div(cols, MQ_NT), (unsigned)nb), MQ_NT, 0, stream,
                  (const XT *)x, rows, cols, mo, mn, partial);
    }
    MQ_LAUNCH(norm_finish_kernel, (unsigned)mq_ceil_div(cols, MQ_NT), MQ_NT, 0, stream, partial, nb,
              cols, scale / (double)rows, out);
    MQ_CHECK_LAUNCH(what);
    return MQ_OK;
}

extern "C" int mq_col_mean(const void *x, int x_dtype, int64_t rows, int64_t cols, double scale,
                           double *out, void *ws, int64_t ws_bytes, void *stream) {
    if (x_dtype == MQ_F64)
        return norm_reduce<double, false>(x, rows, cols, nullptr, nullptr, scale, out, ws, ws_bytes, stream, "mq_col_mean");
    return norm_reduce<float, false>(x, rows, cols, nullptr, nullptr, scale, out, ws, ws_bytes, stream, "mq_col_mean");
}

extern "C" int mq_norm_cov(const void *x, int x_dtype, int64_t rows, int64_t cols,
                           const double *mean_old, const double *mean_new, double scale,
                           double *out, void *ws, int64_t ws_bytes, void *stream) {
    MQ_REQUIRE(mean_old && mean_new, MQ_ERR_INVALID, "mq_norm_cov: null mean");
    if (x_dtype == MQ_F64)
        return norm_reduce<double, true>(x, rows, cols, mean_old, mean_new, scale, out, ws, ws_bytes, stream, "mq_norm_cov");
    return norm_reduce<float, true>(x, rows, cols, mean_old, mean_new, scale, out, ws, ws_bytes, stream, "mq_norm_cov");
}

template <typename XT, typename OT>
__global__ void __launch_bounds__(MQ_NT)
norm_apply_kernel(const XT *__restrict__ x, int64_t n, int64_t cols, const double *__restrict__ mean,
                  const double *__restrict__ var, OT *__restrict__ out, int fuse_tanh) {
    for (int64_t i = (int64_t)blockIdx.x * MQ_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * MQ_NT) {
        const int64_t c = cols == 1 ? 0 : i % cols;
        const double den = fmax(1e-8, sqrt(var[c]));
        const double z = ((double)x[i] - mean[c]) / den;
        if (fuse_tanh) out[i] = (OT)tanhf((float)z);      // Trajectory casts to f32 first (ppo.py:24-25)
        else out[i] = (OT)z;
    }
}

extern "C" int mq_norm_apply(const void *x, int x_dtype, int64_t rows, int64_t cols,
                             const double *mean, const double *var, void *out, int out_dtype,
                             int fuse_tanh, void *stream) {
    MQ_REQUIRE(x && mean && var && out, MQ_ERR_INVALID, "mq_norm_apply: null pointer");
    MQ_REQUIRE(rows >= 0 && cols >= 1, MQ_ERR_SHAPE, "mq_norm_apply: bad shape");
    const int64_t n = rows * cols;
    if (n == 0) return MQ_OK;
    int64_t nb = mq_ceil_div(n, MQ_NT);
    const int64_t cap = (int64_t)mq_sm_count() * 16;
    if (nb > cap) nb = cap;
#define MQ_APPLY_GO(XT, OT)                                                                   \
    do {                                                                                      \
        auto kfn = norm_apply_kernel<XT, OT>;                                                 \
        MQ_LAUNCH(kfn, (unsigned)nb, MQ_NT, 0, stream, (const XT *)x, n, cols, mean, var,     \
                  (OT *)out, fuse_tanh);                                                      \
    } while (0)
    if (x_dtype == MQ_F64 && out_dtype == MQ_F64) MQ_APPLY_GO(double, double);
    else if (x_dtype == MQ_F64) MQ_APPLY_GO(double, float);
    else if (out_dtype == MQ_F64) MQ_APPLY_GO(float, double);
    else MQ_APPLY_GO(float, float);
#undef MQ_APPLY_GO
    MQ_CHECK_LAUNCH("mq_norm_apply");
    return MQ_OK;
}
