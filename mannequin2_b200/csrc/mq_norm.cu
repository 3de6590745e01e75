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

// one CTA per output: thread t adds row blocks t, t + 256, ... then the CTA adds its 256 sums, all in a fixed order
// (one thread walking ~1,000 partials with dependent loads took 34 us: as long as the pass over the data itself)
__global__ void __launch_bounds__(MQ_NT)
norm_finish_kernel(const double *__restrict__ partial, int nb, int64_t cols, double scale,
                   double *__restrict__ out) {
    __shared__ double red[MQ_NT / 32];
    const int64_t c = blockIdx.x;
    double s = 0.0;
    for (int b = threadIdx.x; b < nb; b += MQ_NT) s += partial[(int64_t)b * cols + c];
    s = mq_block_sum(s, red);
    if (threadIdx.x == 0) out[c] = s * scale;
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
    MQ_LAUNCH(norm_finish_kernel, (unsigned)cols, MQ_NT, 0, stream, partial, nb,
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

// ---- one statistics pass -----------------------------------------------------------------------------------
// RunningNormalize.update (_normalize.py:40-57) needs mean(x) and mean((x - mean_old)(x - mean_new)), where mean_new
// depends on mean(x): two passes as written.  With the shift m = mean_old (known before the pass)
//     s1 = mean(x - m),  s2 = mean((x - m)^2)
//     mean(x) = m + s1,  mean((x - mean_old)(x - mean_new)) = s2 - (mean_new - m) s1
// both come out of ONE pass, in fp64, without the cancellation of raw sum(x^2) (m tracks the data).
// cols == 1: 128-bit loads, four independent fp64 accumulator pairs per thread, fixed-order two-stage reduction.
template <typename XT>
__global__ void __launch_bounds__(MQ_NT)
norm_stats1_kernel(const XT *__restrict__ x, int64_t rows, const double *__restrict__ shift, double *__restrict__ partial) {
    __shared__ double red[MQ_NT / 32];
    const double m = shift[0];
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    const int64_t tid = (int64_t)blockIdx.x * MQ_NT + threadIdx.x, nth = (int64_t)gridDim.x * MQ_NT;
    if (sizeof(XT) == 4 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        const int64_t nq = rows >> 2;
        int64_t q = tid;
        for (; q + nth < nq; q += 2 * nth) {                 // two independent 16-byte loads in flight
            const float4 u = x4[q], w = x4[q + nth];
            const double u0 = (double)u.x - m, u1 = (double)u.y - m, u2 = (double)u.z - m, u3 = (double)u.w - m;
            const double w0 = (double)w.x - m, w1 = (double)w.y - m, w2 = (double)w.z - m, w3 = (double)w.w - m;
            a0 += (u0 + u1) + (u2 + u3);
            a1 += (w0 + w1) + (w2 + w3);
            b0 = fma(u0, u0, fma(u1, u1, fma(u2, u2, fma(u3, u3, b0))));
            b1 = fma(w0, w0, fma(w1, w1, fma(w2, w2, fma(w3, w3, b1))));
        }
        for (; q < nq; q += nth) {
            const float4 u = x4[q];
            const double u0 = (double)u.x - m, u1 = (double)u.y - m, u2 = (double)u.z - m, u3 = (double)u.w - m;
            a0 += (u0 + u1) + (u2 + u3);
            b0 = fma(u0, u0, fma(u1, u1, fma(u2, u2, fma(u3, u3, b0))));
        }
        for (int64_t i = (nq << 2) + tid; i < rows; i += nth) {
            const double v = (double)x[i] - m;
            a0 += v; b0 = fma(v, v, b0);
        }
    } else {
        for (int64_t i = tid; i < rows; i += nth) {
            const double v = (double)x[i] - m;
            a0 += v; b0 = fma(v, v, b0);
        }
    }
    const double sa = mq_block_sum(a0 + a1, red);
    const double sb = mq_block_sum(b0 + b1, red);
    if (threadIdx.x == 0) { partial[blockIdx.x] = sa; partial[gridDim.x + blockIdx.x] = sb; }
}

// cols > 1: one thread per column, blockIdx.y splits the rows; partial[stat][row block][col]
template <typename XT>
__global__ void __launch_bounds__(MQ_NT)
norm_statsN_kernel(const XT *__restrict__ x, int64_t rows, int64_t cols, const double *__restrict__ shift,
                   double *__restrict__ partial) {
    const int64_t c = (int64_t)blockIdx.x * MQ_NT + threadIdx.x;
    if (c >= cols) return;
    const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
    const int64_t r0 = (int64_t)blockIdx.y * per;
    const int64_t r1 = r0 + per < rows ? r0 + per : rows;
    const double m = shift[c];
    double a = 0.0, b = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
        const double v = (double)x[r * cols + c] - m;
        a += v; b = fma(v, v, b);
    }
    partial[(int64_t)blockIdx.y * cols + c] = a;
    partial[((int64_t)gridDim.y + blockIdx.y) * cols + c] = b;
}

// out[stat * cols + c] = scale * sum over row blocks (fixed order)
__global__ void __launch_bounds__(MQ_NT)
norm_stats_finish_kernel(const double *__restrict__ partial, int nb, int64_t cols, double scale, double *__restrict__ out) {
    __shared__ double red[MQ_NT / 32];
    const int64_t e = blockIdx.x;                          // one CTA per (statistic, column)
    const int64_t stat = e / cols, c = e - stat * cols;
    double s = 0.0;
    for (int b = threadIdx.x; b < nb; b += MQ_NT) s += partial[(stat * nb + b) * cols + c];
    s = mq_block_sum(s, red);
    if (threadIdx.x == 0) out[e] = s * scale;
}

static int64_t norm_stats_blocks(int64_t rows, int64_t cols) {
    if (cols > 1) return norm_blocks(rows);
    int64_t nb = mq_ceil_div(rows, (int64_t)MQ_NT * 32);
    const int64_t cap = (int64_t)mq_sm_count() * 8;
    if (nb > cap) nb = cap;
    return nb < 1 ? 1 : nb;
}

extern "C" int64_t mq_norm_stats_ws_bytes(int64_t rows, int64_t cols) {
    return 2 * norm_stats_blocks(rows, cols) * (cols < 1 ? 1 : cols) * (int64_t)sizeof(double);
}

extern "C" int mq_norm_stats(const void *x, int x_dtype, int64_t rows, int64_t cols, const double *shift, double scale,
                             double *out, void *ws, int64_t ws_bytes, void *stream) {
    MQ_REQUIRE(x && shift && out && ws, MQ_ERR_INVALID, "mq_norm_stats: null pointer");
    MQ_REQUIRE(rows >= 1 && cols >= 1, MQ_ERR_SHAPE, "mq_norm_stats: empty input");
    MQ_REQUIRE(ws_bytes >= mq_norm_stats_ws_bytes(rows, cols), MQ_ERR_SHAPE, "mq_norm_stats: workspace too small");
    const int nb = (int)norm_stats_blocks(rows, cols);
    double *partial = (double *)ws;
    if (cols == 1) {
        if (x_dtype == MQ_F64) { auto k = norm_stats1_kernel<double>; MQ_LAUNCH(k, (unsigned)nb, MQ_NT, 0, stream, (const double *)x, rows, shift, partial); }
        else { auto k = norm_stats1_kernel<float>; MQ_LAUNCH(k, (unsigned)nb, MQ_NT, 0, stream, (const float *)x, rows, shift, partial); }
    } else {
        const dim3 grid((unsigned)mq_ceil_div(cols, MQ_NT), (unsigned)nb);
        if (x_dtype == MQ_F64) { auto k = norm_statsN_kernel<double>; MQ_LAUNCH(k, grid, MQ_NT, 0, stream, (const double *)x, rows, cols, shift, partial); }
        else { auto k = norm_statsN_kernel<float>; MQ_LAUNCH(k, grid, MQ_NT, 0, stream, (const float *)x, rows, cols, shift, partial); }
    }
    MQ_LAUNCH(norm_stats_finish_kernel, (unsigned)(2 * cols), MQ_NT, 0, stream, partial, nb, cols,
              scale / (double)rows, out);
    MQ_CHECK_LAUNCH("mq_norm_stats");
    return MQ_OK;
}

// statistics -> running state (_normalize.py:9-25,44-57): stats = [s1[cols], s2[cols]] taken with shift = mean (the value
// BEFORE this update).  mean += coef_mean * s1;  var += coef_var * (var_scale * (s2 - (mean_new - mean_old) s1) - var)
// when coef_var > 0.
__global__ void __launch_bounds__(MQ_NT)
norm_update_kernel(double *__restrict__ mean, double *__restrict__ var, const double *__restrict__ stats, int64_t cols,
                   double coef_mean, double coef_var, double var_scale) {
    for (int64_t c = (int64_t)blockIdx.x * MQ_NT + threadIdx.x; c < cols; c += (int64_t)gridDim.x * MQ_NT) {
        const double s1 = stats[c], s2 = stats[cols + c];
        const double mo = mean[c];
        const double mn = mo + coef_mean * s1;          // RunningMean.update with the batch mean mo + s1
        mean[c] = mn;
        if (coef_var > 0.0) {
            const double cov = s2 - (mn - mo) * s1;
            const double v = var[c];
            var[c] = v + coef_var * (cov * var_scale - v);
        }
    }
}

extern "C" int mq_norm_update_from_stats(double *mean, double *var, const double *stats, int64_t cols, double coef_mean,
                                         double coef_var, double var_scale, void *stream) {
    MQ_REQUIRE(mean && var && stats, MQ_ERR_INVALID, "mq_norm_update_from_stats: null pointer");
    MQ_REQUIRE(cols >= 1, MQ_ERR_SHAPE, "mq_norm_update_from_stats: bad shape");
    MQ_LAUNCH(norm_update_kernel, (unsigned)mq_ceil_div(cols, MQ_NT), MQ_NT, 0, stream, mean, var, stats, cols, coef_mean,
              coef_var, var_scale);
    MQ_CHECK_LAUNCH("mq_norm_update_from_stats");
    return MQ_OK;
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

// scalar statistic (the advantage normaliser of ppo.py:24-25), float32 in and out: 128-bit loads / stores, the divisor
// hoisted out of the loop (as a reciprocal: <= 1 ulp of fp64 before the cast to float32)
__global__ void __launch_bounds__(MQ_NT)
norm_apply1_f32_kernel(const float *__restrict__ x, int64_t n, const double *__restrict__ mean,
                       const double *__restrict__ var, float *__restrict__ out, int fuse_tanh) {
    const double m = mean[0];
    const double inv = 1.0 / fmax(1e-8, sqrt(var[0]));
    const int64_t tid = (int64_t)blockIdx.x * MQ_NT + threadIdx.x, nth = (int64_t)gridDim.x * MQ_NT;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    float4 *o4 = reinterpret_cast<float4 *>(out);
    const int64_t nq = n >> 2;
    auto apply4 = [&](const float4 u) {
        float4 r;
        r.x = (float)(((double)u.x - m) * inv); r.y = (float)(((double)u.y - m) * inv);
        r.z = (float)(((double)u.z - m) * inv); r.w = (float)(((double)u.w - m) * inv);
        if (fuse_tanh) { r.x = tanhf(r.x); r.y = tanhf(r.y); r.z = tanhf(r.z); r.w = tanhf(r.w); }
        return r;
    };
    int64_t q = tid;
    for (; q + 3 * nth < nq; q += 4 * nth) {           // four independent loads in flight per thread (one input stream)
        const float4 u0 = x4[q], u1 = x4[q + nth], u2 = x4[q + 2 * nth], u3 = x4[q + 3 * nth];
        o4[q] = apply4(u0);
        o4[q + nth] = apply4(u1);
        o4[q + 2 * nth] = apply4(u2);
        o4[q + 3 * nth] = apply4(u3);
    }
    for (; q < nq; q += nth) o4[q] = apply4(x4[q]);
    for (int64_t i = (nq << 2) + tid; i < n; i += nth) {
        const float z = (float)(((double)x[i] - m) * inv);
        out[i] = fuse_tanh ? tanhf(z) : z;
    }
}

extern "C" int mq_norm_apply(const void *x, int x_dtype, int64_t rows, int64_t cols,
                             const double *mean, const double *var, void *out, int out_dtype,
                             int fuse_tanh, void *stream) {
    MQ_REQUIRE(x && mean && var && out, MQ_ERR_INVALID, "mq_norm_apply: null pointer");
    MQ_REQUIRE(rows >= 0 && cols >= 1, MQ_ERR_SHAPE, "mq_norm_apply: bad shape");
    const int64_t n = rows * cols;
    if (n == 0) return MQ_OK;
    if (cols == 1 && x_dtype == MQ_F32 && out_dtype == MQ_F32 &&
        (((uintptr_t)x | (uintptr_t)out) & 15) == 0) {
        int64_t nb1 = mq_ceil_div(n, (int64_t)MQ_NT * 4);
        const int64_t cap1 = (int64_t)mq_sm_count() * 16;
        if (nb1 > cap1) nb1 = cap1;
        MQ_LAUNCH(norm_apply1_f32_kernel, (unsigned)nb1, MQ_NT, 0, stream, (const float *)x, n, mean, var, (float *)out,
                  fuse_tanh);
        MQ_CHECK_LAUNCH("mq_norm_apply");
        return MQ_OK;
    }
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
