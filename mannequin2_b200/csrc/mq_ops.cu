// Per-op kernels: trajectory gather, element-wise primitives, distribution heads.
//
// These back (a) the general layer graph, where every node of
// mannequin/basicnet.py / mannequin/backprop.py is one launch, and (b) the
// layered MLP path for nets too wide for the fused shared-memory kernels.
// All are HBM-bound streaming kernels: coalesced, grid-stride, 4 B/element/pass.
#include "mq_common.cuh"

static inline unsigned ew_blocks(int64_t n) {
    int64_t nb = mq_ceil_div(n, MQ_NT);
    const int64_t cap = (int64_t)mq_sm_count() * 16;
    if (nb > cap) nb = cap;
    return (unsigned)(nb < 1 ? 1 : nb);
}

#define EW_LOOP(i, n) \
    for (int64_t i = (int64_t)blockIdx.x * MQ_NT + threadIdx.x; i < (n); i += (int64_t)gridDim.x * MQ_NT)

// ---- mannequin/_trajectory.py:79-86: t[idx] with an integer array ----------------
// One warp moves whole rows with 16-byte (or 4-byte) words; negative indices wrap
// like NumPy.  Out-of-range indices are rejected on the host before the launch.
template <typename W>
__global__ void __launch_bounds__(MQ_NT)
gather_rows_kernel(const W *__restrict__ src, const int32_t *__restrict__ idx, int64_t n_idx,
                   int64_t row_words, int64_t n_src_rows, W *__restrict__ dst) {
    const int64_t total = n_idx * row_words;
    EW_LOOP(e, total) {
        const int64_t r = e / row_words, w = e - r * row_words;
        int64_t s = idx[r];
        if (s < 0) s += n_src_rows;
        dst[e] = src[s * row_words + w];
    }
}

extern "C" int mq_gather_rows(const void *src, const int32_t *idx, int64_t n_idx, int64_t row_bytes,
                              int64_t n_src_rows, void *dst, void *stream) {
    MQ_REQUIRE(n_idx >= 0 && row_bytes > 0 && n_src_rows > 0, MQ_ERR_SHAPE, "mq_gather_rows: bad sizes");
    if (n_idx == 0) return MQ_OK;
    MQ_REQUIRE(src && idx && dst, MQ_ERR_INVALID, "mq_gather_rows: null pointer");
    const bool a16 = (row_bytes % 16 == 0) && ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
    if (a16) {
        auto kfn = gather_rows_kernel<float4>;
        MQ_LAUNCH(kfn, ew_blocks(n_idx * (row_bytes / 16)), MQ_NT, 0, stream, (const float4 *)src, idx,
                  n_idx, row_bytes / 16, n_src_rows, (float4 *)dst);
    } else if (row_bytes % 4 == 0) {
        auto kfn = gather_rows_kernel<float>;
        MQ_LAUNCH(kfn, ew_blocks(n_idx * (row_bytes / 4)), MQ_NT, 0, stream, (const float *)src, idx,
                  n_idx, row_bytes / 4, n_src_rows, (float *)dst);
    } else {
        auto kfn = gather_rows_kernel<unsigned char>;
        MQ_LAUNCH(kfn, ew_blocks(n_idx * row_bytes), MQ_NT, 0, stream, (const unsigned char *)src, idx,
                  n_idx, row_bytes, n_src_rows, (unsigned char *)dst);
    }
    MQ_CHECK_LAUNCH("mq_gather_rows");
    return MQ_OK;
}

// ---- mannequin/backprop.py:11-43: add / multiply with trailing-shape broadcast ----
__global__ void __launch_bounds__(MQ_NT)
binary_kernel(const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out,
              int64_t n, int64_t period, int op) {
    EW_LOOP(i, n) {
        const float bv = b[period == n ? i : i % period];
        out[i] = op == 0 ? a[i] + bv : a[i] * bv;
    }
}

extern "C" int mq_binary(const float *a, const float *b, float *out, int64_t n, int64_t b_period,
                         int op, void *stream) {
    MQ_REQUIRE(n >= 0 && b_period >= 1 && (op == 0 || op == 1), MQ_ERR_INVALID, "mq_binary: bad arguments");
    if (n == 0) return MQ_OK;
    MQ_REQUIRE(a && b && out, MQ_ERR_INVALID, "mq_binary: null pointer");
    MQ_REQUIRE(n % b_period == 0, MQ_ERR_SHAPE, "mq_binary: shapes do not broadcast");
    MQ_LAUNCH(binary_kernel, ew_blocks(n), MQ_NT, 0, stream, a, b, out, n, b_period, op);
    MQ_CHECK_LAUNCH("mq_binary");
    return MQ_OK;
}

extern "C" int mq_add_bias(const float *x, const float *b, float *y, int64_t rows, int64_t cols,
                           void *stream) {
    return mq_binary(x, b, y, rows * cols, cols, 0, stream);
}

// out[c] = scale * sum_r g[r,c] * a[r,c]; block (32 cols x 8 row-lanes), fixed order.
__global__ void __launch_bounds__(MQ_NT)
col_sum_kernel(const float *__restrict__ g, const float *__restrict__ a, int64_t rows, int64_t cols,
               float scale, float *__restrict__ out) {
    __shared__ float part[8][33];
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int64_t c = (int64_t)blockIdx.x * 32 + cx;
    float acc = 0.0f;
    if (c < cols)
        for (int64_t r = ry; r < rows; r += 8) {
            const float gv = g[r * cols + c];
            acc += a ? gv * a[r * cols + c] : gv;
        }
    part[ry][cx] = acc;
    __syncthreads();
    if (ry == 0 && c < cols) {
        float s = part[0][cx];
#pragma unroll
        for (int k = 1; k < 8; ++k) s += part[k][cx];
        out[c] = s * scale;
    }
}

// The same for tall matrices (thousands of rows): block = 8 columns x 32 row-lanes, so that a 4096 x 784 bias gradient is
// 98 CTAs of 128 independent 32-byte loads per thread instead of 25 CTAs of 512; fixed order (four interleaved partial
// sums per thread, then the 32 row-lanes in order).
__global__ void __launch_bounds__(MQ_NT)
col_sum_tall_kernel(const float *__restrict__ g, const float *__restrict__ a, int64_t rows, int64_t cols,
                    float scale, float *__restrict__ out) {
    __shared__ float part[32][9];
    const int cx = threadIdx.x & 7, ry = threadIdx.x >> 3;
    const int64_t c = (int64_t)blockIdx.x * 8 + cx;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    if (c < cols) {
        int64_t r = ry;
        for (; r + 96 < rows; r += 128) {
            const float g0 = g[r * cols + c], g1 = g[(r + 32) * cols + c], g2 = g[(r + 64) * cols + c],
                        g3 = g[(r + 96) * cols + c];
            if (a) {
                s0 += g0 * a[r * cols + c]; s1 += g1 * a[(r + 32) * cols + c];
                s2 += g2 * a[(r + 64) * cols + c]; s3 += g3 * a[(r + 96) * cols + c];
            } else {
                s0 += g0; s1 += g1; s2 += g2; s3 += g3;
            }
        }
        for (; r < rows; r += 32) s0 += a ? g[r * cols + c] * a[r * cols + c] : g[r * cols + c];
    }
    part[ry][cx] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (ry == 0 && c < cols) {
        float s = part[0][cx];
#pragma unroll
        for (int k = 1; k < 32; ++k) s += part[k][cx];
        out[c] = s * scale;
    }
}

extern "C" int mq_col_sum(const float *g, const float *a, int64_t rows, int64_t cols, float scale,
                          float *out, void *stream) {
    MQ_REQUIRE(rows >= 0 && cols >= 1, MQ_ERR_SHAPE, "mq_col_sum: bad shape");
    MQ_REQUIRE(g && out, MQ_ERR_INVALID, "mq_col_sum: null pointer");
    if (rows >= 1024) {
        MQ_LAUNCH(col_sum_tall_kernel, (unsigned)mq_ceil_div(cols, 8), MQ_NT, 0, stream, g, a, rows, cols, scale, out);
        MQ_CHECK_LAUNCH("mq_col_sum");
        return MQ_OK;
    }
    MQ_LAUNCH(col_sum_kernel, (unsigned)mq_ceil_div(cols, 32), MQ_NT, 0, stream, g, a, rows, cols, scale, out);
    MQ_CHECK_LAUNCH("mq_col_sum");
    return MQ_OK;
}

// ---- mannequin/backprop.py:53-60: relu / tanh ---------------------------------------
__global__ void __launch_bounds__(MQ_NT)
act_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t n, int act, float leak) {
    EW_LOOP(i, n) y[i] = mq_act_apply(x[i], act, leak);
}
__global__ void __launch_bounds__(MQ_NT)
act_bwd_kernel(const float *__restrict__ y, const float *__restrict__ g, float *__restrict__ out,
               int64_t n, int act, float leak) {
    EW_LOOP(i, n) out[i] = mq_act_grad(y[i], g[i], act, leak);
}

// 128-bit versions (16-byte aligned arrays, n % 4 == 0): the elementwise kernels are pure HBM streams
// (the forward pass has ONE input stream: four independent 128-bit loads per thread are issued before the first use, or a
//  thread keeps 16 bytes in flight and 2048 threads per SM cannot cover the HBM latency: 81 % -> 9x % of the copy peak)
__device__ __forceinline__ float4 act_apply4(const float4 v, int act, float leak) {
    return make_float4(mq_act_apply(v.x, act, leak), mq_act_apply(v.y, act, leak), mq_act_apply(v.z, act, leak),
                       mq_act_apply(v.w, act, leak));
}
__global__ void __launch_bounds__(MQ_NT)
act_fwd4_kernel(const float4 *__restrict__ x, float4 *__restrict__ y, int64_t n4, int act, float leak) {
    const int64_t nth = (int64_t)gridDim.x * MQ_NT;
    int64_t i = (int64_t)blockIdx.x * MQ_NT + threadIdx.x;
    for (; i + 3 * nth < n4; i += 4 * nth) {
        const float4 v0 = x[i], v1 = x[i + nth], v2 = x[i + 2 * nth], v3 = x[i + 3 * nth];
        y[i] = act_apply4(v0, act, leak);
        y[i + nth] = act_apply4(v1, act, leak);
        y[i + 2 * nth] = act_apply4(v2, act, leak);
        y[i + 3 * nth] = act_apply4(v3, act, leak);
    }
    for (; i < n4; i += nth) y[i] = act_apply4(x[i], act, leak);
}
__global__ void __launch_bounds__(MQ_NT)
act_bwd4_kernel(const float4 *__restrict__ y, const float4 *__restrict__ g, float4 *__restrict__ out, int64_t n4, int act,
                float leak) {
    EW_LOOP(i, n4) {
        const float4 v = y[i], w = g[i];
        out[i] = make_float4(mq_act_grad(v.x, w.x, act, leak), mq_act_grad(v.y, w.y, act, leak),
                             mq_act_grad(v.z, w.z, act, leak), mq_act_grad(v.w, w.w, act, leak));
    }
}
static inline bool ew_vec4(int64_t n, const void *a, const void *b, const void *c) {
    return (n & 3) == 0 && (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0;
}

extern "C" int mq_act_forward(const float *x, float *y, int64_t n, int act, float leak, void *stream) {
    MQ_REQUIRE(act >= 0 && act <= 2, MQ_ERR_INVALID, "mq_act_forward: unknown activation %d", act);
    if (n <= 0) return MQ_OK;
    if (ew_vec4(n, x, y, nullptr)) {
        MQ_LAUNCH(act_fwd4_kernel, ew_blocks(n / 4), MQ_NT, 0, stream, (const float4 *)x, (float4 *)y, n / 4, act, leak);
        MQ_CHECK_LAUNCH("mq_act_forward");
        return MQ_OK;
    }
    MQ_LAUNCH(act_fwd_kernel, ew_blocks(n), MQ_NT, 0, stream, x, y, n, act, leak);
    MQ_CHECK_LAUNCH("mq_act_forward");
    return MQ_OK;
}
extern "C" int mq_act_backward(const float *y, const float *g, float *out, int64_t n, int act,
                               float leak, void *stream) {
    MQ_REQUIRE(act >= 0 && act <= 2, MQ_ERR_INVALID, "mq_act_backward: unknown activation %d", act);
    if (n <= 0) return MQ_OK;
    if (ew_vec4(n, y, g, out)) {
        MQ_LAUNCH(act_bwd4_kernel, ew_blocks(n / 4), MQ_NT, 0, stream, (const float4 *)y, (const float4 *)g, (float4 *)out,
                  n / 4, act, leak);
        MQ_CHECK_LAUNCH("mq_act_backward");
        return MQ_OK;
    }
    MQ_LAUNCH(act_bwd_kernel, ew_blocks(n), MQ_NT, 0, stream, y, g, out, n, act, leak);
    MQ_CHECK_LAUNCH("mq_act_backward");
    return MQ_OK;
}

// ---- mnist/vae.py:24-34: Clip ------------------------------------------------------
__global__ void __launch_bounds__(MQ_NT)
clip_fwd_kernel(const float *__restrict__ x, float *__restrict__ y, int64_t n, float lo, float hi) {
    EW_LOOP(i, n) y[i] = fminf(fmaxf(x[i], lo), hi);
}
__global__ void __launch_bounds__(MQ_NT)
clip_bwd_kernel(const float *__restrict__ x, const float *__restrict__ g, float *__restrict__ out,
                int64_t n, float lo, float hi) {
    EW_LOOP(i, n) {
        const float v = x[i], gv = g[i];
        const bool kill = (v < lo && gv < 0.0f) || (v > hi && gv > 0.0f);
        out[i] = kill ? 0.0f : gv;
    }
}
extern "C" int mq_clip_forward(const float *x, float *y, int64_t n, float lo, float hi, void *stream) {
    if (n <= 0) return MQ_OK;
    MQ_LAUNCH(clip_fwd_kernel, ew_blocks(n), MQ_NT, 0, stream, x, y, n, lo, hi);
    MQ_CHECK_LAUNCH("mq_clip_forward");
    return MQ_OK;
}
extern "C" int mq_clip_backward(const float *x, const float *g, float *out, int64_t n, float lo,
                                float hi, void *stream) {
    if (n <= 0) return MQ_OK;
    MQ_LAUNCH(clip_bwd_kernel, ew_blocks(n), MQ_NT, 0, stream, x, g, out, n, lo, hi);
    MQ_CHECK_LAUNCH("mq_clip_backward");
    return MQ_OK;
}

// ---- mannequin/distrib.py:41-76: diagonal Gaussian ------------------------------------
#define MQ_HALF_LOG_2PI 0.91893853320467274178f

__global__ void __launch_bounds__(MQ_NT)
gauss_logp_kernel(const float *__restrict__ mu, const float *__restrict__ ls, int64_t period,
                  const float *__restrict__ s, int64_t rows, int64_t a, float *__restrict__ logp) {
    EW_LOOP(r, rows) {
        float acc = 0.0f;
        for (int64_t j = 0; j < a; ++j) {
            const int64_t e = r * a + j;
            const float l = ls[e % period];
            const float z = (s[e] - mu[e]) * expf(-l);
            acc += l + 0.5f * z * z;
        }
        logp[r] = -(float)a * MQ_HALF_LOG_2PI - acc;
    }
}

// wide actions (the 784-pixel decoder of mnist/vae.py): one warp per row, lanes along the action axis
__global__ void __launch_bounds__(MQ_NT)
gauss_logp_wide_kernel(const float *__restrict__ mu, const float *__restrict__ ls, int64_t period,
                       const float *__restrict__ s, int64_t rows, int64_t a, float *__restrict__ logp) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < rows; r += nwarps) {
        float acc = 0.0f;
        for (int64_t j = lane; j < a; j += 32) {
            const int64_t e = r * a + j;
            const float l = ls[e % period];
            const float z = (s[e] - mu[e]) * expf(-l);
            acc += l + 0.5f * z * z;
        }
        acc = mq_warp_sum(acc);
        if (lane == 0) logp[r] = -(float)a * MQ_HALF_LOG_2PI - acc;
    }
}

__global__ void __launch_bounds__(MQ_NT)
gauss_logp_bwd_kernel(const float *__restrict__ mu, const float *__restrict__ ls, int64_t period,
                      const float *__restrict__ s, const float *__restrict__ g, int64_t n, int64_t a,
                      float *__restrict__ d_mu, float *__restrict__ d_ls) {
    EW_LOOP(e, n) {
        const float inv = expf(-ls[e % period]);
        const float z = (s[e] - mu[e]) * inv;
        const float gv = g[e / a];
        d_mu[e] = gv * z * inv;
        d_ls[e] = gv * (z * z - 1.0f);
    }
}

__global__ void __launch_bounds__(MQ_NT)
gauss_sample_kernel(const float *__restrict__ mu, const float *__restrict__ ls, int64_t period,
                    const float *__restrict__ unit, int64_t n, float *__restrict__ out,
                    float *__restrict__ noise) {
    EW_LOOP(e, n) {
        const float nz = unit[e] * expf(ls[e % period]);
        noise[e] = nz;
        out[e] = mu[e] + nz;
    }
}

extern "C" int mq_gauss_logprob(const float *mu, const float *logstd, int64_t logstd_period,
                                const float *sample, int64_t rows, int64_t a, float *logp,
                                void *stream) {
    MQ_REQUIRE(rows >= 0 && a >= 1 && logstd_period >= 1, MQ_ERR_SHAPE, "mq_gauss_logprob: bad shape");
    if (rows == 0) return MQ_OK;
    MQ_REQUIRE(mu && logstd && sample && logp, MQ_ERR_INVALID, "mq_gauss_logprob: null pointer");
    if (a >= 64)
        MQ_LAUNCH(gauss_logp_wide_kernel, ew_blocks(rows * 32), MQ_NT, 0, stream, mu, logstd, logstd_period, sample, rows,
                  a, logp);
    else
        MQ_LAUNCH(gauss_logp_kernel, ew_blocks(rows), MQ_NT, 0, stream, mu, logstd, logstd_period, sample, rows, a, logp);
    MQ_CHECK_LAUNCH("mq_gauss_logprob");
    return MQ_OK;
}

extern "C" int mq_gauss_logprob_bwd(const float *mu, const float *logstd, int64_t logstd_period,
                                    const float *sample, const float *g, int64_t rows, int64_t a,
                                    float *d_mu, float *d_logstd_rows, void *stream) {
    MQ_REQUIRE(rows >= 0 && a >= 1 && logstd_period >= 1, MQ_ERR_SHAPE, "mq_gauss_logprob_bwd: bad shape");
    if (rows == 0) return MQ_OK;
    MQ_REQUIRE(mu && logstd && sample && g && d_mu && d_logstd_rows, MQ_ERR_INVALID,
               "mq_gauss_logprob_bwd: null pointer");
    MQ_LAUNCH(gauss_logp_bwd_kernel, ew_blocks(rows * a), MQ_NT, 0, stream, mu, logstd, logstd_period,
              sample, g, rows * a, a, d_mu, d_logstd_rows);
    MQ_CHECK_LAUNCH("mq_gauss_logprob_bwd");
    return MQ_OK;
}

// Gaussian entropy (north_star names it, the reference has no such function):
// H[r] = sum_a logstd[r,a] + A/2 (1 + ln 2 pi);  dH/dlogstd = 1, dH/dmean = 0.
__global__ void __launch_bounds__(MQ_NT)
gauss_entropy_kernel(const float *__restrict__ ls, int64_t rows, int64_t a, float *__restrict__ h) {
    EW_LOOP(r, rows) {
        float acc = 0.0f;
        for (int64_t j = 0; j < a; ++j) acc += ls[r * a + j];
        h[r] = acc + (float)a * (0.5f + MQ_HALF_LOG_2PI);
    }
}
__global__ void __launch_bounds__(MQ_NT)
gauss_entropy_bwd_kernel(const float *__restrict__ g, int64_t n, int64_t a, float *__restrict__ d_ls) {
    EW_LOOP(e, n) d_ls[e] = g[e / a];
}

extern "C" int mq_gauss_entropy(const float *logstd, const float *g, int64_t rows, int64_t a, float *h,
                                float *d_logstd, void *stream) {
    MQ_REQUIRE(rows >= 0 && a >= 1, MQ_ERR_SHAPE, "mq_gauss_entropy: bad shape");
    if (rows == 0) return MQ_OK;
    MQ_REQUIRE((h && logstd) || (d_logstd && g), MQ_ERR_INVALID, "mq_gauss_entropy: null pointer");
    if (h) MQ_LAUNCH(gauss_entropy_kernel, ew_blocks(rows), MQ_NT, 0, stream, logstd, rows, a, h);
    if (d_logstd) MQ_LAUNCH(gauss_entropy_bwd_kernel, ew_blocks(rows * a), MQ_NT, 0, stream, g, rows * a, a, d_logstd);
    MQ_CHECK_LAUNCH("mq_gauss_entropy");
    return MQ_OK;
}

extern "C" int mq_gauss_sample(const float *mu, const float *logstd, int64_t logstd_period,
                               const float *unit, int64_t rows, int64_t a, float *out, float *noise,
                               void *stream) {
    MQ_REQUIRE(rows >= 0 && a >= 1 && logstd_period >= 1, MQ_ERR_SHAPE, "mq_gauss_sample: bad shape");
    if (rows == 0) return MQ_OK;
    MQ_REQUIRE(mu && logstd && unit && out && noise, MQ_ERR_INVALID, "mq_gauss_sample: null pointer");
    MQ_LAUNCH(gauss_sample_kernel, ew_blocks(rows * a), MQ_NT, 0, stream, mu, logstd, logstd_period,
              unit, rows * a, out, noise);
    MQ_CHECK_LAUNCH("mq_gauss_sample");
    return MQ_OK;
}

// ---- mnist/vae.py:13-22: KL(N(mean, exp(logstd)) || N(0,1)) as written there -----------
__global__ void __launch_bounds__(MQ_NT)
dkl_fwd_kernel(const float *__restrict__ mean, const float *__restrict__ ls, int64_t rows, int64_t d,
               float *__restrict__ out) {
    EW_LOOP(r, rows) {
        float acc = 0.0f;
        for (int64_t j = 0; j < d; ++j) {
            const float m = mean[r * d + j], l = ls[r * d + j];
            acc += expf(l) - l + m * m;
        }
        out[r] = 0.5f * (acc - (float)d);
    }
}
__global__ void __launch_bounds__(MQ_NT)
dkl_bwd_kernel(const float *__restrict__ mean, const float *__restrict__ ls,
               const float *__restrict__ g, int64_t n, int64_t d, float *__restrict__ d_mean,
               float *__restrict__ d_ls) {
    EW_LOOP(e, n) {
        const float gv = g[e / d];
        d_mean[e] = gv * mean[e];
        d_ls[e] = 0.5f * gv * (expf(ls[e]) - 1.0f);
    }
}
extern "C" int mq_dkl_forward(const float *mean, const float *logstd, int64_t rows, int64_t d,
                              float *out, void *stream) {
    MQ_REQUIRE(rows >= 0 && d >= 1, MQ_ERR_SHAPE, "mq_dkl_forward: bad shape");
    if (rows == 0) return MQ_OK;
    MQ_LAUNCH(dkl_fwd_kernel, ew_blocks(rows), MQ_NT, 0, stream, mean, logstd, rows, d, out);
    MQ_CHECK_LAUNCH("mq_dkl_forward");
    return MQ_OK;
}
extern "C" int mq_dkl_backward(const float *mean, const float *logstd, const float *g, int64_t rows,
                               int64_t d, float *d_mean, float *d_logstd, void *stream) {
    MQ_REQUIRE(rows >= 0 && d >= 1, MQ_ERR_SHAPE, "mq_dkl_backward: bad shape");
    if (rows == 0) return MQ_OK;
    MQ_LAUNCH(dkl_bwd_kernel, ew_blocks(rows * d), MQ_NT, 0, stream, mean, logstd, g, rows * d, d, d_mean, d_logstd);
    MQ_CHECK_LAUNCH("mq_dkl_backward");
    return MQ_OK;
}

// ---- mnist/mlp.py:10-14,33: dy = onehot - softmax(logits) - l2*logits ------------------
__global__ void __launch_bounds__(MQ_NT)
softmax_ce_kernel(const float *__restrict__ logits, const float *__restrict__ onehot,
                  const int32_t *__restrict__ row_idx, int64_t rows, int64_t cols, float l2,
                  float *__restrict__ dy) {
    EW_LOOP(r, rows) {
        const float *z = logits + r * cols;
        const float *t = onehot + (row_idx ? (int64_t)row_idx[r] : r) * cols;
        float mx = z[0];
        for (int64_t j = 1; j < cols; ++j) mx = fmaxf(mx, z[j]);
        float sum = 0.0f;
        for (int64_t j = 0; j < cols; ++j) sum += expf(z[j] - mx);
        const float inv = 1.0f / sum;
        for (int64_t j = 0; j < cols; ++j) dy[r * cols + j] = t[j] - expf(z[j] - mx) * inv - l2 * z[j];
    }
}
extern "C" int mq_softmax_ce_grad(const float *logits, const float *onehot, const int32_t *row_idx,
                                  int64_t rows, int64_t cols, float l2, float *dy, void *stream) {
    MQ_REQUIRE(rows >= 0 && cols >= 1, MQ_ERR_SHAPE, "mq_softmax_ce_grad: bad shape");
    if (rows == 0) return MQ_OK;
    MQ_REQUIRE(logits && onehot && dy, MQ_ERR_INVALID, "mq_softmax_ce_grad: null pointer");
    MQ_LAUNCH(softmax_ce_kernel, ew_blocks(rows), MQ_NT, 0, stream, logits, onehot, row_idx, rows, cols, l2, dy);
    MQ_CHECK_LAUNCH("mq_softmax_ce_grad");
    return MQ_OK;
}

// ---- benchmarks/ppo.py:34-37: ratio, clip rule, times advantage -------------------------
__device__ __forceinline__ float mq_ppo_rule(float logp, float logp_old, float adv, float lo, float hi) {
    const float ratio = expf(logp - logp_old);
    const bool kill = (ratio > hi && adv > 0.0f) || (ratio < lo && adv < 0.0f);
    return kill ? 0.0f : ratio * adv;
}
__global__ void __launch_bounds__(MQ_NT)
ppo_weight_kernel(const float *__restrict__ logp, const float *__restrict__ logp_old,
                  const float *__restrict__ adv, const int32_t *__restrict__ row_idx, int64_t rows,
                  float lo, float hi, float *__restrict__ w) {
    EW_LOOP(r, rows) {
        const int64_t s = row_idx ? (int64_t)row_idx[r] : r;
        w[r] = mq_ppo_rule(logp[r], logp_old[s], adv[s], lo, hi);
    }
}
extern "C" int mq_ppo_weight(const float *logp, const float *logp_old, const float *adv,
                             const int32_t *row_idx, int64_t rows, float lo, float hi, float *w,
                             void *stream) {
    MQ_REQUIRE(rows >= 0, MQ_ERR_SHAPE, "mq_ppo_weight: bad shape");
    if (rows == 0) return MQ_OK;
    MQ_REQUIRE(logp && logp_old && adv && w, MQ_ERR_INVALID, "mq_ppo_weight: null pointer");
    MQ_LAUNCH(ppo_weight_kernel, ew_blocks(rows), MQ_NT, 0, stream, logp, logp_old, adv, row_idx, rows, lo, hi, w);
    MQ_CHECK_LAUNCH("mq_ppo_weight");
    return MQ_OK;
}

// ---- categorical log-prob: mannequin/distrib.py:24-32 ---------------------------------------------
// logp[r] = log softmax(logits[r])[sample[r]];  d_logits[r] = g[r] * (onehot(sample[r]) - softmax(logits[r]))
// (g == NULL: forward only; logp / d_logits may each be NULL).  One thread per row.
__global__ void __launch_bounds__(MQ_NT)
discrete_logp_kernel(const float *__restrict__ logits, const int32_t *__restrict__ sample, const float *__restrict__ g,
                     int64_t rows, int64_t n, float *__restrict__ logp, float *__restrict__ d_logits) {
    EW_LOOP(r, rows) {
        const float *y = logits + r * n;
        float mx = y[0];
        for (int64_t k = 1; k < n; ++k) mx = fmaxf(mx, y[k]);
        float sum = 0.0f;
        for (int64_t k = 0; k < n; ++k) sum += expf(y[k] - mx);
        const float lse = mx + logf(sum);
        const int32_t a = sample[r];
        if (logp) logp[r] = y[a] - lse;
        if (d_logits) {
            const float gv = g[r];
            for (int64_t k = 0; k < n; ++k) d_logits[r * n + k] = gv * ((k == a ? 1.0f : 0.0f) - expf(y[k] - lse));
        }
    }
}

extern "C" int mq_discrete_logprob(const float *logits, const int32_t *sample, const float *g, int64_t rows, int64_t n,
                                   float *logp, float *d_logits, void *stream) {
    MQ_REQUIRE(rows >= 0 && n >= 1, MQ_ERR_SHAPE, "mq_discrete_logprob: bad shape");
    if (rows == 0) return MQ_OK;
    MQ_REQUIRE(logits && sample && (logp || d_logits) && (!d_logits || g), MQ_ERR_INVALID,
               "mq_discrete_logprob: null pointer");
    MQ_LAUNCH(discrete_logp_kernel, ew_blocks(rows), MQ_NT, 0, stream, logits, sample, g, rows, n, logp, d_logits);
    MQ_CHECK_LAUNCH("mq_discrete_logprob");
    return MQ_OK;
}

// onehot[r, idx[r]] = 1 (rows x n, zero elsewhere): actions as the target rows of the categorical heads
__global__ void __launch_bounds__(MQ_NT)
onehot_kernel(const int32_t *__restrict__ idx, int64_t rows, int64_t n, float *__restrict__ out) {
    EW_LOOP(e, rows * n) { out[e] = (e % n == idx[e / n]) ? 1.0f : 0.0f; }
}

extern "C" int mq_onehot(const int32_t *idx, int64_t rows, int64_t n, float *out, void *stream) {
    MQ_REQUIRE(rows >= 0 && n >= 1, MQ_ERR_SHAPE, "mq_onehot: bad shape");
    if (rows == 0) return MQ_OK;
    MQ_REQUIRE(idx && out, MQ_ERR_INVALID, "mq_onehot: null pointer");
    MQ_LAUNCH(onehot_kernel, ew_blocks(rows * n), MQ_NT, 0, stream, idx, rows, n, out);
    MQ_CHECK_LAUNCH("mq_onehot");
    return MQ_OK;
}

// ---- clipped-momentum ascent of mnist/vae.py:64-69 ------------------------------------------------------
// g = g1 (+ g2 on the first n2 entries: the encoder parameters come first); mom = clip(decay*mom + (1-decay)*g, +-clip);
// theta += lr * mom
__global__ void __launch_bounds__(MQ_NT)
momentum_step_kernel(float *__restrict__ theta, float *__restrict__ mom, const float *__restrict__ g1,
                     const float *__restrict__ g2, int64_t n2, int64_t n, float decay, float clip, float lr) {
    EW_LOOP(e, n) {
        float g = g1[e];
        if (g2 && e < n2) g += g2[e];
        float m = mom[e] * decay + g * (1.0f - decay);
        m = fminf(fmaxf(m, -clip), clip);
        mom[e] = m;
        theta[e] += lr * m;
    }
}

extern "C" int mq_momentum_step(float *theta, float *mom, const float *g1, const float *g2, int64_t n2, int64_t n,
                                float decay, float clip, float lr, void *stream) {
    MQ_REQUIRE(n >= 0 && n2 >= 0 && n2 <= n, MQ_ERR_SHAPE, "mq_momentum_step: bad sizes");
    if (n == 0) return MQ_OK;
    MQ_REQUIRE(theta && mom && g1, MQ_ERR_INVALID, "mq_momentum_step: null pointer");
    MQ_LAUNCH(momentum_step_kernel, ew_blocks(n), MQ_NT, 0, stream, theta, mom, g1, g2, n2, n, decay, clip, lr);
    MQ_CHECK_LAUNCH("mq_momentum_step");
    return MQ_OK;
}
