// One fused plan for the training step of mnist/vae.py (BASELINE configs[4]: 784-400-(20+20)-400-(784+784), batch 4096).
//
// The reference step (mnist/vae.py:57-71) evaluates the shared graph twice -- decoder.logprob with upstream +1, dkl with
// upstream -1 -- adds the second gradient onto the encoder's slice of the first and takes a clipped-momentum step.  Every
// node downstream of the latent heads is linear in its upstream gradient, so ONE forward pass and ONE backward pass with the
// two upstream gradients added at (mean, logstd) of the encoder give the same parameter gradient up to fp32 rounding (the
// Clip rule of vae.py:26-34 depends on the sign of each upstream gradient and is applied to the two contributions
// SEPARATELY before they are added, as the reference does).  On the layer graph the step was ~110 launches; here it is
//
//   pack (weights in both operand layouts, the batch)                                      1 launch
//   encoder:  h = tanh(x W + b) per hidden layer (GEMM, epilogue emits h as fp32 AND as the operand planes the next
//             product and the dW product read), [mean | logstd] heads as ONE product           n_hidden + 1
//   latent:   Clip, reparameterised sample (vae.py:50 / distrib.py:48-51), DKL (vae.py:13-24)       1 (+ 1 pack of z)
//   decoder:  hidden layers, [mean | logstd] heads as one 400 x 1568 product                   n_hidden + 1
//   head:     Gaussian log-prob of the batch (distrib.py:61-68), Clip rule, dZ written as operand planes       1
//   backward: per layer dW = a^T dZ (+ bias row; split along the batch, fixed-order finish: 2 launches) and
//             dA = dZ W^T with tanh' in the epilogue, which again emits operand planes               3 per layer
//   latent backward, encoder backward, clipped-momentum step (vae.py:66-69)
//
// = 24 launches for one hidden layer per side, all products on the TMA-fed tcgen05 GEMM (mq_gemm_tma.cu) in the operand
// format `tc` names (default fp16x2: fp32 products).  Bit-reproducible: fixed split-K order, no atomics.
//
// Parameter layout (the flat vector of decoder.logprob, DFS post-order of mannequin/basicnet.py:66-77):
//   encoder hidden layers (W [in, H], b [H]) ..., W_mu [H, L], b_mu [L], W_ls [H, L], b_ls [L],
//   decoder hidden layers (W [in, H], b [H]) ..., W_m [H, D], b_m [D], W_s [H, D], b_s [D].
#include "mq_gemm.cuh"

#ifndef MQ_EMU
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#define VAE_MAX_HIDDEN 4
#define VAE_HALF_LOG_2PI 0.91893853320467274178f

namespace {

struct VaeDims {
    int64_t B, D, H, L, n;          // batch, pixels, hidden width, latent width, hidden layers per side
    int64_t Lp, Dp;                 // head widths rounded up to 8: the logstd head starts at column Lp / Dp of a head product
    int planes, mode;
    int64_t Bp, Bk;                 // batch as rows (multiple of 128) and as a contraction (multiple of 64)
};

static inline int64_t rp(int64_t rows) { return mq_round_up(rows, 128); }
static inline int64_t kp(int64_t k) { return mq_round_up(k, 64); }

struct VaeOffsets {
    int64_t we[VAE_MAX_HIDDEN], be[VAE_MAX_HIDDEN], w_mu, b_mu, w_ls, b_ls;
    int64_t wd[VAE_MAX_HIDDEN], bd[VAE_MAX_HIDDEN], w_m, b_m, w_s, b_s, total;
};

static VaeOffsets vae_offsets(const VaeDims &d) {
    VaeOffsets o;
    int64_t pos = 0;
    for (int i = 0; i < d.n; ++i) {
        const int64_t in = i == 0 ? d.D : d.H;
        o.we[i] = pos; pos += in * d.H;
        o.be[i] = pos; pos += d.H;
    }
    o.w_mu = pos; pos += d.H * d.L; o.b_mu = pos; pos += d.L;
    o.w_ls = pos; pos += d.H * d.L; o.b_ls = pos; pos += d.L;
    for (int i = 0; i < d.n; ++i) {
        const int64_t in = i == 0 ? d.L : d.H;
        o.wd[i] = pos; pos += in * d.H;
        o.bd[i] = pos; pos += d.H;
    }
    o.w_m = pos; pos += d.H * d.D; o.b_m = pos; pos += d.D;
    o.w_s = pos; pos += d.H * d.D; o.b_s = pos; pos += d.D;
    o.total = pos;
    return o;
}

// Workspace: one bump allocation, the same walk sizes it and hands out the pointers.
struct VaeWs {
    // operand planes: *k = K-major operand [planes][rows_pad][k_pad] of the tensor itself, *t = of its transpose
    void *xk, *xt, *zk, *zt;
    void *ek[VAE_MAX_HIDDEN], *et[VAE_MAX_HIDDEN], *dk[VAE_MAX_HIDDEN], *dt[VAE_MAX_HIDDEN];     // hidden activations, encoder / decoder
    float *ea[VAE_MAX_HIDDEN], *da[VAE_MAX_HIDDEN];                                               // the same in fp32 (for tanh')
    void *wf_e[VAE_MAX_HIDDEN], *wb_e[VAE_MAX_HIDDEN], *wf_d[VAE_MAX_HIDDEN], *wb_d[VAE_MAX_HIDDEN];   // weights: forward (W^T rows) / backward (W rows)
    void *wf_eh, *wb_eh, *wf_dh, *wb_dh;                                                        // the two heads side by side
    float *ch, *mu, *ls, *ls_raw, *noise, *z, *cd, *dz, *dzh;
    void *gdk, *gdt, *ghk, *ght;                                                                // dZ of the decoder / encoder heads
    void *gk_d[VAE_MAX_HIDDEN], *gt_d[VAE_MAX_HIDDEN], *gk_e[VAE_MAX_HIDDEN], *gt_e[VAE_MAX_HIDDEN];   // dZ of the hidden layers
    float *partials;
    int64_t bytes;
};

static int64_t vae_dw_splits(int64_t m, int64_t n, int64_t k, int64_t *k_per_split) {
    const int sms = mq_sm_count();
    const int64_t tiles = mq_ceil_div(m, 128) * mq_ceil_div(n, 128);
    int64_t splits = sms / tiles;
    if (splits > k / 256) splits = k / 256;
    if (splits < 1) splits = 1;
    *k_per_split = mq_round_up(mq_ceil_div(k, splits), 64);
    return mq_ceil_div(k, *k_per_split);
}

static VaeWs vae_ws(const VaeDims &d, void *base) {
    VaeWs w;
    memset(&w, 0, sizeof(w));
    uintptr_t p = ((uintptr_t)base + 1023) & ~(uintptr_t)1023;
    auto take = [&](int64_t bytes) { void *r = (void *)p; p += (uintptr_t)mq_round_up(bytes, 1024); return r; };
    auto planes = [&](int64_t rows_pad, int64_t k_pad) { return take((int64_t)d.planes * rows_pad * k_pad * 2); };
    auto f32 = [&](int64_t count) { return (float *)take(count * 4); };
    w.xk = planes(d.Bp, kp(d.D)); w.xt = planes(rp(d.D + 1), d.Bk);
    w.zk = planes(d.Bp, kp(d.L)); w.zt = planes(rp(d.L + 1), d.Bk);
    for (int i = 0; i < d.n; ++i) {
        w.ek[i] = planes(d.Bp, kp(d.H)); w.et[i] = planes(rp(d.H + 1), d.Bk); w.ea[i] = f32(d.B * d.H);
        w.dk[i] = planes(d.Bp, kp(d.H)); w.dt[i] = planes(rp(d.H + 1), d.Bk); w.da[i] = f32(d.B * d.H);
        const int64_t in_e = i == 0 ? d.D : d.H, in_d = i == 0 ? d.L : d.H;
        w.wf_e[i] = planes(rp(d.H), kp(in_e)); w.wb_e[i] = planes(rp(in_e), kp(d.H));
        w.wf_d[i] = planes(rp(d.H), kp(in_d)); w.wb_d[i] = planes(rp(in_d), kp(d.H));
        w.gk_d[i] = planes(d.Bp, kp(d.H)); w.gt_d[i] = planes(rp(d.H), d.Bk);
        w.gk_e[i] = planes(d.Bp, kp(d.H)); w.gt_e[i] = planes(rp(d.H), d.Bk);
    }
    w.wf_eh = planes(rp(2 * d.Lp), kp(d.H)); w.wb_eh = planes(rp(d.H), kp(2 * d.Lp));
    w.wf_dh = planes(rp(2 * d.Dp), kp(d.H)); w.wb_dh = planes(rp(d.H), kp(2 * d.Dp));
    w.ch = f32(d.B * 2 * d.Lp); w.mu = f32(d.B * d.L); w.ls = f32(d.B * d.L); w.ls_raw = f32(d.B * d.L);
    w.noise = f32(d.B * d.L); w.z = f32(d.B * d.L); w.cd = f32(d.B * 2 * d.Dp); w.dz = f32(d.B * d.L); w.dzh = f32(d.B * 2 * d.Lp);
    w.gdk = planes(d.Bp, kp(2 * d.Dp)); w.gdt = planes(rp(2 * d.Dp), d.Bk);
    w.ghk = planes(d.Bp, kp(2 * d.Lp)); w.ght = planes(rp(2 * d.Lp), d.Bk);
    // split-K partials of the largest dW product
    int64_t most = 0, kps;
    const int64_t shapes[4][2] = {{d.D + 1, d.H}, {d.H + 1, d.H}, {d.H + 1, 2 * d.Lp}, {d.H + 1, 2 * d.Dp}};
    for (int s = 0; s < 4; ++s) {
        const int64_t sp = vae_dw_splits(shapes[s][0], shapes[s][1], d.B, &kps);
        if (sp * shapes[s][0] * shapes[s][1] > most) most = sp * shapes[s][0] * shapes[s][1];
    }
    const int64_t lh = vae_dw_splits(d.L + 1, d.H, d.B, &kps) * (d.L + 1) * d.H;
    if (lh > most) most = lh;
    w.partials = f32(most);
    w.bytes = (int64_t)(p - (uintptr_t)base);
    return w;
}

static int vae_dims(const mq_vae_t *v, VaeDims *d, const char *what) {
    MQ_REQUIRE(v, MQ_ERR_INVALID, "%s: null descriptor", what);
    MQ_REQUIRE(v->batch >= 1 && v->d_img >= 1 && v->hidden >= 1 && v->latent >= 1, MQ_ERR_SHAPE, "%s: bad sizes", what);
    MQ_REQUIRE(v->n_hidden >= 1 && v->n_hidden <= VAE_MAX_HIDDEN, MQ_ERR_UNSUPPORTED, "%s: 1..%d hidden layers per side", what, VAE_MAX_HIDDEN);
    // the activation planes are written 8 columns at a time by the GEMM epilogue
    MQ_REQUIRE(v->hidden % 8 == 0, MQ_ERR_UNSUPPORTED, "%s: hidden width must be a multiple of 8", what);
    MQ_REQUIRE(v->batch < (1LL << 30) && v->d_img < (1LL << 24) && v->hidden < (1LL << 24) && v->latent < (1LL << 24), MQ_ERR_UNSUPPORTED, "%s: too large", what);
    d->B = v->batch; d->D = v->d_img; d->H = v->hidden; d->L = v->latent; d->n = v->n_hidden;
    d->Lp = mq_round_up(d->L, 8); d->Dp = mq_round_up(d->D, 8);
    d->mode = (v->tc & (MQ_TC_PLANES_MASK | MQ_TC_FP16)) ? (v->tc & (MQ_TC_PLANES_MASK | MQ_TC_FP16)) : (2 | MQ_TC_FP16);
    d->planes = d->mode & MQ_TC_PLANES_MASK;
    MQ_REQUIRE(d->planes >= 1 && (!(d->mode & MQ_TC_FP16) || d->planes == 2), MQ_ERR_INVALID, "%s: operand format", what);
    d->Bp = rp(d->B); d->Bk = kp(d->B);
    return MQ_OK;
}

// ---- side stream for the weight-gradient products -----------------------------------------------------------------------
// Nothing but the final momentum step reads a weight gradient, while the input-gradient products form the critical path of the
// backward pass and several of them are small (32-64 CTAs).  The dW products (+ their finishes) therefore run on a side stream,
// forked after the kernel that produces their operand and joined before the step ends: event record / wait pairs, which a CUDA
// graph capture of the calling stream turns into fork / join edges.  Stream and events are created once by mq_vae_ws_init
// (outside any capture); without them the plan runs serially on the caller's stream.
struct VaeAsync { cudaStream_t side; cudaEvent_t ev[2 * VAE_MAX_HIDDEN + 4]; bool ok; };
static VaeAsync g_async = {};

static void vae_async_init() {
    if (g_async.ok) return;
    if (cudaStreamCreateWithFlags(&g_async.side, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return; }
    for (auto &e : g_async.ev)
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return; }
    g_async.ok = true;
}

// ---- kernels ---------------------------------------------------------------------------------------------------------

// the row of ones (plane 0 only) under the transposed activations written by the GEMM epilogue: bias gradients as one
// more output row of dW = a^T dZ
__global__ void __launch_bounds__(MQ_NT)
vae_ones_row_kernel(unsigned short *__restrict__ row, int64_t n, unsigned short one) {
    for (int64_t e = (int64_t)blockIdx.x * MQ_NT + threadIdx.x; e < n; e += (int64_t)gridDim.x * MQ_NT) row[e] = one;
}

__device__ __forceinline__ float vae_clip_rule(float raw, float g, float lo, float hi) {      // vae.py:28-32
    if (raw < lo && g < 0.0f) return 0.0f;
    if (raw > hi && g > 0.0f) return 0.0f;
    return g;
}

// Encoder heads -> latent sample.  ch = h [W_mu | W_ls] (columns 0 .. L-1 and Lp .. Lp+L-1), biases added here.
//   ls = clip(raw)  (vae.py:26-34), noise = unit * exp(ls), z = mu + noise  (distrib.py:48-51),
//   dkl = 0.5 (sum(exp(ls) - ls + mu^2) - L)  (vae.py:13-22).  One warp per row.
__global__ void __launch_bounds__(MQ_NT)
vae_latent_fwd_kernel(const float *__restrict__ ch, const float *__restrict__ b_mu, const float *__restrict__ b_ls,
                      const float *__restrict__ unit, int64_t rows, int L, int Lp, float lo, float hi, float *__restrict__ mu,
                      float *__restrict__ ls, float *__restrict__ ls_raw, float *__restrict__ noise, float *__restrict__ z,
                      float *__restrict__ dkl) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * MQ_NT + threadIdx.x) >> 5;      // one warp per row, lanes along the latent axis
    if (r >= rows) return;
    float acc = 0.0f;
    for (int j = lane; j < L; j += 32) {
        const float m = ch[r * 2 * Lp + j] + b_mu[j];
        const float raw = ch[r * 2 * Lp + Lp + j] + b_ls[j];
        const float l = fminf(fmaxf(raw, lo), hi);
        const float e = expf(l);
        const float nz = unit[r * L + j] * e;
        mu[r * L + j] = m; ls[r * L + j] = l; ls_raw[r * L + j] = raw; noise[r * L + j] = nz; z[r * L + j] = m + nz;
        acc += e - l + m * m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dkl[r] = 0.5f * (acc - (float)L);
}

// Upstream gradient of the encoder heads: the logprob pass (+1) arrives through the sample (distrib.py:51: (g, g * noise)),
// the dkl pass with upstream -1 (vae.py:62): d/dmean = mean, d/dlogstd = 0.5 (exp(logstd) - 1).  Clip rule per pass.
// dz already carries the factor `up` = 1 / batch of the head kernel; the dkl terms get it here.
__global__ void __launch_bounds__(MQ_NT)
vae_latent_bwd_kernel(const float *__restrict__ dz, const float *__restrict__ mu, const float *__restrict__ ls,
                      const float *__restrict__ ls_raw, const float *__restrict__ noise, int64_t n, int L, int Lp, float lo, float hi,
                      float up, float *__restrict__ dzh) {
    const int64_t e = (int64_t)blockIdx.x * MQ_NT + threadIdx.x;
    if (e >= n) return;
    const int64_t r = e / L;
    const int j = (int)(e - r * L);
    const float g = dz[e];
    const float g1 = vae_clip_rule(ls_raw[e], g * noise[e], lo, hi);
    const float g2 = vae_clip_rule(ls_raw[e], -0.5f * up * (expf(ls[e]) - 1.0f), lo, hi);
    dzh[r * 2 * Lp + j] = g - up * mu[e];
    dzh[r * 2 * Lp + Lp + j] = g1 + g2;
}

// Decoder heads -> log-prob of the batch and dZ as OPERAND PLANES.  cd = h [W_m | W_s] (columns 0 .. D-1 and Dp .. Dp+D-1).
//   logprob = -D/2 ln 2 pi - sum(ls + 0.5 ((x - mean) / exp(ls))^2)  (distrib.py:61-68) with ls = clip(raw);
//   d/dmean = (x - mean) exp(-2 ls), d/dls = ((x - mean) exp(-ls))^2 - 1, then the Clip rule, times the upstream
//   gradient `up` = 1 / batch: parameter gradients are batch MEANS (basicnet.py:231-236), and dividing here instead of
//   after the dW products keeps dZ -- up to exp(12) (x - mean)^2 at the lower clip -- inside the fp16 range of the
//   fp16x2 operand format for any realistic batch.
// dZ (batch x 2 Dp, the largest tensor of the step) is never stored as fp32: a block owns 16 batch rows and walks the pixels
// in blocks of 128; 16 threads take a row and each writes its 8 + 8 values straight into the K-major planes the dA product
// reads, and, through a shared-memory tile, thread-per-pixel into the planes of dZ^T the dW product reads (the block's 16
// batch rows are one 32-byte sector of a row of dZ^T).
#define VAE_HEAD_ROWS 16
#define VAE_HEAD_COLS 128
template <int P, int F>
__global__ void __launch_bounds__(MQ_NT)
vae_head_emit_kernel(const float *__restrict__ cd, const float *__restrict__ b_m, const float *__restrict__ b_s,
                     const float *__restrict__ x, int64_t rows, int D, int Dp, float lo, float hi, float up, float *__restrict__ logprob,
                     __nv_bfloat16 *__restrict__ out_k, int64_t k_rows_pad, int64_t k_kpad, __nv_bfloat16 *__restrict__ out_t, int64_t t_rows_pad,
                     int64_t t_kpad) {
    __shared__ float tile[2][VAE_HEAD_ROWS][VAE_HEAD_COLS + 1];           // [mean | logstd][batch row][pixel]
    const int row = threadIdx.x >> 4, grp = threadIdx.x & 15;             // phase 1: 16 threads per batch row, 8 pixels each
    const int64_t b0 = (int64_t)blockIdx.x * VAE_HEAD_ROWS, b = b0 + row;
    const bool row_ok = b < rows;
    float acc = 0.0f;
    for (int j0 = 0; j0 < D; j0 += VAE_HEAD_COLS) {
        const int j = j0 + grp * 8;
        float dm[8], dl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { dm[i] = 0.0f; dl[i] = 0.0f; }
        if (row_ok && j < D) {
            const float *cm = cd + b * 2 * Dp + j, *cl = cm + Dp, *xr = x + b * D + j;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (j + i < D) {
                    const float m = cm[i] + b_m[j + i];
                    const float raw = cl[i] + b_s[j + i];
                    const float l = fminf(fmaxf(raw, lo), hi);
                    const float inv = expf(-l);
                    const float zz = (xr[i] - m) * inv;
                    acc += l + 0.5f * zz * zz;
                    dm[i] = up * (zz * inv);
                    dl[i] = up * vae_clip_rule(raw, zz * zz - 1.0f, lo, hi);
                }
            }
            uint32_t wm[4][P], wl[4][P];
#pragma unroll
            for (int i = 0; i < 4; ++i) { gm_split_pair<P, F>(dm[2 * i], dm[2 * i + 1], wm[i]); gm_split_pair<P, F>(dl[2 * i], dl[2 * i + 1], wl[i]); }
#pragma unroll
            for (int pl = 0; pl < P; ++pl) {
                __nv_bfloat16 *o = out_k + ((int64_t)pl * k_rows_pad + b) * k_kpad + j;
                *reinterpret_cast<uint4 *>(o) = make_uint4(wm[0][pl], wm[1][pl], wm[2][pl], wm[3][pl]);
                *reinterpret_cast<uint4 *>(o + Dp) = make_uint4(wl[0][pl], wl[1][pl], wl[2][pl], wl[3][pl]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { tile[0][row][grp * 8 + i] = dm[i]; tile[1][row][grp * 8 + i] = dl[i]; }
        __syncthreads();
        // phase 2: a thread per (head, pixel): the 16 batch rows of the block are one 32-byte sector of a row of dZ^T (with 8
        // rows per block every store was half a sector and the kernel took as long as the fp32 round trip it replaces)
        {
            const int tcol = threadIdx.x & (VAE_HEAD_COLS - 1), h = threadIdx.x >> 7;
            if (j0 + tcol < D) {
                const int64_t c = (int64_t)h * Dp + j0 + tcol;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if (b0 + q * 8 >= t_kpad) continue;
                    uint32_t w[4][P];
#pragma unroll
                    for (int i = 0; i < 4; ++i) gm_split_pair<P, F>(tile[h][q * 8 + 2 * i][tcol], tile[h][q * 8 + 2 * i + 1][tcol], w[i]);
#pragma unroll
                    for (int pl = 0; pl < P; ++pl)
                        *reinterpret_cast<uint4 *>(out_t + ((int64_t)pl * t_rows_pad + c) * t_kpad + b0 + q * 8) = make_uint4(w[0][pl], w[1][pl], w[2][pl], w[3][pl]);
                }
            }
        }
        __syncthreads();
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    if (grp == 0 && row_ok) logprob[b] = -(float)D * VAE_HALF_LOG_2PI - acc;
}

// dW (+ bias row) = alpha * sum over the K splits, in split order; columns [0, nd0) go to dst0 (row stride nd0), columns
// [off1, off1 + nd1) to dst1: the two heads of one product land in their own (W, b) slices of the gradient vector.
__global__ void __launch_bounds__(MQ_NT)
vae_dw_finish_kernel(const float *__restrict__ partial, int splits, int64_t m, int64_t n, float alpha, float *__restrict__ dst0,
                     int64_t nd0, float *__restrict__ dst1, int64_t off1, int64_t nd1) {
    const int64_t total = m * n;
    for (int64_t e = (int64_t)blockIdx.x * MQ_NT + threadIdx.x; e < total; e += (int64_t)gridDim.x * MQ_NT) {
        const int64_t r = e / n, c = e - r * n;
        float *dst;
        if (c < nd0) dst = dst0 + r * nd0 + c;
        else if (dst1 && c >= off1 && c < off1 + nd1) dst = dst1 + r * nd1 + (c - off1);
        else continue;
        float s = 0.0f;
#pragma unroll 8
        for (int z = 0; z < splits; ++z) s += __ldg(partial + (int64_t)z * total + e);      // split order: bit-reproducible
        *dst = alpha * s;
    }
}

// ---- host helpers ----------------------------------------------------------------------------------------------------
static GmPackJob vae_job(const float *base, int64_t rs, int64_t cs, int64_t rows, int64_t k, int64_t ones_row, void *out,
                         int64_t rows_pad, int64_t k_pad, int64_t row_off, int64_t k_off, bool whole) {
    GmPackJob jb;
    memset(&jb, 0, sizeof(jb));
    jb.base = base; jb.rs = rs; jb.cs = cs; jb.rows = rows; jb.k = k; jb.ones_row = ones_row; jb.out = out;
    jb.rows_pad = rows_pad; jb.k_pad = k_pad; jb.row_off = row_off; jb.k_off = k_off;
    jb.rows_cover = whole ? rows_pad : rows;
    jb.k_cover = whole ? k_pad : mq_round_up(k, 8);
    return jb;
}

struct VaeGemm {
    const void *pa; int64_t a_rows_pad; const void *pb; int64_t b_rows_pad;
    int64_t m, n, k;
};

// C = epilogue(A B) on packed operands, optionally emitting operand planes of C / C^T
static int vae_product(const VaeDims &d, const VaeGemm &p, float *c, int64_t ldc, int epi, const float *bias, const float *y,
                       void *out_k, int64_t k_rows_pad, int64_t k_kpad, void *out_t, int64_t t_rows_pad, int64_t t_kpad, void *stream) {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.c = c; g.m = p.m; g.n = p.n; g.k = p.k; g.k_per_split = kp(p.k); g.alpha = 1.0f; g.epi = epi; g.act = MQ_ACT_TANH; g.bias = bias; g.y = y;
    GmEmit em;
    memset(&em, 0, sizeof(em));
    em.out_k = out_k; em.k_rows_pad = k_rows_pad; em.k_kpad = k_kpad; em.out_t = out_t; em.t_rows_pad = t_rows_pad; em.t_kpad = t_kpad; em.ldc = ldc;
    return mq_gemm_tma_packed(p.pa, p.a_rows_pad, p.pb, p.b_rows_pad, kp(p.k), g, em, d.mode, 1, stream);
}

// dW = alpha A B split along K (the batch), finished in fixed order into one or two destinations
static int vae_dw(const VaeDims &d, const VaeGemm &p, float *partials, float alpha, float *dst0, int64_t nd0, float *dst1, int64_t off1,
                  int64_t nd1, void *stream) {
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.c = partials; g.m = p.m; g.n = p.n; g.k = p.k; g.alpha = 1.0f; g.epi = EPI_SCALE;
    const int64_t splits = vae_dw_splits(p.m, p.n, p.k, &g.k_per_split);
    GmEmit em;
    memset(&em, 0, sizeof(em));
    const int rc = mq_gemm_tma_packed(p.pa, p.a_rows_pad, p.pb, p.b_rows_pad, kp(p.k), g, em, d.mode, splits, stream);
    if (rc != MQ_OK) return rc;
    int64_t nb = mq_ceil_div(p.m * p.n, MQ_NT);
    if (nb > (int64_t)mq_sm_count() * 8) nb = (int64_t)mq_sm_count() * 8;
    MQ_LAUNCH(vae_dw_finish_kernel, (unsigned)nb, MQ_NT, 0, stream, (const float *)partials, (int)splits, p.m, p.n, alpha, dst0, nd0, dst1, off1, nd1);
    return MQ_OK;
}

}  // namespace

#define VAE_TRY(expr) do { const int rc_ = (expr); if (rc_ != MQ_OK) return rc_; } while (0)

extern "C" int64_t mq_vae_n_params(const mq_vae_t *v) {
    VaeDims d;
    if (vae_dims(v, &d, "mq_vae_n_params") != MQ_OK) return -1;
    return vae_offsets(d).total;
}

extern "C" int64_t mq_vae_ws_bytes(const mq_vae_t *v) {
    VaeDims d;
    if (vae_dims(v, &d, "mq_vae_ws_bytes") != MQ_OK) return -1;
    return vae_ws(d, nullptr).bytes + 2048;
}

extern "C" int mq_vae_ws_init(const mq_vae_t *v, void *ws, int64_t ws_bytes, void *stream) {
    VaeDims d;
    VAE_TRY(vae_dims(v, &d, "mq_vae_ws_init"));
    MQ_REQUIRE(ws && ws_bytes >= vae_ws(d, nullptr).bytes + 2048, MQ_ERR_INVALID, "mq_vae_ws_init: workspace smaller than mq_vae_ws_bytes()");
    // every plane image is zero outside what the packs and the GEMM epilogues write (padding rows and columns)
    if (cudaMemsetAsync(ws, 0, (size_t)ws_bytes, (cudaStream_t)stream) != cudaSuccess) return mq_check_launch("mq_vae_ws_init");
    vae_async_init();
    const VaeWs w = vae_ws(d, ws);
    const unsigned short one = (d.mode & MQ_TC_FP16) ? (unsigned short)0x3C00 : (unsigned short)0x3F80;      // 1.0 as fp16 / bf16
    for (int i = 0; i < d.n; ++i) {
        unsigned short *re = (unsigned short *)w.et[i] + d.H * d.Bk, *rd = (unsigned short *)w.dt[i] + d.H * d.Bk;
        MQ_LAUNCH(vae_ones_row_kernel, (unsigned)mq_ceil_div(d.B, MQ_NT), MQ_NT, 0, stream, re, d.B, one);
        MQ_LAUNCH(vae_ones_row_kernel, (unsigned)mq_ceil_div(d.B, MQ_NT), MQ_NT, 0, stream, rd, d.B, one);
    }
    MQ_CHECK_LAUNCH("mq_vae_ws_init");
    return MQ_OK;
}

// grad = d(sum logprob)/B - d(sum dkl)/B w.r.t. the flat parameters, i.e. grad1 with grad2 added onto the encoder's slice
// (vae.py:57-66; parameter gradients are batch means, basicnet.py:231-236), logprob[B] and dkl[B] the two values.
extern "C" int mq_vae_grad(const mq_vae_t *v, const float *theta, const float *x, const float *unit, float *grad, float *logprob,
                           float *dkl, void *ws, int64_t ws_bytes, void *stream) {
    VaeDims d;
    VAE_TRY(vae_dims(v, &d, "mq_vae_grad"));
    MQ_REQUIRE(theta && x && unit && grad && logprob && dkl && ws, MQ_ERR_INVALID, "mq_vae_grad: null pointer");
    MQ_REQUIRE(ws_bytes >= vae_ws(d, nullptr).bytes + 2048, MQ_ERR_INVALID, "mq_vae_grad: workspace smaller than mq_vae_ws_bytes()");
    const VaeWs w = vae_ws(d, ws);
    const VaeOffsets o = vae_offsets(d);
    const int64_t B = d.B, D = d.D, H = d.H, L = d.L, Lp = d.Lp, Dp = d.Dp;
    const int n = (int)d.n;
    const float lo = v->clip_lo, hi = v->clip_hi, inv_b = 1.0f / (float)B;

    // the side stream, ordered after everything issued to `stream` so far (or `stream` itself: serial plan)
    int n_forks = 0;
    auto side_after_main = [&]() -> void * {
        if (!g_async.ok) return stream;
        cudaEventRecord(g_async.ev[n_forks], (cudaStream_t)stream);
        cudaStreamWaitEvent(g_async.side, g_async.ev[n_forks], 0);
        ++n_forks;
        return (void *)g_async.side;
    };

    // ---- 1. weights in both operand layouts + the batch: one launch ---------------------------------------------------
    // forward operand of W [in, out]: rows = out, K = in  (B^T(n, k) = W[k, n]: rows are the unit-stride axis)
    // backward operand (dA = dZ W^T):  rows = in,  K = out (plain rows of W)
    {
        GmPackJob jobs[GM_MAX_PACK_JOBS];
        int nj = 0;
        jobs[nj++] = vae_job(x, D, 1, B, D, 0, w.xk, d.Bp, kp(D), 0, 0, true);
        jobs[nj++] = vae_job(x, 1, D, D + 1, B, D, w.xt, rp(D + 1), d.Bk, 0, 0, true);          // x^T with the row of ones
        for (int i = 0; i < n; ++i) {
            const int64_t in_e = i == 0 ? D : H, in_d = i == 0 ? L : H;
            jobs[nj++] = vae_job(theta + o.we[i], 1, H, H, in_e, 0, w.wf_e[i], rp(H), kp(in_e), 0, 0, true);
            if (i > 0) jobs[nj++] = vae_job(theta + o.we[i], H, 1, in_e, H, 0, w.wb_e[i], rp(in_e), kp(H), 0, 0, true);
            jobs[nj++] = vae_job(theta + o.wd[i], 1, H, H, in_d, 0, w.wf_d[i], rp(H), kp(in_d), 0, 0, true);
            jobs[nj++] = vae_job(theta + o.wd[i], H, 1, in_d, H, 0, w.wb_d[i], rp(in_d), kp(H), 0, 0, true);
        }
        if (nj + 8 > GM_MAX_PACK_JOBS) {          // deep nets: flush what we have
            VAE_TRY(mq_gemm_tma_pack_multi(jobs, nj, d.mode, stream));
            nj = 0;
        }
        // heads side by side: forward rows [0, L) | [Lp, Lp + L), backward K [0, L) | [Lp, Lp + L)
        jobs[nj++] = vae_job(theta + o.w_mu, 1, L, L, H, 0, w.wf_eh, rp(2 * Lp), kp(H), 0, 0, false);
        jobs[nj++] = vae_job(theta + o.w_ls, 1, L, L, H, 0, w.wf_eh, rp(2 * Lp), kp(H), Lp, 0, false);
        jobs[nj++] = vae_job(theta + o.w_mu, L, 1, H, L, 0, w.wb_eh, rp(H), kp(2 * Lp), 0, 0, false);
        jobs[nj++] = vae_job(theta + o.w_ls, L, 1, H, L, 0, w.wb_eh, rp(H), kp(2 * Lp), 0, Lp, false);
        jobs[nj++] = vae_job(theta + o.w_m, 1, D, D, H, 0, w.wf_dh, rp(2 * Dp), kp(H), 0, 0, false);
        jobs[nj++] = vae_job(theta + o.w_s, 1, D, D, H, 0, w.wf_dh, rp(2 * Dp), kp(H), Dp, 0, false);
        jobs[nj++] = vae_job(theta + o.w_m, D, 1, H, D, 0, w.wb_dh, rp(H), kp(2 * Dp), 0, 0, false);
        jobs[nj++] = vae_job(theta + o.w_s, D, 1, H, D, 0, w.wb_dh, rp(H), kp(2 * Dp), 0, Dp, false);
        VAE_TRY(mq_gemm_tma_pack_multi(jobs, nj, d.mode, stream));
    }

    // ---- 2. encoder ------------------------------------------------------------------------------------------------------
    for (int i = 0; i < n; ++i) {
        const int64_t in = i == 0 ? D : H;
        const VaeGemm p = {i == 0 ? w.xk : w.ek[i - 1], d.Bp, w.wf_e[i], rp(H), B, H, in};
        VAE_TRY(vae_product(d, p, w.ea[i], H, EPI_BIAS_ACT, theta + o.be[i], nullptr, w.ek[i], d.Bp, kp(H), w.et[i], rp(H + 1), d.Bk, stream));
    }
    {
        const VaeGemm p = {w.ek[n - 1], d.Bp, w.wf_eh, rp(2 * Lp), B, 2 * Lp, H};
        VAE_TRY(vae_product(d, p, w.ch, 2 * Lp, EPI_SCALE, nullptr, nullptr, nullptr, 0, 0, nullptr, 0, 0, stream));
    }
    MQ_LAUNCH(vae_latent_fwd_kernel, (unsigned)mq_ceil_div(B * 32, MQ_NT), MQ_NT, 0, stream, (const float *)w.ch, theta + o.b_mu, theta + o.b_ls, unit, B,
              (int)L, (int)Lp, lo, hi, w.mu, w.ls, w.ls_raw, w.noise, w.z, dkl);
    {
        GmPackJob jobs[2];
        jobs[0] = vae_job(w.z, L, 1, B, L, 0, w.zk, d.Bp, kp(L), 0, 0, true);
        jobs[1] = vae_job(w.z, 1, L, L + 1, B, L, w.zt, rp(L + 1), d.Bk, 0, 0, true);
        VAE_TRY(mq_gemm_tma_pack_multi(jobs, 2, d.mode, stream));
    }

    // ---- 3. decoder and its head -------------------------------------------------------------------------------------------
    for (int i = 0; i < n; ++i) {
        const int64_t in = i == 0 ? L : H;
        const VaeGemm p = {i == 0 ? w.zk : w.dk[i - 1], d.Bp, w.wf_d[i], rp(H), B, H, in};
        VAE_TRY(vae_product(d, p, w.da[i], H, EPI_BIAS_ACT, theta + o.bd[i], nullptr, w.dk[i], d.Bp, kp(H), w.dt[i], rp(H + 1), d.Bk, stream));
    }
    {
        const VaeGemm p = {w.dk[n - 1], d.Bp, w.wf_dh, rp(2 * Dp), B, 2 * Dp, H};
        VAE_TRY(vae_product(d, p, w.cd, 2 * Dp, EPI_SCALE, nullptr, nullptr, nullptr, 0, 0, nullptr, 0, 0, stream));
    }
    {
        const unsigned nb = (unsigned)mq_ceil_div(B, VAE_HEAD_ROWS);
        __nv_bfloat16 *gk = (__nv_bfloat16 *)w.gdk, *gt = (__nv_bfloat16 *)w.gdt;
#define VAE_HEAD_GO(P_, F_)                                                                                                            \
        do { auto kh = vae_head_emit_kernel<P_, F_>;                                                                                     \
             MQ_LAUNCH(kh, nb, MQ_NT, 0, stream, (const float *)w.cd, theta + o.b_m, theta + o.b_s, x, B, (int)D, (int)Dp, lo, hi, inv_b, \
                       logprob, gk, d.Bp, kp(2 * Dp), gt, rp(2 * Dp), d.Bk); } while (0)
        if (d.planes == 3) VAE_HEAD_GO(3, 0);
        else if (d.planes == 2 && (d.mode & MQ_TC_FP16)) VAE_HEAD_GO(2, 1);
        else if (d.planes == 2) VAE_HEAD_GO(2, 0);
        else VAE_HEAD_GO(1, 0);
#undef VAE_HEAD_GO
    }

    // ---- 4. decoder backward -----------------------------------------------------------------------------------------------
    {
        const VaeGemm pw = {w.dt[n - 1], rp(H + 1), w.gdt, rp(2 * Dp), H + 1, 2 * Dp, B};
        VAE_TRY(vae_dw(d, pw, w.partials, 1.0f, grad + o.w_m, D, grad + o.w_s, Dp, D, side_after_main()));
        const VaeGemm px = {w.gdk, d.Bp, w.wb_dh, rp(H), B, H, 2 * Dp};
        VAE_TRY(vae_product(d, px, nullptr, 0, EPI_ACT_GRAD, nullptr, w.da[n - 1], w.gk_d[n - 1], d.Bp, kp(H), w.gt_d[n - 1], rp(H), d.Bk, stream));
    }
    for (int i = n - 1; i >= 0; --i) {
        const int64_t in = i == 0 ? L : H;
        const VaeGemm pw = {i == 0 ? w.zt : w.dt[i - 1], rp(in + 1), w.gt_d[i], rp(H), in + 1, H, B};
        VAE_TRY(vae_dw(d, pw, w.partials, 1.0f, grad + o.wd[i], H, nullptr, 0, 0, side_after_main()));
        const VaeGemm px = {w.gk_d[i], d.Bp, w.wb_d[i], rp(in), B, in, H};
        if (i > 0)
            VAE_TRY(vae_product(d, px, nullptr, 0, EPI_ACT_GRAD, nullptr, w.da[i - 1], w.gk_d[i - 1], d.Bp, kp(H), w.gt_d[i - 1], rp(H), d.Bk, stream));
        else
            VAE_TRY(vae_product(d, px, w.dz, L, EPI_SCALE, nullptr, nullptr, nullptr, 0, 0, nullptr, 0, 0, stream));
    }

    // ---- 5. latent backward, encoder backward ----------------------------------------------------------------------------------
    MQ_LAUNCH(vae_latent_bwd_kernel, (unsigned)mq_ceil_div(B * L, MQ_NT), MQ_NT, 0, stream, (const float *)w.dz, (const float *)w.mu, (const float *)w.ls,
              (const float *)w.ls_raw, (const float *)w.noise, B * L, (int)L, (int)Lp, lo, hi, inv_b, w.dzh);
    {
        GmPackJob jobs[2];
        jobs[0] = vae_job(w.dzh, 2 * Lp, 1, B, 2 * Lp, 0, w.ghk, d.Bp, kp(2 * Lp), 0, 0, true);
        jobs[1] = vae_job(w.dzh, 1, 2 * Lp, 2 * Lp, B, 0, w.ght, rp(2 * Lp), d.Bk, 0, 0, true);
        VAE_TRY(mq_gemm_tma_pack_multi(jobs, 2, d.mode, stream));
    }
    {
        const VaeGemm pw = {w.et[n - 1], rp(H + 1), w.ght, rp(2 * Lp), H + 1, 2 * Lp, B};
        VAE_TRY(vae_dw(d, pw, w.partials, 1.0f, grad + o.w_mu, L, grad + o.w_ls, Lp, L, side_after_main()));
        const VaeGemm px = {w.ghk, d.Bp, w.wb_eh, rp(H), B, H, 2 * Lp};
        VAE_TRY(vae_product(d, px, nullptr, 0, EPI_ACT_GRAD, nullptr, w.ea[n - 1], n > 1 ? w.gk_e[n - 1] : nullptr, d.Bp, kp(H), w.gt_e[n - 1], rp(H), d.Bk,
                            stream));
    }
    for (int i = n - 1; i >= 0; --i) {
        const int64_t in = i == 0 ? D : H;
        const VaeGemm pw = {i == 0 ? w.xt : w.et[i - 1], rp(in + 1), w.gt_e[i], rp(H), in + 1, H, B};
        VAE_TRY(vae_dw(d, pw, w.partials, 1.0f, grad + o.we[i], H, nullptr, 0, 0, side_after_main()));
        if (i > 0) {
            const VaeGemm px = {w.gk_e[i], d.Bp, w.wb_e[i], rp(H), B, H, H};
            VAE_TRY(vae_product(d, px, nullptr, 0, EPI_ACT_GRAD, nullptr, w.ea[i - 1], i > 1 ? w.gk_e[i - 1] : nullptr, d.Bp, kp(H), w.gt_e[i - 1], rp(H), d.Bk,
                                stream));
        }
    }
    if (n_forks > 0) {                   // join: the caller's stream continues after the last weight gradient
        cudaEventRecord(g_async.ev[n_forks], g_async.side);
        cudaStreamWaitEvent((cudaStream_t)stream, g_async.ev[n_forks], 0);
    }
    MQ_CHECK_LAUNCH("mq_vae_grad");
    return MQ_OK;
}

// One training step of vae.py:57-71: the gradient above, then mom = clip(decay * mom + (1 - decay) * grad, +-clip),
// theta += lr * mom (mq_momentum_step).  `grad` is scratch of n_params floats.
extern "C" int mq_vae_step(const mq_vae_t *v, float *theta, float *mom, const float *x, const float *unit, float *grad, float *logprob,
                           float *dkl, float decay, float clip, float lr, void *ws, int64_t ws_bytes, void *stream) {
    VAE_TRY(mq_vae_grad(v, theta, x, unit, grad, logprob, dkl, ws, ws_bytes, stream));
    MQ_REQUIRE(mom, MQ_ERR_INVALID, "mq_vae_step: null momentum");
    return mq_momentum_step(theta, mom, grad, nullptr, 0, mq_vae_n_params(v), decay, clip, lr, stream);
}

#else
extern "C" int64_t mq_vae_n_params(const mq_vae_t *) { return -1; }
extern "C" int64_t mq_vae_ws_bytes(const mq_vae_t *) { return -1; }
extern "C" int mq_vae_ws_init(const mq_vae_t *, void *, int64_t, void *) { return MQ_ERR_UNSUPPORTED; }
extern "C" int mq_vae_grad(const mq_vae_t *, const float *, const float *, const float *, float *, float *, float *, void *, int64_t, void *) { return MQ_ERR_UNSUPPORTED; }
extern "C" int mq_vae_step(const mq_vae_t *, float *, float *, const float *, const float *, float *, float *, float *, float, float, float, void *, int64_t, void *) { return MQ_ERR_UNSUPPORTED; }
#endif
