// mnist/mlp.py:24-43 at the reference's own batch size (128) as ONE persistent launch per block of optimiser steps.
//
// Reference sequence per step (mnist/mlp.py:31-35): draw 128 rows by index, Function.evaluate (mannequin/basicnet.py:192-207)
// through Affine(784 -> 128) + LReLU [, Affine(128 -> 128) + LReLU], Affine(128 -> 10), the softmax / L2 head
// (mlp.py:10-14: dy = onehot - softmax(out) - 0.01 out), backprop (basicnet.py:209-257, backprop.py:11-60) with the batch mean
// (basicnet.py:219), Adam.apply_gradient (mannequin/_adam.py:23-53).  The layered path (mlp.py of this package) runs that
// as 11 launches per step; at batch 128 a step is 65 MFLOP and 3 MB, so launch gaps and dependent-kernel latency ARE the
// step time (54 us).  Here a block of steps is one cooperative launch of 128 CTAs that meet at two grid-wide barriers per
// step:
//
//   F0    the first layer is split 8 x 16: CTA (c, q) owns 8 of its 128 output columns and 1/8 of the 784 inputs.  Every
//         CTA keeps ITS slice of W0 (and that slice's Adam moments) in shared memory for the whole launch -- the 401 KB
//         layer never moves -- and multiplies the matching slice of the gathered minibatch (prefetched one step ahead with
//         cp.async) by it; the 128 x 8 partial sums go to L2 (4 KB per CTA).
//         (A thread-block cluster of the 8 K slices with a distributed-shared-memory reduction was the first design: only
//         15 clusters of 8 one-CTA-per-SM blocks are co-resident on a B200, one short of the 16 it needs.)
//   ----- grid barrier
//   MID   CTA g owns batch ROW g: h1 = act(sum of the 8 K-slice partials of row g, in slice order, + b0), h2 = act(h1 W1 +
//         b1), logits, head, dh2, dh1 are all row-local, so the middle of the network (two matrix-vector products forward,
//         two backward, against a shared-memory copy of W1) needs no barrier.
//   ----- grid barrier
//   D     parameter gradients + Adam, partitioned by OWNER: the W0 slice against the minibatch slice the CTA still holds
//         (dW0 = x^T dh1), a 16 x 8 block of W1, row g of the output layer, the biases; the owner applies the optimiser
//         step in place (moments never leave its shared memory) and publishes W1 / W2 / b for the next step's MID.
//
// All sums run in a fixed order: the result is bit-reproducible.  fp32 FFMA arithmetic, fp32 or fp64 optimiser state.
#include "mq_common.cuh"

#ifndef MQ_EMU
#define MF_B 128        // minibatch rows = CTAs
#define MF_H 128        // hidden width
#define MF_G 128        // CTAs of the launch
#define MF_CL 8         // K slices of the first layer
#define MF_NS 8         // first-layer columns per CTA
#define MF_KW 100       // widest K slice (multiple of 4)
#define MF_NO 16        // most outputs
#define MF_NT 256

struct MfArgs {
    mq_net_t net;
    float *theta;
    void *value, *m, *v;
    const float *x, *y;
    const int32_t *idx;             // [steps][128]
    const double *coefs;            // [steps][4]: c1, c2, lr, eps
    int steps, kw;
    float l2;
    float *pg;                      // [8][128][128] K-slice partial sums of the first layer
    float *h1, *h2, *dy, *dh2, *dh1;
    unsigned int *bar;
    long long *prof;                // developer aid (mq_debug_set_phase_buffer): time stamps of the last step, CTAs 0 and 127
};

template <typename ST>
__device__ __forceinline__ void mf_adam(ST &val, ST &m, ST &v, float g, ST c1, ST c2, ST lr, ST eps);
template <>
__device__ __forceinline__ void mf_adam<float>(float &val, float &m, float &v, float g, float c1, float c2, float lr, float eps) {
    m += c1 * (g - m);                                   // the arithmetic of adam_kernel (mq_adam.cu), fp32 state
    v += c2 * (g * g - v);
    val += lr * (m / (eps + sqrtf(v)));
}
template <>
__device__ __forceinline__ void mf_adam<double>(double &val, double &m, double &v, float g32, double c1, double c2, double lr,
                                                double eps) {
    const double g = (double)g32;                        // every rounding spelled out: NumPy's sequence of ufunc calls
    m = __dadd_rn(m, __dmul_rn(c1, __dsub_rn(g, m)));
    v = __dadd_rn(v, __dmul_rn(c2, __dsub_rn(__dmul_rn(g, g), v)));
    val = __dadd_rn(val, __dmul_rn(lr, __ddiv_rn(m, __dadd_rn(eps, __dsqrt_rn(v)))));
}

__device__ __forceinline__ void mf_cp16(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void mf_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void mf_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// bulk asynchronous copies (TMA engine): one thread issues, completion is counted in bytes on an mbarrier -- no per-thread
// load/store-unit slots are held while the data is in flight (W1: one 64 KB copy instead of 4,096 cp.async of 16 bytes)
__device__ __forceinline__ uint32_t mf_s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mf_mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mf_s32(b)), "r"(count));
}
__device__ __forceinline__ void mf_mbar_expect(uint64_t *b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mf_s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mf_bulk(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(mf_s32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(mf_s32(b)) : "memory");
}
__device__ __forceinline__ void mf_mbar_wait(uint64_t *b, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(mf_s32(b)), "r"(parity) : "memory");
    }
}
// loads that must be ISSUED where they are written (ptxas otherwise sinks them next to their first use, a phase later)
__device__ __forceinline__ int32_t mf_ldg_s32(const int32_t *p) { int32_t v; asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ float mf_ldg_f32(const float *p) { float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }

// all CTAs of the launch have arrived (and their global writes are visible) -- monotonic counter, target = arrivals so far
__device__ __forceinline__ void mf_grid_barrier(unsigned int *bar, unsigned int &epoch) {
    __syncthreads();
    epoch += MF_G;
    if (threadIdx.x == 0) {
        // arrive without waiting for the old value (one L2 round trip less than atomicAdd); release is cumulative over the
        // writes of the whole CTA ordered before it by the barrier above
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
        } while (seen < epoch);
    }
    __syncthreads();
}

__device__ __forceinline__ long long mf_now() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define MF_MARK(i) do { if (a.prof && tid == 0 && (g == 0 || g == MF_G - 1) && s == a.steps / 2) a.prof[(g ? 16 : 0) + (i)] = mf_now(); } while (0)

template <typename ST>
__global__ void __launch_bounds__(MF_NT, 1)
mlp_block_kernel(const __grid_constant__ MfArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x;                           // CTA = batch row in MID
    const int c = g / MF_CL, q = g % MF_CL;             // column slice and K slice of the first layer
    const mq_net_t &net = a.net;
    const int L = net.n_layers, n_in = net.n_in, n_out = net.n_out;
    const int kw = a.kw;
    const int k0 = q * kw;
    const int kn = n_in - k0 < kw ? (n_in - k0 > 0 ? n_in - k0 : 0) : kw;      // this slice's width (multiple of 4)
    const int kb = g / 16, jb = g % 16;                 // the 16 x 8 block of W1 this CTA owns (L == 3)
    const int lastl = L - 1;
    const float inv_b = 1.0f / (float)MF_B;

    // ---- shared memory carve-up ------------------------------------------------------------------------------
    unsigned char *sp = smem_raw;
    auto take = [&](size_t bytes) { unsigned char *r = sp; sp += (bytes + 15) & ~(size_t)15; return r; };
    float *xs = (float *)take(2 * MF_B * MF_KW * 4);                 // [2][128][100] gathered minibatch slice
    float *w0f = (float *)take(MF_KW * MF_NS * 4);                   // [kw][8] compute copy of the W0 slice
    float *W1s = (float *)take(MF_H * MF_H * 4);                     // [128][128] (L == 3)
    float *W2s = (float *)take(MF_H * MF_NO * 4);                    // [128][n_out]
    float *b1s = (float *)take(MF_H * 4);
    float *b2s = (float *)take(MF_NO * 4);
    float *h1r = (float *)take(MF_H * 4);
    float *h2r = (float *)take(MF_H * 4);
    float *dh2r = (float *)take(MF_H * 4);
    float *tmp = (float *)take(MF_NT * 4);
    float *part8 = (float *)take(8 * MF_H * 4);                       // [8][128] partial h2 sums of the 8 k groups
    float *lg = (float *)take(MF_NO * 4);
    float *dyr = (float *)take(MF_NO * 4);
    // operands of phase D live in the W1 buffer (MID is done with it by then, the next MID reloads it)
    float *d1s = W1s;                                                // [128][8] dh1 columns of this cluster
    float *hk = d1s + MF_B * MF_NS;                                  // [128][16] h1 columns of the W1 block
    float *dj = hk + MF_B * 16;                                      // [128][8] dh2 columns of the W1 block
    float *dys = dj + MF_B * MF_NS;                                  // [128][16] dy
    float *hcol = dys + MF_B * MF_NO;                                // column g of the last hidden layer
    float *c1s = hcol + MF_B, *c2s = c1s + MF_B;                     // column g of dh1 / dh2 (bias gradients)
    // optimiser state of the parameters this CTA owns: value, m, v
    ST *sW0 = (ST *)take(3 * MF_KW * MF_NS * sizeof(ST));            // [3][kw * 8]
    ST *sB0 = (ST *)take(3 * sizeof(ST) * 2);                        // [3]: b0[g]
    ST *sW1 = (ST *)take(3 * 128 * sizeof(ST));                      // [3][16 * 8]
    ST *sWl = (ST *)take(3 * MF_NO * sizeof(ST));                    // [3][n_out]: row g of the output layer
    ST *sB1 = (ST *)take(3 * sizeof(ST) * 2);                        // [3]: b1[g] (L == 3)
    ST *sBl = (ST *)take(3 * MF_NO * sizeof(ST));                    // [3][n_out]: output bias (CTA 0)

    ST *gval = (ST *)a.value, *gm = (ST *)a.m, *gv = (ST *)a.v;
    const int w0o = net.w_off[0], b0o = net.b_off[0];
    const int w1o = L == 3 ? net.w_off[1] : 0, b1o = L == 3 ? net.b_off[1] : 0;
    const int wlo = net.w_off[lastl], blo = net.b_off[lastl];
    const int act0 = net.act[0];
    const float leak0 = net.leak[0];
    const int act1 = L == 3 ? net.act[1] : MQ_ACT_NONE;
    const float leak1 = L == 3 ? net.leak[1] : 0.0f;

    // ---- load the owned parameters and their moments ----------------------------------------------------------------
    for (int i = tid; i < kn * MF_NS; i += MF_NT) {
        const int k = i / MF_NS, n = i % MF_NS;
        const int p = w0o + (k0 + k) * MF_H + c * MF_NS + n;
        sW0[i] = gval[p]; sW0[MF_KW * MF_NS + i] = gm[p]; sW0[2 * MF_KW * MF_NS + i] = gv[p];
        w0f[i] = (float)gval[p];
    }
    if (tid == 0) { const int p = b0o + g; sB0[0] = gval[p]; sB0[1] = gm[p]; sB0[2] = gv[p]; }
    if (L == 3 && tid < 128) {
        const int p = w1o + (16 * kb + tid / 8) * MF_H + 8 * jb + tid % 8;
        sW1[tid] = gval[p]; sW1[128 + tid] = gm[p]; sW1[256 + tid] = gv[p];
        if (tid == 0) { const int pb = b1o + g; sB1[0] = gval[pb]; sB1[1] = gm[pb]; sB1[2] = gv[pb]; }
    }
    if (tid < n_out) {
        const int p = wlo + g * n_out + tid;
        sWl[tid] = gval[p]; sWl[MF_NO + tid] = gm[p]; sWl[2 * MF_NO + tid] = gv[p];
        if (g == 0) { const int pb = blo + tid; sBl[tid] = gval[pb]; sBl[MF_NO + tid] = gm[pb]; sBl[2 * MF_NO + tid] = gv[pb]; }
    }

    // minibatch slice of step s -> xs[s & 1] (16-byte asynchronous copies; rows are 16-byte aligned: n_in % 4 == 0)
    // minibatch slice of step s -> xs[s & 1]: thread (row, half) copies <= 13 chunks of 16 bytes with cp.async (rows are
    // 16-byte aligned: n_in % 4 == 0); the row's gather index was loaded a phase earlier.  (One bulk copy per row on the TMA
    // engine was measured slower: 128 copies of 400 bytes per CTA queue up behind each other.)
    __shared__ __align__(8) uint64_t bar_w;
    if (tid == 0) {
        mf_mbar_init(&bar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    auto prefetch_x = [&](int s, int32_t src_row) {
        const int r = tid >> 1, half = tid & 1;
        const int cpr = kn >> 2;                        // 16-byte chunks per row
        const int ch0 = half * ((cpr + 1) >> 1), ch1 = half ? cpr : ((cpr + 1) >> 1);
        float *dst = xs + (s & 1) * MF_B * MF_KW + r * MF_KW;
        const float *src = a.x + (int64_t)src_row * n_in + k0;
        for (int ch = ch0; ch < ch1; ++ch) mf_cp16(dst + 4 * ch, src + 4 * ch);
        mf_cp_commit();
    };
    prefetch_x(0, mf_ldg_s32(a.idx + (tid >> 1)));
    int32_t next_row = a.steps > 1 ? mf_ldg_s32(a.idx + MF_B + (tid >> 1)) : 0;     // gather index of this thread's row, next step
    uint32_t par_w = 0u;
    unsigned int epoch = 0;

    for (int s = 0; s < a.steps; ++s) {
        const ST c1 = (ST)a.coefs[4 * s], c2 = (ST)a.coefs[4 * s + 1], lr = (ST)a.coefs[4 * s + 2], eps = (ST)a.coefs[4 * s + 3];
        const float *xcur = xs + (s & 1) * MF_B * MF_KW;
        float *h1g = a.h1;
        MF_MARK(0);
        if (a.prof && tid == 0 && g == 0 && s == 0) a.prof[14] = mf_now();
        if (s + 1 < a.steps) { prefetch_x(s + 1, next_row); mf_cp_wait<1>(); } else { mf_cp_wait<0>(); }
        // label entry of row g (lane = output) for the head, and the gather index of the step after the next: both far ahead of use
        float ylab = 0.0f;
        if (warp == 0 && lane < n_out) ylab = mf_ldg_f32(a.y + (int64_t)mf_ldg_s32(a.idx + (int64_t)s * MF_B + g) * n_out + lane);
        if (s + 2 < a.steps) next_row = mf_ldg_s32(a.idx + (int64_t)(s + 2) * MF_B + (tid >> 1));
        __syncthreads();
        MF_MARK(1);

        // ---- F0: partial[r][n] = sum_k x[r][k0 + k] W0[k0 + k][8 c + n] ------------------------------------------------
        {
            const int r = tid >> 1, nh = (tid & 1) * 4;
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
            const float *xr = xcur + r * MF_KW;
            for (int k = 0; k < kn; k += 4) {
                const float4 xv = *reinterpret_cast<const float4 *>(xr + k);
                const float xe[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float4 w = *reinterpret_cast<const float4 *>(w0f + (k + e) * MF_NS + nh);
                    acc[0] = fmaf(xe[e], w.x, acc[0]); acc[1] = fmaf(xe[e], w.y, acc[1]);
                    acc[2] = fmaf(xe[e], w.z, acc[2]); acc[3] = fmaf(xe[e], w.w, acc[3]);
                }
            }
            *reinterpret_cast<float4 *>(a.pg + ((int64_t)q * MF_B + r) * MF_H + c * MF_NS + nh) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
        MF_MARK(2);
        mf_grid_barrier(a.bar, epoch);
        MF_MARK(3);

        // ---- MID: row g through the rest of the net and back (row-local) ---------------------------------------------
        if (L == 3 && tid == 0) {                       // W1 (64 KB, published by its owners before the barrier): one bulk copy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // (the buffer was read / written by ordinary accesses)
            mf_mbar_expect(&bar_w, (uint32_t)(MF_H * MF_H * 4));
            mf_bulk(W1s, a.theta + w1o, (uint32_t)(MF_H * MF_H * 4), &bar_w);
        }
        if (tid < MF_H) {                                // the 8 K-slice partials of row g, in slice order
            float sum = 0.0f;
#pragma unroll
            for (int qq = 0; qq < MF_CL; ++qq) sum += __ldcg(a.pg + ((int64_t)qq * MF_B + g) * MF_H + tid);
            const float h = mq_act_apply(sum + __ldcg(a.theta + b0o + tid), act0, leak0);
            h1r[tid] = h;
            h1g[g * MF_H + tid] = h;
            if (L == 3) b1s[tid] = __ldcg(a.theta + b1o + tid);
        }
        for (int i = tid; i < MF_H * n_out; i += MF_NT) W2s[i] = __ldcg(a.theta + wlo + i);
        if (tid < n_out) b2s[tid] = __ldcg(a.theta + blo + tid);
        if (L == 3) { mf_mbar_wait(&bar_w, par_w); par_w ^= 1u; }
        __syncthreads();
        MF_MARK(7);
        const float *hl = h1r;                           // last hidden activations of this row
        if (L == 3) {
            // h2 = act(h1 W1 + b1): thread (chunk j4 of 4 outputs, group of 16 k) -- a warp reads the 32 chunks of a row
            const int j4 = tid & 31, kg = tid >> 5;
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k = kg * 16 + i;
                const float4 w = *reinterpret_cast<const float4 *>(W1s + k * MF_H + 4 * j4);
                const float hv = h1r[k];
                acc[0] = fmaf(hv, w.x, acc[0]); acc[1] = fmaf(hv, w.y, acc[1]);
                acc[2] = fmaf(hv, w.z, acc[2]); acc[3] = fmaf(hv, w.w, acc[3]);
            }
            *reinterpret_cast<float4 *>(part8 + kg * MF_H + 4 * j4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            __syncthreads();
            if (tid < MF_H) {
                float sum = 0.0f;
#pragma unroll
                for (int q8 = 0; q8 < 8; ++q8) sum += part8[q8 * MF_H + tid];
                h2r[tid] = mq_act_apply(sum + b1s[tid], act1, leak1);
            }
            __syncthreads();
            hl = h2r;
        }
        MF_MARK(8);
        for (int o = warp; o < n_out; o += MF_NT / 32) {  // logits: one warp per output
            float acc = 0.0f;
#pragma unroll
            for (int j = lane; j < MF_H; j += 32) acc = fmaf(hl[j], W2s[j * n_out + o], acc);
            acc = mq_warp_sum(acc);
            if (lane == 0) lg[o] = acc + b2s[o];
        }
        __syncthreads();
        MF_MARK(9);
        if (warp == 0) {                                 // mnist/mlp.py:10-14: dy = onehot - softmax(out) - l2 out (lane = output)
            const bool on = lane < n_out;
            const float z = on ? lg[lane] : -INFINITY;
            float mx = z;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float e = on ? expf(z - mx) : 0.0f;
            const float sum = mq_warp_sum(e);
            if (on) {
                const float d = ylab - e * (1.0f / sum) - a.l2 * z;
                dyr[lane] = d;
                a.dy[g * MF_NO + lane] = d;
            }
        }
        __syncthreads();
        MF_MARK(10);
        if (tid < MF_H) {                                // gradient at the last hidden layer's pre-activation
            float acc = 0.0f;
            for (int o = 0; o < n_out; ++o) acc = fmaf(dyr[o], W2s[tid * n_out + o], acc);
            const float dz = mq_act_grad(hl[tid], acc, L == 3 ? act1 : act0, L == 3 ? leak1 : leak0);
            if (L == 3) { dh2r[tid] = dz; a.dh2[g * MF_H + tid] = dz; a.h2[g * MF_H + tid] = h2r[tid]; }
            else a.dh1[g * MF_H + tid] = dz;
        }
        if (L == 3) {
            __syncthreads();
            MF_MARK(11);
            {   // dh1[k] = act'(h1[k]) sum_j dh2[j] W1[k][j]: thread (k, half) walks half of row k in 16-byte chunks, starting at
                // chunk k (rows are 512 bytes apart: eight consecutive k starting at the same chunk would share a bank group)
                const int k = tid & 127, half = tid >> 7;
                float acc = 0.0f;
#pragma unroll 4
                for (int i = half * 16; i < half * 16 + 16; ++i) {
                    const int j4 = (i + k) & 31;
                    const float4 w = *reinterpret_cast<const float4 *>(W1s + k * MF_H + 4 * j4);
                    const float4 d = *reinterpret_cast<const float4 *>(dh2r + 4 * j4);
                    acc = fmaf(d.x, w.x, acc); acc = fmaf(d.y, w.y, acc); acc = fmaf(d.z, w.z, acc); acc = fmaf(d.w, w.w, acc);
                }
                tmp[tid] = acc;
                __syncthreads();
                if (tid < MF_H) a.dh1[g * MF_H + tid] = mq_act_grad(h1r[tid], tmp[tid] + tmp[tid + 128], act0, leak0);
            }
        }
        MF_MARK(4);
        mf_grid_barrier(a.bar, epoch);
        MF_MARK(5);

        // ---- D: parameter gradients (batch mean, basicnet.py:219) + Adam, by owner ------------------------------------
        {   // operands of the three products
            const float4 v4 = __ldcg(reinterpret_cast<const float4 *>(a.dh1 + (tid >> 1) * MF_H + c * MF_NS + (tid & 1) * 4));
            *reinterpret_cast<float4 *>(d1s + (tid >> 1) * MF_NS + (tid & 1) * 4) = v4;
            if (L == 3) {
                for (int i = tid; i < MF_B * 4; i += MF_NT) {
                    const int r = i >> 2, f = (i & 3) * 4;
                    *reinterpret_cast<float4 *>(hk + r * 16 + f) = __ldcg(reinterpret_cast<const float4 *>(h1g + r * MF_H + 16 * kb + f));
                }
                const float4 u4 = __ldcg(reinterpret_cast<const float4 *>(a.dh2 + (tid >> 1) * MF_H + 8 * jb + (tid & 1) * 4));
                *reinterpret_cast<float4 *>(dj + (tid >> 1) * MF_NS + (tid & 1) * 4) = u4;
            }
            for (int i = tid; i < MF_B * MF_NO / 4; i += MF_NT)
                *reinterpret_cast<float4 *>(dys + 4 * i) = __ldcg(reinterpret_cast<const float4 *>(a.dy + 4 * i));
            if (tid < MF_B) {
                hcol[tid] = __ldcg((L == 3 ? a.h2 : h1g) + tid * MF_H + g);
                c1s[tid] = __ldcg(a.dh1 + tid * MF_H + g);
                if (L == 3) c2s[tid] = __ldcg(a.dh2 + tid * MF_H + g);
            }
        }
        __syncthreads();
        MF_MARK(12);
        if (tid < 2 * kn) {                              // dW0[k][4 nh ..] = (1/B) sum_r x[r][k] dh1[r][..]
            const int k = tid >> 1, nh = (tid & 1) * 4;
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll 8
            for (int r = 0; r < MF_B; ++r) {
                const float xv = xcur[r * MF_KW + k];
                const float4 d = *reinterpret_cast<const float4 *>(d1s + r * MF_NS + nh);
                acc[0] = fmaf(xv, d.x, acc[0]); acc[1] = fmaf(xv, d.y, acc[1]);
                acc[2] = fmaf(xv, d.z, acc[2]); acc[3] = fmaf(xv, d.w, acc[3]);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = k * MF_NS + nh + e;
                ST val = sW0[i], mm = sW0[MF_KW * MF_NS + i], vv = sW0[2 * MF_KW * MF_NS + i];
                mf_adam<ST>(val, mm, vv, acc[e] * inv_b, c1, c2, lr, eps);
                sW0[i] = val; sW0[MF_KW * MF_NS + i] = mm; sW0[2 * MF_KW * MF_NS + i] = vv;
                w0f[i] = (float)val;
            }
        }
        MF_MARK(13);
        if (L == 3 && tid < 128) {                       // the 16 x 8 block of W1: dW1[k][j] = (1/B) sum_r h1[r][k] dh2[r][j]
            const int kk = tid >> 3, jj = tid & 7;
            float acc = 0.0f;
#pragma unroll 8
            for (int r = 0; r < MF_B; ++r) acc = fmaf(hk[r * 16 + kk], dj[r * MF_NS + jj], acc);
            ST val = sW1[tid], mm = sW1[128 + tid], vv = sW1[256 + tid];
            mf_adam<ST>(val, mm, vv, acc * inv_b, c1, c2, lr, eps);
            sW1[tid] = val; sW1[128 + tid] = mm; sW1[256 + tid] = vv;
            a.theta[w1o + (16 * kb + kk) * MF_H + 8 * jb + jj] = (float)val;
        } else if (tid >= 128 && tid < 128 + n_out) {    // row g of the output layer: dW[g][o] = (1/B) sum_r hl[r][g] dy[r][o]
            const int o = tid - 128;
            float acc = 0.0f;
#pragma unroll 8
            for (int r = 0; r < MF_B; ++r) acc = fmaf(hcol[r], dys[r * MF_NO + o], acc);
            ST val = sWl[o], mm = sWl[MF_NO + o], vv = sWl[2 * MF_NO + o];
            mf_adam<ST>(val, mm, vv, acc * inv_b, c1, c2, lr, eps);
            sWl[o] = val; sWl[MF_NO + o] = mm; sWl[2 * MF_NO + o] = vv;
            a.theta[wlo + g * n_out + o] = (float)val;
        } else if (tid >= 160 && tid < 192) {            // b1[g] = (1/B) sum_r dh2[r][g]  (one warp)
            if (L == 3) {
                float acc = 0.0f;
                for (int r = lane; r < MF_B; r += 32) acc += c2s[r];
                acc = mq_warp_sum(acc);
                if (lane == 0) {
                    ST val = sB1[0], mm = sB1[1], vv = sB1[2];
                    mf_adam<ST>(val, mm, vv, acc * inv_b, c1, c2, lr, eps);
                    sB1[0] = val; sB1[1] = mm; sB1[2] = vv;
                    a.theta[b1o + g] = (float)val;
                }
            }
        } else if (tid >= 192 && tid < 224) {            // b0[g] = (1/B) sum_r dh1[r][g]  (one warp)
            float acc = 0.0f;
            for (int r = lane; r < MF_B; r += 32) acc += c1s[r];
            acc = mq_warp_sum(acc);
            if (lane == 0) {
                ST val = sB0[0], mm = sB0[1], vv = sB0[2];
                mf_adam<ST>(val, mm, vv, acc * inv_b, c1, c2, lr, eps);
                sB0[0] = val; sB0[1] = mm; sB0[2] = vv;
                a.theta[b0o + g] = (float)val;
            }
        } else if (g == 0 && tid >= 224 && tid < 224 + n_out) {      // output bias: db[o] = (1/B) sum_r dy[r][o]
            const int o = tid - 224;
            float acc = 0.0f;
            for (int r = 0; r < MF_B; ++r) acc += dys[r * MF_NO + o];
            ST val = sBl[o], mm = sBl[MF_NO + o], vv = sBl[2 * MF_NO + o];
            mf_adam<ST>(val, mm, vv, acc * inv_b, c1, c2, lr, eps);
            sBl[o] = val; sBl[MF_NO + o] = mm; sBl[2 * MF_NO + o] = vv;
            a.theta[blo + o] = (float)val;
        }
        __syncthreads();
        MF_MARK(6);
        if (a.prof && tid == 0 && g == 0 && s == a.steps - 1) a.prof[15] = mf_now();
    }

    // ---- write the owned parameters and moments back -------------------------------------------------------------------
    for (int i = tid; i < kn * MF_NS; i += MF_NT) {
        const int k = i / MF_NS, n = i % MF_NS;
        const int p = w0o + (k0 + k) * MF_H + c * MF_NS + n;
        gval[p] = sW0[i]; gm[p] = sW0[MF_KW * MF_NS + i]; gv[p] = sW0[2 * MF_KW * MF_NS + i];
        a.theta[p] = (float)sW0[i];
    }
    if (tid == 0) { const int p = b0o + g; gval[p] = sB0[0]; gm[p] = sB0[1]; gv[p] = sB0[2]; }
    if (L == 3 && tid < 128) {
        const int p = w1o + (16 * kb + tid / 8) * MF_H + 8 * jb + tid % 8;
        gval[p] = sW1[tid]; gm[p] = sW1[128 + tid]; gv[p] = sW1[256 + tid];
        if (tid == 0) { const int pb = b1o + g; gval[pb] = sB1[0]; gm[pb] = sB1[1]; gv[pb] = sB1[2]; }
    }
    if (tid < n_out) {
        const int p = wlo + g * n_out + tid;
        gval[p] = sWl[tid]; gm[p] = sWl[MF_NO + tid]; gv[p] = sWl[2 * MF_NO + tid];
        if (g == 0) { const int pb = blo + tid; gval[pb] = sBl[tid]; gm[pb] = sBl[MF_NO + tid]; gv[pb] = sBl[2 * MF_NO + tid]; }
    }
}

long long *mq_phase_prof_ptr();

static size_t mf_smem_bytes(size_t st) {
    auto r16 = [](size_t b) { return (b + 15) & ~(size_t)15; };
    size_t b = 0;
    b += r16(2 * MF_B * MF_KW * 4) + r16(MF_KW * MF_NS * 4) + r16(MF_H * MF_H * 4) + r16(MF_H * MF_NO * 4);
    b += r16(MF_H * 4) + r16(MF_NO * 4) + 3 * r16(MF_H * 4) + r16(MF_NT * 4) + r16(8 * MF_H * 4) + 2 * r16(MF_NO * 4);
    b += r16(3 * MF_KW * MF_NS * st) + r16(3 * st * 2) + r16(3 * 128 * st) + r16(3 * MF_NO * st) + r16(3 * st * 2) + r16(3 * MF_NO * st);
    return b;
}

static bool mf_fits(const mq_net_t *net, int64_t batch) {
    if (!net || batch != MF_B) return false;
    const int L = net->n_layers;
    if (L != 2 && L != 3) return false;
    if (net->logstd_off >= 0 || net->n_out > MF_NO || net->n_out < 1) return false;
    if (net->n_in % 4 != 0 || net->n_in < 4 || net->n_in > MF_CL * MF_KW) return false;
    for (int l = 0; l + 1 < L; ++l)
        if (net->out[l] != MF_H) return false;
    if (net->act[L - 1] != MQ_ACT_NONE) return false;
    return true;
}

extern "C" int64_t mq_mlp_train_block_ws_bytes(const mq_net_t *net, int64_t batch) {
    if (!mf_fits(net, batch)) return 0;
    // first-layer partials, h1, h2, dh2, dh1, dy, barrier word
    return (int64_t)((MF_CL + 4) * MF_B * MF_H + MF_B * MF_NO) * 4 + 256;
}

// A block of `steps` optimiser steps of mnist/mlp.py:31-35 (batch 128 drawn through idx_table, softmax / L2 head,
// Adam) as ONE launch.  theta: fp32 model copy; value / m / v: the optimiser state (fp32 or fp64, as mq_adam_step);
// coefs_dev: [steps][4] doubles {c1, c2, lr, eps} (the host's RunningMean decay schedule, _adam.py:14-21).
// MQ_ERR_UNSUPPORTED when the net or the batch does not fit (anything but batch 128, hidden width 128, two or three layers).
extern "C" int mq_mlp_train_block(const mq_net_t *net, float *theta, void *value, void *m, void *v, int state_dtype,
                                  const float *x, const float *y_onehot, const int32_t *idx_table, int64_t batch, int32_t steps,
                                  const double *coefs_dev, float l2, void *ws, int64_t ws_bytes, void *stream) {
    if (!mf_fits(net, batch)) return MQ_ERR_UNSUPPORTED;
    MQ_REQUIRE(theta && value && m && v && x && y_onehot && idx_table && coefs_dev && ws, MQ_ERR_INVALID, "mq_mlp_train_block: null pointer");
    MQ_REQUIRE(steps >= 0, MQ_ERR_SHAPE, "mq_mlp_train_block: negative step count");
    MQ_REQUIRE(state_dtype == MQ_F32 || state_dtype == MQ_F64, MQ_ERR_INVALID, "mq_mlp_train_block: bad state dtype");
    MQ_REQUIRE(ws_bytes >= mq_mlp_train_block_ws_bytes(net, batch), MQ_ERR_INVALID, "mq_mlp_train_block: workspace too small");
    MQ_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)ws & 15) == 0 && ((uintptr_t)theta & 15) == 0, MQ_ERR_INVALID,
               "mq_mlp_train_block: x, theta and the workspace must be 16-byte aligned");
    if (steps == 0) return MQ_OK;
    if (net->n_layers == 3) MQ_REQUIRE(net->w_off[1] % 4 == 0, MQ_ERR_UNSUPPORTED, "mq_mlp_train_block: W1 is not 16-byte aligned in theta");
    MfArgs a;
    memset(&a, 0, sizeof(a));
    a.net = *net;
    a.theta = theta; a.value = value; a.m = m; a.v = v;
    a.x = x; a.y = y_onehot; a.idx = idx_table; a.coefs = coefs_dev;
    a.steps = steps; a.l2 = l2;
    a.kw = (int)mq_round_up(mq_ceil_div(net->n_in, MF_CL), 4);
    float *w = (float *)ws;
    a.pg = w; w += MF_CL * MF_B * MF_H;
    a.h1 = w; w += MF_B * MF_H;
    a.h2 = w; w += MF_B * MF_H;
    a.dh2 = w; w += MF_B * MF_H;
    a.dh1 = w; w += MF_B * MF_H;
    a.dy = w; w += MF_B * MF_NO;
    a.bar = (unsigned int *)w;
    a.prof = mq_phase_prof_ptr();
    cudaError_t e = cudaMemsetAsync(a.bar, 0, 256, (cudaStream_t)stream);
    if (e != cudaSuccess) { mq_set_error("mq_mlp_train_block: %s", cudaGetErrorString(e)); return MQ_ERR_CUDA; }
    const size_t smem = mf_smem_bytes(state_dtype == MQ_F64 ? 8 : 4);
    MQ_REQUIRE((int64_t)smem <= mq_smem_optin(), MQ_ERR_UNSUPPORTED, "mq_mlp_train_block: shared memory");
    // all 128 CTAs must be resident at once (they meet at grid-wide barriers): a cooperative launch, which the runtime
    // refuses when they would not be
    void (*kfn)(MfArgs) = state_dtype == MQ_F64 ? mlp_block_kernel<double> : mlp_block_kernel<float>;
    e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(MF_G); cfg.blockDim = dim3(MF_NT); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeCooperative;
        at[0].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kfn, a);
    }
    if (e != cudaSuccess) { mq_set_error("mq_mlp_train_block: %s", cudaGetErrorString(e)); return MQ_ERR_CUDA; }
    MQ_CHECK_LAUNCH("mq_mlp_train_block");
    return MQ_OK;
}

#else   // MQ_EMU: no clusters in the host emulation
extern "C" int64_t mq_mlp_train_block_ws_bytes(const mq_net_t *, int64_t) { return 0; }
extern "C" int mq_mlp_train_block(const mq_net_t *, float *, void *, void *, void *, int, const float *, const float *, const int32_t *,
                                  int64_t, int32_t, const double *, float, void *, int64_t, void *) { return MQ_ERR_UNSUPPORTED; }
#endif
