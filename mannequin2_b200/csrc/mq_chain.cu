// Persistent optimiser chain ("row-owner warps"): n_steps dependent
//   {gather minibatch, forward, loss-gradient head, backward, Adam}
// steps of a small MLP in ONE launch of ONE CTA.  This is the latency-critical kernel of
// benchmarks/ppo.py:29-40 (300 policy steps) and benchmarks/gae.py:71-73 (300 value steps):
// every step depends on the previous one and is only ~2 MFLOP, so the design goal is few
// barriers and no idle lanes rather than FLOP/s.
//
//   * a 64-row tile; warp w owns rows 8w..8w+7 through forward, head and the
//     back-propagation of the signal (dX): rows are independent there, so those phases need
//     only __syncwarp -- the CTA synchronises 3 times per step instead of ~15;
//   * wide layers: lane <-> output column (lane, lane+32, ...), 8 rows x CPL columns of FFMA
//     accumulators per lane, the 8 row values are one broadcast 2x128-bit load;
//     narrow layers (<= 8 outputs, e.g. the 3 action means): lane <-> (row, k-slice), the
//     k-slices are combined with warp shuffles;
//   * dW/db: the whole CTA, 4x4 (or 1x4 for small layers) register tiles over (in,out),
//     contraction over the 64 rows with unit-stride 128-bit loads; the logstd gradient is a
//     warp-shuffle row sum.  Fixed summation order => bit-reproducible;
//   * Adam (fp32 moments) lives in shared memory and is merged with the refresh of the
//     transposed weight copies; weights never leave the SM during the launch;
//   * the gathers of step s+1 (Trajectory indexing, mannequin/_trajectory.py:79-86) are
//     issued into registers during step s; their index rows are fetched two steps ahead.
//
// Shared memory (floats), natural layouts: W_l [in][npad] | b_l [npad] | logstd, G same,
// M, V same; Wt_l [out][kpad+1] for l >= 1; activations A_l [width][68]; dZ_l [out_l][68]; aux: target/hterm [n_out][68] x 2, misc [2][68]; idx [2][64].
#include "mq_common.cuh"

// Developer aid: per-phase cycle counters (buffer is NULL in normal use).
#ifndef MQ_EMU
#define CH_MARK(i) do { if (prof && threadIdx.x == 0) { long long now_ = clock64(); prof[i] += now_ - t_last; t_last = now_; } } while (0)
#else
#define CH_MARK(i) do { } while (0)
#endif
long long *mq_phase_prof_ptr();

// A configuration is encoded in ONE template integer CF = NW + 256 * RW:
//   NW warps per CTA (NT threads), RW rows per warp, a tile of TB = RW*NW rows,
//   padded row stride TBP (= 4 mod 32).
#define CH_CF(nw, rw) ((nw) + 256 * (rw))
template <int CF> struct ChainCfg {
    static constexpr int NWARPS = CF & 255;
    static constexpr int RW = CF >> 8;
    static constexpr int NT = 32 * NWARPS;
    static constexpr int TB = RW * NWARPS;
    static constexpr int TBP = TB + 4;
};
#define CH_PFX 8                      // prefetch registers per thread: x elements
#define CH_PFH 2                      //   head matrix elements
#define CH_HALF_LOG_2PI 0.91893853320467274178f

struct ChainPlan {
    mq_net_t net;
    int nw, rw, with_state;
    int npad[MQ_MAX_LAYERS], ldwt[MQ_MAX_LAYERS];
    int sw_off[MQ_MAX_LAYERS], sb_off[MQ_MAX_LAYERS], swt_off[MQ_MAX_LAYERS];
    int sls_off, p_pad, wt_total;
    int act_off[MQ_MAX_LAYERS + 1], act_total, dz_off[MQ_MAX_LAYERS], dz_total, aux_off;
    int64_t smem_bytes;
};

static bool chain_plan(const mq_net_t *net, ChainPlan *p, int nw, int rw, int with_state) {
    memset(p, 0, sizeof(*p));
    p->net = *net;
    p->nw = nw;
    p->rw = rw;
    p->with_state = with_state;
    const int CH_NT = 32 * nw, CH_TB = rw * nw, CH_TBP = CH_TB + 4;
    int pos = 0, wt = 0, maxw = 4;
    for (int l = 0; l < net->n_layers; ++l) {
        if (net->out[l] > 128) return false;
        p->npad[l] = (int)mq_round_up(net->out[l], 4);
        p->ldwt[l] = (int)mq_round_up(net->in[l], 4) + 4;
        p->sw_off[l] = pos; pos += net->in[l] * p->npad[l];
        p->sb_off[l] = pos; pos += p->npad[l];
        if (l >= 1) { p->swt_off[l] = wt; wt += net->out[l] * p->ldwt[l]; }
        if (net->out[l] > maxw) maxw = net->out[l];
    }
    if ((int64_t)net->n_in * CH_TB > (int64_t)CH_PFX * CH_NT) return false;
    if ((int64_t)net->n_out * CH_TB > (int64_t)CH_PFH * CH_NT) return false;
    p->sls_off = pos;
    if (net->logstd_off >= 0) pos += (int)mq_round_up(net->n_out, 4);
    p->p_pad = pos;
    p->wt_total = (int)mq_round_up(wt, 4);
    int a = 0;
    p->act_off[0] = a; a += net->n_in * CH_TBP;
    for (int l = 0; l < net->n_layers; ++l) { p->act_off[l + 1] = a; a += net->out[l] * CH_TBP; }
    p->aux_off = a; a += (2 * net->n_out + 2) * CH_TBP;
    p->act_total = a;
    int dz = 0;
    for (int l = 0; l < net->n_layers; ++l) { p->dz_off[l] = dz; dz += net->out[l] * CH_TBP; }
    p->dz_total = dz;
    (void)maxw;
    const int64_t floats = (with_state ? 4LL : 2LL) * p->p_pad + p->wt_total + p->act_total + p->dz_total + 2 * CH_TB;
    p->smem_bytes = floats * (int64_t)sizeof(float);
    return p->smem_bytes <= mq_smem_optin();
}

struct ChainSmem {
    float *W, *G, *M, *V, *Wt, *A, *DZ;
    int *I;
    __device__ __forceinline__ ChainSmem(const ChainPlan &p, unsigned char *base) {
        W = (float *)base;
        G = W + p.p_pad;
        M = G + p.p_pad;
        V = M + p.p_pad;
        Wt = p.with_state ? V + p.p_pad : M;
        A = Wt + p.wt_total;
        DZ = A + p.act_total;
        I = (int *)(DZ + p.dz_total);
    }
};

// flat parameter index -> padded shared index (used once per launch)
template <int NW, typename F>
__device__ __forceinline__ void chain_for_each_param(const ChainPlan &p, F f) {
    constexpr int CH_NT = ChainCfg<NW>::NT, CH_TB = ChainCfg<NW>::TB, CH_TBP = ChainCfg<NW>::TBP, CH_RW = ChainCfg<NW>::RW;
    (void)CH_NT; (void)CH_TB; (void)CH_TBP; (void)CH_RW;
    for (int l = 0; l < p.net.n_layers; ++l) {
        const int K = p.net.in[l], N = p.net.out[l];
        for (int e = threadIdx.x; e < K * N; e += CH_NT) {
            const int k = e / N, n = e - k * N;
            f(p.net.w_off[l] + e, p.sw_off[l] + k * p.npad[l] + n);
        }
        for (int n = threadIdx.x; n < N; n += CH_NT) f(p.net.b_off[l] + n, p.sb_off[l] + n);
    }
    if (p.net.logstd_off >= 0)
        for (int a = threadIdx.x; a < p.net.n_out; a += CH_NT) f(p.net.logstd_off + a, p.sls_off + a);
}

template <int R> struct RowsV {
    static __device__ __forceinline__ void ld(const float *p, float *a) {
        if (R >= 4) {
#pragma unroll
            for (int q = 0; q < R / 4; ++q) {
                const float4 u = *reinterpret_cast<const float4 *>(p + 4 * q);
                a[4 * q] = u.x; a[4 * q + 1] = u.y; a[4 * q + 2] = u.z; a[4 * q + 3] = u.w;
            }
        } else if (R == 2) {
            const float2 u = *reinterpret_cast<const float2 *>(p);
            a[0] = u.x; a[1] = u.y;
        } else {
            a[0] = p[0];
        }
    }
    static __device__ __forceinline__ void st(float *p, const float *a) {
        if (R >= 4) {
#pragma unroll
            for (int q = 0; q < R / 4; ++q)
                *reinterpret_cast<float4 *>(p + 4 * q) = make_float4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
        } else if (R == 2) {
            *reinterpret_cast<float2 *>(p) = make_float2(a[0], a[1]);
        } else {
            p[0] = a[0];
        }
    }
};

// Warp-local GEMM for the warp's RW rows r0..r0+RW-1:
//   out[n][r] = epi( sum_k in[k][r] * w[k*ldw + n] )
// The 32 lanes form an LR x LC grid; a lane owns R = RW/LR rows and the 4 columns 4*lc..4*lc+3:
// per k one (broadcast) row-vector load and ONE 128-bit weight load feed 4R FFMAs.
//   BWD == false: epi = act(acc + bias[n]);  BWD == true: epi = act'(saved[n][r]) * acc
template <int NW, int LC, bool BWD>
__device__ __forceinline__ void warp_gemm(const float *__restrict__ in, int K, const float *__restrict__ w,
                                          int ldw, int N, float *__restrict__ out,
                                          const float *__restrict__ saved, int r0,
                                          const float *__restrict__ bias, int act, float leak) {
    constexpr int CH_TBP = ChainCfg<NW>::TBP, CH_RW = ChainCfg<NW>::RW;
    constexpr int LR = 32 / LC, R = (CH_RW / LR) > 0 ? (CH_RW / LR) : 1;
    const int lane = threadIdx.x & 31;
    const int lc = lane % LC, lr = lane / LC;
    const bool row_ok = (lr * R) < CH_RW;                  // lanes beyond the warp's rows idle (R == 1 case)
    const int rbase = r0 + (row_ok ? lr * R : 0);
    const int npad4 = (N + 3) >> 2;
    const int cb = (lc < npad4 ? lc : npad4 - 1) * 4;       // clamped column block (results discarded)
    float acc[4][R];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < R; ++i) acc[j][i] = 0.0f;
    const float *ip = in + rbase;
    const float *wp = w + cb;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        float a[R];
        RowsV<R>::ld(ip + k * CH_TBP, a);
        const float4 wv = *reinterpret_cast<const float4 *>(wp + k * ldw);
#pragma unroll
        for (int i = 0; i < R; ++i) {
            acc[0][i] = fmaf(a[i], wv.x, acc[0][i]);
            acc[1][i] = fmaf(a[i], wv.y, acc[1][i]);
            acc[2][i] = fmaf(a[i], wv.z, acc[2][i]);
            acc[3][i] = fmaf(a[i], wv.w, acc[3][i]);
        }
    }
    if (!row_ok || lc >= npad4) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = lc * 4 + j;
        if (n >= N) continue;
        float res[R];
        if (BWD) {
            float y[R];
            RowsV<R>::ld(saved + n * CH_TBP + rbase, y);
#pragma unroll
            for (int i = 0; i < R; ++i) res[i] = mq_act_grad(y[i], acc[j][i], act, leak);
        } else {
            const float bn = bias[n];
#pragma unroll
            for (int i = 0; i < R; ++i) res[i] = mq_act_apply(acc[j][i] + bn, act, leak);
        }
        RowsV<R>::st(out + n * CH_TBP + rbase, res);
    }
}

// Narrow forward layer (N <= 8): lane = (k-slice kq, row r); k-slices merged by shuffles.
template <int NW>
__device__ __forceinline__ void warp_gemm_narrow(const float *__restrict__ in, int K,
                                                 const float *__restrict__ w, int ldw, int N,
                                                 float *__restrict__ out, int r0,
                                                 const float *__restrict__ bias, int act, float leak) {
    constexpr int CH_NT = ChainCfg<NW>::NT, CH_TB = ChainCfg<NW>::TB, CH_TBP = ChainCfg<NW>::TBP, CH_RW = ChainCfg<NW>::RW;
    (void)CH_NT; (void)CH_TB; (void)CH_TBP; (void)CH_RW;
    const int lane = threadIdx.x & 31;
    constexpr int KQ = 32 / CH_RW;
    const int r = lane % CH_RW, kq = lane / CH_RW;
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = 0.0f;
    for (int k = kq; k < K; k += KQ) {
        const float a = in[k * CH_TBP + r0 + r];
        const float *wr = w + k * ldw;
#pragma unroll
        for (int n = 0; n < 8; ++n)
            if (n < N) acc[n] = fmaf(a, wr[n], acc[n]);
    }
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        if (n < N) {
#pragma unroll
            for (int o = CH_RW; o < 32; o <<= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
        }
    }
    if (kq == 0) {
#pragma unroll
        for (int n = 0; n < 8; ++n)
            if (n < N) out[n * CH_TBP + r0 + r] = mq_act_apply(acc[n] + bias[n], act, leak);
    }
}

template <int NW, bool BWD>
__device__ __forceinline__ void warp_layer(const float *in, int K, const float *w, int ldw, int N,
                                           float *out, const float *saved, int r0, const float *bias,
                                           int act, float leak) {
    if (!BWD && N <= 8) { warp_gemm_narrow<NW>(in, K, w, ldw, N, out, r0, bias, act, leak); return; }
    if (N <= 16) warp_gemm<NW, 4, BWD>(in, K, w, ldw, N, out, saved, r0, bias, act, leak);
    else if (N <= 32) warp_gemm<NW, 8, BWD>(in, K, w, ldw, N, out, saved, r0, bias, act, leak);
    else if (N <= 64) warp_gemm<NW, 16, BWD>(in, K, w, ldw, N, out, saved, r0, bias, act, leak);
    else warp_gemm<NW, 32, BWD>(in, K, w, ldw, N, out, saved, r0, bias, act, leak);
}

// Loss-gradient rule for the warp's rows (lanes 0..7, one row each).
template <int NW>
__device__ __forceinline__ void warp_head(const ChainPlan &p, const mq_head_t &h, const float *sW,
                                          const float *sAout, float *aux, float *dzout, int r0, int n_valid,
                                          int64_t grow0, float *__restrict__ out, float *__restrict__ logp_out) {
    constexpr int CH_NT = ChainCfg<NW>::NT, CH_TB = ChainCfg<NW>::TB, CH_TBP = ChainCfg<NW>::TBP, CH_RW = ChainCfg<NW>::RW;
    (void)CH_NT; (void)CH_TB; (void)CH_TBP; (void)CH_RW;
    const int lane = threadIdx.x & 31;
    if (lane >= CH_RW) return;
    const int N = p.net.n_out, r = r0 + lane;
    const int lastact = p.net.act[p.net.n_layers - 1];
    const float lastleak = p.net.leak[p.net.n_layers - 1];
    const float *tgt = aux;
    float *hterm = aux + N * CH_TBP;
    const float *misc = aux + 2 * N * CH_TBP;
    if (r >= n_valid) {
        for (int n = 0; n < N; ++n) { dzout[n * CH_TBP + r] = 0.0f; hterm[n * CH_TBP + r] = 0.0f; }
        return;
    }
    if (out)
        for (int n = 0; n < N; ++n) out[(grow0 + r) * N + n] = sAout[n * CH_TBP + r];
    if (h.kind == MQ_HEAD_DY) {
        for (int n = 0; n < N; ++n)
            dzout[n * CH_TBP + r] = mq_act_grad(sAout[n * CH_TBP + r], tgt[n * CH_TBP + r], lastact, lastleak);
    } else if (h.kind == MQ_HEAD_DIFF) {
        for (int n = 0; n < N; ++n) {
            const float y = sAout[n * CH_TBP + r];
            dzout[n * CH_TBP + r] = mq_act_grad(y, tgt[n * CH_TBP + r] - y, lastact, lastleak);
        }
    } else if (h.kind == MQ_HEAD_SOFTMAX_CE) {
        float mx = sAout[r];
        for (int n = 1; n < N; ++n) mx = fmaxf(mx, sAout[n * CH_TBP + r]);
        float sum = 0.0f;
        for (int n = 0; n < N; ++n) sum += expf(sAout[n * CH_TBP + r] - mx);
        const float inv = 1.0f / sum;
        for (int n = 0; n < N; ++n) {
            const float y = sAout[n * CH_TBP + r];
            const float dy = tgt[n * CH_TBP + r] - expf(y - mx) * inv - h.p0 * y;
            dzout[n * CH_TBP + r] = mq_act_grad(y, dy, lastact, lastleak);
        }
    } else if (mq_head_discrete(h.kind)) {
        // categorical log-prob (distrib.py:24-32): target row is the one-hot action
        float mx = sAout[r];
        for (int n = 1; n < N; ++n) mx = fmaxf(mx, sAout[n * CH_TBP + r]);
        float sum = 0.0f;
        for (int n = 0; n < N; ++n) sum += expf(sAout[n * CH_TBP + r] - mx);
        const float lse = mx + logf(sum);
        float lp = 0.0f;
        for (int n = 0; n < N; ++n) lp += tgt[n * CH_TBP + r] * (sAout[n * CH_TBP + r] - lse);
        if (logp_out) logp_out[grow0 + r] = lp;
        float w;
        if (h.kind == MQ_HEAD_DISCRETE_PPO) {
            const float adv = misc[r];
            const float ratio = expf(lp - misc[CH_TBP + r]);
            const bool kill = (ratio > h.p1 && adv > 0.0f) || (ratio < h.p0 && adv < 0.0f);
            w = kill ? 0.0f : ratio * adv;
        } else {
            w = misc[r];
        }
        for (int n = 0; n < N; ++n) {
            const float y = sAout[n * CH_TBP + r];
            dzout[n * CH_TBP + r] = mq_act_grad(y, w * (tgt[n * CH_TBP + r] - expf(y - lse)), lastact, lastleak);
            hterm[n * CH_TBP + r] = 0.0f;
        }
    } else {
        float acc = 0.0f;
        for (int a = 0; a < N; ++a) {
            const float ls = sW[p.sls_off + a];
            const float z = (tgt[a * CH_TBP + r] - sAout[a * CH_TBP + r]) * expf(-ls);
            acc += ls + 0.5f * z * z;
        }
        const float lp = -(float)N * CH_HALF_LOG_2PI - acc;
        if (logp_out) logp_out[grow0 + r] = lp;
        float wgt;
        if (h.kind == MQ_HEAD_GAUSS_PPO) {
            const float adv = misc[r];
            const float ratio = expf(lp - misc[CH_TBP + r]);
            const bool kill = (ratio > h.p1 && adv > 0.0f) || (ratio < h.p0 && adv < 0.0f);
            wgt = kill ? 0.0f : ratio * adv;
        } else {
            wgt = misc[r];
        }
        for (int a = 0; a < N; ++a) {
            const float inv = expf(-sW[p.sls_off + a]);
            const float y = sAout[a * CH_TBP + r];
            const float z = (tgt[a * CH_TBP + r] - y) * inv;
            hterm[a * CH_TBP + r] = wgt * (z * z - 1.0f) + h.ent;
            dzout[a * CH_TBP + r] = mq_act_grad(y, wgt * z * inv, lastact, lastleak);
        }
    }
}

// G[k*ldw + n] += sum_r X[k][r] dZ[n][r], Gb[n] += sum_r dZ[n][r]  (CTA-wide, no barrier inside).
// 4x4 register tile: k = kg + i*KG, n = ng + j*NG (interleaved -> conflict-free 128-bit loads).
// Small layers: S neighbouring lanes share a tile, each takes 64/S rows, and the partial sums
// are combined with warp shuffles in a fixed order.
template <int NW, int S>
__device__ __forceinline__ void chain_dw(const float *__restrict__ X, int K, const float *__restrict__ dZ,
                                         int N, float *__restrict__ G, float *__restrict__ Gb, int ldw) {
    constexpr int CH_NT = ChainCfg<NW>::NT, CH_TB = ChainCfg<NW>::TB, CH_TBP = ChainCfg<NW>::TBP, CH_RW = ChainCfg<NW>::RW;
    (void)CH_NT; (void)CH_TB; (void)CH_TBP; (void)CH_RW;
    const int KG = (K + 3) >> 2, NG = ldw >> 2;
    const int items = KG * NG;
    // Lane layout inside a warp: the row-split index s is the SLOW index (lanes 0..LPS-1 have
    // s = 0, ...), so a quarter-warp reads 8 different rows, 4 banks apart: conflict-free.
    constexpr int PER = CH_NT / S, SPAN = CH_TB / S, LPS = 32 / S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int s = lane / LPS;
    for (int base = 0; base < items; base += PER) {
        const int item_raw = base + warp * LPS + (lane % LPS);
        const bool active = item_raw < items;
        const int item = active ? item_raw : items - 1;
        const int kg = item / NG, ng = item - kg * NG;
        const float *xp[4], *zp[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            xp[i] = X + min(kg + i * KG, K - 1) * CH_TBP + s * SPAN;
            zp[i] = dZ + min(ng + i * NG, N - 1) * CH_TBP + s * SPAN;
        }
        float acc[4][4], bsum[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            bsum[j] = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][j] = 0.0f;
        }
#pragma unroll 2
        for (int r = 0; r < SPAN; r += 4) {
            float4 xv[4], zv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                xv[i] = *reinterpret_cast<const float4 *>(xp[i] + r);
                zv[i] = *reinterpret_cast<const float4 *>(zp[i] + r);
            }
            if (kg == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) bsum[j] += (zv[j].x + zv[j].y) + (zv[j].z + zv[j].w);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(xv[i].x, zv[j].x, acc[i][j]);
                    acc[i][j] = fmaf(xv[i].y, zv[j].y, acc[i][j]);
                    acc[i][j] = fmaf(xv[i].z, zv[j].z, acc[i][j]);
                    acc[i][j] = fmaf(xv[i].w, zv[j].w, acc[i][j]);
                }
        }
        if (S > 1) {
#pragma unroll
            for (int o = LPS; o < 32; o <<= 1) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    bsum[j] += __shfl_xor_sync(0xffffffffu, bsum[j], o);
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], o);
                }
            }
        }
        if (active && s == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = kg + i * KG;
                if (k >= K) continue;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = ng + j * NG;
                    if (n < N) G[k * ldw + n] += acc[i][j];
                }
            }
            if (kg == 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = ng + j * NG;
                    if (n < N) Gb[n] += bsum[j];
                }
            }
        }
    }
}

template <int NW>
__device__ __forceinline__ void chain_dw_any(const float *X, int K, const float *dZ, int N, float *G,
                                             float *Gb, int ldw) {
    constexpr int CH_NT = ChainCfg<NW>::NT, CH_TB = ChainCfg<NW>::TB, CH_TBP = ChainCfg<NW>::TBP, CH_RW = ChainCfg<NW>::RW;
    (void)CH_NT; (void)CH_TB; (void)CH_TBP; (void)CH_RW;
    const int items = ((K + 3) >> 2) * (ldw >> 2);
    if (items * 2 > CH_NT) chain_dw<NW, 1>(X, K, dZ, N, G, Gb, ldw);
    else if (items * 4 > CH_NT) chain_dw<NW, 2>(X, K, dZ, N, G, Gb, ldw);
    else if (items * 8 > CH_NT) chain_dw<NW, 4>(X, K, dZ, N, G, Gb, ldw);
    else chain_dw<NW, 8>(X, K, dZ, N, G, Gb, ldw);
}

template <int NW>
__device__ __forceinline__ void chain_rowsum(const float *__restrict__ Z, int N, float *__restrict__ G) {
    constexpr int CH_NT = ChainCfg<NW>::NT, CH_TB = ChainCfg<NW>::TB, CH_TBP = ChainCfg<NW>::TBP, CH_RW = ChainCfg<NW>::RW;
    (void)CH_NT; (void)CH_TB; (void)CH_TBP; (void)CH_RW;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int n = warp; n < N; n += CH_NT / 32) {
        float s = 0.0f;
        for (int r = lane * 4; r < CH_TB; r += 128) {
            const float4 v = *reinterpret_cast<const float4 *>(Z + n * CH_TBP + r);
            s += (v.x + v.y) + (v.z + v.w);
        }
        s = mq_warp_sum(s);
        if (lane == 0) G[n] += s;
    }
}

__device__ __forceinline__ float ch_sqrt(float x) {
#ifdef MQ_EMU
    return sqrtf(x);
#else
    float r;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
#endif
}

// Adam over the whole padded parameter block (padding entries stay exactly 0), 4 elements
// in flight per thread; then the transposed copies are rebuilt (layers >= 1).
template <int NW>
__device__ __forceinline__ void chain_adam(const ChainPlan &p, float *__restrict__ W, float *__restrict__ G,
                                           float *__restrict__ M, float *__restrict__ V, float inv_mb,
                                           float c1, float c2, float lr, float eps) {
    constexpr int CH_NT = ChainCfg<NW>::NT, CH_TB = ChainCfg<NW>::TB, CH_TBP = ChainCfg<NW>::TBP, CH_RW = ChainCfg<NW>::RW;
    (void)CH_NT; (void)CH_TB; (void)CH_TBP; (void)CH_RW;
    for (int e0 = threadIdx.x; e0 < p.p_pad; e0 += 4 * CH_NT) {
        float w[4], g[4], m[4], v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * CH_NT;
            if (e < p.p_pad) { w[u] = W[e]; g[u] = G[e] * inv_mb; m[u] = M[e]; v[u] = V[e]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * CH_NT;
            if (e >= p.p_pad) continue;
            m[u] += c1 * (g[u] - m[u]);                         // _normalize.py:24-25 via _adam.py:29-30
            v[u] += c2 * (g[u] * g[u] - v[u]);
            w[u] += lr * __fdividef(m[u], eps + ch_sqrt(v[u]));  // ascent, _adam.py:42,50-53
            W[e] = w[u]; M[e] = m[u]; V[e] = v[u]; G[e] = 0.0f;
        }
    }
}

template <int NW>
__device__ __forceinline__ void chain_transposes(const ChainPlan &p, const float *__restrict__ W,
                                                 float *__restrict__ Wt) {
    constexpr int CH_NT = ChainCfg<NW>::NT, CH_TB = ChainCfg<NW>::TB, CH_TBP = ChainCfg<NW>::TBP, CH_RW = ChainCfg<NW>::RW;
    (void)CH_NT; (void)CH_TB; (void)CH_TBP; (void)CH_RW;
    for (int l = 1; l < p.net.n_layers; ++l) {
        const int K = p.net.in[l], N = p.net.out[l], ld = p.npad[l], ldt = p.ldwt[l];
        const float *src = W + p.sw_off[l];
        float *dst = Wt + p.swt_off[l];
        if (CH_NT % ld == 0) {
            const int n = threadIdx.x % ld, kstep = CH_NT / ld;
            if (n < N) {
#pragma unroll 4
                for (int k = threadIdx.x / ld; k < K; k += kstep) dst[n * ldt + k] = src[k * ld + n];
            }
        } else {
            for (int e = threadIdx.x; e < K * N; e += CH_NT) {
                const int k = e / N, n = e - k * N;
                dst[n * ldt + k] = src[k * ld + n];
            }
        }
    }
}

template <int NW>
__global__ void __launch_bounds__(ChainCfg<NW>::NT)
chain_train_kernel(ChainPlan p, float *__restrict__ theta, float *__restrict__ value,
                   float *__restrict__ am, float *__restrict__ av, const float *__restrict__ x,
                   mq_head_t head, const int32_t *__restrict__ idx_table, int n_steps, int mb,
                   double beta1, double beta2, double tw1, double tw2, double lr, double eps,
                   long long *prof) {
    constexpr int CH_NT = ChainCfg<NW>::NT, CH_TB = ChainCfg<NW>::TB, CH_TBP = ChainCfg<NW>::TBP, CH_RW = ChainCfg<NW>::RW;
    (void)CH_NT; (void)CH_TB; (void)CH_TBP; (void)CH_RW;
    MQ_SHARED_BYTES(smem);
    ChainSmem sm(p, smem);
    long long t_last = 0;
#ifndef MQ_EMU
    if (prof && threadIdx.x == 0) t_last = clock64();
#endif
    const int tid = threadIdx.x, warp = tid >> 5;
    const int L = p.net.n_layers, K0 = p.net.n_in, NO = p.net.n_out;
    const bool gauss = head.kind == MQ_HEAD_GAUSS_LOGP || head.kind == MQ_HEAD_GAUSS_PPO;
    const int ntiles = (mb + CH_TB - 1) / CH_TB;
    const float inv_mb = 1.0f / (float)mb;
    const float *hmat = (head.kind == MQ_HEAD_DY) ? head.dy : head.target;
    const float *hv0 = mq_head_ppo(head.kind) ? head.adv : head.dy;
    const bool prob = mq_head_prob(head.kind);
    float *aux = sm.A + p.aux_off;

    // ---- load state: value/m/v (fp32) -> shared, padded entries zero
    for (int e = tid; e < 4 * p.p_pad + p.wt_total; e += CH_NT) sm.W[e] = 0.0f;
    __syncthreads();
    chain_for_each_param<NW>(p, [&](int flat, int pad) {
        sm.W[pad] = value[flat]; sm.M[pad] = am[flat]; sm.V[pad] = av[flat];
    });
    __syncthreads();
    for (int l = 1; l < L; ++l) {
        const int K = p.net.in[l], N = p.net.out[l];
        for (int e = tid; e < K * N; e += CH_NT) {
            const int k = e / N, n = e - k * N;
            sm.Wt[p.swt_off[l] + n * p.ldwt[l] + k] = sm.W[p.sw_off[l] + k * p.npad[l] + n];
        }
    }

    // ---- software pipeline of the gathers: idx two units ahead, data one unit ahead
    const int n_units = n_steps * ntiles;                       // unit u = (step, tile)
    auto unit_row = [&](int u, int r) -> int {                  // position in the index table or -1
        const int step = u / ntiles, tile = u - step * ntiles;
        const int gr = tile * CH_TB + r;
        return (gr < mb) ? step * mb + gr : -1;
    };
    if (tid < CH_TB) {
        const int a = unit_row(0, tid);
        sm.I[tid] = (a >= 0) ? idx_table[a] : -1;
        const int b = (n_units > 1) ? unit_row(1, tid) : -1;
        sm.I[CH_TB + tid] = (b >= 0) ? idx_table[b] : -1;
    }
    __syncthreads();
    float px[CH_PFX], ph[CH_PFH], pm = 0.0f;
    int pidx = -1;
    auto issue_loads = [&](int u) {
        const int *I = sm.I + (u & 1) * CH_TB;
#pragma unroll
        for (int q = 0; q < CH_PFX; ++q) {
            const int e = tid + q * CH_NT;
            px[q] = 0.0f;
            if (e < CH_TB * K0) {
                const int r = e / K0, k = e - r * K0;
                const int src = I[r];
                if (src >= 0) px[q] = x[(int64_t)src * K0 + k];
            }
        }
#pragma unroll
        for (int q = 0; q < CH_PFH; ++q) {
            const int e = tid + q * CH_NT;
            ph[q] = 0.0f;
            if (e < CH_TB * NO) {
                const int r = e / NO, n = e - r * NO;
                const int src = I[r];
                if (src >= 0) ph[q] = hmat[(int64_t)src * NO + n];
            }
        }
        pm = 0.0f;
        if (prob && tid < 2 * CH_TB) {
            const int which = tid / CH_TB, r = tid - which * CH_TB;
            const int src = I[r];
            if (src >= 0) {
                if (which == 0) pm = hv0[src];
                else if (mq_head_ppo(head.kind)) pm = head.logp_old[src];
            }
        }
    };
    issue_loads(0);

    int u = 0;
    for (int step = 0; step < n_steps; ++step) {
        for (int tile = 0; tile < ntiles; ++tile, ++u) {
            const int n_valid = min(CH_TB, mb - tile * CH_TB);
            // ---- stage the prefetched unit into shared memory
#pragma unroll
            for (int q = 0; q < CH_PFX; ++q) {
                const int e = tid + q * CH_NT;
                if (e < CH_TB * K0) { const int r = e / K0, k = e - r * K0; sm.A[k * CH_TBP + r] = px[q]; }
            }
#pragma unroll
            for (int q = 0; q < CH_PFH; ++q) {
                const int e = tid + q * CH_NT;
                if (e < CH_TB * NO) { const int r = e / NO, n = e - r * NO; aux[n * CH_TBP + r] = ph[q]; }
            }
            if (prob && tid < 2 * CH_TB) {
                const int which = tid / CH_TB, r = tid - which * CH_TB;
                aux[2 * NO * CH_TBP + which * CH_TBP + r] = pm;
            }
            if (tid < CH_TB && u >= 1 && u + 1 < n_units) sm.I[((u + 1) & 1) * CH_TB + tid] = pidx;
            __syncthreads();                                                        // barrier 1
            CH_MARK(0);
            // ---- start the gathers of the next unit; index rows two units ahead
            if (u + 1 < n_units) issue_loads(u + 1);
            if (tid < CH_TB && u + 2 < n_units) {
                const int a = unit_row(u + 2, tid);
                pidx = (a >= 0) ? idx_table[a] : -1;
            }
            // ---- warp-local: forward, head, back-propagated signal for rows r0..r0+7
            const int r0 = warp * CH_RW;
            for (int l = 0; l < L; ++l) {
                warp_layer<NW, false>(sm.A + p.act_off[l], p.net.in[l], sm.W + p.sw_off[l], p.npad[l],
                                  p.net.out[l], sm.A + p.act_off[l + 1], nullptr, r0, sm.W + p.sb_off[l],
                                  p.net.act[l], p.net.leak[l]);
                __syncwarp();
                CH_MARK(20 + l);
            }
            warp_head<NW>(p, head, sm.W, sm.A + p.act_off[L], aux, sm.DZ + p.dz_off[L - 1], r0, n_valid, 0, nullptr, nullptr);
            __syncwarp();
            CH_MARK(2);
            for (int l = L - 1; l >= 1; --l) {
                // dZ_{l-1} = (dZ_l W_l^T) * act'(A_l); one buffer per layer (dW runs later)
                warp_layer<NW, true>(sm.DZ + p.dz_off[l], p.net.out[l], sm.Wt + p.swt_off[l], p.ldwt[l],
                                 p.net.in[l], sm.DZ + p.dz_off[l - 1], sm.A + p.act_off[l], r0,
                                 nullptr, p.net.act[l - 1], p.net.leak[l - 1]);
                __syncwarp();
                CH_MARK(10 + 3 * l);
            }
            __syncthreads();                                                        // barrier 2
            CH_MARK(6);
            // ---- CTA-wide: weight / bias / logstd gradients (disjoint outputs, no barrier)
            if (gauss) chain_rowsum<NW>(aux + NO * CH_TBP, NO, sm.G + p.sls_off);
            for (int l = L - 1; l >= 0; --l) {
                chain_dw_any<NW>(sm.A + p.act_off[l], p.net.in[l], sm.DZ + p.dz_off[l], p.net.out[l],
                             sm.G + p.sw_off[l], sm.G + p.sb_off[l], p.npad[l]);
                CH_MARK(9 + 3 * l);
            }
            __syncthreads();                                                        // barrier 3
            CH_MARK(7);
        }
        // ---- Adam (fp32 moments on chip) merged with the refresh of the transposed copies
        tw1 = tw1 * beta1 + 1.0;                       // _normalize.py:18-23 with weight 1
        tw2 = tw2 * beta2 + 1.0;
        const float c1 = (float)(1.0 / tw1), c2 = (float)(1.0 / tw2), flr = (float)lr, feps = (float)eps;
        chain_adam<NW>(p, sm.W, sm.G, sm.M, sm.V, inv_mb, c1, c2, flr, feps);
        __syncthreads();
        CH_MARK(5);
        chain_transposes<NW>(p, sm.W, sm.Wt);
        CH_MARK(4);
        // no barrier needed here: the next unit's barrier 1 orders these writes before any read
    }
    __syncthreads();
    chain_for_each_param<NW>(p, [&](int flat, int pad) {
        const float w = sm.W[pad];
        value[flat] = w; theta[flat] = w; am[flat] = sm.M[pad]; av[flat] = sm.V[pad];
    });
}

// Gradient of one (large) minibatch with the same row-owner-warp structure: every CTA strides
// over tiles of 8*NW rows, accumulates dW/db/dlogstd in shared memory and writes ONE partial
// (or, for a single CTA, the scaled gradient).  Used by mq_fused_grad when the net fits.
template <int NW>
__global__ void __launch_bounds__(ChainCfg<NW>::NT)
chain_grad_kernel(ChainPlan p, const float *__restrict__ theta, const float *__restrict__ x,
                  const int32_t *__restrict__ idx, int64_t n_rows, mq_head_t head, float scale,
                  float *__restrict__ grad_or_partial, int direct, float *__restrict__ out,
                  float *__restrict__ logp) {
    constexpr int CH_NT = ChainCfg<NW>::NT, CH_TB = ChainCfg<NW>::TB, CH_TBP = ChainCfg<NW>::TBP, CH_RW = ChainCfg<NW>::RW;
    MQ_SHARED_BYTES(smem);
    ChainSmem sm(p, smem);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int L = p.net.n_layers, K0 = p.net.n_in, NO = p.net.n_out;
    const bool gauss = head.kind == MQ_HEAD_GAUSS_LOGP || head.kind == MQ_HEAD_GAUSS_PPO;
    const float *hmat = (head.kind == MQ_HEAD_DY) ? head.dy : head.target;
    const float *hv0 = mq_head_ppo(head.kind) ? head.adv : head.dy;
    const bool prob = mq_head_prob(head.kind);
    float *aux = sm.A + p.aux_off;

    for (int e = tid; e < 2 * p.p_pad + p.wt_total; e += CH_NT) sm.W[e] = 0.0f;
    __syncthreads();
    chain_for_each_param<NW>(p, [&](int flat, int pad) { sm.W[pad] = theta[flat]; });
    __syncthreads();
    chain_transposes<NW>(p, sm.W, sm.Wt);

    const int64_t ntiles = (n_rows + CH_TB - 1) / CH_TB;
    const int n_units = ntiles > blockIdx.x ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    auto unit_src = [&](int u, int r) -> int {          // source row of (unit, row) or -1 past the end
        const int64_t gr = ((int64_t)blockIdx.x + (int64_t)u * gridDim.x) * CH_TB + r;
        if (gr >= n_rows) return -1;
        return idx ? idx[gr] : (int)gr;
    };
    if (tid < CH_TB) {
        sm.I[tid] = n_units > 0 ? unit_src(0, tid) : -1;
        sm.I[CH_TB + tid] = n_units > 1 ? unit_src(1, tid) : -1;
    }
    __syncthreads();
    float px[CH_PFX], ph[CH_PFH], pm = 0.0f;
    int pidx = -1;
    auto issue_loads = [&](int u) {
        const int *I = sm.I + (u & 1) * CH_TB;
#pragma unroll
        for (int q = 0; q < CH_PFX; ++q) {
            const int e = tid + q * CH_NT;
            px[q] = 0.0f;
            if (e < CH_TB * K0) {
                const int r = e / K0, k = e - r * K0;
                const int src = I[r];
                if (src >= 0) px[q] = x[(int64_t)src * K0 + k];
            }
        }
#pragma unroll
        for (int q = 0; q < CH_PFH; ++q) {
            const int e = tid + q * CH_NT;
            ph[q] = 0.0f;
            if (e < CH_TB * NO) {
                const int r = e / NO, n = e - r * NO;
                const int src = I[r];
                if (src >= 0) ph[q] = hmat[(int64_t)src * NO + n];
            }
        }
        pm = 0.0f;
        if (prob && tid < 2 * CH_TB) {
            const int which = tid / CH_TB, r = tid - which * CH_TB;
            const int src = I[r];
            if (src >= 0) {
                if (which == 0) pm = hv0[src];
                else if (mq_head_ppo(head.kind)) pm = head.logp_old[src];
            }
        }
    };
    if (n_units > 0) issue_loads(0);

    for (int u = 0; u < n_units; ++u) {
        const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)u * gridDim.x) * CH_TB;
        const int n_valid = (int)(n_rows - row0 < CH_TB ? n_rows - row0 : CH_TB);
#pragma unroll
        for (int q = 0; q < CH_PFX; ++q) {
            const int e = tid + q * CH_NT;
            if (e < CH_TB * K0) { const int r = e / K0, k = e - r * K0; sm.A[k * CH_TBP + r] = px[q]; }
        }
#pragma unroll
        for (int q = 0; q < CH_PFH; ++q) {
            const int e = tid + q * CH_NT;
            if (e < CH_TB * NO) { const int r = e / NO, n = e - r * NO; aux[n * CH_TBP + r] = ph[q]; }
        }
        if (prob && tid < 2 * CH_TB) {
            const int which = tid / CH_TB, r = tid - which * CH_TB;
            aux[2 * NO * CH_TBP + which * CH_TBP + r] = pm;
        }
        if (tid < CH_TB && u >= 1 && u + 1 < n_units) sm.I[((u + 1) & 1) * CH_TB + tid] = pidx;
        __syncthreads();                                                            // barrier 1
        if (u + 1 < n_units) issue_loads(u + 1);
        if (tid < CH_TB && u + 2 < n_units) pidx = unit_src(u + 2, tid);
        const int r0 = warp * CH_RW;
        for (int l = 0; l < L; ++l) {
            warp_layer<NW, false>(sm.A + p.act_off[l], p.net.in[l], sm.W + p.sw_off[l], p.npad[l],
                                  p.net.out[l], sm.A + p.act_off[l + 1], nullptr, r0, sm.W + p.sb_off[l],
                                  p.net.act[l], p.net.leak[l]);
            __syncwarp();
        }
        warp_head<NW>(p, head, sm.W, sm.A + p.act_off[L], aux, sm.DZ + p.dz_off[L - 1], r0, n_valid, row0,
                      out, logp);
        __syncwarp();
        for (int l = L - 1; l >= 1; --l) {
            warp_layer<NW, true>(sm.DZ + p.dz_off[l], p.net.out[l], sm.Wt + p.swt_off[l], p.ldwt[l],
                                 p.net.in[l], sm.DZ + p.dz_off[l - 1], sm.A + p.act_off[l], r0, nullptr,
                                 p.net.act[l - 1], p.net.leak[l - 1]);
            __syncwarp();
        }
        __syncthreads();                                                            // barrier 2
        if (gauss) chain_rowsum<NW>(aux + NO * CH_TBP, NO, sm.G + p.sls_off);
        for (int l = L - 1; l >= 0; --l)
            chain_dw_any<NW>(sm.A + p.act_off[l], p.net.in[l], sm.DZ + p.dz_off[l], p.net.out[l],
                             sm.G + p.sw_off[l], sm.G + p.sb_off[l], p.npad[l]);
        __syncthreads();                                                            // barrier 3
    }
    float *dst = direct ? grad_or_partial : grad_or_partial + (int64_t)blockIdx.x * p.net.n_params;
    const float sc = direct ? scale : 1.0f;
    chain_for_each_param<NW>(p, [&](int flat, int pad) { dst[flat] = sm.G[pad] * sc; });
}

// Launch helper used by mq_fused_grad.  *grid_out receives the number of CTAs (= partials).
int mq_chain_grad(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                  int64_t batch, const mq_head_t *head, float scale, float *grad, float *out, float *logp,
                  void *ws, int64_t ws_bytes, int *grid_out, void *stream) {
    ChainPlan p;
    const int sms = mq_sm_count();
    // 16 warps / 128-row tiles when the batch feeds every SM with them, else 8 warps / 64 rows
    int nw = (batch >= (int64_t)sms * 128) ? 16 : 8;
    int rw = 8;
    if (!chain_plan(net, &p, nw, rw, 0)) {
        if (chain_plan(net, &p, 8, 8, 0)) { nw = 8; rw = 8; }
        else return MQ_ERR_UNSUPPORTED;
    }
    const int tb = rw * nw;
    int64_t grid = mq_ceil_div(batch, tb);
    if (grid > sms) grid = sms;
    if (grid < 1) grid = 1;
    const int direct = grid == 1;
    if (!direct && (!ws || ws_bytes < grid * net->n_params * (int64_t)sizeof(float))) {
        mq_set_error("mq_fused_grad: workspace too small");
        return MQ_ERR_SHAPE;
    }
    float *dst = direct ? grad : (float *)ws;
#ifndef MQ_EMU
#define MQ_CHAIN_SMEM(K)                                                                              \
    do {                                                                                              \
        cudaError_t e = cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                             (int)p.smem_bytes);                                      \
        if (e != cudaSuccess) {                                                                       \
            cudaGetLastError();                                                                       \
            mq_set_error("mq_fused_grad: cannot reserve %lld bytes of shared memory (%s)",            \
                         (long long)p.smem_bytes, cudaGetErrorString(e));                             \
            return MQ_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)
#else
#define MQ_CHAIN_SMEM(K) do { } while (0)
#endif
    if (rw == 16) {
        auto kfn = chain_grad_kernel<CH_CF(8, 16)>;
        MQ_CHAIN_SMEM(kfn);
        MQ_LAUNCH(kfn, (unsigned)grid, 256, p.smem_bytes, stream, p, theta, x, row_idx, batch, *head, scale, dst,
                  direct, out, logp);
    } else if (nw == 16) {
        auto kfn = chain_grad_kernel<CH_CF(16, 8)>;
        MQ_CHAIN_SMEM(kfn);
        MQ_LAUNCH(kfn, (unsigned)grid, 512, p.smem_bytes, stream, p, theta, x, row_idx, batch, *head, scale, dst,
                  direct, out, logp);
    } else {
        auto kfn = chain_grad_kernel<CH_CF(8, 8)>;
        MQ_CHAIN_SMEM(kfn);
        MQ_LAUNCH(kfn, (unsigned)grid, 256, p.smem_bytes, stream, p, theta, x, row_idx, batch, *head, scale, dst,
                  direct, out, logp);
    }
#undef MQ_CHAIN_SMEM
    *grid_out = (int)grid;
    MQ_CHECK_LAUNCH("mq_fused_grad");
    return MQ_OK;
}

// Host side: returns MQ_ERR_UNSUPPORTED when the net does not fit this kernel (the caller
// then uses the tile-based kernel of mq_fused.cu).
int mq_chain_train_f32(const mq_net_t *net, float *theta, float *value, float *m, float *v, const float *x,
                       const mq_head_t *head, const int32_t *idx_table, int32_t n_steps, int32_t mb,
                       double beta1, double beta2, const double *tw, double lr, double eps, void *stream) {
    ChainPlan p;
    if (!chain_plan(net, &p, 8, 8, 1)) return MQ_ERR_UNSUPPORTED;
    auto kfn = chain_train_kernel<CH_CF(8, 8)>;
#ifndef MQ_EMU
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)p.smem_bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        mq_set_error("mq_fused_train: cannot reserve %lld bytes of shared memory (%s)",
                     (long long)p.smem_bytes, cudaGetErrorString(e));
        return MQ_ERR_CUDA;
    }
#endif
    MQ_LAUNCH(kfn, 1, 256, p.smem_bytes, stream, p, theta, value, m, v, x, *head, idx_table,
              (int)n_steps, (int)mb, beta1, beta2, tw[0], tw[1], lr, eps, mq_phase_prof_ptr());
    MQ_CHECK_LAUNCH("mq_fused_train");
    return MQ_OK;
}
