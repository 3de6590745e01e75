// Fused small-MLP kernels: the whole Affine/activation stack of a policy- or
// value-sized net lives in shared memory, a CTA pushes a tile of batch rows
// forward through every layer, evaluates the loss-gradient rule ("head") and
// walks back, accumulating dW/db on chip.  Activations never touch HBM.
//
// Replaces, per launch, the Python/NumPy sequence
//   Function.evaluate     mannequin/basicnet.py:192-207   (matmul, add, tanh/relu per layer)
//   head                  benchmarks/gae.py:22, benchmarks/ppo.py:34-37, mnist/mlp.py:33,
//                         mannequin/distrib.py:61-68
//   backprop              mannequin/basicnet.py:209-257 + mannequin/backprop.py:11-60
//   (mq_fused_train also) mannequin/_adam.py:23-53 and Trajectory.__getitem__ gathers
//
// Shared-memory layout (floats), tbp = tile_rows + 4 (so that tbp = 4 mod 32):
//   sW   layer l as [in][npad] + bias[npad] (+ logstd); weight COLUMNS ARE PERMUTED,
//        pos(n) = (n % CG)*4 + n / CG with CG = npad/4, so that the four columns a thread
//        owns (cg, cg+CG, cg+2CG, cg+3CG) are one 128-bit word AND neighbouring threads write
//        neighbouring activation rows (conflict-free epilogue stores)
//   sWt  transposed copy of layers >= 1, [out][kpad+4], columns permuted the same way
//   sG   gradient accumulators, same layout as sW (so Adam is a straight loop over sW/sG)
//   sA_l activations, TRANSPOSED [width_l][tbp]: a thread reads its RM rows of one feature
//        with one vector load; dW (a contraction over rows) reads both operands unit-stride
//   DZ   two ping-pong buffers for the back-propagated signal, so dW_l and dX_l need no
//        barrier between them
// FFMA register tiles: forward / g W^T use (RM rows x 4 cols) per thread, dW a 4x4 tile over
// (in,out) with the row range split across thread groups when the layer is small; the bias
// gradient rides along in the dW tiles that own input row 0; the logstd gradient is a
// warp-shuffle row sum.  Every reduction has a fixed order: bit-reproducible run to run.
#include "mq_common.cuh"

// Developer aid: per-phase cycle counters (buffer is NULL in normal use).
struct PhaseProf { long long *buf; long long last; };
#ifndef MQ_EMU
#define PHASE_MARK(pp, i) do { if ((pp).buf && threadIdx.x == 0) { long long now_ = clock64(); (pp).buf[i] += now_ - (pp).last; (pp).last = now_; } } while (0)
#else
#define PHASE_MARK(pp, i) do { } while (0)
#endif

#define HALF_LOG_2PI 0.91893853320467274178f
#define DW_SLOTS 20                  // scratch floats per thread: 16 dW partials + 4 bias partials
#define TRAIN_NT 256                 // threads of the persistent training CTA

struct FusedPlan {
    mq_net_t net;
    int tb, tbp;
    int npad[MQ_MAX_LAYERS], kpad[MQ_MAX_LAYERS], ldwt[MQ_MAX_LAYERS];
    int sw_off[MQ_MAX_LAYERS], sb_off[MQ_MAX_LAYERS], swt_off[MQ_MAX_LAYERS];
    int sls_off;
    int p_pad, wt_total;
    int act_off[MQ_MAX_LAYERS + 1];
    int aux_off;                     // head inputs: target[n_out][tbp], hterm[n_out][tbp], misc[2][tbp]
    int act_total;
    int dz_floats;                   // one ping-pong buffer
    int scratch_floats;
    int with_grad;
    int64_t smem_bytes;
};

static int64_t plan_build(const mq_net_t *net, int tb, int with_grad, FusedPlan *p, int smem_state, int nt) {
    memset(p, 0, sizeof(*p));
    p->net = *net;
    p->tb = tb;
    p->tbp = tb + 4;
    p->with_grad = with_grad;
    int pos = 0, wt = 0, maxw = 4;
    for (int l = 0; l < net->n_layers; ++l) {
        p->npad[l] = (int)mq_round_up(net->out[l], 4);
        p->kpad[l] = (int)mq_round_up(net->in[l], 4);
        p->ldwt[l] = p->kpad[l] + 4;
        p->sw_off[l] = pos; pos += net->in[l] * p->npad[l];
        p->sb_off[l] = pos; pos += p->npad[l];
        if (with_grad && l >= 1) { p->swt_off[l] = wt; wt += net->out[l] * p->ldwt[l]; }
        if (net->out[l] > maxw) maxw = net->out[l];
    }
    p->sls_off = pos;
    if (net->logstd_off >= 0) pos += (int)mq_round_up(net->n_out, 4);
    p->p_pad = pos;
    p->wt_total = wt;
    int a = 0;
    p->act_off[0] = a; a += net->n_in * p->tbp;
    for (int l = 0; l < net->n_layers; ++l) { p->act_off[l + 1] = a; a += net->out[l] * p->tbp; }
    p->aux_off = a; a += (2 * net->n_out + 2) * p->tbp;
    p->act_total = a;
    p->dz_floats = maxw * p->tbp;
    p->scratch_floats = nt * DW_SLOTS;
    int64_t floats = (int64_t)p->p_pad + p->act_total;
    if (with_grad) floats += (int64_t)p->wt_total + p->p_pad + 2 * p->dz_floats + p->scratch_floats;
    if (smem_state) floats += 2 * p->p_pad;
    p->smem_bytes = floats * (int64_t)sizeof(float);
    return p->smem_bytes;
}


static int plan_choose(const mq_net_t *net, int64_t batch, int with_grad, FusedPlan *p, int smem_state = 0,
                       int nt = MQ_NT) {
    const int64_t lim = mq_smem_optin();
    int tb = 128;
    if (batch <= 32) tb = 32; else if (batch <= 64) tb = 64;
    // forward only (a few thousand rows per time step of a rollout): smaller tiles until every SM has one
    if (!with_grad) while (tb > 32 && batch / tb < mq_sm_count()) tb >>= 1;
    for (; tb >= 32; tb >>= 1)
        if (plan_build(net, tb, with_grad, p, smem_state, nt) <= lim) return tb;
    return 0;
}

struct Smem {
    float *W, *A, *Wt, *G, *DZ, *S, *M, *V;
    __device__ __forceinline__ Smem(const FusedPlan &p, unsigned char *base) {
        W = (float *)base;
        A = W + p.p_pad;
        Wt = A + p.act_total;
        G = Wt + p.wt_total;
        DZ = G + p.p_pad;
        S = DZ + 2 * p.dz_floats;
        M = S + p.scratch_floats;
        V = M + p.p_pad;
    }
};

__device__ __forceinline__ int perm_pos(int n, int cg_count) { return (n % cg_count) * 4 + n / cg_count; }

// flat parameter j -> padded shared index (runs once per launch: divisions are fine here)
template <int NT, typename F>
__device__ __forceinline__ void for_each_param(const FusedPlan &p, F f) {
    for (int l = 0; l < p.net.n_layers; ++l) {
        const int K = p.net.in[l], N = p.net.out[l], CG = p.npad[l] >> 2;
        for (int e = threadIdx.x; e < K * N; e += NT) {
            const int k = e / N, n = e - k * N;
            f(p.net.w_off[l] + e, p.sw_off[l] + k * p.npad[l] + perm_pos(n, CG));
        }
        for (int n = threadIdx.x; n < N; n += NT) f(p.net.b_off[l] + n, p.sb_off[l] + n);
    }
    if (p.net.logstd_off >= 0)
        for (int a = threadIdx.x; a < p.net.n_out; a += NT) f(p.net.logstd_off + a, p.sls_off + a);
}

__device__ __forceinline__ int flat_to_pad(const FusedPlan &p, int j) {
    for (int l = 0; l < p.net.n_layers; ++l) {
        const int N = p.net.out[l];
        if (j < p.net.b_off[l]) {
            const int e = j - p.net.w_off[l];
            const int k = e / N, n = e - k * N;
            return p.sw_off[l] + k * p.npad[l] + perm_pos(n, p.npad[l] >> 2);
        }
        if (j < p.net.b_off[l] + N) return p.sb_off[l] + (j - p.net.b_off[l]);
    }
    return p.sls_off + (j - p.net.logstd_off);
}

// sWt[n][pos(k)] = sW[k][pos(n)] for layers >= 1.  n is the fast index (<= 4-way conflicts on
// both sides); when N divides NT every thread keeps its n and walks k without any division.
template <int NT>
__device__ void refresh_transposes(const FusedPlan &p, const float *sW, float *sWt) {
    for (int l = 1; l < p.net.n_layers; ++l) {
        const int K = p.net.in[l], N = p.net.out[l], CG = p.npad[l] >> 2, CGK = p.kpad[l] >> 2;
        const float *src = sW + p.sw_off[l];
        float *dst = sWt + p.swt_off[l];
        if (NT % N == 0) {
            const int n = threadIdx.x % N, kstep = NT / N;
            const int srcn = perm_pos(n, CG), dstn = n * p.ldwt[l];
            int k = threadIdx.x / N;
            int kq = k / CGK, kr = k - kq * CGK;
            for (; k < K; k += kstep) {
                dst[dstn + kr * 4 + kq] = src[k * p.npad[l] + srcn];
                kr += kstep;
                while (kr >= CGK) { kr -= CGK; ++kq; }
            }
        } else {
            for (int e = threadIdx.x; e < K * N; e += NT) {
                const int k = e / N, n = e - k * N;
                dst[n * p.ldwt[l] + perm_pos(k, CGK)] = src[k * p.npad[l] + perm_pos(n, CG)];
            }
        }
    }
}

// ---- weights: flat theta (+ optional per-member noise) -> padded shared layout --------
template <int NT>
__device__ void load_weights(const FusedPlan &p, const float *__restrict__ theta,
                             const float *__restrict__ noise, float *sW, float *sWt) {
    for (int e = threadIdx.x; e < p.p_pad; e += NT) sW[e] = 0.0f;
    if (sWt) for (int e = threadIdx.x; e < p.wt_total; e += NT) sWt[e] = 0.0f;
    __syncthreads();
    for_each_param<NT>(p, [&](int flat, int pad) {
        float w = theta[flat];
        if (noise) w += noise[flat];
        sW[pad] = w;
    });
    __syncthreads();
    if (sWt) { refresh_transposes<NT>(p, sW, sWt); __syncthreads(); }
}

// ---- input tile (optionally gathered) -> sA0[k][r]; head inputs -> aux ------------------
template <int NT>
__device__ void load_rows(const FusedPlan &p, const float *__restrict__ x,
                          const int32_t *__restrict__ idx, int64_t row0, int64_t n_rows, float *sA0) {
    const int K = p.net.n_in;
    for (int e = threadIdx.x; e < p.tb * K; e += NT) {
        const int r = e / K, k = e - r * K;
        const int64_t gr = row0 + r;
        float v = 0.0f;
        if (gr < n_rows) {
            const int64_t src = idx ? (int64_t)idx[gr] : gr;
            v = x[src * K + k];
        }
        sA0[k * p.tbp + r] = v;
    }
}

template <int NT>
__device__ void load_head_inputs(const FusedPlan &p, const mq_head_t &h, const int32_t *__restrict__ idx,
                                 int64_t row0, int64_t n_rows, float *aux) {
    const int N = p.net.n_out, tbp = p.tbp;
    const float *mat = (h.kind == MQ_HEAD_DY) ? h.dy : h.target;
    for (int e = threadIdx.x; e < p.tb * N; e += NT) {
        const int r = e / N, n = e - r * N;
        const int64_t gr = row0 + r;
        float v = 0.0f;
        if (gr < n_rows) v = mat[(idx ? (int64_t)idx[gr] : gr) * N + n];
        aux[n * tbp + r] = v;
    }
    if (mq_head_prob(h.kind)) {
        float *misc = aux + 2 * N * tbp;
        const float *v0 = mq_head_ppo(h.kind) ? h.adv : h.dy;
        for (int e = threadIdx.x; e < 2 * p.tb; e += NT) {
            const int which = e / p.tb, r = e - which * p.tb;
            const int64_t gr = row0 + r;
            float v = 0.0f;
            if (gr < n_rows) {
                const int64_t src = idx ? (int64_t)idx[gr] : gr;
                if (which == 0) v = v0[src];
                else if (mq_head_ppo(h.kind)) v = h.logp_old[src];
            }
            misc[which * tbp + r] = v;
        }
    }
}

template <int RM> struct RowVec;
template <> struct RowVec<1> {
    static __device__ __forceinline__ void ld(const float *p, float *a) { a[0] = p[0]; }
    static __device__ __forceinline__ void st(float *p, const float *a) { p[0] = a[0]; }
};
template <> struct RowVec<2> {
    static __device__ __forceinline__ void ld(const float *p, float *a) {
        float2 v = *reinterpret_cast<const float2 *>(p); a[0] = v.x; a[1] = v.y; }
    static __device__ __forceinline__ void st(float *p, const float *a) {
        *reinterpret_cast<float2 *>(p) = make_float2(a[0], a[1]); }
};
template <> struct RowVec<4> {
    static __device__ __forceinline__ void ld(const float *p, float *a) {
        float4 v = *reinterpret_cast<const float4 *>(p); a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }
    static __device__ __forceinline__ void st(float *p, const float *a) {
        *reinterpret_cast<float4 *>(p) = make_float4(a[0], a[1], a[2], a[3]); }
};
template <> struct RowVec<8> {
    static __device__ __forceinline__ void ld(const float *p, float *a) { RowVec<4>::ld(p, a); RowVec<4>::ld(p + 4, a + 4); }
    static __device__ __forceinline__ void st(float *p, const float *a) { RowVec<4>::st(p, a); RowVec<4>::st(p + 4, a + 4); }
};

// out[n][r] = epi( sum_k in[k][r] * w[k*ldw + pos(n)] ),  in: [K][tbp], out: [N][tbp]
//   BWD == false: out = act(acc + bias[n])
//   BWD == true : out = act'(saved[n][r]) * acc
template <int NT, int RM, bool BWD>
__device__ void tile_gemm(const float *__restrict__ in, int K, const float *__restrict__ w, int ldw,
                          int N, int CG, float *__restrict__ out, const float *__restrict__ saved, int tb,
                          int tbp, const float *__restrict__ bias, int act, float leak) {
    const int items = CG * (tb / RM);
    for (int item = threadIdx.x; item < items; item += NT) {
        const int cg = item % CG, rg = item / CG;
        const float *ip = in + rg * RM;
        const float *wp = w + cg * 4;
        float acc[4][RM];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int i = 0; i < RM; ++i) acc[j][i] = 0.0f;
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            float a[RM];
            RowVec<RM>::ld(ip + k * tbp, a);
            const float4 wv = *reinterpret_cast<const float4 *>(wp + k * ldw);
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                acc[0][i] = fmaf(a[i], wv.x, acc[0][i]);
                acc[1][i] = fmaf(a[i], wv.y, acc[1][i]);
                acc[2][i] = fmaf(a[i], wv.z, acc[2][i]);
                acc[3][i] = fmaf(a[i], wv.w, acc[3][i]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = cg + j * CG;
            if (n >= N) continue;
            float *op = out + n * tbp + rg * RM;
            float res[RM];
            if (BWD) {
                float y[RM];
                RowVec<RM>::ld(saved + n * tbp + rg * RM, y);
#pragma unroll
                for (int i = 0; i < RM; ++i) res[i] = mq_act_grad(y[i], acc[j][i], act, leak);
            } else {
                const float bn = bias[n];
#pragma unroll
                for (int i = 0; i < RM; ++i) res[i] = mq_act_apply(acc[j][i] + bn, act, leak);
            }
            RowVec<RM>::st(op, res);
        }
    }
}

template <int NT, bool BWD>
__device__ __forceinline__ void tile_gemm_any(const float *in, int K, const float *w, int ldw, int N,
                                              int CG, float *out, const float *saved, int tb, int tbp,
                                              const float *bias, int act, float leak) {
    int rm = 8;
    while (rm > 1 && (tb / rm) * CG < NT) rm >>= 1;
    if (rm == 8) tile_gemm<NT, 8, BWD>(in, K, w, ldw, N, CG, out, saved, tb, tbp, bias, act, leak);
    else if (rm == 4) tile_gemm<NT, 4, BWD>(in, K, w, ldw, N, CG, out, saved, tb, tbp, bias, act, leak);
    else if (rm == 2) tile_gemm<NT, 2, BWD>(in, K, w, ldw, N, CG, out, saved, tb, tbp, bias, act, leak);
    else tile_gemm<NT, 1, BWD>(in, K, w, ldw, N, CG, out, saved, tb, tbp, bias, act, leak);
}

template <int NT>
__device__ void forward_tile(const FusedPlan &p, const float *sW, float *sA, PhaseProf &pp) {
    for (int l = 0; l < p.net.n_layers; ++l) {
        tile_gemm_any<NT, false>(sA + p.act_off[l], p.net.in[l], sW + p.sw_off[l], p.npad[l], p.net.out[l],
                                 p.npad[l] >> 2, sA + p.act_off[l + 1], nullptr, p.tb, p.tbp,
                                 sW + p.sb_off[l], p.net.act[l], p.net.leak[l]);
        __syncthreads();
        PHASE_MARK(pp, 20 + l);
    }
}

// G[k*ldw + pos(n)] += sum_r X[k][r] * dZ[n][r]  and  Gb[n] += sum_r dZ[n][r].
// Thread tile: k = kg + i*KG, n = ng + j*NG (interleaved, so neighbouring threads read
// neighbouring rows 4 banks apart: conflict free).  The tiles with kg == 0 also carry the
// bias sums.  Small layers split the row range over S thread groups; partials go through
// `scratch` laid out [group][slot][item] and are summed in a fixed order by ALL threads.
template <int NT>
__device__ void tile_dw(const float *__restrict__ X, int K, const float *__restrict__ dZ, int N,
                        float *G, float *Gb, int ldw, int tb, int tbp, float *scratch) {
    const int KG = (K + 3) >> 2, NG = ldw >> 2;
    const int items = KG * NG;
    int S = 1;
    if (items < NT) {
        while (S * 2 * items <= NT && S * 2 <= tb / 4) S <<= 1;
    }
    const int span = tb / S;
    const int nloop = (S > 1) ? 1 : (items + NT - 1) / NT;
    for (int it = 0; it < nloop; ++it) {
        const int lin = threadIdx.x + it * NT;
        const int s = (S > 1) ? lin / items : 0;
        const int item = (S > 1) ? lin - s * items : lin;
        const bool active = (S > 1) ? (s < S) : (item < items);
        float acc[4][4], bsum[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f; bsum[i] = 0.0f; }
        const int kg = item / NG, ng = item - kg * NG;
        if (active) {
            const float *xp[4], *zp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = min(kg + i * KG, K - 1), n = min(ng + i * NG, N - 1);
                xp[i] = X + k * tbp;
                zp[i] = dZ + n * tbp;
            }
            for (int r = s * span; r < (s + 1) * span; r += 4) {
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
        }
        if (S == 1) {
            if (active) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = kg + i * KG;
                    if (k >= K) continue;
                    float4 *gp = reinterpret_cast<float4 *>(G + k * ldw + ng * 4);
                    float4 g = *gp;
                    if (ng < N) g.x += acc[i][0];
                    if (ng + NG < N) g.y += acc[i][1];
                    if (ng + 2 * NG < N) g.z += acc[i][2];
                    if (ng + 3 * NG < N) g.w += acc[i][3];
                    *gp = g;
                }
                if (kg == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (ng + j * NG < N) Gb[ng + j * NG] += bsum[j];
                }
            }
        } else {
            if (active) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) scratch[(s * DW_SLOTS + i * 4 + j) * items + item] = acc[i][j];
                if (kg == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) scratch[(s * DW_SLOTS + 16 + j) * items + item] = bsum[j];
                }
            }
            __syncthreads();
            for (int o = threadIdx.x; o < DW_SLOTS * items; o += NT) {
                const int slot = o / items, im = o - slot * items;
                const int kg2 = im / NG, ng2 = im - kg2 * NG;
                if (slot < 16) {
                    const int i = slot >> 2, j = slot & 3;
                    const int k = kg2 + i * KG, n = ng2 + j * NG;
                    if (k >= K || n >= N) continue;
                    float sum = 0.0f;
                    for (int q = 0; q < S; ++q) sum += scratch[(q * DW_SLOTS + slot) * items + im];
                    G[k * ldw + ng2 * 4 + j] += sum;
                } else {
                    const int n = ng2 + (slot - 16) * NG;
                    if (kg2 != 0 || n >= N) continue;
                    float sum = 0.0f;
                    for (int q = 0; q < S; ++q) sum += scratch[(q * DW_SLOTS + slot) * items + im];
                    Gb[n] += sum;
                }
            }
            __syncthreads();
        }
    }
}

// G[n] += sum_r Z[n][r]   one warp per row of Z, warp-shuffle reduction (logstd gradient)
template <int NT>
__device__ void tile_rowsum(const float *__restrict__ Z, int N, float *G, int tb, int tbp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int n = warp; n < N; n += NT / 32) {
        float s = 0.0f;
        for (int r = lane * 4; r < tb; r += 128) {
            const float4 v = *reinterpret_cast<const float4 *>(Z + n * tbp + r);
            s += (v.x + v.y) + (v.z + v.w);
        }
        s = mq_warp_sum(s);
        if (lane == 0) G[n] += s;
    }
}

// Loss-gradient rule on the output tile (one thread per row); all inputs are already in
// shared memory (aux).  Writes d(loss)/d(pre-activation of the last layer) to dzout and the
// per-row logstd terms to aux.hterm.
template <int NT>
__device__ void head_tile(const FusedPlan &p, const mq_head_t &h, const float *sW, const float *sAout,
                          float *aux, float *dzout, int64_t row0, int64_t n_rows,
                          float *__restrict__ out, float *__restrict__ logp_out) {
    const int N = p.net.n_out, tbp = p.tbp;
    const int lastact = p.net.act[p.net.n_layers - 1];
    const float lastleak = p.net.leak[p.net.n_layers - 1];
    const float *tgt = aux;
    float *hterm = aux + N * tbp;
    const float *misc = aux + 2 * N * tbp;
    for (int r = threadIdx.x; r < p.tb; r += NT) {
        const int64_t gr = row0 + r;
        const bool valid = gr < n_rows;
        if (valid && out)
            for (int n = 0; n < N; ++n) out[gr * N + n] = sAout[n * tbp + r];
        if (!valid) {
            for (int n = 0; n < N; ++n) { dzout[n * tbp + r] = 0.0f; hterm[n * tbp + r] = 0.0f; }
            continue;
        }
        if (h.kind == MQ_HEAD_DY) {
            for (int n = 0; n < N; ++n)
                dzout[n * tbp + r] = mq_act_grad(sAout[n * tbp + r], tgt[n * tbp + r], lastact, lastleak);
        } else if (h.kind == MQ_HEAD_DIFF) {
            for (int n = 0; n < N; ++n) {
                const float y = sAout[n * tbp + r];
                dzout[n * tbp + r] = mq_act_grad(y, tgt[n * tbp + r] - y, lastact, lastleak);
            }
        } else if (h.kind == MQ_HEAD_SOFTMAX_CE) {
            float mx = sAout[r];
            for (int n = 1; n < N; ++n) mx = fmaxf(mx, sAout[n * tbp + r]);
            float sum = 0.0f;
            for (int n = 0; n < N; ++n) sum += expf(sAout[n * tbp + r] - mx);
            const float inv = 1.0f / sum;
            for (int n = 0; n < N; ++n) {
                const float y = sAout[n * tbp + r];
                const float dy = tgt[n * tbp + r] - expf(y - mx) * inv - h.p0 * y;
                dzout[n * tbp + r] = mq_act_grad(y, dy, lastact, lastleak);
            }
        } else if (mq_head_discrete(h.kind)) {
            // categorical log-prob (distrib.py:24-32): target row is the one-hot action
            float mx = sAout[r];
            for (int n = 1; n < N; ++n) mx = fmaxf(mx, sAout[n * tbp + r]);
            float sum = 0.0f;
            for (int n = 0; n < N; ++n) sum += expf(sAout[n * tbp + r] - mx);
            const float lse = mx + logf(sum);
            float lp = 0.0f;
            for (int n = 0; n < N; ++n) lp += tgt[n * tbp + r] * (sAout[n * tbp + r] - lse);
            if (logp_out) logp_out[gr] = lp;
            float w;
            if (h.kind == MQ_HEAD_DISCRETE_PPO) {
                const float adv = misc[r];
                const float ratio = expf(lp - misc[tbp + r]);
                const bool kill = (ratio > h.p1 && adv > 0.0f) || (ratio < h.p0 && adv < 0.0f);
                w = kill ? 0.0f : ratio * adv;
            } else {
                w = misc[r];
            }
            for (int n = 0; n < N; ++n) {
                const float y = sAout[n * tbp + r];
                dzout[n * tbp + r] = mq_act_grad(y, w * (tgt[n * tbp + r] - expf(y - lse)), lastact, lastleak);
                hterm[n * tbp + r] = 0.0f;
            }
        } else {
            // diagonal Gaussian: distrib.py:61-68 and its closed-form derivative
            float acc = 0.0f;
            for (int a = 0; a < N; ++a) {
                const float ls = sW[p.sls_off + a];
                const float z = (tgt[a * tbp + r] - sAout[a * tbp + r]) * expf(-ls);
                acc += ls + 0.5f * z * z;
            }
            const float lp = -(float)N * HALF_LOG_2PI - acc;
            if (logp_out) logp_out[gr] = lp;
            float w;
            if (h.kind == MQ_HEAD_GAUSS_PPO) {
                const float adv = misc[r];
                const float ratio = expf(lp - misc[tbp + r]);
                const bool kill = (ratio > h.p1 && adv > 0.0f) || (ratio < h.p0 && adv < 0.0f);
                w = kill ? 0.0f : ratio * adv;
            } else {
                w = misc[r];
            }
            for (int a = 0; a < N; ++a) {
                const float inv = expf(-sW[p.sls_off + a]);
                const float y = sAout[a * tbp + r];
                const float z = (tgt[a * tbp + r] - y) * inv;
                hterm[a * tbp + r] = w * (z * z - 1.0f) + h.ent;
                dzout[a * tbp + r] = mq_act_grad(y, w * z * inv, lastact, lastleak);
            }
        }
    }
}

// Backward over one tile.  On entry dZ of the last layer is in DZ buffer 0.
template <int NT>
__device__ void backward_tile(const FusedPlan &p, const Smem &sm, bool gauss, PhaseProf &pp) {
    const int L = p.net.n_layers;
    float *sA = sm.A, *sG = sm.G;
    if (gauss) tile_rowsum<NT>(sA + p.aux_off + p.net.n_out * p.tbp, p.net.n_out, sG + p.sls_off, p.tb, p.tbp);
    int cur = 0;
    for (int l = L - 1; l >= 0; --l) {
        const float *dZ = sm.DZ + cur * p.dz_floats;
        if (l > 0) {
            // dZ_{l-1} = (dZ_l W_l^T) * act'(A_l)  -> other ping-pong buffer (no barrier vs dW_l)
            tile_gemm_any<NT, true>(dZ, p.net.out[l], sm.Wt + p.swt_off[l], p.ldwt[l], p.net.in[l],
                                    p.kpad[l] >> 2, sm.DZ + (cur ^ 1) * p.dz_floats, sA + p.act_off[l],
                                    p.tb, p.tbp, nullptr, p.net.act[l - 1], p.net.leak[l - 1]);
            PHASE_MARK(pp, 10 + 3 * l);
        }
        tile_dw<NT>(sA + p.act_off[l], p.net.in[l], dZ, p.net.out[l], sG + p.sw_off[l], sG + p.sb_off[l],
                    p.npad[l], p.tb, p.tbp, sm.S);
        __syncthreads();
        PHASE_MARK(pp, 9 + 3 * l);
        cur ^= 1;
    }
}

// ---- kernels ---------------------------------------------------------------------------
static long long *g_phase_prof = nullptr;
extern "C" int mq_debug_set_phase_buffer(void *dev_buf) { g_phase_prof = (long long *)dev_buf; return MQ_OK; }
long long *mq_phase_prof_ptr() { return g_phase_prof; }

__global__ void __launch_bounds__(MQ_NT)
fused_forward_kernel(FusedPlan p, const float *__restrict__ theta, const float *__restrict__ x,
                     const int32_t *__restrict__ idx, int64_t n_rows, float *__restrict__ out,
                     const float *__restrict__ sample, float *__restrict__ logp) {
    MQ_SHARED_BYTES(smem);
    Smem sm(p, smem);
    float *sW = sm.W, *sA = sm.A;
    load_weights<MQ_NT>(p, theta, nullptr, sW, nullptr);
    const int64_t ntiles = (n_rows + p.tb - 1) / p.tb;
    const int N = p.net.n_out, tbp = p.tbp;
    PhaseProf pp = {nullptr, 0};
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * p.tb;
        load_rows<MQ_NT>(p, x, idx, row0, n_rows, sA + p.act_off[0]);
        __syncthreads();
        forward_tile<MQ_NT>(p, sW, sA, pp);
        const float *sO = sA + p.act_off[p.net.n_layers];
        for (int r = threadIdx.x; r < p.tb; r += MQ_NT) {
            const int64_t gr = row0 + r;
            if (gr >= n_rows) continue;
            if (out) for (int n = 0; n < N; ++n) out[gr * N + n] = sO[n * tbp + r];
            if (logp && p.net.logstd_off < 0) {          // categorical: sample row is the one-hot action
                const int64_t src = idx ? (int64_t)idx[gr] : gr;
                float mx = sO[r];
                for (int n = 1; n < N; ++n) mx = fmaxf(mx, sO[n * tbp + r]);
                float sum = 0.0f;
                for (int n = 0; n < N; ++n) sum += expf(sO[n * tbp + r] - mx);
                const float lse = mx + logf(sum);
                float lp = 0.0f;
                for (int n = 0; n < N; ++n) lp += sample[src * N + n] * (sO[n * tbp + r] - lse);
                logp[gr] = lp;
            } else if (logp) {
                const int64_t src = idx ? (int64_t)idx[gr] : gr;
                float acc = 0.0f;
                for (int a = 0; a < N; ++a) {
                    const float ls = sW[p.sls_off + a];
                    const float z = (sample[src * N + a] - sO[a * tbp + r]) * expf(-ls);
                    acc += ls + 0.5f * z * z;
                }
                logp[gr] = -(float)N * HALF_LOG_2PI - acc;
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(MQ_NT)
fused_grad_kernel(FusedPlan p, const float *__restrict__ theta, const float *__restrict__ x,
                  const int32_t *__restrict__ idx, int64_t n_rows, mq_head_t head, float scale,
                  float *__restrict__ grad_or_partial, int direct, float *__restrict__ out,
                  float *__restrict__ logp) {
    MQ_SHARED_BYTES(smem);
    Smem sm(p, smem);
    load_weights<MQ_NT>(p, theta, nullptr, sm.W, sm.Wt);
    for (int e = threadIdx.x; e < p.p_pad; e += MQ_NT) sm.G[e] = 0.0f;
    const bool gauss = head.kind == MQ_HEAD_GAUSS_LOGP || head.kind == MQ_HEAD_GAUSS_PPO;
    const int64_t ntiles = (n_rows + p.tb - 1) / p.tb;
    PhaseProf pp = {nullptr, 0};
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * p.tb;
        load_rows<MQ_NT>(p, x, idx, row0, n_rows, sm.A + p.act_off[0]);
        load_head_inputs<MQ_NT>(p, head, idx, row0, n_rows, sm.A + p.aux_off);
        __syncthreads();
        forward_tile<MQ_NT>(p, sm.W, sm.A, pp);
        head_tile<MQ_NT>(p, head, sm.W, sm.A + p.act_off[p.net.n_layers], sm.A + p.aux_off, sm.DZ, row0,
                         n_rows, out, logp);
        __syncthreads();
        backward_tile<MQ_NT>(p, sm, gauss, pp);
    }
    float *dst = direct ? grad_or_partial : grad_or_partial + (int64_t)blockIdx.x * p.net.n_params;
    const float sc = direct ? scale : 1.0f;
    const float *sG = sm.G;
    for_each_param<MQ_NT>(p, [&](int flat, int pad) { dst[flat] = sG[pad] * sc; });
}

// grad[j] = scale * sum_b partial[b][j]: 32 parameters x 8 groups of CTAs per block, group sums
// combined in fixed order (deterministic), 4 loads in flight per thread.
__global__ void __launch_bounds__(MQ_NT)
fused_reduce_kernel(const float *__restrict__ partial, int nblocks, int n_params, float scale,
                    float *__restrict__ grad) {
    __shared__ float part[8][33];
    const int jx = threadIdx.x & 31, gy = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + jx;
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    if (j < n_params) {
        int b = gy;
        for (; b + 24 < nblocks; b += 32) {
            s0 += partial[(int64_t)b * n_params + j];
            s1 += partial[(int64_t)(b + 8) * n_params + j];
            s2 += partial[(int64_t)(b + 16) * n_params + j];
            s3 += partial[(int64_t)(b + 24) * n_params + j];
        }
        for (; b < nblocks; b += 8) s0 += partial[(int64_t)b * n_params + j];
    }
    part[gy][jx] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (gy == 0 && j < n_params) {
        float s = part[0][jx];
#pragma unroll
        for (int k = 1; k < 8; ++k) s += part[k][jx];
        grad[j] = s * scale;
    }
}

template <typename ST> struct AdamMath;
template <> struct AdamMath<double> {
    static __device__ __forceinline__ void step(double &val, double &m, double &v, double g, double c1,
                                                double c2, double lr, double eps) {
        m = __dadd_rn(m, __dmul_rn(c1, __dsub_rn(g, m)));
        v = __dadd_rn(v, __dmul_rn(c2, __dsub_rn(__dmul_rn(g, g), v)));
        val = __dadd_rn(val, __dmul_rn(lr, __ddiv_rn(m, __dadd_rn(eps, __dsqrt_rn(v)))));
    }
};
template <> struct AdamMath<float> {
    static __device__ __forceinline__ void step(float &val, float &m, float &v, float g, float c1,
                                                float c2, float lr, float eps) {
        m += c1 * (g - m);
        v += c2 * (g * g - v);
        val += lr * (m / (eps + sqrtf(v)));
    }
};

// One CTA runs n_steps dependent optimiser steps: weights, their transposes, the gradient
// accumulators and (fp32 state) the Adam moments stay in shared memory for the whole launch.
template <typename ST, bool ONCHIP, int NT>
__global__ void __launch_bounds__(NT)
fused_train_kernel(FusedPlan p, float *__restrict__ theta, ST *__restrict__ value, ST *__restrict__ am,
                   ST *__restrict__ av, const float *__restrict__ x, mq_head_t head,
                   const int32_t *__restrict__ idx_table, int n_steps, int mb, double beta1,
                   double beta2, double tw1, double tw2, double lr, double eps, long long *prof) {
    MQ_SHARED_BYTES(smem);
    Smem sm(p, smem);
    float *sW = sm.W, *sG = sm.G, *sM = sm.M, *sV = sm.V;
    PhaseProf pp = {prof, 0};
#ifndef MQ_EMU
    if (prof && threadIdx.x == 0) pp.last = clock64();
#endif
    load_weights<NT>(p, theta, nullptr, sW, sm.Wt);
    for (int e = threadIdx.x; e < p.p_pad; e += NT) sG[e] = 0.0f;
    if (ONCHIP) {
        for (int e = threadIdx.x; e < 2 * p.p_pad; e += NT) sM[e] = 0.0f;
        __syncthreads();
        for_each_param<NT>(p, [&](int flat, int pad) {
            sW[pad] = (float)value[flat]; sM[pad] = (float)am[flat]; sV[pad] = (float)av[flat];
        });
        __syncthreads();
        refresh_transposes<NT>(p, sW, sm.Wt);
    }
    const bool gauss = head.kind == MQ_HEAD_GAUSS_LOGP || head.kind == MQ_HEAD_GAUSS_PPO;
    const int ntiles = (mb + p.tb - 1) / p.tb;
    const float inv_mb = 1.0f / (float)mb;
    __syncthreads();
    for (int step = 0; step < n_steps; ++step) {
        const int32_t *idx = idx_table + (int64_t)step * mb;
        for (int tile = 0; tile < ntiles; ++tile) {
            const int64_t row0 = (int64_t)tile * p.tb;
            load_rows<NT>(p, x, idx, row0, mb, sm.A + p.act_off[0]);
            load_head_inputs<NT>(p, head, idx, row0, mb, sm.A + p.aux_off);
            __syncthreads();
            PHASE_MARK(pp, 0);
            forward_tile<NT>(p, sW, sm.A, pp);
            head_tile<NT>(p, head, sW, sm.A + p.act_off[p.net.n_layers], sm.A + p.aux_off, sm.DZ, row0, mb,
                          nullptr, nullptr);
            __syncthreads();
            PHASE_MARK(pp, 2);
            backward_tile<NT>(p, sm, gauss, pp);
        }
        tw1 = tw1 * beta1 + 1.0;                       // _normalize.py:18-23 with weight 1
        tw2 = tw2 * beta2 + 1.0;
        const ST c1 = (ST)(1.0 / tw1), c2 = (ST)(1.0 / tw2);
        if (ONCHIP) {
            // fp32 state: moments in shared memory, the value IS the weight copy sW.  Straight
            // loop over the padded layout (padding stays 0: g = m = v = 0).
            for (int e = threadIdx.x; e < p.p_pad; e += NT) {
                const float g = sG[e] * inv_mb;        // batch mean (basicnet.py:242-249)
                sG[e] = 0.0f;
                float val = sW[e], m = sM[e], v = sV[e];
                AdamMath<float>::step(val, m, v, g, (float)c1, (float)c2, (float)lr, (float)eps);
                sW[e] = val; sM[e] = m; sV[e] = v;
            }
        } else {
            // state in global memory (L2 resident): issue the loads of 4 elements together
            const int P = p.net.n_params;
            for (int j0 = threadIdx.x; j0 < P; j0 += 4 * NT) {
                ST val[4], m[4], v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u * NT;
                    if (j < P) { val[u] = value[j]; m[u] = am[j]; v[u] = av[j]; }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int j = j0 + u * NT;
                    if (j >= P) continue;
                    const int pad = flat_to_pad(p, j);
                    const float g = sG[pad] * inv_mb;
                    sG[pad] = 0.0f;
                    AdamMath<ST>::step(val[u], m[u], v[u], (ST)g, c1, c2, (ST)lr, (ST)eps);
                    value[j] = val[u]; am[j] = m[u]; av[j] = v[u];
                    const float w = (float)val[u];     // load_params: fp64 -> fp32 (basicnet.py:56)
                    theta[j] = w;
                    sW[pad] = w;
                }
            }
        }
        __syncthreads();
        PHASE_MARK(pp, 4);
        refresh_transposes<NT>(p, sW, sm.Wt);
        __syncthreads();
        PHASE_MARK(pp, 5);
    }
    if (ONCHIP) {
        for_each_param<NT>(p, [&](int flat, int pad) {
            value[flat] = (ST)sW[pad]; am[flat] = (ST)sM[pad]; av[flat] = (ST)sV[pad];
            theta[flat] = sW[pad];
        });
    }
}

// One CTA per population member: theta + noise[p] is built in shared memory, the shared
// observation batch streams through it, the return is reduced on chip.
__global__ void __launch_bounds__(MQ_NT)
es_eval_kernel(FusedPlan p, const float *__restrict__ theta, const float *__restrict__ noise,
               int64_t n_members, const float *__restrict__ obs, int64_t n_obs,
               const float *__restrict__ ret_coef, float *__restrict__ returns) {
    MQ_SHARED_BYTES(smem);
    __shared__ float red[MQ_NT / 32];
    Smem sm(p, smem);
    float *sW = sm.W, *sA = sm.A;
    const int N = p.net.n_out, tbp = p.tbp;
    const int64_t ntiles = (n_obs + p.tb - 1) / p.tb;
    PhaseProf pp = {nullptr, 0};
    for (int64_t member = blockIdx.x; member < n_members; member += gridDim.x) {
        load_weights<MQ_NT>(p, theta, noise + member * p.net.n_params, sW, nullptr);
        float acc = 0.0f;
        for (int64_t tile = 0; tile < ntiles; ++tile) {
            const int64_t row0 = tile * p.tb;
            load_rows<MQ_NT>(p, obs, nullptr, row0, n_obs, sA + p.act_off[0]);
            __syncthreads();
            forward_tile<MQ_NT>(p, sW, sA, pp);
            const float *sO = sA + p.act_off[p.net.n_layers];
            for (int r = threadIdx.x; r < p.tb; r += MQ_NT)
                if (row0 + r < n_obs)
                    for (int a = 0; a < N; ++a) acc = fmaf(sO[a * tbp + r], ret_coef[a], acc);
            __syncthreads();
        }
        acc = mq_block_sum(acc, red);
        if (threadIdx.x == 0) returns[member] = acc;
        __syncthreads();
    }
}

// out[j] = scale * sum_p noise[p][j] * w[p]; blockIdx.y splits the members, fixed order
__global__ void __launch_bounds__(MQ_NT)
es_wsum_kernel(const float *__restrict__ noise, const double *__restrict__ w, int64_t n_members,
               int64_t n_params, double *__restrict__ partial) {
    const int64_t j = (int64_t)blockIdx.x * MQ_NT + threadIdx.x;
    if (j >= n_params) return;
    const int64_t per = (n_members + gridDim.y - 1) / gridDim.y;
    const int64_t p0 = (int64_t)blockIdx.y * per;
    const int64_t p1 = p0 + per < n_members ? p0 + per : n_members;
    double acc = 0.0;
    for (int64_t q = p0; q < p1; ++q) acc += (double)noise[q * n_params + j] * w[q];
    partial[(int64_t)blockIdx.y * n_params + j] = acc;
}
__global__ void __launch_bounds__(MQ_NT)
es_wsum_finish_kernel(const double *__restrict__ partial, int nsplit, int64_t n_params, double scale,
                      double *__restrict__ out) {
    for (int64_t j = (int64_t)blockIdx.x * MQ_NT + threadIdx.x; j < n_params; j += (int64_t)gridDim.x * MQ_NT) {
        double s = 0.0;
        for (int b = 0; b < nsplit; ++b) s += partial[(int64_t)b * n_params + j];
        out[j] = s * scale;
    }
}

// row-owner-warp persistent chain (mq_chain.cu); MQ_ERR_UNSUPPORTED if the net does not fit it
int mq_chain_train_f32(const mq_net_t *net, float *theta, float *value, float *m, float *v, const float *x,
                       const mq_head_t *head, const int32_t *idx_table, int32_t n_steps, int32_t mb,
                       double beta1, double beta2, const double *tw, double lr, double eps, void *stream);

int mq_chain_grad(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                  int64_t batch, const mq_head_t *head, float scale, float *grad, float *out, float *logp,
                  void *ws, int64_t ws_bytes, int *grid_out, void *stream);
int mq_tc_launch(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                 int64_t batch, const mq_head_t *head, int with_grad, float *out, float *logp,
                 void *ws, int64_t ws_bytes, int *grid_out, void *stream);
int mq_tc3_launch(const mq_net_t *net, const int32_t *row_idx, int64_t batch, const mq_head_t *head, const float *theta,
                  int with_grad, float *out, float *logp, void *ws, int64_t ws_bytes, int *grid_out, void *stream);      // mq_tc3.cu
int64_t mq_tc3_scratch_bytes();
int mq_tc_es_launch(const mq_net_t *net, const float *theta, const float *noise, int64_t n_members, const float *obs,
                    int64_t n_obs, const float *ret_coef, float *returns, void *stream);

// ---- host entry points -----------------------------------------------------------------
static int fused_net_check(const mq_net_t *net, const char *what) {
    MQ_REQUIRE(net, MQ_ERR_INVALID, "%s: null net", what);
    MQ_REQUIRE(net->n_layers >= 1 && net->n_layers <= MQ_MAX_LAYERS, MQ_ERR_SHAPE, "%s: n_layers", what);
    int k = net->n_in;
    for (int l = 0; l < net->n_layers; ++l) {
        MQ_REQUIRE(net->in[l] == k && net->out[l] >= 1, MQ_ERR_SHAPE, "%s: layer %d shape", what, l);
        MQ_REQUIRE(net->act[l] >= 0 && net->act[l] <= 2, MQ_ERR_INVALID, "%s: layer %d activation", what, l);
        k = net->out[l];
    }
    MQ_REQUIRE(k == net->n_out, MQ_ERR_SHAPE, "%s: n_out mismatch", what);
    return MQ_OK;
}

static int head_check(const mq_net_t *net, const mq_head_t *h, const char *what) {
    MQ_REQUIRE(h, MQ_ERR_INVALID, "%s: null head", what);
    switch (h->kind) {
        case MQ_HEAD_DY: MQ_REQUIRE(h->dy, MQ_ERR_INVALID, "%s: head DY needs dy", what); break;
        case MQ_HEAD_DIFF:
        case MQ_HEAD_SOFTMAX_CE: MQ_REQUIRE(h->target, MQ_ERR_INVALID, "%s: head needs target", what); break;
        case MQ_HEAD_GAUSS_LOGP:
            MQ_REQUIRE(h->target && h->dy && net->logstd_off >= 0, MQ_ERR_INVALID,
                       "%s: Gaussian head needs sample, weights and logstd parameters", what);
            break;
        case MQ_HEAD_GAUSS_PPO:
            MQ_REQUIRE(h->target && h->adv && h->logp_old && net->logstd_off >= 0, MQ_ERR_INVALID,
                       "%s: PPO head needs sample, adv, logp_old and logstd parameters", what);
            break;
        case MQ_HEAD_DISCRETE_LOGP:
            MQ_REQUIRE(h->target && h->dy && net->logstd_off < 0, MQ_ERR_INVALID,
                       "%s: categorical head needs one-hot actions and weights (and a net without logstd)", what);
            break;
        case MQ_HEAD_DISCRETE_PPO:
            MQ_REQUIRE(h->target && h->adv && h->logp_old && net->logstd_off < 0, MQ_ERR_INVALID,
                       "%s: categorical PPO head needs one-hot actions, adv and logp_old", what);
            break;
        default: mq_set_error("%s: unknown head kind %d", what, h->kind); return MQ_ERR_INVALID;
    }
    return MQ_OK;
}

#ifndef MQ_EMU
template <typename K>
static int set_smem(K kfn, int64_t bytes, const char *what) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        mq_set_error("%s: cannot reserve %lld bytes of shared memory (%s)", what, (long long)bytes,
                     cudaGetErrorString(e));
        return MQ_ERR_CUDA;
    }
    return MQ_OK;
}
#else
template <typename K> static int set_smem(K, int64_t, const char *) { return MQ_OK; }
#endif

extern "C" int mq_fused_tile_rows(const mq_net_t *net, int with_grad) {
    if (fused_net_check(net, "mq_fused_tile_rows") != MQ_OK) return 0;
    FusedPlan p;
    return plan_choose(net, 1 << 20, with_grad, &p);
}

// CTAs per launch: as many as can be co-resident (shared memory bound), never more than tiles.
// Tiles are strided by the grid size; bid and bid+148 land on the same SM, so the per-SM load
// stays balanced (ceil vs floor of tiles/148) and the summation order stays fixed.
static int64_t fused_grid(int64_t ntiles, int64_t smem_bytes) {
    const int64_t sms = mq_sm_count();
    int64_t per_sm = (228 * 1024) / (smem_bytes + 1024);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;
    const int64_t cap = sms * per_sm;
    return ntiles < cap ? (ntiles < 1 ? 1 : ntiles) : cap;
}

extern "C" int64_t mq_fused_ws_bytes(const mq_net_t *net, int64_t batch) {
    (void)batch;
    // per-CTA partial gradients (up to two CTAs per SM) + room for an operand image the tensor-core launch builds itself
    return (2 * (int64_t)mq_sm_count() + 1) * (net->n_params + 64) * (int64_t)sizeof(float) + 256 + mq_tc3_scratch_bytes();
}

extern "C" int mq_fused_forward_tc(const mq_net_t *net, const float *theta, const float *x,
                                   const int32_t *row_idx, int64_t batch, float *out, const float *sample,
                                   float *logp, int32_t tc, void *stream) {
    int rc = fused_net_check(net, "mq_fused_forward");
    if (rc != MQ_OK) return rc;
    MQ_REQUIRE(batch >= 0, MQ_ERR_SHAPE, "mq_fused_forward: negative batch");
    if (batch == 0) return MQ_OK;
    MQ_REQUIRE(theta && x, MQ_ERR_INVALID, "mq_fused_forward: null pointer");
    MQ_REQUIRE(!logp || sample, MQ_ERR_INVALID, "mq_fused_forward: logp needs sample");
    {   // large batches: tcgen05 / TMEM kernel (mq_tc.cu), forward + outputs only
        mq_head_t fh;
        memset(&fh, 0, sizeof(fh));
        fh.kind = !logp ? MQ_HEAD_DY : (net->logstd_off >= 0 ? MQ_HEAD_GAUSS_LOGP : MQ_HEAD_DISCRETE_LOGP);
        fh.target = sample;
        fh.tc = tc;
        rc = mq_tc_launch(net, theta, x, row_idx, batch, &fh, 0, out, logp, nullptr, 0, nullptr, stream);
        if (rc != MQ_ERR_UNSUPPORTED) return rc;
    }
    FusedPlan p;
    MQ_REQUIRE(plan_choose(net, batch, 0, &p) > 0, MQ_ERR_UNSUPPORTED,
               "mq_fused_forward: net does not fit in shared memory");
    rc = set_smem(fused_forward_kernel, p.smem_bytes, "mq_fused_forward");
    if (rc != MQ_OK) return rc;
    const int64_t grid = fused_grid(mq_ceil_div(batch, p.tb), p.smem_bytes);
    MQ_LAUNCH(fused_forward_kernel, (unsigned)grid, MQ_NT, p.smem_bytes, stream, p, theta, x, row_idx,
              batch, out, sample, logp);
    MQ_CHECK_LAUNCH("mq_fused_forward");
    return MQ_OK;
}

// Forward pass from PACKED RECORDS (mq_pack_records: x | target | w0 | w1 per row) on the record-fed tensor-core kernel
// (csrc/mq_tc3.cu, forward only): the two passes over the whole rollout of a PPO update -- value predictions, gae.py:43-45,
// and the old log-probabilities, ppo.py:27 -- in the operand format `tc` names, two tiles per SM in flight.  The sample of
// the log-probability is the record's target.  tc_image: operand image of the CURRENT parameters (mq_tc_image_build).
// MQ_ERR_UNSUPPORTED when the net does not fit that kernel (callers then use mq_fused_forward on the separate arrays).
extern "C" int mq_fused_forward_rec(const mq_net_t *net, const float *theta, const float *records, int32_t rec_w,
                                    const void *tc_image, const int32_t *row_idx, int64_t batch, float *out, float *logp,
                                    int32_t tc, void *stream) {
    int rc = fused_net_check(net, "mq_fused_forward_rec");
    if (rc != MQ_OK) return rc;
    MQ_REQUIRE(batch >= 0, MQ_ERR_SHAPE, "mq_fused_forward_rec: negative batch");
    if (batch == 0) return MQ_OK;
    MQ_REQUIRE(records && tc_image && rec_w > 0 && (out || logp), MQ_ERR_INVALID, "mq_fused_forward_rec: null pointer");
    mq_head_t fh;
    memset(&fh, 0, sizeof(fh));
    fh.kind = !logp ? MQ_HEAD_DY : (net->logstd_off >= 0 ? MQ_HEAD_GAUSS_LOGP : MQ_HEAD_DISCRETE_LOGP);
    fh.records = records; fh.rec_w = rec_w; fh.tc_image = tc_image; fh.tc = tc;
    return mq_tc3_launch(net, row_idx, batch, &fh, theta, 0, out, logp, nullptr, 0, nullptr, stream);
}

extern "C" int mq_fused_forward(const mq_net_t *net, const float *theta, const float *x,
                                const int32_t *row_idx, int64_t batch, float *out, const float *sample,
                                float *logp, void *stream) {
    return mq_fused_forward_tc(net, theta, x, row_idx, batch, out, sample, logp, 0, stream);
}

extern "C" int mq_fused_grad(const mq_net_t *net, const float *theta, const float *x,
                             const int32_t *row_idx, int64_t batch, const mq_head_t *head, float scale,
                             float *grad, float *out, float *logp, void *ws, int64_t ws_bytes,
                             void *stream) {
    int rc = fused_net_check(net, "mq_fused_grad");
    if (rc != MQ_OK) return rc;
    rc = head_check(net, head, "mq_fused_grad");
    if (rc != MQ_OK) return rc;
    MQ_REQUIRE(batch >= 1, MQ_ERR_SHAPE, "mq_fused_grad: empty batch");
    MQ_REQUIRE(theta && x && grad, MQ_ERR_INVALID, "mq_fused_grad: null pointer");
    {   // large batches: tcgen05 / TMEM kernel (mq_tc.cu)
        int tgrid = 0;
        rc = mq_tc_launch(net, theta, x, row_idx, batch, head, 1, out, logp, ws, ws_bytes, &tgrid, stream);
        if (rc == MQ_OK) {
            MQ_LAUNCH(fused_reduce_kernel, (unsigned)mq_ceil_div(net->n_params, 32), MQ_NT, 0, stream,
                      (const float *)ws, tgrid, net->n_params, scale, grad);
            MQ_CHECK_LAUNCH("mq_fused_grad");
            return MQ_OK;
        }
        if (rc != MQ_ERR_UNSUPPORTED) return rc;
    }
    {
        int cgrid = 0;
        rc = mq_chain_grad(net, theta, x, row_idx, batch, head, scale, grad, out, logp, ws, ws_bytes, &cgrid,
                           stream);
        if (rc == MQ_OK) {
            if (cgrid > 1) {
                MQ_LAUNCH(fused_reduce_kernel, (unsigned)mq_ceil_div(net->n_params, 32), MQ_NT, 0, stream,
                          (const float *)ws, cgrid, net->n_params, scale, grad);
                MQ_CHECK_LAUNCH("mq_fused_grad");
            }
            return MQ_OK;
        }
        if (rc != MQ_ERR_UNSUPPORTED) return rc;
    }
    FusedPlan p;
    MQ_REQUIRE(plan_choose(net, batch, 1, &p) > 0, MQ_ERR_UNSUPPORTED,
               "mq_fused_grad: net does not fit in shared memory");
    rc = set_smem(fused_grad_kernel, p.smem_bytes, "mq_fused_grad");
    if (rc != MQ_OK) return rc;
    const int64_t grid = fused_grid(mq_ceil_div(batch, p.tb), p.smem_bytes);
    if (grid == 1) {
        MQ_LAUNCH(fused_grad_kernel, 1, MQ_NT, p.smem_bytes, stream, p, theta, x, row_idx, batch, *head,
                  scale, grad, 1, out, logp);
    } else {
        MQ_REQUIRE(ws && ws_bytes >= grid * net->n_params * (int64_t)sizeof(float), MQ_ERR_SHAPE,
                   "mq_fused_grad: workspace too small");
        MQ_LAUNCH(fused_grad_kernel, (unsigned)grid, MQ_NT, p.smem_bytes, stream, p, theta, x, row_idx,
                  batch, *head, scale, (float *)ws, 0, out, logp);
        MQ_LAUNCH(fused_reduce_kernel, (unsigned)mq_ceil_div(net->n_params, 32), MQ_NT, 0, stream,
                  (const float *)ws, (int)grid, net->n_params, scale, grad);
    }
    MQ_CHECK_LAUNCH("mq_fused_grad");
    return MQ_OK;
}

extern "C" int mq_fused_grad_partials(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                                      int64_t batch, const mq_head_t *head, float *out, float *logp, void *ws,
                                      int64_t ws_bytes, int32_t *n_parts_out, void *stream) {
    int rc = fused_net_check(net, "mq_fused_grad_partials");
    if (rc != MQ_OK) return rc;
    rc = head_check(net, head, "mq_fused_grad_partials");
    if (rc != MQ_OK) return rc;
    MQ_REQUIRE(batch >= 1, MQ_ERR_SHAPE, "mq_fused_grad_partials: empty batch");
    MQ_REQUIRE(theta && x && ws && n_parts_out, MQ_ERR_INVALID, "mq_fused_grad_partials: null pointer");
    int tgrid = 0;
    rc = mq_tc_launch(net, theta, x, row_idx, batch, head, 1, out, logp, ws, ws_bytes, &tgrid, stream);
    if (rc == MQ_OK) { *n_parts_out = tgrid; return MQ_OK; }
    if (rc != MQ_ERR_UNSUPPORTED) return rc;
    // other kernels: one already summed, unscaled vector at the head of the workspace
    const int64_t head_bytes = mq_round_up((int64_t)net->n_params * (int64_t)sizeof(float), 256);
    MQ_REQUIRE(ws_bytes > head_bytes, MQ_ERR_INVALID, "mq_fused_grad_partials: workspace too small");
    rc = mq_fused_grad(net, theta, x, row_idx, batch, head, 1.0f, (float *)ws, out, logp, (char *)ws + head_bytes,
                       ws_bytes - head_bytes, stream);
    if (rc == MQ_OK) *n_parts_out = 1;
    return rc;
}

extern "C" int mq_fused_train(const mq_net_t *net, float *theta, void *value, void *m, void *v,
                              int state_dtype, const float *x, const mq_head_t *head,
                              const int32_t *idx_table, int32_t n_steps, int32_t mb, double beta1,
                              double beta2, double *tw_host, double lr, double eps, void *stream) {
    int rc = fused_net_check(net, "mq_fused_train");
    if (rc != MQ_OK) return rc;
    rc = head_check(net, head, "mq_fused_train");
    if (rc != MQ_OK) return rc;
    MQ_REQUIRE(n_steps >= 0 && mb >= 1, MQ_ERR_SHAPE, "mq_fused_train: bad step/minibatch count");
    if (n_steps == 0) return MQ_OK;
    MQ_REQUIRE(theta && value && m && v && x && idx_table && tw_host, MQ_ERR_INVALID,
               "mq_fused_train: null pointer");
    FusedPlan p;
#define MQ_TRAIN_GO(ST, ONCHIP)                                                                      \
    do {                                                                                             \
        auto kfn = fused_train_kernel<ST, ONCHIP, TRAIN_NT>;                                         \
        rc = set_smem(kfn, p.smem_bytes, "mq_fused_train");                                          \
        if (rc != MQ_OK) return rc;                                                                  \
        MQ_LAUNCH(kfn, 1, TRAIN_NT, p.smem_bytes, stream, p, theta, (ST *)value, (ST *)m, (ST *)v,   \
                  x, *head, idx_table, (int)n_steps, (int)mb, beta1, beta2, tw_host[0], tw_host[1],  \
                  lr, eps, g_phase_prof);                                                            \
    } while (0)
    if (state_dtype == MQ_F64) {
        MQ_REQUIRE(plan_choose(net, mb, 1, &p, 0, TRAIN_NT) > 0, MQ_ERR_UNSUPPORTED,
                   "mq_fused_train: net does not fit in shared memory");
        MQ_TRAIN_GO(double, false);
    } else if (state_dtype == MQ_F32) {
        rc = mq_chain_train_f32(net, theta, (float *)value, (float *)m, (float *)v, x, head, idx_table,
                                n_steps, mb, beta1, beta2, tw_host, lr, eps, stream);
        if (rc != MQ_OK && rc != MQ_ERR_UNSUPPORTED) return rc;
        if (rc == MQ_OK) {
            // launched by the row-owner-warp kernel
        } else if (plan_choose(net, mb, 1, &p, 1, TRAIN_NT) > 0) {
            MQ_TRAIN_GO(float, true);
        } else {
            MQ_REQUIRE(plan_choose(net, mb, 1, &p, 0, TRAIN_NT) > 0, MQ_ERR_UNSUPPORTED,
                       "mq_fused_train: net does not fit in shared memory");
            MQ_TRAIN_GO(float, false);
        }
    } else {
        mq_set_error("mq_fused_train: unknown state dtype %d", state_dtype);
        return MQ_ERR_INVALID;
    }
#undef MQ_TRAIN_GO
    MQ_CHECK_LAUNCH("mq_fused_train");
    // the total-weight recurrence is data independent: mirror it on the host
    for (int s = 0; s < n_steps; ++s) {
        tw_host[0] = tw_host[0] * beta1 + 1.0;
        tw_host[1] = tw_host[1] * beta2 + 1.0;
    }
    return MQ_OK;
}

extern "C" int mq_es_eval(const mq_net_t *net, const float *theta, const float *noise,
                          int64_t n_members, const float *obs, int64_t n_obs, const float *ret_coef,
                          float *returns, void *stream) {
    int rc = fused_net_check(net, "mq_es_eval");
    if (rc != MQ_OK) return rc;
    MQ_REQUIRE(n_members >= 0 && n_obs >= 1, MQ_ERR_SHAPE, "mq_es_eval: bad sizes");
    if (n_members == 0) return MQ_OK;
    MQ_REQUIRE(theta && noise && obs && ret_coef && returns, MQ_ERR_INVALID, "mq_es_eval: null pointer");
    {   // long observation batches: population evaluation on the tensor pipe (mq_tc.cu)
        const int trc = mq_tc_es_launch(net, theta, noise, n_members, obs, n_obs, ret_coef, returns, stream);
        if (trc != MQ_ERR_UNSUPPORTED) return trc;
    }
    FusedPlan p;
    MQ_REQUIRE(plan_choose(net, n_obs, 0, &p) > 0, MQ_ERR_UNSUPPORTED,
               "mq_es_eval: net does not fit in shared memory");
    rc = set_smem(es_eval_kernel, p.smem_bytes, "mq_es_eval");
    if (rc != MQ_OK) return rc;
    int64_t grid = n_members;
    const int64_t cap = (int64_t)mq_sm_count() * 8;
    if (grid > cap) grid = cap;
    MQ_LAUNCH(es_eval_kernel, (unsigned)grid, MQ_NT, p.smem_bytes, stream, p, theta, noise, n_members,
              obs, n_obs, ret_coef, returns);
    MQ_CHECK_LAUNCH("mq_es_eval");
    return MQ_OK;
}

static int es_splits(int64_t n_members) {
    int64_t s = mq_ceil_div(n_members, 64);
    if (s > 32) s = 32;
    return (int)(s < 1 ? 1 : s);
}

extern "C" int64_t mq_es_ws_bytes(int64_t n_members, int64_t n_params) {
    return (int64_t)es_splits(n_members) * n_params * (int64_t)sizeof(double) + 64;
}

extern "C" int mq_es_weighted_sum(const float *noise, const double *w, int64_t n_members,
                                  int64_t n_params, double scale, double *out, void *ws,
                                  int64_t ws_bytes, void *stream) {
    MQ_REQUIRE(n_members >= 1 && n_params >= 1, MQ_ERR_SHAPE, "mq_es_weighted_sum: bad sizes");
    MQ_REQUIRE(noise && w && out && ws, MQ_ERR_INVALID, "mq_es_weighted_sum: null pointer");
    MQ_REQUIRE(ws_bytes >= mq_es_ws_bytes(n_members, n_params), MQ_ERR_SHAPE,
               "mq_es_weighted_sum: workspace too small");
    const int ns = es_splits(n_members);
    MQ_LAUNCH(es_wsum_kernel, dim3((unsigned)mq_ceil_div(n_params, MQ_NT), (unsigned)ns), MQ_NT, 0, stream,
              noise, w, n_members, n_params, (double *)ws);
    MQ_LAUNCH(es_wsum_finish_kernel, (unsigned)mq_ceil_div(n_params, MQ_NT), MQ_NT, 0, stream,
              (const double *)ws, ns, n_params, scale, out);
    MQ_CHECK_LAUNCH("mq_es_weighted_sum");
    return MQ_OK;
}
