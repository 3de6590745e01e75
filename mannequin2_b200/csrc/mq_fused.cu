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
// Layout in shared memory (floats):
//   sW   weights, layer l as [in][out padded to 4] + bias + logstd     (natural order)
//   sWt  transposed copy of layers >= 1, [out][in padded to 4]         (for g W^T)
//   sG   gradient accumulators, same layout as sW
//   sA_l activations, TRANSPOSED: [width_l][tile_rows + 4]; rows contiguous so a thread
//        reads its 8 rows of one feature with two 128-bit loads, and dW (a contraction
//        over rows) reads both operands with unit stride.
// FFMA register tiles: forward / g W^T use (RM rows x 4 cols) per thread, dW uses 4x4
// over (in,out) with the row range split across thread groups when the layer is small;
// bias and logstd gradients are warp-shuffle row sums.  All reductions have a fixed
// order, so results are bit-reproducible run to run.
#include "mq_common.cuh"

#define HALF_LOG_2PI 0.91893853320467274178f
#define DW_SCRATCH_FLOATS (MQ_NT * 16)

struct FusedPlan {
    mq_net_t net;
    int tb, tbp;
    int npad[MQ_MAX_LAYERS], kpad[MQ_MAX_LAYERS];
    int sw_off[MQ_MAX_LAYERS], sb_off[MQ_MAX_LAYERS], swt_off[MQ_MAX_LAYERS];
    int sls_off;
    int p_pad, wt_total;
    int act_off[MQ_MAX_LAYERS + 1];
    int head_off, act_total;
    int with_grad;
    int64_t smem_bytes;
};

static int64_t plan_build(const mq_net_t *net, int tb, int with_grad, FusedPlan *p) {
    memset(p, 0, sizeof(*p));
    p->net = *net;
    p->tb = tb;
    p->tbp = tb + 4;
    p->with_grad = with_grad;
    int pos = 0, wt = 0;
    for (int l = 0; l < net->n_layers; ++l) {
        p->npad[l] = (int)mq_round_up(net->out[l], 4);
        p->kpad[l] = (int)mq_round_up(net->in[l], 4);
        p->sw_off[l] = pos; pos += net->in[l] * p->npad[l];
        p->sb_off[l] = pos; pos += p->npad[l];
        if (with_grad && l >= 1) { p->swt_off[l] = wt; wt += net->out[l] * p->kpad[l]; }
    }
    p->sls_off = pos;
    if (net->logstd_off >= 0) pos += (int)mq_round_up(net->n_out, 4);
    p->p_pad = pos;
    p->wt_total = wt;
    int a = 0;
    p->act_off[0] = a; a += net->n_in * p->tbp;
    for (int l = 0; l < net->n_layers; ++l) { p->act_off[l + 1] = a; a += net->out[l] * p->tbp; }
    p->head_off = a; a += net->n_out * p->tbp;
    p->act_total = a;
    int64_t floats = (int64_t)p->p_pad + p->act_total;
    if (with_grad) floats += (int64_t)p->wt_total + p->p_pad + DW_SCRATCH_FLOATS;
    p->smem_bytes = floats * (int64_t)sizeof(float);
    return p->smem_bytes;
}

static int plan_choose(const mq_net_t *net, int64_t batch, int with_grad, FusedPlan *p) {
    const int64_t lim = mq_smem_optin();
    int tb = 128;
    if (batch <= 32) tb = 32; else if (batch <= 64) tb = 64;
    for (; tb >= 32; tb >>= 1)
        if (plan_build(net, tb, with_grad, p) <= lim) return tb;
    return 0;
}

__device__ __forceinline__ float *smem_W(const FusedPlan &p, unsigned char *base) { return (float *)base; }
__device__ __forceinline__ float *smem_A(const FusedPlan &p, unsigned char *base) {
    return (float *)base + p.p_pad;
}
__device__ __forceinline__ float *smem_Wt(const FusedPlan &p, unsigned char *base) {
    return (float *)base + p.p_pad + p.act_total;
}
__device__ __forceinline__ float *smem_G(const FusedPlan &p, unsigned char *base) {
    return (float *)base + p.p_pad + p.act_total + p.wt_total;
}
__device__ __forceinline__ float *smem_S(const FusedPlan &p, unsigned char *base) {
    return (float *)base + p.p_pad + p.act_total + p.wt_total + p.p_pad;
}

// ---- weights: flat theta (+ optional per-member noise) -> padded shared layout --------
__device__ void load_weights(const FusedPlan &p, const float *__restrict__ theta,
                             const float *__restrict__ noise, float *sW, float *sWt) {
    const int t = threadIdx.x;
    for (int e = t; e < p.p_pad; e += MQ_NT) sW[e] = 0.0f;
    if (sWt) for (int e = t; e < p.wt_total; e += MQ_NT) sWt[e] = 0.0f;
    __syncthreads();
    for (int l = 0; l < p.net.n_layers; ++l) {
        const int K = p.net.in[l], N = p.net.out[l];
        for (int e = t; e < K * N; e += MQ_NT) {
            const int k = e / N, n = e - k * N;
            float w = theta[p.net.w_off[l] + e];
            if (noise) w += noise[p.net.w_off[l] + e];
            sW[p.sw_off[l] + k * p.npad[l] + n] = w;
            if (sWt && l >= 1) sWt[p.swt_off[l] + n * p.kpad[l] + k] = w;
        }
        for (int n = t; n < N; n += MQ_NT) {
            float b = theta[p.net.b_off[l] + n];
            if (noise) b += noise[p.net.b_off[l] + n];
            sW[p.sb_off[l] + n] = b;
        }
    }
    if (p.net.logstd_off >= 0)
        for (int a = t; a < p.net.n_out; a += MQ_NT) {
            float v = theta[p.net.logstd_off + a];
            if (noise) v += noise[p.net.logstd_off + a];
            sW[p.sls_off + a] = v;
        }
    __syncthreads();
}

// ---- input tile: x[row, :] (optionally gathered) -> sA0[k][r] -------------------------
__device__ void load_rows(const FusedPlan &p, const float *__restrict__ x,
                          const int32_t *__restrict__ idx, int64_t row0, int64_t n_rows, float *sA0) {
    const int K = p.net.n_in;
    for (int e = threadIdx.x; e < p.tb * K; e += MQ_NT) {
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

template <int RM> struct RowVec;
template <> struct RowVec<1> { static __device__ __forceinline__ void ld(const float *p, float *a) { a[0] = p[0]; } };
template <> struct RowVec<2> { static __device__ __forceinline__ void ld(const float *p, float *a) {
    float2 v = *reinterpret_cast<const float2 *>(p); a[0] = v.x; a[1] = v.y; } };
template <> struct RowVec<4> { static __device__ __forceinline__ void ld(const float *p, float *a) {
    float4 v = *reinterpret_cast<const float4 *>(p); a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; } };
template <> struct RowVec<8> { static __device__ __forceinline__ void ld(const float *p, float *a) {
    RowVec<4>::ld(p, a); RowVec<4>::ld(p + 4, a + 4); } };

// out[n][r] = epi( sum_k in[k][r] * w[k*ldw + n] ),  in: [K][tbp], out: [N][tbp]
//   BWD == false: out = act(acc + bias[n])
//   BWD == true : out = act'(out) * acc        (in place on the saved activation)
template <int RM, bool BWD>
__device__ void tile_gemm(const float *__restrict__ in, int K, const float *__restrict__ w, int ldw,
                          int N, float *out, int tb, int tbp, const float *bias, int act, float leak) {
    const int CG = (N + 3) >> 2;
    const int items = CG * (tb / RM);
    for (int item = threadIdx.x; item < items; item += MQ_NT) {
        const int cg = item % CG, rg = item / CG;
        const float *ip = in + rg * RM;
        const float *wp = w + cg * 4;
        float acc[RM][4];
#pragma unroll
        for (int i = 0; i < RM; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f; }
#pragma unroll 4
        for (int k = 0; k < K; ++k) {
            float a[RM];
            RowVec<RM>::ld(ip + k * tbp, a);
            const float4 wv = *reinterpret_cast<const float4 *>(wp + k * ldw);
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                acc[i][0] = fmaf(a[i], wv.x, acc[i][0]);
                acc[i][1] = fmaf(a[i], wv.y, acc[i][1]);
                acc[i][2] = fmaf(a[i], wv.z, acc[i][2]);
                acc[i][3] = fmaf(a[i], wv.w, acc[i][3]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = cg * 4 + j;
            if (n >= N) continue;
            float *op = out + n * tbp + rg * RM;
            const float bn = BWD ? 0.0f : bias[n];
#pragma unroll
            for (int i = 0; i < RM; ++i) {
                if (BWD) op[i] = mq_act_grad(op[i], acc[i][j], act, leak);
                else op[i] = mq_act_apply(acc[i][j] + bn, act, leak);
            }
        }
    }
}

template <bool BWD>
__device__ __forceinline__ void tile_gemm_any(const float *in, int K, const float *w, int ldw, int N,
                                              float *out, int tb, int tbp, const float *bias, int act,
                                              float leak) {
    const int CG = (N + 3) >> 2;
    int rm = 8;
    while (rm > 1 && (tb / rm) * CG < MQ_NT) rm >>= 1;
    if (rm == 8) tile_gemm<8, BWD>(in, K, w, ldw, N, out, tb, tbp, bias, act, leak);
    else if (rm == 4) tile_gemm<4, BWD>(in, K, w, ldw, N, out, tb, tbp, bias, act, leak);
    else if (rm == 2) tile_gemm<2, BWD>(in, K, w, ldw, N, out, tb, tbp, bias, act, leak);
    else tile_gemm<1, BWD>(in, K, w, ldw, N, out, tb, tbp, bias, act, leak);
}

__device__ void forward_tile(const FusedPlan &p, const float *sW, float *sA) {
    for (int l = 0; l < p.net.n_layers; ++l) {
        tile_gemm_any<false>(sA + p.act_off[l], p.net.in[l], sW + p.sw_off[l], p.npad[l], p.net.out[l],
                             sA + p.act_off[l + 1], p.tb, p.tbp, sW + p.sb_off[l], p.net.act[l],
                             p.net.leak[l]);
        __syncthreads();
    }
}

// G[k*ldw + n] += sum_r X[k][r] * dZ[n][r]     (4x4 register tile over (k,n))
__device__ void tile_dw(const float *__restrict__ X, int K, const float *__restrict__ dZ, int N,
                        float *G, int ldw, int tb, int tbp, float *scratch) {
    const int KG = (K + 3) >> 2, NG = (N + 3) >> 2;
    const int items = KG * NG;
    int S = 1;
    if (items < MQ_NT) {
        while (S * 2 * items <= MQ_NT && S * 2 <= tb / 4) S <<= 1;
    }
    const int span = tb / S;
    const int nloop = (S > 1) ? 1 : (items + MQ_NT - 1) / MQ_NT;
    for (int it = 0; it < nloop; ++it) {
        const int lin = threadIdx.x + it * MQ_NT;
        const int s = (S > 1) ? lin / items : 0;
        const int item = (S > 1) ? lin % items : lin;
        const bool active = (S > 1) ? (s < S) : (item < items);
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f; }
        const int kg = item / NG, ng = item - kg * NG;
        if (active) {
            const float *xp[4], *zp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = min(kg * 4 + i, K - 1), n = min(ng * 4 + i, N - 1);
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
                    const int k = kg * 4 + i;
                    if (k >= K) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = ng * 4 + j;
                        if (n < N) G[k * ldw + n] += acc[i][j];
                    }
                }
            }
        } else {
            if (active) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) scratch[(s * items + item) * 16 + i * 4 + j] = acc[i][j];
            }
            __syncthreads();
            if (active && s == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int k = kg * 4 + i;
                    if (k >= K) continue;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = ng * 4 + j;
                        if (n >= N) continue;
                        float sum = 0.0f;
                        for (int q = 0; q < S; ++q) sum += scratch[(q * items + item) * 16 + i * 4 + j];
                        G[k * ldw + n] += sum;
                    }
                }
            }
            __syncthreads();
        }
    }
}

// G[n] += sum_r Z[n][r]   one warp per row of Z, warp-shuffle reduction
__device__ void tile_rowsum(const float *__restrict__ Z, int N, float *G, int tb, int tbp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int n = warp; n < N; n += MQ_NT / 32) {
        float s = 0.0f;
        for (int r = lane * 4; r < tb; r += 128) {
            const float4 v = *reinterpret_cast<const float4 *>(Z + n * tbp + r);
            s += (v.x + v.y) + (v.z + v.w);
        }
        s = mq_warp_sum(s);
        if (lane == 0) G[n] += s;
    }
}

// Loss-gradient rule on the output tile (one thread per row).  On return sA_out holds
// d(loss)/d(pre-activation of the last layer); sH holds per-row logstd terms.
__device__ void head_tile(const FusedPlan &p, const mq_head_t &h, const float *sW, float *sAout,
                          float *sH, const int32_t *__restrict__ idx, int64_t row0, int64_t n_rows,
                          float *__restrict__ out, float *__restrict__ logp_out) {
    const int N = p.net.n_out, tbp = p.tbp;
    const int lastact = p.net.act[p.net.n_layers - 1];
    const float lastleak = p.net.leak[p.net.n_layers - 1];
    for (int r = threadIdx.x; r < p.tb; r += MQ_NT) {
        const int64_t gr = row0 + r;
        const bool valid = gr < n_rows;
        const int64_t src = (valid && idx) ? (int64_t)idx[gr] : gr;
        if (valid && out)
            for (int n = 0; n < N; ++n) out[gr * N + n] = sAout[n * tbp + r];
        if (!valid) {
            for (int n = 0; n < N; ++n) { sAout[n * tbp + r] = 0.0f; sH[n * tbp + r] = 0.0f; }
            continue;
        }
        if (h.kind == MQ_HEAD_DY) {
            for (int n = 0; n < N; ++n) {
                const float y = sAout[n * tbp + r];
                sAout[n * tbp + r] = mq_act_grad(y, h.dy[src * N + n], lastact, lastleak);
            }
        } else if (h.kind == MQ_HEAD_DIFF) {
            for (int n = 0; n < N; ++n) {
                const float y = sAout[n * tbp + r];
                sAout[n * tbp + r] = mq_act_grad(y, h.target[src * N + n] - y, lastact, lastleak);
            }
        } else if (h.kind == MQ_HEAD_SOFTMAX_CE) {
            float mx = sAout[r];
            for (int n = 1; n < N; ++n) mx = fmaxf(mx, sAout[n * tbp + r]);
            float sum = 0.0f;
            for (int n = 0; n < N; ++n) sum += expf(sAout[n * tbp + r] - mx);
            const float inv = 1.0f / sum;
            for (int n = 0; n < N; ++n) {
                const float y = sAout[n * tbp + r];
                const float dy = h.target[src * N + n] - expf(y - mx) * inv - h.p0 * y;
                sAout[n * tbp + r] = mq_act_grad(y, dy, lastact, lastleak);
            }
        } else {
            // diagonal Gaussian: distrib.py:61-68 and its closed-form derivative
            float acc = 0.0f;
            for (int a = 0; a < N; ++a) {
                const float ls = sW[p.sls_off + a];
                const float z = (h.target[src * N + a] - sAout[a * tbp + r]) * expf(-ls);
                acc += ls + 0.5f * z * z;
            }
            const float lp = -(float)N * HALF_LOG_2PI - acc;
            if (logp_out) logp_out[gr] = lp;
            float w;
            if (h.kind == MQ_HEAD_GAUSS_PPO) {
                const float adv = h.adv[src];
                const float ratio = expf(lp - h.logp_old[src]);
                const bool kill = (ratio > h.p1 && adv > 0.0f) || (ratio < h.p0 && adv < 0.0f);
                w = kill ? 0.0f : ratio * adv;
            } else {
                w = h.dy[src];
            }
            for (int a = 0; a < N; ++a) {
                const float inv = expf(-sW[p.sls_off + a]);
                const float y = sAout[a * tbp + r];
                const float z = (h.target[src * N + a] - y) * inv;
                sH[a * tbp + r] = w * (z * z - 1.0f);
                sAout[a * tbp + r] = mq_act_grad(y, w * z * inv, lastact, lastleak);
            }
        }
    }
}

__device__ void backward_tile(const FusedPlan &p, const float *sWt, float *sA, float *sG,
                              float *scratch, bool gauss) {
    const int L = p.net.n_layers;
    if (gauss) tile_rowsum(sA + p.head_off, p.net.n_out, sG + p.sls_off, p.tb, p.tbp);
    for (int l = L - 1; l >= 0; --l) {
        const float *dZ = sA + p.act_off[l + 1];
        tile_rowsum(dZ, p.net.out[l], sG + p.sb_off[l], p.tb, p.tbp);
        tile_dw(sA + p.act_off[l], p.net.in[l], dZ, p.net.out[l], sG + p.sw_off[l], p.npad[l], p.tb,
                p.tbp, scratch);
        __syncthreads();
        if (l > 0) {
            tile_gemm_any<true>(dZ, p.net.out[l], sWt + p.swt_off[l], p.kpad[l], p.net.in[l],
                                sA + p.act_off[l], p.tb, p.tbp, nullptr, p.net.act[l - 1],
                                p.net.leak[l - 1]);
            __syncthreads();
        }
    }
}

// flat[j] = f(padded accumulator) for every parameter j
template <typename F>
__device__ __forceinline__ void for_each_param(const FusedPlan &p, F f) {
    for (int l = 0; l < p.net.n_layers; ++l) {
        const int K = p.net.in[l], N = p.net.out[l];
        for (int e = threadIdx.x; e < K * N; e += MQ_NT) {
            const int k = e / N, n = e - k * N;
            f(p.net.w_off[l] + e, p.sw_off[l] + k * p.npad[l] + n, l, k, n);
        }
        for (int n = threadIdx.x; n < N; n += MQ_NT) f(p.net.b_off[l] + n, p.sb_off[l] + n, -1, 0, n);
    }
    if (p.net.logstd_off >= 0)
        for (int a = threadIdx.x; a < p.net.n_out; a += MQ_NT) f(p.net.logstd_off + a, p.sls_off + a, -1, 0, a);
}

// ---- kernels ---------------------------------------------------------------------------
__global__ void __launch_bounds__(MQ_NT)
fused_forward_kernel(FusedPlan p, const float *__restrict__ theta, const float *__restrict__ x,
                     const int32_t *__restrict__ idx, int64_t n_rows, float *__restrict__ out,
                     const float *__restrict__ sample, float *__restrict__ logp) {
    MQ_SHARED_BYTES(smem);
    float *sW = smem_W(p, smem), *sA = smem_A(p, smem);
    load_weights(p, theta, nullptr, sW, nullptr);
    const int64_t ntiles = (n_rows + p.tb - 1) / p.tb;
    const int N = p.net.n_out, tbp = p.tbp;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * p.tb;
        load_rows(p, x, idx, row0, n_rows, sA + p.act_off[0]);
        __syncthreads();
        forward_tile(p, sW, sA);
        const float *sO = sA + p.act_off[p.net.n_layers];
        for (int r = threadIdx.x; r < p.tb; r += MQ_NT) {
            const int64_t gr = row0 + r;
            if (gr >= n_rows) continue;
            if (out) for (int n = 0; n < N; ++n) out[gr * N + n] = sO[n * tbp + r];
            if (logp) {
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
    float *sW = smem_W(p, smem), *sA = smem_A(p, smem), *sWt = smem_Wt(p, smem);
    float *sG = smem_G(p, smem), *sS = smem_S(p, smem);
    load_weights(p, theta, nullptr, sW, sWt);
    for (int e = threadIdx.x; e < p.p_pad; e += MQ_NT) sG[e] = 0.0f;
    const bool gauss = head.kind == MQ_HEAD_GAUSS_LOGP || head.kind == MQ_HEAD_GAUSS_PPO;
    const int64_t ntiles = (n_rows + p.tb - 1) / p.tb;
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * p.tb;
        load_rows(p, x, idx, row0, n_rows, sA + p.act_off[0]);
        __syncthreads();
        forward_tile(p, sW, sA);
        head_tile(p, head, sW, sA + p.act_off[p.net.n_layers], sA + p.head_off, idx, row0, n_rows, out, logp);
        __syncthreads();
        backward_tile(p, sWt, sA, sG, sS, gauss);
        __syncthreads();
    }
    float *dst = direct ? grad_or_partial : grad_or_partial + (int64_t)blockIdx.x * p.net.n_params;
    const float sc = direct ? scale : 1.0f;
    for_each_param(p, [&](int flat, int pad, int, int, int) { dst[flat] = sG[pad] * sc; });
}

__global__ void __launch_bounds__(MQ_NT)
fused_reduce_kernel(const float *__restrict__ partial, int nblocks, int n_params, float scale,
                    float *__restrict__ grad) {
    for (int j = blockIdx.x * MQ_NT + threadIdx.x; j < n_params; j += gridDim.x * MQ_NT) {
        float s = 0.0f;
        for (int b = 0; b < nblocks; ++b) s += partial[(int64_t)b * n_params + j];
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

// One CTA runs n_steps dependent optimiser steps: weights, their transposes and the
// gradient accumulators stay in shared memory for the whole launch.
template <typename ST>
__global__ void __launch_bounds__(MQ_NT)
fused_train_kernel(FusedPlan p, float *__restrict__ theta, ST *__restrict__ value, ST *__restrict__ am,
                   ST *__restrict__ av, const float *__restrict__ x, mq_head_t head,
                   const int32_t *__restrict__ idx_table, int n_steps, int mb, double beta1,
                   double beta2, double tw1, double tw2, double lr, double eps,
                   double *__restrict__ tw_out) {
    MQ_SHARED_BYTES(smem);
    float *sW = smem_W(p, smem), *sA = smem_A(p, smem), *sWt = smem_Wt(p, smem);
    float *sG = smem_G(p, smem), *sS = smem_S(p, smem);
    load_weights(p, theta, nullptr, sW, sWt);
    for (int e = threadIdx.x; e < p.p_pad; e += MQ_NT) sG[e] = 0.0f;
    const bool gauss = head.kind == MQ_HEAD_GAUSS_LOGP || head.kind == MQ_HEAD_GAUSS_PPO;
    const int ntiles = (mb + p.tb - 1) / p.tb;
    const float inv_mb = 1.0f / (float)mb;
    __syncthreads();
    for (int step = 0; step < n_steps; ++step) {
        const int32_t *idx = idx_table + (int64_t)step * mb;
        for (int tile = 0; tile < ntiles; ++tile) {
            const int64_t row0 = (int64_t)tile * p.tb;
            load_rows(p, x, idx, row0, mb, sA + p.act_off[0]);
            __syncthreads();
            forward_tile(p, sW, sA);
            head_tile(p, head, sW, sA + p.act_off[p.net.n_layers], sA + p.head_off, idx, row0, mb,
                      nullptr, nullptr);
            __syncthreads();
            backward_tile(p, sWt, sA, sG, sS, gauss);
            __syncthreads();
        }
        tw1 = tw1 * beta1 + 1.0;                       // _normalize.py:18-23 with weight 1
        tw2 = tw2 * beta2 + 1.0;
        const ST c1 = (ST)(1.0 / tw1), c2 = (ST)(1.0 / tw2);
        for_each_param(p, [&](int flat, int pad, int l, int k, int n) {
            const float g = sG[pad] * inv_mb;          // batch mean (basicnet.py:242-249)
            sG[pad] = 0.0f;
            ST val = value[flat], m = am[flat], v = av[flat];
            AdamMath<ST>::step(val, m, v, (ST)g, c1, c2, (ST)lr, (ST)eps);
            value[flat] = val; am[flat] = m; av[flat] = v;
            const float w = (float)val;                // load_params: fp64 -> fp32 (basicnet.py:56)
            theta[flat] = w;
            sW[pad] = w;
            if (l >= 1) sWt[p.swt_off[l] + n * p.kpad[l] + k] = w;
        });
        __syncthreads();
    }
    if (threadIdx.x == 0 && tw_out) { tw_out[0] = tw1; tw_out[1] = tw2; }
}

// One CTA per population member: theta + noise[p] is built in shared memory, the shared
// observation batch streams through it, the return is reduced on chip.
__global__ void __launch_bounds__(MQ_NT)
es_eval_kernel(FusedPlan p, const float *__restrict__ theta, const float *__restrict__ noise,
               int64_t n_members, const float *__restrict__ obs, int64_t n_obs,
               const float *__restrict__ ret_coef, float *__restrict__ returns) {
    MQ_SHARED_BYTES(smem);
    __shared__ float red[MQ_NT / 32];
    float *sW = smem_W(p, smem), *sA = smem_A(p, smem);
    const int N = p.net.n_out, tbp = p.tbp;
    const int64_t ntiles = (n_obs + p.tb - 1) / p.tb;
    for (int64_t member = blockIdx.x; member < n_members; member += gridDim.x) {
        load_weights(p, theta, noise + member * p.net.n_params, sW, nullptr);
        float acc = 0.0f;
        for (int64_t tile = 0; tile < ntiles; ++tile) {
            const int64_t row0 = tile * p.tb;
            load_rows(p, obs, nullptr, row0, n_obs, sA + p.act_off[0]);
            __syncthreads();
            forward_tile(p, sW, sA);
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

static int64_t fused_grid(int64_t ntiles) {
    const int64_t sms = mq_sm_count();
    return ntiles < sms ? (ntiles < 1 ? 1 : ntiles) : sms;
}

extern "C" int64_t mq_fused_ws_bytes(const mq_net_t *net, int64_t batch) {
    (void)batch;
    return (int64_t)mq_sm_count() * net->n_params * (int64_t)sizeof(float) + 256;
}

extern "C" int mq_fused_forward(const mq_net_t *net, const float *theta, const float *x,
                                const int32_t *row_idx, int64_t batch, float *out, const float *sample,
                                float *logp, void *stream) {
    int rc = fused_net_check(net, "mq_fused_forward");
    if (rc != MQ_OK) return rc;
    MQ_REQUIRE(batch >= 0, MQ_ERR_SHAPE, "mq_fused_forward: negative batch");
    if (batch == 0) return MQ_OK;
    MQ_REQUIRE(theta && x, MQ_ERR_INVALID, "mq_fused_forward: null pointer");
    MQ_REQUIRE(!logp || (sample && net->logstd_off >= 0), MQ_ERR_INVALID,
               "mq_fused_forward: logp needs sample and logstd parameters");
    FusedPlan p;
    MQ_REQUIRE(plan_choose(net, batch, 0, &p) > 0, MQ_ERR_UNSUPPORTED,
               "mq_fused_forward: net does not fit in shared memory");
    rc = set_smem(fused_forward_kernel, p.smem_bytes, "mq_fused_forward");
    if (rc != MQ_OK) return rc;
    const int64_t grid = fused_grid(mq_ceil_div(batch, p.tb));
    MQ_LAUNCH(fused_forward_kernel, (unsigned)grid, MQ_NT, p.smem_bytes, stream, p, theta, x, row_idx,
              batch, out, sample, logp);
    MQ_CHECK_LAUNCH("mq_fused_forward");
    return MQ_OK;
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
    FusedPlan p;
    MQ_REQUIRE(plan_choose(net, batch, 1, &p) > 0, MQ_ERR_UNSUPPORTED,
               "mq_fused_grad: net does not fit in shared memory");
    rc = set_smem(fused_grad_kernel, p.smem_bytes, "mq_fused_grad");
    if (rc != MQ_OK) return rc;
    const int64_t grid = fused_grid(mq_ceil_div(batch, p.tb));
    if (grid == 1) {
        MQ_LAUNCH(fused_grad_kernel, 1, MQ_NT, p.smem_bytes, stream, p, theta, x, row_idx, batch, *head,
                  scale, grad, 1, out, logp);
    } else {
        MQ_REQUIRE(ws && ws_bytes >= grid * net->n_params * (int64_t)sizeof(float), MQ_ERR_SHAPE,
                   "mq_fused_grad: workspace too small");
        MQ_LAUNCH(fused_grad_kernel, (unsigned)grid, MQ_NT, p.smem_bytes, stream, p, theta, x, row_idx,
                  batch, *head, scale, (float *)ws, 0, out, logp);
        MQ_LAUNCH(fused_reduce_kernel, (unsigned)mq_ceil_div(net->n_params, MQ_NT), MQ_NT, 0, stream,
                  (const float *)ws, (int)grid, net->n_params, scale, grad);
    }
    MQ_CHECK_LAUNCH("mq_fused_grad");
    return MQ_OK;
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
    MQ_REQUIRE(plan_choose(net, mb, 1, &p) > 0, MQ_ERR_UNSUPPORTED,
               "mq_fused_train: net does not fit in shared memory");
    if (state_dtype == MQ_F64) {
        auto kfn = fused_train_kernel<double>;
        rc = set_smem(kfn, p.smem_bytes, "mq_fused_train");
        if (rc != MQ_OK) return rc;
        MQ_LAUNCH(kfn, 1, MQ_NT, p.smem_bytes, stream, p, theta, (double *)value, (double *)m,
                  (double *)v, x, *head, idx_table, (int)n_steps, (int)mb, beta1, beta2, tw_host[0],
                  tw_host[1], lr, eps, (double *)nullptr);
    } else if (state_dtype == MQ_F32) {
        auto kfn = fused_train_kernel<float>;
        rc = set_smem(kfn, p.smem_bytes, "mq_fused_train");
        if (rc != MQ_OK) return rc;
        MQ_LAUNCH(kfn, 1, MQ_NT, p.smem_bytes, stream, p, theta, (float *)value, (float *)m, (float *)v,
                  x, *head, idx_table, (int)n_steps, (int)mb, beta1, beta2, tw_host[0], tw_host[1], lr,
                  eps, (double *)nullptr);
    } else {
        mq_set_error("mq_fused_train: unknown state dtype %d", state_dtype);
        return MQ_ERR_INVALID;
    }
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
