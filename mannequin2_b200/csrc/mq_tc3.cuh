// Operand image of the tensor-core gradient kernel (mq_tc3.cu): the bf16 planes of the weights in the kernel's
// shared-memory layout, the biases and the Gaussian head parameters, as ONE contiguous byte image in global memory.
// The gradient kernel fetches it with a single bulk copy (cp.async.bulk, TMA engine) instead of converting theta in
// every CTA of every launch; mq_tc_image_build writes it from theta, and the optimiser step kernel (mq_step.cu)
// refreshes the entries of the parameters it updates, so in the PPO loop no launch converts weights at all.
//
// Operand formats (T3Img::F, selected per call by MQ_TC_FP16): F = 0, bf16 planes x = x0 + .. + x_{P-1}; F = 1 (two planes
// only), fp16 planes with the LOW plane scaled: x = x0 + 2^-11 x1, x0 = fp16(x), x1 = fp16(2^11 (x - x0)).  An fp16 keeps
// 11 significant bits, the residual x - x0 is exact in fp32 and its scaled copy keeps 11 more bits and a sign, so the
// two planes carry x to 2^-24 relative (fp32 itself) for |x| < 65504 (above: +-inf, the gradient turns non-finite), and
// a product needs three plane pairs where three bf16 planes need six.  The scaled pairs accumulate in their own TMEM
// columns and are folded in with one FMA by the epilogues.
//
// Layout for P planes, layer l with padded sizes kp (inputs) x np (outputs):
//   forward image   [k chunk of 8][plane][n][8 k]   K-major B operand of Z = A W, planes side by side along N
//   backward image  [plane][k chunk of 8][n][8 k]   MN-major B operand of dA = dZ W^T, planes side by side along N'
//                   (layers l >= 1 only)
//   bias            float[3][64]
//   misc            float[32]: logstd[16], exp(-logstd)[16]
#pragma once
#include "mq_tc.cuh"
#ifndef MQ_EMU
#include <cuda_fp16.h>
#endif

#define T3_MAXL 3
#define T3_HID 64
#define T3_NO 16

#define T3_LOW_SCALE 2048.0f               // 2^11: scale of the low fp16 plane
#define T3_LOW_UNSCALE 4.8828125e-4f       // 2^-11

struct T3Img {
    int P, F, L, k0;
    int kp[T3_MAXL], np[T3_MAXL];
    int in[T3_MAXL], out[T3_MAXL], w_off[T3_MAXL], b_off[T3_MAXL];
    int logstd_off, n_out, n_params;
    int off_wf[T3_MAXL], off_wb[T3_MAXL], wb_plane[T3_MAXL];
    int off_bias, off_misc, bytes;
};

// 0 when the net does not fit the kernel
static inline int t3_img_build_layout(const mq_net_t *net, int planes, int fmt, T3Img *g) {
    memset(g, 0, sizeof(*g));
    const int L = net->n_layers;
    if (L < 2 || L > T3_MAXL) return 0;
    if (fmt != 0 && (fmt != 1 || planes != 2)) return 0;
    if (net->n_in > 32 || net->n_out > T3_NO) return 0;
    for (int l = 0; l + 1 < L; ++l) {
        if (net->out[l] > T3_HID) return 0;
        if (net->act[l] != net->act[0] || (net->act[l] != MQ_ACT_TANH && net->act[l] != MQ_ACT_LRELU)) return 0;
    }
    if (net->logstd_off >= 0 && net->n_out > 8) return 0;
    g->P = planes; g->F = fmt; g->L = L; g->k0 = net->n_in <= 16 ? 16 : 32;
    g->logstd_off = net->logstd_off; g->n_out = net->n_out; g->n_params = net->n_params;
    int off = 0;
    for (int l = 0; l < L; ++l) {
        g->kp[l] = l == 0 ? g->k0 : T3_HID;
        g->np[l] = l == L - 1 ? T3_NO : T3_HID;
        g->in[l] = net->in[l]; g->out[l] = net->out[l]; g->w_off[l] = net->w_off[l]; g->b_off[l] = net->b_off[l];
    }
    for (int l = 0; l < L; ++l) { g->off_wf[l] = off; off += planes * g->np[l] * g->kp[l] * 2; }
    for (int l = 1; l < L; ++l) {
        g->wb_plane[l] = g->np[l] * g->kp[l] * 2;
        g->off_wb[l] = off; off += planes * g->wb_plane[l];
    }
    g->off_bias = off; off += T3_MAXL * 64 * 4;
    g->off_misc = off; off += 32 * 4;
    g->bytes = (off + 15) / 16 * 16;
    return 1;
}

#ifndef MQ_EMU
// planes of one value: bf16 w[0] + w[1] + .. (round to nearest, residuals exact in fp32), or fp16 w[0] + 2^-11 w[1]
template <int P, int F>
__device__ __forceinline__ void t3_split1(float v, uint16_t *w) {
    if (F == 1) {
        const __half h0 = __float2half_rn(v);
        const __half h1 = __float2half_rn((v - __half2float(h0)) * T3_LOW_SCALE);
        w[0] = *reinterpret_cast<const uint16_t *>(&h0);
        w[1] = *reinterpret_cast<const uint16_t *>(&h1);
        return;
    }
#pragma unroll
    for (int pl = 0; pl < P; ++pl) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        w[pl] = *reinterpret_cast<const uint16_t *>(&h);
        v -= __bfloat162float(h);
    }
}

// entries of parameter j (value val) in the image
template <int P, int F>
__device__ __forceinline__ void t3_img_store_param(const T3Img &g, unsigned char *img, int j, float val) {
    if (g.logstd_off >= 0 && j >= g.logstd_off) {
        const int a = j - g.logstd_off;
        float *misc = reinterpret_cast<float *>(img + g.off_misc);
        misc[a] = val;
        misc[16 + a] = expf(-val);
        return;
    }
#pragma unroll
    for (int l = 0; l < T3_MAXL; ++l) {
        if (l >= g.L) break;
        const int wj = j - g.w_off[l];
        if (wj >= 0 && wj < g.in[l] * g.out[l]) {
            const int k = wj / g.out[l], n = wj - k * g.out[l];
            uint16_t w[P];
            t3_split1<P, F>(val, w);
            const int np = g.np[l];
            unsigned char *df = img + g.off_wf[l] + (k >> 3) * (P * np * 16) + n * 16 + (k & 7) * 2;
#pragma unroll
            for (int pl = 0; pl < P; ++pl) *reinterpret_cast<uint16_t *>(df + pl * np * 16) = w[pl];
            if (l > 0) {
                unsigned char *db = img + g.off_wb[l] + (k >> 3) * (np * 16) + n * 16 + (k & 7) * 2;
#pragma unroll
                for (int pl = 0; pl < P; ++pl) *reinterpret_cast<uint16_t *>(db + pl * g.wb_plane[l]) = w[pl];
            }
            return;
        }
        const int bj = j - g.b_off[l];
        if (bj >= 0 && bj < g.out[l]) {
            reinterpret_cast<float *>(img + g.off_bias)[l * 64 + bj] = val;
            return;
        }
    }
}

__device__ __forceinline__ void t3_img_store_param_rt(const T3Img &g, unsigned char *img, int j, float val) {
    if (g.P == 3) t3_img_store_param<3, 0>(g, img, j, val);
    else if (g.P == 2 && g.F == 1) t3_img_store_param<2, 1>(g, img, j, val);
    else if (g.P == 2) t3_img_store_param<2, 0>(g, img, j, val);
    else t3_img_store_param<1, 0>(g, img, j, val);
}
#endif
