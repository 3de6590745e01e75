// Generic-stride FP32 GEMM with fused epilogues, and the layered MLP path on top.
//
// This is the path for Affine layers too large for the shared-memory-resident
// fused kernels (mq_fused.cu), e.g. the 784x128 first layer of mnist/mlp.py,
// and for `matmul` nodes of the general graph (mannequin/backprop.py:45-51).
// FFMA register tiles (4x4 or 2x2 per thread), operand tiles staged in shared
// memory, optional deterministic split-K, optional row gather on the batch axis
// (mannequin/_trajectory.py:79-86 fused into the operand load).
//
// Epilogues (all mannequin/backprop.py):
//   EPI_SCALE    C = alpha*acc                         (dW = x^T g / B, lines 48-51)
//   EPI_BIAS_ACT C = act(acc + bias[n])                (add 11-19, relu 53-56, tanh 58-60)
//   EPI_ACT_GRAD C = act'(Y[m,n]) * acc                (g W^T through the previous activation)
#include "mq_gemm.cuh"

template <int BM, int BN, bool SPLIT>
__global__ void __launch_bounds__(MQ_NT)
gemm_kernel(GemmArgs g) {
    constexpr int TM = BM / 16, TN = BN / 16;
    __shared__ __align__(16) float As[GEMM_BK][BM + 4];
    __shared__ __align__(16) float Bs[GEMM_BK][BN + 4];
    const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
    const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
    const int64_t k_begin = (int64_t)blockIdx.z * g.k_per_split;
    const int64_t k_end = (k_begin + g.k_per_split < g.k) ? k_begin + g.k_per_split : g.k;
    constexpr int LA = BM * GEMM_BK / MQ_NT, LB = BN * GEMM_BK / MQ_NT;
    float ra[LA], rb[LB];
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    const bool a_kfast = (g.a_cs == 1);
    const bool b_nfast = (g.b_cs == 1);

    auto load_tiles = [&](int64_t kt) {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            const int e = t + i * MQ_NT;
            const int mm = a_kfast ? e / GEMM_BK : e % BM;
            const int kk = a_kfast ? e % GEMM_BK : e / BM;
            int64_t gm = m0 + mm, gk = kt + kk;
            float v = 0.0f;
            if (gm < g.m && gk < k_end) {
                if (g.a_ones_row > 0 && gm >= g.a_ones_row) {
                    v = 1.0f;
                } else {
                    if (g.idx) { if (g.idx_on_k) gk = g.idx[gk]; else gm = g.idx[gm]; }
                    v = g.a[gm * g.a_rs + gk * g.a_cs];
                }
            }
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            const int e = t + i * MQ_NT;
            const int kk = b_nfast ? e / BN : e % GEMM_BK;
            const int nn = b_nfast ? e % BN : e / GEMM_BK;
            const int64_t gk = kt + kk, gn = n0 + nn;
            rb[i] = (gk < k_end && gn < g.n) ? g.b[gk * g.b_rs + gn * g.b_cs] : 0.0f;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < LA; ++i) {
            const int e = t + i * MQ_NT;
            const int mm = a_kfast ? e / GEMM_BK : e % BM;
            const int kk = a_kfast ? e % GEMM_BK : e / BM;
            As[kk][mm] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < LB; ++i) {
            const int e = t + i * MQ_NT;
            const int kk = b_nfast ? e / BN : e % GEMM_BK;
            const int nn = b_nfast ? e % BN : e / GEMM_BK;
            Bs[kk][nn] = rb[i];
        }
    };

    if (k_begin < k_end) {
        load_tiles(k_begin);
        store_tiles();
        __syncthreads();
        for (int64_t kt = k_begin; kt < k_end; kt += GEMM_BK) {
            const bool more = kt + GEMM_BK < k_end;
            if (more) load_tiles(kt + GEMM_BK);
#pragma unroll
            for (int kk = 0; kk < GEMM_BK; ++kk) {
                float av[TM], bv[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) av[i] = As[kk][ty * TM + i];
#pragma unroll
                for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
            if (more) { store_tiles(); __syncthreads(); }
        }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int64_t gm = m0 + ty * TM + i;
        if (gm >= g.m) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int64_t gn = n0 + tx * TN + j;
            if (gn >= g.n) continue;
            if (SPLIT) g.c[((int64_t)blockIdx.z * g.m + gm) * g.n + gn] = acc[i][j];
            else g.c[gm * g.n + gn] = gemm_epilogue(g, acc[i][j], gm, gn);
        }
    }
}

__global__ void __launch_bounds__(MQ_NT)
gemm_splitk_finish_kernel(GemmArgs g, const float *__restrict__ partial, int splits, float *__restrict__ c) {
    const int64_t total = g.m * g.n;
    for (int64_t e = (int64_t)blockIdx.x * MQ_NT + threadIdx.x; e < total; e += (int64_t)gridDim.x * MQ_NT) {
        float s = 0.0f;
        for (int z = 0; z < splits; ++z) s += partial[(int64_t)z * total + e];
        c[e] = gemm_epilogue(g, s, e / g.n, e % g.n);
    }
}

int mq_gemm_tma_launch(const GemmArgs &g, int mode, int64_t splits, void *pack, void *stream);        // mq_gemm_tma.cu (mode = planes | MQ_TC_FP16)
int64_t mq_gemm_tma_pack_bytes(int64_t m, int64_t n, int64_t k, int planes);

// split-K of the tensor-core GEMM: products with few 128 x 128 output tiles and a long contraction (dW = x^T g)
static int64_t gemm_tc_splits(int64_t m, int64_t n, int64_t k, int64_t *k_per_split) {
    const int sms = mq_sm_count();
    const int64_t tiles = mq_ceil_div(m, 128) * mq_ceil_div(n, 128);
    int64_t splits = 1;
    if (2 * tiles <= sms && k >= 1024) {
        splits = sms / tiles;
        if (splits > k / 256) splits = k / 256;
        if (splits < 1) splits = 1;
    }
    *k_per_split = mq_round_up(mq_ceil_div(k > 0 ? k : 1, splits), 64);
    return mq_ceil_div(k > 0 ? k : 1, *k_per_split);
}

static bool gemm_tc_wanted(const GemmArgs &g, int32_t tc) {
    if ((tc & MQ_TC_OFF) || g.m <= 0 || g.n <= 0 || g.k <= 0) return false;
    if (tc & MQ_TC_FORCE) return true;
    return g.m * g.n * g.k >= (1LL << 26) && g.n >= 64 && g.k >= 64 && g.m >= 128;
}

// bytes of workspace with which gemm_run can take the tensor-core path for this shape (packed planes + split-K partials)
static int64_t gemm_tc_ws_bytes(int64_t m, int64_t n, int64_t k, int planes) {
    int64_t kps;
    const int64_t splits = gemm_tc_splits(m, n, k, &kps);
    return mq_gemm_tma_pack_bytes(m, n, k, planes) + (splits > 1 ? mq_round_up(splits * m * n * (int64_t)sizeof(float), 256) : 0);
}

extern "C" int64_t mq_gemm_ws_bytes(int64_t m, int64_t n, int64_t k) {
    if (m <= 0 || n <= 0 || k <= 0) return 256;
    // FFMA split-K partials (at most 64 splits, only for few-tile products) or the tensor-core path, whichever is larger
    const int64_t ffma = (mq_ceil_div(m, 32) * mq_ceil_div(n, 32) < mq_sm_count() && k >= 512) ? 64 * m * n * (int64_t)sizeof(float) : 0;
    const int64_t tc = gemm_tc_ws_bytes(m, n, k, 3);
    return (ffma > tc ? ffma : tc) + 256;
}

// ---- skinny products of a narrow output layer at large batch -------------------------------------------------------
// The 128 -> 10 layer of mnist/mlp.py at batch 65,536 is three products with one dimension <= 16: through the 64 x 64
// tiles of gemm_kernel they did 6x the arithmetic of the layer and ran at a tenth of the HBM rate the operands allow
// (49 + 67 + 46 us against ~6 us of traffic each).  One small kernel per shape, all bound by reading or writing the
// batch x width activations once:
//   SKINNY_N   C[m, n <= 16]  = epi(A[m, k] B[k, n])     warp per row, lanes along k, B in shared memory
//   SKINNY_K   C[m, n]        = epi(A[m, k <= 16] B[k, n])     warp per row, lanes along n, B in shared memory
//   SKINNY_DW  C[m <= 512, n <= 16] = alpha A[m, k] B[k, n] with k = the batch and A stored k-major (dW = x^T g):
//              a block per slab of k, a thread per row m with its n sums in registers, fixed-order finish over the slabs
#define SKINNY_MAX 16
#define SKINNY_SMEM_FLOATS 4096
#define SKINNY_ROWS 128                  // rows of A per block (skinny N): a thread per row, A staged 32 columns at a time

// C[m, n <= 16] = epi(A[m, k] B[k, n]).  A block stages 128 rows x 32 columns of A in shared memory with coalesced loads
// (every thread has 16 independent loads in flight), then thread r walks its row against B (shared memory, broadcast reads).
__global__ void __launch_bounds__(SKINNY_ROWS)
gemm_skinny_n_kernel(GemmArgs g) {
    __shared__ float bs[SKINNY_SMEM_FLOATS];                 // B[k][n]
    __shared__ float as[SKINNY_ROWS][33];
    const int n = (int)g.n, k = (int)g.k;
    for (int e = threadIdx.x; e < k * n; e += SKINNY_ROWS) bs[e] = g.b[(int64_t)(e / n) * g.b_rs + (int64_t)(e % n) * g.b_cs];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    for (int64_t r0 = (int64_t)blockIdx.x * SKINNY_ROWS; r0 < g.m; r0 += (int64_t)gridDim.x * SKINNY_ROWS) {
        float acc[SKINNY_MAX];
#pragma unroll
        for (int j = 0; j < SKINNY_MAX; ++j) acc[j] = 0.0f;
        for (int kb = 0; kb < k; kb += 32) {
            __syncthreads();
            // warp w loads rows w, w + 4, ...: 32 consecutive columns of a row per instruction
            float v[SKINNY_ROWS / 4];
#pragma unroll
            for (int i = 0; i < SKINNY_ROWS / 4; ++i) {
                const int64_t r = r0 + wrp + 4 * i;
                v[i] = (r < g.m && kb + lane < k) ? __ldg(g.a + (g.idx ? (int64_t)g.idx[r] : r) * g.a_rs + (int64_t)(kb + lane) * g.a_cs) : 0.0f;
            }
#pragma unroll
            for (int i = 0; i < SKINNY_ROWS / 4; ++i) as[wrp + 4 * i][lane] = v[i];
            __syncthreads();
            const int kn = k - kb < 32 ? k - kb : 32;
            for (int kk = 0; kk < kn; ++kk) {
                const float a = as[threadIdx.x][kk];
                const float *brow = bs + (kb + kk) * n;
#pragma unroll
                for (int j = 0; j < SKINNY_MAX; ++j) if (j < n) acc[j] = fmaf(a, brow[j], acc[j]);
            }
        }
        const int64_t r = r0 + threadIdx.x;
        if (r < g.m) {
#pragma unroll
            for (int j = 0; j < SKINNY_MAX; ++j) if (j < n) g.c[r * g.n + j] = gemm_epilogue(g, acc[j], r, j);
        }
    }
}

// C[m, n] = epi(A[m, k <= 16] B[k, n]): elementwise over C, four consecutive columns per thread (n % 4 == 0): every load is
// independent of every other, the k values of a row are shared by the threads of the row through L1.
__global__ void __launch_bounds__(MQ_NT)
gemm_skinny_k_kernel(GemmArgs g) {
    __shared__ __align__(16) float bs[SKINNY_SMEM_FLOATS];   // B[k][n]
    const int n = (int)g.n, k = (int)g.k, nq = n >> 2;
    for (int e = threadIdx.x; e < k * n; e += MQ_NT) bs[e] = g.b[(int64_t)(e / n) * g.b_rs + (int64_t)(e % n) * g.b_cs];
    __syncthreads();
    const int64_t total = g.m * nq;
    for (int64_t e = (int64_t)blockIdx.x * MQ_NT + threadIdx.x; e < total; e += (int64_t)gridDim.x * MQ_NT) {
        const int64_t r = e / nq;
        const int c0 = (int)(e - r * nq) * 4;
        const float *arow = g.a + (g.idx ? (int64_t)g.idx[r] : r) * g.a_rs;
        float a[SKINNY_MAX];
#pragma unroll
        for (int j = 0; j < SKINNY_MAX; ++j) a[j] = j < k ? __ldg(arow + (int64_t)j * g.a_cs) : 0.0f;
        float4 y4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (g.epi == EPI_ACT_GRAD) y4 = *reinterpret_cast<const float4 *>(g.y + r * g.n + c0);
        float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int j = 0; j < SKINNY_MAX; ++j) {
            if (j < k) {
                const float4 b4 = *reinterpret_cast<const float4 *>(bs + j * n + c0);
                acc[0] = fmaf(a[j], b4.x, acc[0]); acc[1] = fmaf(a[j], b4.y, acc[1]);
                acc[2] = fmaf(a[j], b4.z, acc[2]); acc[3] = fmaf(a[j], b4.w, acc[3]);
            }
        }
        float4 o;
        if (g.epi == EPI_ACT_GRAD) {
            o.x = mq_act_grad(y4.x, acc[0], g.act, g.leak); o.y = mq_act_grad(y4.y, acc[1], g.act, g.leak);
            o.z = mq_act_grad(y4.z, acc[2], g.act, g.leak); o.w = mq_act_grad(y4.w, acc[3], g.act, g.leak);
        } else if (g.epi == EPI_BIAS_ACT) {
            o.x = mq_act_apply(acc[0] + g.bias[c0], g.act, g.leak); o.y = mq_act_apply(acc[1] + g.bias[c0 + 1], g.act, g.leak);
            o.z = mq_act_apply(acc[2] + g.bias[c0 + 2], g.act, g.leak); o.w = mq_act_apply(acc[3] + g.bias[c0 + 3], g.act, g.leak);
        } else {
            o = make_float4(g.alpha * acc[0], g.alpha * acc[1], g.alpha * acc[2], g.alpha * acc[3]);
        }
        *reinterpret_cast<float4 *>(g.c + r * g.n + c0) = o;
    }
}

// partial[slab][m][n]; A(m, kk) = a[m + col(kk) * a_cs] (rows of A are the unit-stride axis), B(kk, j) = b[kk * b_rs + j].
// Thread = row m; the loop over the slab is unrolled by 8 with the loads of a step issued before its arithmetic.
__global__ void __launch_bounds__(MQ_NT)
gemm_skinny_dw_kernel(GemmArgs g, int64_t k_per_slab, float *__restrict__ partial) {
    __shared__ float bs[8 * SKINNY_MAX];
    const int n = (int)g.n;
    const int64_t k0 = (int64_t)blockIdx.x * k_per_slab;
    const int64_t k1 = k0 + k_per_slab < g.k ? k0 + k_per_slab : g.k;
    float acc[2][SKINNY_MAX];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int j = 0; j < SKINNY_MAX; ++j) acc[p][j] = 0.0f;
    for (int64_t kb = k0; kb < k1; kb += 8) {
        __syncthreads();
        if (threadIdx.x < 8 * n) {
            const int kk = threadIdx.x / n, j = threadIdx.x - kk * n;
            bs[kk * SKINNY_MAX + j] = kb + kk < k1 ? __ldg(g.b + (kb + kk) * g.b_rs + j) : 0.0f;
        }
        float a[2][8];
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const int64_t m = threadIdx.x + p * MQ_NT;
            const bool ones = g.a_ones_row > 0 && m >= g.a_ones_row;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = 0.0f;
                if (m < g.m && kb + i < k1) {
                    const int64_t sk = (g.idx && g.idx_on_k) ? (int64_t)g.idx[kb + i] : kb + i;
                    v = ones ? 1.0f : __ldg(g.a + m + sk * g.a_cs);
                }
                a[p][i] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < SKINNY_MAX; ++j)
                if (j < n) {
                    const float bv = bs[i * SKINNY_MAX + j];
                    acc[0][j] = fmaf(a[0][i], bv, acc[0][j]);
                    acc[1][j] = fmaf(a[1][i], bv, acc[1][j]);
                }
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int64_t m = threadIdx.x + p * MQ_NT;
        if (m < g.m) {
            float *dst = partial + ((int64_t)blockIdx.x * g.m + m) * n;
#pragma unroll
            for (int j = 0; j < SKINNY_MAX; ++j) if (j < n) dst[j] = acc[p][j];
        }
    }
}

// c[e] = alpha * sum over the slabs of partial[slab][e], a warp per element: lanes take slabs lane, lane + 32, ... (all loads
// independent), then a fixed butterfly -- bit-reproducible
__global__ void __launch_bounds__(MQ_NT)
gemm_skinny_finish_kernel(const float *__restrict__ partial, int slabs, int64_t total, float alpha, float *__restrict__ c) {
    const int lane = threadIdx.x & 31;
    const int64_t e = ((int64_t)blockIdx.x * MQ_NT + threadIdx.x) >> 5;
    if (e >= total) return;
    float s = 0.0f;
    for (int z = lane; z < slabs; z += 32) s += __ldg(partial + (int64_t)z * total + e);
    s = mq_warp_sum(s);
    if (lane == 0) c[e] = alpha * s;
}

#ifndef MQ_EMU
// ---- streaming forms of the three skinny products (dense fp32 operands, the shapes of a narrow output layer) -------------
// Each reads / writes the batch x width activations ONCE at HBM rate; the general kernels above remain for strided or
// gathered operands and other widths.

// SKINNY_N, A rows dense (a_cs = 1, 16-byte aligned rows), k <= 128, k % 4 == 0: a WARP per row, lane l holds columns
// 4l .. 4l+3 of the row (one 128-bit load: the row is one coalesced 512-byte read) and rows 4l .. 4l+3 of B in registers;
// the 32 partial sums of each of the (<= 16) outputs are added by a transposing butterfly (halving the live outputs at
// every level: 16 shuffles per row instead of 5 per output).  Four rows per warp in flight.  Fixed order: bit-reproducible.
__global__ void __launch_bounds__(128, 4)
gemm_skinny_n_warp_kernel(GemmArgs g) {
    const int n = (int)g.n, k = (int)g.k;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * 128 + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * 128) >> 5;
    float b[4][SKINNY_MAX];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < SKINNY_MAX; ++j)
            b[i][j] = (4 * lane + i < k && j < n) ? __ldg(g.b + (int64_t)(4 * lane + i) * g.b_rs + (int64_t)j * g.b_cs) : 0.0f;
    const bool has = 4 * lane < k;
    const int slot = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    for (int64_t r0 = warp * 4; r0 < g.m; r0 += nwarps * 4) {
        float4 a[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t r = r0 + q;
            a[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (has && r < g.m) a[q] = __ldg(reinterpret_cast<const float4 *>(g.a + (g.idx ? (int64_t)g.idx[r] : r) * g.a_rs) + lane);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float p[SKINNY_MAX];
#pragma unroll
            for (int j = 0; j < SKINNY_MAX; ++j)
                p[j] = fmaf(a[q].w, b[3][j], fmaf(a[q].z, b[2][j], fmaf(a[q].y, b[1][j], a[q].x * b[0][j])));
            float p8[8], p4[4], p2[2];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const bool up = (lane & 16) != 0;
                const float got = __shfl_xor_sync(0xffffffffu, up ? p[i] : p[i + 8], 16);
                p8[i] = (up ? p[i + 8] : p[i]) + got;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool up = (lane & 8) != 0;
                const float got = __shfl_xor_sync(0xffffffffu, up ? p8[i] : p8[i + 4], 8);
                p4[i] = (up ? p8[i + 4] : p8[i]) + got;
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const bool up = (lane & 4) != 0;
                const float got = __shfl_xor_sync(0xffffffffu, up ? p4[i] : p4[i + 2], 4);
                p2[i] = (up ? p4[i + 2] : p4[i]) + got;
            }
            const bool up = (lane & 2) != 0;
            float v = (up ? p2[1] : p2[0]) + __shfl_xor_sync(0xffffffffu, up ? p2[0] : p2[1], 2);
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            const int64_t r = r0 + q;
            if (r < g.m && (lane & 1) == 0 && slot < n) g.c[r * g.n + slot] = gemm_epilogue(g, v, r, slot);
        }
    }
}

// SKINNY_N, thread-per-row form (the default; the warp-per-row kernel above measured 293 warp instructions per row -- all 16
// output slots computed and reduced, 128 registers -- and stays for n > 12... it is the fallback when shared memory cannot be
// raised): a block of 128 threads owns 128 rows; THREAD r fetches ITS row (k * 4 <= 512 bytes, gathered rows included) with one
// bulk copy (TMA engine) into a padded shared-memory row (pitch 132 floats: 128-bit reads of consecutive rows fall in different
// bank groups), all 128 copies complete on one mbarrier; B sits in shared memory padded to NP columns and is read as
// broadcast 128-bit words.  ~53 warp instructions per row instead of 293; no block barrier in the row loop (a thread re-arms
// only its own row).
#define SKN_PITCH 132
template <int NP>
__global__ void __launch_bounds__(128)
gemm_skinny_n_rows_kernel(GemmArgs g) {
    extern __shared__ __align__(128) float skn_smem[];
    __shared__ __align__(8) uint64_t bar_mem;
    float *as = skn_smem, *bs = skn_smem + 128 * SKN_PITCH;     // A tile [128][132], B [k][NP]
    const int n = (int)g.n, k = (int)g.k;
    for (int e = threadIdx.x; e < k * NP; e += 128) {
        const int kk = e / NP, j = e - kk * NP;
        bs[e] = j < n ? __ldg(g.b + (int64_t)kk * g.b_rs + (int64_t)j * g.b_cs) : 0.0f;
    }
    const unsigned bar = (unsigned)__cvta_generic_to_shared(&bar_mem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(128u));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    float *my = as + threadIdx.x * SKN_PITCH;
    const unsigned my_s = (unsigned)__cvta_generic_to_shared(my);
    unsigned phase = 0;
    for (int64_t r0 = (int64_t)blockIdx.x * 128; r0 < g.m; r0 += (int64_t)gridDim.x * 128) {
        const int64_t r = r0 + threadIdx.x;
        if (r < g.m) {
            const float *src = g.a + (g.idx ? (int64_t)g.idx[r] : r) * g.a_rs;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // my reads of the row's previous contents are done
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)k * 4u) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(my_s), "l"(src), "r"((unsigned)k * 4u), "r"(bar) : "memory");
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
        }
        unsigned ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
        } while (!ok);
        phase ^= 1u;
        if (r >= g.m) continue;
        float acc[NP];
#pragma unroll
        for (int j = 0; j < NP; ++j) acc[j] = 0.0f;
        for (int q = 0; q < k; q += 4) {
            const float4 a4 = *reinterpret_cast<const float4 *>(my + q);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j4 = 0; j4 < NP; j4 += 4) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(bs + (q + i) * NP + j4);
                    acc[j4] = fmaf(av[i], b4.x, acc[j4]); acc[j4 + 1] = fmaf(av[i], b4.y, acc[j4 + 1]);
                    acc[j4 + 2] = fmaf(av[i], b4.z, acc[j4 + 2]); acc[j4 + 3] = fmaf(av[i], b4.w, acc[j4 + 3]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NP; ++j) if (j < n) g.c[r * g.n + j] = gemm_epilogue(g, acc[j], r, j);
    }
}

// SKINNY_K, n <= 128 (one column quad per lane): a WARP per row, lane l owns columns 4l .. 4l+3 with ITS quad of every row of B
// in registers (no shared memory, no staging barrier); the row's k values are one coalesced load (lane j loads A[r][j]) handed
// round by shuffles; four rows per warp in flight.  The item-per-thread kernel below measured 490 warp instructions per row
// (a 64-bit division per item, 16 predicated k steps with their shared-memory loads): instruction-issue bound at 44 us for
// 70 MB of traffic.
template <int KP, int U>                                      // k <= KP rows of B in registers, U rows of A per warp in flight
__global__ void __launch_bounds__(MQ_NT, 2)
gemm_skinny_k_warp_kernel(GemmArgs g) {
    const int n = (int)g.n, k = (int)g.k, nq = n >> 2;
    const int lane = threadIdx.x & 31;
    const bool col = lane < nq;
    float4 b[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) {
        b[j] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (col && j < k) {
            const float *bp = g.b + (int64_t)j * g.b_rs + (int64_t)(4 * lane) * g.b_cs;
            b[j] = make_float4(__ldg(bp), __ldg(bp + g.b_cs), __ldg(bp + 2 * g.b_cs), __ldg(bp + 3 * g.b_cs));
        }
    }
    float4 bias4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (col && g.epi == EPI_BIAS_ACT) bias4 = make_float4(g.bias[4 * lane], g.bias[4 * lane + 1], g.bias[4 * lane + 2], g.bias[4 * lane + 3]);
    const int64_t warp = ((int64_t)blockIdx.x * MQ_NT + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * MQ_NT) >> 5;
    for (int64_t r0 = warp * U; r0 < g.m; r0 += nwarps * U) {
        float av[U];
        float4 y4[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t r = r0 + u;
            av[u] = 0.0f;
            y4[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (r < g.m) {
                if (lane < k) av[u] = __ldg(g.a + (g.idx ? (int64_t)g.idx[r] : r) * g.a_rs + (int64_t)lane * g.a_cs);
                if (col && g.epi == EPI_ACT_GRAD) y4[u] = __ldg(reinterpret_cast<const float4 *>(g.y + r * g.n) + lane);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            // no branch on the row or on k in here: rows past the end compute on zeros and skip the store, rows j >= k of B are
            // zero registers -- the U rows' FMA chains are then free to interleave (with the branches the kernel issued 207
            // warp instructions per row at 43 % issue activity: four dependent 10-step chains per row, one row at a time)
            const int64_t r = r0 + u;
            float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
            for (int j = 0; j < KP; ++j) {
                const float a = __shfl_sync(0xffffffffu, av[u], j);
                acc.x = fmaf(a, b[j].x, acc.x); acc.y = fmaf(a, b[j].y, acc.y);
                acc.z = fmaf(a, b[j].z, acc.z); acc.w = fmaf(a, b[j].w, acc.w);
            }
            float4 o;
            if (g.epi == EPI_ACT_GRAD) {
                o.x = mq_act_grad(y4[u].x, acc.x, g.act, g.leak); o.y = mq_act_grad(y4[u].y, acc.y, g.act, g.leak);
                o.z = mq_act_grad(y4[u].z, acc.z, g.act, g.leak); o.w = mq_act_grad(y4[u].w, acc.w, g.act, g.leak);
            } else if (g.epi == EPI_BIAS_ACT) {
                o.x = mq_act_apply(acc.x + bias4.x, g.act, g.leak); o.y = mq_act_apply(acc.y + bias4.y, g.act, g.leak);
                o.z = mq_act_apply(acc.z + bias4.z, g.act, g.leak); o.w = mq_act_apply(acc.w + bias4.w, g.act, g.leak);
            } else {
                o = make_float4(g.alpha * acc.x, g.alpha * acc.y, g.alpha * acc.z, g.alpha * acc.w);
            }
            if (col && r < g.m) *(reinterpret_cast<float4 *>(g.c + r * g.n) + lane) = o;
        }
    }
}

// SKINNY_DW, both operands dense (A = x^T with x [k][md] contiguous, B [k][n] contiguous; md <= 128, md % 4 == 0, k % 4 == 0):
// a block per slab of k; tiles of 64 batch rows of x and of B arrive in shared memory by bulk copies (TMA engine, two stages,
// one issuing thread, completion counted on an mbarrier), thread (m, half) adds the rows of its parity into n sums in
// registers (x conflict-free along m, B broadcast), halves combined in a fixed order; the bias-gradient row (a_ones_row)
// is the column sum of B, kept by the m = 0 threads.
#define SKDW_TILE 64
#define SKDW_STAGE (SKDW_TILE * 128 * 4 + SKDW_TILE * SKINNY_MAX * 4)
template <int NP>                                             // n <= NP sums per thread (columns n .. NP-1 read neighbouring words of
__global__ void __launch_bounds__(MQ_NT)                      // the stage and are never stored: no predicate in the inner loop)
gemm_skinny_dw_tma_kernel(GemmArgs g, int64_t k_per_slab, float *__restrict__ partial) {
    extern __shared__ __align__(128) unsigned char sk_smem[];
    __shared__ __align__(8) uint64_t full[2];
    const int n = (int)g.n, md = (int)(g.a_ones_row > 0 ? g.a_ones_row : g.m);
    const int64_t k0 = (int64_t)blockIdx.x * k_per_slab;
    const int64_t k1 = k0 + k_per_slab < g.k ? k0 + k_per_slab : g.k;
    const int ntiles = (int)((k1 - k0 + SKDW_TILE - 1) / SKDW_TILE);
    const int m = threadIdx.x & 127, half = threadIdx.x >> 7;
    auto bar = [&](int s) { return (unsigned)__cvta_generic_to_shared(&full[s]); };
    auto issue = [&](int t) {
        const int s = t & 1;
        const int64_t kb = k0 + (int64_t)t * SKDW_TILE;
        const unsigned rows = (unsigned)(k1 - kb < SKDW_TILE ? k1 - kb : SKDW_TILE);
        const unsigned ba = rows * (unsigned)md * 4u, bb = rows * (unsigned)n * 4u;
        const unsigned dst = (unsigned)__cvta_generic_to_shared(sk_smem + (size_t)s * SKDW_STAGE);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar(s)), "r"(ba + bb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(g.a + kb * md), "r"(ba), "r"(bar(s)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst + SKDW_TILE * 128 * 4), "l"(g.b + kb * n), "r"(bb), "r"(bar(s)) : "memory");
    };
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar(0)), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar(1)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;");
        if (ntiles > 0) issue(0);
        if (ntiles > 1) issue(1);
    }
    __syncthreads();
    float acc[NP], ones = 0.0f;                              // ones: column threadIdx.x of B summed over the slab (threads < n)
#pragma unroll
    for (int j = 0; j < NP; ++j) acc[j] = 0.0f;
    for (int t = 0; t < ntiles; ++t) {
        const int s = t & 1;
        unsigned ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar(s)), "r"((unsigned)((t >> 1) & 1)) : "memory");
        } while (!ok);
        const int64_t kb = k0 + (int64_t)t * SKDW_TILE;
        const int rows = (int)(k1 - kb < SKDW_TILE ? k1 - kb : SKDW_TILE);
        const float *xa = reinterpret_cast<const float *>(sk_smem + (size_t)s * SKDW_STAGE);
        const float *gb = xa + SKDW_TILE * 128;
        if (m < md) {
#pragma unroll 4
            for (int r = half; r < rows; r += 2) {
                const float a = xa[r * md + m];
                const float *gr = gb + r * n;
#pragma unroll
                for (int j = 0; j < NP; ++j) acc[j] = fmaf(a, gr[j], acc[j]);
            }
        }
        if ((int)threadIdx.x < n)
            for (int r = 0; r < rows; ++r) ones += gb[r * n + threadIdx.x];
        __syncthreads();                                   // every read of stage s is done
        if (threadIdx.x == 0 && t + 2 < ntiles) issue(t + 2);
    }
    // halves in a fixed order through the (now idle) first stage
    float *red = reinterpret_cast<float *>(sk_smem);
    if (half == 1) {
#pragma unroll
        for (int j = 0; j < NP; ++j) red[m * SKINNY_MAX + j] = acc[j];
    }
    __syncthreads();
    if (half == 0) {
        float *dst = partial + (int64_t)blockIdx.x * g.m * n;
#pragma unroll
        for (int j = 0; j < NP; ++j)
            if (j < n && m < md) dst[(int64_t)m * n + j] = acc[j] + red[m * SKINNY_MAX + j];
        if ((int)threadIdx.x < n && g.a_ones_row > 0) dst[g.a_ones_row * n + threadIdx.x] = ones;
    }
}
#endif

// 1 = handled.  Shapes outside the kernels' limits fall through to the general kernels.
static int gemm_skinny(const GemmArgs &g, void *ws, int64_t ws_bytes, void *stream) {
#ifdef MQ_EMU
    (void)g; (void)ws; (void)ws_bytes; (void)stream;
    return 0;
#else
    const int sms = mq_sm_count();
    if (g.m >= 2048 && g.n <= SKINNY_MAX && g.k >= 32 && g.k * g.n <= SKINNY_SMEM_FLOATS && !(g.idx && g.idx_on_k) && g.a_ones_row == 0) {
        if (g.a_cs == 1 && g.k <= 128 && (g.k & 3) == 0 && (g.a_rs & 3) == 0 && ((uintptr_t)g.a & 15) == 0) {
            const int np = g.n <= 12 ? 12 : 16;
            const int smem = (128 * SKN_PITCH + (int)g.k * np) * (int)sizeof(float);
            auto kr = np == 12 ? gemm_skinny_n_rows_kernel<12> : gemm_skinny_n_rows_kernel<16>;
            if (cudaFuncSetAttribute(kr, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) == cudaSuccess) {
                int64_t nbr = mq_ceil_div(g.m, 128);
                if (nbr > (int64_t)sms * 3) nbr = (int64_t)sms * 3;          // three blocks (3 x 72 KB) per SM
                MQ_LAUNCH(kr, (unsigned)nbr, 128, smem, stream, g);
                return 1;
            }
            cudaGetLastError();
            int64_t nbw = mq_ceil_div(g.m, 16);               // 4 warps x 4 rows per block and pass
            if (nbw > (int64_t)sms * 16) nbw = (int64_t)sms * 16;
            MQ_LAUNCH(gemm_skinny_n_warp_kernel, (unsigned)nbw, 128, 0, stream, g);
            return 1;
        }
        int64_t nb = mq_ceil_div(g.m, SKINNY_ROWS);
        if (nb > (int64_t)sms * 8) nb = (int64_t)sms * 8;
        MQ_LAUNCH(gemm_skinny_n_kernel, (unsigned)nb, SKINNY_ROWS, 0, stream, g);
        return 1;
    }
    if (g.m >= 2048 && g.k <= SKINNY_MAX && g.n >= 32 && (g.n & 3) == 0 && g.k * g.n <= SKINNY_SMEM_FLOATS && !(g.idx && g.idx_on_k) &&
        g.a_ones_row == 0 && (((uintptr_t)g.c | (uintptr_t)g.y) & 15) == 0) {
        if (g.n <= 128) {
            int64_t nbw = mq_ceil_div(g.m, 64);               // 8 warps x 8 (4) rows per block and pass
            if (nbw > (int64_t)sms * 2) nbw = (int64_t)sms * 2;
            if (g.k <= 10) { auto kk = gemm_skinny_k_warp_kernel<10, 8>; MQ_LAUNCH(kk, (unsigned)nbw, MQ_NT, 0, stream, g); }
            else { auto kk = gemm_skinny_k_warp_kernel<SKINNY_MAX, 4>; MQ_LAUNCH(kk, (unsigned)nbw, MQ_NT, 0, stream, g); }
            return 1;
        }
        int64_t nb = mq_ceil_div(g.m * (g.n >> 2), MQ_NT);
        if (nb > (int64_t)sms * 16) nb = (int64_t)sms * 16;
        MQ_LAUNCH(gemm_skinny_k_kernel, (unsigned)nb, MQ_NT, 0, stream, g);
        return 1;
    }
    if (ws && g.k >= 4096 && g.n <= SKINNY_MAX && g.m <= 2 * MQ_NT && g.a_rs == 1 && g.b_cs == 1 && g.epi == EPI_SCALE && !(g.idx && !g.idx_on_k)) {
        const int64_t md = g.a_ones_row > 0 ? g.a_ones_row : g.m;
        if (!g.idx && md <= 128 && (md & 3) == 0 && (g.k & 3) == 0 && g.a_cs == md && g.b_rs == g.n && (g.a_ones_row == 0 || g.m == md + 1) &&
            (((uintptr_t)g.a | (uintptr_t)g.b) & 15) == 0) {
            int64_t sl = (int64_t)sms * 3;                     // three blocks per SM (2 x 36 KB of stages each), all resident
            if (sl > g.k / 64) sl = g.k / 64;
            const int64_t per4 = mq_round_up(mq_ceil_div(g.k, sl), 8);
            sl = mq_ceil_div(g.k, per4);
            auto kd = g.n <= 10 ? gemm_skinny_dw_tma_kernel<10> : gemm_skinny_dw_tma_kernel<SKINNY_MAX>;
            if (sl * g.m * g.n * (int64_t)sizeof(float) <= ws_bytes &&
                cudaFuncSetAttribute(kd, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * SKDW_STAGE) == cudaSuccess) {
                MQ_LAUNCH(kd, (unsigned)sl, MQ_NT, 2 * SKDW_STAGE, stream, g, per4, (float *)ws);
                const int64_t tot = g.m * g.n;
                MQ_LAUNCH(gemm_skinny_finish_kernel, (unsigned)mq_ceil_div(tot * 32, MQ_NT), MQ_NT, 0, stream, (const float *)ws, (int)sl, tot, g.alpha, g.c);
                return 1;
            }
        }
        int64_t slabs = (int64_t)sms * 8;
        if (slabs > g.k / 32) slabs = g.k / 32;
        while (slabs > 1 && slabs * g.m * g.n * (int64_t)sizeof(float) > ws_bytes) --slabs;
        const int64_t per = mq_round_up(mq_ceil_div(g.k, slabs), 8);
        slabs = mq_ceil_div(g.k, per);
        if (slabs * g.m * g.n * (int64_t)sizeof(float) > ws_bytes) return 0;
        MQ_LAUNCH(gemm_skinny_dw_kernel, (unsigned)slabs, MQ_NT, 0, stream, g, per, (float *)ws);
        const int64_t total = g.m * g.n;
        MQ_LAUNCH(gemm_skinny_finish_kernel, (unsigned)mq_ceil_div(total * 32, MQ_NT), MQ_NT, 0, stream, (const float *)ws, (int)slabs, total, g.alpha, g.c);
        return 1;
    }
    return 0;
#endif
}

// Decide kernel, tile size and split-K; `ws` may be NULL (then no split and no tensor-core path).
static int gemm_run(GemmArgs g, void *ws, int64_t ws_bytes, int32_t tc, void *stream, const char *what) {
    if (g.m <= 0 || g.n <= 0) return MQ_OK;
    const int sms = mq_sm_count();
    const int planes = (tc & MQ_TC_PLANES_MASK) ? (tc & MQ_TC_PLANES_MASK) : 3;
    if (ws && gemm_tc_wanted(g, tc) && ws_bytes >= gemm_tc_ws_bytes(g.m, g.n, g.k, planes)) {
        // large problems: TMA-fed tcgen05 kernel on 128 x 128 tiles (mq_gemm_tma.cu), same epilogues / gather / split-K
        GemmArgs gt = g;
        const int64_t splits = gemm_tc_splits(g.m, g.n, g.k, &gt.k_per_split);
        const int64_t pack_bytes = mq_gemm_tma_pack_bytes(g.m, g.n, g.k, planes);
        float *partials = (float *)((char *)ws + pack_bytes);
        if (splits > 1) gt.c = partials;
        const int rc = mq_gemm_tma_launch(gt, planes | (tc & MQ_TC_FP16), splits, ws, stream);
        if (rc == MQ_OK) {
            if (splits > 1) {
                int64_t nb = mq_ceil_div(g.m * g.n, MQ_NT);
                if (nb > sms * 8) nb = sms * 8;
                MQ_LAUNCH(gemm_splitk_finish_kernel, (unsigned)nb, MQ_NT, 0, stream, g, (const float *)partials, (int)splits, g.c);
            }
            MQ_CHECK_LAUNCH(what);
            return MQ_OK;
        }
        if (rc != MQ_ERR_UNSUPPORTED) return rc;
    }
    MQ_REQUIRE(!(tc & MQ_TC_FORCE), MQ_ERR_INVALID, "%s: MQ_TC_FORCE needs a workspace of mq_gemm_ws_bytes()", what);
    if (gemm_skinny(g, ws, ws_bytes, stream)) {
        MQ_CHECK_LAUNCH(what);
        return MQ_OK;
    }
    const bool big = mq_ceil_div(g.m, 64) * mq_ceil_div(g.n, 64) >= sms / 2;
    const int bm = big ? 64 : 32;
    const int64_t tiles = mq_ceil_div(g.m, bm) * mq_ceil_div(g.n, bm);
    int64_t splits = 1;
    if (ws && tiles < sms && g.k >= 512) {
        splits = mq_ceil_div(2 * sms, tiles);
        if (splits > g.k / 128) splits = g.k / 128;
        if (splits > 64) splits = 64;
        while (splits > 1 && splits * g.m * g.n * (int64_t)sizeof(float) > ws_bytes) --splits;
        if (splits < 1) splits = 1;
    }
    g.k_per_split = mq_round_up(mq_ceil_div(g.k > 0 ? g.k : 1, splits), GEMM_BK);
    splits = mq_ceil_div(g.k > 0 ? g.k : 1, g.k_per_split);
    dim3 grid((unsigned)mq_ceil_div(g.n, bm), (unsigned)mq_ceil_div(g.m, bm), (unsigned)splits);
    float *c_final = g.c;
    if (splits > 1) {
        g.c = (float *)ws;
        if (big) { auto kfn = gemm_kernel<64, 64, true>; MQ_LAUNCH(kfn, grid, MQ_NT, 0, stream, g); }
        else { auto kfn = gemm_kernel<32, 32, true>; MQ_LAUNCH(kfn, grid, MQ_NT, 0, stream, g); }
        int64_t nb = mq_ceil_div(g.m * g.n, MQ_NT);
        if (nb > sms * 8) nb = sms * 8;
        MQ_LAUNCH(gemm_splitk_finish_kernel, (unsigned)nb, MQ_NT, 0, stream, g, (const float *)ws,
                  (int)splits, c_final);
    } else {
        if (big) { auto kfn = gemm_kernel<64, 64, false>; MQ_LAUNCH(kfn, grid, MQ_NT, 0, stream, g); }
        else { auto kfn = gemm_kernel<32, 32, false>; MQ_LAUNCH(kfn, grid, MQ_NT, 0, stream, g); }
    }
    MQ_CHECK_LAUNCH(what);
    return MQ_OK;
}

extern "C" int mq_gemm_ex(const float *a, int64_t a_rs, int64_t a_cs, const float *b, int64_t b_rs, int64_t b_cs, float *c,
                          int64_t m, int64_t n, int64_t k, float alpha, int32_t tc, void *ws, int64_t ws_bytes, void *stream) {
    MQ_REQUIRE(m >= 0 && n >= 0 && k >= 0, MQ_ERR_SHAPE, "mq_gemm: negative size");
    MQ_REQUIRE((m == 0 || n == 0) || (a && b && c), MQ_ERR_INVALID, "mq_gemm: null pointer");
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.a = a; g.a_rs = a_rs; g.a_cs = a_cs; g.b = b; g.b_rs = b_rs; g.b_cs = b_cs;
    g.c = c; g.m = m; g.n = n; g.k = k; g.alpha = alpha; g.epi = EPI_SCALE;
    return gemm_run(g, ws, ws ? ws_bytes : 0, tc, stream, "mq_gemm");
}

extern "C" int mq_gemm(const float *a, int64_t a_rs, int64_t a_cs, const float *b, int64_t b_rs,
                       int64_t b_cs, float *c, int64_t m, int64_t n, int64_t k, float alpha,
                       void *stream) {
    return mq_gemm_ex(a, a_rs, a_cs, b, b_rs, b_cs, c, m, n, k, alpha, 0, nullptr, 0, stream);
}

// The same with a workspace (mq_gemm_ws_bytes): large products run on the TMA-fed tcgen05 kernel, products with few output
// tiles and a long contraction (dW = x^T g over a batch of thousands) are split along K, deterministically.
extern "C" int mq_gemm_ws(const float *a, int64_t a_rs, int64_t a_cs, const float *b, int64_t b_rs,
                          int64_t b_cs, float *c, int64_t m, int64_t n, int64_t k, float alpha, void *ws,
                          int64_t ws_bytes, void *stream) {
    return mq_gemm_ex(a, a_rs, a_cs, b, b_rs, b_cs, c, m, n, k, alpha, 0, ws, ws_bytes, stream);
}

// ---- layered MLP path --------------------------------------------------------------
static int net_check(const mq_net_t *net, const char *what) {
    MQ_REQUIRE(net, MQ_ERR_INVALID, "%s: null net", what);
    MQ_REQUIRE(net->n_layers >= 1 && net->n_layers <= MQ_MAX_LAYERS, MQ_ERR_SHAPE,
               "%s: n_layers=%d out of range", what, net->n_layers);
    int k = net->n_in;
    for (int l = 0; l < net->n_layers; ++l) {
        MQ_REQUIRE(net->in[l] == k && net->out[l] >= 1 && k >= 1, MQ_ERR_SHAPE,
                   "%s: layer %d shape mismatch", what, l);
        MQ_REQUIRE(net->act[l] >= 0 && net->act[l] <= 2, MQ_ERR_INVALID, "%s: layer %d activation", what, l);
        k = net->out[l];
    }
    MQ_REQUIRE(k == net->n_out, MQ_ERR_SHAPE, "%s: n_out mismatch", what);
    return MQ_OK;
}

static int64_t net_max_width(const mq_net_t *net) {
    int64_t w = net->n_in;
    for (int l = 0; l < net->n_layers; ++l) if (net->out[l] > w) w = net->out[l];
    return w;
}

extern "C" int64_t mq_mlp_acts_floats(const mq_net_t *net, int64_t batch) {
    int64_t s = 0;
    for (int l = 0; l < net->n_layers; ++l) s += (int64_t)net->out[l] * batch;
    return s;
}

// per-product workspace of the layered path: the largest of forward / dW / dX over the layers
static int64_t mlp_gemm_ws(const mq_net_t *net, int64_t batch) {
    int64_t w = 256;
    for (int l = 0; l < net->n_layers; ++l) {
        const int64_t f = mq_gemm_ws_bytes(batch, net->out[l], net->in[l]);
        const int64_t dw = mq_gemm_ws_bytes(net->in[l] + 1, net->out[l], batch);
        const int64_t dx = mq_gemm_ws_bytes(batch, net->in[l], net->out[l]);
        if (f > w) w = f;
        if (dw > w) w = dw;
        if (dx > w) w = dx;
    }
    return mq_round_up(w, 1024);
}

// [gradient buffers: nbuf x batch x max width floats][workspace of the dW products][workspace of the forward / dX products]
static int64_t mlp_scratch_bytes(const mq_net_t *net, int64_t batch) {
    const int64_t nbuf = net->n_layers > 2 ? net->n_layers : 2;     // one gradient buffer per layer (overlapped backward)
    return mq_round_up(nbuf * batch * net_max_width(net) * (int64_t)sizeof(float), 1024) + 2 * mlp_gemm_ws(net, batch) + 256;
}
// MQ_TC_KEEP_X: the operand planes of the gathered input batch (+ a column of ones), packed by the forward pass and read
// again by dW_0 = x^T g of the backward pass: [3 planes at most][batch padded to 128][(in + 1) padded to 128] 16-bit words
static int64_t mlp_xplanes_rows(int64_t batch) { return mq_round_up(batch, 128); }
static int64_t mlp_xplanes_pitch(const mq_net_t *net) { return mq_round_up((int64_t)net->in[0] + 1, 128); }
static int64_t mlp_xplanes_bytes(const mq_net_t *net, int64_t batch) {
    return mq_round_up(3 * mlp_xplanes_rows(batch) * mlp_xplanes_pitch(net) * 2, 1024) + 1024;
}
static void *mlp_xplanes_ptr(const mq_net_t *net, int64_t batch, void *ws) {
    return (void *)(((uintptr_t)ws + (uintptr_t)mlp_scratch_bytes(net, batch) + 1023) & ~(uintptr_t)1023);
}
// the forward product of layer 0 takes the tensor-core path (and, with MQ_TC_KEEP_X, leaves the planes of x behind)
static bool mlp_keeps_x(const mq_net_t *net, int64_t batch, int32_t tc, void *ws) {
    if (!ws || !(tc & MQ_TC_KEEP_X)) return false;
    GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.m = batch; g.n = net->out[0]; g.k = net->in[0];
    return gemm_tc_wanted(g, tc);
}

// [gradient buffers: nbuf x batch x max width floats][workspace of the dW products][workspace of the forward / dX products]
// [planes of the input batch, MQ_TC_KEEP_X]
extern "C" int64_t mq_mlp_ws_bytes(const mq_net_t *net, int64_t batch) {
    return mlp_scratch_bytes(net, batch) + mlp_xplanes_bytes(net, batch);
}

extern "C" int mq_mlp_forward(const mq_net_t *net, const float *theta, const float *x,
                              const int32_t *row_idx, int64_t batch, float *acts, void *stream) {
    return mq_mlp_forward_ws(net, theta, x, row_idx, batch, acts, nullptr, 0, stream);
}

// The same with a workspace: layers whose output has fewer tiles than the GPU has SMs and a long contraction (the
// 784 -> 128 layer of mnist/mlp.py at batch 128: 16 tiles) are split along K, deterministically.
extern "C" int mq_mlp_forward_ws(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                                 int64_t batch, float *acts, void *ws, int64_t ws_bytes, void *stream) {
    return mq_mlp_forward_tc(net, theta, x, row_idx, batch, acts, 0, ws, ws_bytes, stream);
}

// The same with per-call MQ_TC_* flags for the layers that run on the tcgen05 GEMM (0 = three bf16 planes, any range;
// 2 | MQ_TC_FP16 = fp16x2: the same fp32 bar in half the tensor work and 2/3 of the packed bytes, for |operand| < 65504).
extern "C" int mq_mlp_forward_tc(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                                 int64_t batch, float *acts, int32_t tc, void *ws, int64_t ws_bytes, void *stream) {
    int rc = net_check(net, "mq_mlp_forward");
    if (rc != MQ_OK) return rc;
    MQ_REQUIRE(batch >= 0, MQ_ERR_SHAPE, "mq_mlp_forward: negative batch");
    if (batch == 0) return MQ_OK;
    MQ_REQUIRE(theta && x && acts, MQ_ERR_INVALID, "mq_mlp_forward: null pointer");
    const float *inp = x;
    float *out = acts;
    for (int l = 0; l < net->n_layers; ++l) {
        GemmArgs g;
        memset(&g, 0, sizeof(g));
        g.a = inp; g.a_rs = net->in[l]; g.a_cs = 1;
        g.idx = (l == 0) ? row_idx : nullptr; g.idx_on_k = 0;
        g.b = theta + net->w_off[l]; g.b_rs = net->out[l]; g.b_cs = 1;
        g.c = out; g.m = batch; g.n = net->out[l]; g.k = net->in[l];
        g.alpha = 1.0f; g.epi = EPI_BIAS_ACT; g.act = net->act[l]; g.leak = net->leak[l];
        g.bias = theta + net->b_off[l];
#ifndef MQ_EMU
        if (l == 0 && ws && ws_bytes >= mq_mlp_ws_bytes(net, batch) && mlp_keeps_x(net, batch, tc, ws)) {
            // pack the gathered batch ONCE per step: this product reads the planes K-major, dW_0 of the backward pass MN-major
            const int planes = (tc & MQ_TC_PLANES_MASK) ? (tc & MQ_TC_PLANES_MASK) : 3;
            GmPackJob jb;
            memset(&jb, 0, sizeof(jb));
            jb.base = x; jb.rs = net->in[0]; jb.cs = 1; jb.idx = row_idx; jb.idx_on_k = 0;
            jb.rows = batch; jb.k = (int64_t)net->in[0] + 1; jb.ones_col = net->in[0];
            jb.rows_pad = jb.rows_cover = mlp_xplanes_rows(batch);
            jb.k_pad = jb.k_cover = mlp_xplanes_pitch(net);
            jb.out = mlp_xplanes_ptr(net, batch, ws);
            rc = mq_gemm_tma_pack_multi(&jb, 1, planes | (tc & MQ_TC_FP16), stream);
            if (rc == MQ_OK) { g.a_planes = jb.out; g.a_planes_rows_pad = jb.rows_pad; g.a_planes_pitch = jb.k_pad; g.a_planes_mn = 0; }
            else if (rc != MQ_ERR_UNSUPPORTED) return rc;
        }
#endif
        rc = gemm_run(g, ws, ws ? ws_bytes : 0, tc, stream, "mq_mlp_forward");
        if (rc != MQ_OK) return rc;
        inp = out;
        out += (int64_t)net->out[l] * batch;
    }
    return MQ_OK;
}

extern "C" int mq_mlp_backward(const mq_net_t *net, const float *theta, const float *x,
                               const int32_t *row_idx, const float *acts, float *dy, int64_t batch,
                               float scale, float *grad, float *dx, void *ws, int64_t ws_bytes,
                               void *stream) {
    return mq_mlp_backward_overlap(net, theta, x, row_idx, acts, dy, batch, scale, grad, dx, ws, ws_bytes, stream, nullptr,
                                   nullptr);
}

// side_stream != NULL: the parameter-gradient products (dW_l, db_l: nothing downstream but the optimiser needs them) run
// on side_stream while the chain g_l -> g_{l-1} continues on `stream`; events[l] (l < L) orders dW_l after g_l,
// events[L] joins side_stream back into `stream`.  Works inside a stream capture (the graph gets the fork / join edges).
extern "C" int mq_mlp_backward_overlap(const mq_net_t *net, const float *theta, const float *x,
                                       const int32_t *row_idx, const float *acts, float *dy, int64_t batch,
                                       float scale, float *grad, float *dx, void *ws, int64_t ws_bytes,
                                       void *stream, void *side_stream, void *const *events) {
    return mq_mlp_backward_tc(net, theta, x, row_idx, acts, dy, batch, scale, grad, dx, 0, ws, ws_bytes, stream, side_stream, events);
}

// The same with per-call MQ_TC_* flags (see mq_mlp_forward_tc).
extern "C" int mq_mlp_backward_tc(const mq_net_t *net, const float *theta, const float *x,
                                  const int32_t *row_idx, const float *acts, float *dy, int64_t batch,
                                  float scale, float *grad, float *dx, int32_t tc, void *ws, int64_t ws_bytes,
                                  void *stream, void *side_stream, void *const *events) {
    int rc = net_check(net, "mq_mlp_backward");
    if (rc != MQ_OK) return rc;
    MQ_REQUIRE(batch >= 1, MQ_ERR_SHAPE, "mq_mlp_backward: empty batch");
    MQ_REQUIRE(theta && x && acts && dy && grad && ws, MQ_ERR_INVALID, "mq_mlp_backward: null pointer");
    MQ_REQUIRE(ws_bytes >= mq_mlp_ws_bytes(net, batch), MQ_ERR_SHAPE, "mq_mlp_backward: workspace too small");
    const int L = net->n_layers;
    const int64_t maxw = net_max_width(net);
    const int64_t nbuf = L > 2 ? L : 2;
    const int64_t per = mlp_gemm_ws(net, batch);
    char *split_ws = (char *)ws + mq_round_up(nbuf * batch * maxw * (int64_t)sizeof(float), 1024);     // dW products (side stream)
    char *main_ws = split_ws + per;                                                                     // dX products
    const int64_t split_bytes = per;
    MQ_REQUIRE(!side_stream || events, MQ_ERR_INVALID, "mq_mlp_backward_overlap: events missing");
#ifdef MQ_EMU
    side_stream = nullptr;
#endif
    void *w_stream = side_stream ? side_stream : stream;     // where the dW / db products go

    // offsets of each layer's saved output inside acts
    int64_t off[MQ_MAX_LAYERS];
    int64_t s = 0;
    for (int l = 0; l < L; ++l) { off[l] = s; s += (int64_t)net->out[l] * batch; }

    float *g_cur = dy;
    if (net->act[L - 1] != MQ_ACT_NONE) {
        rc = mq_act_backward(acts + off[L - 1], dy, dy, batch * net->out[L - 1], net->act[L - 1],
                             net->leak[L - 1], stream);
        if (rc != MQ_OK) return rc;
    }
    for (int l = L - 1; l >= 0; --l) {
        const float *inp = (l == 0) ? x : acts + off[l - 1];
#ifndef MQ_EMU
        if (side_stream) {                                   // g_l is complete on `stream`: let the side stream see it
            if (cudaEventRecord((cudaEvent_t)events[l], (cudaStream_t)stream) != cudaSuccess ||
                cudaStreamWaitEvent((cudaStream_t)side_stream, (cudaEvent_t)events[l], 0) != cudaSuccess) {
                mq_set_error("mq_mlp_backward_overlap: %s", cudaGetErrorString(cudaGetLastError()));
                return MQ_ERR_CUDA;
            }
        }
#endif
        // dW = scale * inp^T g      A(m=in, k=batch) = inp[k*in + m]
        GemmArgs g;
        memset(&g, 0, sizeof(g));
        g.a = inp; g.a_rs = 1; g.a_cs = net->in[l];
        g.idx = (l == 0) ? row_idx : nullptr; g.idx_on_k = 1;
        g.b = g_cur; g.b_rs = net->out[l]; g.b_cs = 1;
        g.c = grad + net->w_off[l]; g.m = net->in[l]; g.n = net->out[l]; g.k = batch;
        g.alpha = scale; g.epi = EPI_SCALE;
        // the bias gradient 1^T g is one more row of the same product (a row of ones appended to x^T), written straight
        // into b's slot behind W
        const bool fuse_bias = net->b_off[l] == net->w_off[l] + net->in[l] * net->out[l];
        if (fuse_bias) { g.m = net->in[l] + 1; g.a_ones_row = net->in[l]; }
        if (l == 0 && mlp_keeps_x(net, batch, tc, ws)) {     // the forward pass of this step left the planes of x (+ ones) behind
            g.a_planes = mlp_xplanes_ptr(net, batch, ws); g.a_planes_rows_pad = mlp_xplanes_rows(batch);
            g.a_planes_pitch = mlp_xplanes_pitch(net); g.a_planes_mn = 1;
        }
        rc = gemm_run(g, split_ws, split_bytes, tc, w_stream, "mq_mlp_backward(dW)");
        if (rc != MQ_OK) return rc;
        if (!fuse_bias) {
            rc = mq_col_sum(g_cur, nullptr, batch, net->out[l], scale, grad + net->b_off[l], w_stream);
            if (rc != MQ_OK) return rc;
        }
        if (l > 0 || dx) {
            // g_prev = (g W^T) * act'(saved output of layer l-1)     B(k=out, n=in) = W[n*out + k]
            float *dst = (l == 0) ? dx : (float *)ws + (int64_t)((l - 1) % nbuf) * batch * maxw;   // g_{l-1}: own buffer
            memset(&g, 0, sizeof(g));
            g.a = g_cur; g.a_rs = net->out[l]; g.a_cs = 1;
            g.b = theta + net->w_off[l]; g.b_rs = 1; g.b_cs = net->out[l];
            g.c = dst; g.m = batch; g.n = net->in[l]; g.k = net->out[l];
            g.alpha = 1.0f;
            if (l > 0 && net->act[l - 1] != MQ_ACT_NONE) {
                g.epi = EPI_ACT_GRAD; g.act = net->act[l - 1]; g.leak = net->leak[l - 1];
                g.y = acts + off[l - 1];
            } else {
                g.epi = EPI_SCALE;
            }
            rc = gemm_run(g, main_ws, per, tc, stream, "mq_mlp_backward(dX)");
            if (rc != MQ_OK) return rc;
            g_cur = dst;
        }
    }
#ifndef MQ_EMU
    if (side_stream) {
        if (cudaEventRecord((cudaEvent_t)events[L], (cudaStream_t)side_stream) != cudaSuccess ||
            cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)events[L], 0) != cudaSuccess) {
            mq_set_error("mq_mlp_backward_overlap: %s", cudaGetErrorString(cudaGetLastError()));
            return MQ_ERR_CUDA;
        }
    }
#endif
    return MQ_OK;
}
