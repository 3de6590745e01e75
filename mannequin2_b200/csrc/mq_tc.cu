// Tensor-core (tcgen05 / TMEM) fused forward + head + backward for small MLPs at LARGE batch.
//
// Replaces, for a whole minibatch in one launch, the same reference sequence as mq_fused.cu
//   Function.evaluate     mannequin/basicnet.py:192-207
//   head                  benchmarks/gae.py:22, benchmarks/ppo.py:34-37, mannequin/distrib.py:61-68
//   backprop              mannequin/basicnet.py:209-257 + mannequin/backprop.py:11-60
// when the batch is large enough for the Affine layers to be real dense contractions
// (65,536-row minibatches of BASELINE configs[2]): every product -- Z = A W, dA = dZ W^T,
// dW = A^T dZ and the bias column sums -- is a tcgen05.mma whose accumulator lives in TMEM.
//
// fp32 parity on bf16 tensor cores: every operand is split exactly into THREE bf16 planes
// (x = x0 + x1 + x2, 24 mantissa bits) and a product is the six plane pairs with i + j <= 2,
// accumulated in fp32 in TMEM; the dropped terms are below 2^-24 |a||b|.
//
// One CTA (256 threads) per SM walks tiles of <= 128 batch rows:
//   smem  operands as 8 x 16 B core matrices (SWIZZLE_NONE).  An activation buffer is
//         [plane][feature chunk of 8][row] so that the SAME bytes serve as the K-major A operand
//         of the next layer / of dA (rows = M) and as the MN-major operand of dW (rows = K);
//         weights are staged once per launch as [plane][k chunk][n], K-major for the forward
//         product and MN-major (same bytes) for dA = dZ W^T.
//   TMEM  lane = batch row.  Z_l is overwritten in place by the fp32 activation (kept for the
//         backward pass), dA_l has its own columns, the dW / db accumulators (M = 64: row i in
//         lane i%16 + 32*(i/16)) persist across all tiles of the CTA and are read once.
//   dZ_l overwrites A_{l+1} in shared memory in place (same thread, same element), after the
//   commit that covers the last MMA reading A_{l+1}.
// MMAs are issued by one thread; completion is a tcgen05.commit on one mbarrier per phase.
// Per-CTA partial gradients go to the workspace and are summed in fixed order by
// fused_reduce_kernel (mq_fused.cu): bit-reproducible run to run.
#include "mq_common.cuh"

#ifndef MQ_EMU
#include <cuda_bf16.h>

#define HALF_LOG_2PI 0.91893853320467274178f
#define TC_NT 256
#define TC_ROWS 128              // MMA M
#define TC_HID 64                // padded hidden width
#define TC_NO 16                 // padded output width
#define TC_PLANES 3
#define TC_CHUNK_BYTES 2048      // one 8-feature chunk of all 128 rows (128 rows x 16 B)
#define TC_MAXL 3

struct TcPlan {
    mq_net_t net;
    int L;
    int k0;                      // padded input width (16 or 32)
    int kp[TC_MAXL], np[TC_MAXL];
    int tile_rows;               // rows per tile, multiple of 16, <= 128
    int with_grad;
    // shared-memory byte offsets
    int off_dzl, off_act[TC_MAXL], act_plane[TC_MAXL], off_ones, off_w[TC_MAXL], w_plane[TC_MAXL];
    int off_bias, off_misc;
    int smem_bytes;
    // TMEM columns
    int col_z[TC_MAXL], col_da[TC_MAXL], col_dw[TC_MAXL], col_db[TC_MAXL];
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// SWIZZLE_NONE shared-memory matrix descriptor, split in two words so that stepping through K / planes is
// one 32-bit add: lo = start >> 4 | (leading byte offset >> 4) << 16, hi = stride byte offset >> 4 |
// descriptor version 1 (Blackwell) at bit 46.
__device__ __forceinline__ uint32_t tc_desc_lo(uint32_t addr, uint32_t lbo) { return ((addr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ uint32_t tc_desc_hi(uint32_t sbo) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14); }

// kind::f16 instruction descriptor: D fp32, A/B bf16, majors (0 = K, 1 = MN), N >> 3, M >> 4
__device__ __forceinline__ uint32_t tc_idesc(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi,
                                       uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void tc_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

__device__ __forceinline__ void tc_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// every thread: make generic-proxy smem writes visible to the tensor core, order TMEM accesses, meet
__device__ __forceinline__ void tc_publish() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

__device__ __forceinline__ void tc_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tc_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tc_st32(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// ---- exact three-way bf16 split -------------------------------------------------------------
// x = p0 + p1 + p2 with each p a bf16 (round to nearest): the residuals are exact in fp32.
__device__ __forceinline__ void tc_split_pair(float a, float b, uint32_t &w0, uint32_t &w1, uint32_t &w2) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(a, b);
    w0 = *reinterpret_cast<uint32_t *>(&h0);
    float ra = a - __uint_as_float(w0 << 16), rb = b - __uint_as_float(w0 & 0xFFFF0000u);
    __nv_bfloat162 h1 = __floats2bfloat162_rn(ra, rb);
    w1 = *reinterpret_cast<uint32_t *>(&h1);
    ra -= __uint_as_float(w1 << 16);
    rb -= __uint_as_float(w1 & 0xFFFF0000u);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(ra, rb);
    w2 = *reinterpret_cast<uint32_t *>(&h2);
}

// store 8 consecutive features of one row as one 16-byte chunk in each of the three planes
__device__ __forceinline__ void tc_store_chunk(unsigned char *buf, int plane_bytes, int chunk, int row,
                                               const float *v) {
    uint32_t w[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tc_split_pair(v[2 * j], v[2 * j + 1], w[0][j], w[1][j], w[2][j]);
    unsigned char *p = buf + chunk * TC_CHUNK_BYTES + (row >> 3) * 128 + (row & 7) * 16;
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
        *reinterpret_cast<uint4 *>(p + pl * plane_bytes) = make_uint4(w[pl][0], w[pl][1], w[pl][2], w[pl][3]);
}

// ---- MMA issue helpers (one thread) -----------------------------------------------------------
// (called by the elected lane of a warp inside warp-uniform control flow, so that descriptors live in uniform registers)
struct TcOperand {
    uint32_t lo, hi;      // descriptor words of plane 0, K step 0
    uint32_t plane16;     // (bytes between planes) >> 4
    uint32_t k16;         // (bytes per K step of 16 elements) >> 4
    int planes;           // 3, or 1 for the ones matrix
};

// D (+)= A * B over the plane pairs with i + j <= 2 (largest terms first)
__device__ __forceinline__ void tc_product(uint32_t d_tmem, const TcOperand &a, const TcOperand &b,
                                           uint32_t idesc, int ksteps, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
#pragma unroll
    for (int s = 0; s < 3; ++s) {
#pragma unroll
        for (int i = 0; i <= s; ++i) {
            const int j = s - i;
            if (i >= a.planes || j >= b.planes) continue;
            uint32_t alo = a.lo + i * a.plane16, blo = b.lo + j * b.plane16;
#pragma unroll 1
            for (int k = 0; k < ksteps; ++k) {
                tc_mma(d_tmem, alo, a.hi, blo, b.hi, idesc, acc);
                alo += a.k16;
                blo += b.k16;
                acc = 1u;
            }
        }
    }
}

// activation buffer viewed K-major (M = rows, K = features): LBO = chunk stride, SBO = row-group stride
__device__ __forceinline__ TcOperand tc_act_k(uint32_t base, int plane) {
    return TcOperand{tc_desc_lo(base, TC_CHUNK_BYTES), tc_desc_hi(128u), (uint32_t)plane >> 4, (2u * TC_CHUNK_BYTES) >> 4, 3};
}
// activation buffer viewed MN-major (M or N = features, K = rows): LBO = row-group stride, SBO = chunk stride
__device__ __forceinline__ TcOperand tc_act_mn(uint32_t base, int plane) {
    return TcOperand{tc_desc_lo(base, 128u), tc_desc_hi(TC_CHUNK_BYTES), (uint32_t)plane >> 4, 256u >> 4, 3};
}
// weight image [k chunk][n][8 k] viewed K-major (N = out, K = in): forward
__device__ __forceinline__ TcOperand tc_w_k(uint32_t base, int plane, int np) {
    return TcOperand{tc_desc_lo(base, (uint32_t)np * 16u), tc_desc_hi(128u), (uint32_t)plane >> 4, ((uint32_t)np * 32u) >> 4, 3};
}
// same bytes viewed MN-major (N' = in, K' = out): dA = dZ W^T
__device__ __forceinline__ TcOperand tc_w_mn(uint32_t base, int plane, int np) {
    return TcOperand{tc_desc_lo(base, 128u), tc_desc_hi((uint32_t)np * 16u), (uint32_t)plane >> 4, 256u >> 4, 3};
}

// ---- the kernel ---------------------------------------------------------------------------
template <int K0>
__global__ void __launch_bounds__(TC_NT, 1)
tc_grad_kernel(TcPlan p, const float *__restrict__ theta, const float *__restrict__ x,
               const int32_t *__restrict__ idx, int64_t n_rows, mq_head_t head,
               float *__restrict__ partial, float *__restrict__ out, float *__restrict__ logp_out, long long *prof) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_mbar;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp-uniform for the compiler too
    // developer aid: per-phase cycles of CTA 0 (prof is NULL in normal use)
    long long t_last = 0;
    const bool profiling = prof != nullptr && blockIdx.x == 0 && tid == 0;
    if (profiling) t_last = clock64();
#define TC_MARK(i) do { if (profiling) { long long now_ = clock64(); prof[i] += now_ - t_last; t_last = now_; } } while (0)
    const int L = p.L;
    const int n_in = p.net.n_in, n_out = p.net.n_out;
    constexpr int XCH = K0 / 16;                       // 8-feature chunks per thread of the input tile

    // ---- one-time setup: TMEM, barrier, operand images of the weights --------------------------
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&s_mbar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = tid; i < p.smem_bytes / 16; i += TC_NT) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    {
        __nv_bfloat16 one = __float2bfloat16_rn(1.0f);
        uint16_t ob = *reinterpret_cast<uint16_t *>(&one);
        for (int i = tid; i < TC_CHUNK_BYTES / 2; i += TC_NT) reinterpret_cast<uint16_t *>(smem + p.off_ones)[i] = ob;
    }
    float *sBias = reinterpret_cast<float *>(smem + p.off_bias);      // [L][64]
    float *sMisc = reinterpret_cast<float *>(smem + p.off_misc);      // logstd[16], exp(-logstd)[16]
    for (int l = 0; l < L; ++l) {
        const int in = p.net.in[l], on = p.net.out[l], np = p.np[l];
        const float *w = theta + p.net.w_off[l];
        for (int e = tid; e < in * on; e += TC_NT) {
            const int k = e / on, n = e - k * on;
            uint32_t w0, w1, w2;
            tc_split_pair(w[e], 0.0f, w0, w1, w2);
            unsigned char *dst = smem + p.off_w[l] + (k >> 3) * (np * 16) + (n >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2;
            *reinterpret_cast<uint16_t *>(dst) = (uint16_t)(w0 & 0xFFFFu);
            *reinterpret_cast<uint16_t *>(dst + p.w_plane[l]) = (uint16_t)(w1 & 0xFFFFu);
            *reinterpret_cast<uint16_t *>(dst + 2 * p.w_plane[l]) = (uint16_t)(w2 & 0xFFFFu);
        }
        for (int n = tid; n < on; n += TC_NT) sBias[l * 64 + n] = theta[p.net.b_off[l] + n];
    }
    if (p.net.logstd_off >= 0 && tid < n_out) {
        const float ls = theta[p.net.logstd_off + tid];
        sMisc[tid] = ls;
        sMisc[16 + tid] = expf(-ls);
    }
    tc_publish();
    TC_MARK(0);
    const uint32_t tm = __shfl_sync(0xffffffffu, s_tmem, 0);
    const uint32_t mbar = tc_smem_u32(&s_mbar);
    const uint32_t sbase = tc_smem_u32(smem);
    uint32_t parity = 0;
    bool pending = false;                               // a committed MMA batch nobody waited for yet
    bool dw_started = false;                            // dW / db accumulators hold data

    const int row = (warp & 3) * 32 + lane;             // TMEM lane == batch row of the tile
    const int half = warp >> 2;                         // which 32 of the 64 columns this thread owns
    const uint32_t tlane = tm + ((uint32_t)((warp & 3) * 32) << 16);
    const int tr = p.tile_rows;
    const int64_t ntiles = (n_rows + tr - 1) / tr;
    const int dw_ksteps = tr / 16;
    const bool gauss = head.kind == MQ_HEAD_GAUSS_LOGP || head.kind == MQ_HEAD_GAUSS_PPO;
    const int lastact = p.net.act[L - 1];
    const float lastleak = p.net.leak[L - 1];

    // input-tile staging: thread (row = tid & 127) owns chunks (tid >> 7) + 2 j
    const int xrow = tid & 127;
    float xr[XCH][8];
    auto load_x = [&](int64_t tile) {
        const int64_t gr = tile * tr + xrow;
        const bool valid = xrow < tr && gr < n_rows;
        const int64_t src = valid ? (idx ? (int64_t)idx[gr] : gr) : 0;
        const float *xp = x + src * n_in;
#pragma unroll
        for (int j = 0; j < XCH; ++j) {
            const int c = (tid >> 7) + 2 * j;
#pragma unroll
            for (int f = 0; f < 8; ++f) {
                const int col = c * 8 + f;
                xr[j][f] = (valid && col < n_in) ? __ldg(xp + col) : 0.0f;
            }
        }
    };
    if ((int64_t)blockIdx.x < ntiles) load_x(blockIdx.x);

    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t row0 = tile * tr;
        const int64_t gr = row0 + row;
        const bool rvalid = row < tr && gr < n_rows;
        const int64_t hsrc = rvalid ? (idx ? (int64_t)idx[gr] : gr) : 0;
        if (pending) { tc_wait(mbar, parity); parity ^= 1; pending = false; }
        TC_MARK(1);
        // ---- stage the input tile ---------------------------------------------------------------
#pragma unroll
        for (int j = 0; j < XCH; ++j)
            tc_store_chunk(smem + p.off_act[0], p.act_plane[0], (tid >> 7) + 2 * j, xrow, xr[j]);
        tc_publish();
        TC_MARK(2);
        // ---- forward ------------------------------------------------------------------------------
        float tgt[TC_NO];                               // head inputs, fetched while the MMAs run
        float h_adv = 0.0f, h_lpold = 0.0f;
        for (int l = 0; l < L; ++l) {
            if (warp == 0) {
              if (tc_elect_one()) {
                tc_product(tm + p.col_z[l], tc_act_k(sbase + p.off_act[l], p.act_plane[l]),
                           tc_w_k(sbase + p.off_w[l], p.w_plane[l], p.np[l]),
                           tc_idesc(TC_ROWS, p.np[l], 0, 0), p.kp[l] / 16, false);
                tc_commit(mbar);
              }
              __syncwarp();
            }
            TC_MARK(3);
            if (l == 0) {
                // overlap global loads with the first products: next tile's rows, this tile's head inputs
                if (tile + gridDim.x < ntiles) load_x(tile + gridDim.x);
                if (half == 0) {
                    const float *mat = (head.kind == MQ_HEAD_DY) ? head.dy : head.target;
#pragma unroll
                    for (int n = 0; n < TC_NO; ++n) tgt[n] = (rvalid && n < n_out && mat) ? __ldg(mat + hsrc * n_out + n) : 0.0f;
                    if (head.kind == MQ_HEAD_GAUSS_PPO) {
                        h_adv = rvalid ? __ldg(head.adv + hsrc) : 0.0f;
                        h_lpold = rvalid ? __ldg(head.logp_old + hsrc) : 0.0f;
                    } else if (head.kind == MQ_HEAD_GAUSS_LOGP) {
                        h_adv = (rvalid && head.dy) ? __ldg(head.dy + hsrc) : 0.0f;
                    }
                }
            }
            TC_MARK(4);
            tc_wait(mbar, parity); parity ^= 1;
            TC_MARK(5 + l);
            if (l < L - 1) {
                // A_{l+1} = act(Z_l + b_l): fp32 copy back to TMEM, three bf16 planes to shared memory
                float v[32];
                tc_ld32(tlane + p.col_z[l] + half * 32, v);
                const float *b = sBias + l * 64 + half * 32;
                const int act = p.net.act[l];
                const float leak = p.net.leak[l];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = mq_act_apply(v[j] + b[j], act, leak);
                if (p.with_grad) tc_st32(tlane + p.col_z[l] + half * 32, v);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    tc_store_chunk(smem + p.off_act[l + 1], p.act_plane[l + 1], half * 4 + c, row, v + 8 * c);
                tc_publish();
                TC_MARK(8 + l);
            }
        }
        // ---- head: loss-gradient rule on the output row (one thread per row, warps 0-3) ------------
        if (half == 0) {
            float y[TC_NO], dz[TC_NO];
            tc_ld16(tlane + p.col_z[L - 1], y);
            const float *b = sBias + (L - 1) * 64;
#pragma unroll
            for (int n = 0; n < TC_NO; ++n) {
                y[n] = (n < n_out) ? mq_act_apply(y[n] + b[n], lastact, lastleak) : 0.0f;
                dz[n] = 0.0f;
            }
            if (rvalid && out) {
#pragma unroll
                for (int n = 0; n < TC_NO; ++n) if (n < n_out) out[gr * n_out + n] = y[n];
            }
            if (rvalid) {
                if (head.kind == MQ_HEAD_DY) {
#pragma unroll
                    for (int n = 0; n < TC_NO; ++n) if (n < n_out) dz[n] = mq_act_grad(y[n], tgt[n], lastact, lastleak);
                } else if (head.kind == MQ_HEAD_DIFF) {
#pragma unroll
                    for (int n = 0; n < TC_NO; ++n) if (n < n_out) dz[n] = mq_act_grad(y[n], tgt[n] - y[n], lastact, lastleak);
                } else if (head.kind == MQ_HEAD_SOFTMAX_CE) {
                    float mx = y[0];
#pragma unroll
                    for (int n = 1; n < TC_NO; ++n) if (n < n_out) mx = fmaxf(mx, y[n]);
                    float sum = 0.0f;
#pragma unroll
                    for (int n = 0; n < TC_NO; ++n) if (n < n_out) sum += expf(y[n] - mx);
                    const float inv = 1.0f / sum;
#pragma unroll
                    for (int n = 0; n < TC_NO; ++n)
                        if (n < n_out) dz[n] = mq_act_grad(y[n], tgt[n] - expf(y[n] - mx) * inv - head.p0 * y[n], lastact, lastleak);
                } else if (gauss) {
                    // diagonal Gaussian (distrib.py:61-68): n_out <= 8; columns 8.. carry the logstd terms
                    float acc = 0.0f;
#pragma unroll
                    for (int a = 0; a < 8; ++a)
                        if (a < n_out) {
                            const float z = (tgt[a] - y[a]) * sMisc[16 + a];
                            acc += sMisc[a] + 0.5f * z * z;
                        }
                    const float lp = -(float)n_out * HALF_LOG_2PI - acc;
                    if (logp_out) logp_out[gr] = lp;
                    float w;
                    if (head.kind == MQ_HEAD_GAUSS_PPO) {
                        const float ratio = expf(lp - h_lpold);
                        const bool kill = (ratio > head.p1 && h_adv > 0.0f) || (ratio < head.p0 && h_adv < 0.0f);
                        w = kill ? 0.0f : ratio * h_adv;
                    } else {
                        w = h_adv;
                    }
#pragma unroll
                    for (int a = 0; a < 8; ++a)
                        if (a < n_out) {
                            const float inv = sMisc[16 + a];
                            const float z = (tgt[a] - y[a]) * inv;
                            dz[8 + a] = w * (z * z - 1.0f);
                            dz[a] = mq_act_grad(y[a], w * z * inv, lastact, lastleak);
                        }
                }
            }
            if (p.with_grad) {
                tc_store_chunk(smem + p.off_dzl, TC_NO / 8 * TC_CHUNK_BYTES, 0, row, dz);
                tc_store_chunk(smem + p.off_dzl, TC_NO / 8 * TC_CHUNK_BYTES, 1, row, dz + 8);
            }
        }
        if (!p.with_grad) { tc_publish(); continue; }
        tc_publish();
        TC_MARK(11);
        // ---- backward -----------------------------------------------------------------------------
        for (int l = L - 1; l >= 0; --l) {
            // dZ_l lives in dzl (l == L-1) or in act[l+1] (in place of A_{l+1})
            const uint32_t dz_base = sbase + (l == L - 1 ? p.off_dzl : p.off_act[l + 1]);
            const int dz_plane = (l == L - 1) ? TC_NO / 8 * TC_CHUNK_BYTES : p.act_plane[l + 1];
            if (warp == 0) {
              if (tc_elect_one()) {
                if (l > 0)      // dA_l = dZ_l W_l^T  (M = rows, N = in_l = 64, K = out_l)
                    tc_product(tm + p.col_da[l], tc_act_k(dz_base, dz_plane), tc_w_mn(sbase + p.off_w[l], p.w_plane[l], p.np[l]),
                               tc_idesc(TC_ROWS, p.kp[l], 0, 1), p.np[l] / 16, false);
                if (l > 0)      // dW_l += A_l^T dZ_l   (M = in_l = 64, N = out_l, K = rows)
                    tc_product(tm + p.col_dw[l], tc_act_mn(sbase + p.off_act[l], p.act_plane[l]), tc_act_mn(dz_base, dz_plane),
                               tc_idesc(64, p.np[l], 1, 1), dw_ksteps, dw_started);
                else            // dW_0^T += dZ_0^T A_0 (M = out_0 = 64, N = k0, K = rows)
                    tc_product(tm + p.col_dw[0], tc_act_mn(dz_base, dz_plane), tc_act_mn(sbase + p.off_act[0], p.act_plane[0]),
                               tc_idesc(64, K0, 1, 1), dw_ksteps, dw_started);
                // db_l += dZ_l^T 1        (M = 64 feature rows of dZ_l, N = 8 identical columns)
                tc_product(tm + p.col_db[l], tc_act_mn(dz_base, dz_plane),
                           TcOperand{tc_desc_lo(sbase + (uint32_t)p.off_ones, 128u), tc_desc_hi(TC_CHUNK_BYTES), 0u, 256u >> 4, 1},
                           tc_idesc(64, 8, 1, 1), dw_ksteps, dw_started);
                tc_commit(mbar);
              }
              __syncwarp();
            }
            TC_MARK(12);
            if (l == 0) { pending = true; break; }
            tc_wait(mbar, parity); parity ^= 1;
            TC_MARK(13 + l);
            // dZ_{l-1} = dA_l * act'(A_l), written over A_l
            float a[32], g[32];
            tc_ld32(tlane + p.col_z[l - 1] + half * 32, a);
            tc_ld32(tlane + p.col_da[l] + half * 32, g);
            const int act = p.net.act[l - 1];
            const float leak = p.net.leak[l - 1];
#pragma unroll
            for (int j = 0; j < 32; ++j) g[j] = mq_act_grad(a[j], g[j], act, leak);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                tc_store_chunk(smem + p.off_act[l], p.act_plane[l], half * 4 + c, row, g + 8 * c);
            tc_publish();
            TC_MARK(16 + l);
        }
        dw_started = true;
    }
    if (pending) { tc_wait(mbar, parity); parity ^= 1; }
    TC_MARK(19);

    // ---- per-CTA partial gradient: TMEM accumulators -> workspace -----------------------------------
    if (p.with_grad) {
        float *dst = partial + (int64_t)blockIdx.x * p.net.n_params;
        if (!dw_started) {
            for (int j = tid; j < p.net.n_params; j += TC_NT) dst[j] = 0.0f;
        } else if (warp < 4) {
            // M = 64 accumulators: row i is lane i%16 of warp i/16
            const int i = warp * 16 + lane;
            const bool lv = lane < 16;
            for (int l = 0; l < L; ++l) {
                const int in = p.net.in[l], on = p.net.out[l];
                for (int c0 = 0; c0 < (l == 0 ? K0 : p.np[l]); c0 += 16) {
                    float v[16];
                    tc_ld16(tlane + p.col_dw[l] + c0, v);
                    if (!lv) continue;
                    if (l == 0) {           // D[o][k] = dW_0[k][o]
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (i < on && c0 + j < in) dst[p.net.w_off[0] + (c0 + j) * on + i] = v[j];
                    } else {                // D[k][o]
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (i < in && c0 + j < on) dst[p.net.w_off[l] + i * on + c0 + j] = v[j];
                    }
                }
                float v[16];
                tc_ld16(tlane + p.col_db[l], v);     // 8 identical columns (+ 8 unrelated ones)
                if (lv && i < on) dst[p.net.b_off[l] + i] = v[0];
                if (l == L - 1 && p.net.logstd_off >= 0 && lv && i >= 8 && i < 8 + n_out)
                    dst[p.net.logstd_off + i - 8] = gauss ? v[0] : 0.0f;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TC_MARK(20);
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}

// ---- host side ------------------------------------------------------------------------------------
static int tc_plan_build(const mq_net_t *net, int64_t batch, int with_grad, const mq_head_t *head, TcPlan *p) {
    memset(p, 0, sizeof(*p));
    p->net = *net;
    const int L = net->n_layers;
    if (L < 2 || L > TC_MAXL) return 0;
    if (net->n_in > 32 || net->n_out > TC_NO) return 0;
    for (int l = 0; l + 1 < L; ++l) if (net->out[l] > TC_HID) return 0;
    const bool gauss = head && (head->kind == MQ_HEAD_GAUSS_LOGP || head->kind == MQ_HEAD_GAUSS_PPO);
    if ((gauss || net->logstd_off >= 0) && net->n_out > 8) return 0;
    p->L = L;
    p->k0 = net->n_in <= 16 ? 16 : 32;
    for (int l = 0; l < L; ++l) {
        p->kp[l] = l == 0 ? p->k0 : TC_HID;
        p->np[l] = l == L - 1 ? TC_NO : TC_HID;
    }
    p->with_grad = with_grad;
    // balance: every CTA of a full wave gets the same number of tiles
    const int64_t sms = mq_sm_count();
    const int64_t waves = mq_ceil_div(batch, sms * TC_ROWS);
    int64_t tr = mq_round_up(mq_ceil_div(batch, sms * waves), 16);
    if (tr > TC_ROWS) tr = TC_ROWS;
    if (tr < 16) tr = 16;
    p->tile_rows = (int)tr;
    int off = 0;
    p->off_dzl = off; off += TC_PLANES * (TC_NO / 8) * TC_CHUNK_BYTES;
    for (int l = 0; l < L; ++l) {
        p->act_plane[l] = (p->kp[l] / 8) * TC_CHUNK_BYTES;
        p->off_act[l] = off; off += TC_PLANES * p->act_plane[l];
    }
    p->off_ones = off; off += TC_CHUNK_BYTES;
    for (int l = 0; l < L; ++l) {
        p->w_plane[l] = p->np[l] * p->kp[l] * 2;
        p->off_w[l] = off; off += TC_PLANES * p->w_plane[l];
    }
    p->off_bias = off; off += TC_MAXL * 64 * 4;
    p->off_misc = off; off += 32 * 4;
    p->smem_bytes = (int)mq_round_up(off, 16);
    int col = 0;
    for (int l = 0; l < L; ++l) { p->col_z[l] = col; col += p->np[l]; }
    for (int l = 1; l < L; ++l) { p->col_da[l] = col; col += TC_HID; }
    for (int l = 0; l < L; ++l) { p->col_dw[l] = col; col += (l == 0 ? p->k0 : p->np[l]); }
    for (int l = 0; l < L; ++l) { p->col_db[l] = col; col += 16; }
    if (col > 512) return 0;
    if (p->smem_bytes + 1024 > mq_smem_optin()) return 0;
    return 1;
}

long long *mq_phase_prof_ptr();
static int g_tc_mode = 1;            // 0: never, 1: automatic (large batches), 2: whenever the net fits
extern "C" int mq_debug_set_tc_mode(int mode) { g_tc_mode = mode; return MQ_OK; }

// Launches the tensor-core kernel when the net fits and the batch is large enough.
// Returns MQ_ERR_UNSUPPORTED (without an error message side effect that matters) otherwise.
// with_grad: partials go to ws ([grid][n_params]); *grid_out = CTAs launched.
int mq_tc_launch(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                 int64_t batch, const mq_head_t *head, int with_grad, float *out, float *logp,
                 void *ws, int64_t ws_bytes, int *grid_out, void *stream) {
    if (g_tc_mode == 0) return MQ_ERR_UNSUPPORTED;
    if (g_tc_mode == 1 && batch < 8192) return MQ_ERR_UNSUPPORTED;
    TcPlan p;
    if (!tc_plan_build(net, batch, with_grad, head, &p)) return MQ_ERR_UNSUPPORTED;
    const int64_t ntiles = mq_ceil_div(batch, p.tile_rows);
    const int64_t sms = mq_sm_count();
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    if (with_grad)
        MQ_REQUIRE(ws && ws_bytes >= (int64_t)grid * net->n_params * (int64_t)sizeof(float), MQ_ERR_INVALID,
                   "mq_tc_launch: workspace too small");
    cudaError_t e;
    if (p.k0 == 16) {
        e = cudaFuncSetAttribute(tc_grad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes);
        if (e == cudaSuccess)
            tc_grad_kernel<16><<<grid, TC_NT, p.smem_bytes, (cudaStream_t)stream>>>(p, theta, x, row_idx, batch, *head,
                                                                                   (float *)ws, out, logp, mq_phase_prof_ptr());
    } else {
        e = cudaFuncSetAttribute(tc_grad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes);
        if (e == cudaSuccess)
            tc_grad_kernel<32><<<grid, TC_NT, p.smem_bytes, (cudaStream_t)stream>>>(p, theta, x, row_idx, batch, *head,
                                                                                   (float *)ws, out, logp, mq_phase_prof_ptr());
    }
    if (e != cudaSuccess) {
        mq_set_error("mq_tc_launch: %s", cudaGetErrorString(e));
        return MQ_ERR_CUDA;
    }
    MQ_CHECK_LAUNCH("mq_tc_launch");
    if (grid_out) *grid_out = grid;
    return MQ_OK;
}
#else   // MQ_EMU: the host emulation has no tensor cores
int mq_tc_launch(const mq_net_t *, const float *, const float *, const int32_t *, int64_t, const mq_head_t *, int,
                 float *, float *, void *, int64_t, int *, void *) { return MQ_ERR_UNSUPPORTED; }
extern "C" int mq_debug_set_tc_mode(int) { return MQ_OK; }
#endif
