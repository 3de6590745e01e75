// Tensor-core (tcgen05 / TMEM) fused forward + head + backward for small MLPs at LARGE batch, third generation.
//
// Same reference sequence as mq_fused.cu / mq_tc.cu, for a whole minibatch in one launch:
//   gather                mannequin/_trajectory.py:79-86 (traj.o[idx], traj.a[idx], traj.r[idx]), benchmarks/ppo.py:30-33
//   Function.evaluate     mannequin/basicnet.py:192-207
//   head                  benchmarks/gae.py:22, benchmarks/ppo.py:34-37, mannequin/distrib.py:24-32,61-68
//   backprop              mannequin/basicnet.py:209-257 + mannequin/backprop.py:11-60
//
// What is new against mq_tc.cu (which stays for calls that pass separate arrays and for the forward-only passes):
//   * PRECISION MODES (template P = bf16 planes per operand, selected per call through mq_head_t.tc): 3 planes = exact
//     fp32 products (six plane pairs), 2 planes = 16 mantissa bits (three pairs), 1 plane = plain bf16 (one pair, MUFU
//     tanh).  The tolerances are stated in include/mq_b200.h and asserted by tests/kernel_cases.py: case_tc_modes.
//   * TWO TILES IN FLIGHT per SM (template NS = 2, for P <= 2 where two 128-row tiles fit in shared memory): the 16
//     worker warps are two groups of 8, each group owns a tile SLOT (its own operand buffers and Z / dA columns in
//     TMEM); the issuer warp alternates between the slots, so one slot's products run under the other slot's
//     epilogue.  Hand-offs are mbarriers only (workers -> issuer: one arrive per warp; tensor pipe -> workers:
//     tcgen05.commit), there is no CTA-wide barrier inside a tile.  Both slots accumulate dW / db into the SAME TMEM
//     accumulators (MMAs execute in issue order, one issuing thread: bit-reproducible).
//   * PACKED RECORDS: a minibatch row is ONE aligned record x | target | w0 | w1 (mq_pack_records; 64 bytes = two
//     sectors for the 11-obs / 3-act policy) fetched with 128-bit loads one tile ahead, instead of four gathers of 44 + 12
//     + 4 + 4 bytes at sector granularity.
//   * OPERAND IMAGE: the bf16 planes of the weights, the biases and the Gaussian head parameters arrive as ONE bulk
//     copy (cp.async.bulk on the TMA engine, completion counted on an mbarrier) of an image kept up to date by the
//     optimiser-step kernel (mq_tc3.cuh), instead of being converted from theta by every CTA of every launch.
//   * the fp32 activations for the backward pass are rebuilt from their planes in shared memory (no TMEM copy).
#include "mq_tc3.cuh"

#ifndef MQ_EMU

#define HALF_LOG_2PI 0.91893853320467274178f
#define T3_NT 512                // worker threads (16 warps)
#define T3_NTI 544               // + the issuer warp
#define T3_ROWS 128
#define T3_CB 2048               // one 8-feature chunk of all 128 rows (128 rows x 16 B)
#define T3_FLUSH_TILES 4         // tiles accumulated in TMEM before the dW accumulators are folded into memory (the tensor
                                 // pipe truncates when it adds into its accumulator: mq_gemm_tc.cu)
#define T3_MAXLD 2               // 128-bit record loads per thread and tile

struct T3Plan {
    mq_net_t net;
    T3Img img;
    int L, tile_rows, with_grad;
    int rec_w, q4, hin_w;
    // byte offsets inside a slot (slot s starts at s * slot_bytes); act[l]: planes P-1 .. 0, then the ones chunk
    int off_dzl, dzl_plane, off_act[T3_MAXL], act_plane[T3_MAXL], off_hin, slot_bytes;
    int x_bufs, x_stride, hin_stride;   // input tile (operand planes + head inputs) double-buffered when shared memory allows
    int off_const, smem_bytes;
    // TMEM columns
    int col_za[2], col_acc[T3_MAXL];
    int col_save[2];             // fp32 hidden activations kept in TMEM for the backward pass (-1: rebuilt from the planes)
};

// One net's share of a launch.  A launch carries up to TWO jobs (CTAs [0, grid) run job 0, the rest job 1): the value and
// the policy minibatch gradient of one PPO step (benchmarks/gae.py:71-73 and benchmarks/ppo.py:29-40 are independent of
// each other) as ONE grid, each on half of the SMs -- at 8,192 rows per GPU a net has one tile per SM and a launch is one
// tile's chain of dependent phases long, so two nets side by side cost the time of one.
struct T3Job {
    T3Plan p;
    const unsigned char *image;
    const float *rec;
    const int32_t *idx;
    int64_t n_rows;
    mq_head_t head;
    float *partial, *out, *logp_out;
    int grid;
};
struct T3Jobs { T3Job job[2]; };

// ---- small PTX helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ void t3_ld16_nowait(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void t3_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void t3_st16_nowait(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void t3_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void t3_mbar_init(uint64_t *b, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void t3_mbar_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
// generic-proxy shared-memory writes and TMEM reads of this warp are done: tell the issuer (one arrive per warp)
__device__ __forceinline__ void t3_warp_ready(uint32_t mbar, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) t3_mbar_arrive(mbar);
}
__device__ __forceinline__ float4 t3_ldg128(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float t3_tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- plane arithmetic ----------------------------------------------------------------------------------------
// two values -> P words of packed bf16 pairs (low half = a), w[0] + w[1] + .. = the values (exact for P = 3)
template <int P, int F>
__device__ __forceinline__ void t3_split_pair(float a, float b, uint32_t *w) {
    if (F == 1) {                                       // fp16 planes, the low one scaled by 2^11 (mq_tc3.cuh)
        const __half2 h0 = __floats2half2_rn(a, b);
        const float2 f0 = __half22float2(h0);
        const __half2 h1 = __floats2half2_rn((a - f0.x) * T3_LOW_SCALE, (b - f0.y) * T3_LOW_SCALE);
        w[0] = *reinterpret_cast<const uint32_t *>(&h0);
        w[1] = *reinterpret_cast<const uint32_t *>(&h1);
        return;
    }
#pragma unroll
    for (int pl = 0; pl < P; ++pl) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        w[pl] = *reinterpret_cast<const uint32_t *>(&h);
        if (pl + 1 < P) {
            a -= __uint_as_float(w[pl] << 16);
            b -= __uint_as_float(w[pl] & 0xFFFF0000u);
        }
    }
}

// address of (chunk, row) in plane 0 of a buffer whose planes are stored P-1 .. 0
template <int P, int CB>
__device__ __forceinline__ unsigned char *t3_chunk_ptr(unsigned char *buf, int plane_bytes, int chunk, int row) {
    return buf + (P - 1) * plane_bytes + chunk * CB + (row >> 3) * 128 + (row & 7) * 16;
}

// 8 consecutive features of one row -> one 16-byte chunk in each plane
template <int P, int F, int CB>
__device__ __forceinline__ void t3_store_chunk(unsigned char *buf, int plane_bytes, int chunk, int row, const float *v) {
    uint32_t w[4][P];
#pragma unroll
    for (int j = 0; j < 4; ++j) t3_split_pair<P, F>(v[2 * j], v[2 * j + 1], w[j]);
    unsigned char *q = t3_chunk_ptr<P, CB>(buf, plane_bytes, chunk, row);
#pragma unroll
    for (int pl = 0; pl < P; ++pl)
        *reinterpret_cast<uint4 *>(q - pl * plane_bytes) = make_uint4(w[0][pl], w[1][pl], w[2][pl], w[3][pl]);
}

// the value the planes represent (exact fp32 for P = 3, the rounded operand otherwise)
template <int P, int F, int CB>
__device__ __forceinline__ void t3_load_chunk(unsigned char *buf, int plane_bytes, int chunk, int row, float *v) {
    const unsigned char *q = t3_chunk_ptr<P, CB>(buf, plane_bytes, chunk, row);
    uint4 w[P];
#pragma unroll
    for (int pl = 0; pl < P; ++pl) w[pl] = *reinterpret_cast<const uint4 *>(q - pl * plane_bytes);
    if (F == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t x0 = j == 0 ? w[0].x : (j == 1 ? w[0].y : (j == 2 ? w[0].z : w[0].w));
            const uint32_t x1 = j == 0 ? w[P - 1].x : (j == 1 ? w[P - 1].y : (j == 2 ? w[P - 1].z : w[P - 1].w));
            const float2 f0 = __half22float2(*reinterpret_cast<const __half2 *>(&x0));
            const float2 f1 = __half22float2(*reinterpret_cast<const __half2 *>(&x1));
            v[2 * j] = fmaf(f1.x, T3_LOW_UNSCALE, f0.x);
            v[2 * j + 1] = fmaf(f1.y, T3_LOW_UNSCALE, f0.y);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float lo = 0.0f, hi = 0.0f;
#pragma unroll
        for (int pl = 0; pl < P; ++pl) {
            const uint32_t x = j == 0 ? w[pl].x : (j == 1 ? w[pl].y : (j == 2 ? w[pl].z : w[pl].w));
            lo = pl == 0 ? __uint_as_float(x << 16) : lo + __uint_as_float(x << 16);
            hi = pl == 0 ? __uint_as_float(x & 0xFFFF0000u) : hi + __uint_as_float(x & 0xFFFF0000u);
        }
        v[2 * j] = lo;
        v[2 * j + 1] = hi;
    }
}

template <int P, int F, int ACT>
__device__ __forceinline__ float t3_act(float z, float leak) {
    if (ACT == MQ_ACT_TANH) {
        if (P == 1) return t3_tanh_fast(z);             // MUFU.TANH: 2^-11, the size of a bf16 rounding
        if (P == 2) {                                   // 1 - 2 / (2^(2 log2(e) |z|) + 1): ~1.2e-7 ABSOLUTE (no small-|z| branch):
            float t, r;                                 // inside the 16-bit operands of bf16x2; for fp16x2 it is the fp32 bar's
                                                        // absolute term (tests/kernel_cases.py: atol 3e-6 max|ref|), half the work
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(z) * 2.8853900817779268f));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
            return copysignf(fmaf(r, -2.0f, 1.0f), z);
        }
        return tc_tanh(z);
    }
    return z < 0.0f ? z * leak : z;
}

// the value of an accumulator spread over NB column blocks: block 0 (+ block 1, scaled back for the fp16 format)
template <int NB, int F>
__device__ __forceinline__ float t3_blocks(float b0, float b1) {
    if (NB == 1) return b0;
    return F == 1 ? fmaf(b1, T3_LOW_UNSCALE, b0) : b0 + b1;
}

// ---- MMA issue helpers (elected lane of the issuer warp, warp-uniform control flow) ---------------------------------
// activation plane at `addr` viewed K-major (M = rows, K = features)
__device__ __forceinline__ TcView t3_rows_k(uint32_t addr, uint32_t cb) {
    return TcView{tc_desc_lo(addr, cb), tc_desc_hi(128u), (2u * cb) >> 4};
}
// activation buffer from `addr` on viewed MN-major (M or N = features, planes side by side; K = rows)
__device__ __forceinline__ TcView t3_feat_mn(uint32_t addr, uint32_t cb) {
    return TcView{tc_desc_lo(addr, 128u), tc_desc_hi(cb), 256u >> 4};
}

// D[128 x (NB blocks of nb)] = sum over plane pairs i + j < P of A_i B_j.  A: row-major planes (plane 0 at a0_addr, plane i
// i * a_plane bytes BELOW it); B: an image whose planes 0, 1, .. sit side by side along N (b_plane16 = stride >> 4).
// F = 1 (fp16, scaled low planes): column block 0 = A_0 B_0, block 1 = A_0 B_1' + A_1' B_0 (to be scaled by 2^-11).
template <int P, int F>
__device__ __forceinline__ void t3_rows_product(uint32_t d_tmem, uint32_t a0_addr, uint32_t a_plane, uint32_t cb, const TcView &b,
                                                int b_mn, uint32_t b_plane16, int nb, int ksteps) {
    const TcView a0 = t3_rows_k(a0_addr, cb);
    const uint32_t ap16 = a_plane >> 4;
    const uint32_t id2 = tc_idesc_fmt(T3_ROWS, 2 * nb, 0, b_mn, F);
    const uint32_t id1 = tc_idesc_fmt(T3_ROWS, nb, 0, b_mn, F);
    uint32_t alo = a0.lo, blo = b.lo;
#pragma unroll 1
    for (int k = 0; k < ksteps; ++k) {
        const uint32_t acc = k > 0 ? 1u : 0u;
        if (P == 1) {
            tc_mma(d_tmem, alo, a0.hi, blo, b.hi, id1, acc);                              // A_0 x B_0
        } else if (P == 2) {
            tc_mma(d_tmem, alo, a0.hi, blo, b.hi, id2, acc);                              // A_0 x [B_0|B_1]
            tc_mma(d_tmem + (F == 1 ? nb : 0), alo - ap16, a0.hi, blo, b.hi, id1, 1u);    // A_1 x B_0
        } else {
            tc_mma(d_tmem, alo, a0.hi, blo, b.hi, id2, acc);                              // A_0 x [B_0|B_1]
            tc_mma(d_tmem, alo - ap16, a0.hi, blo, b.hi, id2, 1u);                        // A_1 x [B_0|B_1]
            tc_mma(d_tmem, alo, a0.hi, blo + 2 * b_plane16, b.hi, id1, 1u);               // A_0 x B_2
            tc_mma(d_tmem, alo - 2 * ap16, a0.hi, blo, b.hi, id1, 1u);                    // A_2 x B_0
        }
        alo += a0.k16;
        blo += b.k16;
    }
}

// acc[64 x (NB blocks of nb, + extra)] (+)= sum over plane pairs of Da_i^T Db_j over the rows of the tile.  Da, Db:
// activation buffers (planes P-1 .. 0 in memory, buffer start at *_base); Da's features are M, Db's planes side by side
// (+ `extra` columns behind plane 0: the ones chunk) are N.  Column blocks for P >= 2: [Db_1 | Db_0 | extra].
// F = 1: block [0, nb) = Da_0 Db_1' + Da_1' Db_0 (scaled), [nb, 2 nb) = Da_0 Db_0, [2 nb, +8) = Da_0 e, [2 nb + 8, +8) = Da_1' e.
template <int P, int F>
__device__ __forceinline__ void t3_feat_product(uint32_t acc_tmem, uint32_t da_base, uint32_t da_plane, uint32_t db_base,
                                                uint32_t db_plane, uint32_t cb, int nb, int extra, int ksteps, bool accumulate) {
    const TcView a = t3_feat_mn(da_base, cb), b = t3_feat_mn(db_base, cb);
    const uint32_t ap16 = da_plane >> 4, bp16 = db_plane >> 4;
    const uint32_t id_all = tc_idesc_fmt(64, 2 * nb + extra, 1, 1, F), id_one = tc_idesc_fmt(64, nb + extra, 1, 1, F),
                   id_blk = tc_idesc_fmt(64, nb, 1, 1, F), id_ex = tc_idesc_fmt(64, 8, 1, 1, F);
    uint32_t alo = a.lo, blo = b.lo;
#pragma unroll 1
    for (int k = 0; k < ksteps; ++k) {
        const uint32_t acc = (accumulate || k > 0) ? 1u : 0u;
        if (P == 1) {
            tc_mma(acc_tmem, alo, a.hi, blo, b.hi, id_one, acc);                                        // Da_0 x [Db_0|e]
        } else if (P == 2 && F == 1) {
            tc_mma(acc_tmem, alo + ap16, a.hi, blo, b.hi, id_all, acc);                                 // Da_0 x [Db_1'|Db_0|e]
            tc_mma(acc_tmem, alo, a.hi, blo + bp16, b.hi, id_blk, 1u);                                  // Da_1' x Db_0
            if (extra) tc_mma(acc_tmem + 2 * nb + 8, alo, a.hi, blo + 2 * bp16, b.hi, id_ex, acc);      // Da_1' x e
        } else if (P == 2) {
            tc_mma(acc_tmem, alo + ap16, a.hi, blo, b.hi, id_all, acc);                                 // Da_0 x [Db_1|Db_0|e]
            tc_mma(acc_tmem + nb, alo, a.hi, blo + bp16, b.hi, id_one, 1u);                             // Da_1 x [Db_0|e]
        } else {
            tc_mma(acc_tmem, alo + 2 * ap16, a.hi, blo + bp16, b.hi, id_all, acc);                      // Da_0 x [Db_1|Db_0|e]
            tc_mma(acc_tmem, alo + ap16, a.hi, blo + bp16, b.hi, id_all, 1u);                           // Da_1 x [Db_1|Db_0|e]
            tc_mma(acc_tmem + nb, alo, a.hi, blo + 2 * bp16, b.hi, id_one, 1u);                         // Da_2 x [Db_0|e]
            tc_mma(acc_tmem, alo + 2 * ap16, a.hi, blo, b.hi, id_blk, 1u);                              // Da_0 x Db_2
        }
        alo += a.k16;
        blo += b.k16;
    }
}

// forward weight image [k chunk][plane][n][8 k] of P planes from `addr` on, K-major (N = plane * np + n, K = in)
template <int P>
__device__ __forceinline__ TcView t3_wf_k(uint32_t addr, int np) {
    return TcView{tc_desc_lo(addr, (uint32_t)P * (uint32_t)np * 16u), tc_desc_hi(128u), (2u * (uint32_t)P * (uint32_t)np * 16u) >> 4};
}

// ---- the kernel ------------------------------------------------------------------------------------------------
template <int P, int F, int NS, int K0, int HACT>
__global__ void __launch_bounds__(T3_NTI, 1)
tc3_grad_kernel(const __grid_constant__ T3Jobs launch, long long *prof) {
    const bool second = blockIdx.x >= (unsigned)launch.job[0].grid;
    const T3Job &J = second ? launch.job[1] : launch.job[0];
    const T3Plan &p = J.p;
    const mq_head_t &head = J.head;
    const unsigned char *__restrict__ image = J.image;
    const float *__restrict__ rec = J.rec;
    const int32_t *__restrict__ idx = J.idx;
    const int64_t n_rows = J.n_rows;
    float *__restrict__ partial = J.partial, *__restrict__ out = J.out, *__restrict__ logp_out = J.logp_out;
    const int bid = (int)blockIdx.x - (second ? launch.job[0].grid : 0), gdim = J.grid;       // this CTA within its job
    extern __shared__ __align__(1024) unsigned char smem[];
    // developer aid (build with MQ_TC_PROFILE=1, tools/tc3_trace.py): time stamps of CTA 0 -- worker thread 0 logs into
    // prof[0..255], the issuing lane into prof[256..511]
#ifdef MQ_TC_PROFILE
    int prof_n = 0;
    const bool prof_on = prof != nullptr && blockIdx.x == 0;   // (first CTA of the launch)
#define T3_MARK(tag) do { if (prof_on && (threadIdx.x == 0 || threadIdx.x == T3_NT) && prof_n < 127) { \
        long long *q_ = prof + (threadIdx.x == 0 ? 0 : 256) + 2 * prof_n; q_[0] = (tag); q_[1] = clock64(); ++prof_n; } } while (0)
#else
    (void)prof;
#define T3_MARK(tag) do { } while (0)
#endif
    T3_MARK(1);
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_full[2], s_dw[2], s_ready[2], s_readyx[2], s_img;
    __shared__ float s_rb[T3_NT / 32][8];               // per-warp column sums of the output layer's dZ (bias / logstd gradients)
    constexpr int ROWS = T3_ROWS;             // rows of a slot
    constexpr int CB = T3_CB;                 // one 8-feature chunk of all rows of a slot
    constexpr int GT = T3_NT / NS;            // threads of a slot group
    constexpr int GW = GT / 32;               // warps of a slot group
    constexpr int TPR = GW / 4;               // threads per tile row
    constexpr int CPT = T3_HID / TPR;         // hidden columns per thread
    constexpr int DPT = T3_NO / TPR;          // dZ columns of the output layer per thread
    constexpr int NB = P > 1 ? 2 : 1;         // column blocks the epilogue adds
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const bool worker = warp < T3_NT / 32;
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may touch
    const int slot = worker ? warp / GW : 0;
    const int cgi = (warp - slot * GW) >> 2;            // which CPT of the 64 columns this thread owns
    const int row = quarter * 32 + lane;                // TMEM lane == row of the tile
    const int gtid = tid - slot * GT;                   // index within the slot group
    const int L = p.L;
    const int n_in = p.net.n_in, n_out = p.net.n_out;

    // ---- one-time setup (on-chip work that overlaps the tail of the previous kernel: PDL) ----------------------
    mq_pdl_trigger();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            t3_mbar_init(&s_full[s], 1); t3_mbar_init(&s_dw[s], 1); t3_mbar_init(&s_ready[s], GW); t3_mbar_init(&s_readyx[s], GW);
        }
        t3_mbar_init(&s_img, 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (tid < (T3_NT / 32) * 8) (&s_rb[0][0])[tid] = 0.0f;
    // clear the slots except the hidden activation buffers (every tile rewrites all of their rows)
    if (worker) {
        for (int s = 0; s < NS; ++s) {
            uint4 *z = reinterpret_cast<uint4 *>(smem + s * p.slot_bytes);
            for (int i = tid; i < p.off_act[1] / 16; i += T3_NT) z[i] = make_uint4(0, 0, 0, 0);
        }
    }
    __syncthreads();
    if (worker) {
        const uint32_t ob = F == 1 ? 0x3C00u : 0x3F80u;                // 1.0 as fp16 / bf16
        for (int s = 0; s < NS; ++s)
            for (int l = 0; l < L; ++l)
                for (int xb = 0; xb < (l == 0 ? p.x_bufs : 1); ++xb) {
                    uint32_t *ones = reinterpret_cast<uint32_t *>(smem + s * p.slot_bytes + p.off_act[l] + xb * p.x_stride + P * p.act_plane[l]);
                    for (int i = tid; i < CB / 4; i += T3_NT) ones[i] = ob | (ob << 16);
                }
    }
    const uint32_t sbase = tc_smem_u32(smem);
    const uint32_t img_bar = tc_smem_u32(&s_img);
    T3_MARK(2);
    // The operand image is written by the predecessor (the optimiser-step kernel): it is fetched after griddepcontrol.wait.
    // The minibatch indices and the packed records are NOT (they date from before the update's first step), so the workers
    // start their first gathers before waiting -- under the tail of the step kernel.
    auto wait_and_fetch_image = [&]() {
        mq_pdl_wait();
        T3_MARK(3);
        if (tid == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(img_bar), "r"((uint32_t)p.img.bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sbase + p.off_const), "l"(image), "r"((uint32_t)p.img.bytes), "r"(img_bar) : "memory");
        }
        tc_wait(img_bar, 0);                            // weights, biases, head parameters have landed
        T3_MARK(4);
    };
    const float *sBias = reinterpret_cast<const float *>(smem + p.off_const + p.img.off_bias);      // [L][64]
    const float *sMisc = reinterpret_cast<const float *>(smem + p.off_const + p.img.off_misc);      // logstd[16], exp(-logstd)[16]
    unsigned char *sl = smem + slot * p.slot_bytes;                                                 // this group's slot
    float *sHin0 = reinterpret_cast<float *>(sl + p.off_hin);                                       // [x_bufs][128][hin_w]
    const int tr = p.tile_rows;
    const int64_t ntiles = (n_rows + tr - 1) / tr;
    const int dw_ksteps = tr / 16;
    const bool gauss = mq_head_gauss(head.kind), prob = mq_head_prob(head.kind);
    const int lastact = p.net.act[L - 1];
    const float lastleak = p.net.leak[L - 1];
    const int jobs = p.with_grad ? 2 * L : L;
    // tiles of this CTA: ordinal i is tile blockIdx.x + i * gdim; slot s takes the ordinals i % NS == s
    const int n_i = bid < ntiles ? (int)((ntiles - bid + gdim - 1) / gdim) : 0;

    tc_publish();                                       // zeroed buffers / ones chunks visible to the tensor pipe, TMEM address
    const uint32_t tm = __shfl_sync(0xffffffffu, s_tmem, 0);

    if (!worker) {
        // ---- issuer warp ---------------------------------------------------------------------------------------
        wait_and_fetch_image();
        uint32_t rdy_par[2] = {0u, 0u}, rdx_par[2] = {0u, 0u};
        int tcount[2] = {0, 0};                          // tiles issued per slot: selects the input buffer
        for (int i0 = 0; i0 < n_i; i0 += T3_FLUSH_TILES) {
            const int i1 = i0 + T3_FLUSH_TILES < n_i ? i0 + T3_FLUSH_TILES : n_i;
            int cnt[2];
            cnt[0] = (i1 - i0 + NS - 1) / NS;
            cnt[1] = NS == 2 ? (i1 - i0) / 2 : 0;
            bool acc_started[T3_MAXL] = {false, false, false};
            for (int k = 0; k < cnt[0] * jobs; ++k) {
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    if (k >= cnt[s] * jobs) continue;
                    const int job = k % jobs;
                    const uint32_t sb = sbase + s * p.slot_bytes;
                    const uint32_t za = tm + p.col_za[s];
                    const uint32_t full = tc_smem_u32(&s_full[s]), dwb = tc_smem_u32(&s_dw[s]);
                    const uint32_t xo = (p.x_bufs == 2 && (tcount[s] & 1)) ? (uint32_t)p.x_stride : 0u;     // this tile's input buffer
                    // job 0 waits for the staged input tile (its own barrier: with two input buffers the next tile is staged
                    // right behind the last backward epilogue, and two arrivals of one warp must not fall into one phase)
                    if (job == 0) { tc_wait(tc_smem_u32(&s_readyx[s]), rdx_par[s]); rdx_par[s] ^= 1u; }
                    else { tc_wait(tc_smem_u32(&s_ready[s]), rdy_par[s]); rdy_par[s] ^= 1u; }
                    T3_MARK(100 + 10 * s + job);
                    if (tc_elect_one()) {
                        if (job < L) {
                            const int l = job;
                            t3_rows_product<P, F>(za, sb + p.off_act[l] + (l == 0 ? xo : 0u) + (P - 1) * p.act_plane[l], p.act_plane[l], CB,
                                               t3_wf_k<P>(sbase + p.off_const + p.img.off_wf[l], p.img.np[l]), 0, (uint32_t)p.img.np[l],
                                               p.img.np[l], p.img.kp[l] / 16);
                            tc_commit(full);
                        } else {
                            const int l = 2 * L - 1 - job;
                            if (l == L - 1) {
                                // dA_l = dZ_l W_l^T (K = 16 outputs), dW_l (+)= A_l^T dZ_l (direct form); db_l / dlogstd = 1^T dZ_l are
                                // summed by the workers (warp shuffles in the head phase), not by the tensor pipe
                                t3_rows_product<P, F>(za, sb + p.off_dzl + (P - 1) * p.dzl_plane, p.dzl_plane, CB,
                                                   tc_wb_mn(sbase + p.off_const + p.img.off_wb[l], p.img.np[l]), 1,
                                                   (uint32_t)p.img.wb_plane[l] >> 4, p.img.kp[l], 1);
                                tc_commit(full);             // dA alone: the epilogue's arithmetic runs under the dW products
                                t3_feat_product<P, F>(tm + p.col_acc[l], sb + p.off_act[l], p.act_plane[l], sb + p.off_dzl,
                                                   p.dzl_plane, CB, T3_NO, 0, dw_ksteps, acc_started[l]);
                                acc_started[l] = true;
                                tc_commit(dwb);
                            } else {
                                // dZ_l sits in act[l+1].  dA_l = dZ_l W_l^T (l > 0); dW_l^T, db_l (+)= dZ_l^T [A_l | 1]
                                if (l > 0) {
                                    t3_rows_product<P, F>(za, sb + p.off_act[l + 1] + (P - 1) * p.act_plane[l + 1], p.act_plane[l + 1], CB,
                                                       tc_wb_mn(sbase + p.off_const + p.img.off_wb[l], p.img.np[l]), 1,
                                                       (uint32_t)p.img.wb_plane[l] >> 4, p.img.kp[l], p.img.np[l] / 16);
                                    tc_commit(full);
                                }
                                t3_feat_product<P, F>(tm + p.col_acc[l], sb + p.off_act[l + 1], p.act_plane[l + 1],
                                                   sb + p.off_act[l] + (l == 0 ? xo : 0u), p.act_plane[l], CB, p.img.kp[l], 8, dw_ksteps,
                                                   acc_started[l]);
                                acc_started[l] = true;
                                // dW_0 (no dA) also completes on the dW barrier, NOT on `full`: with two input buffers the next
                                // tile's F0 may be committed before a slow warp has looked at dW_0's completion, and two
                                // completions of one barrier inside one wait would leave that warp waiting for a third
                                tc_commit(dwb);
                            }
                        }
                    }
                    T3_MARK(200 + 10 * s + job);
                    if (job == jobs - 1) ++tcount[s];
                    __syncwarp();
                }
            }
            if (p.with_grad) {
                // flush: the workers fold the accumulators into memory between these two barriers
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
                __syncthreads();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
        }
    } else {
        // ---- worker warps -----------------------------------------------------------------------------------
        const uint32_t full = tc_smem_u32(&s_full[slot]), dwb = tc_smem_u32(&s_dw[slot]), rdy = tc_smem_u32(&s_ready[slot]),
                       rdx = tc_smem_u32(&s_readyx[slot]);
        uint32_t par_full = 0u, par_dw = 0u;
        const uint32_t tlane = tm + ((uint32_t)(quarter * 32) << 16);
        const uint32_t za = tlane + p.col_za[slot];
        const bool wlive = row - lane < tr;             // this warp owns rows of the tile
        // gather bookkeeping: this thread's (row, 128-bit word) items of a tile
        int item_rq[T3_MAXLD];
#pragma unroll
        for (int j = 0; j < T3_MAXLD; ++j) {
            const int item = gtid + j * GT;
            const int r = item / p.q4, q = item - r * p.q4;
            item_rq[j] = r < ROWS ? ((r << 8) | q) : -1;
        }
        int cur_idx[T3_MAXLD];
        float4 recs[T3_MAXLD];
        auto fetch_idx = [&](int ord) {                 // source row (or -1) of this thread's items of tile ordinal `ord`
            const int64_t tile = (int64_t)bid + (int64_t)ord * gdim;
#pragma unroll
            for (int j = 0; j < T3_MAXLD; ++j) {
                int v = -1;
                if (item_rq[j] >= 0 && ord < n_i) {
                    const int r = item_rq[j] >> 8;
                    const int64_t gr = tile * tr + r;
                    if (r < tr && gr < n_rows) v = idx ? tc_ldg_s32(idx + gr) : (int)gr;
                }
                cur_idx[j] = v;
            }
        };
        auto load_recs = [&]() {
#pragma unroll
            for (int j = 0; j < T3_MAXLD; ++j) {
                float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (cur_idx[j] >= 0) v = t3_ldg128(rec + (int64_t)cur_idx[j] * p.rec_w + 4 * (item_rq[j] & 255));
                recs[j] = v;
            }
        };
        auto stage = [&](int buf) {                     // records -> X operand planes and head inputs of input buffer `buf`
            unsigned char *xbuf = sl + p.off_act[0] + buf * p.x_stride;
            float *sHin = sHin0 + buf * p.hin_stride;
#pragma unroll
            for (int j = 0; j < T3_MAXLD; ++j) {
                if (item_rq[j] < 0) continue;
                const int r = item_rq[j] >> 8, e0 = 4 * (item_rq[j] & 255);
                const float v[4] = {recs[j].x, recs[j].y, recs[j].z, recs[j].w};
                if (e0 + 3 < n_in) {
                    uint32_t w0[P], w1[P];
                    t3_split_pair<P, F>(v[0], v[1], w0);
                    t3_split_pair<P, F>(v[2], v[3], w1);
                    unsigned char *d = t3_chunk_ptr<P, CB>(xbuf, p.act_plane[0], e0 >> 3, r) + (e0 & 7) * 2;
#pragma unroll
                    for (int pl = 0; pl < P; ++pl) *reinterpret_cast<uint2 *>(d - pl * p.act_plane[0]) = make_uint2(w0[pl], w1[pl]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int f = e0 + e;
                        if (f < n_in) {
                            uint32_t w[P];
                            t3_split_pair<P, F>(v[e], 0.0f, w);
                            unsigned char *d = t3_chunk_ptr<P, CB>(xbuf, p.act_plane[0], f >> 3, r) + (f & 7) * 2;
#pragma unroll
                            for (int pl = 0; pl < P; ++pl) *reinterpret_cast<uint16_t *>(d - pl * p.act_plane[0]) = (uint16_t)(w[pl] & 0xFFFFu);
                        } else if (f - n_in < p.hin_w) {
                            sHin[r * p.hin_w + f - n_in] = v[e];
                        }
                    }
                }
            }
        };

        bool pending = false;                           // a committed dW_0 product nobody waited for yet
        bool flushed = false;
        auto flush_grad = [&]() {
            // TMEM accumulators -> this CTA's partial gradient (set on the first flush, added afterwards).  M = 64
            // accumulators: row i is lane i % 16 of the warps of quarter i / 16; the four warps of a quarter split the columns
            float *dst = partial + (int64_t)bid * p.net.n_params;
            const bool add = flushed;
            const int cg4 = warp >> 2;
            const int i = quarter * 16 + lane;
            const bool lv = lane < 16;
            for (int l = 0; l < L; ++l) {
                T3_MARK(92 + l);
                const int in = p.net.in[l], on = p.net.out[l];
                const int nb = (l == L - 1) ? T3_NO : p.img.kp[l];
                for (int c0 = cg4 * 16; c0 < nb; c0 += 64) {
                    float v0[16], v1[16];
                    t3_ld16_nowait(tlane + p.col_acc[l] + c0, v0);
                    if (NB == 2) t3_ld16_nowait(tlane + p.col_acc[l] + nb + c0, v1);
                    t3_ld_wait();
                    if (!lv) continue;
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        // l < L-1: acc[o][k] = dW_l[k][o]; last layer: acc[k][o]
                        const bool ok = (l < L - 1) ? (i < on && c0 + j < in) : (i < in && c0 + j < on);
                        if (!ok) continue;
                        float *q = dst + p.net.w_off[l] + ((l < L - 1) ? (c0 + j) * on + i : i * on + c0 + j);
                        const float val = NB == 2 ? t3_blocks<NB, F>(v1[j], v0[j]) : v0[j];    // (fp16 format: block 0 is the scaled one)
                        *q = add ? *q + val : val;
                    }
                }
                if (l < L - 1 && cg4 == 1) {    // the ones column(s) behind the blocks
                    float v[16];
                    t3_ld16_nowait(tlane + p.col_acc[l] + NB * nb - (F == 1 ? 0 : 8), v);
                    t3_ld_wait();
                    const float bsum = F == 1 ? fmaf(v[8], T3_LOW_UNSCALE, v[0]) : v[8];
                    if (lv && i < on) { float *q = dst + p.net.b_off[l] + i; *q = add ? *q + bsum : bsum; }
                }
            }
            if (warp == 0 && lane < T3_NO) {    // last layer: column sums of dZ, gathered per warp in the head phases
                // column j belongs to the warps whose column group is j / DPT (every slot, every lane quarter): fixed order
                const int j = lane, cgj = j / DPT, kj = j - cgj * DPT;
                float sv = 0.0f;
                for (int w = 0; w < T3_NT / 32; ++w)
                    if (((w % GW) >> 2) == cgj) sv += s_rb[w][kj];
                // (s_rb holds running totals over all tiles of the CTA: written, not added)
                if (j < n_out) dst[p.net.b_off[L - 1] + j] = sv;
                if (p.net.logstd_off >= 0 && j >= 8 && j < 8 + n_out) dst[p.net.logstd_off + j - 8] = gauss ? sv : 0.0f;
            }
            flushed = true;
        };

        fetch_idx(slot);
        load_recs();
        fetch_idx(slot + NS);
        wait_and_fetch_image();
        int tcount = 0;                                 // tiles this slot has started: selects the input buffer
        bool staged = false;                            // this tile's input was staged (and announced) behind the previous tile
        for (int i0 = 0; i0 < n_i; i0 += T3_FLUSH_TILES) {
            const int i1 = i0 + T3_FLUSH_TILES < n_i ? i0 + T3_FLUSH_TILES : n_i;
            for (int ord = i0 + slot; ord < i1; ord += NS) {
                const int64_t tile = (int64_t)bid + (int64_t)ord * gdim;
                const int64_t gr = tile * tr + row;
                const bool rvalid = row < tr && gr < n_rows;
                // ---- stage the input tile (unless it already was: two input buffers) ---------------------------------
                const int buf = p.x_bufs == 2 ? (tcount & 1) : 0;
                const float *sHin = sHin0 + buf * p.hin_stride;
                T3_MARK(10);
                if (!staged) {
                    if (pending) { tc_wait(dwb, par_dw); par_dw ^= 1u; pending = false; }          // dW_0 read the buffer
                    stage(buf);
                    t3_warp_ready(rdx, lane);
                } else if (pending) {                   // (consume dW_0's phase of the dW barrier: its next phase is a whole
                    tc_wait(dwb, par_dw); par_dw ^= 1u; pending = false;       // forward pass away)
                }
                staged = false;
                ++tcount;
                T3_MARK(11);
                // ---- forward -----------------------------------------------------------------------------------
                for (int l = 0; l < L - 1; ++l) {
                    if (l == L - 2) {
                        // records of this slot's NEXT tile and the source rows of the one after: issued here so that they
                        // have the remaining phases of this tile to arrive
                        load_recs();
                        fetch_idx(ord + 2 * NS);
                    }
                    tc_wait(full, par_full); par_full ^= 1u;
                    T3_MARK(20 + l);
                    if (wlive) {
                        // A_{l+1} = act(Z_l + b_l) as bf16 planes
                        const float leak = p.net.leak[l];
                        unsigned char *abuf = sl + p.off_act[l + 1];
#pragma unroll
                        for (int h = 0; h < CPT / 16; ++h) {
                            const int c0 = cgi * CPT + 16 * h;
                            float v[16], v2[16];
                            t3_ld16_nowait(za + c0, v);
                            if (NB == 2) t3_ld16_nowait(za + T3_HID + c0, v2);
                            t3_ld_wait();
                            T3_MARK(300 + l);
                            const float4 *b4 = reinterpret_cast<const float4 *>(sBias + l * 64 + c0);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float4 bb = b4[j];
                                v[4 * j + 0] = t3_act<P, F, HACT>(t3_blocks<NB, F>(v[4 * j + 0], v2[4 * j + 0]) + bb.x, leak);
                                v[4 * j + 1] = t3_act<P, F, HACT>(t3_blocks<NB, F>(v[4 * j + 1], v2[4 * j + 1]) + bb.y, leak);
                                v[4 * j + 2] = t3_act<P, F, HACT>(t3_blocks<NB, F>(v[4 * j + 2], v2[4 * j + 2]) + bb.z, leak);
                                v[4 * j + 3] = t3_act<P, F, HACT>(t3_blocks<NB, F>(v[4 * j + 3], v2[4 * j + 3]) + bb.w, leak);
                            }
                            if (p.with_grad && p.col_save[slot] >= 0) t3_st16_nowait(tlane + p.col_save[slot] + l * T3_HID + c0, v);
                            T3_MARK(310 + l);
                            t3_store_chunk<P, F, CB>(abuf, p.act_plane[l + 1], c0 / 8, row, v);
                            t3_store_chunk<P, F, CB>(abuf, p.act_plane[l + 1], c0 / 8 + 1, row, v + 8);
                            T3_MARK(320 + l);
                        }
                        if (p.with_grad && p.col_save[slot] >= 0) t3_st_wait();
                    }
                    t3_warp_ready(rdy, lane);
                    T3_MARK(30 + l);
                }
                // ---- head: loss-gradient rule on the output row.  All TPR threads of a row evaluate it (inputs come
                // from shared memory) and each splits / stores its own DPT of the 16 dZ columns.
                tc_wait(full, par_full); par_full ^= 1u;
                T3_MARK(40);
                if (wlive) {
                    float y[T3_NO], dz[T3_NO];
                    {
                        float y2[T3_NO];
                        t3_ld16_nowait(za, y);
                        if (NB == 2) t3_ld16_nowait(za + T3_NO, y2);
                        t3_ld_wait();
                        T3_MARK(330);
                        const float *b = sBias + (L - 1) * 64;
                        if (lastact == MQ_ACT_NONE) {          // (the usual case, decided once: no per-element activation code)
#pragma unroll
                            for (int n = 0; n < T3_NO; ++n) {
                                y[n] = (n < n_out) ? t3_blocks<NB, F>(y[n], y2[n]) + b[n] : 0.0f;
                                dz[n] = 0.0f;
                            }
                        } else {
#pragma unroll
                            for (int n = 0; n < T3_NO; ++n) {
                                y[n] = (n < n_out) ? mq_act_apply(t3_blocks<NB, F>(y[n], y2[n]) + b[n], lastact, lastleak) : 0.0f;
                                dz[n] = 0.0f;
                            }
                        }
                    }
                    float tgt[T3_NO];
                    {
                        const float4 *h4 = reinterpret_cast<const float4 *>(sHin + row * p.hin_w);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float4 t = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                            if (4 * q < p.hin_w) t = h4[q];
                            tgt[4 * q] = t.x; tgt[4 * q + 1] = t.y; tgt[4 * q + 2] = t.z; tgt[4 * q + 3] = t.w;
                        }
                    }
                    const float h_adv = sHin[row * p.hin_w + n_out], h_lpold = sHin[row * p.hin_w + n_out + 1];
                    if (rvalid && out && cgi == 0) {
#pragma unroll
                        for (int n = 0; n < T3_NO; ++n) if (n < n_out) out[gr * n_out + n] = y[n];
                    }
                    if (rvalid) {
                        if (head.kind == MQ_HEAD_DY) {
#pragma unroll
                            for (int n = 0; n < T3_NO; ++n) if (n < n_out) dz[n] = mq_act_grad(y[n], tgt[n], lastact, lastleak);
                        } else if (head.kind == MQ_HEAD_DIFF) {
#pragma unroll
                            for (int n = 0; n < T3_NO; ++n) if (n < n_out) dz[n] = mq_act_grad(y[n], tgt[n] - y[n], lastact, lastleak);
                        } else if (head.kind == MQ_HEAD_SOFTMAX_CE) {
                            float mx = y[0];
#pragma unroll
                            for (int n = 1; n < T3_NO; ++n) if (n < n_out) mx = fmaxf(mx, y[n]);
                            float sum = 0.0f;
#pragma unroll
                            for (int n = 0; n < T3_NO; ++n) if (n < n_out) sum += expf(y[n] - mx);
                            const float inv = 1.0f / sum;
#pragma unroll
                            for (int n = 0; n < T3_NO; ++n)
                                if (n < n_out) dz[n] = mq_act_grad(y[n], tgt[n] - expf(y[n] - mx) * inv - head.p0 * y[n], lastact, lastleak);
                        } else if (mq_head_discrete(head.kind)) {
                            // categorical log-prob (distrib.py:24-32): the target row is the one-hot action
                            float mx = y[0];
#pragma unroll
                            for (int n = 1; n < T3_NO; ++n) if (n < n_out) mx = fmaxf(mx, y[n]);
                            float sum = 0.0f;
#pragma unroll
                            for (int n = 0; n < T3_NO; ++n) if (n < n_out) sum += expf(y[n] - mx);
                            const float lse = mx + logf(sum);
                            float lp = 0.0f;
#pragma unroll
                            for (int n = 0; n < T3_NO; ++n) if (n < n_out) lp += tgt[n] * (y[n] - lse);
                            if (logp_out && cgi == TPR - 1) logp_out[gr] = lp;
                            float w;
                            if (head.kind == MQ_HEAD_DISCRETE_PPO) {
                                const float ratio = expf(lp - h_lpold);
                                const bool kill = (ratio > head.p1 && h_adv > 0.0f) || (ratio < head.p0 && h_adv < 0.0f);
                                w = kill ? 0.0f : ratio * h_adv;
                            } else {
                                w = h_adv;
                            }
#pragma unroll
                            for (int n = 0; n < T3_NO; ++n)
                                if (n < n_out) dz[n] = mq_act_grad(y[n], w * (tgt[n] - expf(y[n] - lse)), lastact, lastleak);
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
                            if (logp_out && cgi == TPR - 1) logp_out[gr] = lp;
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
                                    dz[8 + a] = w * (z * z - 1.0f) + head.ent;
                                    dz[a] = mq_act_grad(y[a], w * z * inv, lastact, lastleak);
                                }
                        }
                    }
                    T3_MARK(331);
                    if (p.with_grad) {
                        float sv[DPT];
#pragma unroll
                        for (int k = 0; k < DPT; ++k) sv[k] = 0.0f;
#pragma unroll
                        for (int q = 0; q < TPR; ++q)
                            if (cgi == q) {
#pragma unroll
                                for (int k = 0; k < DPT; ++k) sv[k] = dz[DPT * q + k];
                            }
                        // db_{L-1} / dlogstd = column sums of dZ over the rows: this warp's 32 rows by shuffles, accumulated per
                        // warp in shared memory in tile order (fixed order: bit-reproducible); invalid rows hold zeros
                        // (transposing butterfly: every exchange halves the columns a lane still carries -- DPT - 1 + the
                        // remaining steps shuffles instead of 5 DPT)
                        {
                            float cs[DPT];
#pragma unroll
                            for (int k = 0; k < DPT; ++k) cs[k] = sv[k];
                            int col = 0;
#pragma unroll
                            for (int o = 16, nn = DPT; o > 0; o >>= 1) {
                                if (nn > 1) {
                                    nn >>= 1;
                                    const bool up = (lane & o) != 0;
#pragma unroll
                                    for (int k = 0; k < DPT / 2; ++k)
                                        if (k < nn) {
                                            const float keep = up ? cs[k + nn] : cs[k], give = up ? cs[k] : cs[k + nn];
                                            cs[k] = keep + __shfl_xor_sync(0xffffffffu, give, o);
                                        }
                                    col += up ? nn : 0;
                                } else {
                                    cs[0] += __shfl_xor_sync(0xffffffffu, cs[0], o);
                                }
                            }
                            if ((lane & (32 / DPT - 1)) == 0) s_rb[warp][col] += cs[0];
                        }
                        T3_MARK(332);
                        unsigned char *dbuf = sl + p.off_dzl;
                        if (DPT == 8) {
                            t3_store_chunk<P, F, CB>(dbuf, p.dzl_plane, cgi, row, sv);
                        } else {
                            uint32_t wa[P], wb[P];
                            t3_split_pair<P, F>(sv[0], sv[1], wa);
                            t3_split_pair<P, F>(sv[2], sv[3], wb);
                            unsigned char *d = t3_chunk_ptr<P, CB>(dbuf, p.dzl_plane, cgi >> 1, row) + (cgi & 1) * 8;
#pragma unroll
                            for (int pl = 0; pl < P; ++pl) *reinterpret_cast<uint2 *>(d - pl * p.dzl_plane) = make_uint2(wa[pl], wb[pl]);
                        }
                    }
                }
                if (!p.with_grad) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    continue;
                }
                T3_MARK(333);
                t3_warp_ready(rdy, lane);
                T3_MARK(41);
                // ---- backward ----------------------------------------------------------------------------------
                for (int l = L - 1; l >= 1; --l) {
                    tc_wait(full, par_full); par_full ^= 1u;
                    T3_MARK(50 + l);
                    // dZ_{l-1} = dA_l * act'(A_l), written over A_l.  Only dA_l has been waited for: the arithmetic and
                    // the split into planes run while the tensor pipe still works on dW_l / db_l, which READ A_l -- so the
                    // stores wait for the second barrier.
                    uint32_t w[CPT / 8][4][P];
                    unsigned char *abuf = sl + p.off_act[l];
                    if (wlive) {
                        const float leak = p.net.leak[l - 1];
#pragma unroll
                        for (int h = 0; h < CPT / 16; ++h) {
                            const int c0 = cgi * CPT + 16 * h;
                            float a[16], g[16], g2[16];
                            t3_ld16_nowait(za + c0, g);
                            if (NB == 2) t3_ld16_nowait(za + T3_HID + c0, g2);
                            if (p.col_save[slot] >= 0) {
                                t3_ld16_nowait(tlane + p.col_save[slot] + (l - 1) * T3_HID + c0, a);
                            } else {
                                t3_load_chunk<P, F, CB>(abuf, p.act_plane[l], c0 / 8, row, a);
                                t3_load_chunk<P, F, CB>(abuf, p.act_plane[l], c0 / 8 + 1, row, a + 8);
                            }
                            t3_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j) g[j] = mq_act_grad(a[j], t3_blocks<NB, F>(g[j], g2[j]), HACT, leak);
#pragma unroll
                            for (int c = 0; c < 2; ++c)
#pragma unroll
                                for (int j = 0; j < 4; ++j) t3_split_pair<P, F>(g[8 * c + 2 * j], g[8 * c + 2 * j + 1], w[2 * h + c][j]);
                        }
                    }
                    T3_MARK(60 + l);
                    tc_wait(dwb, par_dw); par_dw ^= 1u;
                    T3_MARK(70 + l);
                    if (wlive) {
#pragma unroll
                        for (int c = 0; c < CPT / 8; ++c) {
                            unsigned char *d = t3_chunk_ptr<P, CB>(abuf, p.act_plane[l], cgi * (CPT / 8) + c, row);
#pragma unroll
                            for (int pl = 0; pl < P; ++pl)
                                *reinterpret_cast<uint4 *>(d - pl * p.act_plane[l]) = make_uint4(w[c][0][pl], w[c][1][pl], w[c][2][pl], w[c][3][pl]);
                        }
                    }
                    t3_warp_ready(rdy, lane);
                    T3_MARK(80 + l);
                }
                pending = true;                         // dW_0 is in flight
                // two input buffers: the slot's next tile of this round is staged now, under dW_0 -- its F0 queues right
                // behind dW_0 on the tensor pipe and nobody waits for dW_0 by itself
                if (p.x_bufs == 2 && ord + NS < i1) {
                    stage(tcount & 1);
                    t3_warp_ready(rdx, lane);
                    staged = true;
                }
            }
            if (p.with_grad) {
                if (pending) { tc_wait(dwb, par_dw); par_dw ^= 1u; pending = false; }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                T3_MARK(90);
                flush_grad();
                T3_MARK(91);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncthreads();
            }
        }
        if (p.with_grad && !flushed) {                  // a CTA without tiles still owns a partial vector
            float *dst = partial + (int64_t)bid * p.net.n_params;
            for (int j = tid; j < p.net.n_params; j += T3_NT) dst[j] = 0.0f;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}

// ---- operand image and records ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(MQ_NT)
tc3_image_kernel(T3Img g, const float *__restrict__ theta, unsigned char *__restrict__ image) {
    const int j = blockIdx.x * MQ_NT + threadIdx.x;
    if (j < g.n_params) t3_img_store_param_rt(g, image, j, theta[j]);
}

// rec[r] = x[r, :n_in] | tgt[r, :n_out] | w0[r] | w1[r] | 0 ...   (one thread per 128-bit word of a record)
__global__ void __launch_bounds__(MQ_NT)
pack_records_kernel(const float *__restrict__ x, const float *__restrict__ tgt, const float *__restrict__ w0,
                    const float *__restrict__ w1, int64_t rows, int n_in, int n_out, int rec_w, float *__restrict__ rec) {
    const int q4 = rec_w / 4;
    const int64_t item = (int64_t)blockIdx.x * MQ_NT + threadIdx.x;
    if (item >= rows * q4) return;
    const int64_t r = item / q4;
    const int q = (int)(item - r * q4);
    float v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const int f = 4 * q + e;
        float val = 0.0f;
        if (f < n_in) val = x[r * n_in + f];
        else if (f < n_in + n_out) val = tgt ? tgt[r * n_out + f - n_in] : 0.0f;
        else if (f == n_in + n_out) val = w0 ? w0[r] : 0.0f;
        else if (f == n_in + n_out + 1) val = w1 ? w1[r] : 0.0f;
        v[e] = val;
    }
    reinterpret_cast<float4 *>(rec)[item] = make_float4(v[0], v[1], v[2], v[3]);
}

// ---- host side ----------------------------------------------------------------------------------------------------
long long *mq_phase_prof_ptr();
static int t3_planes(int32_t tc) { const int pl = tc & MQ_TC_PLANES_MASK; return pl == 0 ? 3 : pl; }
static int t3_fmt(int32_t tc) { return (tc & MQ_TC_FP16) ? 1 : 0; }

extern "C" int32_t mq_record_width(const mq_net_t *net) {
    if (!net) return 0;
    return (int32_t)mq_round_up(net->n_in + net->n_out + 2, 16);
}

// one (ns = 1) or two (ns = 2) 128-row tile slots
static int t3_plan_try(const mq_net_t *net, int64_t batch, int with_grad, int planes, int fmt, int rec_w, int ns, int x_bufs,
                       int64_t sms, T3Plan *p) {
    const int rows = T3_ROWS;
    memset(p, 0, sizeof(*p));
    p->net = *net;
    if (!t3_img_build_layout(net, planes, fmt, &p->img)) return 0;
    const int L = net->n_layers, k0 = p->img.k0;
    const int cb = rows * 16;
    p->L = L;
    p->with_grad = with_grad;
    p->rec_w = rec_w; p->q4 = rec_w / 4;
    if (rec_w != mq_record_width(net) || rec_w > 64) return 0;
    if (ns == 2 && k0 != 16) return 0;
    if (rows * p->q4 > T3_MAXLD * (T3_NT / ns)) return 0;     // every thread's record words fit its registers
    p->hin_w = (int)mq_round_up(net->n_out + 2, 4);
    // per-slot layout: [dzl][act0 + ones][hin] cleared once, then [act1 + ones][act2 + ones]
    int off = 0;
    p->dzl_plane = 2 * cb;
    p->off_dzl = off; off += planes * p->dzl_plane;
    p->act_plane[0] = (k0 / 8) * cb;
    p->x_bufs = x_bufs;
    p->x_stride = planes * p->act_plane[0] + cb;
    p->hin_stride = rows * p->hin_w;                          // floats
    p->off_act[0] = off; off += x_bufs * p->x_stride;
    p->off_hin = off; off += x_bufs * rows * p->hin_w * 4;
    off = (int)mq_round_up(off, 1024);
    for (int l = 1; l < L; ++l) {
        p->act_plane[l] = (T3_HID / 8) * cb;
        p->off_act[l] = off; off += planes * p->act_plane[l] + cb;
    }
    p->slot_bytes = (int)mq_round_up(off, 1024);
    const int64_t limit = mq_smem_optin() - 256;         // static shared memory: TMEM address + seven mbarriers
    p->off_const = ns * p->slot_bytes;
    p->smem_bytes = (int)mq_round_up(p->off_const + p->img.bytes, 16);
    if (p->smem_bytes > limit) return 0;
    // TMEM columns: Z / dA of each slot (NB blocks of 64), then the shared dW / db accumulators
    const int nbk = planes > 1 ? 2 : 1;
    const int ones_cols = fmt == 1 ? 16 : 8;              // bias-gradient columns behind the dW blocks (fp16: main + scaled)
    int col = 0;
    for (int s = 0; s < ns; ++s) { p->col_za[s] = col; col += nbk * T3_HID; }
    {   // fp32 copy of the hidden activations for the backward pass, when the columns are there
        int acc_cols = 0;
        for (int l = 0; l < L; ++l) acc_cols += (l == L - 1) ? nbk * T3_NO : nbk * p->img.kp[l] + ones_cols;
        const bool fits = with_grad && col + ns * (L - 1) * T3_HID + acc_cols <= 512;
        for (int s = 0; s < 2; ++s) p->col_save[s] = -1;
        if (fits)
            for (int s = 0; s < ns; ++s) { p->col_save[s] = col; col += (L - 1) * T3_HID; }
    }
    for (int l = 0; l < L; ++l) { p->col_acc[l] = col; col += (l == L - 1) ? nbk * T3_NO : nbk * p->img.kp[l] + ones_cols; }
    if (col > 512) return 0;
    // balance: every tile slot of a full wave gets the same number of rows
    const int64_t waves = mq_ceil_div(batch, sms * ns * rows);
    int64_t tr = mq_round_up(mq_ceil_div(batch, sms * ns * waves), 16);
    if (tr > rows) tr = rows;
    if (tr < 16) tr = 16;
    p->tile_rows = (int)tr;
    return 1;
}

// Two tile slots (two tiles in flight) for one and two planes when they fit next to the operand image, else one slot.
// Measured and dropped for the exact three-plane mode, whose two 128-row tiles do not fit (profiles/r2_tc3_modes.md):
// two 64-row slots fed by half-filled M = 128 instructions (74 us against 57 us at 65,536 rows: the doubled tensor work
// is not hidden), and publishing the activations K chunk by K chunk so that layer l + 1's products start under layer l's
// epilogue (61 us against 57 us: four fence / arrive rounds per phase cost more than the ~2 k cycles of product time
// they hide per tile).
static int t3_plan_build(const mq_net_t *net, int64_t batch, int with_grad, int32_t tc, int rec_w, int64_t sms, T3Plan *p,
                         int *ns_out) {
    const int planes = t3_planes(tc), fmt = t3_fmt(tc);
    // preference: two slots; within a slot count, two input buffers (the next tile is staged under the current tile's last
    // product) when shared memory allows
    // A batch of at most one 128-row tile per SM gains nothing from a second slot (there is no second tile to overlap with)
    // and would split its tile into two half-filled ones: one slot, all 16 warps on the tile's epilogues.
    const bool single_wave = mq_ceil_div(batch, T3_ROWS) <= sms;
    for (int pass = 0; pass < 2; ++pass) {
        const int ns = (pass == 0) == single_wave ? 1 : 2;
        for (int xb = 2; xb >= 1; --xb)
            if (t3_plan_try(net, batch, with_grad, planes, fmt, rec_w, ns, xb, sms, p)) {
                *ns_out = ns;
                return 1;
            }
    }
    return 0;
}

extern "C" int64_t mq_tc_image_bytes(const mq_net_t *net, int32_t tc) {
    T3Img g;
    if (!net || !t3_img_build_layout(net, t3_planes(tc), t3_fmt(tc), &g)) return 0;
    return g.bytes;
}

static int t3_image_build(const T3Img &g, const float *theta, void *image, void *stream) {
    cudaError_t e = cudaMemsetAsync(image, 0, (size_t)g.bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) { mq_set_error("mq_tc_image_build: %s", cudaGetErrorString(e)); return MQ_ERR_CUDA; }
    MQ_LAUNCH(tc3_image_kernel, (unsigned)mq_ceil_div(g.n_params, MQ_NT), MQ_NT, 0, stream, g, theta, (unsigned char *)image);
    MQ_CHECK_LAUNCH("mq_tc_image_build");
    return MQ_OK;
}

extern "C" int mq_tc_image_build(const mq_net_t *net, const float *theta, int32_t tc, void *image, void *stream) {
    MQ_REQUIRE(net && theta && image, MQ_ERR_INVALID, "mq_tc_image_build: null pointer");
    MQ_REQUIRE(((uintptr_t)image & 15) == 0, MQ_ERR_INVALID, "mq_tc_image_build: image must be 16-byte aligned");
    T3Img g;
    MQ_REQUIRE(t3_img_build_layout(net, t3_planes(tc), t3_fmt(tc), &g), MQ_ERR_UNSUPPORTED,
               "mq_tc_image_build: net does not fit the tensor-core kernel (or MQ_TC_FP16 without two planes)");
    return t3_image_build(g, theta, image, stream);
}

extern "C" int mq_pack_records(const mq_net_t *net, const float *x, const float *tgt, const float *w0, const float *w1,
                               int64_t rows, float *rec, void *stream) {
    MQ_REQUIRE(net && x && rec, MQ_ERR_INVALID, "mq_pack_records: null pointer");
    MQ_REQUIRE(rows >= 0, MQ_ERR_SHAPE, "mq_pack_records: negative row count");
    MQ_REQUIRE(((uintptr_t)rec & 15) == 0, MQ_ERR_INVALID, "mq_pack_records: records must be 16-byte aligned");
    if (rows == 0) return MQ_OK;
    const int rec_w = mq_record_width(net);
    const int64_t items = rows * (rec_w / 4);
    MQ_LAUNCH(pack_records_kernel, (unsigned)mq_ceil_div(items, MQ_NT), MQ_NT, 0, stream, x, tgt, w0, w1, rows, net->n_in,
              net->n_out, rec_w, rec);
    MQ_CHECK_LAUNCH("mq_pack_records");
    return MQ_OK;
}

// bytes mq_tc3_launch may use behind the partial gradients for an image it has to build itself
int64_t mq_tc3_scratch_bytes() { return 96 * 1024; }

// One net's job on `sms` SMs: plan, grid, operand image (built into the workspace when the caller has none).
static int t3_job_build(const mq_net_t *net, const int32_t *row_idx, int64_t batch, const mq_head_t *head, const float *theta,
                        int with_grad, float *out, float *logp, void *ws, int64_t ws_bytes, int64_t sms, T3Job *j, int *ns_out,
                        void *stream) {
    if (!t3_plan_build(net, batch, with_grad, head->tc, head->rec_w, sms, &j->p, ns_out)) return MQ_ERR_UNSUPPORTED;
    MQ_REQUIRE(((uintptr_t)head->records & 15) == 0, MQ_ERR_INVALID, "mq_tc3_launch: records must be 16-byte aligned");
    const int64_t ntiles = mq_ceil_div(batch, j->p.tile_rows);
    j->grid = (int)(ntiles < sms ? ntiles : sms);
    const int64_t part_bytes = with_grad ? mq_round_up((int64_t)j->grid * net->n_params * (int64_t)sizeof(float), 256) : 0;
    if (with_grad) MQ_REQUIRE(ws && ws_bytes >= part_bytes, MQ_ERR_INVALID, "mq_tc3_launch: workspace too small");
    const void *image = head->tc_image;
    if (!image) {
        MQ_REQUIRE(theta && ws && ws_bytes >= part_bytes + j->p.img.bytes, MQ_ERR_INVALID,
                   "mq_tc3_launch: no operand image and no room in the workspace to build one");
        void *scratch = (char *)ws + part_bytes;
        const int rc = t3_image_build(j->p.img, theta, scratch, stream);
        if (rc != MQ_OK) return rc;
        image = scratch;
    }
    MQ_REQUIRE(((uintptr_t)image & 15) == 0, MQ_ERR_INVALID, "mq_tc3_launch: image must be 16-byte aligned");
    j->image = (const unsigned char *)image;
    j->rec = head->records; j->idx = row_idx; j->n_rows = batch; j->head = *head;
    j->partial = (float *)ws; j->out = out; j->logp_out = logp;
    return MQ_OK;
}

static int t3_launch_jobs(const T3Jobs &jobs, int total_grid, int planes, int fmt, int ns, int k0, int act, int smem_bytes,
                          void *stream) {
    cudaError_t e = cudaSuccess;
#define T3_LAUNCH(P_, F_, NS_, K0_, ACT_)                                                                                   \
    do {                                                                                                                    \
        e = cudaFuncSetAttribute(tc3_grad_kernel<P_, F_, NS_, K0_, ACT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes); \
        if (e == cudaSuccess)                                                                                               \
            e = mq_launch_pdl(tc3_grad_kernel<P_, F_, NS_, K0_, ACT_>, dim3(total_grid), dim3(T3_NTI), (size_t)smem_bytes, stream, \
                              jobs, mq_phase_prof_ptr());                                                                   \
    } while (0)
#define T3_LAUNCH_ACT(P_, F_, NS_, K0_)                                      \
    do {                                                                     \
        if (act == MQ_ACT_TANH) T3_LAUNCH(P_, F_, NS_, K0_, MQ_ACT_TANH);    \
        else T3_LAUNCH(P_, F_, NS_, K0_, MQ_ACT_LRELU);                      \
    } while (0)
#define T3_LAUNCH_P(P_, F_)                                              \
    do {                                                                 \
        if (ns == 2) T3_LAUNCH_ACT(P_, F_, 2, 16);                       \
        else if (k0 == 16) T3_LAUNCH_ACT(P_, F_, 1, 16);                 \
        else T3_LAUNCH_ACT(P_, F_, 1, 32);                               \
    } while (0)
    if (planes == 3) { if (k0 == 16) T3_LAUNCH_ACT(3, 0, 1, 16); else T3_LAUNCH_ACT(3, 0, 1, 32); }
    else if (planes == 2 && fmt == 1) T3_LAUNCH_P(2, 1);
    else if (planes == 2) T3_LAUNCH_P(2, 0);
    else T3_LAUNCH_P(1, 0);
#undef T3_LAUNCH_P
#undef T3_LAUNCH_ACT
#undef T3_LAUNCH
    if (e != cudaSuccess) {
        mq_set_error("mq_tc3_launch: %s", cudaGetErrorString(e));
        return MQ_ERR_CUDA;
    }
    MQ_CHECK_LAUNCH("mq_tc3_launch");
    return MQ_OK;
}

// Launches the record-fed tensor-core gradient kernel.  MQ_ERR_UNSUPPORTED when the net does not fit.
// Partials go to ws ([grid][n_params]); *grid_out = CTAs launched.
int mq_tc3_launch(const mq_net_t *net, const int32_t *row_idx, int64_t batch, const mq_head_t *head, const float *theta,
                  int with_grad, float *out, float *logp, void *ws, int64_t ws_bytes, int *grid_out, void *stream) {
    T3Jobs local;
    memset(&local, 0, sizeof(local));
    int ns = 1;
    const int rc = t3_job_build(net, row_idx, batch, head, theta, with_grad, out, logp, ws, ws_bytes, mq_sm_count(), &local.job[0],
                                &ns, stream);
    if (rc != MQ_OK) return rc;
    const int rc2 = t3_launch_jobs(local, local.job[0].grid, t3_planes(head->tc), t3_fmt(head->tc), ns, local.job[0].p.img.k0, net->act[0],
                                   local.job[0].p.smem_bytes, stream);
    if (rc2 == MQ_OK && grid_out) *grid_out = local.job[0].grid;
    return rc2;
}

// Two independent minibatch gradients (two nets of the same shape class: planes, slots, padded input width and hidden
// activation must agree) as ONE launch, each on half of the SMs.  MQ_ERR_UNSUPPORTED when they cannot share a launch.
extern "C" int mq_fused_grad_partials2(const mq_net_t *net_a, const float *theta_a, const int32_t *idx_a, const mq_head_t *head_a,
                                       void *ws_a, int64_t ws_a_bytes, int32_t *n_parts_a, const mq_net_t *net_b,
                                       const float *theta_b, const int32_t *idx_b, const mq_head_t *head_b, void *ws_b,
                                       int64_t ws_b_bytes, int32_t *n_parts_b, int64_t batch, void *stream) {
    MQ_REQUIRE(net_a && net_b && head_a && head_b && ws_a && ws_b && n_parts_a && n_parts_b, MQ_ERR_INVALID,
               "mq_fused_grad_partials2: null pointer");
    MQ_REQUIRE(batch >= 1, MQ_ERR_SHAPE, "mq_fused_grad_partials2: empty batch");
    if (!head_a->records || !head_b->records || ((head_a->tc | head_b->tc) & MQ_TC_OFF)) return MQ_ERR_UNSUPPORTED;
    if (t3_planes(head_a->tc) != t3_planes(head_b->tc) || t3_fmt(head_a->tc) != t3_fmt(head_b->tc) || net_a->act[0] != net_b->act[0])
        return MQ_ERR_UNSUPPORTED;
    T3Jobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    const int64_t sms = mq_sm_count();
    const int64_t half = sms / 2 > 0 ? sms / 2 : 1;
    int ns_a = 1, ns_b = 1;
    int rc = t3_job_build(net_a, idx_a, batch, head_a, theta_a, 1, nullptr, nullptr, ws_a, ws_a_bytes, half, &jobs.job[0], &ns_a, stream);
    if (rc != MQ_OK) return rc;
    rc = t3_job_build(net_b, idx_b, batch, head_b, theta_b, 1, nullptr, nullptr, ws_b, ws_b_bytes, sms - half, &jobs.job[1], &ns_b, stream);
    if (rc != MQ_OK) return rc;
    if (ns_a != ns_b || jobs.job[0].p.img.k0 != jobs.job[1].p.img.k0) return MQ_ERR_UNSUPPORTED;
    const int smem = jobs.job[0].p.smem_bytes > jobs.job[1].p.smem_bytes ? jobs.job[0].p.smem_bytes : jobs.job[1].p.smem_bytes;
    rc = t3_launch_jobs(jobs, jobs.job[0].grid + jobs.job[1].grid, t3_planes(head_a->tc), t3_fmt(head_a->tc), ns_a, jobs.job[0].p.img.k0, net_a->act[0],
                        smem, stream);
    if (rc != MQ_OK) return rc;
    *n_parts_a = jobs.job[0].grid;
    *n_parts_b = jobs.job[1].grid;
    return MQ_OK;
}

#else   // MQ_EMU: the host emulation has no tensor cores
extern "C" int32_t mq_record_width(const mq_net_t *net) { return net ? (int32_t)mq_round_up(net->n_in + net->n_out + 2, 16) : 0; }
extern "C" int64_t mq_tc_image_bytes(const mq_net_t *, int32_t) { return 0; }
extern "C" int mq_tc_image_build(const mq_net_t *, const float *, int32_t, void *, void *) { return MQ_ERR_UNSUPPORTED; }
extern "C" int mq_pack_records(const mq_net_t *, const float *, const float *, const float *, const float *, int64_t, float *,
                               void *) { return MQ_ERR_UNSUPPORTED; }
int64_t mq_tc3_scratch_bytes() { return 0; }
int mq_tc3_launch(const mq_net_t *, const int32_t *, int64_t, const mq_head_t *, const float *, int, float *, float *, void *,
                  int64_t, int *, void *) { return MQ_ERR_UNSUPPORTED; }
extern "C" int mq_fused_grad_partials2(const mq_net_t *, const float *, const int32_t *, const mq_head_t *, void *, int64_t, int32_t *,
                                       const mq_net_t *, const float *, const int32_t *, const mq_head_t *, void *, int64_t, int32_t *,
                                       int64_t, void *) { return MQ_ERR_UNSUPPORTED; }
#endif
