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
// An MMA of this size costs max(~32 cycles, operand bytes / 128 B per cycle, M*N/256) (measured,
// tools/tc_mma_bench.cu), so the plane pairs are NOT issued one by one: planes are laid out next to
// each other and one instruction covers several pairs through a wide N
//   forward / dA   A_0 x [W_0|W_1], A_1 x [W_0|W_1], A_0 x W_2, A_2 x W_0   (two 64-column blocks that
//                  the epilogue adds)
//   dW^T, db       dZ_0 x [A_2|A_1|A_0|1], dZ_1 x [A_1|A_0|1], dZ_2 x [A_0|1] into one accumulator
//                  whose three column blocks (+ the ones column = bias gradient) are added at the end.
//
// One CTA (512 threads) per SM walks tiles of <= 128 batch rows:
//   smem  operands as 8 x 16 B core matrices (SWIZZLE_NONE).  An activation buffer is
//         [plane 2,1,0][feature chunk of 8][row] + one chunk of ones, so that the SAME bytes serve as
//         the K-major A operand of the next layer / of dA (rows = M) and, planes side by side, as the
//         MN-major operand of dW (rows = K).  Weights are staged once per launch in two images:
//         [k chunk][plane][n] (K-major B of the forward product, planes side by side along N) and
//         [plane][k chunk][n] (MN-major B of dA = dZ W^T, planes side by side along N').
//   TMEM  lane = batch row for Z / dA (128 columns, reused by every layer); the dW / db accumulators
//         (M = 64: row i in lane i%16 + 32*(i/16)) persist across all tiles of the CTA.
//   dZ_l overwrites A_{l+1} in shared memory in place (same thread, same element), after the
//   commit that covers the last MMA reading A_{l+1}.
// MMAs are issued by one elected lane inside warp-uniform control flow (descriptors stay in uniform
// registers); completion is a tcgen05.commit on one mbarrier per phase.
// Per-CTA partial gradients go to the workspace and are summed in fixed order by
// fused_reduce_kernel (mq_fused.cu): bit-reproducible run to run.
#include "mq_tc.cuh"

#ifndef MQ_EMU

#define HALF_LOG_2PI 0.91893853320467274178f
#define TC_NT 512                // worker threads (16 warps)
#define TC_NTI 544               // + one warp that only issues MMAs (tc_grad_kernel)
#define TC_ROWS 128              // MMA M of the row-major products
#define TC_HID 64                // padded hidden width
#define TC_NO 16                 // padded output width
#define TC_CHUNK_BYTES 2048      // one 8-feature chunk of all 128 rows (128 rows x 16 B)
#define TC_MAXL 3
#define TC_FLUSH_TILES 4         // tiles accumulated in TMEM before the dW accumulators are folded into memory
#define TC_HIN 20                // floats of head input per tile row: 16 targets, weight / advantage, old log-prob

struct TcPlan {
    mq_net_t net;
    int L;
    int k0;                      // padded input width (16 or 32)
    int kp[TC_MAXL], np[TC_MAXL];
    int tile_rows;               // rows per tile, multiple of 16, <= 128
    int x_step_q, x_step_r, h_step_q, h_step_r;   // TC_NT / n_in, TC_NT % n_in, TC_NT / n_out, TC_NT % n_out
    int with_grad;
    // shared-memory byte offsets.  act[l]: planes 2,1,0 then the ones chunk; dzl: planes 2,1,0
    int off_dzl, off_act[TC_MAXL], act_plane[TC_MAXL], off_wf[TC_MAXL], off_wb[TC_MAXL], wb_plane[TC_MAXL];
    int off_bias, off_misc, off_hin, off_idx;
    int smem_bytes;
    // TMEM columns
    int col_za, col_save, col_acc[TC_MAXL], col_rb;
};

// address of (chunk, row) in plane 0 of an activation buffer (planes are stored 2,1,0)
__device__ __forceinline__ unsigned char *tc_chunk_ptr(unsigned char *buf, int plane_bytes, int chunk, int row) {
    return buf + 2 * plane_bytes + chunk * TC_CHUNK_BYTES + (row >> 3) * 128 + (row & 7) * 16;
}

// store 8 consecutive features of one row as one 16-byte chunk in each of the three planes
__device__ __forceinline__ void tc_store_chunk(unsigned char *buf, int plane_bytes, int chunk, int row, const float *v) {
    uint32_t w[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tc_split_pair(v[2 * j], v[2 * j + 1], w[0][j], w[1][j], w[2][j]);
    unsigned char *p = tc_chunk_ptr(buf, plane_bytes, chunk, row);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
        *reinterpret_cast<uint4 *>(p - pl * plane_bytes) = make_uint4(w[pl][0], w[pl][1], w[pl][2], w[pl][3]);
}

// the exact fp32 values back from the three planes
__device__ __forceinline__ void tc_load_chunk(const unsigned char *buf, int plane_bytes, int chunk, int row, float *v) {
    const unsigned char *p = tc_chunk_ptr(const_cast<unsigned char *>(buf), plane_bytes, chunk, row);
    const uint4 a = *reinterpret_cast<const uint4 *>(p);
    const uint4 b = *reinterpret_cast<const uint4 *>(p - plane_bytes);
    const uint4 c = *reinterpret_cast<const uint4 *>(p - 2 * plane_bytes);
    const uint32_t wa[4] = {a.x, a.y, a.z, a.w}, wb[4] = {b.x, b.y, b.z, b.w}, wc[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[2 * j] = (__uint_as_float(wa[j] << 16) + __uint_as_float(wb[j] << 16)) + __uint_as_float(wc[j] << 16);
        v[2 * j + 1] = (__uint_as_float(wa[j] & 0xFFFF0000u) + __uint_as_float(wb[j] & 0xFFFF0000u)) +
                       __uint_as_float(wc[j] & 0xFFFF0000u);
    }
}

// ---- MMA issue helpers ----------------------------------------------------------------------------
// (called by the elected lane of a warp inside warp-uniform control flow, so that descriptors live in
// uniform registers).  A view is the two descriptor words of its first K step plus the K-step advance.

// activation buffer plane at `addr` viewed K-major (M = rows, K = features): LBO = chunk stride, SBO = row-group stride
__device__ __forceinline__ TcView tc_rows_k(uint32_t addr) {
    return TcView{tc_desc_lo(addr, TC_CHUNK_BYTES), tc_desc_hi(128u), (2u * TC_CHUNK_BYTES) >> 4};
}
// activation buffer from `addr` on viewed MN-major (M or N = features, planes side by side; K = rows)
__device__ __forceinline__ TcView tc_feat_mn(uint32_t addr) {
    return TcView{tc_desc_lo(addr, 128u), tc_desc_hi(TC_CHUNK_BYTES), 256u >> 4};
}
// D[128 x (two blocks of nb)] = sum over plane pairs of A_i B_j, A row-major planes (2,1,0 in memory, `a_plane`
// bytes apart, a0 = plane 0), B a weight image whose planes sit side by side along N (b_plane16 = plane stride >> 4
// in descriptor units).  Four instructions per K step.
__device__ __forceinline__ void tc_rows_product(uint32_t d_tmem, uint32_t a0_addr, uint32_t a_plane, const TcView &b,
                                                int b_mn, uint32_t b_plane16, int nb, int ksteps) {
    const TcView a0 = tc_rows_k(a0_addr);
    const uint32_t ap16 = a_plane >> 4;
    const uint32_t id2 = tc_idesc(TC_ROWS, 2 * nb, 0, b_mn);
    const uint32_t id1 = tc_idesc(TC_ROWS, nb, 0, b_mn);
    uint32_t alo = a0.lo, blo = b.lo;
#pragma unroll 1
    for (int k = 0; k < ksteps; ++k) {
        tc_mma(d_tmem, alo, a0.hi, blo, b.hi, id2, k > 0 ? 1u : 0u);             // A_0 x [B_0|B_1]
        tc_mma(d_tmem, alo - ap16, a0.hi, blo, b.hi, id2, 1u);                    // A_1 x [B_0|B_1]
        tc_mma(d_tmem, alo, a0.hi, blo + 2 * b_plane16, b.hi, id1, 1u);           // A_0 x B_2
        tc_mma(d_tmem, alo - 2 * ap16, a0.hi, blo, b.hi, id1, 1u);                // A_2 x B_0
        alo += a0.k16;
        blo += b.k16;
    }
}

// acc[64 x (2 nb + extra)] (+)= sum over plane pairs of Da_i^T Db_j over the rows of the tile: Da, Db activation
// buffers (planes 2,1,0 in memory), Da's features are M, Db's planes 1,0 side by side (+ `extra` columns behind
// plane 0) are N.  Column blocks: [Db_1 | Db_0 | extra]; Da_0 x Db_2 is added into the first block (the final
// epilogue adds the blocks anyway).  Four instructions per K step of 16 rows.
__device__ __forceinline__ void tc_feat_product(uint32_t acc_tmem, uint32_t da_base, uint32_t da_plane, uint32_t db_base,
                                                uint32_t db_plane, int nb, int extra, int ksteps, bool accumulate) {
    const TcView a = tc_feat_mn(da_base), b = tc_feat_mn(db_base);
    const uint32_t ap16 = da_plane >> 4, bp16 = db_plane >> 4;
    const uint32_t id_all = tc_idesc(64, 2 * nb + extra, 1, 1), id_one = tc_idesc(64, nb + extra, 1, 1), id_blk = tc_idesc(64, nb, 1, 1);
    uint32_t alo = a.lo, blo = b.lo;
#pragma unroll 1
    for (int k = 0; k < ksteps; ++k) {
        tc_mma(acc_tmem, alo + 2 * ap16, a.hi, blo + bp16, b.hi, id_all, (accumulate || k > 0) ? 1u : 0u);   // Da_0 x [Db_1|Db_0|e]
        tc_mma(acc_tmem, alo + ap16, a.hi, blo + bp16, b.hi, id_all, 1u);                                    // Da_1 x [Db_1|Db_0|e]
        tc_mma(acc_tmem + nb, alo, a.hi, blo + 2 * bp16, b.hi, id_one, 1u);                                  // Da_2 x [Db_0|e]
        tc_mma(acc_tmem, alo + 2 * ap16, a.hi, blo, b.hi, id_blk, 1u);                                       // Da_0 x Db_2
        alo += a.k16;
        blo += b.k16;
    }
}

// ---- the kernel ---------------------------------------------------------------------------
template <int K0, int HACT>
__global__ void __launch_bounds__(TC_NTI, 1)
tc_grad_kernel(TcPlan p, const float *__restrict__ theta, const float *__restrict__ x,
               const int32_t *__restrict__ idx, int64_t n_rows, mq_head_t head,
               float *__restrict__ partial, float *__restrict__ out, float *__restrict__ logp_out, long long *prof) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ __align__(8) uint64_t s_mbar_dw;      // completion of the dW / db products of a backward phase (see below)
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp-uniform for the compiler too
    // Warps 0..15 are WORKERS (gathers, epilogues, head); warp 16 only ISSUES the MMAs.  A thread that issues MMAs is held up
    // by the tensor pipe's back-pressure for about as long as the products take, so a worker warp doing it delayed every
    // epilogue that could have run under the dW products.  The issuer meets the workers at the publish barriers only.
    const bool worker = warp < TC_NT / 32;
    const int stid = worker ? tid : (1 << 28);                    // start index of the workers' strided loops
    // developer aid (compile with -DMQ_TC_PROFILE): per-phase cycles of CTA 0 and every CTA's start / end time
#ifdef MQ_TC_PROFILE
    long long t_last = 0;
    const bool profiling = prof != nullptr && blockIdx.x == 0 && tid == 0;
    if (profiling) t_last = clock64();
    if (prof != nullptr && tid == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); prof[64 + 2 * blockIdx.x] = (long long)gt; }
#define TC_MARK(i) do { if (profiling) { long long now_ = clock64(); prof[i] += now_ - t_last; t_last = now_; } } while (0)
#else
    (void)prof;
#define TC_MARK(i) do { } while (0)
#endif
    const int L = p.L;
    const int n_in = p.net.n_in, n_out = p.net.n_out;
    constexpr int XCHUNKS = K0 / 8;

    // ---- one-time setup: TMEM, barrier, operand images of the weights --------------------------
    // (launched with programmatic stream serialisation: everything up to mq_pdl_wait() is on-chip work that overlaps
    //  with the tail of the previous kernel in the stream -- in the PPO loop the reduce + Adam kernel that writes theta)
    mq_pdl_trigger();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&s_mbar)), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&s_mbar_dw)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    TC_MARK(21);
    // clear shared memory except the hidden activation buffers (every tile rewrites all of their rows)
    for (int i = stid; i < p.off_act[1] / 16; i += TC_NT) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = p.off_wf[0] / 16 + stid; i < p.smem_bytes / 16; i += TC_NT) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    TC_MARK(22);
    {
        __nv_bfloat16 one = __float2bfloat16_rn(1.0f);
        const uint32_t ob = *reinterpret_cast<uint16_t *>(&one);
        for (int l = 0; l < L; ++l) {
            uint32_t *ones = reinterpret_cast<uint32_t *>(smem + p.off_act[l] + 3 * p.act_plane[l]);
            for (int i = stid; i < TC_CHUNK_BYTES / 4; i += TC_NT) ones[i] = ob | (ob << 16);
        }
    }
    float *sBias = reinterpret_cast<float *>(smem + p.off_bias);      // [L][64]
    float *sMisc = reinterpret_cast<float *>(smem + p.off_misc);      // logstd[16], exp(-logstd)[16]
    float *sHin = reinterpret_cast<float *>(smem + p.off_hin);        // head inputs of the tile: [128][TC_HIN]
    int *sIdx = reinterpret_cast<int *>(smem + p.off_idx);            // source rows of three consecutive tiles: [3][128]
    TC_MARK(23);
    mq_pdl_wait();                        // first global read below: the predecessor's writes are complete and visible
    // weight images: one work item = 8 consecutive inputs (one 16-byte chunk) of one output column
    {
        int items[TC_MAXL + 1];
        items[0] = 0;
        for (int l = 0; l < L; ++l) items[l + 1] = items[l] + (p.kp[l] / 8) * p.net.out[l];
#pragma unroll 2
        for (int it = stid; it < items[L]; it += TC_NT) {
            int l = 0;
#pragma unroll
            for (int q = 1; q < TC_MAXL; ++q) if (q < L && it >= items[q]) l = q;
            const int on = p.net.out[l], np = p.np[l], in = p.net.in[l];
            const int local = it - items[l];
            const int kc = local / on, n = local - kc * on;
            const float *w = theta + p.net.w_off[l] + (kc * 8) * on + n;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (kc * 8 + j < in) ? __ldg(w + j * on) : 0.0f;
            uint32_t q[3][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) tc_split_pair(v[2 * j], v[2 * j + 1], q[0][j], q[1][j], q[2][j]);
            unsigned char *df = smem + p.off_wf[l] + kc * (3 * np * 16) + n * 16;          // [k chunk][plane][n][8 k]
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
                *reinterpret_cast<uint4 *>(df + pl * np * 16) = make_uint4(q[pl][0], q[pl][1], q[pl][2], q[pl][3]);
            if (l > 0 && p.with_grad) {
                unsigned char *db = smem + p.off_wb[l] + kc * (np * 16) + n * 16;           // [plane][k chunk][n][8 k]
#pragma unroll
                for (int pl = 0; pl < 3; ++pl)
                    *reinterpret_cast<uint4 *>(db + pl * p.wb_plane[l]) = make_uint4(q[pl][0], q[pl][1], q[pl][2], q[pl][3]);
            }
        }
        for (int l = 0; l < L; ++l)
            for (int n = stid; n < p.net.out[l]; n += TC_NT) sBias[l * 64 + n] = __ldg(theta + p.net.b_off[l] + n);
        if (p.net.logstd_off >= 0 && tid < n_out) {
            const float ls = __ldg(theta + p.net.logstd_off + tid);
            sMisc[tid] = ls;
            sMisc[16 + tid] = expf(-ls);
        }
    }
    TC_MARK(24);
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may touch
    const int cg = warp >> 2;                           // which 16 of the 64 columns this thread owns
    const int row = quarter * 32 + lane;                // TMEM lane == batch row of the tile
    const int tr = p.tile_rows;
    const int64_t ntiles = (n_rows + tr - 1) / tr;
    const int dw_ksteps = tr / 16;
    const bool gauss = mq_head_gauss(head.kind), prob = mq_head_prob(head.kind);
    const int lastact = p.net.act[L - 1];
    const float lastleak = p.net.leak[L - 1];

    // ---- gathers.  Lanes run along the FEATURES of a row (a warp-wide load touches ~3 rows, not 32), the source
    // row of every tile row is staged in shared memory two tiles ahead, the features of the next tile wait in
    // registers until the MMAs reading the current input tile are done.
    constexpr int XE = K0 / 4;                          // elements of the 128 x n_in input tile per thread
    constexpr int HE = TC_NO * TC_ROWS / TC_NT;         // elements of the 128 x n_out target tile per thread
    float xr[XE];
    int xrf[XE], hrf[HE];                               // (row << 8 | feature) of this thread's elements, -1: none
    {
        int r = tid / n_in, f = tid - r * n_in;
#pragma unroll
        for (int j = 0; j < XE; ++j) {                  // e = tid + j * TC_NT, stepped without divisions
            xrf[j] = (worker && r < TC_ROWS) ? ((r << 8) | f) : -1;
            r += p.x_step_q; f += p.x_step_r;
            if (f >= n_in) { f -= n_in; ++r; }
        }
        r = tid / n_out; f = tid - r * n_out;
#pragma unroll
        for (int j = 0; j < HE; ++j) {
            hrf[j] = (worker && r < TC_ROWS) ? ((r << 8) | f) : -1;
            r += p.h_step_q; f += p.h_step_r;
            if (f >= n_out) { f -= n_out; ++r; }
        }
    }
    // gathered values are only LOADED while the first products run and are written to shared memory one epilogue later,
    // so that nobody waits for HBM
    int pend_idx = -1;
    float pend_t[TC_NO * TC_ROWS / TC_NT], pend_a = 0.0f, pend_l = 0.0f;
    (void)0;
    auto fetch_idx = [&](int64_t tile) {                // threads 0..127: source row (or -1) of every row of `tile`
        pend_idx = -1;
        if (tid < TC_ROWS) {
            const int64_t gr = tile * tr + tid;
            if (tile < ntiles && tid < tr && gr < n_rows) pend_idx = idx ? tc_ldg_s32(idx + gr) : (int)gr;
        }
    };
    auto commit_idx = [&](int slot) { if (tid < TC_ROWS) sIdx[slot * TC_ROWS + tid] = pend_idx; };
    auto load_x = [&](int slot) {
#pragma unroll
        for (int j = 0; j < XE; ++j) {
            float v = 0.0f;
            if (xrf[j] >= 0) {
                const int src = sIdx[slot * TC_ROWS + (xrf[j] >> 8)];
                if (src >= 0) v = tc_ldg_f32(x + (int64_t)src * n_in + (xrf[j] & 255));
            }
            xr[j] = v;
        }
    };
    auto store_x = [&]() {
#pragma unroll
        for (int j = 0; j < XE; ++j) {
            if (xrf[j] < 0) continue;
            const int r = xrf[j] >> 8, f = xrf[j] & 255;
            uint32_t w0, w1, w2;
            tc_split_pair(xr[j], 0.0f, w0, w1, w2);
            unsigned char *d = tc_chunk_ptr(smem + p.off_act[0], p.act_plane[0], f >> 3, r) + (f & 7) * 2;
            *reinterpret_cast<uint16_t *>(d) = (uint16_t)(w0 & 0xFFFFu);
            *reinterpret_cast<uint16_t *>(d - p.act_plane[0]) = (uint16_t)(w1 & 0xFFFFu);
            *reinterpret_cast<uint16_t *>(d - 2 * p.act_plane[0]) = (uint16_t)(w2 & 0xFFFFu);
        }
    };
    auto fetch_head_inputs = [&](int slot) {            // this tile's targets / weights
        const float *mat = (head.kind == MQ_HEAD_DY) ? head.dy : head.target;
#pragma unroll
        for (int j = 0; j < HE; ++j) {
            float v = 0.0f;
            if (mat && hrf[j] >= 0) {
                const int src = sIdx[slot * TC_ROWS + (hrf[j] >> 8)];
                if (src >= 0) v = tc_ldg_f32(mat + (int64_t)src * n_out + (hrf[j] & 255));
            }
            pend_t[j] = v;
        }
        pend_a = 0.0f; pend_l = 0.0f;
        if (prob && tid < TC_ROWS) {
            const int src = sIdx[slot * TC_ROWS + tid];
            if (src >= 0) {
                if (mq_head_ppo(head.kind)) { pend_a = tc_ldg_f32(head.adv + src); pend_l = tc_ldg_f32(head.logp_old + src); }
                else if (head.dy) pend_a = tc_ldg_f32(head.dy + src);
            }
        }
    };
    auto commit_head_inputs = [&]() {
#pragma unroll
        for (int j = 0; j < HE; ++j)
            if (hrf[j] >= 0) sHin[(hrf[j] >> 8) * TC_HIN + (hrf[j] & 255)] = pend_t[j];
        if (prob && tid < TC_ROWS) { sHin[tid * TC_HIN + 16] = pend_a; sHin[tid * TC_HIN + 17] = pend_l; }
    };
    TC_MARK(25);
    fetch_idx(blockIdx.x); commit_idx(0);
    fetch_idx((int64_t)blockIdx.x + gridDim.x); commit_idx(1);
    tc_publish();
    TC_MARK(0);
    const uint32_t tm = __shfl_sync(0xffffffffu, s_tmem, 0);
    const uint32_t mbar = tc_smem_u32(&s_mbar);
    const uint32_t mbar_dw = tc_smem_u32(&s_mbar_dw);
    uint32_t parity_dw = 0;
    const uint32_t sbase = tc_smem_u32(smem);
    const uint32_t tlane = tm + ((uint32_t)(quarter * 32) << 16);
    uint32_t parity = 0;
    bool pending = false;                               // a committed MMA batch nobody waited for yet
    bool dw_started = false;                            // dW / db accumulators hold data
    load_x(0);
    fetch_head_inputs(0);
    commit_head_inputs();

    // TMEM accumulators -> this CTA's partial gradient in the workspace (set on the first flush, added afterwards).
    // The tensor pipe truncates when it adds into its fp32 accumulator (see mq_gemm_tc.cu), so a CTA that walks many
    // tiles folds its accumulators into memory every TC_FLUSH_TILES tiles instead of only once at the end.
    bool flushed = false;
    auto flush_grad = [&]() {
        float *dst = partial + (int64_t)blockIdx.x * p.net.n_params;
        const bool add = flushed;
        if (worker) {
        // M = 64 accumulators: row i is lane i%16 of the warps of quarter i/16; the four warps of a quarter
        // split the columns of a block
        const int i = quarter * 16 + lane;
        const bool lv = lane < 16;
        for (int l = 0; l < L; ++l) {
            const int in = p.net.in[l], on = p.net.out[l];
            const int nb = (l == L - 1) ? TC_NO : p.kp[l];          // columns per block
            for (int c0 = cg * 16; c0 < nb; c0 += 64) {
                float v0[16], v1[16];
                tc_ld16(tlane + p.col_acc[l] + c0, v0);
                tc_ld16(tlane + p.col_acc[l] + nb + c0, v1);
                if (!lv) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    // l < L-1: acc[o][k] = dW_l[k][o]; last layer: acc[k][o]
                    const bool ok = (l < L - 1) ? (i < on && c0 + j < in) : (i < in && c0 + j < on);
                    if (!ok) continue;
                    float *q = dst + p.net.w_off[l] + ((l < L - 1) ? (c0 + j) * on + i : i * on + c0 + j);
                    const float val = v0[j] + v1[j];
                    *q = add ? *q + val : val;
                }
            }
            if (l < L - 1 && cg == 1) {     // the ones column behind the two blocks
                float v[16];
                tc_ld16(tlane + p.col_acc[l] + 2 * nb - 8, v);
                if (lv && i < on) { float *q = dst + p.net.b_off[l] + i; *q = add ? *q + v[8] : v[8]; }
            }
        }
        if (cg == 2 && quarter == 0) {      // last layer: row 0 of the ones product = column sums of dZ (two blocks)
            float v0[16], v1[16];
            tc_ld16(tlane + p.col_rb, v0);
            tc_ld16(tlane + p.col_rb + TC_NO, v1);
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float sv = v0[j] + v1[j];
                    if (j < n_out) { float *q = dst + p.net.b_off[L - 1] + j; *q = add ? *q + sv : sv; }
                    if (p.net.logstd_off >= 0 && j >= 8 && j < 8 + n_out) {
                        float *q = dst + p.net.logstd_off + j - 8;
                        *q = !gauss ? 0.0f : (add ? *q + sv : sv);
                    }
                }
            }
        }
        }
        flushed = true;
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    };

    if (!worker) {
        // ---- issuer warp: the same barrier sequence as the workers, MMAs instead of epilogues -----------------------
        int ord = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ord) {
            tc_publish();                                               // input tile staged
            for (int l = 0; l < L; ++l) {
                if (tc_elect_one()) {
                    tc_rows_product(tm + p.col_za, sbase + p.off_act[l] + 2 * p.act_plane[l], p.act_plane[l],
                                    tc_wf_k(sbase + p.off_wf[l], p.np[l]), 0, (uint32_t)p.np[l], p.np[l], p.kp[l] / 16);
                    tc_commit(mbar);
                }
                __syncwarp();
                tc_publish();                                           // A_{l+1} (l < L-1) or dZ_{L-1} (head) published
            }
            if (!p.with_grad) continue;
            for (int l = L - 1; l >= 0; --l) {
                if (tc_elect_one()) {
                    if (l == L - 1) {
                        // dA_l = dZ_l W_l^T (K = 16 outputs), dW_l (+)= A_l^T dZ_l (direct form), db_l / dlogstd (+)= 1^T dZ_l
                        tc_rows_product(tm + p.col_za, sbase + p.off_dzl + 2 * (2 * TC_CHUNK_BYTES), 2 * TC_CHUNK_BYTES,
                                        tc_wb_mn(sbase + p.off_wb[l], p.np[l]), 1, (uint32_t)p.wb_plane[l] >> 4, p.kp[l], 1);
                        tc_commit(mbar);                 // dA alone: the epilogue's arithmetic runs under the dW products
                        tc_feat_product(tm + p.col_acc[l], sbase + p.off_act[l], p.act_plane[l], sbase + p.off_dzl,
                                        2 * TC_CHUNK_BYTES, TC_NO, 0, dw_ksteps, dw_started);
                        const TcView ones = tc_feat_mn(sbase + p.off_act[l] + 3 * p.act_plane[l]);
                        const TcView dzl = tc_feat_mn(sbase + p.off_dzl);
                        const uint32_t dzp16 = (2u * TC_CHUNK_BYTES) >> 4;
                        uint32_t alo = ones.lo, blo = dzl.lo;
#pragma unroll 1
                        for (int k = 0; k < dw_ksteps; ++k) {    // rows 0-7 of the result: column sums of [dZ_1|dZ_0] (+ dZ_2)
                            tc_mma(tm + p.col_rb, alo, ones.hi, blo + dzp16, dzl.hi, tc_idesc(64, 2 * TC_NO, 1, 1), (dw_started || k > 0) ? 1u : 0u);
                            tc_mma(tm + p.col_rb, alo, ones.hi, blo, dzl.hi, tc_idesc(64, TC_NO, 1, 1), 1u);
                            alo += ones.k16;
                            blo += dzl.k16;
                        }
                    } else {
                        // dZ_l sits in act[l+1].  dA_l = dZ_l W_l^T (l > 0); dW_l^T, db_l (+)= dZ_l^T [A_l | 1]
                        if (l > 0) {
                            tc_rows_product(tm + p.col_za, sbase + p.off_act[l + 1] + 2 * p.act_plane[l + 1], p.act_plane[l + 1],
                                            tc_wb_mn(sbase + p.off_wb[l], p.np[l]), 1, (uint32_t)p.wb_plane[l] >> 4, p.kp[l],
                                            p.np[l] / 16);
                            tc_commit(mbar);
                        }
                        tc_feat_product(tm + p.col_acc[l], sbase + p.off_act[l + 1], p.act_plane[l + 1], sbase + p.off_act[l],
                                        p.act_plane[l], p.kp[l], 8, dw_ksteps, dw_started);
                    }
                    tc_commit(l == 0 ? mbar : mbar_dw);  // l == 0 has no dA: its dW product is the phase
                }
                __syncwarp();
                if (l == 0) break;
                tc_publish();                                           // dZ_{l-1} published
            }
            dw_started = true;
            if ((ord + 1) % TC_FLUSH_TILES == 0 && tile + gridDim.x < ntiles) {
                flush_grad();                                           // barrier only
                dw_started = false;
            }
        }
        if (p.with_grad && dw_started) flush_grad();
    } else {
    int ord = 0;                                        // ordinal of the tile within this CTA
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ord) {
        const int64_t row0 = tile * tr;
        const int64_t gr = row0 + row;
        const bool rvalid = row < tr && gr < n_rows;
        if (pending) { tc_wait(mbar, parity); parity ^= 1; pending = false; }
        TC_MARK(1);
        // ---- stage the input tile ---------------------------------------------------------------
        store_x();
        tc_publish();
        TC_MARK(2);
        // ---- forward ------------------------------------------------------------------------------
        for (int l = 0; l < L; ++l) {
            if (l == L - 1) {
                // gathers for the NEXT tile (rows, head inputs) and the source rows of the tile after it: issued here so
                // that they have the last product and the whole head phase to arrive -- the next publish (a MEMBAR)
                // would otherwise sit out the HBM latency
                fetch_head_inputs((ord + 1) % 3);
                load_x((ord + 1) % 3);
                fetch_idx(tile + 2 * (int64_t)gridDim.x);
            }
            TC_MARK(4);
            tc_wait(mbar, parity); parity ^= 1;
            TC_MARK(5 + l);
            if (l < L - 1) {
                // A_{l+1} = act(Z_l + b_l), Z_l = the two column blocks: fp32 copy to TMEM for the backward pass,
                // three bf16 planes to shared memory
                float v[16], v2[16];
                tc_ld16(tlane + p.col_za + cg * 16, v);
                tc_ld16(tlane + p.col_za + TC_HID + cg * 16, v2);
                const float4 *b4 = reinterpret_cast<const float4 *>(sBias + l * 64 + cg * 16);
                const float leak = p.net.leak[l];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 bb = b4[j];
                    v[4 * j + 0] = tc_act<HACT>((v[4 * j + 0] + v2[4 * j + 0]) + bb.x, leak);
                    v[4 * j + 1] = tc_act<HACT>((v[4 * j + 1] + v2[4 * j + 1]) + bb.y, leak);
                    v[4 * j + 2] = tc_act<HACT>((v[4 * j + 2] + v2[4 * j + 2]) + bb.z, leak);
                    v[4 * j + 3] = tc_act<HACT>((v[4 * j + 3] + v2[4 * j + 3]) + bb.w, leak);
                }
                if (p.with_grad) tc_st16(tlane + p.col_save + l * TC_HID + cg * 16, v);
                tc_store_chunk(smem + p.off_act[l + 1], p.act_plane[l + 1], cg * 2, row, v);
                tc_store_chunk(smem + p.off_act[l + 1], p.act_plane[l + 1], cg * 2 + 1, row, v + 8);
                tc_publish();
                TC_MARK(8 + l);
            }
        }
        // ---- head: loss-gradient rule on the output row.  All four threads of a row evaluate it (inputs come from
        // shared memory) and each splits / stores its own 4 of the 16 dZ columns.
        {
            float y[TC_NO], dz[TC_NO];
            {
                float y2[TC_NO];
                tc_ld16(tlane + p.col_za, y);
                tc_ld16(tlane + p.col_za + TC_NO, y2);
                const float *b = sBias + (L - 1) * 64;
#pragma unroll
                for (int n = 0; n < TC_NO; ++n) {
                    y[n] = (n < n_out) ? mq_act_apply((y[n] + y2[n]) + b[n], lastact, lastleak) : 0.0f;
                    dz[n] = 0.0f;
                }
            }
            float tgt[TC_NO];
            {
                const float4 *h4 = reinterpret_cast<const float4 *>(sHin + row * TC_HIN);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = h4[q];
                    tgt[4 * q] = t.x; tgt[4 * q + 1] = t.y; tgt[4 * q + 2] = t.z; tgt[4 * q + 3] = t.w;
                }
            }
            const float h_adv = sHin[row * TC_HIN + 16], h_lpold = sHin[row * TC_HIN + 17];
            if (rvalid && out && cg == 1) {
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
                } else if (mq_head_discrete(head.kind)) {
                    // categorical log-prob (distrib.py:24-32): the target row is the one-hot action
                    float mx = y[0];
#pragma unroll
                    for (int n = 1; n < TC_NO; ++n) if (n < n_out) mx = fmaxf(mx, y[n]);
                    float sum = 0.0f;
#pragma unroll
                    for (int n = 0; n < TC_NO; ++n) if (n < n_out) sum += expf(y[n] - mx);
                    const float lse = mx + logf(sum);
                    float lp = 0.0f;
#pragma unroll
                    for (int n = 0; n < TC_NO; ++n) if (n < n_out) lp += tgt[n] * (y[n] - lse);
                    if (logp_out && cg == 2) logp_out[gr] = lp;
                    float w;
                    if (head.kind == MQ_HEAD_DISCRETE_PPO) {
                        const float ratio = expf(lp - h_lpold);
                        const bool kill = (ratio > head.p1 && h_adv > 0.0f) || (ratio < head.p0 && h_adv < 0.0f);
                        w = kill ? 0.0f : ratio * h_adv;
                    } else {
                        w = h_adv;
                    }
#pragma unroll
                    for (int n = 0; n < TC_NO; ++n)
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
                    if (logp_out && cg == 2) logp_out[gr] = lp;
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
            if (p.with_grad) {
                float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (cg == q) { s0 = dz[4 * q]; s1 = dz[4 * q + 1]; s2 = dz[4 * q + 2]; s3 = dz[4 * q + 3]; }
                uint32_t a0, a1, a2, b0, b1, b2;
                tc_split_pair(s0, s1, a0, a1, a2);
                tc_split_pair(s2, s3, b0, b1, b2);
                unsigned char *d = tc_chunk_ptr(smem + p.off_dzl, 2 * TC_CHUNK_BYTES, cg >> 1, row) + (cg & 1) * 8;
                *reinterpret_cast<uint2 *>(d) = make_uint2(a0, b0);
                *reinterpret_cast<uint2 *>(d - 2 * TC_CHUNK_BYTES) = make_uint2(a1, b1);
                *reinterpret_cast<uint2 *>(d - 4 * TC_CHUNK_BYTES) = make_uint2(a2, b2);
            }
        }
        tc_publish();
        commit_head_inputs();               // the head phase is over: its inputs can be replaced by the next tile's
        commit_idx((ord + 2) % 3);
        TC_MARK(11);
        if (!p.with_grad) continue;
        // ---- backward -----------------------------------------------------------------------------
        for (int l = L - 1; l >= 0; --l) {
            TC_MARK(12);
            if (l == 0) { pending = true; break; }
            tc_wait(mbar, parity); parity ^= 1;
            TC_MARK(13 + l);
            // dZ_{l-1} = dA_l * act'(A_l), written over A_l (the fp32 A_l was saved in TMEM by the forward epilogue).
            // Only dA_l has been waited for: the arithmetic and the split into planes run while the tensor pipe still
            // works on dW_l / db_l, which READ A_l -- so the stores wait for the second barrier.
            float a[16], g[16], g2[16];
            tc_ld16(tlane + p.col_za + cg * 16, g);
            tc_ld16(tlane + p.col_za + TC_HID + cg * 16, g2);
            tc_ld16(tlane + p.col_save + (l - 1) * TC_HID + cg * 16, a);
            const float leak = p.net.leak[l - 1];
            uint32_t w[2][3][4];
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] = mq_act_grad(a[j], g[j] + g2[j], HACT, leak);
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int j = 0; j < 4; ++j) tc_split_pair(g[8 * c + 2 * j], g[8 * c + 2 * j + 1], w[c][0][j], w[c][1][j], w[c][2][j]);
            tc_wait(mbar_dw, parity_dw); parity_dw ^= 1;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                unsigned char *d = tc_chunk_ptr(smem + p.off_act[l], p.act_plane[l], cg * 2 + c, row);
#pragma unroll
                for (int pl = 0; pl < 3; ++pl)
                    *reinterpret_cast<uint4 *>(d - pl * p.act_plane[l]) = make_uint4(w[c][pl][0], w[c][pl][1], w[c][pl][2], w[c][pl][3]);
            }
            tc_publish();
            TC_MARK(16 + l);
        }
        dw_started = true;
        if ((ord + 1) % TC_FLUSH_TILES == 0 && tile + gridDim.x < ntiles) {
            if (pending) { tc_wait(mbar, parity); parity ^= 1; pending = false; }
            flush_grad();
            dw_started = false;
        }
    }
    if (pending) { tc_wait(mbar, parity); parity ^= 1; }
    TC_MARK(19);

    // ---- per-CTA partial gradient: TMEM accumulators -> workspace -----------------------------------
    if (p.with_grad) {
        if (dw_started) flush_grad();
        else if (!flushed) {
            float *dst = partial + (int64_t)blockIdx.x * p.net.n_params;
            for (int j = stid; j < p.net.n_params; j += TC_NT) dst[j] = 0.0f;
        }
    }
    }   // workers
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    TC_MARK(20);
#ifdef MQ_TC_PROFILE
    if (prof != nullptr && tid == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); prof[65 + 2 * blockIdx.x] = (long long)gt; }
#endif
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}

// ---- evolution strategies: population evaluation on the tensor pipe ----------------------------------------
// benchmarks/mutation.py:33-37 in population form: member p evaluates policy(theta + noise[p]) on the shared
// observation batch and returns sum_t sum_a action[t][a] * ret_coef[a].  One CTA walks members; per member it builds the
// bf16 plane images of theta + noise[p] (forward image only), pushes the observation tiles through the same three-plane
// products and epilogues as tc_grad_kernel, and reduces the return on chip in a fixed order.
template <int K0, int HACT>
__global__ void __launch_bounds__(TC_NT, 1)
tc_es_kernel(TcPlan p, const float *__restrict__ theta, const float *__restrict__ noise, int64_t n_members,
             const float *__restrict__ obs, int64_t n_obs, const float *__restrict__ ret_coef, float *__restrict__ returns) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_mbar;
    __shared__ float s_red[TC_NT / 32];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int L = p.L;
    const int n_in = p.net.n_in, n_out = p.net.n_out;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"(128u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&s_mbar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = tid; i < p.off_act[1] / 16; i += TC_NT) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    for (int i = p.off_wf[0] / 16 + tid; i < p.smem_bytes / 16; i += TC_NT) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    float *sBias = reinterpret_cast<float *>(smem + p.off_bias);
    tc_publish();
    const uint32_t tm = __shfl_sync(0xffffffffu, s_tmem, 0);
    const uint32_t mbar = tc_smem_u32(&s_mbar);
    const uint32_t sbase = tc_smem_u32(smem);
    uint32_t parity = 0;
    const int quarter = warp & 3, cg = warp >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t tlane = tm + ((uint32_t)(quarter * 32) << 16);
    const int64_t ntiles = (n_obs + TC_ROWS - 1) / TC_ROWS;
    const int lastact = p.net.act[L - 1];
    const float lastleak = p.net.leak[L - 1];
    float coef[TC_NO];
#pragma unroll
    for (int a = 0; a < TC_NO; ++a) coef[a] = a < n_out ? __ldg(ret_coef + a) : 0.0f;

    constexpr int XE = K0 / 4;
    float xr[XE];
    int xrf[XE];
    {
        int r = tid / n_in, f = tid - r * n_in;
#pragma unroll
        for (int j = 0; j < XE; ++j) {
            xrf[j] = r < TC_ROWS ? ((r << 8) | f) : -1;
            r += p.x_step_q; f += p.x_step_r;
            if (f >= n_in) { f -= n_in; ++r; }
        }
    }
    auto load_x = [&](int64_t tile) {
#pragma unroll
        for (int j = 0; j < XE; ++j) {
            float v = 0.0f;
            if (xrf[j] >= 0) {
                const int64_t gr = tile * TC_ROWS + (xrf[j] >> 8);
                if (gr < n_obs) v = __ldg(obs + gr * n_in + (xrf[j] & 255));
            }
            xr[j] = v;
        }
    };
    auto store_x = [&]() {
#pragma unroll
        for (int j = 0; j < XE; ++j) {
            if (xrf[j] < 0) continue;
            const int r = xrf[j] >> 8, f = xrf[j] & 255;
            uint32_t w0, w1, w2;
            tc_split_pair(xr[j], 0.0f, w0, w1, w2);
            unsigned char *d = tc_chunk_ptr(smem + p.off_act[0], p.act_plane[0], f >> 3, r) + (f & 7) * 2;
            *reinterpret_cast<uint16_t *>(d) = (uint16_t)(w0 & 0xFFFFu);
            *reinterpret_cast<uint16_t *>(d - p.act_plane[0]) = (uint16_t)(w1 & 0xFFFFu);
            *reinterpret_cast<uint16_t *>(d - 2 * p.act_plane[0]) = (uint16_t)(w2 & 0xFFFFu);
        }
    };
    int items[TC_MAXL + 1];
    items[0] = 0;
    for (int l = 0; l < L; ++l) items[l + 1] = items[l] + (p.kp[l] / 8) * p.net.out[l];

    for (int64_t member = blockIdx.x; member < n_members; member += gridDim.x) {
        const float *nz = noise + member * p.net.n_params;
        // ---- forward images of theta + noise[member] (all previous MMAs have been waited for) ---------------
#pragma unroll 2
        for (int it = tid; it < items[L]; it += TC_NT) {
            int l = 0;
#pragma unroll
            for (int q = 1; q < TC_MAXL; ++q) if (q < L && it >= items[q]) l = q;
            const int on = p.net.out[l], np = p.np[l], in = p.net.in[l];
            const int local = it - items[l];
            const int kc = local / on, n = local - kc * on;
            const int off = p.net.w_off[l] + (kc * 8) * on + n;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (kc * 8 + j < in) ? __ldg(theta + off + j * on) + __ldg(nz + off + j * on) : 0.0f;
            uint32_t q[3][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) tc_split_pair(v[2 * j], v[2 * j + 1], q[0][j], q[1][j], q[2][j]);
            unsigned char *df = smem + p.off_wf[l] + kc * (3 * np * 16) + n * 16;
#pragma unroll
            for (int pl = 0; pl < 3; ++pl)
                *reinterpret_cast<uint4 *>(df + pl * np * 16) = make_uint4(q[pl][0], q[pl][1], q[pl][2], q[pl][3]);
        }
        for (int l = 0; l < L; ++l)
            for (int n = tid; n < p.net.out[l]; n += TC_NT)
                sBias[l * 64 + n] = __ldg(theta + p.net.b_off[l] + n) + __ldg(nz + p.net.b_off[l] + n);
        load_x(0);
        float acc = 0.0f;
        for (int64_t tile = 0; tile < ntiles; ++tile) {
            store_x();
            tc_publish();                   // also publishes the images / biases on the first tile
            for (int l = 0; l < L; ++l) {
                if (warp == 0) {
                    if (tc_elect_one()) {
                        tc_rows_product(tm, sbase + p.off_act[l] + 2 * p.act_plane[l], p.act_plane[l],
                                        tc_wf_k(sbase + p.off_wf[l], p.np[l]), 0, (uint32_t)p.np[l], p.np[l], p.kp[l] / 16);
                        tc_commit(mbar);
                    }
                    __syncwarp();
                }
                if (l == 0 && tile + 1 < ntiles) load_x(tile + 1);
                tc_wait(mbar, parity); parity ^= 1;
                if (l < L - 1) {
                    float v[16], v2[16];
                    tc_ld16(tlane + cg * 16, v);
                    tc_ld16(tlane + TC_HID + cg * 16, v2);
                    const float4 *b4 = reinterpret_cast<const float4 *>(sBias + l * 64 + cg * 16);
                    const float leak = p.net.leak[l];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 bb = b4[j];
                        v[4 * j + 0] = tc_act<HACT>((v[4 * j + 0] + v2[4 * j + 0]) + bb.x, leak);
                        v[4 * j + 1] = tc_act<HACT>((v[4 * j + 1] + v2[4 * j + 1]) + bb.y, leak);
                        v[4 * j + 2] = tc_act<HACT>((v[4 * j + 2] + v2[4 * j + 2]) + bb.z, leak);
                        v[4 * j + 3] = tc_act<HACT>((v[4 * j + 3] + v2[4 * j + 3]) + bb.w, leak);
                    }
                    tc_store_chunk(smem + p.off_act[l + 1], p.act_plane[l + 1], cg * 2, row, v);
                    tc_store_chunk(smem + p.off_act[l + 1], p.act_plane[l + 1], cg * 2 + 1, row, v + 8);
                    tc_publish();
                }
            }
            if (cg == 0) {                  // actions of this row -> its term of the return
                float y[TC_NO], y2[TC_NO];
                tc_ld16(tlane, y);
                tc_ld16(tlane + TC_NO, y2);
                if (tile * TC_ROWS + row < n_obs) {
                    const float *b = sBias + (L - 1) * 64;
#pragma unroll
                    for (int a = 0; a < TC_NO; ++a)
                        if (a < n_out) acc = fmaf(mq_act_apply((y[a] + y2[a]) + b[a], lastact, lastleak), coef[a], acc);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        }
        // ---- fixed-order block sum -> returns[member]; the barriers also order the next member's image stores
        acc = mq_warp_sum(acc);
        __syncthreads();
        if (lane == 0) s_red[warp] = acc;
        __syncthreads();
        if (tid == 0) {
            float s = 0.0f;
            for (int w = 0; w < TC_NT / 32; ++w) s += s_red[w];
            returns[member] = s;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128u));
}

// ---- host side ------------------------------------------------------------------------------------
static int tc_plan_build(const mq_net_t *net, int64_t batch, int with_grad, const mq_head_t *head, TcPlan *p) {
    memset(p, 0, sizeof(*p));
    p->net = *net;
    const int L = net->n_layers;
    if (L < 2 || L > TC_MAXL) return 0;
    if (net->n_in > 32 || net->n_out > TC_NO) return 0;
    for (int l = 0; l + 1 < L; ++l) {
        if (net->out[l] > TC_HID) return 0;
        if (net->act[l] != net->act[0] || (net->act[l] != MQ_ACT_TANH && net->act[l] != MQ_ACT_LRELU)) return 0;
    }
    const bool gauss = head && mq_head_gauss(head->kind);
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
    p->x_step_q = TC_NT / net->n_in; p->x_step_r = TC_NT % net->n_in;
    p->h_step_q = TC_NT / net->n_out; p->h_step_r = TC_NT % net->n_out;
    int off = 0;
    p->off_dzl = off; off += 3 * 2 * TC_CHUNK_BYTES;
    for (int l = 0; l < L; ++l) {
        p->act_plane[l] = (p->kp[l] / 8) * TC_CHUNK_BYTES;
        p->off_act[l] = off; off += 3 * p->act_plane[l] + TC_CHUNK_BYTES;
    }
    for (int l = 0; l < L; ++l) { p->off_wf[l] = off; off += 3 * p->np[l] * p->kp[l] * 2; }
    for (int l = 1; l < L; ++l) {
        p->wb_plane[l] = p->np[l] * p->kp[l] * 2;
        p->off_wb[l] = off; off += 3 * p->wb_plane[l];
    }
    p->off_bias = off; off += TC_MAXL * 64 * 4;
    p->off_misc = off; off += 32 * 4;
    p->off_hin = off; off += TC_ROWS * TC_HIN * 4;
    p->off_idx = off; off += 3 * TC_ROWS * 4;
    // the ones-row product of the last layer reads 8 chunks from the ones chunk of act[L-1] on: keep that in bounds
    if (off < p->off_act[L - 1] + 3 * p->act_plane[L - 1] + 8 * TC_CHUNK_BYTES)
        off = p->off_act[L - 1] + 3 * p->act_plane[L - 1] + 8 * TC_CHUNK_BYTES;
    p->smem_bytes = (int)mq_round_up(off, 16);
    int col = 0;
    p->col_za = col; col += 2 * TC_HID;
    p->col_save = col; col += (L - 1) * TC_HID;            // fp32 activations of the hidden layers
    for (int l = 0; l < L; ++l) { p->col_acc[l] = col; col += (l == L - 1) ? 2 * TC_NO : 2 * p->kp[l] + 8; }
    p->col_rb = col; col += 2 * TC_NO;
    if (col > 512) return 0;
    if (p->smem_bytes + 1024 > mq_smem_optin()) return 0;
    return 1;
}

long long *mq_phase_prof_ptr();
int mq_tc3_launch(const mq_net_t *net, const int32_t *row_idx, int64_t batch, const mq_head_t *head, const float *theta,
                  int with_grad, float *out, float *logp, void *ws, int64_t ws_bytes, int *grid_out, void *stream);

// Launches a tensor-core kernel when the net fits and the batch is large enough (or head->tc says MQ_TC_FORCE).
// Calls that carry packed records (head->records) go to the record-fed kernel of mq_tc3.cu, which also implements the
// reduced-precision modes; calls with separate arrays keep this file's kernel (exact mode only).
// Returns MQ_ERR_UNSUPPORTED (without an error message side effect that matters) otherwise.
// with_grad: partials go to ws ([grid][n_params]); *grid_out = CTAs launched.
int mq_tc_launch(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                 int64_t batch, const mq_head_t *head, int with_grad, float *out, float *logp,
                 void *ws, int64_t ws_bytes, int *grid_out, void *stream) {
    const int32_t tc = head->tc;
    if (tc & MQ_TC_OFF) return MQ_ERR_UNSUPPORTED;
    if (!(tc & MQ_TC_FORCE) && batch < 8192) return MQ_ERR_UNSUPPORTED;
    if (head->records && with_grad) {
        const int rc3 = mq_tc3_launch(net, row_idx, batch, head, theta, with_grad, out, logp, ws, ws_bytes, grid_out, stream);
        if (rc3 != MQ_ERR_UNSUPPORTED) return rc3;
    }
    const int planes = tc & MQ_TC_PLANES_MASK;
    MQ_REQUIRE(planes == 0 || planes == 3, MQ_ERR_INVALID,
               "mq_tc_launch: the reduced-precision tensor-core modes need packed records (mq_pack_records)");
    TcPlan p;
    if (!tc_plan_build(net, batch, with_grad, head, &p)) return MQ_ERR_UNSUPPORTED;
    const int64_t ntiles = mq_ceil_div(batch, p.tile_rows);
    const int64_t sms = mq_sm_count();
    const int grid = (int)(ntiles < sms ? ntiles : sms);
    if (with_grad)
        MQ_REQUIRE(ws && ws_bytes >= (int64_t)grid * net->n_params * (int64_t)sizeof(float), MQ_ERR_INVALID,
                   "mq_tc_launch: workspace too small");
    cudaError_t e = cudaSuccess;
#define TC_LAUNCH(K0_, ACT_)                                                                                             \
    do {                                                                                                                 \
        e = cudaFuncSetAttribute(tc_grad_kernel<K0_, ACT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes); \
        if (e == cudaSuccess)                                                                                            \
            e = mq_launch_pdl(tc_grad_kernel<K0_, ACT_>, dim3(grid), dim3(TC_NTI), (size_t)p.smem_bytes, stream, p, theta, \
                              x, row_idx, batch, *head, (float *)ws, out, logp, mq_phase_prof_ptr());                    \
    } while (0)
    const int hact = net->act[0];
    if (p.k0 == 16 && hact == MQ_ACT_TANH) TC_LAUNCH(16, MQ_ACT_TANH);
    else if (p.k0 == 16) TC_LAUNCH(16, MQ_ACT_LRELU);
    else if (hact == MQ_ACT_TANH) TC_LAUNCH(32, MQ_ACT_TANH);
    else TC_LAUNCH(32, MQ_ACT_LRELU);
#undef TC_LAUNCH
    if (e != cudaSuccess) {
        mq_set_error("mq_tc_launch: %s", cudaGetErrorString(e));
        return MQ_ERR_CUDA;
    }
    MQ_CHECK_LAUNCH("mq_tc_launch");
    if (grid_out) *grid_out = grid;
    return MQ_OK;
}

// Population evaluation on the tensor pipe (mq_es_eval dispatches here); MQ_ERR_UNSUPPORTED if the net does not fit.
int mq_tc_es_launch(const mq_net_t *net, const float *theta, const float *noise, int64_t n_members, const float *obs,
                    int64_t n_obs, const float *ret_coef, float *returns, void *stream) {
    if (n_obs < 256) return MQ_ERR_UNSUPPORTED;
    if (net->logstd_off >= 0) return MQ_ERR_UNSUPPORTED;
    TcPlan p;
    if (!tc_plan_build(net, n_obs, 0, nullptr, &p)) return MQ_ERR_UNSUPPORTED;
    p.tile_rows = TC_ROWS;
    const int64_t sms = mq_sm_count();
    const int grid = (int)(n_members < sms ? n_members : sms);
    cudaError_t e = cudaSuccess;
#define TC_ES_LAUNCH(K0_, ACT_)                                                                                        \
    do {                                                                                                               \
        e = cudaFuncSetAttribute(tc_es_kernel<K0_, ACT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes); \
        if (e == cudaSuccess)                                                                                          \
            tc_es_kernel<K0_, ACT_><<<grid, TC_NT, p.smem_bytes, (cudaStream_t)stream>>>(p, theta, noise, n_members,  \
                                                                                        obs, n_obs, ret_coef, returns); \
    } while (0)
    const int hact = net->act[0];
    if (p.k0 == 16 && hact == MQ_ACT_TANH) TC_ES_LAUNCH(16, MQ_ACT_TANH);
    else if (p.k0 == 16) TC_ES_LAUNCH(16, MQ_ACT_LRELU);
    else if (hact == MQ_ACT_TANH) TC_ES_LAUNCH(32, MQ_ACT_TANH);
    else TC_ES_LAUNCH(32, MQ_ACT_LRELU);
#undef TC_ES_LAUNCH
    if (e != cudaSuccess) {
        mq_set_error("mq_tc_es_launch: %s", cudaGetErrorString(e));
        return MQ_ERR_CUDA;
    }
    MQ_CHECK_LAUNCH("mq_tc_es_launch");
    return MQ_OK;
}
#else   // MQ_EMU: the host emulation has no tensor cores
int mq_tc_launch(const mq_net_t *, const float *, const float *, const int32_t *, int64_t, const mq_head_t *, int,
                 float *, float *, void *, int64_t, int *, void *) { return MQ_ERR_UNSUPPORTED; }
int mq_tc_es_launch(const mq_net_t *, const float *, const float *, int64_t, const float *, int64_t, const float *, float *,
                    void *) { return MQ_ERR_UNSUPPORTED; }
#endif
