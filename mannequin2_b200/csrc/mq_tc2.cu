// Tensor-core (tcgen05 / TMEM) fused forward + head + backward for small MLPs at LARGE batch, second generation:
// TWO tiles in flight per SM.
//
// Same contract, operand format and arithmetic as mq_tc.cu (reference sequence: Function.evaluate
// mannequin/basicnet.py:192-207, heads benchmarks/gae.py:22 / benchmarks/ppo.py:34-37 / mannequin/distrib.py:24-32,61-68,
// backprop mannequin/basicnet.py:209-257 + mannequin/backprop.py:11-60; fp32 operands as three exact bf16 planes, plane
// pairs i + j <= 2, fp32 accumulation in TMEM).  What changes is the schedule.  A tile's work is a chain of six
// dependent (MMA phase -> epilogue) steps, so with one tile per SM (mq_tc.cu) the tensor pipe idles during epilogues
// and the ALUs idle during MMA phases.  Here a CTA of 512 threads is two independent GROUPS of 8 warps; each group
// walks its own sequence of 64-row tiles (M = 64 MMAs) with its own shared-memory operand buffers, its own issuing
// thread and mbarrier, and its own TMEM columns -- including its own dW / db accumulators, so no MMA of one group
// touches anything of the other and the result stays bit-reproducible.  The groups share the weight images.  While
// one group runs an epilogue the other group's MMAs occupy the tensor pipe.
//
// Budget that makes this fit (the reason for the details below):
//   shared memory  per group 63 chunks of 1 KB (64 rows x 16 B) of activation planes + head inputs ~ 69 KB, x 2, + 66 KB
//                  of weight images = 205 KB (mq_tc.cu: 199 KB for ONE 128-row tile)
//   TMEM           per group 128 columns Z / dA + 128 columns of accumulators = 256; 512 in all.  To fit, (a) the fp32
//                  copy of the activations for the backward pass is NOT kept in TMEM: A_l is rebuilt exactly from its
//                  three planes; (b) the dW accumulators are one column block (six MMAs per K step, N = in + 8) instead
//                  of the wide two-block form.
//   M = 64 TMEM layout: tile row i lives in lane i % 16 of lane quarter i / 16; a warp reads its quarter with the
//   32-lane load and lanes 16..31 take over half of the columns by shuffle, so all 32 threads have epilogue work.
#include "mq_tc.cuh"

#ifndef MQ_EMU

#define HALF_LOG_2PI 0.91893853320467274178f
#define T2_NT 512
#define T2_GNT 256               // threads per group
#define T2_ROWS 64               // MMA M of the row-major products = rows of a tile buffer
#define T2_CB 1024               // one 8-feature chunk of all 64 rows (64 rows x 16 B)
#define T2_HID 64
#define T2_NO 16
#define T2_MAXL 3
#define T2_HIN 20                // floats of head input per tile row: 16 targets, weight / advantage, old log-prob
#define T2_FLUSH_TILES 8         // tiles accumulated in TMEM before the accumulators are folded into memory (truncating adds)
#define T2_GCOLS 256             // TMEM columns per group

struct Tc2Plan {
    mq_net_t net;
    int L, k0;
    int kp[T2_MAXL], np[T2_MAXL];
    int tile_rows;                                  // multiple of 8, <= 64
    int x_step_q, x_step_r, h_step_q, h_step_r;     // T2_GNT / n_in, T2_GNT % n_in, T2_GNT / n_out, T2_GNT % n_out
    int with_grad;
    // byte offsets inside a group's region (group g starts at g * group_bytes); act[l]: planes 2,1,0 then the ones chunk
    int off_dzl, off_act[T2_MAXL], act_plane[T2_MAXL], off_hin, off_idx, group_bytes;
    // shared by both groups
    int off_wf[T2_MAXL], off_wb[T2_MAXL], wb_plane[T2_MAXL], off_bias, off_misc, smem_bytes;
    // TMEM columns relative to the group's base; Z / dA at column 0 (two blocks of 64)
    int col_acc[T2_MAXL], col_rb;
};

// two 32-column loads (this thread's TMEM lane) and ONE wait
__device__ __forceinline__ void t2_ld32x2(uint32_t ta, uint32_t tb, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%64];\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%65];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(ta), "r"(tb) : "memory");
}

__device__ __forceinline__ unsigned char *t2_chunk_ptr(unsigned char *buf, int plane_bytes, int chunk, int row) {
    return buf + 2 * plane_bytes + chunk * T2_CB + (row >> 3) * 128 + (row & 7) * 16;
}
__device__ __forceinline__ void t2_store_chunk(unsigned char *buf, int plane_bytes, int chunk, int row, const float *v) {
    uint32_t w[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tc_split_pair(v[2 * j], v[2 * j + 1], w[0][j], w[1][j], w[2][j]);
    unsigned char *p = t2_chunk_ptr(buf, plane_bytes, chunk, row);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
        *reinterpret_cast<uint4 *>(p - pl * plane_bytes) = make_uint4(w[pl][0], w[pl][1], w[pl][2], w[pl][3]);
}
// the exact fp32 values back from the three planes
__device__ __forceinline__ void t2_load_chunk(const unsigned char *buf, int plane_bytes, int chunk, int row, float *v) {
    const unsigned char *p = t2_chunk_ptr(const_cast<unsigned char *>(buf), plane_bytes, chunk, row);
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

// activation plane at `addr` viewed K-major (M = rows, K = features): LBO = chunk stride, SBO = row-group stride
__device__ __forceinline__ TcView t2_rows_k(uint32_t addr) {
    return TcView{tc_desc_lo(addr, T2_CB), tc_desc_hi(128u), (2u * T2_CB) >> 4};
}
// activation buffer from `addr` on viewed MN-major (M or N = features; K = rows)
__device__ __forceinline__ TcView t2_feat_mn(uint32_t addr) {
    return TcView{tc_desc_lo(addr, 128u), tc_desc_hi(T2_CB), 256u >> 4};
}

// D[64 x (two blocks of nb)] = sum over plane pairs of A_i B_j (see tc_rows_product in mq_tc.cu), M = 64
__device__ __forceinline__ void t2_rows_product(uint32_t d_tmem, uint32_t a0_addr, uint32_t a_plane, const TcView &b,
                                                int b_mn, uint32_t b_plane16, int nb, int ksteps) {
    const TcView a0 = t2_rows_k(a0_addr);
    const uint32_t ap16 = a_plane >> 4;
    const uint32_t id2 = tc_idesc(T2_ROWS, 2 * nb, 0, b_mn);
    const uint32_t id1 = tc_idesc(T2_ROWS, nb, 0, b_mn);
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

// acc[64 x (nb + extra)] (+)= sum over plane pairs of Da_i^T Db_j over the rows of the tile (K = rows, 16 per step):
// Da's features are M, Db's features (+ `extra` columns stored behind plane 0: the ones chunk) are N.  ONE column
// block: six instructions per K step.
__device__ __forceinline__ void t2_feat_product(uint32_t acc_tmem, uint32_t da_base, uint32_t da_plane, uint32_t db_base,
                                                uint32_t db_plane, int nb, int extra, int ksteps, bool accumulate) {
    const TcView a = t2_feat_mn(da_base), b = t2_feat_mn(db_base);
    const uint32_t ap16 = da_plane >> 4, bp16 = db_plane >> 4;
    const uint32_t id_e = tc_idesc(64, nb + extra, 1, 1), id_n = tc_idesc(64, nb, 1, 1);
    uint32_t alo = a.lo, blo = b.lo;
#pragma unroll 1
    for (int k = 0; k < ksteps; ++k) {
        tc_mma(acc_tmem, alo + 2 * ap16, a.hi, blo + 2 * bp16, b.hi, id_e, (accumulate || k > 0) ? 1u : 0u);   // Da_0 x [Db_0|e]
        tc_mma(acc_tmem, alo + ap16, a.hi, blo + 2 * bp16, b.hi, id_e, 1u);                                    // Da_1 x [Db_0|e]
        tc_mma(acc_tmem, alo, a.hi, blo + 2 * bp16, b.hi, id_e, 1u);                                           // Da_2 x [Db_0|e]
        tc_mma(acc_tmem, alo + 2 * ap16, a.hi, blo + bp16, b.hi, id_n, 1u);                                    // Da_0 x Db_1
        tc_mma(acc_tmem, alo + ap16, a.hi, blo + bp16, b.hi, id_n, 1u);                                        // Da_1 x Db_1
        tc_mma(acc_tmem, alo + 2 * ap16, a.hi, blo, b.hi, id_n, 1u);                                           // Da_0 x Db_2
        alo += a.k16;
        blo += b.k16;
    }
}

// one group: make generic-proxy smem writes visible to the tensor core, order TMEM accesses, meet (named barrier 1 + g)
__device__ __forceinline__ void t2_publish(int g) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(T2_GNT) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

template <int K0, int HACT>
__global__ void __launch_bounds__(T2_NT, 1)
tc2_grad_kernel(Tc2Plan p, const float *__restrict__ theta, const float *__restrict__ x,
                const int32_t *__restrict__ idx, int64_t n_rows, mq_head_t head,
                float *__restrict__ partial, float *__restrict__ out, float *__restrict__ logp_out, long long *prof) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_mbar[2];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);       // warp-uniform for the compiler too
    const int g = warp >> 3, gw = warp & 7, gtid = tid & (T2_GNT - 1);
    const int L = p.L;
    const int n_in = p.net.n_in, n_out = p.net.n_out;
    // developer aid (compile with -DMQ_TC_PROFILE): per-phase cycles of both groups of CTA 0, every CTA's start / end time
#ifdef MQ_TC_PROFILE
    long long t_last = 0;
    const bool profiling = prof != nullptr && blockIdx.x == 0 && gtid == 0;
    if (profiling) t_last = clock64();
    if (prof != nullptr && tid == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); prof[64 + 2 * blockIdx.x] = (long long)gt; }
#define T2_MARK(i) do { if (profiling) { long long now_ = clock64(); prof[g * 32 + (i)] += now_ - t_last; t_last = now_; } } while (0)
#else
    (void)prof;
#define T2_MARK(i) do { } while (0)
#endif

    // ---- one-time setup by the whole CTA: TMEM, barriers, operand images of the weights ---------------------------
    // (programmatic dependent launch: everything up to mq_pdl_wait() overlaps with the tail of the previous kernel)
    mq_pdl_trigger();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&s_mbar[0])), "r"(1u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&s_mbar[1])), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    for (int i = tid; i < p.smem_bytes / 16; i += T2_NT) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    {
        __nv_bfloat16 one = __float2bfloat16_rn(1.0f);
        const uint32_t ob = *reinterpret_cast<uint16_t *>(&one);
        for (int gg = 0; gg < 2; ++gg)
            for (int l = 0; l < L; ++l) {
                uint32_t *ones = reinterpret_cast<uint32_t *>(smem + gg * p.group_bytes + p.off_act[l] + 3 * p.act_plane[l]);
                for (int i = tid; i < T2_CB / 4; i += T2_NT) ones[i] = ob | (ob << 16);
            }
    }
    float *sBias = reinterpret_cast<float *>(smem + p.off_bias);      // [L][64]
    float *sMisc = reinterpret_cast<float *>(smem + p.off_misc);      // logstd[16], exp(-logstd)[16]
    mq_pdl_wait();                        // first global read below: the predecessor's writes are complete and visible
    {   // weight images: one work item = 8 consecutive inputs (one 16-byte chunk) of one output column
        int items[T2_MAXL + 1];
        items[0] = 0;
        for (int l = 0; l < L; ++l) items[l + 1] = items[l] + (p.kp[l] / 8) * p.net.out[l];
#pragma unroll 2
        for (int it = tid; it < items[L]; it += T2_NT) {
            int l = 0;
#pragma unroll
            for (int q = 1; q < T2_MAXL; ++q) if (q < L && it >= items[q]) l = q;
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
            for (int n = tid; n < p.net.out[l]; n += T2_NT) sBias[l * 64 + n] = __ldg(theta + p.net.b_off[l] + n);
        if (p.net.logstd_off >= 0 && tid < n_out) {
            const float ls = __ldg(theta + p.net.logstd_off + tid);
            sMisc[tid] = ls;
            sMisc[16 + tid] = expf(-ls);
        }
    }

    // ---- from here on the two groups run independently ---------------------------------------------------------
    unsigned char *gs = smem + g * p.group_bytes;
    float *sHin = reinterpret_cast<float *>(gs + p.off_hin);          // head inputs of the tile: [64][T2_HIN]
    int *sIdx = reinterpret_cast<int *>(gs + p.off_idx);              // source rows of three consecutive tiles: [3][64]
    const int q = warp & 3;                             // TMEM lane quarter this warp may touch
    const int half = gw >> 2;                           // which 32 of the 64 columns this warp reads
    const int sub = lane & 15, hi = lane >> 4;
    const int row = q * 16 + sub;                       // tile row of this thread (M = 64: lane sub of quarter q)
    const int cg = half * 2 + hi;                       // which 16 of the 64 columns this thread owns
    const int tr = p.tile_rows;
    const int64_t ntiles = (n_rows + tr - 1) / tr;
    const int dw_ksteps = (tr + 15) / 16;
    const bool gauss = mq_head_gauss(head.kind), prob = mq_head_prob(head.kind);
    const int lastact = p.net.act[L - 1];
    const float lastleak = p.net.leak[L - 1];
    const int64_t vt0 = (int64_t)blockIdx.x * 2 + g, vstride = (int64_t)gridDim.x * 2;

    // ---- gathers (lanes along the features of a row; source rows staged two tiles ahead; see mq_tc.cu) ----
    constexpr int XE = K0 / 4;                          // elements of the 64 x n_in input tile per thread
    constexpr int HE = T2_NO * T2_ROWS / T2_GNT;        // elements of the 64 x n_out target tile per thread
    float xr[XE];
    int xrf[XE], hrf[HE];                               // (row << 8 | feature) of this thread's elements, -1: none
    {
        int r = gtid / n_in, f = gtid - r * n_in;
#pragma unroll
        for (int j = 0; j < XE; ++j) {
            xrf[j] = r < T2_ROWS ? ((r << 8) | f) : -1;
            r += p.x_step_q; f += p.x_step_r;
            if (f >= n_in) { f -= n_in; ++r; }
        }
        r = gtid / n_out; f = gtid - r * n_out;
#pragma unroll
        for (int j = 0; j < HE; ++j) {
            hrf[j] = r < T2_ROWS ? ((r << 8) | f) : -1;
            r += p.h_step_q; f += p.h_step_r;
            if (f >= n_out) { f -= n_out; ++r; }
        }
    }
    int pend_idx = -1;
    float pend_t[HE], pend_a = 0.0f, pend_l = 0.0f;
    auto fetch_idx = [&](int64_t tile) {                // threads 0..63 of the group: source row (or -1) of every tile row
        pend_idx = -1;
        if (gtid < T2_ROWS) {
            const int64_t gr = tile * tr + gtid;
            if (tile < ntiles && gtid < tr && gr < n_rows) pend_idx = idx ? tc_ldg_s32(idx + gr) : (int)gr;
        }
    };
    auto commit_idx = [&](int slot) { if (gtid < T2_ROWS) sIdx[slot * T2_ROWS + gtid] = pend_idx; };
    auto load_x = [&](int slot) {
#pragma unroll
        for (int j = 0; j < XE; ++j) {
            float v = 0.0f;
            if (xrf[j] >= 0) {
                const int src = sIdx[slot * T2_ROWS + (xrf[j] >> 8)];
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
            unsigned char *d = t2_chunk_ptr(gs + p.off_act[0], p.act_plane[0], f >> 3, r) + (f & 7) * 2;
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
                const int src = sIdx[slot * T2_ROWS + (hrf[j] >> 8)];
                if (src >= 0) v = tc_ldg_f32(mat + (int64_t)src * n_out + (hrf[j] & 255));
            }
            pend_t[j] = v;
        }
        pend_a = 0.0f; pend_l = 0.0f;
        if (prob && gtid < T2_ROWS) {
            const int src = sIdx[slot * T2_ROWS + gtid];
            if (src >= 0) {
                if (mq_head_ppo(head.kind)) { pend_a = tc_ldg_f32(head.adv + src); pend_l = tc_ldg_f32(head.logp_old + src); }
                else if (head.dy) pend_a = tc_ldg_f32(head.dy + src);
            }
        }
    };
    auto commit_head_inputs = [&]() {
#pragma unroll
        for (int j = 0; j < HE; ++j)
            if (hrf[j] >= 0) sHin[(hrf[j] >> 8) * T2_HIN + (hrf[j] & 255)] = pend_t[j];
        if (prob && gtid < T2_ROWS) { sHin[gtid * T2_HIN + 16] = pend_a; sHin[gtid * T2_HIN + 17] = pend_l; }
    };
    fetch_idx(vt0); commit_idx(0);
    fetch_idx(vt0 + vstride); commit_idx(1);
    tc_publish();                                       // whole CTA: weight images, TMEM address, barriers
    const uint32_t tm = __shfl_sync(0xffffffffu, s_tmem, 0) + (uint32_t)(g * T2_GCOLS);
    const uint32_t mbar = tc_smem_u32(&s_mbar[g]);
    const uint32_t sb = tc_smem_u32(gs);                // this group's operand buffers
    const uint32_t sw = tc_smem_u32(smem);              // shared weight images
    const uint32_t tlane = tm + ((uint32_t)(q * 32) << 16);
    uint32_t parity = 0;
    bool pending = false;                               // a committed MMA batch nobody waited for yet
    bool dw_started = false;                            // dW / db accumulators hold data
    load_x(0);
    fetch_head_inputs(0);
    commit_head_inputs();
    T2_MARK(0);

    // this thread's 16 columns of the two-block result at TMEM column 0: the warp reads 32 columns of both blocks for its
    // 32 lanes (lanes 16..31 of a quarter hold no tile row), lanes 16..31 take columns 16..31 of the row of lane - 16
    auto load_blocks = [&](float *v) {
        uint32_t r[64];
        t2_ld32x2(tlane + half * 32, tlane + T2_HID + half * 32, r);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float lo = __uint_as_float(r[j]) + __uint_as_float(r[32 + j]);
            const float up = __uint_as_float(r[16 + j]) + __uint_as_float(r[48 + j]);
            const float o = __shfl_sync(0xffffffffu, up, sub);
            v[j] = hi ? o : lo;
        }
    };

    // TMEM accumulators -> this group's partial gradient in the workspace (set on the first flush, added afterwards)
    bool flushed = false;
    auto flush_grad = [&]() {
        float *dst = partial + (vt0) * p.net.n_params;
        const bool add = flushed;
        const int i = q * 16 + lane;                    // accumulator row of this lane (lanes 0..15 of the quarter)
        const bool lv = lane < 16;
        for (int l = 0; l < L; ++l) {
            const int in = p.net.in[l], on = p.net.out[l];
            const int nb = (l == L - 1) ? T2_NO : p.kp[l];
            for (int c0 = half * 16; c0 < nb; c0 += 32) {
                float v0[16];
                tc_ld16(tlane + p.col_acc[l] + c0, v0);
                if (!lv) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    // l < L-1: acc[o][k] = dW_l[k][o]; last layer: acc[k][o]
                    const bool ok = (l < L - 1) ? (i < on && c0 + j < in) : (i < in && c0 + j < on);
                    if (!ok) continue;
                    float *d = dst + p.net.w_off[l] + ((l < L - 1) ? (c0 + j) * on + i : i * on + c0 + j);
                    *d = add ? *d + v0[j] : v0[j];
                }
            }
            if (l < L - 1 && half == 1) {               // the ones column behind the block
                float v[16];
                tc_ld16(tlane + p.col_acc[l] + nb - 8, v);
                if (lv && i < on) { float *d = dst + p.net.b_off[l] + i; *d = add ? *d + v[8] : v[8]; }
            }
        }
        if (gw == 0) {                                  // last layer: row 0 of the ones product = column sums of dZ
            float v0[16];
            tc_ld16(tlane + p.col_rb, v0);
            if (lane == 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (j < n_out) { float *d = dst + p.net.b_off[L - 1] + j; *d = add ? *d + v0[j] : v0[j]; }
                    if (p.net.logstd_off >= 0 && j >= 8 && j < 8 + n_out) {
                        float *d = dst + p.net.logstd_off + j - 8;
                        *d = !gauss ? 0.0f : (add ? *d + v0[j] : v0[j]);
                    }
                }
            }
        }
        flushed = true;
        t2_publish(g);
    };

    int ord = 0;                                        // ordinal of the tile within this group
    for (int64_t tile = vt0; tile < ntiles; tile += vstride, ++ord) {
        const int64_t gr = tile * tr + row;
        const bool rvalid = row < tr && gr < n_rows;
        if (pending) { tc_wait(mbar, parity); parity ^= 1; pending = false; }
        T2_MARK(1);
        // ---- stage the input tile ---------------------------------------------------------------
        store_x();
        t2_publish(g);
        T2_MARK(2);
        // ---- forward ------------------------------------------------------------------------------
        for (int l = 0; l < L; ++l) {
            if (gw == 0) {
                if (tc_elect_one()) {
                    t2_rows_product(tm, sb + p.off_act[l] + 2 * p.act_plane[l], p.act_plane[l],
                                    tc_wf_k(sw + p.off_wf[l], p.np[l]), 0, (uint32_t)p.np[l], p.np[l], p.kp[l] / 16);
                    tc_commit(mbar);
                }
                __syncwarp();
            }
            T2_MARK(3);
            if (l == L - 1) {
                // gathers for the NEXT tile and the source rows of the tile after it: they have the last product and
                // the whole head phase to arrive
                fetch_head_inputs((ord + 1) % 3);
                load_x((ord + 1) % 3);
                fetch_idx(tile + 2 * vstride);
            }
            T2_MARK(4);
            tc_wait(mbar, parity); parity ^= 1;
            T2_MARK(5 + l);
            if (l < L - 1) {
                // A_{l+1} = act(Z_l + b_l) as three bf16 planes in shared memory
                float v[16];
                load_blocks(v);
                const float4 *b4 = reinterpret_cast<const float4 *>(sBias + l * 64 + cg * 16);
                const float leak = p.net.leak[l];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 bb = b4[j];
                    v[4 * j + 0] = tc_act<HACT>(v[4 * j + 0] + bb.x, leak);
                    v[4 * j + 1] = tc_act<HACT>(v[4 * j + 1] + bb.y, leak);
                    v[4 * j + 2] = tc_act<HACT>(v[4 * j + 2] + bb.z, leak);
                    v[4 * j + 3] = tc_act<HACT>(v[4 * j + 3] + bb.w, leak);
                }
                t2_store_chunk(gs + p.off_act[l + 1], p.act_plane[l + 1], cg * 2, row, v);
                t2_store_chunk(gs + p.off_act[l + 1], p.act_plane[l + 1], cg * 2 + 1, row, v + 8);
                t2_publish(g);
                T2_MARK(8 + l);
            }
        }
        // ---- head: loss-gradient rule on the output row.  All four threads of a row evaluate it (inputs come from
        // shared memory) and each splits / stores its own 4 of the 16 dZ columns.
        {
            float y[T2_NO], dz[T2_NO];
            {
                float y2[T2_NO];
                tc_ld16(tlane, y);
                tc_ld16(tlane + T2_NO, y2);
                const float *b = sBias + (L - 1) * 64;
#pragma unroll
                for (int n = 0; n < T2_NO; ++n) {
                    const float z = __shfl_sync(0xffffffffu, y[n] + y2[n], sub);       // the row lives in lane sub
                    y[n] = (n < n_out) ? mq_act_apply(z + b[n], lastact, lastleak) : 0.0f;
                    dz[n] = 0.0f;
                }
            }
            float tgt[T2_NO];
            {
                const float4 *h4 = reinterpret_cast<const float4 *>(sHin + row * T2_HIN);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 t = h4[k];
                    tgt[4 * k] = t.x; tgt[4 * k + 1] = t.y; tgt[4 * k + 2] = t.z; tgt[4 * k + 3] = t.w;
                }
            }
            const float h_adv = sHin[row * T2_HIN + 16], h_lpold = sHin[row * T2_HIN + 17];
            if (rvalid && out && cg == 1) {
#pragma unroll
                for (int n = 0; n < T2_NO; ++n) if (n < n_out) out[gr * n_out + n] = y[n];
            }
            if (rvalid) {
                if (head.kind == MQ_HEAD_DY) {
#pragma unroll
                    for (int n = 0; n < T2_NO; ++n) if (n < n_out) dz[n] = mq_act_grad(y[n], tgt[n], lastact, lastleak);
                } else if (head.kind == MQ_HEAD_DIFF) {
#pragma unroll
                    for (int n = 0; n < T2_NO; ++n) if (n < n_out) dz[n] = mq_act_grad(y[n], tgt[n] - y[n], lastact, lastleak);
                } else if (head.kind == MQ_HEAD_SOFTMAX_CE) {
                    float mx = y[0];
#pragma unroll
                    for (int n = 1; n < T2_NO; ++n) if (n < n_out) mx = fmaxf(mx, y[n]);
                    float sum = 0.0f;
#pragma unroll
                    for (int n = 0; n < T2_NO; ++n) if (n < n_out) sum += expf(y[n] - mx);
                    const float inv = 1.0f / sum;
#pragma unroll
                    for (int n = 0; n < T2_NO; ++n)
                        if (n < n_out) dz[n] = mq_act_grad(y[n], tgt[n] - expf(y[n] - mx) * inv - head.p0 * y[n], lastact, lastleak);
                } else if (mq_head_discrete(head.kind)) {
                    // categorical log-prob (distrib.py:24-32): the target row is the one-hot action
                    float mx = y[0];
#pragma unroll
                    for (int n = 1; n < T2_NO; ++n) if (n < n_out) mx = fmaxf(mx, y[n]);
                    float sum = 0.0f;
#pragma unroll
                    for (int n = 0; n < T2_NO; ++n) if (n < n_out) sum += expf(y[n] - mx);
                    const float lse = mx + logf(sum);
                    float lp = 0.0f;
#pragma unroll
                    for (int n = 0; n < T2_NO; ++n) if (n < n_out) lp += tgt[n] * (y[n] - lse);
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
                    for (int n = 0; n < T2_NO; ++n)
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
                for (int k = 0; k < 4; ++k)
                    if (cg == k) { s0 = dz[4 * k]; s1 = dz[4 * k + 1]; s2 = dz[4 * k + 2]; s3 = dz[4 * k + 3]; }
                uint32_t a0, a1, a2, b0, b1, b2;
                tc_split_pair(s0, s1, a0, a1, a2);
                tc_split_pair(s2, s3, b0, b1, b2);
                unsigned char *d = t2_chunk_ptr(gs + p.off_dzl, 2 * T2_CB, cg >> 1, row) + (cg & 1) * 8;
                *reinterpret_cast<uint2 *>(d) = make_uint2(a0, b0);
                *reinterpret_cast<uint2 *>(d - 2 * T2_CB) = make_uint2(a1, b1);
                *reinterpret_cast<uint2 *>(d - 4 * T2_CB) = make_uint2(a2, b2);
            }
        }
        t2_publish(g);
        commit_head_inputs();               // the head phase is over: its inputs can be replaced by the next tile's
        commit_idx((ord + 2) % 3);
        T2_MARK(11);
        if (!p.with_grad) continue;
        // ---- backward -----------------------------------------------------------------------------
        for (int l = L - 1; l >= 0; --l) {
            if (gw == 0) {
                if (tc_elect_one()) {
                    if (l == L - 1) {
                        // dA_l = dZ_l W_l^T (K = 16 outputs), dW_l (+)= A_l^T dZ_l (direct form), db_l / dlogstd (+)= 1^T dZ_l
                        t2_rows_product(tm, sb + p.off_dzl + 2 * (2 * T2_CB), 2 * T2_CB,
                                        tc_wb_mn(sw + p.off_wb[l], p.np[l]), 1, (uint32_t)p.wb_plane[l] >> 4, p.kp[l], 1);
                        t2_feat_product(tm + p.col_acc[l], sb + p.off_act[l], p.act_plane[l], sb + p.off_dzl,
                                        2 * T2_CB, T2_NO, 0, dw_ksteps, dw_started);
                        const TcView ones = t2_feat_mn(sb + p.off_act[l] + 3 * p.act_plane[l]);
                        const TcView dzl = t2_feat_mn(sb + p.off_dzl);
                        const uint32_t dzp16 = (2u * T2_CB) >> 4;
                        const uint32_t idr = tc_idesc(64, T2_NO, 1, 1);
                        uint32_t alo = ones.lo, blo = dzl.lo;
#pragma unroll 1
                        for (int k = 0; k < dw_ksteps; ++k) {    // row 0 of the result: column sums of dZ_0 + dZ_1 + dZ_2
                            tc_mma(tm + p.col_rb, alo, ones.hi, blo + 2 * dzp16, dzl.hi, idr, (dw_started || k > 0) ? 1u : 0u);
                            tc_mma(tm + p.col_rb, alo, ones.hi, blo + dzp16, dzl.hi, idr, 1u);
                            tc_mma(tm + p.col_rb, alo, ones.hi, blo, dzl.hi, idr, 1u);
                            alo += ones.k16;
                            blo += dzl.k16;
                        }
                    } else {
                        // dZ_l sits in act[l+1].  dA_l = dZ_l W_l^T (l > 0); dW_l^T, db_l (+)= dZ_l^T [A_l | 1]
                        if (l > 0)
                            t2_rows_product(tm, sb + p.off_act[l + 1] + 2 * p.act_plane[l + 1], p.act_plane[l + 1],
                                            tc_wb_mn(sw + p.off_wb[l], p.np[l]), 1, (uint32_t)p.wb_plane[l] >> 4, p.kp[l],
                                            p.np[l] / 16);
                        t2_feat_product(tm + p.col_acc[l], sb + p.off_act[l + 1], p.act_plane[l + 1], sb + p.off_act[l],
                                        p.act_plane[l], p.kp[l], 8, dw_ksteps, dw_started);
                    }
                    tc_commit(mbar);
                }
                __syncwarp();
            }
            T2_MARK(12);
            if (l == 0) { pending = true; break; }
            tc_wait(mbar, parity); parity ^= 1;
            T2_MARK(13 + l);
            // dZ_{l-1} = dA_l * act'(A_l), written over A_l (A_l is rebuilt exactly from its three planes)
            float a[16], gv[16];
            load_blocks(gv);
            t2_load_chunk(gs + p.off_act[l], p.act_plane[l], cg * 2, row, a);
            t2_load_chunk(gs + p.off_act[l], p.act_plane[l], cg * 2 + 1, row, a + 8);
            const float leak = p.net.leak[l - 1];
#pragma unroll
            for (int j = 0; j < 16; ++j) gv[j] = mq_act_grad(a[j], gv[j], HACT, leak);
            t2_store_chunk(gs + p.off_act[l], p.act_plane[l], cg * 2, row, gv);
            t2_store_chunk(gs + p.off_act[l], p.act_plane[l], cg * 2 + 1, row, gv + 8);
            t2_publish(g);
            T2_MARK(16 + l);
        }
        dw_started = true;
        if ((ord + 1) % T2_FLUSH_TILES == 0 && tile + vstride < ntiles) {
            if (pending) { tc_wait(mbar, parity); parity ^= 1; pending = false; }
            flush_grad();
            dw_started = false;
        }
    }
    if (pending) { tc_wait(mbar, parity); parity ^= 1; }
    T2_MARK(19);

    // ---- per-group partial gradient: TMEM accumulators -> workspace -----------------------------------
    if (p.with_grad) {
        if (dw_started) flush_grad();
        else if (!flushed) {
            float *dst = partial + vt0 * p.net.n_params;
            for (int j = gtid; j < p.net.n_params; j += T2_GNT) dst[j] = 0.0f;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    T2_MARK(20);
#ifdef MQ_TC_PROFILE
    if (prof != nullptr && tid == 0) { unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); prof[65 + 2 * blockIdx.x] = (long long)gt; }
#endif
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}

// ---- host side ------------------------------------------------------------------------------------
static int tc2_plan_build(const mq_net_t *net, int64_t batch, int with_grad, const mq_head_t *head, Tc2Plan *p) {
    memset(p, 0, sizeof(*p));
    p->net = *net;
    const int L = net->n_layers;
    if (L < 2 || L > T2_MAXL) return 0;
    if (net->n_in > 32 || net->n_out > T2_NO) return 0;
    for (int l = 0; l + 1 < L; ++l) {
        if (net->out[l] > T2_HID) return 0;
        if (net->act[l] != net->act[0] || (net->act[l] != MQ_ACT_TANH && net->act[l] != MQ_ACT_LRELU)) return 0;
    }
    const bool gauss = head && mq_head_gauss(head->kind);
    if ((gauss || net->logstd_off >= 0) && net->n_out > 8) return 0;
    p->L = L;
    p->k0 = net->n_in <= 16 ? 16 : 32;
    for (int l = 0; l < L; ++l) {
        p->kp[l] = l == 0 ? p->k0 : T2_HID;
        p->np[l] = l == L - 1 ? T2_NO : T2_HID;
    }
    p->with_grad = with_grad;
    // balance: every group of a full wave gets the same number of tiles
    const int64_t groups = 2 * mq_sm_count();
    const int64_t waves = mq_ceil_div(batch, groups * T2_ROWS);
    int64_t tr = mq_round_up(mq_ceil_div(batch, groups * waves), 8);
    if (tr > T2_ROWS) tr = T2_ROWS;
    if (tr < 8) tr = 8;
    p->tile_rows = (int)tr;
    p->x_step_q = T2_GNT / net->n_in; p->x_step_r = T2_GNT % net->n_in;
    p->h_step_q = T2_GNT / net->n_out; p->h_step_r = T2_GNT % net->n_out;
    int off = 0;
    p->off_dzl = off; off += 3 * 2 * T2_CB;
    for (int l = 0; l < L; ++l) {
        p->act_plane[l] = (p->kp[l] / 8) * T2_CB;
        p->off_act[l] = off; off += 3 * p->act_plane[l] + T2_CB;
    }
    p->off_hin = off; off += T2_ROWS * T2_HIN * 4;
    p->off_idx = off; off += 3 * T2_ROWS * 4;
    p->group_bytes = (int)mq_round_up(off, 1024);
    off = 2 * p->group_bytes;
    for (int l = 0; l < L; ++l) { p->off_wf[l] = off; off += 3 * p->np[l] * p->kp[l] * 2; }
    for (int l = 1; l < L; ++l) {
        p->wb_plane[l] = p->np[l] * p->kp[l] * 2;
        p->off_wb[l] = off; off += 3 * p->wb_plane[l];
    }
    p->off_bias = off; off += T2_MAXL * 64 * 4;
    p->off_misc = off; off += 32 * 4;
    // the ones-row product of the last layer reads 8 chunks from the ones chunk of act[L-1] on: keep that in bounds
    const int reach = p->group_bytes + p->off_act[L - 1] + 3 * p->act_plane[L - 1] + 8 * T2_CB;
    if (off < reach) off = reach;
    p->smem_bytes = (int)mq_round_up(off, 16);
    int col = 2 * T2_HID;                                  // Z / dA: two blocks
    for (int l = 0; l < L; ++l) { p->col_acc[l] = col; col += (l == L - 1) ? T2_NO : p->kp[l] + 8; }
    p->col_rb = col; col += T2_NO;
    if (col > T2_GCOLS) return 0;
    if (p->smem_bytes + 1024 > mq_smem_optin()) return 0;
    return 1;
}

long long *mq_phase_prof_ptr();
// Same contract as mq_tc_launch (mq_tc.cu); *parts_out = partial-gradient rows written (two per CTA).
int mq_tc2_launch(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                  int64_t batch, const mq_head_t *head, int with_grad, float *out, float *logp,
                  void *ws, int64_t ws_bytes, int *parts_out, void *stream) {
    Tc2Plan p;
    if (!tc2_plan_build(net, batch, with_grad, head, &p)) return MQ_ERR_UNSUPPORTED;
    const int64_t ntiles = mq_ceil_div(batch, p.tile_rows);
    const int64_t sms = mq_sm_count();
    const int64_t want = mq_ceil_div(ntiles, 2);
    const int grid = (int)(want < sms ? want : sms);
    if (with_grad)
        MQ_REQUIRE(ws && ws_bytes >= 2 * (int64_t)grid * net->n_params * (int64_t)sizeof(float), MQ_ERR_INVALID,
                   "mq_tc2_launch: workspace too small");
    cudaError_t e = cudaSuccess;
#define T2_LAUNCH(K0_, ACT_)                                                                                              \
    do {                                                                                                                  \
        e = cudaFuncSetAttribute(tc2_grad_kernel<K0_, ACT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem_bytes); \
        if (e == cudaSuccess)                                                                                             \
            e = mq_launch_pdl(tc2_grad_kernel<K0_, ACT_>, dim3(grid), dim3(T2_NT), (size_t)p.smem_bytes, stream, p, theta, \
                              x, row_idx, batch, *head, (float *)ws, out, logp, mq_phase_prof_ptr());                                        \
    } while (0)
    const int hact = net->act[0];
    if (p.k0 == 16 && hact == MQ_ACT_TANH) T2_LAUNCH(16, MQ_ACT_TANH);
    else if (p.k0 == 16) T2_LAUNCH(16, MQ_ACT_LRELU);
    else if (hact == MQ_ACT_TANH) T2_LAUNCH(32, MQ_ACT_TANH);
    else T2_LAUNCH(32, MQ_ACT_LRELU);
#undef T2_LAUNCH
    if (e != cudaSuccess) {
        mq_set_error("mq_tc2_launch: %s", cudaGetErrorString(e));
        return MQ_ERR_CUDA;
    }
    MQ_CHECK_LAUNCH("mq_tc2_launch");
    if (parts_out) *parts_out = 2 * grid;
    return MQ_OK;
}
#else   // MQ_EMU: the host emulation has no tensor cores
int mq_tc2_launch(const mq_net_t *, const float *, const float *, const int32_t *, int64_t, const mq_head_t *, int,
                  float *, float *, void *, int64_t, int *, void *) { return MQ_ERR_UNSUPPORTED; }
#endif
