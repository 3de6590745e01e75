// tcgen05 / TMEM GEMM for WIDE Affine layers (north_star: "tensor-core tiles ... for Affine layers wide and
// batched enough to be real dense contractions"): the 784x400 / 400x784 layers of mnist/vae.py at batch 4096,
// the 784x128 layer of mnist/mlp.py at large batch, and large `matmul` nodes of the general graph
// (mannequin/backprop.py:45-51: y = x W, dx = g W^T, dW = x^T g).
//
// C[M,N] = epilogue(A[M,K] B[K,N]) with fp32 operands of any stride, the epilogues of mq_gemm.cuh and the
// optional row gather / split-K of the FFMA kernel it stands in for (mq_gemm.cu dispatches here for large problems).
//
// fp32 parity on bf16 tensor cores, as in mq_tc.cu: both operands are split exactly into three bf16 planes while
// they are staged into shared memory (global fp32 -> registers -> 3 x 16-byte core-matrix chunks), a product is
// the plane pairs with i + j <= 2 issued as four wide instructions per K step
//     A_0 x [B_0|B_1] (N = 256),  A_1 x [B_0|B_1],  A_0 x B_2 (N = 128),  A_2 x B_0
// into two 128-column fp32 accumulator blocks in TMEM that the epilogue adds.
//
// The tensor pipe TRUNCATES when it adds into its fp32 accumulator (measured: a K = 2003 product accumulated
// entirely in TMEM was low by 1.2e-4 relative, always towards zero), so TMEM only ever accumulates one K block of
// 128 (32 instructions); the blocks are then added in registers (round to nearest) by the warps that own the rows,
// from the accumulator set the tensor pipe is NOT writing (two sets of 256 TMEM columns alternate).
//
// One CTA (512 threads) per 128 x 128 tile of C.  K advances in chunks of 32 through a ring of 3 shared-memory
// stages (60 KB each): while the tensor pipe works on stage s (tcgen05.commit -> the stage's mbarrier when done),
// all warps convert the next chunk, whose global loads were issued one chunk earlier.
#include "mq_gemm.cuh"
#include "mq_tc.cuh"

#ifndef MQ_EMU

#define GT_BM 128
#define GT_BN 128
#define GT_KC 32                      // K per stage: 4 chunks of 8, 2 MMA K steps
#define GT_STAGES 3
#define GT_CPB 4                      // chunks per TMEM accumulation block (K = 128)
// operand stage = [plane 3][row group 16][k chunk 4][8 rows][8 k]: core matrices of 128 B, k chunks 160 B apart and
// row groups 640 B apart (the 32-byte skew makes the 16-byte stores of a quarter warp -- two rows x four k chunks when
// lanes run along k -- land in eight different bank groups); the planes of B are thereby side by side along N
#define GT_KCS 160                    // descriptor leading byte offset: between the k chunks of a row group
#define GT_RGS (4 * GT_KCS)           // descriptor stride byte offset: between 8-row groups
#define GT_PLANE (16 * GT_RGS)
#define GT_A_STAGE (3 * GT_PLANE)
#define GT_B_STAGE (3 * GT_PLANE)
#define GT_STAGE_BYTES (GT_A_STAGE + GT_B_STAGE)
#define GT_SMEM (GT_STAGES * GT_STAGE_BYTES)
#define GT_NT 512

// 8 consecutive k of one row / column -> registers.  KFAST: k is the unit-stride axis of the operand.
template <bool KFAST>
__device__ __forceinline__ void gt_load_item(const float *base, int64_t outer, int64_t outer_stride, int64_t k0,
                                             int64_t k_stride, bool outer_ok, int64_t k_end, bool vec_ok, float *v) {
    if (!outer_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.0f;
        return;
    }
    const float *p = base + outer * outer_stride + k0 * k_stride;
    if (KFAST && vec_ok && k0 + 8 <= k_end) {
        const float4 lo = __ldg(reinterpret_cast<const float4 *>(p)), hi = __ldg(reinterpret_cast<const float4 *>(p) + 1);
        v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (k0 + j < k_end) ? __ldg(p + j * k_stride) : 0.0f;
    }
}

__device__ __forceinline__ void gt_store_item(unsigned char *dst, int plane_stride, const float *v) {
    uint32_t q[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) tc_split_pair(v[2 * j], v[2 * j + 1], q[0][j], q[1][j], q[2][j]);
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
        *reinterpret_cast<uint4 *>(dst + pl * plane_stride) = make_uint4(q[pl][0], q[pl][1], q[pl][2], q[pl][3]);
}

template <bool A_KFAST, bool B_KFAST, bool SPLIT>
__global__ void __launch_bounds__(GT_NT, 1)
gemm_tc_kernel(GemmArgs g) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_bar[GT_STAGES + 2];      // stage free x3, accumulator set ready x2
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int64_t m0 = (int64_t)blockIdx.y * GT_BM, n0 = (int64_t)blockIdx.x * GT_BN;
    const int64_t k_begin = (int64_t)blockIdx.z * g.k_per_split;
    const int64_t k_end = (k_begin + g.k_per_split < g.k) ? k_begin + g.k_per_split : g.k;
    const int n_chunks = (int)((k_end - k_begin + GT_KC - 1) / GT_KC);

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int s = 0; s < GT_STAGES + 2; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&s_bar[s])), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    tc_publish();
    const uint32_t tm = __shfl_sync(0xffffffffu, s_tmem, 0);
    const uint32_t sbase = tc_smem_u32(smem);

    // this thread's A item and B item of a chunk: (outer index within the tile, k chunk 0..3); lanes run along the
    // unit-stride axis of the operand so that the global loads are coalesced (ncu: with one row per lane the L1 tag
    // stage, 64 % busy, was the limiter)
    constexpr int NI = 128 * 4 / GT_NT;
    int a_out[NI], a_kc[NI], b_out[NI], b_kc[NI];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
        const int it = tid + j * GT_NT;
        if (A_KFAST) { a_out[j] = it >> 2; a_kc[j] = it & 3; } else { a_out[j] = it & 127; a_kc[j] = it >> 7; }
        if (B_KFAST) { b_out[j] = it >> 2; b_kc[j] = it & 3; } else { b_out[j] = it & 127; b_kc[j] = it >> 7; }
    }
    // global row of A after the optional gather on the batch axis; validity of the rows / columns
    int64_t a_row[NI];
    bool a_ok[NI], b_ok[NI];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
        const int64_t gm = m0 + a_out[j];
        a_ok[j] = gm < g.m;
        a_row[j] = (a_ok[j] && g.idx && !g.idx_on_k) ? (int64_t)g.idx[gm] : gm;
        b_ok[j] = n0 + b_out[j] < g.n;
    }
    const bool a_vec = A_KFAST && !(g.idx && g.idx_on_k) && (g.a_rs % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.a) & 15) == 0) &&
                       (k_begin % 4 == 0);
    const bool b_vec = B_KFAST && (g.b_cs % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.b) & 15) == 0) && (k_begin % 4 == 0);

    float ra[NI][8], rb[NI][8];
    auto load_chunk = [&](int c) {
        const int64_t kt = k_begin + (int64_t)c * GT_KC;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int64_t k0 = kt + a_kc[j] * 8;
            if (g.idx && g.idx_on_k) {                  // gather along K (dW = x[idx]^T g): element-wise
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    ra[j][e] = (a_ok[j] && k0 + e < k_end) ? __ldg(g.a + a_row[j] * g.a_rs + (int64_t)g.idx[k0 + e] * g.a_cs) : 0.0f;
            } else {
                gt_load_item<A_KFAST>(g.a, a_row[j], g.a_rs, k0, g.a_cs, a_ok[j], k_end, a_vec, ra[j]);
            }
            gt_load_item<B_KFAST>(g.b, n0 + b_out[j], g.b_cs, kt + b_kc[j] * 8, g.b_rs, b_ok[j], k_end, b_vec, rb[j]);
        }
    };
    auto store_chunk = [&](int stage) {
        unsigned char *sa = smem + stage * GT_STAGE_BYTES, *sb = sa + GT_A_STAGE;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            gt_store_item(sa + (a_out[j] >> 3) * GT_RGS + a_kc[j] * GT_KCS + (a_out[j] & 7) * 16, GT_PLANE, ra[j]);
            gt_store_item(sb + (b_out[j] >> 3) * GT_RGS + b_kc[j] * GT_KCS + (b_out[j] & 7) * 16, GT_PLANE, rb[j]);
        }
    };

    // register accumulators of this thread: row quarter*32 + lane, GT_CPT columns
    constexpr int GT_CPT = GT_BN / (GT_NT / 128);
    const int quarter = warp & 3, cgp = warp >> 2;
    const uint32_t tl = tm + ((uint32_t)(quarter * 32) << 16);
    float acc[GT_CPT];
#pragma unroll
    for (int j = 0; j < GT_CPT; ++j) acc[j] = 0.0f;
    uint32_t acc_parity = 0;            // bit s: parity to wait for on accumulator set s
    auto flush_block = [&](int set) {   // acc += the two column blocks of accumulator set `set`
        tc_wait(tc_smem_u32(&s_bar[GT_STAGES + set]), (acc_parity >> set) & 1u);
        acc_parity ^= 1u << set;
#pragma unroll
        for (int q = 0; q < GT_CPT / 16; ++q) {
            float v[16], w[16];
            tc_ld16(tl + set * 256 + cgp * GT_CPT + q * 16, v);
            tc_ld16(tl + set * 256 + GT_BN + cgp * GT_CPT + q * 16, w);
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[q * 16 + j] += v[j] + w[j];
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };

    uint32_t stage_parity = 0;          // bit s: parity to wait for on stage s
    const int n_blocks = (n_chunks + GT_CPB - 1) / GT_CPB;
    if (n_chunks > 0) load_chunk(0);
    for (int c = 0; c < n_chunks; ++c) {
        const int stage = c % GT_STAGES;
        const int blk = c / GT_CPB, set = blk & 1;
        if (c >= GT_STAGES) {           // the MMAs that read this stage three chunks ago must be done
            tc_wait(tc_smem_u32(&s_bar[stage]), (stage_parity >> stage) & 1u);
            stage_parity ^= 1u << stage;
        }
        store_chunk(stage);
        if (c + 1 < n_chunks) load_chunk(c + 1);       // in flight while this chunk is multiplied
        // one chunk into block `blk`, the previous block is complete (or nearly): fold it into the registers while the
        // tensor pipe fills the other set; it is done long before block blk + 1 reuses that set
        if (blk > 0 && c % GT_CPB == 1) flush_block(set ^ 1);
        tc_publish();
        if (warp == 0) {
            if (tc_elect_one()) {
                const uint32_t a0 = sbase + stage * GT_STAGE_BYTES, b0 = a0 + GT_A_STAGE;
                const uint32_t ahi = tc_desc_hi(GT_RGS), bhi = tc_desc_hi(GT_RGS);
                uint32_t alo = tc_desc_lo(a0, GT_KCS), blo = tc_desc_lo(b0, GT_KCS);
                const uint32_t ap16 = GT_PLANE >> 4, bp16 = GT_PLANE >> 4;
                const uint32_t id2 = tc_idesc(GT_BM, 2 * GT_BN, 0, 0), id1 = tc_idesc(GT_BM, GT_BN, 0, 0);
                const uint32_t d = tm + set * 256;
#pragma unroll
                for (int ks = 0; ks < GT_KC / 16; ++ks) {
                    tc_mma(d, alo, ahi, blo, bhi, id2, (c % GT_CPB > 0 || ks > 0) ? 1u : 0u);     // A_0 x [B_0|B_1]
                    tc_mma(d, alo + ap16, ahi, blo, bhi, id2, 1u);                                 // A_1 x [B_0|B_1]
                    tc_mma(d, alo, ahi, blo + 2 * bp16, bhi, id1, 1u);                             // A_0 x B_2
                    tc_mma(d, alo + 2 * ap16, ahi, blo, bhi, id1, 1u);                             // A_2 x B_0
                    alo += (2 * GT_KCS) >> 4;
                    blo += (2 * GT_KCS) >> 4;
                }
                tc_commit(tc_smem_u32(&s_bar[stage]));
                if (c % GT_CPB == GT_CPB - 1 || c == n_chunks - 1) tc_commit(tc_smem_u32(&s_bar[GT_STAGES + set]));
            }
            __syncwarp();
        }
    }
    // ---- epilogue: remaining accumulator blocks, then registers -> C ---------------------------------------
    if (n_blocks > 0) {
        // a block whose flush point (one chunk into the next block) was never reached is still pending
        const int last = n_blocks - 1;
        if (last > 0 && (n_chunks - 1) % GT_CPB < 1) flush_block((last - 1) & 1);
        flush_block(last & 1);
    }
    {
        const int64_t gm = m0 + quarter * 32 + lane;
        if (gm < g.m) {
#pragma unroll
            for (int j = 0; j < GT_CPT; ++j) {
                const int64_t gn = n0 + cgp * GT_CPT + j;
                if (gn >= g.n) continue;
                if (SPLIT) g.c[((int64_t)blockIdx.z * g.m + gm) * g.n + gn] = acc[j];
                else g.c[gm * g.n + gn] = gemm_epilogue(g, acc[j], gm, gn);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}

static int g_gemm_tc_mode = 1;       // 0: never, 1: automatic (large problems), 2: whenever possible
extern "C" int mq_debug_set_gemm_tc_mode(int mode) { g_gemm_tc_mode = mode; return MQ_OK; }

// Launches the tensor-core GEMM if the problem is large enough; MQ_ERR_UNSUPPORTED otherwise.
// splits > 1: g.c points at the split-K partial buffer [splits][M][N] and g.k_per_split is set by the caller.
int mq_gemm_tc_launch(const GemmArgs &g, int64_t splits, void *stream) {
    if (g_gemm_tc_mode == 0) return MQ_ERR_UNSUPPORTED;
    if (g_gemm_tc_mode == 1 && (g.m * g.n * g.k < (1LL << 26) || g.n < 64 || g.k < 64 || g.m < 128)) return MQ_ERR_UNSUPPORTED;
    if (g.m <= 0 || g.n <= 0) return MQ_ERR_UNSUPPORTED;
    const bool a_kfast = g.a_cs == 1, b_kfast = g.b_rs == 1;
    dim3 grid((unsigned)mq_ceil_div(g.n, GT_BN), (unsigned)mq_ceil_div(g.m, GT_BM), (unsigned)splits);
    cudaError_t e = cudaSuccess;
#define GT_GO(AK, BK, SP)                                                                                          \
    do {                                                                                                           \
        e = cudaFuncSetAttribute(gemm_tc_kernel<AK, BK, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM); \
        if (e == cudaSuccess) gemm_tc_kernel<AK, BK, SP><<<grid, GT_NT, GT_SMEM, (cudaStream_t)stream>>>(g);       \
    } while (0)
    if (splits > 1) {
        if (a_kfast && b_kfast) GT_GO(true, true, true); else if (a_kfast) GT_GO(true, false, true);
        else if (b_kfast) GT_GO(false, true, true); else GT_GO(false, false, true);
    } else {
        if (a_kfast && b_kfast) GT_GO(true, true, false); else if (a_kfast) GT_GO(true, false, false);
        else if (b_kfast) GT_GO(false, true, false); else GT_GO(false, false, false);
    }
#undef GT_GO
    if (e != cudaSuccess) { mq_set_error("mq_gemm_tc: %s", cudaGetErrorString(e)); return MQ_ERR_CUDA; }
    return MQ_OK;
}

#else
extern "C" int mq_debug_set_gemm_tc_mode(int) { return MQ_OK; }
int mq_gemm_tc_launch(const GemmArgs &, int64_t, void *) { return MQ_ERR_UNSUPPORTED; }
#endif
