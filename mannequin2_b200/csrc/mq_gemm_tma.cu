// tcgen05 / TMEM GEMM for WIDE Affine layers, fed by TMA (north_star: "tensor-core tiles ... for Affine layers wide and
// batched enough to be real dense contractions"): the 784x400 / 400x784 layers of mnist/vae.py at batch 4096, the 784x128
// layer of mnist/mlp.py at large batch, and large `matmul` nodes of the general graph (mannequin/backprop.py:45-51:
// y = x W, dx = g W^T, dW = x^T g).
//
// C[M,N] = epilogue(A[M,K] B[K,N]) with fp32 operands of any stride, the epilogues of mq_gemm.cuh and the optional row
// gather / split-K of the FFMA kernel it stands in for (mq_gemm.cu dispatches here for large problems).
//
// Two launches per product:
//   1. pack_planes_kernel: each fp32 operand is converted ONCE into P bf16 planes (x = x0 + x1 + x2, as in mq_tc.cu),
//      stored as plain K-contiguous bf16 matrices [plane][rows padded to 128][K padded to 64] in the workspace.  (Round 1
//      converted fp32 -> planes through registers inside every CTA that touched a tile: 3.9x the algorithmic DRAM reads
//      and the shared-memory bandwidth the tensor pipe needs.)
//   2. gemm_tma_kernel: one CTA per 128 x 128 tile of C, warp-specialised:
//        warp 0    TMA producer: cp.async.bulk.tensor.2d loads of 128 x 64 bf16 boxes (128-byte swizzle) of every plane of
//                  A and B into a ring of shared-memory stages, completion counted in bytes on the stage's mbarrier
//        warp 1    MMA issuer: per K step of 16 the plane pairs i + j < P as wide instructions
//                      P = 3: A_0 x [B_0|B_1] (N = 256), A_1 x [B_0|B_1], A_0 x B_2 (N = 128), A_2 x B_0
//                      P = 2: A_0 x [B_0|B_1], A_1 x B_0          P = 1: A_0 x B_0
//                  (the planes of B are consecutive tiles in a stage, so [B_0|B_1] is one descriptor), tcgen05.commit
//                  frees the stage
//        warps 2-9 accumulator owners: the tensor pipe TRUNCATES when it adds into its fp32 accumulator (a K = 2003
//                  product accumulated entirely in TMEM was low by 1.2e-4 relative), so for P >= 2 TMEM only ever
//                  accumulates one K block of 128; the owners add each finished block into registers (round to nearest)
//                  from the accumulator set the tensor pipe is NOT writing (two sets of 256 columns alternate), then
//                  apply the epilogue.  P = 1 (plain bf16, error 2^-9 per operand) accumulates all of K in TMEM.
// The precision modes are those of the fused kernel (include/mq_b200.h MQ_TC_*): 3 planes = fp32 products (the FFMA
// bar), 2 planes = 16-bit operands, 1 plane = bf16; 2 planes with MQ_TC_FP16 = fp16 planes with the low one scaled by 2^11
// (x = x0 + 2^-11 x1: fp32 operands to 2^-24, the FFMA bar again, in three plane pairs instead of six -- block 1 of an
// accumulator set holds exactly the scaled pairs A_0 B_1' + A_1' B_0 and is folded in with one FMA by the owners).
#ifndef MQ_EMU
#include <cuda.h>
#endif
#include <stdlib.h>

#include "mq_gemm.cuh"
#include "mq_tc.cuh"

#ifndef MQ_EMU
#include <cuda_fp16.h>

#define GM_BM 128
#define GM_BN 128
#define GM_BK 64                       // K per stage: one 128-byte swizzle atom of bf16
#define GM_TILE_BYTES (128 * 128)      // one 128 x 64 bf16 box
#define GM_NT 320                      // producer warp, issuer warp, 8 accumulator-owner warps
#define GM_KBLOCK 2                    // stages per TMEM accumulation block (K = 128)

// ---- operand packing -------------------------------------------------------------------------------------------
// A job (GmPackJob, mq_gemm.cuh) writes out[pl][row_off + r][k_off + k] (bf16 / fp16; r < rows_cover, k < k_cover; zero
// outside the source) = plane pl of src(r, k), where src(r, k) = base[row(r) * rs + col(k) * cs] with an optional gather on
// the rows (idx_on_k = 0) or on k (idx_on_k = 1).  A whole operand is one job with cover = (rows_pad, k_pad) and no offsets.
//
// Plain form: one thread converts 8 consecutive k of one row; lanes run along the unit-stride axis of the source.
template <int P, int F>
__device__ __forceinline__ void gm_pack_plain(const GmPackJob &jb, int rows_fast, int64_t item) {
    const int64_t kc_n = jb.k_cover / 8;
    if (item >= jb.rows_cover * kc_n) return;
    int64_t r, kc;
    if (rows_fast) { kc = item / jb.rows_cover; r = item - kc * jb.rows_cover; }
    else { r = item / kc_n; kc = item - r * kc_n; }
    const float *__restrict__ base = jb.base;
    const int32_t *__restrict__ idx = jb.idx;
    const int64_t rs = jb.rs, cs = jb.cs;
    const int64_t k = jb.ones_col > 0 ? jb.ones_col : jb.k;     // columns the source really has
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.0f;
    if (jb.ones_row > 0 && r >= jb.ones_row && r < jb.rows) {   // rows that read as 1.0 (the bias-gradient row of dW = x^T g)
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (kc * 8 + j < k) ? 1.0f : 0.0f;
    } else if (r < jb.rows) {
        const int64_t sr = (idx && !jb.idx_on_k) ? (int64_t)idx[r] : r;
        const int64_t k0 = kc * 8;
        if (cs == 1 && !(idx && jb.idx_on_k) && k0 + 8 <= k && ((reinterpret_cast<uintptr_t>(base) | (uintptr_t)(rs * 4)) & 15) == 0) {
            const float4 lo = __ldg(reinterpret_cast<const float4 *>(base + sr * rs + k0));
            const float4 hi = __ldg(reinterpret_cast<const float4 *>(base + sr * rs + k0) + 1);
            v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (k0 + j < k) {
                    const int64_t sk = (idx && jb.idx_on_k) ? (int64_t)idx[k0 + j] : k0 + j;
                    v[j] = __ldg(base + sr * rs + sk * cs);
                }
        }
        if (jb.ones_col > 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (kc * 8 + j >= jb.ones_col && kc * 8 + j < jb.k) v[j] = 1.0f;
        }
    }
    uint32_t w[4][P];
#pragma unroll
    for (int j = 0; j < 4; ++j) gm_split_pair<P, F>(v[2 * j], v[2 * j + 1], w[j]);
    __nv_bfloat16 *out = (__nv_bfloat16 *)jb.out;
#pragma unroll
    for (int pl = 0; pl < P; ++pl)
        *reinterpret_cast<uint4 *>(out + ((int64_t)pl * jb.rows_pad + jb.row_off + r) * jb.k_pad + jb.k_off + kc * 8) =
            make_uint4(w[0][pl], w[1][pl], w[2][pl], w[3][pl]);
}

// Transposing form for operands whose ROWS are the unit-stride axis of the source (A = x^T of dW = x^T g: src(r, k) = x[k][r];
// B = W of y = x W), through a shared-memory tile: reads run along r (coalesced), writes along k (128 contiguous bytes per
// row and plane instead of 16-byte pieces a whole packed row apart).  Tile = 32 rows x 64 k per CTA of 256 threads; the
// source is addressed as base[row(r) + col(k) * cs].
template <int P, int F>
__device__ __forceinline__ void gm_pack_transposed(const GmPackJob &jb, int64_t bx, int64_t by, float (*tile)[33]) {
    const float *__restrict__ base = jb.base;
    const int32_t *__restrict__ idx = jb.idx;
    const int64_t r0 = bx * 32, k0 = by * 64;
    const int tr = threadIdx.x & 31, tk = threadIdx.x >> 5;
#pragma unroll
    for (int pass = 0; pass < 8; ++pass) {
        const int kk = pass * 8 + tk;
        const int64_t r = r0 + tr, kg = k0 + kk;
        float v = 0.0f;
        if (r < jb.rows && kg < jb.k) {
            if (jb.ones_row > 0 && r >= jb.ones_row) v = 1.0f;
            else {
                const int64_t sr = (idx && !jb.idx_on_k) ? (int64_t)idx[r] : r;
                const int64_t sk = (idx && jb.idx_on_k) ? (int64_t)idx[kg] : kg;
                v = __ldg(base + sr + sk * jb.cs);
            }
        }
        tile[kk][tr] = v;
    }
    __syncthreads();
    const int m = threadIdx.x >> 3, kc = threadIdx.x & 7;
    uint32_t w[4][P];
#pragma unroll
    for (int j = 0; j < 4; ++j) gm_split_pair<P, F>(tile[kc * 8 + 2 * j][m], tile[kc * 8 + 2 * j + 1][m], w[j]);
    const int64_t r = r0 + m;
    __nv_bfloat16 *out = (__nv_bfloat16 *)jb.out;
    if (r < jb.rows_cover && k0 + kc * 8 < jb.k_cover) {
#pragma unroll
        for (int pl = 0; pl < P; ++pl)
            *reinterpret_cast<uint4 *>(out + ((int64_t)pl * jb.rows_pad + jb.row_off + r) * jb.k_pad + jb.k_off + k0 + kc * 8) =
                make_uint4(w[0][pl], w[1][pl], w[2][pl], w[3][pl]);
    }
}

static inline bool gm_job_transposed(const GmPackJob &jb) { return jb.rs == 1 && jb.cs != 1; }
static inline unsigned gm_job_blocks(const GmPackJob &jb) {
    if (gm_job_transposed(jb)) return (unsigned)(mq_ceil_div(jb.rows_cover, 32) * mq_ceil_div(jb.k_cover, 64));
    return (unsigned)mq_ceil_div(jb.rows_cover * (jb.k_cover / 8), MQ_NT);
}

// Up to GM_MAX_PACK_JOBS jobs in ONE launch (the weights of a whole net and its input batch, csrc/mq_vae.cu): a block
// finds its job by the block ranges the launcher filled in.
struct GmPackJobs { int n; GmPackJob job[GM_MAX_PACK_JOBS]; };

template <int P, int F>
__global__ void __launch_bounds__(MQ_NT)
pack_jobs_kernel(const __grid_constant__ GmPackJobs jobs) {
    __shared__ float tile[64][33];
    int j = 0;
    while (j + 1 < jobs.n && blockIdx.x >= jobs.job[j + 1].block0) ++j;
    const GmPackJob &jb = jobs.job[j];
    const int64_t local = (int64_t)blockIdx.x - jb.block0;
    if (jb.rs == 1 && jb.cs != 1) {
        // k blocks fastest: blocks that run together write neighbouring 128-byte pieces of the same packed rows (with the row
        // blocks fastest every piece landed in another DRAM page: 4096^2 floats packed at 2.4 TB/s)
        const int64_t nby = (jb.k_cover + 63) / 64;
        gm_pack_transposed<P, F>(jb, local / nby, local % nby, tile);
    } else {
        gm_pack_plain<P, F>(jb, 0, local * MQ_NT + threadIdx.x);
    }
}

// ---- PTX helpers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ void gm_tma_load(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(mbar) : "memory");
}
__device__ __forceinline__ void gm_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gm_arrive(uint32_t mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}
// K-major operand tile with 128-byte swizzle: 8-row groups 1024 bytes apart; layout type 2 (SWIZZLE_128B) in bits 61-63
__device__ __forceinline__ uint32_t gm_desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | (1u << 16); }
// MN-major operand tile with 128-byte swizzle (the operand's M / N axis is the unit-stride axis: x^T of dW = x^T g read from
// the planes of x): a tile is two blocks of [64 k rows][64 m = 128 bytes] (8 KB each, one TMA box each); leading byte offset
// = distance between the 64-element blocks along M / N (8192), stride byte offset = distance between groups of 8 k rows
// (1024); a K step of 16 is two such groups (2048 bytes).
__device__ __forceinline__ uint32_t gm_desc_lo_mn(uint32_t addr) { return ((addr >> 4) & 0x3FFFu) | ((8192u >> 4) << 16); }
__device__ __forceinline__ uint32_t gm_desc_hi() { return ((1024u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29); }

template <int P>
struct GmCfg {
    static constexpr int STAGES = P == 3 ? 2 : (P == 2 ? 3 : 5);
    static constexpr int STAGE_BYTES = 2 * P * GM_TILE_BYTES;
    static constexpr int EPI_BYTES = 8 * 32 * 33 * 4;              // epilogue staging: 32 rows x 32 columns (+ pad) per owner warp
    static constexpr int SMEM = STAGES * STAGE_BYTES + EPI_BYTES + 1024;       // + slack for the 1024-byte alignment
    static constexpr bool FOLD = P > 1;
    static constexpr int NB = P > 1 ? 2 : 1;                       // column blocks of an accumulator set
    static constexpr int SET_COLS = P > 1 ? 256 : 128;             // two accumulator sets alternate
    static constexpr int TMEM_COLS = 2 * SET_COLS;
};

// PERSISTENT: min(tiles, SMs) CTAs, CTA b takes tiles b, b + gridDim.x, ... (tile = (split z, row tile, column tile),
// column tiles fastest).  The three roles walk the same tile sequence; stage / accumulator-set indices and barrier
// parities follow RUNNING chunk and block counters, so the producer and the issuer run into the next tile while the owners
// still write the previous tile's C (the owners hold it in registers; its accumulator set was released when they read it).
// YPRE: the act' epilogue fetches its saved activations in groups of 8 rows, one group AHEAD (up to 16 rows of a warp in
// flight) instead of 4 rows at the top of their own iteration.  Chosen by the launcher for SHORT contractions (K <= 256), where
// a tile is all epilogue and eight exposed round trips per half made the 65,536 x 128 x 128 dX product of the MLP take 70 us
// against 37 us with the bias epilogue; as the only form it cost every other product registers and code (the fused VAE step
// 0.304 -> 0.311 ms), hence a separate instantiation.
template <int P, int F, bool SPLIT, bool YPRE = false>
__global__ void __launch_bounds__(GM_NT, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, GemmArgs g, GmEmit em,
                int a_rows_pad, int b_rows_pad, int n_splits, int a_mn, int b_mn) {
    using Cfg = GmCfg<P>;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_full[Cfg::STAGES], s_empty[Cfg::STAGES], s_accfull[2], s_accempty[2];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const uint32_t sbase = (tc_smem_u32(smem_raw) + 1023u) & ~1023u;
    const int tiles_n = (int)((g.n + GM_BN - 1) / GM_BN), tiles_m = (int)((g.m + GM_BM - 1) / GM_BM);
    const int64_t tiles_mn = (int64_t)tiles_m * tiles_n, n_tiles = tiles_mn * n_splits;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"((uint32_t)Cfg::TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 32) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&s_full[s])), "r"(1u));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&s_empty[s])), "r"(1u));
        }
        for (int s = 0; s < 2; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&s_accfull[s])), "r"(1u));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&s_accempty[s])), "r"(8u));     // one arrive per owner warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    }
    tc_publish();
    const uint32_t tm = __shfl_sync(0xffffffffu, s_tmem, 0);
    // programmatic dependent launch: everything above (TMEM allocation, barrier init, tensor-map prefetch) may run while the
    // previous kernel of the stream finishes; nothing that kernel wrote is touched before this wait, and the next kernel may
    // be scheduled as soon as SMs free up
    mq_pdl_wait();
    mq_pdl_trigger();

    // tile t -> (m0, n0, first chunk, chunks)
#define GM_TILE_DECODE(t_)                                                                                         \
    const int64_t z_ = (t_) / tiles_mn, rem_ = (t_) - z_ * tiles_mn;                                                \
    const int m0 = (int)(rem_ / tiles_n) * GM_BM, n0 = (int)(rem_ % tiles_n) * GM_BN;                               \
    const int64_t k_begin = z_ * g.k_per_split;                                                                     \
    const int64_t k_end = (k_begin + g.k_per_split < g.k) ? k_begin + g.k_per_split : g.k;                          \
    const int n_chunks = (int)((k_end - k_begin + GM_BK - 1) / GM_BK);                                              \
    const int n_blocks = Cfg::FOLD ? (n_chunks + GM_KBLOCK - 1) / GM_KBLOCK : 1

    if (warp == 0) {
        // ---- TMA producer ----------------------------------------------------------------------------------------
        if (tc_elect_one()) {
            uint32_t gc = 0;                                   // chunks loaded so far, all tiles
            for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                GM_TILE_DECODE(t);
                (void)n_blocks;
                for (int c = 0; c < n_chunks; ++c, ++gc) {
                    const uint32_t s = gc % Cfg::STAGES;
                    if (gc >= (uint32_t)Cfg::STAGES) tc_wait(tc_smem_u32(&s_empty[s]), (gc / Cfg::STAGES - 1) & 1u);
                    const uint32_t full = tc_smem_u32(&s_full[s]);
                    gm_expect_tx(full, (uint32_t)Cfg::STAGE_BYTES);
                    const uint32_t sa = sbase + s * Cfg::STAGE_BYTES, sb = sa + P * GM_TILE_BYTES;
                    const int kc = (int)(k_begin + (int64_t)c * GM_BK);
#pragma unroll
                    for (int pl = 0; pl < P; ++pl) {
                        if (!a_mn) gm_tma_load(sa + pl * GM_TILE_BYTES, &map_a, kc, pl * a_rows_pad + m0, full);
                        else {
                            gm_tma_load(sa + pl * GM_TILE_BYTES, &map_a, m0, pl * a_rows_pad + kc, full);
                            gm_tma_load(sa + pl * GM_TILE_BYTES + 8192, &map_a, m0 + 64, pl * a_rows_pad + kc, full);
                        }
                        if (!b_mn) gm_tma_load(sb + pl * GM_TILE_BYTES, &map_b, kc, pl * b_rows_pad + n0, full);
                        else {
                            gm_tma_load(sb + pl * GM_TILE_BYTES, &map_b, n0, pl * b_rows_pad + kc, full);
                            gm_tma_load(sb + pl * GM_TILE_BYTES + 8192, &map_b, n0 + 64, pl * b_rows_pad + kc, full);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---- MMA issuer --------------------------------------------------------------------------------------------
        const uint32_t hi = gm_desc_hi();
        const uint32_t id2 = tc_idesc_fmt(GM_BM, 2 * GM_BN, a_mn, b_mn, F), id1 = tc_idesc_fmt(GM_BM, GM_BN, a_mn, b_mn, F);
        const uint32_t a_k16 = a_mn ? (2048u >> 4) : 2u, b_k16 = b_mn ? (2048u >> 4) : 2u;     // descriptor advance per K step of 16
        uint32_t gc = 0, gb = 0;                               // chunks / accumulator blocks issued so far, all tiles
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            GM_TILE_DECODE(t);
            (void)m0; (void)n0;
            for (int c = 0; c < n_chunks; ++c, ++gc) {
                const uint32_t s = gc % Cfg::STAGES;
                const uint32_t gbl = gb + (Cfg::FOLD ? (uint32_t)(c / GM_KBLOCK) : 0u), set = gbl & 1u;
                tc_wait(tc_smem_u32(&s_full[s]), (gc / Cfg::STAGES) & 1u);
                const bool block_start = Cfg::FOLD ? (c % GM_KBLOCK == 0) : (c == 0);
                if (block_start && gbl >= 2u) tc_wait(tc_smem_u32(&s_accempty[set]), (gbl / 2u - 1u) & 1u);
                if (tc_elect_one()) {
                    const uint32_t sa = sbase + s * Cfg::STAGE_BYTES, sb = sa + P * GM_TILE_BYTES;
                    const uint32_t d = tm + set * Cfg::SET_COLS;
                    const uint32_t pl16 = GM_TILE_BYTES >> 4;
#pragma unroll
                    for (int ks = 0; ks < GM_BK / 16; ++ks) {
                        const uint32_t alo = (a_mn ? gm_desc_lo_mn(sa) : gm_desc_lo(sa)) + a_k16 * ks;
                        const uint32_t blo = (b_mn ? gm_desc_lo_mn(sb) : gm_desc_lo(sb)) + b_k16 * ks;
                        const uint32_t first = Cfg::FOLD ? ((c % GM_KBLOCK > 0 || ks > 0) ? 1u : 0u) : ((c > 0 || ks > 0) ? 1u : 0u);
                        if (P == 1) {
                            tc_mma(d, alo, hi, blo, hi, id1, first);                              // A_0 x B_0
                        } else if (P == 2) {
                            tc_mma(d, alo, hi, blo, hi, id2, first);                              // A_0 x [B_0|B_1] -> blocks 0 | 1
                            tc_mma(d + GM_BN, alo + pl16, hi, blo, hi, id1, 1u);                  // A_1 x B_0       -> block 1
                        } else {
                            // block 0 receives ONLY the leading term A_0 B_0 (one truncating add per K step), block 1 all the
                            // correction terms (<= 2^-8 of it): the accumulator's round-towards-zero bias is 4x smaller than
                            // with the corrections spread over both blocks, at the same tensor-pipe time (128 + 4 x 64 cycles)
                            tc_mma(d, alo, hi, blo, hi, id2, first);                              // A_0 x [B_0|B_1] -> blocks 0 | 1
                            tc_mma(d + GM_BN, alo + pl16, hi, blo, hi, id1, 1u);                  // A_1 x B_0       -> block 1
                            tc_mma(d + GM_BN, alo + pl16, hi, blo + pl16, hi, id1, 1u);           // A_1 x B_1       -> block 1
                            tc_mma(d + GM_BN, alo, hi, blo + 2 * pl16, hi, id1, 1u);              // A_0 x B_2       -> block 1
                            tc_mma(d + GM_BN, alo + 2 * pl16, hi, blo, hi, id1, 1u);              // A_2 x B_0       -> block 1
                        }
                    }
                    tc_commit(tc_smem_u32(&s_empty[s]));
                    if (c == n_chunks - 1 || (Cfg::FOLD && c % GM_KBLOCK == GM_KBLOCK - 1)) tc_commit(tc_smem_u32(&s_accfull[set]));
                }
                __syncwarp();
            }
            gb += (uint32_t)n_blocks;
        }
    } else {
        // ---- accumulator owners: row = TMEM lane, 64 columns each ----------------------------------------------------
        const int quarter = warp & 3, half = (warp - 2) >> 2;
        const uint32_t tl = tm + ((uint32_t)(quarter * 32) << 16) + half * 64;
        float *stage = reinterpret_cast<float *>(smem_raw + (sbase - tc_smem_u32(smem_raw)) + Cfg::STAGES * Cfg::STAGE_BYTES) + (warp - 2) * (32 * 33);
        const bool emit = !SPLIT && (em.out_k != nullptr || em.out_t != nullptr);
        const int64_t ldc = em.ldc > 0 ? em.ldc : g.n;
        uint32_t gb = 0;
        for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            GM_TILE_DECODE(t);
            (void)n_chunks;
            float acc[64];
#pragma unroll
            for (int j = 0; j < 64; ++j) acc[j] = 0.0f;
            // the tensor pipe TRUNCATES when it adds into its accumulator, so for P >= 2 a set only ever accumulates one K block
            // of 128; the owners add each finished block into registers (round to nearest) from the set the pipe is not writing
            for (int blk = 0; blk < n_blocks; ++blk) {
                const uint32_t gbl = gb + (uint32_t)blk, set = gbl & 1u;
                tc_wait(tc_smem_u32(&s_accfull[set]), (gbl / 2u) & 1u);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float v[16], w[16];
                    tc_ld16(tl + set * Cfg::SET_COLS + q * 16, v);
                    if (Cfg::NB == 2) tc_ld16(tl + set * Cfg::SET_COLS + GM_BN + q * 16, w);
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[q * 16 + j] += Cfg::NB == 2 ? (F == 1 ? fmaf(w[j], GM_LOW_UNSCALE, v[j]) : v[j] + w[j]) : v[j];
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) gm_arrive(tc_smem_u32(&s_accempty[set]));
            }
            gb += (uint32_t)n_blocks;
            // registers -> C through the warp's staging tile, 32 columns at a time, so that a warp writes whole 128-byte row
            // segments instead of 32 rows x 4 bytes per instruction -- and reads bias / the saved activations of the act'
            // epilogue the same way
            // (a real loop over the two halves, the registers picked by selects: unrolled, the epilogue was 5,000
            // instructions of straight-line code executed once per tile, and its instruction fetch cost 6 us per launch)
#pragma unroll 1
            for (int hc = 0; hc < 2; ++hc) {
#pragma unroll
                for (int j = 0; j < 32; ++j) stage[lane * 33 + j] = hc ? acc[32 + j] : acc[j];
                __syncwarp();
                const int64_t gn = n0 + half * 64 + hc * 32 + lane;
                const int64_t gm_base = m0 + quarter * 32;
                const bool col_ok = gn < g.n;
                // rows in groups of 4: what a group reads from global memory (the saved activations of the act' epilogue) is
                // fetched BEFORE its rows are processed -- with eight owner warps per SM a load inside the row loop is an
                // exposed L2 round trip per row
                const float bias_n = (!SPLIT && g.epi == EPI_BIAS_ACT && col_ok) ? __ldg(g.bias + gn) : 0.0f;
                constexpr int RG = YPRE ? 8 : 4;
                const bool want_y = !SPLIT && g.epi == EPI_ACT_GRAD;
                float yn[RG];
#pragma unroll
                for (int i = 0; i < RG; ++i) yn[i] = 0.0f;
                if (YPRE && want_y) {
                    const float *yp = g.y + gm_base * g.n + gn;
#pragma unroll
                    for (int i = 0; i < RG; ++i) yn[i] = (col_ok && gm_base + i < g.m) ? __ldg(yp + i * g.n) : 0.0f;
                }
#pragma unroll 1
                for (int r0 = 0; r0 < 32; r0 += RG) {
                    const int64_t gmr = gm_base + r0;
                    if (gmr >= g.m && !emit) break;
                    float yv[RG];
                    if (YPRE) {
#pragma unroll
                        for (int i = 0; i < RG; ++i) yv[i] = yn[i];
                        if (want_y && r0 + RG < 32) {
                            const float *yp = g.y + (gmr + RG) * g.n + gn;
#pragma unroll
                            for (int i = 0; i < RG; ++i) yn[i] = (col_ok && gmr + RG + i < g.m) ? __ldg(yp + i * g.n) : 0.0f;
                        }
                    } else if (want_y) {
                        const float *yp = g.y + gmr * g.n + gn;
#pragma unroll
                        for (int i = 0; i < RG; ++i) yv[i] = (col_ok && gmr + i < g.m) ? __ldg(yp + i * g.n) : 0.0f;
                    }
                    float *cp = SPLIT ? g.c + (z_ * g.m + gmr) * g.n + gn : (g.c ? g.c + gmr * ldc + gn : nullptr);
                    const int64_t cstep = SPLIT ? g.n : ldc;
#pragma unroll
                    for (int i = 0; i < RG; ++i) {
                        const int r = r0 + i;
                        float v = 0.0f;
                        if (gmr + i < g.m && col_ok) {
                            v = stage[r * 33 + lane];
                            if (!SPLIT) {
                                if (g.epi == EPI_BIAS_ACT) v = mq_act_apply(v + bias_n, g.act, g.leak);
                                else if (g.epi == EPI_ACT_GRAD) v = mq_act_grad(yv[i], v, g.act, g.leak);
                                else v = g.alpha * v;
                            }
                            if (cp) cp[i * cstep] = v;
                        }
                        if (emit) stage[r * 33 + lane] = v;          // the epilogue's value, zero outside C
                    }
                }
                if (emit) {
                    // the same values as operand planes of the NEXT product (GmEmit, mq_gemm.cuh)
                    __syncwarp();
                    if (em.out_k) {            // C as a K-major operand: a lane converts 8 consecutive columns of one row
                        __nv_bfloat16 *out = (__nv_bfloat16 *)em.out_k;
                        const int r8 = lane >> 2, ch = lane & 3;
                        const int64_t gn0 = n0 + half * 64 + hc * 32 + ch * 8;
#pragma unroll 1
                        for (int it = 0; it < 4; ++it) {
                            const int r = it * 8 + r8;
                            const int64_t gm = m0 + quarter * 32 + r;
                            if (gm < g.m && gn0 < g.n) {
                                uint32_t w[4][P];
#pragma unroll
                                for (int j = 0; j < 4; ++j) gm_split_pair<P, F>(stage[r * 33 + ch * 8 + 2 * j], stage[r * 33 + ch * 8 + 2 * j + 1], w[j]);
#pragma unroll
                                for (int pl = 0; pl < P; ++pl)
                                    *reinterpret_cast<uint4 *>(out + ((int64_t)pl * em.k_rows_pad + gm) * em.k_kpad + gn0) =
                                        make_uint4(w[0][pl], w[1][pl], w[2][pl], w[3][pl]);
                            }
                        }
                    }
                    if (em.out_t && gn < g.n) {            // C^T: a lane owns a column and converts its 32 rows, 8 at a time
                        __nv_bfloat16 *out = (__nv_bfloat16 *)em.out_t;
                        const int64_t gm0 = m0 + quarter * 32;
#pragma unroll 1
                        for (int q = 0; q < 4; ++q) {
                            if (gm0 + q * 8 >= em.t_kpad) continue;
                            uint32_t w[4][P];
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                gm_split_pair<P, F>(stage[(q * 8 + 2 * j) * 33 + lane], stage[(q * 8 + 2 * j + 1) * 33 + lane], w[j]);
#pragma unroll
                            for (int pl = 0; pl < P; ++pl)
                                *reinterpret_cast<uint4 *>(out + ((int64_t)pl * em.t_rows_pad + gn) * em.t_kpad + gm0 + q * 8) =
                                    make_uint4(w[0][pl], w[1][pl], w[2][pl], w[3][pl]);
                        }
                    }
                }
                __syncwarp();                  // the next half (or tile) overwrites the staging tile
            }
        }
    }
#undef GM_TILE_DECODE
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"((uint32_t)Cfg::TMEM_COLS));
}

// ---- host side -----------------------------------------------------------------------------------------------------
typedef CUresult (*gm_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static gm_encode_fn gm_encoder() {
    static gm_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &st) == cudaSuccess &&
            st == cudaDriverEntryPointSuccess)
            fn = (gm_encode_fn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// planes stacked along the rows: [P * rows_pad][k_pad] bf16, box 64 (k) x 128 (rows), 128-byte swizzle
static int gm_make_map(CUtensorMap *map, void *base, int64_t rows_total, int64_t k_pad) {
    gm_encode_fn enc = gm_encoder();
    if (!enc) return 0;
    const cuuint64_t dims[2] = {(cuuint64_t)k_pad, (cuuint64_t)rows_total};
    const cuuint64_t strides[1] = {(cuuint64_t)k_pad * 2};
    const cuuint32_t box[2] = {GM_BK, 128};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// MN-major image: [P * k_rows_pad][mn_pad] bf16, box 64 (m or n: one 128-byte swizzle row) x 64 (k rows)
static int gm_make_map_mn(CUtensorMap *map, void *base, int64_t k_rows_total, int64_t mn_pad) {
    gm_encode_fn enc = gm_encoder();
    if (!enc) return 0;
    const cuuint64_t dims[2] = {(cuuint64_t)mn_pad, (cuuint64_t)k_rows_total};
    const cuuint64_t strides[1] = {(cuuint64_t)mn_pad * 2};
    const cuuint32_t box[2] = {64, GM_BK};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// bytes of one packed operand: the larger of its K-major image [rows_pad 128][k_pad 64] and its MN-major image
// [k rows pad 128][rows pad 128] (the launcher picks the form by the source's unit-stride axis)
static int64_t gm_pack_bytes(int planes, int64_t rows, int64_t k) {
    const int64_t kmaj = mq_round_up(rows, 128) * mq_round_up(k, GM_BK), mnmaj = mq_round_up(k, 128) * mq_round_up(rows, 128);
    return (int64_t)planes * (kmaj > mnmaj ? kmaj : mnmaj) * 2;
}

// workspace the packed-operand GEMM needs for an m x n x k product with `planes` planes (without split-K partials)
int64_t mq_gemm_tma_pack_bytes(int64_t m, int64_t n, int64_t k, int planes) {
    return mq_round_up(gm_pack_bytes(planes, m, k), 1024) + mq_round_up(gm_pack_bytes(planes, n, k), 1024) + 1024;
}

static int gm_mode_ok(int mode, int *planes, int *fmt) {
    *planes = mode & MQ_TC_PLANES_MASK; *fmt = (mode & MQ_TC_FP16) ? 1 : 0;      // mode = planes | MQ_TC_FP16
    if (*planes < 1 || *planes > 3) return 0;
    if (*fmt && *planes != 2) return 0;
    return 1;
}

// All jobs in one launch.  Fills in the block ranges; jobs with nothing to write are dropped.
int mq_gemm_tma_pack_multi(GmPackJob *jobs, int n_jobs, int mode, void *stream) {
    int planes, fmt;
    if (!gm_mode_ok(mode, &planes, &fmt)) return MQ_ERR_UNSUPPORTED;
    MQ_REQUIRE(n_jobs >= 0 && n_jobs <= GM_MAX_PACK_JOBS, MQ_ERR_INVALID, "mq_gemm_tma_pack_multi: %d jobs (max %d)", n_jobs, GM_MAX_PACK_JOBS);
    GmPackJobs js;
    memset(&js, 0, sizeof(js));
    unsigned total = 0;
    for (int i = 0; i < n_jobs; ++i) {
        GmPackJob jb = jobs[i];
        MQ_REQUIRE(jb.k_cover % 8 == 0 && jb.k_off % 8 == 0 && jb.k_pad % GM_BK == 0 && jb.rows_pad % 128 == 0 &&
                   jb.row_off + jb.rows_cover <= jb.rows_pad && jb.k_off + jb.k_cover <= jb.k_pad, MQ_ERR_SHAPE,
                   "mq_gemm_tma_pack_multi: job %d geometry", i);
        jb.block0 = total;
        jb.blocks = gm_job_blocks(jb);
        if (jb.blocks == 0) continue;
        total += jb.blocks;
        js.job[js.n++] = jb;
    }
    if (total == 0) return MQ_OK;
    if (planes == 3) { auto kp = pack_jobs_kernel<3, 0>; MQ_LAUNCH(kp, total, MQ_NT, 0, stream, js); }
    else if (planes == 2 && fmt) { auto kp = pack_jobs_kernel<2, 1>; MQ_LAUNCH(kp, total, MQ_NT, 0, stream, js); }
    else if (planes == 2) { auto kp = pack_jobs_kernel<2, 0>; MQ_LAUNCH(kp, total, MQ_NT, 0, stream, js); }
    else { auto kp = pack_jobs_kernel<1, 0>; MQ_LAUNCH(kp, total, MQ_NT, 0, stream, js); }
    return MQ_OK;
}

// a whole operand as one job: rows x k of the source into planes of [rows_pad][k_pad]
static GmPackJob gm_whole_job(const float *base, int64_t rs, int64_t cs, const int32_t *idx, int idx_on_k, int64_t rows,
                              int64_t k, int64_t ones_row, void *out) {
    GmPackJob jb;
    memset(&jb, 0, sizeof(jb));
    jb.base = base; jb.rs = rs; jb.cs = cs; jb.idx = idx; jb.idx_on_k = idx_on_k; jb.rows = rows; jb.k = k;
    jb.ones_row = ones_row;
    jb.rows_pad = jb.rows_cover = mq_round_up(rows, 128);
    jb.k_pad = jb.k_cover = mq_round_up(k, GM_BK);
    jb.out = out;
    return jb;
}

// The product on operands that are already packed: planes of A at `pa` ([planes][a_rows_pad][k_pad]), of B^T at `pb`
// ([planes][b_rows_pad][k_pad]).  g.a / g.b are not read; g.c points at the split-K partial buffer when splits > 1
// (g.k_per_split a multiple of 64; the planes `em` asks for are not written then).
int mq_gemm_tma_packed(const void *pa, int64_t a_rows_pad, const void *pb, int64_t b_rows_pad, int64_t k_pad,
                       const GemmArgs &g, const GmEmit &em, int mode, int64_t splits, void *stream) {
    return mq_gemm_tma_packed_mn(pa, a_rows_pad, 0, 0, pb, b_rows_pad, 0, 0, k_pad, g, em, mode, splits, stream);
}

int mq_gemm_tma_packed_mn(const void *pa, int64_t a_rows_pad, int a_mn, int64_t a_mn_pad, const void *pb, int64_t b_rows_pad,
                          int b_mn, int64_t b_mn_pad, int64_t k_pad, const GemmArgs &g, const GmEmit &em, int mode,
                          int64_t splits, void *stream) {
    int planes, fmt;
    if (!gm_mode_ok(mode, &planes, &fmt)) return MQ_ERR_UNSUPPORTED;
    if (g.m <= 0 || g.n <= 0 || g.k <= 0) return MQ_ERR_UNSUPPORTED;
    if (!gm_encoder()) return MQ_ERR_UNSUPPORTED;
    a_mn = a_mn ? 1 : 0; b_mn = b_mn ? 1 : 0;
    MQ_REQUIRE(a_rows_pad % 128 == 0 && b_rows_pad % 128 == 0 && k_pad % GM_BK == 0 && (((uintptr_t)pa | (uintptr_t)pb) & 127) == 0,
               MQ_ERR_SHAPE, "mq_gemm_tma: packed operand geometry");
    MQ_REQUIRE(a_mn ? (a_rows_pad >= g.k && a_mn_pad % 128 == 0 && a_mn_pad >= g.m) : (a_rows_pad >= g.m && k_pad >= g.k),
               MQ_ERR_SHAPE, "mq_gemm_tma: packed operand geometry (A)");
    MQ_REQUIRE(b_mn ? (b_rows_pad >= g.k && b_mn_pad % 128 == 0 && b_mn_pad >= g.n) : (b_rows_pad >= g.n && k_pad >= g.k),
               MQ_ERR_SHAPE, "mq_gemm_tma: packed operand geometry (B)");
    alignas(64) CUtensorMap map_a, map_b;
    const int ok_a = a_mn ? gm_make_map_mn(&map_a, (void *)pa, planes * a_rows_pad, a_mn_pad) : gm_make_map(&map_a, (void *)pa, planes * a_rows_pad, k_pad);
    const int ok_b = b_mn ? gm_make_map_mn(&map_b, (void *)pb, planes * b_rows_pad, b_mn_pad) : gm_make_map(&map_b, (void *)pb, planes * b_rows_pad, k_pad);
    if (!ok_a || !ok_b) {
        mq_set_error("mq_gemm_tma: cuTensorMapEncodeTiled failed");
        return MQ_ERR_CUDA;
    }
    const int64_t n_tiles = mq_ceil_div(g.n, GM_BN) * mq_ceil_div(g.m, GM_BM) * (splits > 1 ? splits : 1);
    // persistent: one CTA per SM at most (developer aid for A/B measurements: MQ_GEMM_ONE_TILE_PER_CTA=1 launches a CTA per tile)
    static const bool one_tile = getenv("MQ_GEMM_ONE_TILE_PER_CTA") != nullptr;
    const unsigned grid = (unsigned)((one_tile || n_tiles < mq_sm_count()) ? n_tiles : mq_sm_count());
    cudaError_t e = cudaSuccess;
    const bool ypre = splits <= 1 && g.epi == EPI_ACT_GRAD && g.k <= 256;      // short contraction with the act' epilogue
#define GM_GO(P_, F_, SP_)                                                                                                 \
    do {                                                                                                                   \
        auto kfn_ = (!SP_ && ypre) ? gemm_tma_kernel<P_, F_, false, true> : gemm_tma_kernel<P_, F_, SP_, false>;             \
        e = cudaFuncSetAttribute(kfn_, cudaFuncAttributeMaxDynamicSharedMemorySize, GmCfg<P_>::SMEM);                      \
        if (e == cudaSuccess)                                                                                              \
            e = mq_launch_pdl(kfn_, dim3(grid), dim3(GM_NT), (size_t)GmCfg<P_>::SMEM, stream, map_a, map_b, g, em, \
                              (int)a_rows_pad, (int)b_rows_pad, (int)(splits > 1 ? splits : 1), a_mn, b_mn);                    \
    } while (0)
#define GM_GO_P(SP_)                                              \
    do {                                                          \
        if (planes == 3) GM_GO(3, 0, SP_);                        \
        else if (planes == 2 && fmt) GM_GO(2, 1, SP_);            \
        else if (planes == 2) GM_GO(2, 0, SP_);                   \
        else GM_GO(1, 0, SP_);                                    \
    } while (0)
    if (splits > 1) GM_GO_P(true); else GM_GO_P(false);
#undef GM_GO_P
#undef GM_GO
    if (e != cudaSuccess) { mq_set_error("mq_gemm_tma: %s", cudaGetErrorString(e)); return MQ_ERR_CUDA; }
    return MQ_OK;
}

// Launches pack + GEMM.  g.c points at the split-K partial buffer when splits > 1 (g.k_per_split a multiple of 64).
// `pack` = workspace of mq_gemm_tma_pack_bytes() bytes.  MQ_ERR_UNSUPPORTED when tensor maps are not available.
int mq_gemm_tma_launch(const GemmArgs &g, int mode, int64_t splits, void *pack, void *stream) {
    int planes, fmt;
    if (!gm_mode_ok(mode, &planes, &fmt)) return MQ_ERR_UNSUPPORTED;
    if (g.m <= 0 || g.n <= 0 || g.k <= 0) return MQ_ERR_UNSUPPORTED;
    if (g.m >= (1LL << 31) - 256 || g.n >= (1LL << 31) - 256 || g.k >= (1LL << 31) - 256) return MQ_ERR_UNSUPPORTED;
    if (!gm_encoder()) return MQ_ERR_UNSUPPORTED;
    int64_t a_rows_pad = mq_round_up(g.m, 128), b_rows_pad = mq_round_up(g.n, 128);
    int64_t k_pad = mq_round_up(g.k, GM_BK);
    unsigned char *pa = (unsigned char *)(((uintptr_t)pack + 1023) & ~(uintptr_t)1023);
    unsigned char *pb = pa + mq_round_up(gm_pack_bytes(planes, g.m, g.k), 1024);
    if (g.a_planes) {
        // A is already packed (GemmArgs::a_planes): only B is packed here, at the start of the workspace (with the row length
        // of A's image when both are K-major: one tensor-map geometry per product)
        const bool mn = g.a_planes_mn != 0;
        MQ_REQUIRE(((uintptr_t)g.a_planes & 127) == 0 && g.a_planes_rows_pad % 128 == 0 && g.a_planes_pitch % 128 == 0 &&
                   (mn ? (g.a_planes_rows_pad >= g.k && g.a_planes_pitch >= g.m) : (g.a_planes_rows_pad >= g.m && g.a_planes_pitch >= k_pad)),
                   MQ_ERR_SHAPE, "mq_gemm_tma: geometry of the pre-packed operand");
        const int b_mn0 = g.b_cs == 1 && g.b_rs != 1;
        if (!mn) k_pad = g.a_planes_pitch;
        GmPackJob jb;
        int64_t b_mn_pad0 = 0;
        if (b_mn0) {
            jb = gm_whole_job(g.b, g.b_rs, 1, nullptr, 0, g.k, g.n, 0, pa);
            b_mn_pad0 = jb.k_pad = jb.k_cover = mq_round_up(g.n, 128);
            b_rows_pad = jb.rows_pad;
        } else {
            jb = gm_whole_job(g.b, g.b_cs, g.b_rs, nullptr, 0, g.n, g.k, 0, pa);
            jb.k_pad = k_pad;                                  // (columns beyond k_cover are never read: chunks end at k_cover)
            if ((int64_t)planes * jb.rows_pad * jb.k_pad * 2 > mq_gemm_tma_pack_bytes(g.m, g.n, g.k, planes) - 1024) return MQ_ERR_UNSUPPORTED;
        }
        int rc0 = mq_gemm_tma_pack_multi(&jb, 1, mode, stream);
        if (rc0 != MQ_OK) return rc0;
        GmEmit em0;
        memset(&em0, 0, sizeof(em0));
        return mq_gemm_tma_packed_mn(g.a_planes, g.a_planes_rows_pad, mn ? 1 : 0, mn ? g.a_planes_pitch : 0, pa, b_rows_pad, b_mn0, b_mn_pad0,
                                     k_pad, g, em0, mode, splits, stream);
    }
    // A(m, k) = a[row(m) * a_rs + k * a_cs]; B as rows n with K contiguous: B^T(n, k) = b[k * b_rs + n * b_cs]
    // An operand whose M (N) axis is the unit-stride axis of its source (x^T and g of dW = x^T g, W of y = x W) is packed
    // in the SOURCE orientation -- the plain, fully coalesced form: planes [k][m] -- and fed to the tensor pipe MN-major,
    // instead of being transposed through shared memory (the 65,536 x 784 minibatch of the MLP: 213 us at 1.9 TB/s).
    const int a_mn = g.a_rs == 1 && g.a_cs != 1, b_mn = g.b_cs == 1 && g.b_rs != 1;
    int64_t a_mn_pad = 0, b_mn_pad = 0;
    GmPackJob jobs[2];
    if (a_mn) {
        jobs[0] = gm_whole_job(g.a, g.a_cs, 1, g.idx, g.idx ? !g.idx_on_k : 0, g.k, g.m, 0, pa);     // rows = k, columns = m
        if (g.a_ones_row > 0) jobs[0].ones_col = g.a_ones_row;
        a_mn_pad = jobs[0].k_pad = jobs[0].k_cover = mq_round_up(g.m, 128);
        a_rows_pad = jobs[0].rows_pad;
    } else jobs[0] = gm_whole_job(g.a, g.a_rs, g.a_cs, g.idx, g.idx_on_k, g.m, g.k, g.a_ones_row, pa);
    if (b_mn) {
        jobs[1] = gm_whole_job(g.b, g.b_rs, 1, nullptr, 0, g.k, g.n, 0, pb);
        b_mn_pad = jobs[1].k_pad = jobs[1].k_cover = mq_round_up(g.n, 128);
        b_rows_pad = jobs[1].rows_pad;
    } else jobs[1] = gm_whole_job(g.b, g.b_cs, g.b_rs, nullptr, 0, g.n, g.k, 0, pb);
    int rc = mq_gemm_tma_pack_multi(jobs, 2, mode, stream);
    if (rc != MQ_OK) return rc;
    GmEmit em;
    memset(&em, 0, sizeof(em));
    return mq_gemm_tma_packed_mn(pa, a_rows_pad, a_mn, a_mn_pad, pb, b_rows_pad, b_mn, b_mn_pad, k_pad, g, em, mode, splits, stream);
}

#else
int64_t mq_gemm_tma_pack_bytes(int64_t, int64_t, int64_t, int) { return 0; }
int mq_gemm_tma_launch(const GemmArgs &, int, int64_t, void *, void *) { return MQ_ERR_UNSUPPORTED; }
int mq_gemm_tma_pack_multi(GmPackJob *, int, int, void *) { return MQ_ERR_UNSUPPORTED; }
int mq_gemm_tma_packed(const void *, int64_t, const void *, int64_t, int64_t, const GemmArgs &, const GmEmit &, int, int64_t, void *) { return MQ_ERR_UNSUPPORTED; }
int mq_gemm_tma_packed_mn(const void *, int64_t, int, int64_t, const void *, int64_t, int, int64_t, int64_t, const GemmArgs &, const GmEmit &, int, int64_t, void *) { return MQ_ERR_UNSUPPORTED; }
#endif
