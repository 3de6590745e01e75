// tcgen05 / TMEM / mbarrier PTX wrappers and the exact three-way bf16 split shared by the tensor-core kernels
// (mq_tc.cu: fused small-MLP gradient; mq_gemm_tc.cu: wide Affine layers).
#pragma once
#include "mq_common.cuh"
#ifndef MQ_EMU
#include <cuda_bf16.h>

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

// the same with fp16 operands (A / B format 0) when half != 0
__device__ __forceinline__ uint32_t tc_idesc_fmt(int m, int n, int a_mn, int b_mn, int half) {
    const uint32_t ab = half ? 0u : ((1u << 7) | (1u << 10));
    return (1u << 4) | ab | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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

__device__ __forceinline__ void tc_st16(uint32_t taddr, const float *v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr),
          "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// Gather loads that must be ISSUED where they are written (volatile: ptxas would otherwise sink them next to their
// first use, one epilogue later, and expose the full HBM latency there).
__device__ __forceinline__ float tc_ldg_f32(const float *p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int tc_ldg_s32(const int32_t *p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
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

// ---- operand views shared by the fused-MLP kernels (mq_tc.cu, mq_tc2.cu) -----------------------
// A view is the two descriptor words of its first K step plus the K-step advance (descriptor units of 16 bytes).
struct TcView { uint32_t lo, hi, k16; };

// forward weight image [k chunk][plane][n][8 k] from row `addr` on, K-major (N = plane*np + n, K = in)
__device__ __forceinline__ TcView tc_wf_k(uint32_t addr, int np) {
    return TcView{tc_desc_lo(addr, 3u * (uint32_t)np * 16u), tc_desc_hi(128u), (6u * (uint32_t)np * 16u) >> 4};
}
// backward weight image [plane][k chunk][n][8 k] from `addr` on, MN-major (N' = plane*kp + k, K' = n)
__device__ __forceinline__ TcView tc_wb_mn(uint32_t addr, int np) {
    return TcView{tc_desc_lo(addr, 128u), tc_desc_hi((uint32_t)np * 16u), 256u >> 4};
}

// branch-free tanh on raw MUFU ops: 1 - 2/(2^(2 log2(e) |x|) + 1) with the sign of x, odd polynomial below 0.1
// (same formulation and error, ~2e-7 absolute, as mq_tanh)
__device__ __forceinline__ float tc_tanh(float x) {
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(x) * 2.8853900817779268f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
    const float big = copysignf(fmaf(r, -2.0f, 1.0f), x);
    const float x2 = x * x;
    const float poly = x * fmaf(x2, fmaf(x2, fmaf(x2, -0.053968254f, 0.133333333f), -0.333333333f), 1.0f);
    return fabsf(x) < 0.1f ? poly : big;
}

template <int ACT>
__device__ __forceinline__ float tc_act(float z, float leak) {
    if (ACT == MQ_ACT_TANH) return tc_tanh(z);
    return z < 0.0f ? z * leak : z;
}

#endif  // MQ_EMU
