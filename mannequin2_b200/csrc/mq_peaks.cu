// Peak microbenchmarks used by bench.py for the roofline denominators that MEASURED_PEAKS.json does not carry
// ("fp32-FFMA and TF32 peaks to be measured by the harness"):
//   mq_bench_ffma    independent fp32 FFMA chains on every SM                      -> fp32 FFMA TFLOP/s
//   mq_bench_tensor  back-to-back tcgen05.mma M128 N256 from resident shared-memory operands, one CTA per SM,
//                    kind::f16 (bf16) or kind::tf32                               -> tensor-pipe issue peak
// They replace nothing in the reference (which has no device code); they only time the hardware.
#include "mq_tc.cuh"

#ifndef MQ_EMU
__global__ void __launch_bounds__(256) bench_ffma_kernel(int iters, float seed, float *out) {
    float a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = seed + (float)(threadIdx.x + k);
    const float m = 1.0000001f, c = 1e-7f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[k] = fmaf(a[k], m, c);
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k];
    if (s == 123.456f) out[0] = s;            // never true: keeps the chains alive
}

template <int TF32>
__global__ void __launch_bounds__(128, 1) bench_tensor_kernel(int n_mma) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t s_mbar;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    // operands: A 128 rows x 32 bytes of K, B 256 rows x 32 bytes of K, K-major SWIZZLE_NONE core matrices; the values do
    // not matter for the issue rate (zeros)
    for (int i = threadIdx.x; i < (128 + 256) * 32 / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&s_tmem)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&s_mbar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    tc_publish();
    const uint32_t tm = __shfl_sync(0xffffffffu, s_tmem, 0);
    if (warp == 0) {
        if (tc_elect_one()) {
            const uint32_t sb = tc_smem_u32(smem);
            // K-major: LBO = stride between the two 16-byte K halves (rows x 16 B), SBO = 128 (8-row groups)
            const uint32_t alo = tc_desc_lo(sb, 128u * 16u), blo = tc_desc_lo(sb + 128u * 32u, 256u * 16u);
            const uint32_t hi = tc_desc_hi(128u);
            uint32_t idesc = (1u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
            idesc |= TF32 ? ((2u << 7) | (2u << 10)) : ((1u << 7) | (1u << 10));
#pragma unroll 1
            for (int i = 0; i < n_mma; ++i) {
                if (TF32) {
                    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
                                 ::"r"(tm), "r"(alo), "r"(hi), "r"(blo), "r"(hi), "r"(idesc), "r"(1u) : "memory");
                } else {
                    tc_mma(tm, alo, hi, blo, hi, idesc, 1u);
                }
            }
            tc_commit(tc_smem_u32(&s_mbar));
        }
        __syncwarp();
    }
    tc_wait(tc_smem_u32(&s_mbar), 0);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u));
}

extern "C" int mq_bench_ffma(int32_t iters, float *scratch, double *flops_out, void *stream) {
    MQ_REQUIRE(iters > 0 && scratch && flops_out, MQ_ERR_INVALID, "mq_bench_ffma: arguments");
    const int blocks = mq_sm_count() * 8;
    MQ_LAUNCH(bench_ffma_kernel, blocks, 256, 0, stream, (int)iters, 0.5f, scratch);
    MQ_CHECK_LAUNCH("mq_bench_ffma");
    *flops_out = (double)blocks * 256.0 * 8.0 * 4.0 * 2.0 * (double)iters;
    return MQ_OK;
}

extern "C" int mq_bench_tensor(int32_t tf32, int32_t n_mma, double *flops_out, void *stream) {
    MQ_REQUIRE(n_mma > 0 && flops_out, MQ_ERR_INVALID, "mq_bench_tensor: arguments");
    const int blocks = mq_sm_count();
    const size_t smem = (128 + 256) * 32 + 1024;
    if (tf32) MQ_LAUNCH(bench_tensor_kernel<1>, blocks, 128, smem, stream, (int)n_mma);
    else MQ_LAUNCH(bench_tensor_kernel<0>, blocks, 128, smem, stream, (int)n_mma);
    MQ_CHECK_LAUNCH("mq_bench_tensor");
    *flops_out = (double)blocks * (double)n_mma * 2.0 * 128.0 * 256.0 * (tf32 ? 8.0 : 16.0);
    return MQ_OK;
}
#else
extern "C" int mq_bench_ffma(int32_t, float *, double *, void *) { return MQ_ERR_UNSUPPORTED; }
extern "C" int mq_bench_tensor(int32_t, int32_t, double *, void *) { return MQ_ERR_UNSUPPORTED; }
#endif
