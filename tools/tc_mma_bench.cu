// Developer aid: cycles per tcgen05.mma (kind::f16, SWIZZLE_NONE operands) for the shapes mq_tc.cu issues.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct Case { const char *name; uint32_t a_lo_lbo, a_sbo, a_k16, b_lo_lbo, b_sbo, b_k16, a_mn, b_mn, m, n, ksteps, reps; };
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mma(uint32_t d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc) : "memory");
}
__global__ void __launch_bounds__(128) bench(Case c, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_base;
    __shared__ __align__(8) uint64_t mbar;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((uint32_t *)smem)[i] = 0x3c003c00u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t sb = smem_u32(smem);
    long long t0 = 0, t1 = 0, t2 = 0;
    if (warp == 0) {
        if (elect_one()) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (c.a_mn << 15) | (c.b_mn << 16) | ((c.n >> 3) << 17) | ((c.m >> 4) << 24);
            const uint32_t ahi = ((c.a_sbo >> 4) & 0x3FFF) | (1u << 14), bhi = ((c.b_sbo >> 4) & 0x3FFF) | (1u << 14);
            t0 = clock64();
            for (uint32_t r = 0; r < c.reps; ++r) {
                uint32_t alo = ((sb >> 4) & 0x3FFF) | (((c.a_lo_lbo >> 4) & 0x3FFF) << 16);
                uint32_t blo = (((sb + 96 * 1024) >> 4) & 0x3FFF) | (((c.b_lo_lbo >> 4) & 0x3FFF) << 16);
                for (uint32_t k = 0; k < c.ksteps; ++k) { mma(tm, alo, ahi, blo, bhi, idesc, 1u); alo += c.a_k16; blo += c.b_k16; }
            }
            t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
            t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        }
        __syncwarp();
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u));
}
int main() {
    const uint32_t CH = 2048, RG = 128;
    Case cs[] = {
        // name, a_lbo, a_sbo, a_k16, b_lbo, b_sbo, b_k16, a_mn, b_mn, m, n, ksteps, reps
        {"fwd  K/K   M128 N128 (wf image)", CH, RG, 2 * CH / 16, 3 * 64 * 16, RG, 6 * 64 * 16 / 16, 0, 0, 128, 128, 4, 64},
        {"fwd  K/K   M128 N64", CH, RG, 2 * CH / 16, 3 * 64 * 16, RG, 6 * 64 * 16 / 16, 0, 0, 128, 64, 4, 64},
        {"fwd  K/K   M128 N32 (last)", CH, RG, 2 * CH / 16, 3 * 16 * 16, RG, 6 * 16 * 16 / 16, 0, 0, 128, 32, 4, 64},
        {"fwd  K/K   M128 N16 (last)", CH, RG, 2 * CH / 16, 3 * 16 * 16, RG, 6 * 16 * 16 / 16, 0, 0, 128, 16, 4, 64},
        {"dA   K/MN  M128 N128", CH, RG, 2 * CH / 16, RG, 64 * 16, 256 / 16, 0, 1, 128, 128, 4, 64},
        {"dA   K/MN  M128 N64", CH, RG, 2 * CH / 16, RG, 64 * 16, 256 / 16, 0, 1, 128, 64, 4, 64},
        {"dW   MN/MN M64  N136", RG, CH, 256 / 16, RG, CH, 256 / 16, 1, 1, 64, 136, 8, 32},
        {"dW   MN/MN M64  N72", RG, CH, 256 / 16, RG, CH, 256 / 16, 1, 1, 64, 72, 8, 32},
        {"dW   MN/MN M64  N64", RG, CH, 256 / 16, RG, CH, 256 / 16, 1, 1, 64, 64, 8, 32},
        {"dW   MN/MN M64  N40", RG, CH, 256 / 16, RG, CH, 256 / 16, 1, 1, 64, 40, 8, 32},
        {"dW   MN/MN M64  N32", RG, CH, 256 / 16, RG, CH, 256 / 16, 1, 1, 64, 32, 8, 32},
        {"dW   MN/MN M64  N24", RG, CH, 256 / 16, RG, CH, 256 / 16, 1, 1, 64, 24, 8, 32},
        {"dW   MN/MN M64  N16", RG, CH, 256 / 16, RG, CH, 256 / 16, 1, 1, 64, 16, 8, 32},
        {"dW   MN/MN M128 N136", RG, CH, 256 / 16, RG, CH, 256 / 16, 1, 1, 128, 144, 8, 32},
        {"dW   MN/MN M64  N200", RG, CH, 256 / 16, RG, CH, 256 / 16, 1, 1, 64, 200, 8, 32},
    };
    long long *d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (auto &c : cs) {
        bench<<<1, 128, 200 * 1024>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        const double n = (double)c.ksteps * c.reps;
        printf("%-40s issue %.1f cyc/mma, issue+drain %.1f cyc/mma  (floor %d) %s\n", c.name, h[0] / n, h[1] / n,
               (int)((c.m > 128 ? c.m : 128) * c.n / 256), e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
