// Shared helpers for the mq_b200 kernels (sm_100a).
#pragma once

#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <math.h>

#include "../../include/mq_b200.h"

#ifdef MQ_EMU
// Developer-only host emulation (tools/emu): lets the kernel *logic* be checked
// on a machine without a GPU.  Never built into the product library.
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define MQ_SHARED_BYTES(name) extern __shared__ __align__(16) unsigned char name[]
#define MQ_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#endif

#define MQ_NT 256  // threads per CTA used by every kernel in this library

#ifndef __host__
#define __host__
#define __device__
#endif
static __host__ __device__ inline bool mq_head_gauss(int k) { return k == MQ_HEAD_GAUSS_LOGP || k == MQ_HEAD_GAUSS_PPO; }
static __host__ __device__ inline bool mq_head_discrete(int k) { return k == MQ_HEAD_DISCRETE_LOGP || k == MQ_HEAD_DISCRETE_PPO; }
static __host__ __device__ inline bool mq_head_ppo(int k) { return k == MQ_HEAD_GAUSS_PPO || k == MQ_HEAD_DISCRETE_PPO; }
static __host__ __device__ inline bool mq_head_prob(int k) { return mq_head_gauss(k) || mq_head_discrete(k); }

#ifndef MQ_EMU
// Programmatic dependent launch (PDL): a kernel launched with mq_launch_pdl may START while its predecessor in the
// stream is still running; it must execute mq_pdl_wait() before touching anything the predecessor writes (the wait
// returns when the predecessor grid has completed and its writes are visible).  mq_pdl_trigger() in the predecessor
// allows the dependent to be scheduled early.  Both are no-ops in a kernel launched the ordinary way.
__device__ __forceinline__ void mq_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void mq_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
static inline cudaError_t mq_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, void *stream,
                                        Args &&...args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// ---- host-side error plumbing ------------------------------------------------
void mq_set_error(const char *fmt, ...);

#define MQ_REQUIRE(cond, code, ...)        \
    do {                                   \
        if (!(cond)) {                     \
            mq_set_error(__VA_ARGS__);     \
            return (code);                 \
        }                                  \
    } while (0)

int mq_check_launch(const char *what);
#define MQ_CHECK_LAUNCH(what)                  \
    do {                                       \
        int mq_rc_ = mq_check_launch(what);    \
        if (mq_rc_ != MQ_OK) return mq_rc_;    \
    } while (0)

int mq_sm_count();
int64_t mq_smem_optin();

static inline int64_t mq_ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int64_t mq_round_up(int64_t a, int64_t b) { return mq_ceil_div(a, b) * b; }

// ---- device helpers -----------------------------------------------------------
template <typename T>
__device__ __forceinline__ T mq_warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA (MQ_NT threads); result valid in every thread.  `red` must
// hold MQ_NT/32 elements of T in shared memory.  Fixed order => deterministic.
template <typename T>
__device__ __forceinline__ T mq_block_sum(T v, T *red) {
    v = mq_warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    T s = red[0];
#pragma unroll
    for (int w = 1; w < MQ_NT / 32; ++w) s += red[w];
    return s;
}

// Branch-free tanh (so that independent evaluations interleave): odd polynomial for |x| < 0.1
// (truncation < 3e-11), 1 - 2/(exp(2|x|)+1) otherwise (absolute error ~2e-7).
__device__ __forceinline__ float mq_tanh(float x) {
#ifdef MQ_EMU
    return tanhf(x);
#else
    const float ax = fabsf(x);
    const float x2 = x * x;
    const float poly = x * fmaf(x2, fmaf(x2, fmaf(x2, -0.053968254f, 0.133333333f), -0.333333333f), 1.0f);
    const float t = __expf(2.0f * ax);
    const float big = copysignf(1.0f - __fdividef(2.0f, t + 1.0f), x);
    return ax < 0.1f ? poly : big;
#endif
}

__device__ __forceinline__ float mq_act_apply(float z, int act, float leak) {
    if (act == MQ_ACT_TANH) return mq_tanh(z);
    if (act == MQ_ACT_LRELU) return z < 0.0f ? z * leak : z;
    return z;
}

// derivative expressed through the saved OUTPUT y (backprop.py:53-60)
__device__ __forceinline__ float mq_act_grad(float y, float g, int act, float leak) {
    if (act == MQ_ACT_TANH) return g * (1.0f - y * y);
    if (act == MQ_ACT_LRELU) {
        const bool neg = (leak > 0.0f) ? (y < 0.0f) : (leak == 0.0f ? signbit(y) : (y > 0.0f));
        return neg ? g * leak : g;
    }
    return g;
}
