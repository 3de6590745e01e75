// cuda_emu.h -- DEVELOPER-ONLY host emulation of the small CUDA subset the
// mq_b200 FFMA kernels use.  Purpose: check kernel *logic* (indexing, tiling,
// reductions) in the build container, which has no GPU, before spending B200
// time.  One std::thread per CUDA thread, CTAs run one after another.
//
// This is NOT a product path: mannequin2_b200/ never loads an emulated build,
// tests/ and bench.py never use it, and it is not a CPU fallback.  See
// tools/emu/README.md.
#pragma once

#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))
#define __shared__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3_emu { unsigned x, y, z; };
typedef void *cudaStream_t;
typedef int cudaError_t;
#define cudaSuccess 0

struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct __attribute__((aligned(8))) float2 { float x, y; };
struct __attribute__((aligned(16))) double2 { double x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }

struct EmuWarp {
    std::barrier<> bar{32};
    uint64_t slot[32];
};
struct EmuBlock {
    std::barrier<> bar;
    std::vector<std::unique_ptr<EmuWarp>> warps;
    unsigned char *smem;
    explicit EmuBlock(int n) : bar(n) {
        for (int i = 0; i < (n + 31) / 32; ++i) warps.emplace_back(new EmuWarp());
    }
};

struct EmuTls {
    EmuBlock *blk = nullptr;
    int lane = 0, warp = 0;
};
inline thread_local EmuTls emu_tls;
inline thread_local uint3_emu threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

static inline void __syncthreads() { emu_tls.blk->bar.arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
    emu_tls.blk->warps[emu_tls.warp]->bar.arrive_and_wait();
}
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <typename T>
static inline T emu_shfl(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    EmuWarp &w = *emu_tls.blk->warps[emu_tls.warp];
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    w.slot[emu_tls.lane] = bits;
    w.bar.arrive_and_wait();
    uint64_t got = w.slot[src_lane & 31];
    w.bar.arrive_and_wait();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_shfl(v, emu_tls.lane ^ m); }
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, int d) {
    int s = emu_tls.lane + d;
    return emu_shfl(v, s < 32 ? s : emu_tls.lane);
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, int d) {
    int s = emu_tls.lane - d;
    return emu_shfl(v, s >= 0 ? s : emu_tls.lane);
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int s) { return emu_shfl(v, s); }

template <typename T>
static inline T emu_atomic_add(T *p, T v) {
    T old = *p, des;
    do { des = old + v; } while (!__atomic_compare_exchange(p, &old, &des, false, __ATOMIC_SEQ_CST,
                                                            __ATOMIC_SEQ_CST));
    return old;
}
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
static inline float atomicAdd(float *p, float v) { return emu_atomic_add(p, v); }
static inline double atomicAdd(double *p, double v) { return emu_atomic_add(p, v); }

static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
template <typename T>
static inline T __ldg(const T *p) { return *p; }
using std::signbit;
using std::min;
using std::max;

template <typename F>
static inline void emu_launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
    const int nt = (int)(block.x * block.y * block.z);
    void *smem = nullptr;
    if (posix_memalign(&smem, 128, smem_bytes + 128) != 0) abort();
    for (unsigned bz = 0; bz < grid.z; ++bz)
        for (unsigned by = 0; by < grid.y; ++by)
            for (unsigned bx = 0; bx < grid.x; ++bx) {
                EmuBlock blk(nt);
                blk.smem = (unsigned char *)smem;
                std::vector<std::thread> th;
                th.reserve(nt);
                for (int t = 0; t < nt; ++t)
                    th.emplace_back([&, t]() {
                        emu_tls.blk = &blk;
                        emu_tls.lane = t & 31;
                        emu_tls.warp = t >> 5;
                        threadIdx = uint3_emu{t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
                        blockIdx = uint3_emu{bx, by, bz};
                        blockDim = block;
                        gridDim = grid;
                        body();
                    });
                for (auto &x : th) x.join();
            }
    free(smem);
}

#define MQ_SHARED_BYTES(name) unsigned char *name = emu_tls.blk->smem
#define MQ_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu_launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kern(__VA_ARGS__); })
