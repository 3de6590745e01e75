// Developer aid: variants of the single-pass GAE scan (mq_scan.cu) timed side by side on one box.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/_build/gae_variants tools/gae_variants.cu
//   tools/_build/gae_variants [log2 n = 26]
// reg<NT, ITEMS, MINB>     : the library's register kernel (one chunk per CTA, direct 128-bit loads / stores)
// tma<NT, ITEMS, STAGES>   : persistent CTAs, chunks by ticket, bulk-copy (TMA) ring in shared memory, outputs written in
//                            place in the stage and bulk-stored
// Every variant is checked against an fp64 host recurrence.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Lin { float a, b; };
__device__ __forceinline__ Lin lin_then(const Lin &f, const Lin &s) { Lin r; r.a = f.a * s.a; r.b = s.b + s.a * f.b; return r; }
struct Args { const float *rew, *val; const uint8_t *done; float *adv, *ret; float g, gl; };
struct __align__(16) Tile { float a; int ta; float y; int ty; };

__device__ __forceinline__ void st64(void *p, float v, int tag) {
    const unsigned long long w = ((unsigned long long)(unsigned)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ld64(const void *p) {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ void tile_store(Tile *t, float a, float b, float y, int status) {
    if (status == 1) st64(&t->a, a, 1);
    st64(&t->y, status == 1 ? b : y, status);
}
__device__ __forceinline__ void tile_load(const Tile *t, float &a, float &b, int &status) {
    const unsigned long long w1 = ld64(&t->a), w2 = ld64(&t->y);
    const int s1 = (int)(w1 >> 32), s2 = (int)(w2 >> 32);
    a = __uint_as_float((unsigned)w1);
    b = __uint_as_float((unsigned)w2);
    status = (s2 == 2) ? 2 : ((s2 == 1 && s1 == 1) ? 1 : 0);
}
// warp 0: publish, look back, return the value entering chunk c
__device__ __forceinline__ float publish_and_lookback(Tile *__restrict__ tiles, int64_t c, const Lin &total) {
    const int lane = threadIdx.x;
    const bool closed = total.a == 0.0f;
    if (lane == 0) tile_store(&tiles[c], total.a, total.b, total.b, closed ? 2 : 1);
    Lin acc; acc.a = 1.0f; acc.b = 0.0f;
    bool done = false;
    int64_t base = c - 1;
    while (!done) {
        const int64_t q = base - lane;
        float qa = 0.0f, qb = 0.0f;
        if (q >= 0) {
            int st; float ta, tb;
            do { tile_load(&tiles[q], ta, tb, st); } while (st == 0);
            if (st == 2) qb = tb; else { qa = ta; qb = tb; }
        }
        const unsigned term = __ballot_sync(0xffffffffu, qa == 0.0f);
        const int stop = term ? __ffs(term) - 1 : 31;
        Lin win; win.a = 1.0f; win.b = 0.0f;
        for (int l = stop; l >= 0; --l) {
            Lin e;
            e.a = __shfl_sync(0xffffffffu, qa, l);
            e.b = __shfl_sync(0xffffffffu, qb, l);
            win = lin_then(win, e);
        }
        acc = lin_then(win, acc);
        done = term != 0u;
        base -= 32;
    }
    if (lane == 0 && !closed) tile_store(&tiles[c], total.a, total.b, total.b + total.a * acc.b, 2);
    return acc.b;
}
__device__ __forceinline__ Lin warp_incl_scan(Lin v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Lin p; p.a = __shfl_up_sync(0xffffffffu, v.a, o); p.b = __shfl_up_sync(0xffffffffu, v.b, o);
        if (lane >= o) v = lin_then(p, v);
    }
    return v;
}

// ------------------------------------------------------------------------------------------------------------------
template <int NT, int ITEMS, int MINB>
__global__ void __launch_bounds__(NT, MINB) gae_reg_kernel(Args el, int64_t n, Tile *__restrict__ tiles) {
    constexpr int CH = NT * ITEMS;
    __shared__ Lin warp_tot[NT / 32];
    __shared__ float s_yin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c = blockIdx.x;
    const int64_t hi = n - 1 - c * CH;
    const int64_t r1 = hi - (int64_t)threadIdx.x * ITEMS, r0 = r1 - (ITEMS - 1);
    float rw[ITEMS], vl[ITEMS + 1];
    unsigned dn = 0;
    const bool live = r0 >= 0;
    if (live) {
#pragma unroll
        for (int q = 0; q < ITEMS / 4; ++q) {
            const float4 a = *reinterpret_cast<const float4 *>(el.rew + r0 + 4 * q);
            const float4 b = *reinterpret_cast<const float4 *>(el.val + r0 + 4 * q);
            const unsigned d = *reinterpret_cast<const unsigned *>(el.done + r0 + 4 * q);
            rw[4 * q] = a.x; rw[4 * q + 1] = a.y; rw[4 * q + 2] = a.z; rw[4 * q + 3] = a.w;
            vl[4 * q] = b.x; vl[4 * q + 1] = b.y; vl[4 * q + 2] = b.z; vl[4 * q + 3] = b.w;
#pragma unroll
            for (int k = 0; k < 4; ++k) dn |= ((d >> (8 * k)) & 0xFFu) ? (1u << (4 * q + k)) : 0u;
        }
        vl[ITEMS] = el.val[r1 + 1];
    } else {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) { rw[k] = 0.0f; vl[k] = 0.0f; }
        vl[ITEMS] = 0.0f;
    }
    float eb[ITEMS];
    Lin local; local.a = 1.0f; local.b = 0.0f;
#pragma unroll
    for (int k = ITEMS - 1; k >= 0; --k) {
        Lin e;
        if (dn & (1u << k)) { e.a = 0.0f; e.b = rw[k] - vl[k]; }
        else { e.a = el.gl; e.b = rw[k] + el.g * vl[k + 1] - vl[k]; }
        eb[k] = e.b;
        if (live) local = lin_then(local, e);
    }
    const Lin incl = warp_incl_scan(local);
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    Lin wp; wp.a = 1.0f; wp.b = 0.0f;
    Lin all = wp;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
        if (w == warp) wp = all;
        all = lin_then(all, warp_tot[w]);
    }
    Lin prev; prev.a = __shfl_up_sync(0xffffffffu, incl.a, 1); prev.b = __shfl_up_sync(0xffffffffu, incl.b, 1);
    const Lin excl = (lane == 0) ? wp : lin_then(wp, prev);
    if (threadIdx.x < 32) {
        const float yin = publish_and_lookback(tiles, c, all);
        if (threadIdx.x == 0) s_yin = yin;
    }
    __syncthreads();
    if (!live) return;
    float y = excl.b + excl.a * s_yin;
    float av[ITEMS], rt[ITEMS];
#pragma unroll
    for (int k = ITEMS - 1; k >= 0; --k) {
        y = eb[k] + ((dn & (1u << k)) ? 0.0f : el.gl) * y;
        av[k] = y;
        rt[k] = y + vl[k];
    }
#pragma unroll
    for (int q = 0; q < ITEMS / 4; ++q) {
        *reinterpret_cast<float4 *>(el.adv + r0 + 4 * q) = make_float4(av[4 * q], av[4 * q + 1], av[4 * q + 2], av[4 * q + 3]);
        *reinterpret_cast<float4 *>(el.ret + r0 + 4 * q) = make_float4(rt[4 * q], rt[4 * q + 1], rt[4 * q + 2], rt[4 * q + 3]);
    }
}

// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t *b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
    } while (!ok);
}

template <int NT, int ITEMS, int STAGES>
struct TmaCfg {
    static constexpr int CH = NT * ITEMS;
    static constexpr int REW = 0, VAL = CH * 4, DONE = VAL + (CH + 4) * 4, STAGE = DONE + CH;   // bytes, multiples of 16
    static constexpr int SMEM = STAGE * STAGES;
};

// Requires n % 16 == 0 and 16-byte aligned arrays.
template <int NT, int ITEMS, int STAGES>
__global__ void __launch_bounds__(NT) gae_tma_kernel(Args el, int64_t n, int64_t nchunks, Tile *__restrict__ tiles,
                                                     unsigned int *__restrict__ ticket) {
    using C = TmaCfg<NT, ITEMS, STAGES>;
    constexpr int CH = C::CH;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[STAGES];
    __shared__ long long s_chunk[STAGES];
    __shared__ Lin warp_tot[NT / 32];
    __shared__ float s_yin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // producer (thread 0): take the next chunk by ticket and start its three bulk loads into stage s
    auto arm = [&](int s) {
        const int64_t c = (int64_t)atomicAdd(ticket, 1u);
        if (c >= nchunks) { s_chunk[s] = -1; mbar_arrive(&full[s]); return; }
        s_chunk[s] = c;
        unsigned char *st = smem + (size_t)s * C::STAGE;
        const int64_t hi = n - 1 - c * CH, lo_full = hi - (CH - 1);
        const int64_t lo = lo_full > 0 ? lo_full : 0;
        const int off = (int)(lo - lo_full), cnt = (int)(hi - lo + 1);
        const unsigned vbytes = (unsigned)(cnt * 4 + (c > 0 ? 16 : 0));
        if (c == 0) reinterpret_cast<float *>(st + C::VAL)[CH] = el.val[n];      // the one value past the last row
        mbar_expect(&full[s], (unsigned)(cnt * 4) + vbytes + (unsigned)cnt);
        bulk_g2s(st + C::REW + off * 4, el.rew + lo, (unsigned)(cnt * 4), &full[s]);
        bulk_g2s(st + C::VAL + off * 4, el.val + lo, vbytes, &full[s]);
        bulk_g2s(st + C::DONE + off, el.done + lo, (unsigned)cnt, &full[s]);
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
        for (int s = 0; s < STAGES - 1; ++s) arm(s);
    }
    __syncthreads();

    for (int k = 0;; ++k) {
        const int s = k % STAGES;
        mbar_wait(&full[s], (unsigned)((k / STAGES) & 1));
        const int64_t c = s_chunk[s];
        if (c < 0) break;
        unsigned char *st = smem + (size_t)s * C::STAGE;
        float *srew = reinterpret_cast<float *>(st + C::REW), *sval = reinterpret_cast<float *>(st + C::VAL);
        const uint8_t *sdone = st + C::DONE;
        const int64_t hi = n - 1 - c * CH, lo_full = hi - (CH - 1);
        const int i0 = CH - ((int)threadIdx.x + 1) * ITEMS;      // this thread's ITEMS rows in the stage
        const bool live = lo_full + i0 >= 0;
        float rw[ITEMS], vl[ITEMS + 1];
        unsigned dn = 0;
        if (live) {
#pragma unroll
            for (int q = 0; q < ITEMS / 4; ++q) {
                const float4 a = *reinterpret_cast<const float4 *>(srew + i0 + 4 * q);
                const float4 b = *reinterpret_cast<const float4 *>(sval + i0 + 4 * q);
                const unsigned d = *reinterpret_cast<const unsigned *>(sdone + i0 + 4 * q);
                rw[4 * q] = a.x; rw[4 * q + 1] = a.y; rw[4 * q + 2] = a.z; rw[4 * q + 3] = a.w;
                vl[4 * q] = b.x; vl[4 * q + 1] = b.y; vl[4 * q + 2] = b.z; vl[4 * q + 3] = b.w;
#pragma unroll
                for (int j = 0; j < 4; ++j) dn |= ((d >> (8 * j)) & 0xFFu) ? (1u << (4 * q + j)) : 0u;
            }
            vl[ITEMS] = sval[i0 + ITEMS];
        } else {
#pragma unroll
            for (int j = 0; j < ITEMS; ++j) { rw[j] = 0.0f; vl[j] = 0.0f; }
            vl[ITEMS] = 0.0f;
        }
        float eb[ITEMS];
        Lin local; local.a = 1.0f; local.b = 0.0f;
#pragma unroll
        for (int j = ITEMS - 1; j >= 0; --j) {
            Lin e;
            if (dn & (1u << j)) { e.a = 0.0f; e.b = rw[j] - vl[j]; }
            else { e.a = el.gl; e.b = rw[j] + el.g * vl[j + 1] - vl[j]; }
            eb[j] = e.b;
            if (live) local = lin_then(local, e);
        }
        const Lin incl = warp_incl_scan(local);
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();                                           // (A) warp totals visible; every read of the stage done
        Lin wp; wp.a = 1.0f; wp.b = 0.0f;
        Lin all = wp;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
            if (w == warp) wp = all;
            all = lin_then(all, warp_tot[w]);
        }
        Lin prev; prev.a = __shfl_up_sync(0xffffffffu, incl.a, 1); prev.b = __shfl_up_sync(0xffffffffu, incl.b, 1);
        const Lin excl = (lane == 0) ? wp : lin_then(wp, prev);
        if (threadIdx.x < 32) {
            const float yin = publish_and_lookback(tiles, c, all);
            if (threadIdx.x == 0) s_yin = yin;
        }
        __syncthreads();                                           // (B) entering value visible
        if (live) {
            float y = excl.b + excl.a * s_yin;
            float av[ITEMS], rt[ITEMS];
#pragma unroll
            for (int j = ITEMS - 1; j >= 0; --j) {
                y = eb[j] + ((dn & (1u << j)) ? 0.0f : el.gl) * y;
                av[j] = y;
                rt[j] = y + vl[j];
            }
#pragma unroll
            for (int q = 0; q < ITEMS / 4; ++q) {                  // in place: advantages over rewards, returns over values
                *reinterpret_cast<float4 *>(srew + i0 + 4 * q) = make_float4(av[4 * q], av[4 * q + 1], av[4 * q + 2], av[4 * q + 3]);
                *reinterpret_cast<float4 *>(sval + i0 + 4 * q) = make_float4(rt[4 * q], rt[4 * q + 1], rt[4 * q + 2], rt[4 * q + 3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();                                           // (C) outputs of the chunk complete in the stage
        if (threadIdx.x == 0) {
            const int64_t lo = lo_full > 0 ? lo_full : 0;
            const int off = (int)(lo - lo_full), cnt = (int)(hi - lo + 1);
            bulk_s2g(el.adv + lo, srew + off, (unsigned)(cnt * 4));
            bulk_s2g(el.ret + lo, sval + off, (unsigned)(cnt * 4));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the previous chunk's stage has been read out
            arm((k + STAGES - 1) % STAGES);
        }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------------------------
// hyb<NT, ITEMS, MINB, NB>: one CTA = (1 + NB) consecutive chunks of NT * ITEMS rows.  Chunk 0 lives in registers (direct
// 128-bit loads), chunks 1..NB are bulk-copied (TMA) into shared memory at kernel start: the rows a CTA keeps in flight
// are no longer bounded by the register file, and there is one global look-back per (1 + NB) chunks.
template <int NT, int ITEMS, int NB>
struct HybCfg {
    static constexpr int CH = NT * ITEMS;
    static constexpr int REW = 0, VAL = CH * 4, DONE = VAL + (CH + 4) * 4, STAGE = DONE + CH;
    static constexpr int SMEM = STAGE * NB;
    static constexpr int ROWS = CH * (1 + NB);
};

template <int ITEMS>
__device__ __forceinline__ void gae_rows_from_smem(const float *srew, const float *sval, const uint8_t *sdone, int i0,
                                                   float (&rw)[ITEMS], float (&vl)[ITEMS + 1], unsigned &dn) {
    dn = 0;
#pragma unroll
    for (int q = 0; q < ITEMS / 4; ++q) {
        const float4 a = *reinterpret_cast<const float4 *>(srew + i0 + 4 * q);
        const float4 b = *reinterpret_cast<const float4 *>(sval + i0 + 4 * q);
        const unsigned d = *reinterpret_cast<const unsigned *>(sdone + i0 + 4 * q);
        rw[4 * q] = a.x; rw[4 * q + 1] = a.y; rw[4 * q + 2] = a.z; rw[4 * q + 3] = a.w;
        vl[4 * q] = b.x; vl[4 * q + 1] = b.y; vl[4 * q + 2] = b.z; vl[4 * q + 3] = b.w;
#pragma unroll
        for (int j = 0; j < 4; ++j) dn |= ((d >> (8 * j)) & 0xFFu) ? (1u << (4 * q + j)) : 0u;
    }
    vl[ITEMS] = sval[i0 + ITEMS];
}

template <int NT, int ITEMS, int MINB, int NB>
__global__ void __launch_bounds__(NT, MINB) gae_hyb_kernel(Args el, int64_t n, Tile *__restrict__ tiles) {
    using C = HybCfg<NT, ITEMS, NB>;
    constexpr int CH = C::CH;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t full[NB];
    __shared__ Lin warp_tot[1 + NB][NT / 32];
    __shared__ float s_yin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c = blockIdx.x;
    const int64_t hi = n - 1 - c * C::ROWS;                       // newest row of the CTA
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < NB; ++j) mbar_init(&full[j], 1);
        asm volatile("fence.mbarrier_init.release.cluster;");
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            unsigned char *st = smem + (size_t)j * C::STAGE;
            const int64_t hj = hi - (int64_t)(j + 1) * CH, lo_full = hj - (CH - 1);
            const int64_t lo = lo_full > 0 ? lo_full : 0;
            const int cnt = (int)(hj - lo + 1), off = (int)(lo - lo_full);
            if (cnt > 0) {
                mbar_expect(&full[j], (unsigned)(cnt * 9 + 16));
                bulk_g2s(st + C::REW + off * 4, el.rew + lo, (unsigned)(cnt * 4), &full[j]);
                bulk_g2s(st + C::VAL + off * 4, el.val + lo, (unsigned)(cnt * 4 + 16), &full[j]);
                bulk_g2s(st + C::DONE + off, el.done + lo, (unsigned)cnt, &full[j]);
            } else mbar_arrive(&full[j]);
        }
    }
    // ---- chunk 0: registers ------------------------------------------------------------------------
    const int64_t r1 = hi - (int64_t)threadIdx.x * ITEMS, r0 = r1 - (ITEMS - 1);
    float rw[ITEMS], vl[ITEMS + 1];
    unsigned dn = 0;
    const bool live = r0 >= 0;
    if (live) {
#pragma unroll
        for (int q = 0; q < ITEMS / 4; ++q) {
            const float4 a = *reinterpret_cast<const float4 *>(el.rew + r0 + 4 * q);
            const float4 b = *reinterpret_cast<const float4 *>(el.val + r0 + 4 * q);
            const unsigned d = *reinterpret_cast<const unsigned *>(el.done + r0 + 4 * q);
            rw[4 * q] = a.x; rw[4 * q + 1] = a.y; rw[4 * q + 2] = a.z; rw[4 * q + 3] = a.w;
            vl[4 * q] = b.x; vl[4 * q + 1] = b.y; vl[4 * q + 2] = b.z; vl[4 * q + 3] = b.w;
#pragma unroll
            for (int k = 0; k < 4; ++k) dn |= ((d >> (8 * k)) & 0xFFu) ? (1u << (4 * q + k)) : 0u;
        }
        vl[ITEMS] = el.val[r1 + 1];
    } else {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) { rw[k] = 0.0f; vl[k] = 0.0f; }
        vl[ITEMS] = 0.0f;
    }
    __syncthreads();                                               // barriers initialised before anyone waits on them
    float eb[ITEMS];
    Lin local; local.a = 1.0f; local.b = 0.0f;
#pragma unroll
    for (int k = ITEMS - 1; k >= 0; --k) {
        Lin e;
        if (dn & (1u << k)) { e.a = 0.0f; e.b = rw[k] - vl[k]; }
        else { e.a = el.gl; e.b = rw[k] + el.g * vl[k + 1] - vl[k]; }
        eb[k] = e.b;
        if (live) local = lin_then(local, e);
    }
    Lin incl[1 + NB];
    incl[0] = warp_incl_scan(local);
    // ---- chunks 1..NB: shared memory, totals only ------------------------------------------------------
    const int i0 = CH - ((int)threadIdx.x + 1) * ITEMS;
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        const unsigned char *st = smem + (size_t)j * C::STAGE;
        const int64_t lo_full = hi - (int64_t)(j + 2) * CH + 1;
        Lin lj; lj.a = 1.0f; lj.b = 0.0f;
        mbar_wait(&full[j], 0u);
        if (lo_full + i0 >= 0) {
            float rj[ITEMS], vj[ITEMS + 1];
            unsigned dj;
            gae_rows_from_smem<ITEMS>(reinterpret_cast<const float *>(st + C::REW), reinterpret_cast<const float *>(st + C::VAL),
                                      st + C::DONE, i0, rj, vj, dj);
#pragma unroll
            for (int k = ITEMS - 1; k >= 0; --k) {
                Lin e;
                if (dj & (1u << k)) { e.a = 0.0f; e.b = rj[k] - vj[k]; }
                else { e.a = el.gl; e.b = rj[k] + el.g * vj[k + 1] - vj[k]; }
                lj = lin_then(lj, e);
            }
        }
        incl[1 + j] = warp_incl_scan(lj);
    }
    if (lane == 31) {
#pragma unroll
        for (int j = 0; j <= NB; ++j) warp_tot[j][warp] = incl[j];
    }
    __syncthreads();
    Lin excl[1 + NB], tot[1 + NB];
    Lin all; all.a = 1.0f; all.b = 0.0f;
#pragma unroll
    for (int j = 0; j <= NB; ++j) {
        Lin wp; wp.a = 1.0f; wp.b = 0.0f;
        Lin t = wp;
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) {
            if (w == warp) wp = t;
            t = lin_then(t, warp_tot[j][w]);
        }
        Lin prev; prev.a = __shfl_up_sync(0xffffffffu, incl[j].a, 1); prev.b = __shfl_up_sync(0xffffffffu, incl[j].b, 1);
        excl[j] = (lane == 0) ? wp : lin_then(wp, prev);
        tot[j] = t;
        all = lin_then(all, t);
    }
    if (threadIdx.x < 32) {
        const float yin = publish_and_lookback(tiles, c, all);
        if (threadIdx.x == 0) s_yin = yin;
    }
    __syncthreads();
    float yin = s_yin;
    if (live) {
        float y = excl[0].b + excl[0].a * yin;
        float av[ITEMS], rt[ITEMS];
#pragma unroll
        for (int k = ITEMS - 1; k >= 0; --k) {
            y = eb[k] + ((dn & (1u << k)) ? 0.0f : el.gl) * y;
            av[k] = y;
            rt[k] = y + vl[k];
        }
#pragma unroll
        for (int q = 0; q < ITEMS / 4; ++q) {
            *reinterpret_cast<float4 *>(el.adv + r0 + 4 * q) = make_float4(av[4 * q], av[4 * q + 1], av[4 * q + 2], av[4 * q + 3]);
            *reinterpret_cast<float4 *>(el.ret + r0 + 4 * q) = make_float4(rt[4 * q], rt[4 * q + 1], rt[4 * q + 2], rt[4 * q + 3]);
        }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        yin = tot[j].b + tot[j].a * yin;                           // value entering chunk j + 1
        const unsigned char *st = smem + (size_t)j * C::STAGE;
        const int64_t lo_full = hi - (int64_t)(j + 2) * CH + 1;
        const int64_t q0 = lo_full + i0;
        if (q0 >= 0) {
            float rj[ITEMS], vj[ITEMS + 1];
            unsigned dj;
            gae_rows_from_smem<ITEMS>(reinterpret_cast<const float *>(st + C::REW), reinterpret_cast<const float *>(st + C::VAL),
                                      st + C::DONE, i0, rj, vj, dj);
            float y = excl[1 + j].b + excl[1 + j].a * yin;
            float av[ITEMS], rt[ITEMS];
#pragma unroll
            for (int k = ITEMS - 1; k >= 0; --k) {
                const bool dk = dj & (1u << k);
                const float b = dk ? rj[k] - vj[k] : rj[k] + el.g * vj[k + 1] - vj[k];
                y = b + (dk ? 0.0f : el.gl) * y;
                av[k] = y;
                rt[k] = y + vj[k];
            }
#pragma unroll
            for (int q = 0; q < ITEMS / 4; ++q) {
                *reinterpret_cast<float4 *>(el.adv + q0 + 4 * q) = make_float4(av[4 * q], av[4 * q + 1], av[4 * q + 2], av[4 * q + 3]);
                *reinterpret_cast<float4 *>(el.ret + q0 + 4 * q) = make_float4(rt[4 * q], rt[4 * q + 1], rt[4 * q + 2], rt[4 * q + 3]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
__global__ void init_kernel(float *r, float *v, uint8_t *d, int64_t n, float pdone) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i <= n; i += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long x = (unsigned long long)i * 0x9E3779B97F4A7C15ull + 0x1234567ull;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32; x *= 0x94D049BB133111EBull; x ^= x >> 29;
        const float u0 = (float)(x & 0xFFFFFF) / 16777216.0f, u1 = (float)((x >> 24) & 0xFFFFFF) / 16777216.0f;
        const float u2 = (float)((x >> 48) & 0xFFFF) / 65536.0f;
        v[i] = 2.0f * u1 - 1.0f;
        if (i < n) { r[i] = 2.0f * u0 - 1.0f; d[i] = u2 < pdone ? 1 : 0; }
    }
}

static float *d_rew, *d_val, *d_adv, *d_ret;
static uint8_t *d_done;
static void *d_ws;
static int64_t N;
static std::vector<float> h_adv_ref, h_ret_ref;
static int n_sm;

template <typename F>
static void run_variant(const char *name, int64_t chunk, F launch) {
    const int64_t nb = (N + chunk - 1) / chunk;
    const size_t used = (size_t)nb * sizeof(Tile) + 64;
    CK(cudaMemset(d_adv, 0xFF, N * 4)); CK(cudaMemset(d_ret, 0xFF, N * 4));
    CK(cudaMemsetAsync(d_ws, 0, used, 0));
    launch(nb, (Tile *)((char *)d_ws + 64), (unsigned int *)d_ws);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-28s FAILED: %s\n", name, cudaGetErrorString(e)); exit(1); }
    std::vector<float> ha(N), hr(N);
    CK(cudaMemcpy(ha.data(), d_adv, N * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hr.data(), d_ret, N * 4, cudaMemcpyDeviceToHost));
    double ea = 0, er = 0;
    for (int64_t i = 0; i < N; ++i) {
        const double da = fabs((double)ha[i] - h_adv_ref[i]), dr = fabs((double)hr[i] - h_ret_ref[i]);
        if (!(da <= ea)) ea = da;      // NaN-propagating max
        if (!(dr <= er)) er = dr;
    }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    std::vector<float> ts;
    for (int rep = 0; rep < 7; ++rep) {
        CK(cudaEventRecord(e0));
        for (int i = 0; i < 8; ++i) {
            CK(cudaMemsetAsync(d_ws, 0, used, 0));
            launch(nb, (Tile *)((char *)d_ws + 64), (unsigned int *)d_ws);
        }
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        ts.push_back(ms / 8);
    }
    std::sort(ts.begin() + 1, ts.end());
    const float med = ts[1 + (ts.size() - 1) / 2], best = ts[1];
    printf("%-28s median %.4f ms (%.0f GB/s)  best %.4f ms (%.0f GB/s)  max|adv err| %.3g  max|ret err| %.3g\n", name, med,
           17.0 * N / med / 1e6, best, 17.0 * N / best / 1e6, ea, er);
    fflush(stdout);
}

template <int NT, int ITEMS, int MINB>
static void run_reg(const char *name) {
    Args a{d_rew, d_val, d_done, d_adv, d_ret, 0.99f, (float)(0.99 * 0.95)};
    run_variant(name, NT * ITEMS, [&](int64_t nb, Tile *tiles, unsigned int *) {
        gae_reg_kernel<NT, ITEMS, MINB><<<(unsigned)nb, NT>>>(a, N, tiles);
    });
}
template <int NT, int ITEMS, int STAGES>
static void run_tma(const char *name, int ctas_per_sm_cap = 0) {
    using C = TmaCfg<NT, ITEMS, STAGES>;
    Args a{d_rew, d_val, d_done, d_adv, d_ret, 0.99f, (float)(0.99 * 0.95)};
    auto k = gae_tma_kernel<NT, ITEMS, STAGES>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NT, C::SMEM));
    if (ctas_per_sm_cap && occ > ctas_per_sm_cap) occ = ctas_per_sm_cap;
    char nm[128];
    snprintf(nm, sizeof nm, "%s [%d CTAs/SM, %d KB]", name, occ, C::SMEM / 1024);
    run_variant(nm, C::CH, [&](int64_t nb, Tile *tiles, unsigned int *ticket) {
        int64_t grid = (int64_t)n_sm * occ;
        if (grid > nb) grid = nb;
        k<<<(unsigned)grid, NT, C::SMEM>>>(a, N, nb, tiles, ticket);
    });
}

template <int NT, int ITEMS, int MINB, int NB>
static void run_hyb(const char *name) {
    using C = HybCfg<NT, ITEMS, NB>;
    Args a{d_rew, d_val, d_done, d_adv, d_ret, 0.99f, (float)(0.99 * 0.95)};
    auto k = gae_hyb_kernel<NT, ITEMS, MINB, NB>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, NT, C::SMEM));
    char nm[128];
    snprintf(nm, sizeof nm, "%s [%d CTAs/SM, %d rows/SM]", name, occ, occ * C::ROWS);
    run_variant(nm, C::ROWS, [&](int64_t nb, Tile *tiles, unsigned int *) {
        k<<<(unsigned)nb, NT, C::SMEM>>>(a, N, tiles);
    });
}

int main(int argc, char **argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 26;
    const int64_t extra = argc > 2 ? atoll(argv[2]) : 0;          // e.g. 16 * k: a ragged last chunk
    N = ((int64_t)1 << lg) + extra;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    n_sm = prop.multiProcessorCount;
    printf("%s, %d SMs, n = %lld\n", prop.name, n_sm, (long long)N);
    CK(cudaMalloc(&d_rew, N * 4)); CK(cudaMalloc(&d_val, (N + 1) * 4)); CK(cudaMalloc(&d_done, N));
    CK(cudaMalloc(&d_adv, N * 4)); CK(cudaMalloc(&d_ret, N * 4)); CK(cudaMalloc(&d_ws, (N / 512 + 2) * 32 + 64));
    init_kernel<<<n_sm * 8, 256>>>(d_rew, d_val, d_done, N, 0.002f);
    CK(cudaDeviceSynchronize());
    {
        std::vector<float> r(N), v(N + 1); std::vector<uint8_t> d(N);
        CK(cudaMemcpy(r.data(), d_rew, N * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(v.data(), d_val, (N + 1) * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(d.data(), d_done, N, cudaMemcpyDeviceToHost));
        h_adv_ref.resize(N); h_ret_ref.resize(N);
        const double g = (double)0.99f, gl = (double)(float)(0.99 * 0.95);
        double y = 0;
        for (int64_t i = N - 1; i >= 0; --i) {
            const double b = d[i] ? (double)r[i] - v[i] : (double)r[i] + g * v[i + 1] - v[i];
            y = b + (d[i] ? 0.0 : gl) * y;
            h_adv_ref[i] = (float)y; h_ret_ref[i] = (float)(y + v[i]);
        }
    }
    run_reg<256, 8, 8>("reg<256,8,8> (library)");
    run_hyb<256, 4, 6, 4>("hyb<256,4,6,4>");
    run_hyb<256, 4, 8, 3>("hyb<256,4,8,3>");
    run_hyb<256, 4, 5, 5>("hyb<256,4,5,5>");
    run_hyb<256, 4, 4, 6>("hyb<256,4,4,6>");
    run_hyb<256, 4, 4, 7>("hyb<256,4,4,7>");
    run_hyb<256, 4, 3, 8>("hyb<256,4,3,8>");
    run_hyb<512, 4, 3, 4>("hyb<512,4,3,4>");
    run_hyb<512, 4, 4, 3>("hyb<512,4,4,3>");
    run_hyb<512, 4, 2, 6>("hyb<512,4,2,6>");
    run_hyb<128, 4, 12, 5>("hyb<128,4,12,5>");
    run_hyb<128, 4, 10, 6>("hyb<128,4,10,6>");
    run_hyb<128, 4, 8, 8>("hyb<128,4,8,8>");
    run_hyb<1024, 4, 2, 3>("hyb<1024,4,2,3>");
    run_hyb<1024, 4, 1, 6>("hyb<1024,4,1,6>");
    return 0;
}
