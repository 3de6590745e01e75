// Segmented reverse scan of a first-order linear recurrence
//     y[i] = b[i] + a[i] * y[i+1],   y[n] = 0
// used for GAE advantages + returns (benchmarks/gae.py:43-64,69) and for
// `discounted` (mannequin/_utils.py:9-20, mannequin/_trajectory.py:50-55).
//
//   GAE:        a[i] = done[i] ? 0 : gam*lam
//               b[i] = r[i] + (done[i] ? 0 : gam*v[i+1]) - v[i]
//   discounted: a[i] = 1 - 1/h,  b[i] = x[i]/h          (fp64, like the reference)
//
// Episode boundaries are carried by a[i] = 0, so segmentation is exact by
// construction.  Affine maps compose associatively:
//     (a1,b1) then (a2,b2)  =  (a1*a2, b2 + a2*b1)
// ONE pass over HBM (17 B / transition: the algorithmic minimum): a CTA takes the chunk of 2048
// elements of its block index, stages it in shared memory with coalesced loads,
// scans it (8 elements per thread, warp-shuffle scan), publishes the chunk's map, obtains the value
// entering the chunk by looking back over the preceding chunks (decoupled look-back; a chunk that
// contains an episode end has a = 0 and ends the walk at once), publishes the value leaving it and
// writes its outputs with coalesced stores.  Rollouts of <= 2048 steps are a single CTA.
// (The three-launch chunk / carry / apply kernels below remain for that case and for tools/emu.)
#include "mq_common.cuh"

#define SCAN_ITEMS 8
#define SCAN_CHUNK (MQ_NT * SCAN_ITEMS)

template <typename T> struct Lin { T a, b; };

template <typename T>
__device__ __forceinline__ Lin<T> lin_then(const Lin<T> &first, const Lin<T> &second) {
    Lin<T> r;
    r.a = first.a * second.a;
    r.b = second.b + second.a * first.b;
    return r;
}

template <typename T>
__device__ __forceinline__ Lin<T> lin_shfl_up(const Lin<T> &v, int d) {
    Lin<T> r;
    r.a = __shfl_up_sync(0xffffffffu, v.a, d);
    r.b = __shfl_up_sync(0xffffffffu, v.b, d);
    return r;
}

// Exclusive prefix (maps applied before this thread) and CTA total.
template <typename T>
__device__ __forceinline__ void lin_block_scan(Lin<T> local, Lin<T> &excl, Lin<T> &total,
                                               Lin<T> *warp_tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Lin<T> incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Lin<T> p = lin_shfl_up(incl, o);
        if (lane >= o) incl = lin_then(p, incl);
    }
    __syncthreads();
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    Lin<T> wp; wp.a = (T)1; wp.b = (T)0;
    Lin<T> all = wp;
#pragma unroll
    for (int w = 0; w < MQ_NT / 32; ++w) {
        if (w == warp) wp = all;
        all = lin_then(all, warp_tot[w]);
    }
    Lin<T> prev = lin_shfl_up(incl, 1);
    excl = (lane == 0) ? wp : lin_then(wp, prev);
    total = all;
}

struct GaeElems {
    const float *rew, *val;
    const uint8_t *done;
    float *adv, *ret;
    float g, gl;
    __device__ __forceinline__ Lin<float> get(int64_t i) const {
        Lin<float> e;
        const float v0 = val[i];
        if (done[i]) { e.a = 0.0f; e.b = rew[i] - v0; }
        else { e.a = gl; e.b = rew[i] + g * val[i + 1] - v0; }
        return e;
    }
    __device__ __forceinline__ void put(int64_t i, float y) const {
        adv[i] = y;
        ret[i] = y + val[i];
    }
};

struct DiscElems {
    const double *in;
    double *out;
    double keep, step;
    __device__ __forceinline__ Lin<double> get(int64_t i) const {
        Lin<double> e;
        e.a = keep;
        e.b = step * in[i];
        return e;
    }
    __device__ __forceinline__ void put(int64_t i, double y) const { out[i] = y; }
};

// PHASE 0: aggregates only; PHASE 1: final values using carry[blockIdx] (or 0).
template <typename T, typename E, int PHASE>
__global__ void __launch_bounds__(MQ_NT)
scan_chunk_kernel(E el, int64_t n, Lin<T> *__restrict__ agg, const T *__restrict__ carry) {
    __shared__ Lin<T> warp_tot[MQ_NT / 32];
    // reversed position j = n-1-i runs forward through the recurrence
    const int64_t j0 = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * SCAN_ITEMS;
    Lin<T> item[SCAN_ITEMS];
    Lin<T> local; local.a = (T)1; local.b = (T)0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int64_t j = j0 + k;
        if (j < n) item[k] = el.get(n - 1 - j);
        else { item[k].a = (T)1; item[k].b = (T)0; }
        local = lin_then(local, item[k]);
    }
    Lin<T> excl, total;
    lin_block_scan(local, excl, total, warp_tot);
    if (PHASE == 0) {
        if (threadIdx.x == 0) agg[blockIdx.x] = total;
    } else {
        const T y_in = carry ? carry[blockIdx.x] : (T)0;
        T y = excl.b + excl.a * y_in;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) {
            const int64_t j = j0 + k;
            y = item[k].b + item[k].a * y;
            if (j < n) el.put(n - 1 - j, y);
        }
    }
}

// carry[c] = value of the recurrence entering chunk c (one CTA, any nb).
template <typename T>
__global__ void __launch_bounds__(MQ_NT)
scan_carry_kernel(const Lin<T> *__restrict__ agg, int64_t nb, T *__restrict__ carry) {
    __shared__ Lin<T> warp_tot[MQ_NT / 32];
    const int64_t per = (nb + MQ_NT - 1) / MQ_NT;
    const int64_t c0 = (int64_t)threadIdx.x * per;
    Lin<T> local; local.a = (T)1; local.b = (T)0;
    for (int64_t c = c0; c < c0 + per && c < nb; ++c) local = lin_then(local, agg[c]);
    Lin<T> excl, total;
    lin_block_scan(local, excl, total, warp_tot);
    T y = excl.b;                                     // entering value is 0
    for (int64_t c = c0; c < c0 + per && c < nb; ++c) {
        carry[c] = y;
        y = agg[c].b + agg[c].a * y;
    }
}

// ---- single-pass variant -----------------------------------------------------------------------------
#ifndef MQ_EMU
#define SCAN_PAD(i) ((i) + ((i) >> 5))             // conflict-free 8-consecutive-per-thread access

struct GaeStage {
    const float *rew, *val;
    const uint8_t *done;
    float *adv, *ret;
    float g, gl;
    static constexpr int kSmemBytes = (SCAN_PAD(SCAN_CHUNK) + 1) * 4 * 2 + SCAN_CHUNK + 64;
    // memory rows [lo, lo+cnt) -> shared memory; val also at lo+cnt (bootstrap / next state)
    int vec;                     // all five arrays are 16-byte aligned: 128-bit loads / stores on the aligned middle
    __device__ __forceinline__ void stage(unsigned char *sm, int64_t lo, int cnt) const {
        float *sr = reinterpret_cast<float *>(sm), *sv = sr + SCAN_PAD(SCAN_CHUNK) + 1;
        uint8_t *sd = reinterpret_cast<uint8_t *>(sv + SCAN_PAD(SCAN_CHUNK) + 1 + 8);
        if (vec) {
            // rows lo + head, lo + head + 4, ... start 16-byte groups in every array (same misalignment everywhere)
            int head = (int)((4 - (lo & 3)) & 3);
            if (head > cnt) head = cnt;
            const int nq = (cnt - head) >> 2, nqv = (cnt + 1 - head) >> 2;
            const float4 *r4 = reinterpret_cast<const float4 *>(rew + lo + head);
            const float4 *v4 = reinterpret_cast<const float4 *>(val + lo + head);
            const uchar4 *d4 = reinterpret_cast<const uchar4 *>(done + lo + head);
            for (int q = threadIdx.x; q < nq; q += MQ_NT) {
                const float4 a = r4[q];
                const uchar4 dd = d4[q];
                const int i = head + 4 * q;
                sr[SCAN_PAD(i)] = a.x; sr[SCAN_PAD(i + 1)] = a.y; sr[SCAN_PAD(i + 2)] = a.z; sr[SCAN_PAD(i + 3)] = a.w;
                sd[i] = dd.x; sd[i + 1] = dd.y; sd[i + 2] = dd.z; sd[i + 3] = dd.w;
            }
            for (int q = threadIdx.x; q < nqv; q += MQ_NT) {
                const float4 a = v4[q];
                const int i = head + 4 * q;
                sv[SCAN_PAD(i)] = a.x; sv[SCAN_PAD(i + 1)] = a.y; sv[SCAN_PAD(i + 2)] = a.z; sv[SCAN_PAD(i + 3)] = a.w;
            }
            // the unaligned ends (at most 3 + 3 rows; 4 + 4 for val)
            for (int i = threadIdx.x; i < cnt; i += MQ_NT)
                if (i < head || i >= head + 4 * nq) { sr[SCAN_PAD(i)] = rew[lo + i]; sd[i] = done[lo + i]; }
            for (int i = threadIdx.x; i <= cnt; i += MQ_NT)
                if (i < head || i >= head + 4 * nqv) sv[SCAN_PAD(i)] = val[lo + i];
            return;
        }
        for (int i = threadIdx.x; i < cnt; i += MQ_NT) { sr[SCAN_PAD(i)] = rew[lo + i]; sd[i] = done[lo + i]; }
        for (int i = threadIdx.x; i <= cnt; i += MQ_NT) sv[SCAN_PAD(i)] = val[lo + i];
    }
    __device__ __forceinline__ Lin<float> get(const unsigned char *sm, int li) const {
        const float *sr = reinterpret_cast<const float *>(sm), *sv = sr + SCAN_PAD(SCAN_CHUNK) + 1;
        const uint8_t *sd = reinterpret_cast<const uint8_t *>(sv + SCAN_PAD(SCAN_CHUNK) + 1 + 8);
        Lin<float> e;
        const float v0 = sv[SCAN_PAD(li)];
        if (sd[li]) { e.a = 0.0f; e.b = sr[SCAN_PAD(li)] - v0; }
        else { e.a = gl; e.b = sr[SCAN_PAD(li)] + g * sv[SCAN_PAD(li + 1)] - v0; }
        return e;
    }
    __device__ __forceinline__ void put(unsigned char *sm, int li, float y) const {
        reinterpret_cast<float *>(sm)[SCAN_PAD(li)] = y;          // over the reward: already consumed
    }
    __device__ __forceinline__ void flush(const unsigned char *sm, int64_t lo, int cnt) const {
        const float *sr = reinterpret_cast<const float *>(sm), *sv = sr + SCAN_PAD(SCAN_CHUNK) + 1;
        if (vec) {
            int head = (int)((4 - (lo & 3)) & 3);
            if (head > cnt) head = cnt;
            const int nq = (cnt - head) >> 2;
            float4 *a4 = reinterpret_cast<float4 *>(adv + lo + head), *t4 = reinterpret_cast<float4 *>(ret + lo + head);
            for (int q = threadIdx.x; q < nq; q += MQ_NT) {
                const int i = head + 4 * q;
                const float y0 = sr[SCAN_PAD(i)], y1 = sr[SCAN_PAD(i + 1)], y2 = sr[SCAN_PAD(i + 2)], y3 = sr[SCAN_PAD(i + 3)];
                a4[q] = make_float4(y0, y1, y2, y3);
                t4[q] = make_float4(y0 + sv[SCAN_PAD(i)], y1 + sv[SCAN_PAD(i + 1)], y2 + sv[SCAN_PAD(i + 2)],
                                    y3 + sv[SCAN_PAD(i + 3)]);
            }
            for (int i = threadIdx.x; i < cnt; i += MQ_NT)
                if (i < head || i >= head + 4 * nq) {
                    const float y = sr[SCAN_PAD(i)];
                    adv[lo + i] = y;
                    ret[lo + i] = y + sv[SCAN_PAD(i)];
                }
            return;
        }
        for (int i = threadIdx.x; i < cnt; i += MQ_NT) {
            const float y = sr[SCAN_PAD(i)];
            adv[lo + i] = y;
            ret[lo + i] = y + sv[SCAN_PAD(i)];
        }
    }
};

struct DiscStage {
    const double *in;
    double *out;
    double keep, step;
    static constexpr int kSmemBytes = (SCAN_PAD(SCAN_CHUNK) + 1) * 8;
    __device__ __forceinline__ void stage(unsigned char *sm, int64_t lo, int cnt) const {
        double *sx = reinterpret_cast<double *>(sm);
        for (int i = threadIdx.x; i < cnt; i += MQ_NT) sx[SCAN_PAD(i)] = in[lo + i];
    }
    __device__ __forceinline__ Lin<double> get(const unsigned char *sm, int li) const {
        Lin<double> e;
        e.a = keep;
        e.b = step * reinterpret_cast<const double *>(sm)[SCAN_PAD(li)];
        return e;
    }
    __device__ __forceinline__ void put(unsigned char *sm, int li, double y) const {
        reinterpret_cast<double *>(sm)[SCAN_PAD(li)] = y;
    }
    __device__ __forceinline__ void flush(const unsigned char *sm, int64_t lo, int cnt) const {
        const double *sx = reinterpret_cast<const double *>(sm);
        for (int i = threadIdx.x; i < cnt; i += MQ_NT) out[lo + i] = sx[SCAN_PAD(i)];
    }
};

// per-chunk descriptor: status 0 = empty, 1 = map of the chunk alone, 2 = value leaving the chunk
template <typename T> struct ScanTile { T a, b, y; int status; int pad; };
template <> struct __align__(16) ScanTile<float> { float a; int tag_a; float y; int tag_y; };   // {a, tag}, {b or y, tag}

__device__ __forceinline__ int scan_ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void scan_st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// A float descriptor is TWO aligned 8-byte words, each carrying its own copy of the status: {a, status} and {b or y, status}.
// An aligned 8-byte scalar access is single-copy atomic in the PTX memory model (a 16-byte .v4 access is four scalar
// accesses and may tear), so a word and its tag always travel together and no fence / release-acquire pair is needed:
// status 2 only needs the second word (the value leaving the chunk); status 1 is accepted when BOTH words carry tag 1 --
// the map (a, b) is published exactly once per chunk and launch, so two tag-1 words belong to the same publication.
// (The double descriptor of `discounted` is 32 bytes and keeps the release / acquire protocol.)
__device__ __forceinline__ void scan_st64(void *p, float v, int tag) {
    const unsigned long long w = ((unsigned long long)(unsigned)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long scan_ld64(const void *p) {
    unsigned long long w;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
__device__ __forceinline__ void scan_tile_store(ScanTile<float> *t, float a, float b, float y, int status) {
    if (status == 1) scan_st64(&t->a, a, 1);
    scan_st64(&t->y, status == 1 ? b : y, status);
}
__device__ __forceinline__ void scan_tile_load(const ScanTile<float> *t, float &a, float &b, float &y, int &status) {
    const unsigned long long w1 = scan_ld64(&t->a), w2 = scan_ld64(&t->y);
    const int s1 = (int)(w1 >> 32), s2 = (int)(w2 >> 32);
    a = __uint_as_float((unsigned)w1);
    b = y = __uint_as_float((unsigned)w2);
    status = (s2 == 2) ? 2 : ((s2 == 1 && s1 == 1) ? 1 : 0);
}
__device__ __forceinline__ void scan_tile_store(ScanTile<double> *t, double a, double b, double y, int status) {
    t->a = a; t->b = b; t->y = y;
    __threadfence();
    scan_st_release(&t->status, status);
}
__device__ __forceinline__ void scan_tile_load(const ScanTile<double> *t, double &a, double &b, double &y, int &status) {
    status = scan_ld_acquire(&t->status);
    a = t->a; b = t->b; y = t->y;
}

// Warp 0 of the block that owns chunk c: publish the chunk's map (its leaving value at once if the chunk contains an
// episode end), look back over the predecessors and return the value ENTERING the chunk (every lane returns it).
template <typename T>
__device__ __forceinline__ T scan_publish_and_lookback(ScanTile<T> *__restrict__ tiles, int64_t c, const Lin<T> &total) {

    const int lane = threadIdx.x;
    // a chunk that contains an episode end (a = 0) leaves the same value whatever enters it: publish that value
    // at once so that later chunks never wait for this chunk's own look-back
    const bool closed = total.a == (T)0;
    if (lane == 0) scan_tile_store(&tiles[c], total.a, total.b, total.b, closed ? 2 : 1);
    // acc = composition of the chunks already folded in (those nearest to c); identity at first
    Lin<T> acc; acc.a = (T)1; acc.b = (T)0;
    bool done = false;
    int64_t base = c - 1;                          // nearest predecessor not yet folded in
    while (!done) {
        const int64_t q = base - lane;             // lane 0 = nearest
        T qa = (T)0, qb = (T)0;                    // before chunk 0 the recurrence is 0: a terminator
        if (q >= 0) {
            int st;
            T ta, tb, ty;
            do { scan_tile_load(&tiles[q], ta, tb, ty, st); } while (st == 0);
            if (st == 2) qb = ty;                  // value LEAVING chunk q: ignores what entered it
            else { qa = ta; qb = tb; }
        }
        const unsigned term = __ballot_sync(0xffffffffu, qa == (T)0);
        const int stop = term ? __ffs(term) - 1 : 31;
        Lin<T> win; win.a = (T)1; win.b = (T)0;    // chunks base-stop .. base, oldest applied first
        for (int l = stop; l >= 0; --l) {
            Lin<T> e;
            e.a = __shfl_sync(0xffffffffu, qa, l);
            e.b = __shfl_sync(0xffffffffu, qb, l);
            win = lin_then(win, e);
        }
        acc = lin_then(win, acc);
        done = term != 0u;
        base -= 32;
    }
    if (lane == 0 && !closed) scan_tile_store(&tiles[c], total.a, total.b, total.b + total.a * acc.b, 2);
    return acc.b;
}

template <typename T, typename E>
__global__ void __launch_bounds__(MQ_NT, 8)
scan_lookback_kernel(E el, int64_t n, ScanTile<T> *__restrict__ tiles, unsigned int *__restrict__ ticket) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ Lin<T> warp_tot[MQ_NT / 32];
    __shared__ T s_yin;
    // chunk = block index: like cub::DeviceScan this relies on the hardware dispatching blocks in index order, so
    // that every chunk a block waits for belongs to a block that is already resident (a ticket counter costs one
    // same-address atomic per chunk, which at 32,768 chunks was 12 % of the kernel)
    (void)ticket;
    const int64_t c = blockIdx.x;
    // reversed positions [c*CH, (c+1)*CH) = memory rows [lo, hi]
    const int64_t hi = n - 1 - c * SCAN_CHUNK;
    const int64_t lo = hi - (SCAN_CHUNK - 1) > 0 ? hi - (SCAN_CHUNK - 1) : 0;
    const int cnt = (int)(hi - lo + 1);
    el.stage(sm, lo, cnt);
    __syncthreads();
    // (the elements are re-read from shared memory in the output pass instead of being kept in registers:
    //  8 CTAs per SM keep enough bytes in flight to cover the look-back and the phase changes)
    Lin<T> local; local.a = (T)1; local.b = (T)0;
    const int jl0 = threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int li = cnt - 1 - (jl0 + k);          // memory-local index of reversed position jl0 + k
        if (li >= 0) local = lin_then(local, el.get(sm, li));
    }
    Lin<T> excl, total;
    lin_block_scan(local, excl, total, warp_tot);
    // ---- publish this chunk's map, look back for the value entering it (warp 0) --------------------
    if (threadIdx.x < 32) {
        const T yin = scan_publish_and_lookback<T>(tiles, c, total);
        if (threadIdx.x == 0) s_yin = yin;
    }
    __syncthreads();
    T y = excl.b + excl.a * s_yin;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        const int li = cnt - 1 - (jl0 + k);
        if (li >= 0) {
            const Lin<T> e = el.get(sm, li);
            y = e.b + e.a * y;
            el.put(sm, li, y);
        }
    }
    __syncthreads();
    el.flush(sm, lo, cnt);
}

// GAE with the chunk in REGISTERS (no shared-memory staging): thread t owns the 8 consecutive rows hi - 8t - 7 .. hi - 8t.
// ncu on the staged kernel above showed it bound by instruction issue (76 % issue-active, 133 thread instructions per
// transition: padded shared-memory indices, four LDS per element and every element evaluated twice), not by HBM; here an
// element costs ~20.  Requires 16-byte aligned arrays and n % 8 == 0 (rows of a thread start 128-bit groups); any other
// call takes the staged kernel.
__global__ void __launch_bounds__(MQ_NT, 8)
gae_lookback_kernel(GaeStage el, int64_t n, ScanTile<float> *__restrict__ tiles) {
    __shared__ Lin<float> warp_tot[MQ_NT / 32];
    __shared__ float s_yin;
    const int64_t c = blockIdx.x;
    const int64_t hi = n - 1 - c * SCAN_CHUNK;
    const int64_t r1 = hi - (int64_t)threadIdx.x * SCAN_ITEMS, r0 = r1 - (SCAN_ITEMS - 1);   // this thread's rows
    float rw[SCAN_ITEMS], vl[SCAN_ITEMS + 1];
    unsigned dn = 0;                                   // bit k: episode ends at row r0 + k
    if (r0 >= 0) {
        const float4 a0 = *reinterpret_cast<const float4 *>(el.rew + r0), a1 = *reinterpret_cast<const float4 *>(el.rew + r0 + 4);
        const float4 b0 = *reinterpret_cast<const float4 *>(el.val + r0), b1 = *reinterpret_cast<const float4 *>(el.val + r0 + 4);
        const uint2 d8 = *reinterpret_cast<const uint2 *>(el.done + r0);
        vl[8] = el.val[r1 + 1];
        rw[0] = a0.x; rw[1] = a0.y; rw[2] = a0.z; rw[3] = a0.w; rw[4] = a1.x; rw[5] = a1.y; rw[6] = a1.z; rw[7] = a1.w;
        vl[0] = b0.x; vl[1] = b0.y; vl[2] = b0.z; vl[3] = b0.w; vl[4] = b1.x; vl[5] = b1.y; vl[6] = b1.z; vl[7] = b1.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            dn |= ((d8.x >> (8 * k)) & 0xFFu) ? (1u << k) : 0u;
            dn |= ((d8.y >> (8 * k)) & 0xFFu) ? (1u << (4 + k)) : 0u;
        }
    } else {                                           // rows before the start of the data: identity maps
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; ++k) { rw[k] = 0.0f; vl[k] = 0.0f; }
        vl[8] = 0.0f;
    }
    const bool live = r0 >= 0;
    // maps of the 8 rows, newest row first (reversed positions ascend as rows descend); b_k kept for the output pass
    float eb[SCAN_ITEMS];
    Lin<float> local; local.a = 1.0f; local.b = 0.0f;
#pragma unroll
    for (int k = SCAN_ITEMS - 1; k >= 0; --k) {
        Lin<float> e;
        if (dn & (1u << k)) { e.a = 0.0f; e.b = rw[k] - vl[k]; }
        else { e.a = el.gl; e.b = rw[k] + el.g * vl[k + 1] - vl[k]; }
        eb[k] = e.b;
        if (live) local = lin_then(local, e);
    }
    Lin<float> excl, total;
    lin_block_scan(local, excl, total, warp_tot);
    if (threadIdx.x < 32) {
        const float yin = scan_publish_and_lookback<float>(tiles, c, total);
        if (threadIdx.x == 0) s_yin = yin;
    }
    __syncthreads();
    if (!live) return;
    float y = excl.b + excl.a * s_yin;
    float av[SCAN_ITEMS], rt[SCAN_ITEMS];
#pragma unroll
    for (int k = SCAN_ITEMS - 1; k >= 0; --k) {
        y = eb[k] + ((dn & (1u << k)) ? 0.0f : el.gl) * y;
        av[k] = y;
        rt[k] = y + vl[k];
    }
    *reinterpret_cast<float4 *>(el.adv + r0) = make_float4(av[0], av[1], av[2], av[3]);
    *reinterpret_cast<float4 *>(el.adv + r0 + 4) = make_float4(av[4], av[5], av[6], av[7]);
    *reinterpret_cast<float4 *>(el.ret + r0) = make_float4(rt[0], rt[1], rt[2], rt[3]);
    *reinterpret_cast<float4 *>(el.ret + r0 + 4) = make_float4(rt[4], rt[5], rt[6], rt[7]);
}


// GAE, hybrid register / shared-memory form (the default for 16-byte aligned arrays with n % 16 == 0).
// The register kernel above is bound by how many rows an SM keeps in flight: every CTA exposes the latency of its loads
// and of one look-back round trip through L2, and the register file holds only ~16 k rows per SM (ncu: barrier stalls
// dominate, 64 % of the HBM peak).  Here a CTA owns GH_SUB consecutive sub-chunks of 1024 rows (4 rows = one 128-bit load
// per thread, fully coalesced): sub-chunk 0 is loaded into registers, sub-chunks 1.. are fetched by bulk copies (TMA engine,
// one issuing thread, completion counted on an mbarrier) into shared memory at kernel start, so an SM has 32 k rows in
// flight and there is ONE global look-back per 4096 rows.  The CTA first reduces every sub-chunk to its map (the staged
// ones are read twice from shared memory), publishes the composed map, looks back, then writes the outputs of the
// register sub-chunk and walks the staged ones with the value carried on chip.  (Persistent ring-buffered variants with
// bulk stores were measured 1.4-2x SLOWER: one chain of look-back + ticket latencies per CTA cannot be hidden by 2-4
// resident CTAs; tools/gae_variants.cu keeps the variants and their timings are in profiles/r2_gae_variants.txt.)
#define GH_ITEMS 4
#define GH_SUB 4                                   // sub-chunks per CTA: 1 in registers + 3 staged
#define GH_CH (MQ_NT * GH_ITEMS)                   // rows per sub-chunk
#define GH_ROWS (GH_CH * GH_SUB)                   // rows per CTA
#define GH_STAGE (GH_CH * 4 + (GH_CH + 4) * 4 + GH_CH)   // rewards | values (+ the value after the newest row) | done flags
#define GH_SMEM (GH_STAGE * (GH_SUB - 1))

__device__ __forceinline__ unsigned gh_s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gh_bulk_load(void *dst, const void *src, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(gh_s32(dst)), "l"(src), "r"(bytes), "r"(gh_s32(bar)) : "memory");
}
__device__ __forceinline__ void gh_wait(uint64_t *bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(gh_s32(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// the 4 rows of this thread in a staged sub-chunk: maps newest row first, like the register path
struct GhRows { float rw[GH_ITEMS], vl[GH_ITEMS + 1]; unsigned dn; };
__device__ __forceinline__ void gh_rows(const unsigned char *st, int i0, GhRows &r) {
    const float4 a = *reinterpret_cast<const float4 *>(st + (size_t)i0 * 4);
    const float4 b = *reinterpret_cast<const float4 *>(st + GH_CH * 4 + (size_t)i0 * 4);
    const unsigned d = *reinterpret_cast<const unsigned *>(st + GH_CH * 4 + (GH_CH + 4) * 4 + i0);
    r.rw[0] = a.x; r.rw[1] = a.y; r.rw[2] = a.z; r.rw[3] = a.w;
    r.vl[0] = b.x; r.vl[1] = b.y; r.vl[2] = b.z; r.vl[3] = b.w;
    r.vl[4] = *reinterpret_cast<const float *>(st + GH_CH * 4 + (size_t)(i0 + 4) * 4);
    r.dn = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.dn |= ((d >> (8 * k)) & 0xFFu) ? (1u << k) : 0u;
}
__device__ __forceinline__ Lin<float> gh_warp_scan(Lin<float> v) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Lin<float> p = lin_shfl_up(v, o);
        if (lane >= o) v = lin_then(p, v);
    }
    return v;
}

__global__ void __launch_bounds__(MQ_NT, 8)
gae_hybrid_kernel(GaeStage el, int64_t n, ScanTile<float> *__restrict__ tiles) {
    extern __shared__ __align__(128) unsigned char gh_smem[];
    __shared__ uint64_t full[GH_SUB - 1];
    __shared__ Lin<float> warp_tot[GH_SUB][MQ_NT / 32];
    __shared__ float s_yin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t c = blockIdx.x;                                  // (block index = chunk: see scan_lookback_kernel)
    const int64_t hi = n - 1 - c * GH_ROWS;                        // newest row of the CTA
    if (threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < GH_SUB - 1; ++j)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gh_s32(&full[j])), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;");
#pragma unroll
        for (int j = 0; j < GH_SUB - 1; ++j) {
            unsigned char *st = gh_smem + (size_t)j * GH_STAGE;
            // stage slot i <-> row lo_full + i; rows before the start of the data are left out (their threads stay idle)
            const int64_t hj = hi - (int64_t)(j + 1) * GH_CH, lo_full = hj - (GH_CH - 1);
            const int64_t lo = lo_full > 0 ? lo_full : 0;
            const int cnt = (int)(hj - lo + 1), off = (int)(lo - lo_full);
            if (cnt > 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                             ::"r"(gh_s32(&full[j])), "r"((unsigned)(cnt * 9 + 16)) : "memory");
                gh_bulk_load(st + off * 4, el.rew + lo, (unsigned)(cnt * 4), &full[j]);
                gh_bulk_load(st + GH_CH * 4 + off * 4, el.val + lo, (unsigned)(cnt * 4 + 16), &full[j]);   // + val[hj + 1 ..]
                gh_bulk_load(st + GH_CH * 4 + (GH_CH + 4) * 4 + off, el.done + lo, (unsigned)cnt, &full[j]);
            } else {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(gh_s32(&full[j])) : "memory");
            }
        }
    }
    // ---- sub-chunk 0: registers --------------------------------------------------------------------
    const int64_t r0 = hi - (int64_t)threadIdx.x * GH_ITEMS - (GH_ITEMS - 1);
    const bool live = r0 >= 0;
    float vl[GH_ITEMS + 1], eb[GH_ITEMS];
    unsigned dn = 0;
    {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        unsigned d = 0;
        vl[4] = 0.0f;
        if (live) {
            a = *reinterpret_cast<const float4 *>(el.rew + r0);
            b = *reinterpret_cast<const float4 *>(el.val + r0);
            d = *reinterpret_cast<const unsigned *>(el.done + r0);
            vl[4] = el.val[r0 + GH_ITEMS];
        }
        eb[0] = a.x; eb[1] = a.y; eb[2] = a.z; eb[3] = a.w;        // rewards for now
        vl[0] = b.x; vl[1] = b.y; vl[2] = b.z; vl[3] = b.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) dn |= ((d >> (8 * k)) & 0xFFu) ? (1u << k) : 0u;
    }
    __syncthreads();                                               // barriers initialised before anyone waits on them
    Lin<float> incl[GH_SUB];
    {
        Lin<float> local; local.a = 1.0f; local.b = 0.0f;
#pragma unroll
        for (int k = GH_ITEMS - 1; k >= 0; --k) {
            Lin<float> e;
            if (dn & (1u << k)) { e.a = 0.0f; e.b = eb[k] - vl[k]; }
            else { e.a = el.gl; e.b = eb[k] + el.g * vl[k + 1] - vl[k]; }
            eb[k] = e.b;
            if (live) local = lin_then(local, e);
        }
        incl[0] = gh_warp_scan(local);
    }
    // ---- staged sub-chunks: maps only ---------------------------------------------------------------
    const int i0 = GH_CH - ((int)threadIdx.x + 1) * GH_ITEMS;      // this thread's rows in a stage
#pragma unroll
    for (int j = 0; j < GH_SUB - 1; ++j) {
        const int64_t lo_full = hi - (int64_t)(j + 2) * GH_CH + 1;
        Lin<float> lj; lj.a = 1.0f; lj.b = 0.0f;
        gh_wait(&full[j], 0u);
        if (lo_full + i0 >= 0) {
            GhRows r;
            gh_rows(gh_smem + (size_t)j * GH_STAGE, i0, r);
#pragma unroll
            for (int k = GH_ITEMS - 1; k >= 0; --k) {
                Lin<float> e;
                if (r.dn & (1u << k)) { e.a = 0.0f; e.b = r.rw[k] - r.vl[k]; }
                else { e.a = el.gl; e.b = r.rw[k] + el.g * r.vl[k + 1] - r.vl[k]; }
                lj = lin_then(lj, e);
            }
        }
        incl[1 + j] = gh_warp_scan(lj);
    }
    if (lane == 31) {
#pragma unroll
        for (int j = 0; j < GH_SUB; ++j) warp_tot[j][warp] = incl[j];
    }
    __syncthreads();
    Lin<float> excl[GH_SUB], tot[GH_SUB];
    Lin<float> all; all.a = 1.0f; all.b = 0.0f;
#pragma unroll
    for (int j = 0; j < GH_SUB; ++j) {
        Lin<float> wp; wp.a = 1.0f; wp.b = 0.0f;
        Lin<float> t = wp;
#pragma unroll
        for (int w = 0; w < MQ_NT / 32; ++w) {
            if (w == warp) wp = t;
            t = lin_then(t, warp_tot[j][w]);
        }
        const Lin<float> prev = lin_shfl_up(incl[j], 1);
        excl[j] = (lane == 0) ? wp : lin_then(wp, prev);
        tot[j] = t;
        all = lin_then(all, t);
    }
    if (threadIdx.x < 32) {
        const float y0 = scan_publish_and_lookback<float>(tiles, c, all);
        if (threadIdx.x == 0) s_yin = y0;
    }
    __syncthreads();
    float yin = s_yin;
    if (live) {
        float y = excl[0].b + excl[0].a * yin;
        float av[GH_ITEMS], rt[GH_ITEMS];
#pragma unroll
        for (int k = GH_ITEMS - 1; k >= 0; --k) {
            y = eb[k] + ((dn & (1u << k)) ? 0.0f : el.gl) * y;
            av[k] = y;
            rt[k] = y + vl[k];
        }
        *reinterpret_cast<float4 *>(el.adv + r0) = make_float4(av[0], av[1], av[2], av[3]);
        *reinterpret_cast<float4 *>(el.ret + r0) = make_float4(rt[0], rt[1], rt[2], rt[3]);
    }
#pragma unroll
    for (int j = 0; j < GH_SUB - 1; ++j) {
        yin = tot[j].b + tot[j].a * yin;                           // value entering staged sub-chunk j
        const int64_t q0 = hi - (int64_t)(j + 2) * GH_CH + 1 + i0;
        if (q0 >= 0) {
            GhRows r;
            gh_rows(gh_smem + (size_t)j * GH_STAGE, i0, r);
            float y = excl[1 + j].b + excl[1 + j].a * yin;
            float av[GH_ITEMS], rt[GH_ITEMS];
#pragma unroll
            for (int k = GH_ITEMS - 1; k >= 0; --k) {
                const bool dk = (r.dn & (1u << k)) != 0u;
                const float b = dk ? r.rw[k] - r.vl[k] : r.rw[k] + el.g * r.vl[k + 1] - r.vl[k];
                y = b + (dk ? 0.0f : el.gl) * y;
                av[k] = y;
                rt[k] = y + r.vl[k];
            }
            *reinterpret_cast<float4 *>(el.adv + q0) = make_float4(av[0], av[1], av[2], av[3]);
            *reinterpret_cast<float4 *>(el.ret + q0) = make_float4(rt[0], rt[1], rt[2], rt[3]);
        }
    }
}

#else
struct GaeStage { const float *rew, *val; const uint8_t *done; float *adv, *ret; float g, gl; int vec; };
struct DiscStage { const double *in; double *out; double keep, step; };
#endif

extern "C" int64_t mq_scan_ws_bytes(int64_t n) {
    const int64_t nb = mq_ceil_div(n < 1 ? 1 : n, SCAN_CHUNK);
    return nb * 32 + 64;                  // one 32-byte descriptor per chunk (ScanTile<double>)
}

#ifndef MQ_EMU
static bool scan_try_register_kernel(const GaeStage &st, int64_t n, ScanTile<float> *tiles, int64_t nb, void *stream) {
    if (!st.vec || (n & 7) != 0) return false;
    if ((n & 15) == 0) {                                   // bulk copies: 16-byte granules of the done flags as well
        if (cudaFuncSetAttribute(gae_hybrid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GH_SMEM) != cudaSuccess)
            return false;
        MQ_LAUNCH(gae_hybrid_kernel, (unsigned)mq_ceil_div(n, (int64_t)GH_ROWS), MQ_NT, GH_SMEM, stream, st, n, tiles);
        return true;
    }
    MQ_LAUNCH(gae_lookback_kernel, (unsigned)nb, MQ_NT, 0, stream, st, n, tiles);
    return true;
}
static bool scan_try_register_kernel(const DiscStage &, int64_t, ScanTile<double> *, int64_t, void *) { return false; }

template <typename T, typename E, typename S>
static int scan_run(E el, S st, int64_t n, void *ws, int64_t ws_bytes, void *stream, const char *what) {
    if (n == 0) return MQ_OK;
    const int64_t nb = mq_ceil_div(n, SCAN_CHUNK);
    if (nb == 1) {
        auto k1 = scan_chunk_kernel<T, E, 1>;
        MQ_LAUNCH(k1, 1, MQ_NT, 0, stream, el, n, (Lin<T> *)nullptr, (const T *)nullptr);
    } else {
        MQ_REQUIRE(ws && ws_bytes >= mq_scan_ws_bytes(n), MQ_ERR_SHAPE, "%s: workspace too small", what);
        const int64_t used = nb * (int64_t)sizeof(ScanTile<T>) + 64;
        cudaError_t e = cudaMemsetAsync(ws, 0, (size_t)used, (cudaStream_t)stream);
        if (e != cudaSuccess) { mq_set_error("%s: %s", what, cudaGetErrorString(e)); return MQ_ERR_CUDA; }
        unsigned int *ticket = (unsigned int *)ws;
        ScanTile<T> *tiles = (ScanTile<T> *)((char *)ws + 64);
        if (!scan_try_register_kernel(st, n, tiles, nb, stream)) {
            auto k = scan_lookback_kernel<T, S>;
            MQ_LAUNCH(k, (unsigned)nb, MQ_NT, S::kSmemBytes, stream, st, n, tiles, ticket);
        }
    }
    MQ_CHECK_LAUNCH(what);
    return MQ_OK;
}
#else
template <typename T, typename E, typename S>
static int scan_run(E el, S, int64_t n, void *ws, int64_t ws_bytes, void *stream, const char *what) {
    if (n == 0) return MQ_OK;
    const int64_t nb = mq_ceil_div(n, SCAN_CHUNK);
    if (nb == 1) {
        auto k1 = scan_chunk_kernel<T, E, 1>;
        MQ_LAUNCH(k1, 1, MQ_NT, 0, stream, el, n, (Lin<T> *)nullptr, (const T *)nullptr);
    } else {
        MQ_REQUIRE(ws && ws_bytes >= mq_scan_ws_bytes(n), MQ_ERR_SHAPE, "%s: workspace too small", what);
        Lin<T> *agg = (Lin<T> *)ws;
        T *carry = (T *)((char *)ws + nb * 2 * sizeof(double));
        auto k0 = scan_chunk_kernel<T, E, 0>;
        auto k1 = scan_chunk_kernel<T, E, 1>;
        auto kc = scan_carry_kernel<T>;
        MQ_LAUNCH(k0, (unsigned)nb, MQ_NT, 0, stream, el, n, agg, (const T *)nullptr);
        MQ_LAUNCH(kc, 1, MQ_NT, 0, stream, (const Lin<T> *)agg, nb, carry);
        MQ_LAUNCH(k1, (unsigned)nb, MQ_NT, 0, stream, el, n, agg, (const T *)carry);
    }
    MQ_CHECK_LAUNCH(what);
    return MQ_OK;
}
#endif

extern "C" int mq_gae(const float *rew, const float *value, const uint8_t *done, int64_t n,
                      double gam, double lam, float *adv, float *ret, void *ws, int64_t ws_bytes,
                      void *stream) {
    MQ_REQUIRE(n >= 0, MQ_ERR_SHAPE, "mq_gae: negative length");
    MQ_REQUIRE(n == 0 || (rew && value && done && adv && ret), MQ_ERR_INVALID, "mq_gae: null pointer");
    GaeElems el;
    el.rew = rew; el.val = value; el.done = done; el.adv = adv; el.ret = ret;
    el.g = (float)gam;
    el.gl = (float)(gam * lam);           // gae.py:56: python-float product, then fp32
    GaeStage st;
    st.rew = rew; st.val = value; st.done = done; st.adv = adv; st.ret = ret; st.g = el.g; st.gl = el.gl;
    st.vec = (((uintptr_t)rew | (uintptr_t)value | (uintptr_t)done | (uintptr_t)adv | (uintptr_t)ret) & 15) == 0;
    return scan_run<float, GaeElems, GaeStage>(el, st, n, ws, ws_bytes, stream, "mq_gae");
}

extern "C" int mq_discounted(const double *in, int64_t n, double horizon, double *out, void *ws,
                             int64_t ws_bytes, void *stream) {
    MQ_REQUIRE(n >= 0, MQ_ERR_SHAPE, "mq_discounted: negative length");
    const double step = 1.0 / horizon;
    MQ_REQUIRE(step > 1e-6 && step < 1.0, MQ_ERR_SHAPE,
               "mq_discounted: horizon out of range (mannequin/_utils.py:11)");
    MQ_REQUIRE(n == 0 || (in && out), MQ_ERR_INVALID, "mq_discounted: null pointer");
    DiscElems el;
    el.in = in; el.out = out; el.keep = 1.0 - step; el.step = step;
    DiscStage st;
    st.in = in; st.out = out; st.keep = el.keep; st.step = step;
    return scan_run<double, DiscElems, DiscStage>(el, st, n, ws, ws_bytes, stream, "mq_discounted");
}
