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
// Each CTA scans a chunk of 2048 elements (8 per thread, warp-shuffle scan);
// chunk aggregates are scanned by one CTA; a third pass applies the carries.
// Rollouts of <= 2048 steps need a single launch.  HBM-bound: 17 B / transition
// algorithmic; the second read of the inputs is served by the 126 MB L2.
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

extern "C" int64_t mq_scan_ws_bytes(int64_t n) {
    const int64_t nb = mq_ceil_div(n < 1 ? 1 : n, SCAN_CHUNK);
    return nb * 3 * (int64_t)sizeof(double) + 64;
}

template <typename T, typename E>
static int scan_run(E el, int64_t n, void *ws, int64_t ws_bytes, void *stream, const char *what) {
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

extern "C" int mq_gae(const float *rew, const float *value, const uint8_t *done, int64_t n,
                      double gam, double lam, float *adv, float *ret, void *ws, int64_t ws_bytes,
                      void *stream) {
    MQ_REQUIRE(n >= 0, MQ_ERR_SHAPE, "mq_gae: negative length");
    MQ_REQUIRE(n == 0 || (rew && value && done && adv && ret), MQ_ERR_INVALID, "mq_gae: null pointer");
    GaeElems el;
    el.rew = rew; el.val = value; el.done = done; el.adv = adv; el.ret = ret;
    el.g = (float)gam;
    el.gl = (float)(gam * lam);           // gae.py:56: python-float product, then fp32
    return scan_run<float, GaeElems>(el, n, ws, ws_bytes, stream, "mq_gae");
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
    return scan_run<double, DiscElems>(el, n, ws, ws_bytes, stream, "mq_discounted");
}
