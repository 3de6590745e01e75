// Rollout side of the PPO loop for MANY environments stepped in lock step: the step on either
// side of the update path.
//
//   mq_obsnorm_step      NormalizedObservations of mannequin/gym.py:61-84: obs = clip(normalize(obs), -5, 5) where
//                        normalize = RunningNormalize(shape) called on every observation (update, then apply;
//                        mannequin/_normalize.py:34-81).  Here one call sees the E observations of one time step as ONE
//                        batch -- _normalize.py's batch rule (one weight-1 update of the running mean with the batch
//                        mean, d1 * d2 for the variance, the (n-1)/n first-batch case); E = 1 is the reference's
//                        per-step behaviour exactly.  All running statistics are fp64 and live in device memory
//                        (no host round trip per step, so a whole rollout can be captured in a CUDA graph).
//   mq_actor_env_step    one_step of mannequin/gym.py:115-122 for E environments: draw the action of the Gaussian
//                        policy (distrib.py:50-59: mean + exp(logstd) * unit noise), record (obs, act, rew, done),
//                        advance a SYNTHETIC on-device environment and restart finished episodes.  The noise is a
//                        counter-based Philox4x32-10 stream keyed by (seed, step, env): reproducible, no state.
//   mq_randn_philox      the same normal deviates as a plain array (tests, ES noise)
//   mq_bootstrap_segments  the chunk boundary of benchmarks/gae.py:43-64 for E per-environment segments laid end to
//                        end: at the last step of a segment adv = r + gam * v(next state) - v (adv beyond the chunk is 0),
//                        which is the terminal formula with r' = r + gam * v(next): so r is patched and done set.
//
// The environment is NOT the reference's (those are gym's, out of scope and not shippable): a stable linear system
// with quadratic cost, documented in mannequin2_b200/rollout.py.  It exists so that the closed loop can be measured.
#include "mq_common.cuh"

#define RO_MAXD 64
#define RO_NT 1024

// ---- Philox4x32-10 (Salmon et al. 2011), counter (c0..c3), key (k0, k1) ----------------------------------
struct RoPhilox { uint32_t v[4]; };
__host__ __device__ __forceinline__ RoPhilox ro_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    RoPhilox o;
    o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
    return o;
}
// four unit normals from one Philox block (Box-Muller on (0,1] uniforms)
__host__ __device__ __forceinline__ void ro_normal4(const RoPhilox &r, float *z) {
    const float two_pi = 6.283185307179586f;
    for (int h = 0; h < 2; ++h) {
        const float u1 = ((float)(r.v[2 * h] >> 8) + 1.0f) * (1.0f / 16777216.0f);       // (0, 1]
        const float u2 = (float)(r.v[2 * h + 1] >> 8) * (1.0f / 16777216.0f);             // [0, 1)
        const float rad = sqrtf(-2.0f * logf(u1));
        z[2 * h] = rad * cosf(two_pi * u2);
        z[2 * h + 1] = rad * sinf(two_pi * u2);
    }
}
// normal deviate number `j` of stream (seed, a, b): block j / 4, element j % 4
__host__ __device__ __forceinline__ float ro_normal(uint64_t seed, uint32_t a, uint32_t b, uint32_t j) {
    float z[4];
    ro_normal4(ro_philox(j >> 2, a, b, 0u, (uint32_t)seed, (uint32_t)(seed >> 32)), z);
    return z[j & 3];
}

__global__ void __launch_bounds__(MQ_NT)
randn_philox_kernel(uint64_t seed, uint32_t a, uint32_t b0, const int32_t *__restrict__ counter, int64_t n,
                    float *__restrict__ out) {
    const uint32_t b = b0 + (counter ? (uint32_t)counter[0] : 0u);
    const int64_t nblk = (n + 3) / 4;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nblk; q += (int64_t)gridDim.x * blockDim.x) {
        float z[4];
        ro_normal4(ro_philox((uint32_t)q, a, b, 0u, (uint32_t)seed, (uint32_t)(seed >> 32)), z);
        for (int k = 0; k < 4; ++k)
            if (4 * q + k < n) out[4 * q + k] = z[k];
    }
}

extern "C" int mq_randn_philox(uint64_t seed, uint32_t stream_a, uint32_t stream_b, const int32_t *counter, int64_t n,
                               float *out, void *stream) {
    MQ_REQUIRE(n >= 0 && n < (1LL << 33), MQ_ERR_SHAPE, "mq_randn_philox: bad size");
    if (n == 0) return MQ_OK;
    MQ_REQUIRE(out, MQ_ERR_INVALID, "mq_randn_philox: null pointer");
    int64_t nb = mq_ceil_div(mq_ceil_div(n, 4), MQ_NT);
    const int64_t cap = (int64_t)mq_sm_count() * 8;
    if (nb > cap) nb = cap;
    MQ_LAUNCH(randn_philox_kernel, (unsigned)nb, MQ_NT, 0, stream, seed, stream_a, stream_b, counter, n, out);
    MQ_CHECK_LAUNCH("mq_randn_philox");
    return MQ_OK;
}

// ---- observation normaliser: one CTA, fp64, fixed summation order --------------------------------------
// state: mean[d], var[d], tw_mean, tw_var, n_samples (doubles)
__global__ void __launch_bounds__(RO_NT)
obsnorm_step_kernel(const float *__restrict__ raw_g, int rows, int d, double decay, double *__restrict__ state, float clip,
                    float *__restrict__ out, int stage) {
    MQ_SHARED_BYTES(dyn);
    __shared__ double part[RO_NT];
    __shared__ double s_old[RO_MAXD], s_new[RO_MAXD], s_inv[RO_MAXD];
    __shared__ int s_apply;
    const int tid = threadIdx.x;
    const int rpi = RO_NT / d;                          // rows handled per sweep; thread = (row slot, column)
    const int c = tid % d, slot = tid / d;
    const bool active = slot < rpi;
    double *mean = state, *var = state + d, *tw = state + 2 * d;
    // the three sweeps below read the batch from shared memory when it fits (one coalesced, latency-overlapped copy
    // instead of three dependent trips to L2)
    const float *raw = raw_g;
    if (stage) {
        float *sm = reinterpret_cast<float *>(dyn);
        const int n = rows * d;
#pragma unroll 8
        for (int i = tid; i < n; i += RO_NT) sm[i] = raw_g[i];
        raw = sm;
        __syncthreads();
    }
    // pass 1: batch mean per column
    double acc = 0.0;
    if (active)
        for (int r = slot; r < rows; r += rpi) acc += (double)raw[(int64_t)r * d + c];
    part[tid] = acc;
    __syncthreads();
    if (tid < d) {
        double s = 0.0;
        for (int k = 0; k < rpi; ++k) s += part[k * d + tid];
        const double bm = s / (double)rows;
        // RunningMean.update(batch mean, weight 1): _normalize.py:9-25
        const double twm = tw[0] * decay + 1.0;
        const double mo = mean[tid];
        const double mn = mo + (1.0 / twm) * (bm - mo);
        s_old[tid] = mo; s_new[tid] = mn;
        mean[tid] = mn;
    }
    __syncthreads();
    // pass 2: mean of (x - mean_old)(x - mean_new) per column
    acc = 0.0;
    if (active) {
        const double mo = s_old[c], mn = s_new[c];
        for (int r = slot; r < rows; r += rpi) {
            const double x = (double)raw[(int64_t)r * d + c];
            acc += (x - mo) * (x - mn);
        }
    }
    part[tid] = acc;
    __syncthreads();
    const double n_before = tw[2];
    if (tid < d) {
        double s = 0.0;
        for (int k = 0; k < rpi; ++k) s += part[k * d + tid];
        double cov = s / (double)rows;
        double v = var[tid];
        if (n_before >= 1.0) {                          // _normalize.py:48-49
            const double twv = tw[1] * decay + 1.0;
            v += (1.0 / twv) * (cov - v);
        } else if (rows >= 2) {                         // first batch: _normalize.py:50-55
            const double w = ((double)rows - 1.0) / (double)rows;
            const double twv = tw[1] * pow(decay, w) + w;
            v += (w / twv) * (cov / w - v);
        }
        var[tid] = v;
        s_inv[tid] = 1.0 / fmax(1e-8, sqrt(v));
    }
    __syncthreads();
    if (tid == 0) {
        tw[0] = tw[0] * decay + 1.0;
        if (n_before >= 1.0) tw[1] = tw[1] * decay + 1.0;
        else if (rows >= 2) { const double w = ((double)rows - 1.0) / (double)rows; tw[1] = tw[1] * pow(decay, w) + w; }
        tw[2] = n_before + (double)rows;
        s_apply = (n_before + (double)rows) >= 2.0;
    }
    __syncthreads();
    // apply + clip (gym.py:70-71); zeros while fewer than two samples have been seen (_normalize.py:60-61)
    if (active) {
        const double mn = s_new[c], inv = s_inv[c];
        const bool ap = s_apply != 0;
        for (int r = slot; r < rows; r += rpi) {
            double y = 0.0;
            if (ap) {
                y = ((double)raw[(int64_t)r * d + c] - mn) * inv;
                y = fmin(fmax(y, -(double)clip), (double)clip);
            }
            out[(int64_t)r * d + c] = (float)y;
        }
    }
}

#ifndef MQ_EMU
// The same update for a few thousand rows as ONE thread-block cluster: 8 CTAs take 1/8 of the rows each, exchange their
// per-column partial sums through distributed shared memory (every CTA adds the eight partials in rank order, so all
// of them hold bit-identical statistics), and normalise their own rows.  Two cluster barriers instead of a second and
// third kernel launch; the statistics never leave the chip.
#include <cooperative_groups.h>
namespace cg = cooperative_groups;
#define RO_CL 8
#define RO_CNT 256
__global__ void __cluster_dims__(RO_CL, 1, 1) __launch_bounds__(RO_CNT)
obsnorm_cluster_kernel(const float *__restrict__ raw_g, int rows, int d, double decay, double *__restrict__ state,
                       float clip, float *__restrict__ out) {
    MQ_SHARED_BYTES(dyn);
    __shared__ double part[RO_CNT];
    __shared__ double psum[RO_MAXD], pcov[RO_MAXD], s_old[RO_MAXD], s_new[RO_MAXD], s_inv[RO_MAXD];
    cg::cluster_group cl = cg::this_cluster();
    const int rank = (int)cl.block_rank();
    const int tid = threadIdx.x;
    const int per = (rows + RO_CL - 1) / RO_CL;
    const int r0 = rank * per < rows ? rank * per : rows;
    const int nr = (r0 + per <= rows ? per : rows - r0);
    float *sm = reinterpret_cast<float *>(dyn);
    {
        const int n = nr * d;
        const float *src = raw_g + (int64_t)r0 * d;
        const uint32_t bytes = (uint32_t)n * 4u;
        // The slice is contiguous: when it is 16-byte aligned and sized it arrives as bulk asynchronous copies (the TMA
        // engine: cp.async.bulk global -> shared, completion counted in bytes on an mbarrier) issued by one thread, instead
        // of ~22 loads and stores per thread; otherwise a plain strided loop.
        __shared__ __align__(8) uint64_t s_bar;
        const bool bulk = bytes >= 16u && (bytes & 15u) == 0u && (reinterpret_cast<uintptr_t>(src) & 15u) == 0u;
        if (bulk) {
            const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&s_bar);
            if (tid == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1u));
                asm volatile("fence.mbarrier_init.release.cluster;");
            }
            __syncthreads();
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm);
                for (uint32_t off = 0; off < bytes; off += 32768u) {
                    const uint32_t len = bytes - off < 32768u ? bytes - off : 32768u;
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(dst + off), "l"(reinterpret_cast<const char *>(src) + off), "r"(len), "r"(bar) : "memory");
                }
            }
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(0u) : "memory");
        } else {
#pragma unroll 8
            for (int i = tid; i < n; i += RO_CNT) sm[i] = src[i];
        }
    }
    const int rpi = RO_CNT / d;
    const int c = tid % d, slot = tid / d;
    const bool active = slot < rpi;
    double *mean = state, *var = state + d, *tw = state + 2 * d;
    // everything read from `state` is read before the first cluster barrier; rank 0 writes it after the second
    const double tw0 = tw[0], tw1 = tw[1], n_before = tw[2];
    const double mo = tid < d ? mean[tid] : 0.0, v0 = tid < d ? var[tid] : 0.0;
    __syncthreads();
    double acc = 0.0;
    if (active)
        for (int r = slot; r < nr; r += rpi) acc += (double)sm[r * d + c];
    part[tid] = acc;
    __syncthreads();
    if (tid < d) {
        double s = 0.0;
        for (int k = 0; k < rpi; ++k) s += part[k * d + tid];
        psum[tid] = s;
    }
    cl.sync();
    if (tid < d) {
        double s = 0.0;
        for (int r = 0; r < RO_CL; ++r) s += *cl.map_shared_rank(&psum[tid], r);
        const double bm = s / (double)rows;
        const double twm = tw0 * decay + 1.0;
        const double mn = mo + (1.0 / twm) * (bm - mo);
        s_old[tid] = mo; s_new[tid] = mn;
        if (rank == 0) mean[tid] = mn;
    }
    __syncthreads();
    acc = 0.0;
    if (active) {
        const double o = s_old[c], nw = s_new[c];
        for (int r = slot; r < nr; r += rpi) {
            const double x = (double)sm[r * d + c];
            acc += (x - o) * (x - nw);
        }
    }
    part[tid] = acc;
    __syncthreads();
    if (tid < d) {
        double s = 0.0;
        for (int k = 0; k < rpi; ++k) s += part[k * d + tid];
        pcov[tid] = s;
    }
    cl.sync();
    if (tid < d) {
        double s = 0.0;
        for (int r = 0; r < RO_CL; ++r) s += *cl.map_shared_rank(&pcov[tid], r);
        const double cov = s / (double)rows;
        double v = v0;
        if (n_before >= 1.0) {
            const double twv = tw1 * decay + 1.0;
            v += (1.0 / twv) * (cov - v);
        } else if (rows >= 2) {
            const double w = ((double)rows - 1.0) / (double)rows;
            const double twv = tw1 * pow(decay, w) + w;
            v += (w / twv) * (cov / w - v);
        }
        if (rank == 0) var[tid] = v;
        s_inv[tid] = 1.0 / fmax(1e-8, sqrt(v));
    }
    if (tid == 0 && rank == 0) {
        tw[0] = tw0 * decay + 1.0;
        if (n_before >= 1.0) tw[1] = tw1 * decay + 1.0;
        else if (rows >= 2) { const double w = ((double)rows - 1.0) / (double)rows; tw[1] = tw1 * pow(decay, w) + w; }
        tw[2] = n_before + (double)rows;
    }
    cl.sync();                          // nobody leaves (or reuses pcov) while a peer still reads its shared memory
    const bool ap = (n_before + (double)rows) >= 2.0;
    if (active) {
        const double nw = s_new[c], inv = s_inv[c];
        for (int r = slot; r < nr; r += rpi) {
            double y = 0.0;
            if (ap) {
                y = ((double)sm[r * d + c] - nw) * inv;
                y = fmin(fmax(y, -(double)clip), (double)clip);
            }
            out[(int64_t)(r0 + r) * d + c] = (float)y;
        }
    }
}
#endif

extern "C" int mq_obsnorm_step(const float *raw, int64_t rows, int32_t d, double horizon, double *state, float clip,
                               float *out, void *stream) {
    MQ_REQUIRE(rows >= 1 && rows < (1LL << 30) && d >= 1 && d <= RO_MAXD, MQ_ERR_SHAPE, "mq_obsnorm_step: bad shape");
    MQ_REQUIRE(raw && state && out, MQ_ERR_INVALID, "mq_obsnorm_step: null pointer");
    const double decay = horizon > 0.0 ? 1.0 - 1.0 / horizon : 1.0;
#ifndef MQ_EMU
    if (rows >= 512) {
        const int64_t per = mq_ceil_div(rows, RO_CL);
        const int64_t cbytes = per * d * (int64_t)sizeof(float);
        if (cbytes + 24 * 1024 <= mq_smem_optin()) {
            cudaError_t e_ = cudaFuncSetAttribute(obsnorm_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cbytes);
            if (e_ != cudaSuccess) { mq_set_error("mq_obsnorm_step: %s", cudaGetErrorString(e_)); return MQ_ERR_CUDA; }
            MQ_LAUNCH(obsnorm_cluster_kernel, RO_CL, RO_CNT, (size_t)cbytes, stream, raw, (int)rows, (int)d, decay, state,
                      clip, out);
            MQ_CHECK_LAUNCH("mq_obsnorm_step");
            return MQ_OK;
        }
    }
#endif
    const int64_t bytes = rows * d * (int64_t)sizeof(float);
    const int stage = bytes + 24 * 1024 <= mq_smem_optin();
#ifndef MQ_EMU
    if (stage) {
        cudaError_t e_ = cudaFuncSetAttribute(obsnorm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e_ != cudaSuccess) { mq_set_error("mq_obsnorm_step: %s", cudaGetErrorString(e_)); return MQ_ERR_CUDA; }
    }
#endif
    MQ_LAUNCH(obsnorm_step_kernel, 1, RO_NT, stage ? (size_t)bytes : 0, stream, raw, (int)rows, (int)d, decay, state, clip,
              out, stage);
    MQ_CHECK_LAUNCH("mq_obsnorm_step");
    return MQ_OK;
}

// ---- actor + synthetic environment ---------------------------------------------------------------------
// Environment (per env e): hidden state x in R^d.  Observation = obs_scale * x + obs_shift (so that the normaliser has
// work to do).  Step: x' = x + dt * (-damp * x + B a) + sigma * noise;  reward = -(|x'|^2 / d) - ctrl * |a|^2 / A;
// the episode ends after `horizon` steps or, each step, with probability p_end; a finished episode restarts from
// x ~ N(0, 1).  B is d x A, row-major.
#define RO_ENV_NT 64             // environments per CTA: 4,096 environments spread over 64 SMs
// deviates [j0, j0 + n) of stream (seed, a, b), one Philox block per four of them (j0 a multiple of 4)
__device__ __forceinline__ void ro_normals(uint64_t seed, uint32_t a, uint32_t b, uint32_t j0, int n, float *z) {
    for (int q = 0; q * 4 < n; ++q) {
        float t[4];
        ro_normal4(ro_philox((j0 >> 2) + (uint32_t)q, a, b, 0u, (uint32_t)seed, (uint32_t)(seed >> 32)), t);
        for (int k = 0; k < 4; ++k)
            if (q * 4 + k < n) z[q * 4 + k] = t[k];
    }
}

__global__ void __launch_bounds__(RO_ENV_NT)
actor_env_step_kernel(mq_env_t env, int n_env, int t, int t_len, uint64_t seed, uint32_t step_id0,
                      const int32_t *__restrict__ step_base, const float *__restrict__ obs_n,
                      const float *__restrict__ mean, const float *__restrict__ logstd,
                      float *__restrict__ x, int32_t *__restrict__ ep_step,
                      float *__restrict__ obs_buf, float *__restrict__ act_buf, float *__restrict__ rew_buf,
                      uint8_t *__restrict__ done_buf, float *__restrict__ raw_next) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_env) return;
    const int d = env.d, na = env.a;
    // index of this time step in the noise stream: a device-resident base (so that a captured CUDA graph draws fresh
    // noise at every replay) + the step's offset
    const uint32_t step_id = step_id0 + (step_base ? (uint32_t)step_base[0] : 0u);
    float a[16], xs[RO_MAXD], z[RO_MAXD];
    // action = mean + exp(logstd) * unit noise (distrib.py:50-59): stream (seed, step, env), deviates 0 .. A-1
    ro_normals(seed, step_id, (uint32_t)e, 0u, na, z);
    for (int k = 0; k < na; ++k) a[k] = mean[(int64_t)e * na + k] + expf(logstd[k]) * z[k];
    const int64_t slot = (int64_t)e * t_len + t;        // environment-major rollout buffers
    for (int k = 0; k < d; ++k) obs_buf[slot * d + k] = obs_n[(int64_t)e * d + k];
    for (int k = 0; k < na; ++k) act_buf[slot * na + k] = a[k];
    // environment: process noise = deviates 16 .. 16+d-1
    ro_normals(seed, step_id, (uint32_t)e, 16u, d, z);
    float cost = 0.0f, ctrl = 0.0f;
    for (int k = 0; k < na; ++k) ctrl += a[k] * a[k];
    for (int i = 0; i < d; ++i) {
        float f = -env.damp * x[(int64_t)e * d + i];
        for (int k = 0; k < na; ++k) f = fmaf(env.b[i * na + k], a[k], f);
        const float xn = x[(int64_t)e * d + i] + env.dt * f + env.sigma * z[i];
        xs[i] = xn;
        cost += xn * xn;
    }
    const float rew = -(cost / (float)d) - env.ctrl * ctrl / (float)na;
    const int steps = ep_step[e] + 1;
    const RoPhilox u = ro_philox(0xFFFFFFFFu, step_id, (uint32_t)e, 1u, (uint32_t)seed, (uint32_t)(seed >> 32));
    const bool done = steps >= env.horizon || (float)(u.v[0] >> 8) * (1.0f / 16777216.0f) < env.p_end;
    rew_buf[slot] = rew;
    done_buf[slot] = done ? 1 : 0;
    ep_step[e] = done ? 0 : steps;
    if (done) ro_normals(seed, step_id, (uint32_t)e, 16u + RO_MAXD, d, xs);      // restart: deviates 80 .. 80+d-1
    for (int i = 0; i < d; ++i) {
        x[(int64_t)e * d + i] = xs[i];
        raw_next[(int64_t)e * d + i] = env.obs_scale[i] * xs[i] + env.obs_shift[i];
    }
}

extern "C" int mq_actor_env_step(const mq_env_t *env, int64_t n_env, int32_t t, int32_t t_len, uint64_t seed,
                                 uint32_t step_id, const int32_t *step_base, const float *obs_n, const float *mean,
                                 const float *logstd,
                                 float *x, int32_t *ep_step, float *obs_buf, float *act_buf, float *rew_buf,
                                 uint8_t *done_buf, float *raw_next, void *stream) {
    MQ_REQUIRE(env && env->d >= 1 && env->d <= RO_MAXD && env->a >= 1 && env->a <= 16 && env->horizon >= 1,
               MQ_ERR_SHAPE, "mq_actor_env_step: bad environment");
    MQ_REQUIRE(n_env >= 1 && n_env < (1LL << 30) && t >= 0 && t < t_len, MQ_ERR_SHAPE, "mq_actor_env_step: bad sizes");
    MQ_REQUIRE(env->b && env->obs_scale && env->obs_shift && obs_n && mean && logstd && x && ep_step && obs_buf &&
                   act_buf && rew_buf && done_buf && raw_next,
               MQ_ERR_INVALID, "mq_actor_env_step: null pointer");
    MQ_LAUNCH(actor_env_step_kernel, (unsigned)mq_ceil_div(n_env, RO_ENV_NT), RO_ENV_NT, 0, stream, *env, (int)n_env, (int)t,
              (int)t_len, seed, step_id, step_base, obs_n, mean, logstd, x, ep_step, obs_buf, act_buf, rew_buf, done_buf, raw_next);
    MQ_CHECK_LAUNCH("mq_actor_env_step");
    return MQ_OK;
}

// initial hidden states and raw observations: x ~ N(0,1) from stream (seed, 0xFFFFFFFE, env)
__global__ void __launch_bounds__(MQ_NT)
env_reset_kernel(mq_env_t env, int n_env, uint64_t seed, float *__restrict__ x, int32_t *__restrict__ ep_step,
                 float *__restrict__ raw) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_env) return;
    for (int i = 0; i < env.d; ++i) {
        const float xn = ro_normal(seed, 0xFFFFFFFEu, (uint32_t)e, (uint32_t)i);
        x[(int64_t)e * env.d + i] = xn;
        raw[(int64_t)e * env.d + i] = env.obs_scale[i] * xn + env.obs_shift[i];
    }
    ep_step[e] = 0;
}

extern "C" int mq_env_reset(const mq_env_t *env, int64_t n_env, uint64_t seed, float *x, int32_t *ep_step, float *raw,
                            void *stream) {
    MQ_REQUIRE(env && env->d >= 1 && env->d <= RO_MAXD, MQ_ERR_SHAPE, "mq_env_reset: bad environment");
    MQ_REQUIRE(n_env >= 1 && n_env < (1LL << 30), MQ_ERR_SHAPE, "mq_env_reset: bad sizes");
    MQ_REQUIRE(env->obs_scale && env->obs_shift && x && ep_step && raw, MQ_ERR_INVALID, "mq_env_reset: null pointer");
    MQ_LAUNCH(env_reset_kernel, (unsigned)mq_ceil_div(n_env, MQ_NT), MQ_NT, 0, stream, *env, (int)n_env, seed, x, ep_step, raw);
    MQ_CHECK_LAUNCH("mq_env_reset");
    return MQ_OK;
}

// ---- chunk boundary of the per-environment segments (gae.py:43-64) -----------------------------------------
__global__ void __launch_bounds__(MQ_NT)
bootstrap_segments_kernel(float *__restrict__ rew, uint8_t *__restrict__ done, const float *__restrict__ v_next, int n_seg,
                          int seg_len, float gam) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_seg) return;
    const int64_t i = (int64_t)e * seg_len + seg_len - 1;
    if (!done[i]) { rew[i] += gam * v_next[e]; done[i] = 1; }
}

extern "C" int mq_bootstrap_segments(float *rew, uint8_t *done, const float *v_next, int64_t n_seg, int64_t seg_len,
                                     float gam, void *stream) {
    MQ_REQUIRE(n_seg >= 1 && seg_len >= 1 && n_seg < (1LL << 30) && n_seg * seg_len < (1LL << 40), MQ_ERR_SHAPE,
               "mq_bootstrap_segments: bad sizes");
    MQ_REQUIRE(rew && done && v_next, MQ_ERR_INVALID, "mq_bootstrap_segments: null pointer");
    MQ_LAUNCH(bootstrap_segments_kernel, (unsigned)mq_ceil_div(n_seg, MQ_NT), MQ_NT, 0, stream, rew, done, v_next,
              (int)n_seg, (int)seg_len, gam);
    MQ_CHECK_LAUNCH("mq_bootstrap_segments");
    return MQ_OK;
}
