// One optimiser step after the gradient kernel, in ONE launch:
//   fixed-order sum of the per-CTA partial gradients  ->  (multi-GPU) all-reduce over NVLink peer memory
//   ->  two-moment (Adam) update  ->  float32 model copy.
//
// Replaces mannequin/_adam.py:23-53 + mannequin/_normalize.py:9-25 (one apply_gradient call) and, for
// sharded minibatches, the gradient exchange of SURVEY.md 8e, which the reference does not have: every
// rank holds the same parameters and Adam state and applies the identical update to the identical
// summed gradient.
//
// The exchange is a one-shot all-reduce written into the kernel (no NCCL call on this path): a block owns 32 parameters
// and only ever waits for the same block index on the other GPUs, never for another block of its own GPU.  Buffers
// live in symmetric memory mapped into every peer and are double-buffered by step parity (a rank can be at most one
// step ahead of the slowest).  Three forms, all adding the ranks' values in RANK ORDER (bit-identical sums everywhere):
//   low-latency (default, peer.pull == 2): (value, step) as ONE 8-byte word per parameter stored into slot [rank] of every
//       peer's buffer; the receiver polls the words in its own memory until their tag is this step's.  An aligned 8-byte
//       store is atomic: no flags, no fences, one one-way NVLink trip.
//   push (0): plain floats into every peer's slot, one st.release.sys flag per peer, ld.acquire.sys spin, local reads.
//   pull (1): local store + fence, flags, remote loads of every peer's slot.
#include "mq_tc3.cuh"

#ifndef MQ_EMU

template <bool ADAMS>
__device__ __forceinline__ void step_elem(double &val, double &m, double &v, double g, double c1, double c2, double lr,
                                          double eps) {
    m = __dadd_rn(m, __dmul_rn(c1, __dsub_rn(g, m)));
    v = __dadd_rn(v, __dmul_rn(c2, __dsub_rn(__dmul_rn(g, g), v)));
    const double den = __dadd_rn(eps, ADAMS ? v : __dsqrt_rn(v));
    val = __dadd_rn(val, __dmul_rn(lr, __ddiv_rn(m, den)));
}
template <bool ADAMS>
__device__ __forceinline__ void step_elem(float &val, float &m, float &v, float g, float c1, float c2, float lr, float eps) {
    m += c1 * (g - m);
    v += c2 * (g * g - v);
    const float den = eps + (ADAMS ? v : sqrtf(v));
    val += lr * (m / den);
}

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float *p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

// One parameter vector's step.  A launch carries up to TWO (blocks [0, nblk) of job 0, then job 1's): the value and the
// policy net of a PPO update step together -- one launch and, sharded, ONE exchange latency per pair of steps.
template <typename ST>
struct StepJob {
    const float *partial; int n_parts, n; float scale;
    ST *value, *m, *v; float *theta_out, *grad_out;
    ST c1, c2, lr, eps;
    mq_peer_t peer; T3Img img; unsigned char *image;
    int nblk;
};
template <typename ST> struct StepJobs { StepJob<ST> job[2]; };

// 32 parameters x 8 groups of partial vectors per block; group sums combined in fixed order.
template <typename ST, bool ADAMS>
__global__ void __launch_bounds__(MQ_NT)
reduce_step_kernel(const __grid_constant__ StepJobs<ST> launch) {
    const bool second = blockIdx.x >= (unsigned)launch.job[0].nblk;
    const StepJob<ST> &J = second ? launch.job[1] : launch.job[0];
    const int bid = (int)blockIdx.x - (second ? launch.job[0].nblk : 0);
    const float *__restrict__ partial = J.partial;
    const int n_parts = J.n_parts, n = J.n;
    const float scale = J.scale;
    ST *__restrict__ value = J.value, *__restrict__ m = J.m, *__restrict__ v = J.v;
    float *__restrict__ theta_out = J.theta_out, *__restrict__ grad_out = J.grad_out;
    const ST c1 = J.c1, c2 = J.c2, lr = J.lr, eps = J.eps;
    const mq_peer_t &peer = J.peer;
    unsigned char *__restrict__ image = J.image;
    __shared__ float part[8][33];
    const int jx = threadIdx.x & 31, gy = threadIdx.x >> 5;
    const int j = bid * 32 + jx;
    // programmatic dependent launch: this grid may already be resident while the gradient kernel finishes, and the
    // next gradient kernel may start its on-chip prologue while this one runs
    mq_pdl_trigger();
    mq_pdl_wait();
    // group gy sums partial rows gy, gy + 8, ...: 20 independent loads in flight per thread (the whole of a 148-CTA
    // gradient in one round trip to L2), added in a fixed tree
    float tot = 0.0f;
    if (j < n) {
        for (int base = gy; base < n_parts; base += 160) {
            float v[20];
#pragma unroll
            for (int k = 0; k < 20; ++k) {
                const int b = base + 8 * k;
                v[k] = b < n_parts ? partial[(int64_t)b * n + j] : 0.0f;
            }
#pragma unroll
            for (int k = 0; k < 10; ++k) v[k] += v[k + 10];
#pragma unroll
            for (int k = 0; k < 5; ++k) v[k] += v[k + 5];
            tot += ((v[0] + v[1]) + (v[2] + v[3])) + v[4];
        }
    }
    part[gy][jx] = tot;
    __syncthreads();
    if (gy != 0) return;
    float g = part[0][jx];
#pragma unroll
    for (int k = 1; k < 8; ++k) g += part[k][jx];
    if (peer.world > 1 && peer.pull == 2) {
        // ---- low-latency form: value and step tag travel together in ONE 8-byte word (an aligned 8-byte store is
        // atomic), pushed into slot [rank] of every peer's buffer; the receiver polls the words in its own memory until
        // the tag is this step's.  One one-way NVLink trip: no flags, no fences, no release / acquire.
        const int par = (int)(peer.step & 1u);
        const int64_t slot = ((int64_t)par * peer.world + peer.rank) * peer.buf_stride;
        if (j < n) {
            const unsigned long long w = ((unsigned long long)peer.step << 32) | (unsigned long long)__float_as_uint(g);
            for (int r = 0; r < peer.world; ++r) {
                unsigned long long *dst = reinterpret_cast<unsigned long long *>(peer.grad_bufs[r]) + slot + j;
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(w) : "memory");
            }
            g = 0.0f;
            const unsigned long long *mine =
                reinterpret_cast<const unsigned long long *>(peer.grad_bufs[peer.rank]) + (int64_t)par * peer.world * peer.buf_stride;
            for (int r = 0; r < peer.world; ++r) {                       // rank order: bit-identical on every rank
                unsigned long long v;
                do {
                    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine + (int64_t)r * peer.buf_stride + j) : "memory");
                } while ((unsigned)(v >> 32) != peer.step);
                g += __uint_as_float((unsigned)v);
            }
        }
    } else if (peer.world > 1) {
        // ---- one-shot all-reduce over peer memory (push): every rank STORES its 32 sums into slot [rank] of every
        // peer's exchange buffer (remote stores need no round trip), then releases one flag per peer; a rank only ever
        // reads its own memory
        const int par = (int)(peer.step & 1u);
        const int64_t slot = ((int64_t)par * peer.world + peer.rank) * peer.buf_stride;
        if (peer.pull) {                // pull: the own slot only, made visible to the peers that will load it
            if (j < n) peer.grad_bufs[peer.rank][slot + j] = g;
            __threadfence_system();
        } else if (j < n) {
            for (int r = 0; r < peer.world; ++r) peer.grad_bufs[r][slot + j] = g;
        }
        __syncwarp();                   // the lanes' stores happen before the releasing lanes' flag stores (cumulative)
        const int nblk = J.nblk;
        if (jx < peer.world)            // lane r tells rank r that this rank's block has arrived there
            st_release_sys(peer.flag_bufs[jx] + ((int64_t)par * peer.world + peer.rank) * nblk + bid, peer.step);
        if (jx < peer.world) {          // lane r waits for rank r's block
            const uint32_t *f = peer.flag_bufs[peer.rank] + ((int64_t)par * peer.world + jx) * nblk + bid;
            while (ld_acquire_sys(f) != peer.step) { }
        }
        __syncwarp();
        g = 0.0f;
        if (j < n) {
            const int64_t row0 = (int64_t)par * peer.world * peer.buf_stride;
            if (peer.pull) {            // every rank's own slot, through the peer mappings
                for (int r = 0; r < peer.world; ++r)
                    g += ld_relaxed_sys(peer.grad_bufs[r] + row0 + (int64_t)r * peer.buf_stride + j);             // rank order
            } else {
                const float *mine = peer.grad_bufs[peer.rank] + row0;
                for (int r = 0; r < peer.world; ++r) g += ld_relaxed_sys(mine + (int64_t)r * peer.buf_stride + j);   // rank order
            }
        }
    }
    if (j >= n) return;
    g *= scale;
    if (grad_out) grad_out[j] = g;
    ST a = value[j], b = m[j], c = v[j];
    step_elem<ADAMS>(a, b, c, (ST)g, c1, c2, lr, eps);
    value[j] = a; m[j] = b; v[j] = c;
    if (theta_out) theta_out[j] = (float)a;
    // the operand image of the tensor-core gradient kernel follows theta: the next launch copies it, converts nothing
    if (image) t3_img_store_param_rt(J.img, image, j, (float)a);
}

extern "C" int64_t mq_step_blocks(int64_t n) { return mq_ceil_div(n, 32); }

template <typename ST>
static int step_job_fill(const mq_step_t *t, StepJob<ST> *j) {
    MQ_REQUIRE(t->partials && t->value && t->m && t->v, MQ_ERR_INVALID, "mq_reduce_step: null pointer");
    MQ_REQUIRE(t->n_parts >= 1 && t->n >= 1 && t->n < (1LL << 31), MQ_ERR_SHAPE, "mq_reduce_step: sizes");
    memset(j, 0, sizeof(*j));
    j->peer.world = 1;
    if (t->peer && t->peer->world > 1) {
        j->peer = *t->peer;
        MQ_REQUIRE(j->peer.world <= 32 && j->peer.rank >= 0 && j->peer.rank < j->peer.world && j->peer.grad_bufs && j->peer.flag_bufs &&
                       j->peer.buf_stride >= t->n && j->peer.step != 0,
                   MQ_ERR_INVALID, "mq_reduce_step: bad peer descriptor");
    }
    if (t->image) {
        const int planes = (t->tc & MQ_TC_PLANES_MASK) ? (t->tc & MQ_TC_PLANES_MASK) : 3;
        MQ_REQUIRE(t->net && t->net->n_params == t->n && t3_img_build_layout(t->net, planes, (t->tc & MQ_TC_FP16) ? 1 : 0, &j->img), MQ_ERR_INVALID,
                   "mq_reduce_step: the net does not match the parameter vector or does not fit the tensor-core kernel");
    }
    j->partial = t->partials; j->n_parts = t->n_parts; j->n = (int)t->n; j->scale = t->scale;
    j->value = (ST *)t->value; j->m = (ST *)t->m; j->v = (ST *)t->v;
    j->theta_out = t->theta_out; j->grad_out = t->grad_out;
    j->c1 = (ST)t->c1; j->c2 = (ST)t->c2; j->lr = (ST)t->lr; j->eps = (ST)t->eps;
    j->image = (unsigned char *)t->image;
    j->nblk = (int)mq_ceil_div(t->n, 32);
    return MQ_OK;
}

template <typename ST>
static int step_launch(const mq_step_t *a, const mq_step_t *b, int adams, void *stream) {
    StepJobs<ST> jobs;
    int rc = step_job_fill<ST>(a, &jobs.job[0]);
    if (rc != MQ_OK) return rc;
    int blocks = jobs.job[0].nblk;
    if (b) {
        rc = step_job_fill<ST>(b, &jobs.job[1]);
        if (rc != MQ_OK) return rc;
        blocks += jobs.job[1].nblk;
    } else {
        memset(&jobs.job[1], 0, sizeof(jobs.job[1]));
    }
    cudaError_t e = adams ? mq_launch_pdl(reduce_step_kernel<ST, true>, dim3(blocks), dim3(MQ_NT), 0, stream, jobs)
                          : mq_launch_pdl(reduce_step_kernel<ST, false>, dim3(blocks), dim3(MQ_NT), 0, stream, jobs);
    if (e != cudaSuccess) { mq_set_error("mq_reduce_step: %s", cudaGetErrorString(e)); return MQ_ERR_CUDA; }
    MQ_CHECK_LAUNCH("mq_reduce_step");
    return MQ_OK;
}

extern "C" int mq_reduce_step2(const mq_step_t *a, const mq_step_t *b, int state_dtype, int adams, void *stream) {
    MQ_REQUIRE(a, MQ_ERR_INVALID, "mq_reduce_step2: null step");
    if (state_dtype == MQ_F32) return step_launch<float>(a, b, adams, stream);
    if (state_dtype == MQ_F64) return step_launch<double>(a, b, adams, stream);
    mq_set_error("mq_reduce_step: state dtype %d", state_dtype);
    return MQ_ERR_INVALID;
}

extern "C" int mq_reduce_step_image(const float *partials, int32_t n_parts, int64_t n, float scale, void *value, void *m,
                                    void *v, float *theta_out, float *grad_out, int state_dtype, double c1, double c2,
                                    double lr, double eps, int adams, const mq_peer_t *peer, const mq_net_t *net,
                                    int32_t tc, void *image, void *stream) {
    mq_step_t t;
    memset(&t, 0, sizeof(t));
    t.partials = partials; t.n_parts = n_parts; t.n = n; t.scale = scale; t.value = value; t.m = m; t.v = v;
    t.theta_out = theta_out; t.grad_out = grad_out; t.c1 = c1; t.c2 = c2; t.lr = lr; t.eps = eps; t.peer = peer;
    t.net = net; t.tc = tc; t.image = image;
    return mq_reduce_step2(&t, nullptr, state_dtype, adams, stream);
}

extern "C" int mq_reduce_step(const float *partials, int32_t n_parts, int64_t n, float scale, void *value, void *m,
                              void *v, float *theta_out, float *grad_out, int state_dtype, double c1, double c2,
                              double lr, double eps, int adams, const mq_peer_t *peer, void *stream) {
    return mq_reduce_step_image(partials, n_parts, n, scale, value, m, v, theta_out, grad_out, state_dtype, c1, c2, lr, eps,
                                adams, peer, nullptr, 0, nullptr, stream);
}

#else   // MQ_EMU
extern "C" int64_t mq_step_blocks(int64_t n) { return mq_ceil_div(n, 32); }
extern "C" int mq_reduce_step(const float *, int32_t, int64_t, float, void *, void *, void *, float *, float *, int, double,
                              double, double, double, int, const mq_peer_t *, void *) { return MQ_ERR_UNSUPPORTED; }
extern "C" int mq_reduce_step_image(const float *, int32_t, int64_t, float, void *, void *, void *, float *, float *, int,
                                    double, double, double, double, int, const mq_peer_t *, const mq_net_t *, int32_t, void *,
                                    void *) { return MQ_ERR_UNSUPPORTED; }
extern "C" int mq_reduce_step2(const mq_step_t *, const mq_step_t *, int, int, void *) { return MQ_ERR_UNSUPPORTED; }
#endif

// ---- a block of dependent optimiser steps issued by ONE call ------------------------------------------------------------
// benchmarks/ppo.py:29-40 / benchmarks/gae.py:71-73 at large minibatches: per step one gradient launch (per-CTA partials)
// and one reduce / exchange / Adam launch.  The loop lives here so that a 160-step epoch block costs one trip through
// the Python binding instead of 320 (at 8,192 rows per GPU a step is ~25 us of GPU time: the host has to keep ahead).
// Step s reads the mb indices at idx_table + s * idx_stride (idx_stride > mb: this rank's columns of a wider global table).
// coefs_host = [n_steps][2] running-mean coefficients (c1, c2) of mannequin/_normalize.py:18-25; peer (nullable) carries the
// tag of the FIRST step, later steps count up; image (nullable): operand image kept in step with theta (mq_tc3.cuh).
extern "C" int mq_train_steps(const mq_net_t *net, float *theta, const float *x, const mq_head_t *head,
                              const int32_t *idx_table, int64_t idx_stride, int32_t n_steps, int64_t mb, float scale, void *value, void *m,
                              void *v, int state_dtype, const double *coefs_host, double lr, double eps, int adams,
                              const mq_peer_t *peer, void *ws, int64_t ws_bytes, void *stream) {
    MQ_REQUIRE(net && theta && x && head && idx_table && coefs_host && ws, MQ_ERR_INVALID, "mq_train_steps: null pointer");
    MQ_REQUIRE(n_steps >= 0 && mb >= 1 && idx_stride >= mb, MQ_ERR_SHAPE, "mq_train_steps: sizes");
    mq_peer_t pr;
    if (peer) pr = *peer;
    for (int s = 0; s < n_steps; ++s) {
        int32_t n_parts = 0;
        int rc = mq_fused_grad_partials(net, theta, x, idx_table + (int64_t)s * idx_stride, mb, head, nullptr, nullptr, ws, ws_bytes,
                                        &n_parts, stream);
        if (rc != MQ_OK) return rc;
        if (peer) pr.step = peer->step + (uint32_t)s;
        rc = mq_reduce_step_image(ws == nullptr ? nullptr : (const float *)ws, n_parts, net->n_params, scale, value, m, v, theta,
                                  nullptr, state_dtype, coefs_host[2 * s], coefs_host[2 * s + 1], lr, eps, adams,
                                  peer ? &pr : nullptr, head->tc_image ? net : nullptr, head->tc, (void *)head->tc_image, stream);
        if (rc != MQ_OK) return rc;
    }
    return MQ_OK;
}

// The same for TWO nets trained side by side on the same minibatch schedule (the value predictor and the policy of one PPO
// update): per step ONE gradient launch (each net on half of the SMs, mq_fused_grad_partials2) and ONE reduce / exchange /
// Adam launch (mq_reduce_step2).  Falls back to alternating single launches when the nets cannot share a launch.
extern "C" int mq_train_steps2(const mq_train_t *a, const mq_train_t *b, int32_t n_steps, int64_t mb, int state_dtype, int adams,
                               void *stream) {
    MQ_REQUIRE(a && b && a->net && b->net && a->theta && b->theta && a->head && b->head && a->idx_table && b->idx_table &&
                   a->coefs_host && b->coefs_host && a->ws && b->ws,
               MQ_ERR_INVALID, "mq_train_steps2: null pointer");
    MQ_REQUIRE(n_steps >= 0 && mb >= 1 && a->idx_stride >= mb && b->idx_stride >= mb, MQ_ERR_SHAPE, "mq_train_steps2: sizes");
    mq_peer_t pa, pb;
    if (a->peer) pa = *a->peer;
    if (b->peer) pb = *b->peer;
    const mq_train_t *t[2] = {a, b};
    for (int s = 0; s < n_steps; ++s) {
        int32_t parts[2] = {0, 0};
        int rc = mq_fused_grad_partials2(a->net, a->theta, a->idx_table + (int64_t)s * a->idx_stride, a->head, a->ws, a->ws_bytes,
                                         &parts[0], b->net, b->theta, b->idx_table + (int64_t)s * b->idx_stride, b->head, b->ws,
                                         b->ws_bytes, &parts[1], mb, stream);
        if (rc == MQ_ERR_UNSUPPORTED) {
            for (int k = 0; k < 2 && (rc == MQ_OK || rc == MQ_ERR_UNSUPPORTED); ++k)
                rc = mq_fused_grad_partials(t[k]->net, t[k]->theta, t[k]->x, t[k]->idx_table + (int64_t)s * t[k]->idx_stride, mb,
                                            t[k]->head, nullptr, nullptr, t[k]->ws, t[k]->ws_bytes, &parts[k], stream);
        }
        if (rc != MQ_OK) return rc;
        mq_step_t st[2];
        for (int k = 0; k < 2; ++k) {
            memset(&st[k], 0, sizeof(st[k]));
            st[k].partials = (const float *)t[k]->ws; st[k].n_parts = parts[k]; st[k].n = t[k]->net->n_params; st[k].scale = t[k]->scale;
            st[k].value = t[k]->value; st[k].m = t[k]->m; st[k].v = t[k]->v; st[k].theta_out = t[k]->theta;
            st[k].c1 = t[k]->coefs_host[2 * s]; st[k].c2 = t[k]->coefs_host[2 * s + 1]; st[k].lr = t[k]->lr; st[k].eps = t[k]->eps;
            if (t[k]->head->tc_image) { st[k].net = t[k]->net; st[k].tc = t[k]->head->tc; st[k].image = (void *)t[k]->head->tc_image; }
        }
        if (a->peer) { pa.step = a->peer->step + (uint32_t)s; st[0].peer = &pa; }
        if (b->peer) { pb.step = b->peer->step + (uint32_t)s; st[1].peer = &pb; }
        rc = mq_reduce_step2(&st[0], &st[1], state_dtype, adams, stream);
        if (rc != MQ_OK) return rc;
    }
    return MQ_OK;
}
