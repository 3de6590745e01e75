/* mq_b200.h -- C-ABI of the B200-native mannequin2 training hot path.
 *
 * The reference (squirrelinhell/mannequin2) is pure Python + NumPy and has no
 * FFI of its own; its drop-in boundary is the Python API of package
 * `mannequin`.  mannequin2_b200/ mirrors that API and binds the
 * entry points below with ctypes (mannequin2_b200/_lib.py).  Each entry point
 * names the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *     calls only enqueue work, they never synchronise and never allocate;
 *   - every function returns MQ_OK or a negative MQ_ERR_* code, never throws,
 *     never exits; mq_last_error() gives the message for the calling thread;
 *   - dtype selectors: MQ_F32 / MQ_F64.
 */
#ifndef MQ_B200_H
#define MQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MQ_OK               0
#define MQ_ERR_INVALID     (-1)  /* bad argument value            -> ValueError        */
#define MQ_ERR_SHAPE       (-2)  /* shape / size contract broken  -> AssertionError    */
#define MQ_ERR_UNSUPPORTED (-3)  /* net does not fit this kernel  -> NotImplementedError */
#define MQ_ERR_CUDA        (-4)  /* CUDA runtime failure          -> RuntimeError      */

#define MQ_F32 0
#define MQ_F64 1

#define MQ_ACT_NONE  0
#define MQ_ACT_TANH  1   /* mannequin/backprop.py:58-60 */
#define MQ_ACT_LRELU 2   /* mannequin/backprop.py:53-56, leak default 0.1 (basicnet.py:313) */

#define MQ_MAX_LAYERS 8

/* A stack of Affine(+activation) layers over ONE flat float32 parameter vector
 * in the reference's order W1(in x out, row-major), b1, W2, b2, ...
 * (mannequin/basicnet.py:13-23,35-51,303-311).  For Gauss(mean=net) the learnt
 * state-independent logstd is the LAST n_out entries (mannequin/distrib.py:43). */
typedef struct mq_net {
    int32_t n_layers;
    int32_t n_in;
    int32_t n_out;
    int32_t n_params;            /* including logstd when logstd_off >= 0 */
    int32_t logstd_off;          /* -1: no Gaussian head parameters       */
    int32_t in[MQ_MAX_LAYERS];
    int32_t out[MQ_MAX_LAYERS];
    int32_t act[MQ_MAX_LAYERS];
    float   leak[MQ_MAX_LAYERS];
    int32_t w_off[MQ_MAX_LAYERS];
    int32_t b_off[MQ_MAX_LAYERS];
} mq_net_t;

/* Loss-gradient rules evaluated inside the fused kernels ("heads"). */
#define MQ_HEAD_DY         0  /* dy[B,n_out] supplied: Function.evaluate()[1](dy), basicnet.py:209   */
#define MQ_HEAD_DIFF       1  /* dy = target - out: benchmarks/gae.py:20-23                          */
#define MQ_HEAD_SOFTMAX_CE 2  /* dy = onehot - softmax(out) - p0*out: mnist/mlp.py:10-14,31-35       */
#define MQ_HEAD_GAUSS_LOGP 3  /* d/dtheta sum_b w_b logp_b, w supplied: distrib.py:61-68, gae.py:98-99 */
#define MQ_HEAD_GAUSS_PPO  4  /* w_b = clip-rule(exp(logp-logp_old), adv)*adv: benchmarks/ppo.py:33-40 */
#define MQ_HEAD_DISCRETE_LOGP 5 /* categorical: d/dtheta sum_b w_b log softmax(out_b)[a_b], distrib.py:24-32;
                                  the action is a ONE-HOT row of `target`, w = dy                       */
#define MQ_HEAD_DISCRETE_PPO  6 /* the same with the PPO clip weights (benchmarks/ppo.py:33-40)        */

typedef struct mq_head {
    int32_t kind;
    float   p0;                /* SOFTMAX_CE: logit L2 coefficient (0.01); GAUSS_PPO: clip low (0.8) */
    float   p1;                /* GAUSS_PPO: clip high (1.2)                                         */
    float   ent;               /* GAUSS_*: entropy-bonus coefficient c of the objective mean_b(w_b logp_b) + c H,
                                  H = sum_a logstd_a + A/2 (1 + ln 2 pi) (not in the reference, whose
                                  comparator runs with entcoeff = 0): adds c to every d/dlogstd_a.  Default 0.    */
    const float *dy;           /* DY: [rows,n_out]; GAUSS_LOGP / DISCRETE_LOGP: weights [rows]        */
    const float *target;       /* DIFF/SOFTMAX_CE: [rows,n_out]; GAUSS_*: sampled actions [rows,A];
                                  DISCRETE_*: one-hot actions [rows,n_out]                            */
    const float *adv;          /* *_PPO: advantages [rows]                                           */
    const float *logp_old;     /* *_PPO: log-probabilities under the pre-update policy [rows]        */
    /* ---- per-call options of the tensor-core (tcgen05) kernels; all zero = automatic, exact ---- */
    const float *records;      /* packed rows written by mq_pack_records: [rows][rec_w] floats, row =
                                  x[n_in] | target[n_out] | w0 | w1 (w0 = adv or the LOGP weight, w1 = logp_old).
                                  When set, the tensor-core gradient kernel gathers ONE aligned record per minibatch
                                  row (two 32-byte sectors for the 11-obs / 3-act policy) instead of reading x and the
                                  head arrays separately; kernels that do not support records ignore it.            */
    int32_t rec_w;             /* floats per record: mq_record_width(net)                                           */
    int32_t tc;                /* MQ_TC_* flags                                                                      */
    const void *tc_image;      /* operand image of theta (mq_tc_image_build / mq_reduce_step_image) for the planes
                                  selected in `tc`; NULL: the launch builds it from theta into the workspace         */
} mq_head_t;
/* Precision of the tensor-core path (north_star: "a stated looser tolerance for TF32/bf16 tensor-core paths").
 * An fp32 operand is split into bf16 planes x = x0 + x1 + x2; a product uses the plane pairs i + j < planes:
 *   3 planes (default): exact fp32 products, 6 pairs            -- same bar as the FFMA kernels (rtol 1e-5)
 *   2 planes: 16 mantissa bits per operand, 3 pairs             -- |err| <= 3e-5 max|grad|
 *   1 plane : plain bf16 operands, 1 pair, MUFU tanh            -- |err| <= 2e-2 max|grad|
 * MQ_TC_FP16 with 2 planes selects fp16 planes with the low one scaled by 2^11 (x = x0 + 2^-11 x1): 22 significant
 * bits and a sign, i.e. fp32 operands to 2^-24, in THREE pairs -- the exact mode's bar (rtol 1e-5) at the two-plane
 * cost.  Precondition: |operand| < 65504 (inputs, weights, activations, output-layer gradients); beyond it the planes
 * are +-inf and the gradient comes back non-finite (never silently wrong).
 * The tolerances are the ones tests/kernel_cases.py: case_tc_modes asserts. */
#define MQ_TC_PLANES_MASK 0x3   /* 0 = default (3) */
#define MQ_TC_FP16        0x80  /* with 2 planes: fp16 planes, scaled low plane ("fp16x2")                         */
#define MQ_TC_FORCE       0x10  /* use the tensor-core kernel whenever the net fits (default: batch >= 8192) */
#define MQ_TC_OFF         0x20  /* never use it (FFMA kernels)                                                */
#define MQ_TC_NOPIPE      0x40  /* developer aid (A/B measurements): no K-chunk pipelining between layers          */
#define MQ_TC_KEEP_X      0x100 /* mq_mlp_forward_tc / mq_mlp_backward_tc: the forward pass leaves the operand planes of the
                                 * gathered input batch in the workspace and the backward pass of the SAME step (same ws, x,
                                 * row_idx, batch and mode, no other call on ws in between) reads them for dW_0 instead of
                                 * packing x^T again.  Ignored when layer 0 does not run on the tensor-core GEMM.           */
/* When row_idx != NULL, x and every head array are indexed by row_idx[i] (they
 * are whole-rollout arrays, mannequin/_trajectory.py:79-86 fused into the load);
 * otherwise by the minibatch position i. */

/* ---- library -------------------------------------------------------------- */
const char *mq_last_error(void);
/* developer aid: device buffer of 8 int64 that mq_fused_train fills with per-phase cycles */
int mq_debug_set_phase_buffer(void *dev_buf);
/* ---- rollout side: many environments stepped in lock step ------------------------------------
 * NormalizedObservations of mannequin/gym.py:61-84 for the E observations of one time step as ONE batch of
 * RunningNormalize (mannequin/_normalize.py:34-81; E = 1 is the reference's per-step behaviour): out = clip(normalize(raw),
 * +-clip).  state = double[2 d + 3]: mean[d], var[d], total weight of the mean, total weight of the variance, samples seen
 * (all zero initially); horizon <= 0: no forgetting (the reference's choice for observations). */
int mq_obsnorm_step(const float *raw, int64_t rows, int32_t d, double horizon, double *state, float clip, float *out,
                    void *stream);
/* n unit normals of the counter-based Philox4x32-10 stream (seed, stream_a, stream_b + *counter): deviate j = element
 * j % 4 of the Box-Muller pairs of block j / 4.  counter (nullable) is a device word: a captured CUDA graph that bumps it
 * draws fresh noise at every replay.  The rollout kernels draw from the same generator. */
int mq_randn_philox(uint64_t seed, uint32_t stream_a, uint32_t stream_b, const int32_t *counter, int64_t n, float *out,
                    void *stream);
/* Synthetic on-device environment (the reference's environments are gym's: out of scope): hidden state x in R^d,
 * observation = obs_scale * x + obs_shift, x' = x + dt (-damp x + B a) + sigma noise, reward = -|x'|^2 / d - ctrl |a|^2 / A,
 * episode end after `horizon` steps or with probability p_end per step, restart from x ~ N(0, 1). */
typedef struct mq_env {
    int32_t d, a;
    int32_t horizon;
    float dt, damp, sigma, ctrl, p_end;
    const float *b;             /* [d, a] row-major */
    const float *obs_scale;     /* [d] */
    const float *obs_shift;     /* [d] */
} mq_env_t;
int mq_env_reset(const mq_env_t *env, int64_t n_env, uint64_t seed, float *x, int32_t *ep_step, float *raw, void *stream);
/* one_step of mannequin/gym.py:115-122 for n_env environments at time t of a rollout of t_len steps: action = mean +
 * exp(logstd) * noise (mannequin/distrib.py:50-59; noise stream (seed, step_id, env)), records obs_n / action / reward /
 * done at [env * t_len + t] of the environment-major buffers, advances the environment, restarts finished episodes and
 * writes the next raw observations.  The step's index in the noise stream is step_id + *step_base (step_base may be
 * NULL): a device-resident base lets a captured CUDA graph draw fresh noise at every replay. */
int mq_actor_env_step(const mq_env_t *env, int64_t n_env, int32_t t, int32_t t_len, uint64_t seed, uint32_t step_id,
                      const int32_t *step_base, const float *obs_n, const float *mean, const float *logstd, float *x,
                      int32_t *ep_step,
                      float *obs_buf, float *act_buf, float *rew_buf, uint8_t *done_buf, float *raw_next, void *stream);
/* chunk boundary of benchmarks/gae.py:43-64 for n_seg segments of seg_len transitions laid end to end: at a segment's
 * last step, unless the episode really ended there, rew += gam * v_next[seg] and done = 1 (adv = r + gam v(next) - v, the
 * reference's value at a chunk end where the advantage beyond the chunk is 0). */
int mq_bootstrap_segments(float *rew, uint8_t *done, const float *v_next, int64_t n_seg, int64_t seg_len, float gam,
                          void *stream);

/* peak microbenchmarks for bench.py's roofline denominators: *flops_out = FLOPs the launch executes;
 * the caller times it with CUDA events.  mq_bench_tensor: back-to-back tcgen05.mma M128 N256 from resident operands,
 * tf32 = 0: kind::f16 (bf16 inputs), 1: kind::tf32. */
int mq_bench_ffma(int32_t iters, float *scratch, double *flops_out, void *stream);
int mq_bench_tensor(int32_t tf32, int32_t n_mma, double *flops_out, void *stream);
int mq_version(void);
int mq_device_props(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor, int64_t *smem_optin);

/* ---- optimiser: mannequin/_adam.py:7-64 + mannequin/_normalize.py:9-25 ------
 * value += lr * m / (eps + sqrt(v))     (gradient ASCENT; adams!=0: / (eps + v))
 * m += c1*(g-m), v += c2*(g*g-v) with c = weight/total_weight kept by the host.
 * state_dtype MQ_F64 follows the reference operation order bit for bit;
 * theta_out (float32 model copy, basicnet.py:55-66,161-165) may be NULL. */
int mq_adam_step(void *value, void *m, void *v, float *theta_out, const void *grad,
                 int64_t n, int state_dtype, int grad_dtype, double c1, double c2,
                 double lr, double eps, int adams, void *stream);
/* Same step with {c1, c2, lr, eps} read from 4 doubles in DEVICE memory, so that a sequence of
 * steps can be captured once in a CUDA graph and replayed with fresh coefficients. */
int mq_adam_step_dev(void *value, void *m, void *v, float *theta_out, const void *grad, int64_t n,
                     int state_dtype, int grad_dtype, const double *coefs_dev, int adams, void *stream);
/* mean += coef*(x-mean); old_out (nullable) receives the previous mean. */
int mq_running_mean_update(double *mean, const void *x, int x_dtype, int64_t n,
                           double coef, double *old_out, void *stream);

/* ---- running normaliser: mannequin/_normalize.py:34-81 ---------------------- */
int64_t mq_norm_ws_bytes(int64_t rows, int64_t cols);
/* out[c] = scale * mean_r x[r,c]   (scale = 1/world turns a later sum over ranks into the global mean) */
int mq_col_mean(const void *x, int x_dtype, int64_t rows, int64_t cols, double scale, double *out,
                void *ws, int64_t ws_bytes, void *stream);
/* out[c] = scale * mean_r (x[r,c]-mean_old[c])*(x[r,c]-mean_new[c]) */
int mq_norm_cov(const void *x, int x_dtype, int64_t rows, int64_t cols, const double *mean_old,
                const double *mean_new, double scale, double *out, void *ws, int64_t ws_bytes,
                void *stream);
/* out = (x-mean)/max(1e-8,sqrt(var)); fuse_tanh applies benchmarks/ppo.py:25 */
/* ONE statistics pass for RunningNormalize.update (_normalize.py:40-57): out[0..cols) = scale/rows * sum_r (x - shift),
 * out[cols..2cols) = scale/rows * sum_r (x - shift)^2, fp64, fixed order.  With shift = the running mean before the
 * update, mean(x) = shift + s1 and mean((x - mean_old)(x - mean_new)) = s2 - (mean_new - mean_old) s1. */
int64_t mq_norm_stats_ws_bytes(int64_t rows, int64_t cols);
int mq_norm_stats(const void *x, int x_dtype, int64_t rows, int64_t cols, const double *shift, double scale, double *out,
                  void *ws, int64_t ws_bytes, void *stream);
/* stats (taken with shift = mean) -> running state: mean += coef_mean * s1; when coef_var > 0:
 * var += coef_var * (var_scale * (s2 - (mean_new - mean_old) * s1) - var)   (_normalize.py:9-25,44-57) */
int mq_norm_update_from_stats(double *mean, double *var, const double *stats, int64_t cols, double coef_mean,
                              double coef_var, double var_scale, void *stream);
int mq_norm_apply(const void *x, int x_dtype, int64_t rows, int64_t cols, const double *mean,
                  const double *var, void *out, int out_dtype, int fuse_tanh, void *stream);

/* ---- trajectories: mannequin/_trajectory.py:79-86, mannequin/_utils.py:9-20,
 *      benchmarks/gae.py:43-64 ---------------------------------------------- */
int mq_gather_rows(const void *src, const int32_t *idx, int64_t n_idx, int64_t row_bytes,
                   int64_t n_src_rows, void *dst, void *stream);
int64_t mq_scan_ws_bytes(int64_t n);
/* adv[i] = r[i]-v[i] (+ gam*v[i+1] + gam*lam*adv[i+1] unless done[i]); ret = adv+v[:n] */
int mq_gae(const float *rew, const float *value, const uint8_t *done, int64_t n, double gam,
           double lam, float *adv, float *ret, void *ws, int64_t ws_bytes, void *stream);
/* out[i] = out[i+1] + (in[i]-out[i+1])/horizon, fp64 like the reference */
int mq_discounted(const double *in, int64_t n, double horizon, double *out, void *ws,
                  int64_t ws_bytes, void *stream);

/* ---- layered MLP path (any width; activations saved in HBM) ----------------
 * mannequin/basicnet.py:192-259 + mannequin/backprop.py:11-60 for a chain of
 * Affine / LReLU / Tanh.  acts holds every layer's output back to back, layer l
 * as [batch,out[l]] row-major. */
int64_t mq_mlp_acts_floats(const mq_net_t *net, int64_t batch);
int64_t mq_mlp_ws_bytes(const mq_net_t *net, int64_t batch);
int mq_mlp_forward(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                   int64_t batch, float *acts, void *stream);
/* The same; with a workspace (any size; mq_mlp_ws_bytes is plenty) thin layers with a long contraction are split along K */
int mq_mlp_forward_ws(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx, int64_t batch,
                      float *acts, void *ws, int64_t ws_bytes, void *stream);
/* grad[n_params] = scale * sum over batch (scale = 1/batch is the reference's
 * batch MEAN, basicnet.py:242-249); dy is clobbered; dx (nullable) = grad wrt x. */
int mq_mlp_backward(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                    const float *acts, float *dy, int64_t batch, float scale, float *grad, float *dx,
                    void *ws, int64_t ws_bytes, void *stream);
/* The same with the parameter-gradient products on `side_stream` (NULL: plain mq_mlp_backward), overlapping the chain
 * g_l -> g_{l-1} on `stream`: events = n_layers + 1 cudaEvent_t handles owned by the caller (events[l] orders dW_l after
 * g_l, events[n_layers] joins side_stream back into stream).  Capturable in a CUDA graph (fork / join edges). */
int mq_mlp_backward_overlap(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                            const float *acts, float *dy, int64_t batch, float scale, float *grad, float *dx,
                            void *ws, int64_t ws_bytes, void *stream, void *side_stream, void *const *events);

/* mq_mlp_forward_ws / mq_mlp_backward_overlap with per-call MQ_TC_* flags for the layers that run on the tcgen05 GEMM
 * (tc = 0: three bf16 planes, any operand range; 2 | MQ_TC_FP16: fp16x2, the same fp32 bar for |operand| < 65504). */
int mq_mlp_forward_tc(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx, int64_t batch,
                      float *acts, int32_t tc, void *ws, int64_t ws_bytes, void *stream);
int mq_mlp_backward_tc(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                       const float *acts, float *dy, int64_t batch, float scale, float *grad, float *dx, int32_t tc,
                       void *ws, int64_t ws_bytes, void *stream, void *side_stream, void *const *events);

/* mnist/mlp.py:24-43 at the reference's batch size as ONE persistent launch for a block of optimiser steps
 * (csrc/mq_mlp_fused.cu): 128 CTAs = 16 clusters of 8; the first layer stays split over the CTAs' shared memory with its
 * Adam moments, the 8 K-slice partials of a column slice meet through distributed shared memory, the middle of the net is
 * row-local, two grid-wide barriers per step.  Replaces per step: the gather (mlp.py:31-32), Function.evaluate
 * (basicnet.py:192-207), the softmax / L2 head (mlp.py:10-14,33), backprop with the batch mean (basicnet.py:209-257) and
 * Adam.apply_gradient (_adam.py:23-53).  idx_table: device int32 [steps][batch]; coefs_dev: device double [steps][4] =
 * {c1, c2, lr, eps}; y_onehot: float [rows][n_out]; value / m / v: optimiser state (MQ_F32 or MQ_F64), theta: fp32 copy.
 * MQ_ERR_UNSUPPORTED unless batch == 128, two or three layers, hidden width 128, n_in % 4 == 0, n_in <= 800,
 * n_out <= 16, linear output layer (mq_mlp_train_block_ws_bytes returns 0 for such nets). */
int64_t mq_mlp_train_block_ws_bytes(const mq_net_t *net, int64_t batch);
int mq_mlp_train_block(const mq_net_t *net, float *theta, void *value, void *m, void *v, int state_dtype,
                       const float *x, const float *y_onehot, const int32_t *idx_table, int64_t batch, int32_t steps,
                       const double *coefs_dev, float l2, void *ws, int64_t ws_bytes, void *stream);

/* ---- fused small-MLP path (weights resident in shared memory) -------------- */
/* rows per tile the fused kernels would use for this net (0 = does not fit). */
int mq_fused_tile_rows(const mq_net_t *net, int with_grad);
int64_t mq_fused_ws_bytes(const mq_net_t *net, int64_t batch);
/* out[batch,n_out] and/or logp[batch] may be NULL.  logp needs `sample`: sampled actions [batch,A] when the net
 * has logstd parameters (Gaussian), one-hot actions [batch,n_out] otherwise (categorical, distrib.py:24-30). */
int mq_fused_forward(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                     int64_t batch, float *out, const float *sample, float *logp, void *stream);
/* The same with per-call MQ_TC_* flags (mq_fused_forward = flags 0). */
int mq_fused_forward_tc(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                        int64_t batch, float *out, const float *sample, float *logp, int32_t tc, void *stream);
/* The same from packed records (mq_pack_records) on the record-fed tensor-core kernel, forward only: out[batch, n_out] and /
 * or logp[batch] (the sample is the record's target); tc_image = operand image of the current parameters
 * (mq_tc_image_build) in the format `tc` names.  MQ_ERR_UNSUPPORTED when the net does not fit that kernel. */
int mq_fused_forward_rec(const mq_net_t *net, const float *theta, const float *records, int32_t rec_w, const void *tc_image,
                         const int32_t *row_idx, int64_t batch, float *out, float *logp, int32_t tc, void *stream);
/* forward + head + backward in one launch; grad[n_params] = scale * sum_batch */
int mq_fused_grad(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                  int64_t batch, const mq_head_t *head, float scale, float *grad, float *out,
                  float *logp, void *ws, int64_t ws_bytes, void *stream);
/* Same, but stops before the cross-CTA sum: ws receives n_parts UNSCALED partial gradients, [n_parts][n_params]
 * (n_parts = CTAs of the gradient kernel), to be consumed by mq_reduce_step. */
int mq_fused_grad_partials(const mq_net_t *net, const float *theta, const float *x, const int32_t *row_idx,
                           int64_t batch, const mq_head_t *head, float *out, float *logp, void *ws,
                           int64_t ws_bytes, int32_t *n_parts_out, void *stream);

/* ---- packed minibatch records and operand images of the tensor-core gradient kernel (csrc/mq_tc3.cu) --------------
 * Records replace the four separate gathers of benchmarks/ppo.py:30-37 (traj.o[idx], traj.a[idx], traj.r[idx],
 * baseline[idx]; mannequin/_trajectory.py:79-86) by one aligned row per transition, packed once per update. */
int32_t mq_record_width(const mq_net_t *net);           /* floats per record: n_in + n_out + 2 rounded up to 16 */
/* rec[r] = x[r, :n_in] | tgt[r, :n_out] | w0[r] | w1[r] | 0...; tgt / w0 / w1 may be NULL (zeros). */
int mq_pack_records(const mq_net_t *net, const float *x, const float *tgt, const float *w0, const float *w1,
                    int64_t rows, float *rec, void *stream);
/* bytes of the operand image of `net` for the planes in `tc` (0 when the net does not fit the kernel) */
int64_t mq_tc_image_bytes(const mq_net_t *net, int32_t tc);
/* image <- bf16 planes of theta in the kernel's shared-memory layout (weights of basicnet.py:288-311 as K-major /
 * MN-major core matrices, biases, logstd, exp(-logstd)); the gradient kernel copies it with ONE bulk copy. */
int mq_tc_image_build(const mq_net_t *net, const float *theta, int32_t tc, void *image, void *stream);

/* Peer-memory descriptor of the in-kernel all-reduce (one exchange of the flat gradient per
 * optimiser step).  grad_bufs / flag_bufs are DEVICE arrays of `world` pointers: rank r's exchange buffer
 * (float[2][world][buf_stride]) and flag pad (uint32[2][world][mq_step_blocks(n)], zero-initialised), each mapped into
 * this process (symmetric memory).  step must be non-zero and increase by one per call on every rank. */
typedef struct mq_peer {
    int32_t rank, world;
    float *const *grad_bufs;
    uint32_t *const *flag_bufs;
    int64_t buf_stride;
    uint32_t step;
    int32_t pull;                /* 2: low-latency push: (value, step) as ONE 8-byte word per parameter into every peer's slot [rank],
                                    the receiver polls the words in its own memory (no flags, no fences); grad_bufs are then
                                    uint64[2][world][buf_stride];
                                    0: push (remote STORES into every peer's slot [rank], local reads; fastest up to 4 GPUs);
                                    1: pull (local store of the own slot, remote LOADS of every peer's: one release fence
                                    towards local memory instead of one per peer; faster on 8 GPUs) */
} mq_peer_t;
int64_t mq_step_blocks(int64_t n);
/* One launch: g = scale * sum_k partials[k] (fixed order) [-> sum over ranks through peer memory, rank order]
 * -> Adam / Adams step (mannequin/_adam.py:23-53) -> theta_out (float32 model copy, nullable), grad_out (nullable).
 * peer == NULL or world 1: single GPU. */
int mq_reduce_step(const float *partials, int32_t n_parts, int64_t n, float scale, void *value, void *m, void *v,
                   float *theta_out, float *grad_out, int state_dtype, double c1, double c2, double lr,
                   double eps, int adams, const mq_peer_t *peer, void *stream);
/* The same, and the blocks that own a parameter also refresh its entries in the operand image (`image` built for
 * `net` and the planes in `tc`), so that the next gradient launch needs no weight conversion at all. */
int mq_reduce_step_image(const float *partials, int32_t n_parts, int64_t n, float scale, void *value, void *m, void *v,
                         float *theta_out, float *grad_out, int state_dtype, double c1, double c2, double lr,
                         double eps, int adams, const mq_peer_t *peer, const mq_net_t *net, int32_t tc, void *image,
                         void *stream);
/* One parameter vector's optimiser step for mq_reduce_step2 (the arguments of mq_reduce_step_image as a struct). */
typedef struct mq_step {
    const float *partials; int32_t n_parts; int64_t n; float scale;
    void *value, *m, *v; float *theta_out, *grad_out;
    double c1, c2, lr, eps;
    const mq_peer_t *peer;      /* nullable */
    const mq_net_t *net; int32_t tc; void *image;      /* operand image to refresh (image nullable) */
} mq_step_t;
/* mq_reduce_step_image for TWO parameter vectors in ONE launch (b nullable): the value and the policy net of one PPO
 * update step -- sharded, the pair pays one exchange latency instead of two. */
int mq_reduce_step2(const mq_step_t *a, const mq_step_t *b, int state_dtype, int adams, void *stream);
/* Two independent minibatch gradients (benchmarks/gae.py:71-73 and benchmarks/ppo.py:29-40 of the same update) as ONE
 * launch of the record-fed tensor-core kernel, each net on half of the SMs; both heads must carry packed records and the
 * nets must agree in planes, padded input width and hidden activation, else MQ_ERR_UNSUPPORTED.  ws_x receives net x's
 * per-CTA partials ([*n_parts_x][n_params]). */
int mq_fused_grad_partials2(const mq_net_t *net_a, const float *theta_a, const int32_t *idx_a, const mq_head_t *head_a,
                            void *ws_a, int64_t ws_a_bytes, int32_t *n_parts_a, const mq_net_t *net_b, const float *theta_b,
                            const int32_t *idx_b, const mq_head_t *head_b, void *ws_b, int64_t ws_b_bytes, int32_t *n_parts_b,
                            int64_t batch, void *stream);
/* One net's side of mq_train_steps2 (the arguments of mq_train_steps as a struct). */
typedef struct mq_train {
    const mq_net_t *net; float *theta; const float *x; const mq_head_t *head;
    const int32_t *idx_table; int64_t idx_stride; float scale;
    void *value, *m, *v; const double *coefs_host; double lr, eps;
    const mq_peer_t *peer;      /* nullable: descriptor of the FIRST step */
    void *ws; int64_t ws_bytes;
} mq_train_t;
/* mq_train_steps for two nets on the same minibatch schedule: per step one paired gradient launch and one paired
 * reduce / exchange / Adam launch. */
int mq_train_steps2(const mq_train_t *a, const mq_train_t *b, int32_t n_steps, int64_t mb, int state_dtype, int adams,
                    void *stream);
/* n_steps dependent large-minibatch steps {mq_fused_grad_partials, mq_reduce_step_image} issued by one call
 * (benchmarks/ppo.py:29-40, benchmarks/gae.py:71-73 at 65,536-row minibatches): step s reads mb indices at idx_table +
 * s * idx_stride (idx_stride > mb: this rank's columns of the global minibatch table), coefs_host =
 * [n_steps][2] host doubles (c1, c2), peer (nullable) = descriptor of the FIRST step (its tag counts up per step),
 * head->tc_image (nullable) is refreshed by every step. */
int mq_train_steps(const mq_net_t *net, float *theta, const float *x, const mq_head_t *head,
                   const int32_t *idx_table, int64_t idx_stride, int32_t n_steps, int64_t mb, float scale, void *value,
                   void *m, void *v,
                   int state_dtype, const double *coefs_host, double lr, double eps, int adams, const mq_peer_t *peer,
                   void *ws, int64_t ws_bytes, void *stream);
/* n_steps dependent {gather, forward, head, backward, Adam} steps in ONE launch
 * (benchmarks/ppo.py:29-40, benchmarks/gae.py:71-73).  idx_table[n_steps,mb].
 * Adam state (value,m,v: state_dtype) and theta are updated in place; tw[2]
 * holds the two running total weights (host doubles, updated on return). */
int mq_fused_train(const mq_net_t *net, float *theta, void *value, void *m, void *v, int state_dtype,
                   const float *x, const mq_head_t *head, const int32_t *idx_table, int32_t n_steps,
                   int32_t mb, double beta1, double beta2, double *tw_host, double lr, double eps,
                   void *stream);

/* ---- evolution strategies: benchmarks/mutation.py:33-37 -------------------- */
/* returns[p] = sum_t sum_a policy(theta+noise[p])(obs[t])[a]*ret_coef[a] */
int mq_es_eval(const mq_net_t *net, const float *theta, const float *noise, int64_t n_members,
               const float *obs, int64_t n_obs, const float *ret_coef, float *returns, void *stream);
/* out[j] = scale * sum_p noise[p,j]*w[p]   (fixed summation order) */
int mq_es_weighted_sum(const float *noise, const double *w, int64_t n_members, int64_t n_params,
                       double scale, double *out, void *ws, int64_t ws_bytes, void *stream);
int64_t mq_es_ws_bytes(int64_t n_members, int64_t n_params);

/* ---- per-op kernels for the general graph (mannequin/backprop.py:11-60,
 *      mannequin/distrib.py:50-68, mnist/vae.py:13-34) ----------------------- */
/* C[M,N] = alpha * opA(A)[M,K] @ opB(B)[K,N]; element strides in floats */
int mq_gemm(const float *a, int64_t a_rs, int64_t a_cs, const float *b, int64_t b_rs, int64_t b_cs,
            float *c, int64_t m, int64_t n, int64_t k, float alpha, void *stream);
/* The same; with a workspace, products with few output tiles and a long contraction are split along K (deterministic) */
int mq_gemm_ws(const float *a, int64_t a_rs, int64_t a_cs, const float *b, int64_t b_rs, int64_t b_cs,
               float *c, int64_t m, int64_t n, int64_t k, float alpha, void *ws, int64_t ws_bytes, void *stream);
/* workspace with which mq_gemm_ws / mq_gemm_ex take the TMA-fed tcgen05 path for this shape (packed bf16 planes of both
 * operands + split-K partials) */
int64_t mq_gemm_ws_bytes(int64_t m, int64_t n, int64_t k);
/* The same with per-call MQ_TC_* flags: planes (precision mode), MQ_TC_FORCE (tensor-core kernel whatever the size),
 * MQ_TC_OFF (FFMA kernel). */
int mq_gemm_ex(const float *a, int64_t a_rs, int64_t a_cs, const float *b, int64_t b_rs, int64_t b_cs, float *c,
               int64_t m, int64_t n, int64_t k, float alpha, int32_t tc, void *ws, int64_t ws_bytes, void *stream);
int mq_add_bias(const float *x, const float *b, float *y, int64_t rows, int64_t cols, void *stream);
int mq_binary(const float *a, const float *b, float *out, int64_t n, int64_t b_period, int op,
              void *stream);                   /* op 0: a+b, 1: a*b; b is broadcast with period */
/* out[c] = scale * sum_r g[r,c] * (a ? a[r,c] : 1) */
int mq_col_sum(const float *g, const float *a, int64_t rows, int64_t cols, float scale, float *out,
               void *stream);
int mq_act_forward(const float *x, float *y, int64_t n, int act, float leak, void *stream);
int mq_act_backward(const float *y, const float *g, float *out, int64_t n, int act, float leak,
                    void *stream);
int mq_clip_forward(const float *x, float *y, int64_t n, float lo, float hi, void *stream);
int mq_clip_backward(const float *x, const float *g, float *out, int64_t n, float lo, float hi,
                     void *stream);
/* logstd_period = A (shared Params(A)) or rows*A (per-row sub-network) */
int mq_gauss_logprob(const float *mu, const float *logstd, int64_t logstd_period, const float *sample,
                     int64_t rows, int64_t a, float *logp, void *stream);
int mq_gauss_logprob_bwd(const float *mu, const float *logstd, int64_t logstd_period,
                         const float *sample, const float *g, int64_t rows, int64_t a, float *d_mu,
                         float *d_logstd_rows, void *stream);
/* Gaussian entropy, SURVEY a24 (named by the north star; absent from the reference, so pinned analytically and by
 * finite differences only): h[r] = sum_a logstd[r,a] + A/2 (1 + ln 2 pi) when h != NULL; d_logstd[r,a] = g[r] when
 * d_logstd != NULL.  logstd is [rows, A] (rows = 1 for a shared Params(A)). */
int mq_gauss_entropy(const float *logstd, const float *g, int64_t rows, int64_t a, float *h, float *d_logstd,
                     void *stream);
/* mnist/vae.py:64-69: g = g1 (+ g2 on the first n2 entries); mom = clip(decay*mom + (1-decay)*g, +-clip); theta += lr*mom */
int mq_momentum_step(float *theta, float *mom, const float *g1, const float *g2, int64_t n2, int64_t n, float decay,
                     float clip, float lr, void *stream);
/* ---- fused training step of mnist/vae.py (csrc/mq_vae.cu) -------------------------------------------------------------
 * Replaces, for the net of /root/reference/mnist/vae.py:39-55 (n_hidden tanh layers of `hidden` units per side, `latent`
 * Gaussian heads, clipped log-deviations), the two graph evaluations and the update of vae.py:57-71:
 *     logprob, bp = decoder.logprob.evaluate(x, sample=x); grad1 = bp(ones)          (vae.py:60-61)
 *     dkl, bp = dkl.evaluate(x);                           grad2 = bp(-ones)         (vae.py:63-64)
 *     grad1[:len(grad2)] += grad2; momentum = clip(0.9 momentum + 0.1 grad1, -1, 1); theta += 0.001 momentum
 * as one forward and one backward pass (the two upstream gradients meet at the encoder heads; every product on the TMA-fed
 * tcgen05 GEMM with its operands kept as packed planes between products).  theta is the flat parameter vector of
 * decoder.logprob (basicnet.py:66-77 order: encoder hidden (W, b) ..., W_mean, b_mean, W_logstd, b_logstd, decoder hidden
 * (W, b) ..., W_mean, b_mean, W_logstd, b_logstd); x is [batch, d_img] row-major; unit is the N(0,1) draw of
 * distrib.py:49 for the latent sample, [batch, latent].  hidden % 8 == 0, 1 <= n_hidden <= 4, else MQ_ERR_UNSUPPORTED. */
typedef struct mq_vae_t {
    int64_t batch, d_img, hidden, latent;
    int32_t n_hidden;          /* tanh layers per side (vae.py: 2; BASELINE configs[4]: 1)                         */
    int32_t tc;                /* operand format, MQ_TC_* planes (| MQ_TC_FP16); 0 = fp16x2 (fp32 products)        */
    float   clip_lo, clip_hi;  /* Clip of both logstd heads (vae.py:43,54: -6, 0)                                  */
} mq_vae_t;
int64_t mq_vae_n_params(const mq_vae_t *v);                                        /* < 0: bad descriptor */
int64_t mq_vae_ws_bytes(const mq_vae_t *v);
/* once per workspace, before the first step (zeroes the padding of every plane image, writes the bias-gradient rows) */
int mq_vae_ws_init(const mq_vae_t *v, void *ws, int64_t ws_bytes, void *stream);
/* grad[n_params] = grad1 with grad2 added onto the encoder slice; logprob[batch], dkl[batch] the two values */
int mq_vae_grad(const mq_vae_t *v, const float *theta, const float *x, const float *unit, float *grad, float *logprob,
                float *dkl, void *ws, int64_t ws_bytes, void *stream);
/* the whole iteration: mq_vae_grad, then mq_momentum_step(theta, mom, grad, ...) */
int mq_vae_step(const mq_vae_t *v, float *theta, float *mom, const float *x, const float *unit, float *grad, float *logprob,
                float *dkl, float decay, float clip, float lr, void *ws, int64_t ws_bytes, void *stream);

/* categorical log-prob (mannequin/distrib.py:24-32): logp[r] = log softmax(logits[r])[sample[r]] and / or
 * d_logits[r] = g[r] * (onehot(sample[r]) - softmax(logits[r])); g is only read when d_logits != NULL */
int mq_discrete_logprob(const float *logits, const int32_t *sample, const float *g, int64_t rows, int64_t n,
                        float *logp, float *d_logits, void *stream);
/* out[rows,n] = one-hot rows of idx: the `target` of the categorical heads */
int mq_onehot(const int32_t *idx, int64_t rows, int64_t n, float *out, void *stream);
int mq_gauss_sample(const float *mu, const float *logstd, int64_t logstd_period, const float *unit,
                    int64_t rows, int64_t a, float *out, float *noise, void *stream);
int mq_dkl_forward(const float *mean, const float *logstd, int64_t rows, int64_t d, float *out,
                   void *stream);
int mq_dkl_backward(const float *mean, const float *logstd, const float *g, int64_t rows, int64_t d,
                    float *d_mean, float *d_logstd, void *stream);
int mq_softmax_ce_grad(const float *logits, const float *onehot, const int32_t *row_idx,
                       int64_t rows, int64_t cols, float l2, float *dy, void *stream);
/* ppo.py:33-37: w = exp(logp-logp_old[idx]) with the clip rule, times adv[idx] */
int mq_ppo_weight(const float *logp, const float *logp_old, const float *adv, const int32_t *row_idx,
                  int64_t rows, float lo, float hi, float *w, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* MQ_B200_H */
