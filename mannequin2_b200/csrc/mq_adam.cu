// Fused two-moment optimiser step over the flat parameter vector.
//
// Replaces mannequin/_adam.py:23-53 (BaseTwoMoments.apply_gradient, Adam, Adams)
// and the two RunningMean.update calls inside it (mannequin/_normalize.py:9-25):
// five NumPy passes with fp64 temporaries become ONE pass that reads
// {value, m, v, grad} and writes {value, m, v, theta_f32}.
//
// HBM-bound: 24 B/param (fp32 state, +4 B grad, +4 B fp32 model copy) or
// 48 B/param (fp64 state).  128-bit accesses, 4 elements per thread per trip.
//
// The fp64 instantiation spells out every rounding with __d*_rn so that it is
// bit-identical to NumPy's sequence of separate ufunc calls (no FMA contraction).
#include "mq_common.cuh"

template <typename T> struct Vec4;
template <> struct Vec4<float> {
    float v[4];
    __device__ __forceinline__ void load(const float *p) {
        float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float *p) const {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <> struct Vec4<double> {
    double v[4];
    __device__ __forceinline__ void load(const double *p) {
        double2 a = *reinterpret_cast<const double2 *>(p);
        double2 b = *reinterpret_cast<const double2 *>(p + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    __device__ __forceinline__ void store(double *p) const {
        *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2 *>(p + 2) = make_double2(v[2], v[3]);
    }
};

template <bool ADAMS>
__device__ __forceinline__ void adam_elem(double &val, double &m, double &v, double g, double c1,
                                          double c2, double lr, double eps) {
    m = __dadd_rn(m, __dmul_rn(c1, __dsub_rn(g, m)));
    v = __dadd_rn(v, __dmul_rn(c2, __dsub_rn(__dmul_rn(g, g), v)));
    const double den = __dadd_rn(eps, ADAMS ? v : __dsqrt_rn(v));
    val = __dadd_rn(val, __dmul_rn(lr, __ddiv_rn(m, den)));
}

template <bool ADAMS>
__device__ __forceinline__ void adam_elem(float &val, float &m, float &v, float g, float c1,
                                          float c2, float lr, float eps) {
    m += c1 * (g - m);
    v += c2 * (g * g - v);
    const float den = eps + (ADAMS ? v : sqrtf(v));
    val += lr * (m / den);
}

template <typename ST, typename GT, bool ADAMS, bool VEC>
__global__ void __launch_bounds__(MQ_NT)
adam_kernel(ST *__restrict__ value, ST *__restrict__ m, ST *__restrict__ v,
            float *__restrict__ theta_out, const GT *__restrict__ grad, int64_t n, ST c1, ST c2,
            ST lr, ST eps, const double *__restrict__ coefs_dev) {
    if (coefs_dev) {                 // {c1, c2, lr, eps} from device memory (CUDA-graph replay)
        c1 = (ST)coefs_dev[0]; c2 = (ST)coefs_dev[1]; lr = (ST)coefs_dev[2]; eps = (ST)coefs_dev[3];
    }
    const int64_t tid = (int64_t)blockIdx.x * MQ_NT + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * MQ_NT;
    if (VEC) {
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nthreads) {
            const int64_t e = i << 2;
            Vec4<ST> a, b, c;
            Vec4<GT> g;
            a.load(value + e); b.load(m + e); c.load(v + e); g.load(grad + e);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                adam_elem<ADAMS>(a.v[j], b.v[j], c.v[j], (ST)g.v[j], c1, c2, lr, eps);
            a.store(value + e); b.store(m + e); c.store(v + e);
            if (theta_out) {
                Vec4<float> t;
#pragma unroll
                for (int j = 0; j < 4; ++j) t.v[j] = (float)a.v[j];
                t.store(theta_out + e);
            }
        }
        for (int64_t e = (n4 << 2) + tid; e < n; e += nthreads) {
            ST a = value[e], b = m[e], c = v[e];
            adam_elem<ADAMS>(a, b, c, (ST)grad[e], c1, c2, lr, eps);
            value[e] = a; m[e] = b; v[e] = c;
            if (theta_out) theta_out[e] = (float)a;
        }
    } else {
        for (int64_t e = tid; e < n; e += nthreads) {
            ST a = value[e], b = m[e], c = v[e];
            adam_elem<ADAMS>(a, b, c, (ST)grad[e], c1, c2, lr, eps);
            value[e] = a; m[e] = b; v[e] = c;
            if (theta_out) theta_out[e] = (float)a;
        }
    }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <typename ST, typename GT>
static int adam_launch(void *value, void *m, void *v, float *theta_out, const void *grad, int64_t n,
                       double c1, double c2, double lr, double eps, int adams, void *stream,
                       const double *coefs_dev = nullptr) {
    const bool vec = aligned16(value) && aligned16(m) && aligned16(v) && aligned16(grad) &&
                     (theta_out == nullptr || aligned16(theta_out)) && n >= 4;
    const int64_t work = vec ? (n >> 2) : n;
    int64_t blocks = mq_ceil_div(work, MQ_NT);
    const int64_t cap = (int64_t)mq_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
#define MQ_ADAM_GO(A, V)                                                                       \
    do {                                                                                       \
        auto kfn = adam_kernel<ST, GT, A, V>;                                                  \
        MQ_LAUNCH(kfn, (unsigned)blocks, MQ_NT, 0, stream, (ST *)value, (ST *)m, (ST *)v,      \
                  theta_out, (const GT *)grad, n, (ST)c1, (ST)c2, (ST)lr, (ST)eps, coefs_dev); \
    } while (0)
    if (adams) { if (vec) MQ_ADAM_GO(true, true); else MQ_ADAM_GO(true, false); }
    else       { if (vec) MQ_ADAM_GO(false, true); else MQ_ADAM_GO(false, false); }
#undef MQ_ADAM_GO
    MQ_CHECK_LAUNCH("mq_adam_step");
    return MQ_OK;
}

extern "C" int mq_adam_step(void *value, void *m, void *v, float *theta_out, const void *grad,
                            int64_t n, int state_dtype, int grad_dtype, double c1, double c2,
                            double lr, double eps, int adams, void *stream) {
    MQ_REQUIRE(value && m && v && grad, MQ_ERR_INVALID, "mq_adam_step: null pointer");
    MQ_REQUIRE(n >= 0, MQ_ERR_SHAPE, "mq_adam_step: negative size");
    if (n == 0) return MQ_OK;
    if (state_dtype == MQ_F64 && grad_dtype == MQ_F32)
        return adam_launch<double, float>(value, m, v, theta_out, grad, n, c1, c2, lr, eps, adams, stream);
    if (state_dtype == MQ_F64 && grad_dtype == MQ_F64)
        return adam_launch<double, double>(value, m, v, theta_out, grad, n, c1, c2, lr, eps, adams, stream);
    if (state_dtype == MQ_F32 && grad_dtype == MQ_F32)
        return adam_launch<float, float>(value, m, v, theta_out, grad, n, c1, c2, lr, eps, adams, stream);
    mq_set_error("mq_adam_step: unsupported dtype pair (%d,%d)", state_dtype, grad_dtype);
    return MQ_ERR_INVALID;
}

extern "C" int mq_adam_step_dev(void *value, void *m, void *v, float *theta_out, const void *grad,
                                int64_t n, int state_dtype, int grad_dtype, const double *coefs_dev,
                                int adams, void *stream) {
    MQ_REQUIRE(value && m && v && grad && coefs_dev, MQ_ERR_INVALID, "mq_adam_step_dev: null pointer");
    MQ_REQUIRE(n >= 0, MQ_ERR_SHAPE, "mq_adam_step_dev: negative size");
    if (n == 0) return MQ_OK;
    if (state_dtype == MQ_F64 && grad_dtype == MQ_F32)
        return adam_launch<double, float>(value, m, v, theta_out, grad, n, 0, 0, 0, 0, adams, stream, coefs_dev);
    if (state_dtype == MQ_F32 && grad_dtype == MQ_F32)
        return adam_launch<float, float>(value, m, v, theta_out, grad, n, 0, 0, 0, 0, adams, stream, coefs_dev);
    mq_set_error("mq_adam_step_dev: unsupported dtype pair (%d,%d)", state_dtype, grad_dtype);
    return MQ_ERR_INVALID;
}

// mean += coef * (x - mean)   (mannequin/_normalize.py:9-25, the data-dependent part)
template <typename XT>
__global__ void __launch_bounds__(MQ_NT)
running_mean_kernel(double *__restrict__ mean, const XT *__restrict__ x, int64_t n, double coef,
                    double *__restrict__ old_out) {
    for (int64_t i = (int64_t)blockIdx.x * MQ_NT + threadIdx.x; i < n; i += (int64_t)gridDim.x * MQ_NT) {
        const double mu = mean[i];
        if (old_out) old_out[i] = mu;
        mean[i] = __dadd_rn(mu, __dmul_rn(coef, __dsub_rn((double)x[i], mu)));
    }
}

extern "C" int mq_running_mean_update(double *mean, const void *x, int x_dtype, int64_t n,
                                      double coef, double *old_out, void *stream) {
    MQ_REQUIRE(mean && x, MQ_ERR_INVALID, "mq_running_mean_update: null pointer");
    if (n <= 0) return MQ_OK;
    int64_t blocks = mq_ceil_div(n, MQ_NT);
    const int64_t cap = (int64_t)mq_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (x_dtype == MQ_F64) {
        auto kfn = running_mean_kernel<double>;
        MQ_LAUNCH(kfn, (unsigned)blocks, MQ_NT, 0, stream, mean, (const double *)x, n, coef, old_out);
    } else {
        auto kfn = running_mean_kernel<float>;
        MQ_LAUNCH(kfn, (unsigned)blocks, MQ_NT, 0, stream, mean, (const float *)x, n, coef, old_out);
    }
    MQ_CHECK_LAUNCH("mq_running_mean_update");
    return MQ_OK;
}
