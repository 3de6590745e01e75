// Library plumbing: error strings, device properties.
#include <stdarg.h>

#include "mq_common.cuh"

static thread_local char g_err[512] = "";

void mq_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#ifdef MQ_EMU
int mq_check_launch(const char *) { return MQ_OK; }
int mq_sm_count() { return 4; }
int64_t mq_smem_optin() { return 227 * 1024; }
#else
int mq_check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        mq_set_error("%s: %s", what, cudaGetErrorString(e));
        return MQ_ERR_CUDA;
    }
    return MQ_OK;
}

static int g_sm_count = 0;
static int64_t g_smem_optin = 0;

static void mq_query_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return; }
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess) g_sm_count = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) == cudaSuccess)
        g_smem_optin = v;
    cudaGetLastError();
}

int mq_sm_count() {
    if (g_sm_count == 0) mq_query_device();
    return g_sm_count > 0 ? g_sm_count : 148;
}

int64_t mq_smem_optin() {
    if (g_smem_optin == 0) mq_query_device();
    return g_smem_optin > 0 ? g_smem_optin : 227 * 1024;
}
#endif

extern "C" const char *mq_last_error(void) { return g_err; }

extern "C" int mq_version(void) { return 100; }

extern "C" int mq_device_props(int32_t *sm_count, int32_t *cc_major, int32_t *cc_minor,
                               int64_t *smem_optin) {
#ifdef MQ_EMU
    if (sm_count) *sm_count = 4;
    if (cc_major) *cc_major = 0;
    if (cc_minor) *cc_minor = 0;
    if (smem_optin) *smem_optin = mq_smem_optin();
    return MQ_OK;
#else
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        mq_set_error("mq_device_props: no CUDA device (%s)", cudaGetErrorString(e));
        return MQ_ERR_CUDA;
    }
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (sm_count) *sm_count = v;
    cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev);
    if (cc_major) *cc_major = v;
    cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev);
    if (cc_minor) *cc_minor = v;
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem_optin) *smem_optin = v;
    return mq_check_launch("mq_device_props");
#endif
}
