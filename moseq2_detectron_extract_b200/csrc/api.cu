// Error plumbing, version and device queries of the C ABI.
#include "common.cuh"
#include <string.h>

namespace msq {

static thread_local char g_error[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached = v;
        cached_dev = dev;
    }
    return cached;
}

}  // namespace msq

extern "C" int msq_version(void) { return MSQ_VERSION; }

extern "C" const char *msq_last_error(void) { return msq::g_error; }

extern "C" int msq_device_info(int *sms, int *cc_major, int *cc_minor, char *name, int name_len) {
    int dev = 0;
    MSQ_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MSQ_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    if (sms) *sms = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (name && name_len > 0) {
        strncpy(name, prop.name, (size_t)name_len - 1);
        name[name_len - 1] = 0;
    }
    return MSQ_OK;
}
