// core.cu -- version, error string, device info.
#include <cstdarg>

#include "common.cuh"

namespace b200mp {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace b200mp

extern "C" const char* b200mp_version(void) { return B200MP_VERSION; }
extern "C" const char* b200mp_last_error(void) { return b200mp::g_err; }

extern "C" int b200mp_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes) {
    int dev = 0;
    B200MP_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) {
        B200MP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        *sm_count = v;
    }
    if (cc_major) {
        B200MP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
        *cc_major = v;
    }
    if (cc_minor) {
        B200MP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
        *cc_minor = v;
    }
    if (l2_bytes) {
        B200MP_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev));
        *l2_bytes = v;
    }
    return B200MP_OK;
}
