// Error plumbing and small queries of libsedk.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace sedk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
        return SEDK_ERR_CUDA;
    }
    return SEDK_OK;
}

}  // namespace sedk

extern "C" const char* sedk_last_error(void) { return sedk::g_err; }
extern "C" int sedk_version(void) { return 100; }
extern "C" int sedk_device_cc(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
    return major * 10 + minor;
}
extern "C" int sedk_sizeof_crnn_plan(void) { return (int)sizeof(sedk_crnn_plan); }
