// Error plumbing + misc entry points of the C-ABI.
#include <stdarg.h>

#include "common.cuh"

namespace dm {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    if (e == cudaErrorMemoryAllocation) {
        set_error("CUDA out of memory: %s (%s)", cudaGetErrorString(e), what);
        cudaGetLastError();
        return DM_E_OOM;
    }
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    cudaGetLastError();
    return DM_E_CUDA;
}
}  // namespace dm

extern "C" __attribute__((visibility("default"))) const char *dm_last_error(void) { return dm::g_err; }
extern "C" __attribute__((visibility("default"))) int dm_version(void) { return 100; }
extern "C" __attribute__((visibility("default"))) int dm_device_name(char *buf, int len) {
    if (!buf || len <= 0) return DM_E_INVALID;
    buf[0] = 0;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) { cudaGetLastError(); return DM_OK; }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); return DM_OK; }
    strncpy(buf, p.name, (size_t)len - 1);
    buf[len - 1] = 0;
    return DM_OK;
}
