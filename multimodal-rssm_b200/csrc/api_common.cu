#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void mrssm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* mrssm_last_error(void) { return g_err; }
extern "C" int mrssm_abi_version(void) { return MRSSM_ABI_VERSION; }

extern "C" int mrssm_device_ok(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        mrssm_set_error("no CUDA device");
        return 0;
    }
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        mrssm_set_error("device compute capability %d.x, library is sm_100a only", major);
        return 0;
    }
    return 1;
}
