// Error reporting and device checks shared by all entry points.
#include <stdarg.h>
#include <string.h>

#include "rr_common.cuh"

static thread_local char g_last_error[1024] = "";

void rr_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

#include <atomic>
static std::atomic<long long> g_launches{0};
void rr_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" int64_t rr_kernel_launches(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

#include <stdlib.h>
bool rr_pdl_enabled() {
    static const bool on = getenv("RR_NO_PDL") == nullptr;
    return on;
}

extern "C" const char* rr_last_error(void) { return g_last_error; }
extern "C" int rr_abi_version(void) { return RR_ABI_VERSION; }

extern "C" int rr_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        rr_set_error("cudaGetDeviceCount failed: %s (no CPU fallback exists)", cudaGetErrorString(e));
        cudaGetLastError();
        return RR_ERR_CUDA;
    }
    return n;
}

int rr_check_device_arch(int device) {
    int n = rr_device_count();
    if (n < 0) return n;
    if (device < 0 || device >= n) {
        rr_set_error("device %d out of range (%d CUDA devices visible; no CPU fallback exists)", device, n);
        return n == 0 ? RR_ERR_CUDA : RR_ERR_INVALID;
    }
    int major = 0, minor = 0;
    cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    if (e != cudaSuccess) {
        rr_set_error("cudaDeviceGetAttribute failed: %s", cudaGetErrorString(e));
        return RR_ERR_CUDA;
    }
    if (major != 10) {
        rr_set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, major, minor);
        return RR_ERR_ARCH;
    }
    return RR_OK;
}
