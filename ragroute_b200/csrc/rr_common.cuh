// Shared host/device helpers for libragroute_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

#include "../../include/ragroute_b200.h"

// ------------------------------------------------------------------ errors --
void rr_set_error(const char* fmt, ...);

#define RR_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            rr_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),         \
                         __FILE__, __LINE__);                                            \
            return (_e == cudaErrorMemoryAllocation) ? RR_ERR_OOM : RR_ERR_CUDA;         \
        }                                                                                \
    } while (0)

#define RR_REQUIRE(cond, ...)                                                            \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            rr_set_error(__VA_ARGS__);                                                   \
            return RR_ERR_INVALID;                                                       \
        }                                                                                \
    } while (0)

struct rr_device_guard {
    int prev = -1;
    bool ok = false;
    explicit rr_device_guard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = (cudaSetDevice(dev) == cudaSuccess);
    }
    ~rr_device_guard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int rr_check_device_arch(int device);  // RR_OK iff compute capability 10.x
bool rr_pdl_enabled();                  // false when RR_NO_PDL is set (measurement knob)

#ifdef __CUDACC__
// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-stream-serialization attribute
template <class... KArgs, class... Args>
static inline cudaError_t rr_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                        Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = rr_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif
void rr_count_launch(int n);            // process-wide count of this library's kernel launches

constexpr int RR_MAX_DEVICES = 64;  // per-device one-time setup flags
constexpr float RR_NEG_INF = -__builtin_huge_valf();
constexpr uint32_t RR_NO_ROW = 0xFFFFFFFFu;
constexpr unsigned RR_FULL = 0xFFFFFFFFu;

// ---------------------------------------------------- split-plane fp32 rows --
// A corpus row is stored as two 16-bit planes:
//   hi = (bits + 0x8000) >> 16      -- bf16, round-to-nearest (ties away), MMA operand
//   lo = (bits + 0x8000) & 0xFFFF   -- remainder
// so that bits == ((hi << 16) | lo) - 0x8000 exactly: 4 bytes per element in
// total, the scan streams only the 2-byte hi plane, rescoring rebuilds fp32.
__host__ __device__ __forceinline__ void rr_split_f32(float x, uint16_t& hi, uint16_t& lo) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(x) + 0x8000u;
#else
    uint32_t b;
    memcpy(&b, &x, 4);
    b += 0x8000u;
#endif
    hi = (uint16_t)(b >> 16);
    lo = (uint16_t)(b & 0xFFFFu);
}

#ifdef __CUDACC__
// ---- programmatic dependent launch (PDL) for the whole step -------------------------------------------
// Every kernel of a step is launched with cudaLaunchAttributeProgrammaticStreamSerialization and begins with
// rr_pdl_enter(): "launch_dependents" lets the NEXT kernel of the stream be scheduled onto the SMs while this
// one still runs (its blocks then sit at their own wait), "wait" blocks until the PREVIOUS kernel has completed
// and its writes are visible.  Because every kernel executes the wait before it touches global memory and
// before it can exit, completion order stays stream order (transitively), so the step computes exactly what
// plain stream serialisation computes -- only the launch / drain gaps between its ~12 small kernels overlap.
// (The wait is a no-op when the kernel was launched without the attribute.)
__device__ __forceinline__ void rr_pdl_enter() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// two elements packed per 32-bit word in each plane -> two fp32 values
__device__ __forceinline__ float2 rr_join_f32x2(uint32_t hi_w, uint32_t lo_w) {
    uint32_t b0 = __byte_perm(lo_w, hi_w, 0x5410) - 0x8000u;
    uint32_t b1 = __byte_perm(lo_w, hi_w, 0x7632) - 0x8000u;
    return make_float2(__uint_as_float(b0), __uint_as_float(b1));
}

// order-preserving key: larger float <-> larger uint32 (NaNs are never keyed)
__device__ __forceinline__ uint32_t rr_float_key(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float rr_key_float(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
    return __uint_as_float(u);
}

__device__ __forceinline__ double rr_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RR_FULL, v, o);
    return v;
}
__device__ __forceinline__ float rr_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(RR_FULL, v, o);
    return v;
}
__device__ __forceinline__ float rr_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(RR_FULL, v, o));
    return v;
}
#endif
