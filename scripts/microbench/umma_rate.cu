// Microbenchmark: cycles per tcgen05.mma (kind::f16, M=128, K=16) as a function of
// operand source (SS: A from smem, TS: A from TMEM), N, and the number of independent
// accumulator chains the issuing thread round-robins over.  No TMA: operands are
// whatever sits in (zeroed) shared memory / TMEM.  One CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../ragroute_b200/csrc/rr_ptx.cuh"
void rr_set_error(const char*, ...) {}
using namespace rr_ptx;

template <bool TS, int N, int CHAINS>
__global__ void __launch_bounds__(128, 1) k_rate(int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&slot), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, N);
        const uint32_t a_addr = smem_u32(base);            // 128 rows x 64 k = 16 KB
        const uint32_t b_addr = smem_u32(base + 32768);    // up to 256 rows x 64 k = 32 KB
        constexpr int kAcols = TS ? 32 : 0;                // A region in TMEM (4 k-steps x 8 cols)
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
                for (int c = 0; c < CHAINS; ++c) {
                    const uint32_t d = tb + kAcols + c * N;
                    if (TS) umma_bf16_ts(d, tb + k * 8, umma_desc_sw128(b_addr + k * 32), idesc, 1u);
                    else umma_bf16(d, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, 1u);
                }
            }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <bool TS, int N, int CHAINS>
void run(int iters, long long* d_out) {
    static_assert((TS ? 32 : 0) + N * CHAINS <= 512, "tmem");
    cudaFuncSetAttribute(k_rate<TS, N, CHAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) k_rate<TS, N, CHAINS><<<148, 128, 100 * 1024>>>(iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double n_inst = (double)iters * 4 * CHAINS;
    printf("%s N=%3d chains=%d : %8.1f cycles/inst  (%6.1f cycles per N=64-equivalent)  err=%s\n", TS ? "TS" : "SS", N, CHAINS,
           avg / n_inst, avg / n_inst * 64.0 / N, cudaGetErrorString(e));
}

int main() {
    long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
    const int iters = 2000;
    run<false, 64, 1>(iters, d_out);  run<false, 64, 2>(iters, d_out);  run<false, 64, 4>(iters, d_out);
    run<false, 128, 1>(iters, d_out); run<false, 128, 2>(iters, d_out);
    run<false, 256, 1>(iters, d_out); run<false, 256, 2>(iters, d_out);
    run<true, 64, 1>(iters, d_out);   run<true, 64, 2>(iters, d_out);   run<true, 64, 4>(iters, d_out);
    run<true, 128, 1>(iters, d_out);  run<true, 128, 2>(iters, d_out);
    run<true, 256, 1>(iters, d_out);
    run<false, 16, 1>(iters, d_out);  run<false, 32, 1>(iters, d_out);  run<false, 32, 4>(iters, d_out);
    run<true, 32, 1>(iters, d_out);   run<true, 32, 4>(iters, d_out);
    return 0;
}
