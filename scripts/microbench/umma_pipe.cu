// Microbenchmark: steady-state rate of the large-batch scan pipeline without its epilogue.
// Per CTA: warp 0 streams 16 KB corpus chunks (128 rows x 64 k, 128B swizzle) by TMA from an
// L2-resident buffer into a ring of stages; warp 1 issues 4 tcgen05.mma (M=128, N=128, K=16,
// A from TMEM) per chunk.  Variants: TMA on/off (off = MMAs re-read one resident stage),
// MMA on/off (off = stages are released at once).  Reports cycles per 16 KB chunk; the
// tensor pipe needs 256.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_pipe umma_pipe.cu rr_tma_host.cu(see build line)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../ragroute_b200/csrc/rr_index_internal.cuh"
#include "../../ragroute_b200/csrc/rr_ptx.cuh"
using namespace rr_ptx;
void rr_set_error(const char* fmt, ...) { (void)fmt; }

constexpr int kChunk = 16384;

template <bool TMA, bool MMA, int STAGE_KC>
__global__ void __launch_bounds__(192, 1) k_pipe(const __grid_constant__ CUtensorMap tmap, int n_chunks, int rows_total,
                                                 long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    constexpr int kStages = 128 * 1024 / (STAGE_KC * kChunk);
    __shared__ uint64_t bar_full[16], bar_empty[16], bar_done;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(smem_u32(bar_full + i), 1); mbar_init(smem_u32(bar_empty + i), 1); }
        mbar_init(smem_u32(&bar_done), 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = slot;
    const uint32_t x_base = smem_u32(base);
    const int n_it = n_chunks / STAGE_KC;
    long long t0 = clock64();
    if (warp == 0) {
        int stage = 0; uint32_t phase = 0;
        int row = (blockIdx.x * 128) % rows_total;
        for (int it = 0; it < n_it; ++it) {
            mbar_wait(smem_u32(bar_empty + stage), phase ^ 1u);
            if (elect_one_sync()) {
                if (TMA) {
                    mbar_arrive_expect_tx(smem_u32(bar_full + stage), STAGE_KC * kChunk);
                    for (int c = 0; c < STAGE_KC; ++c)
                        tma_load_2d(x_base + stage * STAGE_KC * kChunk + c * kChunk, &tmap, smem_u32(bar_full + stage),
                                    ((it * STAGE_KC + c) % 12) * 64, row, kTmaEvictNormal);
                } else {
                    mbar_arrive(smem_u32(bar_full + stage));
                }
            }
            __syncwarp();
            if ((it * STAGE_KC) % 12 + STAGE_KC >= 12) row = (row + 148 * 128) % rows_total;
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
        int stage = 0; uint32_t phase = 0;
        for (int it = 0; it < n_it; ++it) {
            mbar_wait(smem_u32(bar_full + stage), phase);
            tc_fence_after();
            const uint64_t bd0 = kUmmaDescSw128Hi | (uint64_t)(((x_base + stage * STAGE_KC * kChunk) >> 4) & 0x3FFFu);
            if (elect_one_sync()) {
                if (MMA) {
#pragma unroll
                    for (int c = 0; c < STAGE_KC; ++c) {
                        const uint64_t bd = bd0 + (uint64_t)(c * (kChunk >> 4));
                        const uint32_t at = tb + (uint32_t)(((it * STAGE_KC + c) % 8) * 32);
                        const uint32_t d = tb + 256 + (uint32_t)(((it * STAGE_KC + c) / 12 % 2) * 128);
                        umma_bf16_ts(d, at, bd, idesc, 1u);
                        umma_bf16_ts(d, at + 8, bd + 2, idesc, 1u);
                        umma_bf16_ts(d, at + 16, bd + 4, idesc, 1u);
                        umma_bf16_ts(d, at + 24, bd + 6, idesc, 1u);
                    }
                    umma_commit(smem_u32(bar_empty + stage));
                } else {
                    mbar_arrive(smem_u32(bar_empty + stage));
                }
                if (it == n_it - 1) { if (MMA) umma_commit(smem_u32(&bar_done)); else mbar_arrive(smem_u32(&bar_done)); }
            }
            __syncwarp();
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        mbar_wait(smem_u32(&bar_done), 0);
        if (threadIdx.x == 32) out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tb, 512); }
}

template <bool TMA, bool MMA, int STAGE_KC>
void run(const CUtensorMap& tm, int rows_total, long long* d_out, const char* what) {
    const int n_chunks = 12 * 400;
    cudaFuncSetAttribute(k_pipe<TMA, MMA, STAGE_KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
    for (int rep = 0; rep < 2; ++rep) k_pipe<TMA, MMA, STAGE_KC><<<148, 192, 130 * 1024>>>(tm, n_chunks, rows_total, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0, mx = 0; for (int i = 0; i < 148; ++i) { avg += h[i]; if (h[i] > mx) mx = h[i]; } avg /= 148;
    printf("%-34s stage=%2d KB : %7.1f cycles / 16 KB chunk (max CTA %7.1f)  err=%s\n", what, STAGE_KC * 16, avg / n_chunks,
           mx / n_chunks, cudaGetErrorString(e));
}

int main() {
    const int rows_total = 148 * 128 * 8;  // 151552 rows x 768 x 2 B = 233 MB (> L2: mostly HBM) -- see second buffer below
    long long* d_out; cudaMalloc(&d_out, 148 * sizeof(long long));
    for (int pass = 0; pass < 2; ++pass) {
        const int rows = pass == 0 ? 148 * 128 * 2 : rows_total;  // 58 MB (L2 resident) / 233 MB (HBM)
        void* buf; cudaMalloc(&buf, (size_t)rows * 768 * 2); cudaMemset(buf, 0, (size_t)rows * 768 * 2);
        alignas(64) CUtensorMap tm;
        if (rr_make_tmap_bf16_rows(&tm, buf, rows, 768, 128) != 0) { printf("tmap failed\n"); return 1; }
        printf("--- corpus buffer %d rows (%.0f MB)\n", rows, rows * 1536.0 / 1e6);
        run<true, true, 1>(tm, rows, d_out, "TMA + MMA");
        run<true, true, 2>(tm, rows, d_out, "TMA + MMA");
        run<false, true, 1>(tm, rows, d_out, "MMA only (resident stages)");
        run<false, true, 2>(tm, rows, d_out, "MMA only (resident stages)");
        run<true, false, 1>(tm, rows, d_out, "TMA only");
        run<true, false, 2>(tm, rows, d_out, "TMA only");
        cudaFree(buf);
    }
    return 0;
}
