// B3: cross-source top-k merge (/root/reference/ragroute/rerank.py:3-9 rerank_medrag,
// :28-34 rerank_wikipedia; called at http_server.py:289,293).  Also the W-way merge
// of per-shard top-k lists after the all-gather.
//
// One warp per query, a k-round select: every lane keeps the best live entry of
// its strided slice, a shuffle arg-max picks the round's winner (ties: lower input
// position), the winning lane rescans its slice.  L <= 4096 keys live in shared
// memory; for the reference's shapes (L = S*k <= 128) that is 4 keys per lane.
#include "rr_common.cuh"

namespace {
constexpr int kWarpsPerBlock = 4;
constexpr uint32_t kDead = 0xFFFFFFFFu;  // key slot of an empty / already taken entry

__device__ __forceinline__ void lane_best(const uint32_t* keys, int L, int lane, uint32_t& bk, int& bp) {
    bk = 0;
    bp = -1;
    for (int i = lane; i < L; i += 32) {
        const uint32_t kx = keys[i];
        if (kx != kDead && (bp < 0 || kx > bk)) {  // strict: earlier position wins ties
            bk = kx;
            bp = i;
        }
    }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
rr_merge_topk_kernel(const float* __restrict__ s, const int64_t* __restrict__ id, int64_t B, int L, int k,
                     int descending, float* __restrict__ os, int64_t* __restrict__ oid,
                     int32_t* __restrict__ opos) {
    extern __shared__ uint32_t keys_all[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * kWarpsPerBlock + w;
    if (b >= B) return;
    uint32_t* keys = keys_all + (size_t)w * L;
    const float* sb = s + b * L;
    const int64_t* ib = id + b * L;
    for (int i = lane; i < L; i += 32) {
        const float v = sb[i];
        // order-preserving key; ascending mode ranks by -v; 0xFFFFFFFF is reserved
        uint32_t kx = rr_float_key(descending ? v : -v);
        if (kx == kDead) kx = kDead - 1;
        keys[i] = (ib[i] < 0 || v != v) ? kDead : kx;
    }
    __syncwarp();
    uint32_t bk;
    int bp;
    lane_best(keys, L, lane, bk, bp);
    const float pad = descending ? -3.402823466e+38f : 3.402823466e+38f;
    for (int r = 0; r < k; ++r) {
        // warp arg-max over (key desc, position asc)
        uint32_t wk = bk;
        int wp = bp;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t ok = __shfl_xor_sync(RR_FULL, wk, o);
            const int op = __shfl_xor_sync(RR_FULL, wp, o);
            const bool take = (op >= 0) && (wp < 0 || ok > wk || (ok == wk && op < wp));
            if (take) {
                wk = ok;
                wp = op;
            }
        }
        if (wp < 0) {  // fewer than k live entries: pad the rest like faiss / the oracle
            for (int i = r + lane; i < k; i += 32) {
                os[b * k + i] = pad;
                oid[b * k + i] = -1;
                if (opos) opos[b * k + i] = -1;
            }
            break;
        }
        if (lane == 0) {
            os[b * k + r] = sb[wp];
            oid[b * k + r] = ib[wp];
            if (opos) opos[b * k + r] = wp;
        }
        if ((wp & 31) == lane) {  // owner of the winning slot retires it and rescans
            keys[wp] = kDead;
            lane_best(keys, L, lane, bk, bp);
        }
        __syncwarp();
    }
}
}  // namespace

extern "C" int rr_merge_topk(const float* s_dev, const int64_t* id_dev, int64_t B, int L, int k,
                             int descending, float* os_dev, int64_t* oid_dev, int32_t* opos_dev,
                             void* stream) {
    RR_REQUIRE(B >= 0 && L >= 0 && k > 0, "rr_merge_topk: bad shape B=%lld L=%d k=%d", (long long)B, L, k);
    RR_REQUIRE(L <= RR_MERGE_MAX_L, "rr_merge_topk: L=%d exceeds RR_MERGE_MAX_L=%d", L, RR_MERGE_MAX_L);
    if (B == 0) return RR_OK;
    RR_REQUIRE(os_dev && oid_dev && (L == 0 || (s_dev && id_dev)), "rr_merge_topk: NULL buffer");
    const size_t smem = (size_t)kWarpsPerBlock * L * sizeof(uint32_t);
    if (smem > (48u << 10)) {
        RR_CUDA_CHECK(cudaFuncSetAttribute(rr_merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    }
    const unsigned blocks = (unsigned)((B + kWarpsPerBlock - 1) / kWarpsPerBlock);
    rr_merge_topk_kernel<<<blocks, kWarpsPerBlock * 32, smem, (cudaStream_t)stream>>>(
        s_dev, id_dev, B, L, k, descending, os_dev, oid_dev, opos_dev);
    RR_CUDA_CHECK(cudaGetLastError());
    rr_count_launch(1);
    return RR_OK;
}
