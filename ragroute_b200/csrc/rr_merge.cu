// B3: cross-source top-k merge (/root/reference/ragroute/rerank.py:3-9 rerank_medrag,
// :28-34 rerank_wikipedia; called at http_server.py:289,293).  Also the W-way merge
// of per-shard top-k lists after the all-gather.
//
// One warp per query, a k-round select: every lane keeps the best live entry of
// its strided slice, a warp arg-max picks the round's winner (ties: see tie_before),
// the winning lane rescans its slice.  L <= 4096 keys live in shared
// memory; for the reference's shapes (L = S*k <= 128) that is 4 keys per lane.
#include <algorithm>

#include "rr_common.cuh"
#include "rr_topm.cuh"

namespace {
constexpr int kWarpsPerBlock = 4;
constexpr uint32_t kDead = 0xFFFFFFFFu;  // key slot of an empty / already taken entry
constexpr int kSortMaxL = 256;           // up to this many entries per query: one register sort

// Input = n_seg segments of [B, seg_len] lists (one segment per rank after the all-gather; one segment
// for a plain merge); entry (seg, j) of query b has input position seg * seg_len + j.
struct MergeArgs {
    const float* s;
    const int64_t* id;
    int64_t B;
    int n_seg, seg_len;
    int64_t seg_stride_s, seg_stride_id;  // elements between the starts of two segments
    int k, descending;
    int list_len;  // > 0: also decode the winner's position into (segment, list inside the segment)
    float* os;
    int64_t* oid;
    int32_t* opos;
    int32_t* osrc;
    int32_t* orank;
};

// Order of two entries with EQUAL keys.  Plain merge (list_len == 0): lower input position first.
// Lists with source structure (list_len > 0, what rr_federation_search writes): lower source index, then
// lower id, then lower position -- an order that does not depend on how the rows were sharded over the
// ranks, so a row-sharded federation returns bit for bit what the unsharded one does even when equal
// scores meet at the k-th place (for one segment it coincides with the position order: lists are
// source-major and sorted by (score desc, row asc)).
__device__ __forceinline__ int64_t entry_id(const MergeArgs& a, int64_t b, int pos) {
    const int seg = pos / a.seg_len, j = pos - seg * a.seg_len;
    return a.id[seg * a.seg_stride_id + b * a.seg_len + j];
}
__device__ __forceinline__ bool tie_before(const MergeArgs& a, int64_t b, int i, int j) {
    if (a.list_len > 0) {
        const int si = (i % a.seg_len) / a.list_len, sj = (j % a.seg_len) / a.list_len;
        if (si != sj) return si < sj;
        const int64_t ii = entry_id(a, b, i), ij = entry_id(a, b, j);
        if (ii != ij) return ii < ij;
    }
    return i < j;
}

__device__ __forceinline__ void lane_best(const MergeArgs& a, int64_t b, const uint32_t* keys, int L, int lane,
                                          uint32_t& bk, int& bp) {
    bk = 0;
    bp = -1;
    for (int i = lane; i < L; i += 32) {
        const uint32_t kx = keys[i];
        if (kx != kDead && (bp < 0 || kx > bk || (kx == bk && tie_before(a, b, i, bp)))) {
            bk = kx;
            bp = i;
        }
    }
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32) rr_merge_topk_kernel(const MergeArgs a) {
    rr_pdl_enter();  // see rr_common.cuh: the step's kernels overlap their launch gaps
    extern __shared__ uint32_t keys_all[];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * kWarpsPerBlock + w;
    if (b >= a.B) return;
    const int L = a.n_seg * a.seg_len, k = a.k;
    const int L_pad = L <= 32 ? 32 : L <= 64 ? 64 : L <= 128 ? 128 : 256;   // (fast path only: L <= 256)
    uint32_t* keys = keys_all + (size_t)w * max(L + k, L <= kSortMaxL ? 2 * L_pad : 0);
    const float pad = a.descending ? -3.402823466e+38f : 3.402823466e+38f;
    if (L <= kSortMaxL) {
        // Fast path for the reference's shapes (L = S * k <= 256): ONE register bitonic sort of the (score,
        // position) pairs instead of k dependent select rounds (32 rounds of redux + rescan were 2/3 of the
        // kernel's 11 us).  It orders equal scores by position -- the rule of the plain merge; lists with
        // source structure order them by (source, id, position), so a query whose first k+1 sorted entries
        // contain equal scores (duplicated documents: rare) falls through to the select rounds below.
        float* ks = reinterpret_cast<float*>(keys);
        uint32_t* kr = keys + L_pad;
        int nv = 0;
        for (int i = lane; i < L_pad; i += 32) {
            float sv = RR_NEG_INF;
            uint32_t pos = 0xFFFFFFFFu;
            if (i < L) {
                const int seg = i / a.seg_len, j = i - seg * a.seg_len;
                const float v = a.s[seg * a.seg_stride_s + b * a.seg_len + j];
                const int64_t idv = a.id[seg * a.seg_stride_id + b * a.seg_len + j];
                if (idv >= 0 && v == v) {
                    sv = a.descending ? v : -v;
                    pos = (uint32_t)i;
                    ++nv;
                }
            }
            ks[i] = sv;
            kr[i] = pos;
        }
        nv = __reduce_add_sync(RR_FULL, nv);
        __syncwarp();
        switch (L_pad >> 5) {
            case 1: rr_warp_sort_smem<1>(ks, kr, lane); break;
            case 2: rr_warp_sort_smem<2>(ks, kr, lane); break;
            case 4: rr_warp_sort_smem<4>(ks, kr, lane); break;
            default: rr_warp_sort_smem<8>(ks, kr, lane); break;
        }
        __syncwarp();
        const int n_out = min(k, nv);
        bool tie = false;
        if (a.list_len > 0)
            for (int r = lane; r < n_out; r += 32) tie |= (r + 1 < nv) && ks[r] == ks[r + 1];
        if (!__any_sync(RR_FULL, tie)) {
            for (int r = lane; r < k; r += 32) {
                float sv = pad;
                int64_t iv = -1;
                int pos = -1, src = -1, rank = -1;
                if (r < n_out) {
                    pos = (int)kr[r];
                    const int seg = pos / a.seg_len, j = pos - seg * a.seg_len;
                    sv = a.s[seg * a.seg_stride_s + b * a.seg_len + j];
                    iv = a.id[seg * a.seg_stride_id + b * a.seg_len + j];
                    if (a.list_len > 0) {
                        rank = seg;
                        src = j / a.list_len;
                    }
                }
                a.os[b * k + r] = sv;
                a.oid[b * k + r] = iv;
                if (a.opos) a.opos[b * k + r] = pos;
                if (a.osrc) a.osrc[b * k + r] = src;
                if (a.orank) a.orank[b * k + r] = rank;
            }
            return;
        }
        __syncwarp();  // equal scores near the top: the select rounds below apply the full tie order
    }
    int* wins = reinterpret_cast<int*>(keys + L);  // winner of every round
    for (int i = lane; i < L; i += 32) {
        const int seg = i / a.seg_len, j = i - seg * a.seg_len;
        const float v = a.s[seg * a.seg_stride_s + b * a.seg_len + j];
        const int64_t idv = a.id[seg * a.seg_stride_id + b * a.seg_len + j];
        // order-preserving key; ascending mode ranks by -v; 0xFFFFFFFF is reserved
        uint32_t kx = rr_float_key(a.descending ? v : -v);
        if (kx == kDead) kx = kDead - 1;
        keys[i] = (idv < 0 || v != v) ? kDead : kx;
    }
    __syncwarp();
    uint32_t bk;
    int bp;
    lane_best(a, b, keys, L, lane, bk, bp);
    int n_out = k;
    for (int r = 0; r < k; ++r) {
        // warp arg-max: the largest key (one redux.sync); almost always a single lane holds it.  Key 0 never
        // occurs (it is a NaN pattern, and NaNs are dead), so it stands for "lane has nothing left".
        const uint32_t wk = __reduce_max_sync(RR_FULL, bp >= 0 ? bk : 0u);
        unsigned holders = __ballot_sync(RR_FULL, bp >= 0 && bk == wk);
        if (holders == 0) {  // fewer than k live entries
            n_out = r;
            break;
        }
        int wp = __shfl_sync(RR_FULL, bp, __ffs(holders) - 1);
        holders &= holders - 1;
        while (holders) {  // equal keys in several lanes: the tie order decides (warp-uniform)
            const int cp = __shfl_sync(RR_FULL, bp, __ffs(holders) - 1);
            holders &= holders - 1;
            if (tie_before(a, b, cp, wp)) wp = cp;
        }
        if (lane == 0) wins[r] = wp;
        if ((wp & 31) == lane) {  // owner of the winning slot retires it and rescans
            keys[wp] = kDead;
            lane_best(a, b, keys, L, lane, bk, bp);
        }
        __syncwarp();
    }
    __syncwarp();
    // write the winners out in parallel; pad the rest like faiss / the oracle
    for (int r = lane; r < k; r += 32) {
        float sv = pad;
        int64_t iv = -1;
        int pos = -1, src = -1, rank = -1;
        if (r < n_out) {
            pos = wins[r];
            const int seg = pos / a.seg_len, j = pos - seg * a.seg_len;
            sv = a.s[seg * a.seg_stride_s + b * a.seg_len + j];
            iv = a.id[seg * a.seg_stride_id + b * a.seg_len + j];
            if (a.list_len > 0) {
                rank = seg;
                src = j / a.list_len;
            }
        }
        a.os[b * k + r] = sv;
        a.oid[b * k + r] = iv;
        if (a.opos) a.opos[b * k + r] = pos;
        if (a.osrc) a.osrc[b * k + r] = src;
        if (a.orank) a.orank[b * k + r] = rank;
    }
}

int merge_launch(const MergeArgs& a, cudaStream_t stream) {
    const int L = a.n_seg * a.seg_len;
    const int L_pad = L <= 32 ? 32 : L <= 64 ? 64 : L <= 128 ? 128 : 256;
    const size_t smem = (size_t)kWarpsPerBlock * std::max(L + a.k, L <= kSortMaxL ? 2 * L_pad : 0) * sizeof(uint32_t);
    if (smem > (48u << 10)) {
        RR_CUDA_CHECK(cudaFuncSetAttribute(rr_merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    }
    const unsigned blocks = (unsigned)((a.B + kWarpsPerBlock - 1) / kWarpsPerBlock);
    RR_CUDA_CHECK(rr_launch_pdl(rr_merge_topk_kernel, dim3(blocks), dim3(kWarpsPerBlock * 32), smem, stream, a));
    rr_count_launch(1);
    return RR_OK;
}
}  // namespace

extern "C" int rr_merge_topk(const float* s_dev, const int64_t* id_dev, int64_t B, int L, int k,
                             int descending, float* os_dev, int64_t* oid_dev, int32_t* opos_dev,
                             void* stream) {
    RR_REQUIRE(B >= 0 && L >= 0 && k > 0, "rr_merge_topk: bad shape B=%lld L=%d k=%d", (long long)B, L, k);
    RR_REQUIRE(L <= RR_MERGE_MAX_L && k <= RR_MERGE_MAX_L, "rr_merge_topk: L=%d / k=%d exceeds RR_MERGE_MAX_L=%d", L, k,
               RR_MERGE_MAX_L);
    if (B == 0) return RR_OK;
    RR_REQUIRE(os_dev && oid_dev && (L == 0 || (s_dev && id_dev)), "rr_merge_topk: NULL buffer");
    MergeArgs a = {};
    a.s = s_dev;
    a.id = id_dev;
    a.B = B;
    a.n_seg = 1;
    a.seg_len = std::max(L, 1);
    if (L == 0) a.n_seg = 0;
    a.k = k;
    a.descending = descending;
    a.os = os_dev;
    a.oid = oid_dev;
    a.opos = opos_dev;
    return merge_launch(a, (cudaStream_t)stream);
}

extern "C" int rr_merge_topk_segments(const float* s_dev, const int64_t* id_dev, int64_t B, int n_seg, int seg_len,
                                      int64_t seg_stride_s, int64_t seg_stride_id, int k, int descending,
                                      int list_len, float* os_dev, int64_t* oid_dev, int32_t* opos_dev,
                                      int32_t* osrc_dev, int32_t* orank_dev, void* stream) {
    RR_REQUIRE(B >= 0 && n_seg >= 1 && seg_len >= 1 && k > 0, "rr_merge_topk_segments: bad shape B=%lld n_seg=%d seg_len=%d k=%d",
               (long long)B, n_seg, seg_len, k);
    RR_REQUIRE((int64_t)n_seg * seg_len <= RR_MERGE_MAX_L && k <= RR_MERGE_MAX_L,
               "rr_merge_topk_segments: n_seg * seg_len = %lld / k = %d exceeds RR_MERGE_MAX_L=%d",
               (long long)n_seg * seg_len, k, RR_MERGE_MAX_L);
    RR_REQUIRE(list_len >= 0 && (list_len == 0 || seg_len % list_len == 0),
               "rr_merge_topk_segments: seg_len=%d is not a multiple of list_len=%d", seg_len, list_len);
    RR_REQUIRE(n_seg == 1 || (seg_stride_s >= B * seg_len && seg_stride_id >= B * seg_len),
               "rr_merge_topk_segments: segments overlap");
    if (B == 0) return RR_OK;
    RR_REQUIRE(s_dev && id_dev && os_dev && oid_dev, "rr_merge_topk_segments: NULL buffer");
    MergeArgs a = {};
    a.s = s_dev;
    a.id = id_dev;
    a.B = B;
    a.n_seg = n_seg;
    a.seg_len = seg_len;
    a.seg_stride_s = seg_stride_s;
    a.seg_stride_id = seg_stride_id;
    a.k = k;
    a.descending = descending;
    a.list_len = list_len;
    a.os = os_dev;
    a.oid = oid_dev;
    a.opos = opos_dev;
    a.osrc = osrc_dev;
    a.orank = orank_dev;
    return merge_launch(a, (cudaStream_t)stream);
}
