// Warp-resident top-M lists: lane i of a warp holds entry i of a list sorted by
// (score descending, row ascending); lanes >= M hold the empty entry
// (-inf, RR_NO_ROW).  All functions are warp-collective (all 32 lanes call).
#pragma once
#include "rr_common.cuh"

// strict total order: does (s1, r1) rank before (s2, r2)?
__device__ __forceinline__ bool rr_before(float s1, uint32_t r1, float s2, uint32_t r2) {
    return (s1 > s2) || (s1 == s2 && r1 < r2);
}

// Insert one warp-uniform (v, r) into the sorted list held in (es, er).
__device__ __forceinline__ void rr_list_insert(float& es, uint32_t& er, float v, uint32_t r,
                                               int M, int lane) {
    const bool mine_first = rr_before(es, er, v, r);
    const int pos = __popc(__ballot_sync(RR_FULL, mine_first));  // sorted => prefix of lanes
    const float ps = __shfl_up_sync(RR_FULL, es, 1);
    const uint32_t pr = __shfl_up_sync(RR_FULL, er, 1);
    if (lane == pos) {
        es = v;
        er = r;
    } else if (lane > pos) {
        es = ps;
        er = pr;
    }
    if (lane >= M) {
        es = RR_NEG_INF;
        er = RR_NO_ROW;
    }
}

// compare-exchange with the lane `lane ^ j`; `desc_low` = the lower lane of the pair
// keeps the entry that ranks first.
__device__ __forceinline__ void rr_cmpxchg(float& s, uint32_t& r, int j, bool low_keeps_first,
                                           int lane) {
    const float os = __shfl_xor_sync(RR_FULL, s, j);
    const uint32_t orr = __shfl_xor_sync(RR_FULL, r, j);
    const bool i_am_low = (lane & j) == 0;
    const bool mine_first = rr_before(s, r, os, orr);
    const bool other_first = rr_before(os, orr, s, r);
    // keep mine if (I should hold the first-ranked and mine is first) or
    //              (I should hold the last-ranked and the other is first)
    const bool want_first = (i_am_low == low_keeps_first);
    const bool take_other = want_first ? other_first : mine_first;
    if (take_other) {
        s = os;
        r = orr;
    }
}

// full bitonic sort of 32 entries (one per lane) into ranking order
__device__ __forceinline__ void rr_warp_sort(float& s, uint32_t& r, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
        const bool block_desc = (lane & k) == 0 || k == 32;
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) rr_cmpxchg(s, r, j, block_desc, lane);
    }
}

// Merge a sorted 32-entry candidate set (cs, cr) into the sorted list (es, er),
// keeping the best M (<= 32).
__device__ __forceinline__ void rr_list_merge_sorted(float& es, uint32_t& er, float cs,
                                                     uint32_t cr, int M, int lane) {
    // reverse the candidates so that list ++ reversed forms a bitonic sequence
    const float rs = __shfl_sync(RR_FULL, cs, 31 - lane);
    const uint32_t rr = __shfl_sync(RR_FULL, cr, 31 - lane);
    if (rr_before(rs, rr, es, er)) {  // element-wise best half of the 64
        es = rs;
        er = rr;
    }
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) rr_cmpxchg(es, er, j, true, lane);
    if (lane >= M) {
        es = RR_NEG_INF;
        er = RR_NO_ROW;
    }
}

// Sort 32 * E entries held in shared memory into ranking order with ONE warp: entry lane * E + e lives in
// register e of its lane; stages whose partner distance is below E swap registers, the others exchange
// between lanes (no block-wide barriers: a 128-entry sort is 28 stages of ~30 instructions).
template <int E>
__device__ __noinline__ void rr_warp_sort_smem(float* s, uint32_t* r, int lane) {
    float vs[E];
    uint32_t vr[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        vs[e] = s[lane * E + e];
        vr[e] = r[lane * E + e];
    }
#pragma unroll
    for (int k = 2; k <= 32 * E; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
            for (int e = 0; e < E; ++e) {
                const bool first_low = (((lane * E + e) & k) == 0);  // the lower index keeps the first-ranked
                if (j < E) {
                    if ((e & j) == 0) {
                        const int e2 = e | j;
                        const bool swap = first_low ? rr_before(vs[e2], vr[e2], vs[e], vr[e])
                                                    : rr_before(vs[e], vr[e], vs[e2], vr[e2]);
                        if (swap) {
                            const float ts = vs[e];
                            const uint32_t tr = vr[e];
                            vs[e] = vs[e2];
                            vr[e] = vr[e2];
                            vs[e2] = ts;
                            vr[e2] = tr;
                        }
                    }
                } else {
                    rr_cmpxchg(vs[e], vr[e], j / E, first_low, lane);
                }
            }
        }
    }
#pragma unroll
    for (int e = 0; e < E; ++e) {
        s[lane * E + e] = vs[e];
        r[lane * E + e] = vr[e];
    }
}
