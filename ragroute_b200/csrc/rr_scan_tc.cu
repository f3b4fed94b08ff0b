// The scoring pass: scores = X_bf16 . Q_bf16^T on tcgen05 tensor cores with the
// per-query top-M selection fused into the TMEM epilogue, so no [B, N] score
// matrix is ever written.  Replaces the arithmetic of faiss IndexFlatIP::search
// (/root/reference/ragroute/data_source.py:158,186,203) up to the exact rescoring
// done in rr_index.cu.
//
// Shape of the problem at B <= 64: arithmetic intensity = B flop/byte over the
// bf16 plane, far below the ridge (~215), so the kernel is built around the HBM
// stream, not around MMA throughput:
//   * corpus rows are the UMMA M operand (128 rows x 64 k per stage, 16 KB, TMA,
//     128B swizzle), so even one query fills the tensor core ("swap-AB");
//   * the query tile (<= 64 x d bf16) is the UMMA N operand and stays resident in
//     shared memory for the whole kernel (loaded once by TMA);
//   * a ring of 4-8 corpus stages keeps >= 64 KB in flight per SM;
//   * persistent grid of one CTA per SM, tiles round-robin so neighbouring SMs
//     stream neighbouring rows;
//   * accumulators: 4 TMEM stages of QT fp32 columns; the epilogue of tile i runs
//     under the MMAs of tiles i+1..i+3.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..5 = epilogue
// (each owns the 32 TMEM lanes = 32 corpus rows of its quadrant and keeps, per
// query, a sorted top-16 list in shared memory; see rr_topm.cuh).
#include <algorithm>
#include <atomic>

#include "rr_index_internal.cuh"
#include "rr_ptx.cuh"
#include "rr_topm.cuh"

namespace {

constexpr int kStageBytes = RR_TC_TILE_ROWS * RR_TC_KCHUNK * 2;  // 16 KB
constexpr int kAccStages = 4;
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kDepth = RR_TC_LIST_DEPTH;
constexpr uint64_t kEvictFirst = rr_ptx::kTmaEvictFirst;
constexpr uint64_t kEvictLast = rr_ptx::kTmaEvictLast;

struct TcParams {
    int64_t n_rows;
    int n_tiles;
    int num_kc;   // d_pad / 64
    int nq_tile;  // valid queries in this pass (<= QT)
    int q_row0;   // first row of the pass inside the query buffer
    int stages;
    int chained;  // 1: directly follows another scan of the same step (see the PDL notes in rr_tc_scan_kernel)
    const uint8_t* active;   // [nq_tile]
    const int* any_active;   // device flag: 0 -> nothing routed here, exit at once
    rr_scan_counters* counters;  // profiling: bytes streamed / flops scored by passes that really ran (or NULL)
    float* cand_s;           // [query][list][kDepth]
    uint32_t* cand_r;
    int64_t cand_stride;
};

using namespace rr_ptx;

// Candidate handling for one (epilogue warp, query): `b` = lanes whose score beats
// the list's current tail.  Kept out of line: it runs for a small fraction of the
// (tile, query) pairs.
__device__ __noinline__ void epi_take(float v, unsigned b, uint32_t row, float* ls_q, uint32_t* lr_q,
                                      float* thr_q, int lane) {
    float es = RR_NEG_INF;
    uint32_t er = RR_NO_ROW;
    if (lane < kDepth) {
        es = ls_q[lane];
        er = lr_q[lane];
    }
    if (__popc(b) <= 3) {
        while (b) {
            const int src = __ffs(b) - 1;
            b &= b - 1;
            const float cv = __shfl_sync(RR_FULL, v, src);
            const uint32_t cr = __shfl_sync(RR_FULL, row, src);
            rr_list_insert(es, er, cv, cr, kDepth, lane);
        }
    } else {
        float cs = ((b >> lane) & 1u) ? v : RR_NEG_INF;
        uint32_t cr = ((b >> lane) & 1u) ? row : RR_NO_ROW;
        rr_warp_sort(cs, cr, lane);
        rr_list_merge_sorted(es, er, cs, cr, kDepth, lane);
    }
    if (lane < kDepth) {
        ls_q[lane] = es;
        lr_q[lane] = er;
    }
    const float tail = __shfl_sync(RR_FULL, es, kDepth - 1);
    const uint32_t tail_r = __shfl_sync(RR_FULL, er, kDepth - 1);
    if (lane == 0) *thr_q = (tail_r != RR_NO_ROW) ? tail : RR_NEG_INF;
    __syncwarp();
}

template <int QT>
__global__ void __launch_bounds__(kThreads, 1)
rr_tc_scan_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_q,
                  const TcParams p) {
    // Programmatic dependent launch.  The scans of one step touch disjoint memory, so a CHAINED scan starts at
    // once (its predecessor scan has already waited for everything before the chain) and lets its successor
    // start too; the FIRST scan of a chain waits here for the kernels before it (query preparation) and only
    // then releases its successor.  Every scan waits again before it exits, so it completes after its
    // predecessor and whatever follows the chain sees all of it done.
    if (!p.chained) pdl_wait_primary();
    pdl_launch_dependents();  // the next scan of the step may take SMs as they free up
    if (*p.any_active == 0) {  // no query routed to this source: do not touch the corpus
        pdl_wait_primary();
        return;
    }
    if (p.counters && blockIdx.x == 0 && threadIdx.x == 0) {
        int n_act = 0;
        for (int q = 0; q < p.nq_tile; ++q) n_act += p.active[q] ? 1 : 0;
        atomicAdd(&p.counters->bytes, (unsigned long long)p.n_rows * (unsigned long long)p.num_kc * 128ull);
        atomicAdd(&p.counters->flops,
                  2ull * (unsigned long long)n_act * (unsigned long long)p.n_rows * (unsigned long long)p.num_kc * 64ull);
    }

    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    // carve (all offsets multiples of 1024 for the swizzled tiles)
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                               ~(uintptr_t)1023);
    const int q_bytes = p.num_kc * QT * 128;
    uint8_t* q_sm = base;
    uint8_t* x_sm = base + q_bytes;
    float* ls = reinterpret_cast<float*>(x_sm + p.stages * kStageBytes);      // [4][QT][kDepth]
    uint32_t* lr = reinterpret_cast<uint32_t*>(ls + RR_TC_EPI_WARPS * QT * kDepth);
    float* thr = reinterpret_cast<float*>(lr + RR_TC_EPI_WARPS * QT * kDepth);  // [4][QT]
    uint64_t* bars = reinterpret_cast<uint64_t*>(thr + RR_TC_EPI_WARPS * QT);
    // bars: full[kMaxStages], empty[kMaxStages], acc_full[kAccStages], acc_empty[kAccStages], qfull
    uint64_t* bar_full = bars;
    uint64_t* bar_empty = bars + kMaxStages;
    uint64_t* bar_acc_full = bars + 2 * kMaxStages;
    uint64_t* bar_acc_empty = bar_acc_full + kAccStages;
    uint64_t* bar_q = bar_acc_empty + kAccStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_q + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = (kAccStages * QT < 32) ? 32 : kAccStages * QT;  // 64/128/256

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        tma_prefetch_desc(&tmap_q);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(smem_u32(bar_full + i), 1);
            mbar_init(smem_u32(bar_empty + i), 1);
        }
        for (int i = 0; i < kAccStages; ++i) {
            mbar_init(smem_u32(bar_acc_full + i), 1);
            mbar_init(smem_u32(bar_acc_empty + i), RR_TC_EPI_WARPS);
        }
        mbar_init(smem_u32(bar_q), 1);
        fence_barrier_init();
    } else if (warp == 1) {
        tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Warps 0 and 1 run their loops converged and elect one lane for the single-thread
    // instructions, so that addresses and descriptors stay in uniform registers (an
    // `if (lane == 0)` loop costs ~150 issue cycles per MMA and caps the scan below HBM speed).
    if (warp == 0) {
        // ===================== TMA producer =====================
        const uint32_t x_base = smem_u32(x_sm);
        if (elect_one_sync()) {
            mbar_arrive_expect_tx(smem_u32(bar_q), (uint32_t)q_bytes);
            for (int kc = 0; kc < p.num_kc; ++kc)
                tma_load_2d(smem_u32(q_sm + kc * QT * 128), &tmap_q, smem_u32(bar_q), kc * RR_TC_KCHUNK,
                            p.q_row0, kEvictLast);
        }
        __syncwarp();
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            for (int kc = 0; kc < p.num_kc; ++kc) {
                mbar_wait(smem_u32(bar_empty + stage), phase ^ 1u);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(smem_u32(bar_full + stage), kStageBytes);
                    tma_load_2d(x_base + stage * kStageBytes, &tmap_x, smem_u32(bar_full + stage),
                                kc * RR_TC_KCHUNK, tile * RR_TC_TILE_ROWS, kEvictFirst);
                }
                __syncwarp();
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =======================
        constexpr uint32_t idesc = umma_idesc_bf16(RR_TC_TILE_ROWS, QT);
        mbar_wait(smem_u32(bar_q), 0);
        tc_fence_after();
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        const uint32_t x_base = smem_u32(x_sm);
        const uint32_t q_base = smem_u32(q_sm);
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(smem_u32(bar_acc_empty + acc), acc_phase ^ 1u);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * QT);
            for (int kc = 0; kc < p.num_kc; ++kc) {
                mbar_wait(smem_u32(bar_full + stage), phase);
                tc_fence_after();
                const uint64_t a_desc = kUmmaDescSw128Hi | (uint64_t)(((x_base + stage * kStageBytes) >> 4) & 0x3FFFu);
                const uint64_t b_desc = kUmmaDescSw128Hi | (uint64_t)(((q_base + kc * QT * 128) >> 4) & 0x3FFFu);
                if (elect_one_sync()) {
                    umma_bf16(d_tmem, a_desc, b_desc, idesc, (uint32_t)(kc != 0));
                    umma_bf16(d_tmem, a_desc + 2, b_desc + 2, idesc, 1u);
                    umma_bf16(d_tmem, a_desc + 4, b_desc + 4, idesc, 1u);
                    umma_bf16(d_tmem, a_desc + 6, b_desc + 6, idesc, 1u);
                    umma_commit(smem_u32(bar_empty + stage));  // frees the smem stage when the MMAs retire
                    if (kc == p.num_kc - 1) umma_commit(smem_u32(bar_acc_full + acc));  // accumulator -> epilogue
                }
                __syncwarp();
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (++acc == kAccStages) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
    } else {
        // ===================== epilogue warps 2..5 ===========================
        const int ew = warp - 2;
        const int quad = warp & 3;  // TMEM lanes [32*quad, 32*quad+32) are this warp's
        float* ls_w = ls + ew * QT * kDepth;
        uint32_t* lr_w = lr + ew * QT * kDepth;
        float* thr_w = thr + ew * QT;
        for (int i = lane; i < QT * kDepth; i += 32) {
            ls_w[i] = RR_NEG_INF;
            lr_w[i] = RR_NO_ROW;
        }
        for (int q = lane; q < QT; q += 32)
            thr_w[q] = (q < p.nq_tile && p.active[q]) ? RR_NEG_INF : __builtin_huge_valf();
        __syncwarp();

        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(smem_u32(bar_acc_full + acc), acc_phase);
            tc_fence_after();
            const int64_t row64 = (int64_t)tile * RR_TC_TILE_ROWS + quad * 32 + lane;
            const bool valid = row64 < p.n_rows;
            const uint32_t row = (uint32_t)row64;
            const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * QT);
#pragma unroll 1
            for (int c0 = 0; c0 < QT; c0 += 16) {
                float v[16];
                tmem_ld_x16(t_row + (uint32_t)c0, v);
                unsigned mine = 0;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 t4 = *reinterpret_cast<const float4*>(thr_w + c0 + j4 * 4);
                    mine |= (v[j4 * 4 + 0] > t4.x ? 1u : 0u) << (j4 * 4 + 0);
                    mine |= (v[j4 * 4 + 1] > t4.y ? 1u : 0u) << (j4 * 4 + 1);
                    mine |= (v[j4 * 4 + 2] > t4.z ? 1u : 0u) << (j4 * 4 + 2);
                    mine |= (v[j4 * 4 + 3] > t4.w ? 1u : 0u) << (j4 * 4 + 3);
                }
                if (!valid) mine = 0;
                if (__any_sync(RR_FULL, mine != 0)) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const unsigned b = __ballot_sync(RR_FULL, (mine >> j) & 1u);
                        if (b)
                            epi_take(v[j], b, row, ls_w + (c0 + j) * kDepth, lr_w + (c0 + j) * kDepth,
                                     thr_w + c0 + j, lane);
                    }
                }
            }
            // all TMEM reads of this stage are complete (wait::ld): hand it back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(bar_acc_empty + acc));
            if (++acc == kAccStages) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        // Fold the four warps' lists of every query into ONE list per CTA (the best kDepth of the union: any
        // row dropped here or earlier scores <= the merged tail, so the list keeps its meaning) and publish
        // it: [query][list = cta][kDepth].  A quarter of the candidates for the finalize to select from.
        asm volatile("bar.sync 1, %0;" ::"n"(RR_TC_EPI_WARPS * 32) : "memory");  // the 4 epilogue warps only
        for (int q = ew; q < p.nq_tile; q += RR_TC_EPI_WARPS) {
            float es = RR_NEG_INF;
            uint32_t er = RR_NO_ROW;
            if (lane < kDepth) {
                es = ls[q * kDepth + lane];
                er = lr[q * kDepth + lane];
            }
#pragma unroll
            for (int w2 = 1; w2 < RR_TC_EPI_WARPS; ++w2) {
                float cs = RR_NEG_INF;
                uint32_t cr = RR_NO_ROW;
                if (lane < kDepth) {
                    cs = ls[(w2 * QT + q) * kDepth + lane];
                    cr = lr[(w2 * QT + q) * kDepth + lane];
                }
                rr_list_merge_sorted(es, er, cs, cr, kDepth, lane);
            }
            if (lane < kDepth) {
                const int64_t o = (int64_t)q * p.cand_stride + (int64_t)blockIdx.x * kDepth + lane;
                p.cand_s[o] = es;
                p.cand_r[o] = er;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
    pdl_wait_primary();  // complete only after the previous scan of the chain has
}

// ------------------------------------------------------------------ host ----

struct SmemPlan {
    int stages;
    size_t bytes;
};

SmemPlan plan_smem(int d_pad, int qt) {
    const size_t q_bytes = (size_t)d_pad * 2 * qt;
    const size_t lists = (size_t)RR_TC_EPI_WARPS * qt * kDepth * 8 + (size_t)RR_TC_EPI_WARPS * qt * 4;
    const size_t bars = (2 * kMaxStages + 2 * kAccStages + 1) * 8 + 16;
    const size_t fixed = 1024 /*alignment slack*/ + q_bytes + lists + bars;
    SmemPlan pl{0, 0};
    if (fixed + 2 * (size_t)kStageBytes > (size_t)kSmemLimit) return pl;
    pl.stages = (int)std::min<size_t>(kMaxStages, ((size_t)kSmemLimit - fixed) / kStageBytes);
    pl.bytes = fixed + (size_t)pl.stages * kStageBytes;
    return pl;
}

template <int QT>
int launch_qt(rr_index* h, const CUtensorMap& tq, const TcParams& p, int grid, size_t smem, bool chain,
              cudaStream_t stream) {
    static std::atomic<bool> attr_done[RR_MAX_DEVICES];  // per (instantiation, device); a double set is harmless
    if (!attr_done[h->device].load(std::memory_order_acquire)) {
        RR_CUDA_CHECK(cudaFuncSetAttribute(rr_tc_scan_kernel<QT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kSmemLimit));
        attr_done[h->device].store(true, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = rr_pdl_enabled() ? 1 : 0;
    (void)chain;
    RR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, rr_tc_scan_kernel<QT>, h->tmap_hi, tq, p));
    return RR_OK;
}

}  // namespace

int rr_tc_scan_qtile(const rr_index* h) {
    if (h->d_pad <= 1024) return 64;
    if (h->d_pad <= 2048) return 32;
    if (h->d_pad <= 4096) return 16;
    return 0;
}

int rr_tc_scan_num_lists(const rr_index* h) { return h->sm_count; }

int rr_tc_scan_launch(rr_index* h, const uint16_t* qbf16_dev, int64_t q_rows_total, int q_row0, int nq_tile,
                      const uint8_t* active_dev, const int* any_active_dev, float* cand_s, uint32_t* cand_r,
                      int64_t cand_stride, int* n_lists_out, rr_scan_counters* counters_dev, bool chain,
                      cudaStream_t stream) {
    const int qt = rr_tc_scan_qtile(h);
    RR_REQUIRE(qt > 0 && nq_tile >= 1 && nq_tile <= qt, "rr_tc_scan_launch: bad query tile");
    const SmemPlan pl = plan_smem(h->d_pad, qt);
    RR_REQUIRE(pl.stages >= 2, "rr_tc_scan_launch: shared memory plan failed for d_pad=%d", h->d_pad);
    if (h->tmap_hi_ptr != h->hi || h->tmap_hi_rows != h->n) {
        int rc = rr_make_tmap_bf16_rows(&h->tmap_hi, h->hi, h->n, h->d_pad, RR_TC_TILE_ROWS);
        if (rc != RR_OK) return rc;
        h->tmap_hi_ptr = h->hi;
        h->tmap_hi_rows = h->n;
    }
    alignas(64) CUtensorMap tq;
    int rc = rr_make_tmap_bf16_rows(&tq, qbf16_dev, q_rows_total, h->d_pad, qt);
    if (rc != RR_OK) return rc;

    TcParams p;
    p.n_rows = h->n;
    p.n_tiles = (int)((h->n + RR_TC_TILE_ROWS - 1) / RR_TC_TILE_ROWS);
    p.num_kc = h->d_pad / RR_TC_KCHUNK;
    p.nq_tile = nq_tile;
    p.q_row0 = q_row0;
    p.stages = pl.stages;
    p.chained = (chain && rr_pdl_enabled()) ? 1 : 0;
    p.active = active_dev;
    p.any_active = any_active_dev;
    p.counters = counters_dev;
    p.cand_s = cand_s;
    p.cand_r = cand_r;
    p.cand_stride = cand_stride;
    const int grid = std::min(h->sm_count, p.n_tiles);
    *n_lists_out = grid;  // one merged list per CTA
    switch (qt) {
        case 64: return launch_qt<64>(h, tq, p, grid, pl.bytes, chain, stream);
        case 32: return launch_qt<32>(h, tq, p, grid, pl.bytes, chain, stream);
        default: return launch_qt<16>(h, tq, p, grid, pl.bytes, chain, stream);
    }
}
