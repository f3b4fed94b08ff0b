// The scoring pass for LARGE query batches (more than one 64-query pass):
// scores = Q_bf16 . X_bf16^T with the QUERIES as the stationary operand.
// Replaces the arithmetic of faiss IndexFlatIP::search
// (/root/reference/ragroute/data_source.py:158,186,203) for batched callers, up to the
// exact rescoring done in rr_index.cu.
//
// With B >= 128 queries the arithmetic intensity over the bf16 plane (B flop/byte)
// approaches / exceeds the ridge (~215), so the corpus must be read from HBM once per
// chunk of up to 2048 queries and every byte brought into an SM must be used against as
// many queries as possible:
//   * a tile of 128 queries is stationary for the whole kernel as the UMMA A operand: its
//     first 512 dimensions live in TENSOR MEMORY (TS form of tcgen05.mma: row m of A in
//     TMEM lane m, 2 bf16 per 32-bit column = 256 columns, written once by tcgen05.st),
//     the remaining <= 512 dimensions (d_pad <= 1024) in shared memory (SS form, loaded once by
//     TMA).  The split leaves 256 TMEM columns = TWO accumulator stages, so the epilogue of tile i
//     runs under the MMAs of tile i+1 (with 768 dims in TMEM only one stage fits and the
//     tensor pipe idles ~25 % of the time waiting for the accumulator to drain -- which is
//     the layout of choice when a single query tile makes the launch HBM-bound anyway);
//   * corpus rows stream through shared memory as the B operand, 128 rows x 64 k per
//     16 KB TMA stage (128B swizzle), ring of 12 stages (192 KB in flight per SM); UMMA
//     N = 128 because one tcgen05.mma costs >= 45 cycles whatever its N (measured:
//     scripts/microbench/umma_rate.cu), so N = 64 would waste 30 % of the tensor pipe;
//   * two accumulator stages of 128 fp32 columns (128 queries x 128 corpus rows each);
//   * epilogue orientation: TMEM lane = query, so every epilogue THREAD owns one query:
//     its running threshold sits in a register and the 128 scores of a tile cost one max
//     tree + one compare.  Survivors are first appended (predicated stores, no branches)
//     to the thread's small pending queue in shared memory, then drained in SIMT ROUNDS:
//     in one round every thread with a pending survivor overwrites the minimum of its own
//     unsorted top-32 list and rescans the list for the new minimum (32 independent
//     loads) -- one round serves all 32 queries of the warp at once.  Measured insert
//     costs that made the epilogue the bottleneck before: per-thread SORTED insert ~2000
//     cycles, warp-cooperative replace-the-minimum ~245 cycles per insert (serial).
// CTA c works on query tile (c % nqt) and corpus slice (c / nqt): the nqt CTAs that share
// a slice stream the same rows at the same time, so HBM is read once and the other
// nqt - 1 reads hit L2.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc), 2..5 = A loader + epilogue.
#include <stdlib.h>

#include <algorithm>
#include <atomic>

#include "rr_index_internal.cuh"
#include "rr_ptx.cuh"

namespace {
using namespace rr_ptx;

constexpr int kNT = RR_GEMM_TILE_ROWS;               // corpus rows per MMA (UMMA N)
constexpr int kChunkBytes = kNT * RR_TC_KCHUNK * 2;  // one TMA box: 128 rows x 64 k = 16 KB
// K-chunks per pipeline stage (template parameter SKC): 2 = one barrier wait + one commit per 8 MMAs (32 KB stages
// per query-tile unit); 4 = per 16 MMAs (64 KB stages, CTA pairs in the tensor-bound regime: every stage boundary
// costs the tensor pipe ~120-150 idle cycles, profiles/microbench/r01_umma_pipe.txt)
constexpr int kStageKc = 2;
constexpr int kDepth = RR_GEMM_LIST_DEPTH;
constexpr int kQTile = RR_GEMM_QTILE;  // 128 queries = TMEM lanes
constexpr int kThreads = 192;
constexpr int kMaxStages = 8;
constexpr int kMaxAcc = 2;
constexpr int kMaxATmemKc = ((int)512 - kMaxAcc * RR_GEMM_TILE_ROWS) / 32;  // 8 K-chunks = 512 dims of A in TMEM
constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxDpad = 1024;
constexpr uint32_t kTmemCols = 512;

struct GemmParams {
    int64_t n_rows;
    int n_tiles;     // ceil(n_rows / kNT)
    int num_kc;      // d_pad / 64
    int d_pad;
    int stages;
    int acc_stages;
    int a_tmem_kc;   // leading K-chunks of the query tile held in TMEM
    int a_smem_kc;   // trailing K-chunks held in shared memory
    uint64_t l2_hint;
    uint64_t l2_hint_shared;  // when several query tiles share a corpus slice
    const uint16_t* qbf16;    // [ceil(nq / 128) * 128][d_pad] bf16: routed queries first, then zero rows
    rr_chunk_ctl* ctl;        // n_active (in), n_lists (out)
    rr_scan_counters* counters;
    // COLLECT mode (second pass for queries whose certificate failed): fixed thresholds, every row that
    // beats its query's threshold is appended to that query's row list
    const float* theta;     // [compact query] rows with score > theta are collected
    uint32_t* col_rows;     // [compact query][col_cap]
    int* col_cnt;           // [compact query] (may exceed col_cap: overflow, the caller falls back)
    int col_cap;
    int min_active, max_active;  // run only if min_active <= routed queries <= max_active (see rr_gemm_scan_launch)
    int chained;                 // 1: directly follows another scan of the same step (PDL, see rr_scan_tc.cu)
    int trace;                   // RR_TRACE: block 0 records a timeline (debugging aid)
    // lockstep of the units that share a corpus slice (see the producer): progress[unit] = progress_base + tiles
    // issued, progress_base = launch number << 32 (entries of earlier launches read as "not started"); NULL = off
    unsigned long long* progress;
    unsigned long long progress_base;
    int lock_window;  // tiles a unit's producer may run ahead of the slowest unit of its slice
    float* cand_s;  // [compact query][slice][kDepth]
    uint32_t* cand_r;
    int64_t cand_stride;
};

// Timeline tracing (debugging aid): a library built with -DRR_TRACE_BUILD (RR_BUILD_TRACE=1 for
// __graft_entry__.build_cuda) and run with RR_TRACE=1 has block 0 record (tag, tile, globaltimer) events of its
// three roles into a device array that rr_debug_read_trace copies out.
constexpr int kTraceCap = 4096;
__device__ unsigned long long g_trace[kTraceCap * 2];
__device__ unsigned int g_trace_n;
#ifdef RR_TRACE_BUILD
__device__ __forceinline__ void trace_ev(int on, int tag, int tile) {
    if (!on) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    const unsigned i = atomicAdd(&g_trace_n, 1u);
    if (i < kTraceCap) {
        g_trace[2 * i] = ((unsigned long long)tag << 32) | (unsigned)tile;
        g_trace[2 * i + 1] = t;
    }
}
#else
// (compiled out of the product build: a lane-dependent branch in the MMA issuer's converged loop takes its
// descriptors out of the uniform registers -- measured ~10 % of the tensor rate at batch 2048)
#define trace_ev(on, tag, tile) ((void)0)
#endif

// Per-thread state: unsorted top-kDepth list of the thread's query in shared memory, entry e of
// thread t at [e * kQTile + t] (conflict-free across a warp); `thr` = the list's minimum score
// (-inf while an empty slot remains), `pos` = where that minimum sits.
// Drain the pending queues of a warp: round i inserts the i-th pending survivor of every thread
// that has one.  Out of line: it runs for a fraction of the tiles.
constexpr int kPend = 16;  // pending slots per thread
// Lockstep: a unit's producer does not run more than lock_window corpus tiles ahead of the slowest unit sharing its
// slice, polling at most kLockPolls times per tile (bounded: a sibling whose CTA has not started yet -- the
// previous scan of the chain still holds its SM -- delays a tile by ~50 us at most, it never blocks it).
constexpr int kLockPolls = 48;
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// (a max, not a store: when two chunks of one batch are scanned back to back, the second launch's units start
// while the first's are still running and publish to the same entries; an entry never goes back to the older launch)
__device__ __forceinline__ void red_max_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// COLLECT mode: the pending survivors go to the query's global row list (thresholds are fixed).
__device__ __noinline__ void epi_collect(int cnt, int qg, const GemmParams& p, const uint32_t* pr, int t) {
    for (int i = 0; i < cnt; ++i) {
        const int pos = atomicAdd(p.col_cnt + qg, 1);
        if (pos < p.col_cap) p.col_rows[(int64_t)qg * p.col_cap + pos] = pr[i * kQTile + t];
    }
}

template <int DEPTH>
__device__ __noinline__ void epi_drain(int cnt, float& thr, int& pos, float* ls, uint32_t* lr, const float* ps,
                                       const uint32_t* pr, int t) {
    int i = 0;
    while (__any_sync(RR_FULL, i < cnt)) {
        if (i < cnt) {
            const float v = ps[i * kQTile + t];
            if (v > thr) {  // an earlier insert of this drain may have raised the threshold
                ls[pos * kQTile + t] = v;
                lr[pos * kQTile + t] = pr[i * kQTile + t];
                // new minimum of the list: four independent chains (a single chain of DEPTH - 1 dependent
                // compare-selects made a round cost ~900 cycles: measured with RR_TRACE on a 125 k-row corpus)
                float m[4];
                int am[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    m[c] = ls[c * kQTile + t];
                    am[c] = c;
                }
#pragma unroll
                for (int e = 4; e < DEPTH; e += 4) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float sv = ls[(e + c) * kQTile + t];
                        if (sv < m[c]) {
                            m[c] = sv;
                            am[c] = e + c;
                        }
                    }
                }
                if (m[1] < m[0]) {
                    m[0] = m[1];
                    am[0] = am[1];
                }
                if (m[3] < m[2]) {
                    m[2] = m[3];
                    am[2] = am[3];
                }
                if (m[2] < m[0]) {
                    m[0] = m[2];
                    am[0] = am[2];
                }
                thr = m[0];
                pos = am[0];
            }
        }
        ++i;
    }
    __syncwarp();
}

static_assert(kNT == RR_TC_TILE_ROWS, "both scan kernels share the 128-row TMA box of the hi plane");

// CG = 1: one CTA per work unit (128 queries x 128-row corpus tiles).
// CG = 2: a CTA PAIR per work unit (cluster of 2, tcgen05 cta_group::2): 256 queries (128 per
//         CTA, each in its own TMEM) against the same 128-row corpus tiles, of which each CTA
//         streams only 64 rows -- per SM half the L2->smem traffic and half the shared-memory
//         operand reads per flop, which is what bounds CG = 1.
// COLLECT = false: per-query top-DEPTH lists (first pass).  COLLECT = true: fixed per-query thresholds, every
// row above them is reported (second pass for the queries whose certificate failed).
// DEPTH = entries per (corpus slice, query) list: 32 for large shards; 8 for small ones (<= ~16 k rows per
// list), where a cold list of 32 lets ~32 (1 + ln(rows / 32)) rows through the insert path -- ~170 SIMT rounds
// per warp on a 125 k-row corpus, 80 us of a launch whose HBM time is 29 us (RR_TRACE timeline) -- while
// 148 lists x 8 still hold far more than the k' = 128 candidates the finalize wants.
template <int CG, bool COLLECT, int DEPTH, int SKC>
__global__ void __launch_bounds__(kThreads, 1)
rr_gemm_scan_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_q,
                    const GemmParams p) {
    // Work split, decided on the device from the number of ROUTED queries: nqt query tiles of
    // 128 * CG (compacted, so unrouted queries cost nothing) x n_slices corpus slices.
    if (!p.chained) pdl_wait_primary();  // first scan of a chain: the query preparation before it must be complete
    pdl_launch_dependents();  // the next scan of the step may take SMs as they free up
    const int n_active = p.ctl->n_active;
    // no query routed to this source: do not touch the corpus; outside [min_active, max_active]: the launch of
    // the other CTA-group variant right before / after this one takes the chunk
    if (n_active == 0 || n_active < p.min_active || n_active > p.max_active) {
        pdl_wait_primary();
        return;
    }
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;  // 0 = leader (issues the MMAs)
    const int unit = (int)blockIdx.x / CG;                       // work unit = CTA or CTA pair
    const int n_units = (int)gridDim.x / CG;
    const int nqt = (n_active + kQTile * CG - 1) / (kQTile * CG);
    const int n_slices = max(1, min(n_units / nqt, p.n_tiles));
    const int qt = (unit % nqt) * CG + (int)rank;  // this CTA's 128-query tile
    const int slice = unit / nqt;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (!COLLECT) p.ctl->n_lists = n_slices;
        if (p.counters) {
            atomicAdd(&p.counters->bytes, (unsigned long long)p.n_rows * (unsigned long long)p.num_kc * 128ull);
            atomicAdd(&p.counters->flops,
                      2ull * (unsigned long long)n_active * (unsigned long long)p.n_rows * (unsigned long long)p.d_pad);
        }
    }
    if (slice >= n_slices) {  // leftover CTAs (gridDim.x is not a multiple of nqt)
        pdl_wait_primary();
        return;
    }
    const uint64_t l2_hint = (nqt > 1) ? p.l2_hint_shared : p.l2_hint;
    // corpus rows / bytes of THIS CTA per tile chunk: all 128 rows (16 KB) or, in a pair, its 64 (8 KB)
    constexpr int kRowsCta = kNT / CG;
    constexpr int kChunkB = kChunkBytes / CG;
    constexpr int kStageB = SKC * kChunkB;

    extern __shared__ __align__(1024) uint8_t smem_dyn[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) &
                                               ~(uintptr_t)1023);
    uint8_t* x_sm = base;
    uint8_t* a_sm = x_sm + p.stages * kStageB;  // [a_smem_kc][128 queries x 64 k], 128B swizzle
    float* ls = reinterpret_cast<float*>(a_sm + p.a_smem_kc * kChunkBytes);  // lists   [kDepth][128 queries]
    uint32_t* lr = reinterpret_cast<uint32_t*>(ls + kDepth * kQTile);
    float* ps = reinterpret_cast<float*>(lr + kDepth * kQTile);              // pending [kPend][128 queries]
    uint32_t* pr = reinterpret_cast<uint32_t*>(ps + kPend * kQTile);
    uint64_t* bars = reinterpret_cast<uint64_t*>(pr + kPend * kQTile);
    uint64_t* bar_full = bars;
    uint64_t* bar_empty = bars + kMaxStages;
    uint64_t* bar_acc_full = bars + 2 * kMaxStages;
    uint64_t* bar_acc_empty = bar_acc_full + kMaxAcc;
    uint64_t* bar_a = bar_acc_empty + kMaxAcc;   // query tile written to TMEM (4 epilogue warps)
    uint64_t* bar_aq = bar_a + 1;                // query tile tail landed in shared memory (TMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_aq + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
#ifdef RR_TRACE_BUILD
    const int tr = (p.trace && blockIdx.x == 0 && lane == 0) ? 1 : 0;
#endif
    trace_ev(tr && warp == 0, 0, 0);  // kernel start

    // Barrier protocol for a pair (CG = 2): the LEADER owns bar_full (armed by its producer with the
    // bytes of both CTAs; both producers' TMA complete on it), bar_acc_empty and bar_a (both CTAs'
    // epilogue warps arrive on them remotely); bar_empty and bar_acc_full exist in both CTAs and
    // are signalled in both at once by the leader's multicast tcgen05.commit.
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_x);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(smem_u32(bar_full + i), 1);
            mbar_init(smem_u32(bar_empty + i), 1);
        }
        for (int i = 0; i < p.acc_stages; ++i) {
            mbar_init(smem_u32(bar_acc_full + i), 1);
            mbar_init(smem_u32(bar_acc_empty + i), 4 * CG);
        }
        mbar_init(smem_u32(bar_a), (4 + (p.a_smem_kc > 0 ? 1 : 0)) * CG);
        mbar_init(smem_u32(bar_aq), 1);
        fence_barrier_init();
    } else if (warp == 1) {
        if (CG == 2) tmem_alloc_2cta(smem_u32(tmem_slot), kTmemCols);
        else tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all();  // the peer's barriers exist before anything is sent to them
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t a_cols = (uint32_t)p.a_tmem_kc * 32u;  // TMEM columns of the query tile
    const uint32_t acc_base = tmem_base + a_cols;

    // Warps 0 and 1 run their loops converged (all 32 lanes wait on the barriers) and elect one
    // lane for the single-thread instructions: with warp-uniform control flow the compiler keeps
    // addresses and descriptors in uniform registers -- an `if (lane == 0)` loop costs ~150
    // issue cycles per MMA (measured) and makes the issuing thread the bottleneck.
    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t x_base = smem_u32(x_sm);
        if (p.a_smem_kc > 0) {
            // this CTA's query-tile tail -> its own shared memory; then tell the leader's MMA warp
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(smem_u32(bar_aq), (uint32_t)(p.a_smem_kc * kChunkBytes));
                for (int c = 0; c < p.a_smem_kc; ++c)
                    tma_load_2d(smem_u32(a_sm + c * kChunkBytes), &tmap_q, smem_u32(bar_aq),
                                (p.a_tmem_kc + c) * RR_TC_KCHUNK, qt * kQTile, kTmaEvictLast);
            }
            __syncwarp();
            mbar_wait(smem_u32(bar_aq), 0);
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(bar_a), 0));
                else mbar_arrive(smem_u32(bar_a));
            }
        }
        // Lockstep (nqt > 1): the nqt units of a slice stream the same tiles and only the first read of a tile
        // should come from HBM.  Free running they drift apart -- a fraction of a percent in speed is tens of
        // tiles over a 25 ms launch, more than L2 holds of the 18 x nqt streams -- and every unit ends up reading
        // its own copy from HBM (ncu on the routed federation at batch 2048: 2.63x the algorithmic DRAM bytes).
        // So each leader producer publishes how many tiles it has issued and waits (bounded) while the slowest
        // unit of its slice is more than lock_window tiles behind.  The kernel ends with its slowest unit either
        // way; the peer CTA of a pair follows its leader through the shared ring.
        const bool lockstep = nqt > 1 && p.progress != nullptr && rank == 0;
        const unsigned long long* prog_sib = p.progress + slice * nqt + min(lane, nqt - 1);
        unsigned long long issued = p.progress_base;
        const unsigned long long window = (unsigned long long)p.lock_window;
        for (int tile = slice; tile < p.n_tiles; tile += n_slices) {
            unsigned long long sib = 0;
            if (lockstep) {
                if (elect_one_sync()) red_max_u64(p.progress + unit, issued);
                sib = ld_relaxed_u64(prog_sib);  // in flight under this tile's TMA issue; looked at after it
            }
            for (int kc = 0; kc < p.num_kc; kc += SKC) {
                const int nk = min(SKC, p.num_kc - kc);
                mbar_wait(smem_u32(bar_empty + stage), phase ^ 1u);
                trace_ev(tr && kc == 0, 1, tile);  // producer: first stage of the tile is free
                if (elect_one_sync()) {
                    if (CG == 2) {
                        const uint32_t full_leader = mapa_shared(smem_u32(bar_full + stage), 0);
                        if (rank == 0) mbar_arrive_expect_tx(smem_u32(bar_full + stage), (uint32_t)(nk * kChunkBytes));
                        for (int c = 0; c < nk; ++c)
                            tma_load_2d_2cta(x_base + stage * kStageB + c * kChunkB, &tmap_x, full_leader,
                                             (kc + c) * RR_TC_KCHUNK, tile * kNT + (int)rank * kRowsCta, l2_hint);
                    } else {
                        mbar_arrive_expect_tx(smem_u32(bar_full + stage), (uint32_t)(nk * kChunkBytes));
                        for (int c = 0; c < nk; ++c)
                            tma_load_2d(x_base + stage * kStageB + c * kChunkB, &tmap_x,
                                        smem_u32(bar_full + stage), (kc + c) * RR_TC_KCHUNK, tile * kNT, l2_hint);
                    }
                }
                __syncwarp();
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            ++issued;
            if (lockstep) {
#pragma unroll 1
                for (int poll = 0; poll < kLockPolls; ++poll) {
                    if (!__any_sync(RR_FULL, sib + window < issued)) break;
                    __nanosleep(400);
                    sib = ld_relaxed_u64(prog_sib);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =======================
        constexpr uint32_t idesc = umma_idesc_bf16(kQTile * CG, kNT);
        if (rank == 0) {  // only the leader of a pair issues
        mbar_wait(smem_u32(bar_a), 0);  // the query tiles are in TMEM and their tails in shared memory
        tc_fence_after();
        trace_ev(tr, 2, 0);  // MMA warp: query tile in place
        int stage = 0, acc = 0;
        uint32_t phase = 0, acc_phase = 0;
        const uint32_t x_base = smem_u32(x_sm);
        const uint32_t a_base = smem_u32(a_sm);
        for (int tile = slice; tile < p.n_tiles; tile += n_slices) {
            mbar_wait(smem_u32(bar_acc_empty + acc), acc_phase ^ 1u);
            tc_fence_after();
            trace_ev(tr, 3, tile);  // MMA warp: accumulator free
            const uint32_t d_tmem = acc_base + (uint32_t)(acc * kNT);
            for (int kc = 0; kc < p.num_kc; kc += SKC) {
                const int nk = min(SKC, p.num_kc - kc);
                mbar_wait(smem_u32(bar_full + stage), phase);
                tc_fence_after();
                trace_ev(tr && (kc == 0 || kc + SKC >= p.num_kc), 4, tile);  // MMA warp: first / last stage of the tile landed
                const uint64_t b_desc = kUmmaDescSw128Hi | (uint64_t)(((x_base + stage * kStageB) >> 4) & 0x3FFFu);
                const uint32_t a_tmem = tmem_base + (uint32_t)(kc * 32);
                const uint64_t a_desc =
                    kUmmaDescSw128Hi | (uint64_t)(((a_base + (kc - p.a_tmem_kc) * kChunkBytes) >> 4) & 0x3FFFu);
                if (elect_one_sync()) {
                    // a_tmem_kc is even whenever part of the tile lives in shared memory, so a stage is
                    // all-TMEM or all-smem
                    if (kc < p.a_tmem_kc) {  // A from tensor memory
#pragma unroll
                        for (int c = 0; c < SKC; ++c) {
                            if (c < nk) {
                                const uint32_t at = a_tmem + c * 32;
                                const uint64_t bd = b_desc + (uint64_t)(c * (kChunkB >> 4));
                                if (CG == 2) {
                                    umma_bf16_ts_2cta(d_tmem, at, bd, idesc, (uint32_t)((kc | c) != 0));
                                    umma_bf16_ts_2cta(d_tmem, at + 8, bd + 2, idesc, 1u);
                                    umma_bf16_ts_2cta(d_tmem, at + 16, bd + 4, idesc, 1u);
                                    umma_bf16_ts_2cta(d_tmem, at + 24, bd + 6, idesc, 1u);
                                } else {
                                    umma_bf16_ts(d_tmem, at, bd, idesc, (uint32_t)((kc | c) != 0));
                                    umma_bf16_ts(d_tmem, at + 8, bd + 2, idesc, 1u);
                                    umma_bf16_ts(d_tmem, at + 16, bd + 4, idesc, 1u);
                                    umma_bf16_ts(d_tmem, at + 24, bd + 6, idesc, 1u);
                                }
                            }
                        }
                    } else {  // A from shared memory
#pragma unroll
                        for (int c = 0; c < SKC; ++c) {
                            if (c < nk) {
                                const uint64_t ad = a_desc + (uint64_t)(c * (kChunkBytes >> 4));
                                const uint64_t bd = b_desc + (uint64_t)(c * (kChunkB >> 4));
                                if (CG == 2) {
                                    umma_bf16_2cta(d_tmem, ad, bd, idesc, 1u);
                                    umma_bf16_2cta(d_tmem, ad + 2, bd + 2, idesc, 1u);
                                    umma_bf16_2cta(d_tmem, ad + 4, bd + 4, idesc, 1u);
                                    umma_bf16_2cta(d_tmem, ad + 6, bd + 6, idesc, 1u);
                                } else {
                                    umma_bf16(d_tmem, ad, bd, idesc, 1u);
                                    umma_bf16(d_tmem, ad + 2, bd + 2, idesc, 1u);
                                    umma_bf16(d_tmem, ad + 4, bd + 4, idesc, 1u);
                                    umma_bf16(d_tmem, ad + 6, bd + 6, idesc, 1u);
                                }
                            }
                        }
                    }
                    if (CG == 2) {  // both CTAs' stage / accumulator barriers at once
                        umma_commit_2cta(smem_u32(bar_empty + stage));
                        if (kc + SKC >= p.num_kc) umma_commit_2cta(smem_u32(bar_acc_full + acc));
                    } else {
                        umma_commit(smem_u32(bar_empty + stage));  // frees the smem stage when the MMAs retire
                        if (kc + SKC >= p.num_kc) umma_commit(smem_u32(bar_acc_full + acc));  // accumulator complete
                    }
                }
                __syncwarp();
                if (++stage == p.stages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (++acc == p.acc_stages) {
                acc = 0;
                acc_phase ^= 1u;
            }
        }
        }  // leader
    } else {
        // ============ warps 2..5: load the query tile into TMEM, then epilogue ============
        const int quad = warp & 3;           // TMEM lanes [32*quad, 32*quad+32) are this warp's
        const int t = quad * 32 + lane;      // query row inside the tile = TMEM lane
        const int qg = qt * kQTile + t;      // query inside the chunk
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quad * 32) << 16);
        {
            // The warp's 32 query rows go into its 32 TMEM lanes, a_cols * 4 bytes each.  tcgen05.st wants row i in
            // the registers of lane i, but a lane reading its own row straight from global memory makes every load
            // instruction touch 32 different lines (9 us for the tile, RR_TRACE).  So: coalesced loads (half a warp
            // per 256-byte row piece), a transpose through shared memory (the list / pending arrays, not yet in
            // use: 32 rows x (256 + 16) bytes per warp), and each lane picks up its own row.
            constexpr int kPitch = 272;  // bytes per staged row piece: conflict-free 16-byte reads of one's own row
            static_assert(4 * 32 * kPitch <= (kDepth + kPend) * kQTile * 8, "staging aliases the list arrays");
            uint8_t* stg = reinterpret_cast<uint8_t*>(ls) + quad * (32 * kPitch);
            const uint8_t* tile0 = reinterpret_cast<const uint8_t*>(p.qbf16 + (size_t)(qt * kQTile + quad * 32) * p.d_pad);
            const size_t row_bytes = (size_t)p.d_pad * 2;
            const int half = lane >> 4, l16 = lane & 15;
            for (uint32_t c = 0; c < a_cols; c += 64) {  // 64 TMEM columns = 256 bytes of every row
                const bool two = c + 32 < a_cols;  // (a_cols is a multiple of 32: the last piece may be 128 bytes)
                uint4 v[16];
                if (two || l16 < 8) {
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        v[u] = __ldg(reinterpret_cast<const uint4*>(tile0 + (size_t)(2 * u + half) * row_bytes + 4 * c) + l16);
#pragma unroll
                    for (int u = 0; u < 16; ++u)
                        *reinterpret_cast<uint4*>(stg + (2 * u + half) * kPitch + 16 * l16) = v[u];
                }
                __syncwarp();
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (u < 4 || two) {
                        const uint4 lo4 = *reinterpret_cast<const uint4*>(stg + lane * kPitch + 32 * u);
                        const uint4 hi4 = *reinterpret_cast<const uint4*>(stg + lane * kPitch + 32 * u + 16);
                        const uint32_t r[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
                        tmem_st_x8(lane_addr + c + 8 * u, r);
                    }
                }
                __syncwarp();  // the staging rows are rewritten by the next piece
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(bar_a), 0));
                else mbar_arrive(smem_u32(bar_a));
            }
            trace_ev(tr && warp == 2, 8, 0);  // epilogue: own part of the query tile stored
        }
        // the transposing stage of ALL four warps aliases the list arrays initialised next: wait for the slowest
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int e = 0; e < DEPTH; ++e) {
            ls[e * kQTile + t] = RR_NEG_INF;
            lr[e * kQTile + t] = RR_NO_ROW;
        }
        float thr = (qg < n_active) ? (COLLECT ? p.theta[qg] : RR_NEG_INF) : __builtin_huge_valf();
        int pos = 0;  // slot of the list's minimum (an empty slot while there is one)
        int cnt = 0;  // pending survivors

        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = slice; tile < p.n_tiles; tile += n_slices) {
            mbar_wait(smem_u32(bar_acc_full + acc), acc_phase);
            tc_fence_after();
            trace_ev(tr && warp == 2, 5, tile);  // epilogue: accumulator complete
            const uint32_t t_acc = lane_addr + a_cols + (uint32_t)(acc * kNT);
            uint32_t r[4][32];
            tmem_ld_x32_nowait(t_acc, r[0]);
            tmem_ld_x32_nowait(t_acc + 32, r[1]);
            tmem_ld_x32_nowait(t_acc + 64, r[2]);
            tmem_ld_x32_nowait(t_acc + 96, r[3]);
            tmem_ld_wait();
            trace_ev(tr && warp == 2, 6, tile);  // epilogue: tile in registers
            // the registers hold the tile: hand the accumulator back before selecting
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(bar_acc_empty + acc), 0));
                else mbar_arrive(smem_u32(bar_acc_empty + acc));
            }
            if (++acc == p.acc_stages) {
                acc = 0;
                acc_phase ^= 1u;
            }
            const int64_t row0 = (int64_t)tile * kNT;
            if (row0 + kNT > p.n_rows) {  // last, partial tile: rows past the end never qualify
                const int valid = (int)(p.n_rows - row0);
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (c * 32 + j >= valid) r[c][j] = __float_as_uint(RR_NEG_INF);
            }
            if (!COLLECT && tile == slice && qg < n_active) {
                // First tile of the launch: the list is empty, so all 128 scores would "survive" and go
                // through 128 one-by-one inserts while the tensor pipe and the TMA ring wait.  Seed the
                // list with the best (DEPTH 32: two best) of each group of 128 / DEPTH * ... columns instead
                // (branch-free) and send only what beats their minimum through the regular path.
                static_assert((DEPTH == 32 || DEPTH == 8) && kNT == 128, "seeding fills 2 x 16 or 1 x 8 groups");
                float mn = __builtin_huge_valf();
                int mn_e = 0;
                if (DEPTH == 32) {
#pragma unroll
                    for (int g = 0; g < 16; ++g) {
                        float a1 = RR_NEG_INF, a2 = RR_NEG_INF;
                        int j1 = 0, j2 = 0;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float v = __uint_as_float(r[g >> 2][(g & 3) * 8 + j]);
                            if (v > a1) {
                                a2 = a1;
                                j2 = j1;
                                a1 = v;
                                j1 = j;
                            } else if (v > a2) {
                                a2 = v;
                                j2 = j;
                            }
                        }
                        ls[(2 * g) * kQTile + t] = a1;
                        lr[(2 * g) * kQTile + t] = (a1 > RR_NEG_INF) ? (uint32_t)(row0 + g * 8 + j1) : RR_NO_ROW;
                        ls[(2 * g + 1) * kQTile + t] = a2;
                        lr[(2 * g + 1) * kQTile + t] = (a2 > RR_NEG_INF) ? (uint32_t)(row0 + g * 8 + j2) : RR_NO_ROW;
                        if (a1 < mn) {
                            mn = a1;
                            mn_e = 2 * g;
                        }
                        if (a2 < mn) {
                            mn = a2;
                            mn_e = 2 * g + 1;
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j == j1 || j == j2) r[g >> 2][(g & 3) * 8 + j] = __float_as_uint(RR_NEG_INF);
                    }
                } else {
#pragma unroll
                    for (int g = 0; g < 8; ++g) {  // the best of each group of 16 columns
                        float a1 = RR_NEG_INF;
                        int j1 = 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float v = __uint_as_float(r[g >> 1][(g & 1) * 16 + j]);
                            if (v > a1) {
                                a1 = v;
                                j1 = j;
                            }
                        }
                        ls[g * kQTile + t] = a1;
                        lr[g * kQTile + t] = (a1 > RR_NEG_INF) ? (uint32_t)(row0 + g * 16 + j1) : RR_NO_ROW;
                        if (a1 < mn) {
                            mn = a1;
                            mn_e = g;
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (j == j1) r[g >> 1][(g & 1) * 16 + j] = __float_as_uint(RR_NEG_INF);
                    }
                }
                thr = mn;  // -inf (and pos = an empty slot) when the tile had fewer than DEPTH rows
                pos = mn_e;
            }
            // maxima of the 16 groups of 8 columns, then of the tile: hits are sparse (a few
            // per warp and tile), so whole groups are skipped with one vote each
            float m8[16];
#pragma unroll
            for (int g = 0; g < 16; ++g) {
                const uint32_t* rg = &r[g >> 2][(g & 3) * 8];
                const float a0 = fmaxf(__uint_as_float(rg[0]), __uint_as_float(rg[1]));
                const float a1 = fmaxf(__uint_as_float(rg[2]), __uint_as_float(rg[3]));
                const float a2 = fmaxf(__uint_as_float(rg[4]), __uint_as_float(rg[5]));
                const float a3 = fmaxf(__uint_as_float(rg[6]), __uint_as_float(rg[7]));
                m8[g] = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
            }
            float mx = m8[0];
#pragma unroll
            for (int g = 1; g < 16; ++g) mx = fmaxf(mx, m8[g]);
            if (__any_sync(RR_FULL, mx > thr)) {
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                    if (__any_sync(RR_FULL, m8[g] > thr)) {
                        // append this group's survivors to the pending queue (predicated, branch-free)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float v = __uint_as_float(r[g >> 2][(g & 3) * 8 + j]);
                            if (v > thr) {
                                ps[cnt * kQTile + t] = v;
                                pr[cnt * kQTile + t] = (uint32_t)(row0 + g * 8 + j);
                                ++cnt;
                            }
                        }
                        // a group adds at most 8 entries per thread: drain before the queue could overflow
                        if (__any_sync(RR_FULL, cnt > kPend - 8)) {
                            if (COLLECT) epi_collect(cnt, qg, p, pr, t);
                            else epi_drain<DEPTH>(cnt, thr, pos, ls, lr, ps, pr, t);
                            cnt = 0;
                        }
                    }
                }
                // one drain per tile serves every query of the warp that had survivors
                if (__any_sync(RR_FULL, cnt > 0)) {
                    if (COLLECT) epi_collect(cnt, qg, p, pr, t);
                    else epi_drain<DEPTH>(cnt, thr, pos, ls, lr, ps, pr, t);
                    cnt = 0;
                }
            }
        }
        trace_ev(tr && warp == 2, 7, 0);  // epilogue: all tiles selected
        // publish [compact query][list = slice][kDepth]: the list's minimum goes to the LAST slot (the tail
        // that bounds every dropped row; an empty slot there = nothing was dropped), the rest unsorted
        if (!COLLECT && qg < n_active) {
            const int64_t o = (int64_t)qg * p.cand_stride + (int64_t)slice * DEPTH;
            int w = 0;
            for (int e = 0; e < DEPTH; ++e) {
                if (e == pos) continue;
                p.cand_s[o + w] = ls[e * kQTile + t];
                p.cand_r[o + w] = lr[e * kQTile + t];
                ++w;
            }
            p.cand_s[o + DEPTH - 1] = ls[pos * kQTile + t];
            p.cand_r[o + DEPTH - 1] = lr[pos * kQTile + t];
        }
    }

    tc_fence_before();
    if (CG == 2) cluster_sync_all();  // neither CTA leaves while the other may still signal it
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_2cta(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
    trace_ev(tr && warp == 0, 9, 0);  // kernel end
    pdl_wait_primary();  // complete only after the previous scan of the chain has
}

}  // namespace

extern "C" int rr_debug_read_trace(unsigned long long* out, int cap) {
    unsigned int n = 0;
    RR_CUDA_CHECK(cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(n)));
    const int m = (int)std::min<unsigned>(n, (unsigned)std::min(cap, kTraceCap));
    if (m > 0) RR_CUDA_CHECK(cudaMemcpyFromSymbol(out, g_trace, (size_t)m * 16));
    const unsigned int zero = 0;
    RR_CUDA_CHECK(cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero)));
    return m;
}

bool rr_gemm_scan_supported(const rr_index* h) {
    // the 128-query tile lives in TMEM (up to 768 dims next to one accumulator stage, 512 next to two) and
    // shared memory (up to 512 dims): 1024 dimensions at most
    return h->d_pad <= kMaxDpad;
}

int rr_gemm_scan_max_lists(const rr_index* h) { return h->sm_count; }

template <int CG, bool COLLECT, int DEPTH, int SKC = kStageKc>
static int launch_cg(rr_index* h, const CUtensorMap& tx, const CUtensorMap& tq, GemmParams p, bool chain,
                     cudaStream_t stream) {
    const size_t stage_b = (size_t)SKC * kChunkBytes / CG;
    const size_t fixed = 1024 + (size_t)p.a_smem_kc * kChunkBytes + (size_t)(kDepth + kPend) * kQTile * 8 +
                         (2 * kMaxStages + 2 * kMaxAcc + 2) * 8 + 16;
    p.stages = (int)std::min<size_t>(kMaxStages, ((size_t)kSmemLimit - fixed) / stage_b);
    RR_REQUIRE(p.stages >= 2, "rr_gemm_scan: d_pad=%d leaves room for %d TMA stage(s) with %d CTA(s) per query tile",
               p.d_pad, p.stages, CG);
    const size_t smem = fixed + (size_t)p.stages * stage_b;
    static std::atomic<bool> attr_done[RR_MAX_DEVICES];  // per (instantiation, device); a double set is harmless
    if (!attr_done[h->device].load(std::memory_order_acquire)) {
        RR_CUDA_CHECK(cudaFuncSetAttribute(rr_gemm_scan_kernel<CG, COLLECT, DEPTH, SKC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           kSmemLimit));
        attr_done[h->device].store(true, std::memory_order_release);
    }
    // every SM gets a CTA; the kernel itself splits them into query tiles x corpus slices
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(h->sm_count / CG * CG));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = rr_pdl_enabled() ? 2 : 1;
    p.chained = (chain && rr_pdl_enabled()) ? 1 : 0;
    RR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, rr_gemm_scan_kernel<CG, COLLECT, DEPTH, SKC>, tx, tq, p));
    return RR_OK;
}

int rr_gemm_scan_list_depth(const rr_index* h, int nq) {
    // Short lists for small shards (see rr_gemm_scan_kernel) -- but only when the chunk is ONE query tile (<= 128
    // queries on single CTAs, <= 256 on pairs), so that every query gets a list from each of the 148 / 74 corpus
    // slices: with t query tiles the slices (and lists per query) shrink t-fold and 8-entry lists no longer hold
    // the k' = 128 candidates the finalize wants.  "gemm_depth" (testing knob) forces 8 or 32.
    if (h->gemm_depth == 8 || h->gemm_depth == 32) return h->gemm_depth;
    // (at least ~4 corpus tiles per list: with fewer rows per list than that a list's 8th-best is too close to the
    // corpus-wide k-th best and certificates start to fail -- seen at 8 GPUs on 15 k-row shards -- while the
    // cold-start cost that short lists avoid is small there anyway)
    const bool small_shard = h->n <= (int64_t)h->sm_count * 16384 && h->n >= (int64_t)h->sm_count * 4 * kNT;
    return (small_shard && nq <= 2 * kQTile) ? 8 : kDepth;
}

static int rr_gemm_scan_launch_one(rr_index* h, const uint16_t* qbf16_dev, int nq, int cta_group, rr_chunk_ctl* ctl_dev,
                                   float* cand_s, uint32_t* cand_r, int64_t cand_stride, rr_scan_counters* counters_dev,
                                   const rr_collect_args* collect, int min_active, int max_active, bool chain,
                                   cudaStream_t stream);

int rr_gemm_scan_launch(rr_index* h, const uint16_t* qbf16_dev, int nq, int cta_group, rr_chunk_ctl* ctl_dev,
                        float* cand_s, uint32_t* cand_r, int64_t cand_stride, rr_scan_counters* counters_dev,
                        const rr_collect_args* collect, bool chain, cudaStream_t stream) {
    RR_REQUIRE(rr_gemm_scan_supported(h), "rr_gemm_scan_launch: d_pad=%d leaves no TMEM for accumulators", h->d_pad);
    RR_REQUIRE(nq >= 1 && nq <= RR_GEMM_QCHUNK, "rr_gemm_scan_launch: nq=%d out of range", nq);
    RR_REQUIRE(cta_group >= 0 && cta_group <= 2, "rr_gemm_scan_launch: cta_group must be 0 (decide on the device), 1 or 2");
    if (cta_group == 0) {
        // Two launches, of which the device runs one: up to 128 ROUTED queries one CTA per corpus slice
        // (a single query tile: HBM-bound, 148 slices), more than that CTA pairs.  The host only knows the
        // batch size, not how many of its queries the router sent to this source.
        // (the pair variant may take over below a full single tile: "gemm_pair_min" routed queries, <= 129)
        static const char* pair_min_env = getenv("RR_PAIR_MIN");  // A/B switch for measurements
        const int pair_min = std::min(kQTile + 1, std::max(2, pair_min_env ? atoi(pair_min_env) : h->gemm_pair_min));
        int rc = rr_gemm_scan_launch_one(h, qbf16_dev, nq, 1, ctl_dev, cand_s, cand_r, cand_stride, counters_dev, collect,
                                         1, pair_min - 1, chain, stream);
        if (rc != RR_OK) return rc;
        return rr_gemm_scan_launch_one(h, qbf16_dev, nq, 2, ctl_dev, cand_s, cand_r, cand_stride, counters_dev, collect,
                                       pair_min, 0x7FFFFFFF, true, stream);
    }
    return rr_gemm_scan_launch_one(h, qbf16_dev, nq, cta_group, ctl_dev, cand_s, cand_r, cand_stride, counters_dev, collect,
                                   1, 0x7FFFFFFF, chain, stream);
}

static int rr_gemm_scan_launch_one(rr_index* h, const uint16_t* qbf16_dev, int nq, int cta_group, rr_chunk_ctl* ctl_dev,
                                   float* cand_s, uint32_t* cand_r, int64_t cand_stride, rr_scan_counters* counters_dev,
                                   const rr_collect_args* collect, int min_active, int max_active, bool chain,
                                   cudaStream_t stream) {
    // corpus tensor maps: 128-row boxes (shared with the small-batch kernel) or, for CTA pairs, 64-row boxes
    if (h->tmap_hi_ptr != h->hi || h->tmap_hi_rows != h->n) {
        int rc = rr_make_tmap_bf16_rows(&h->tmap_hi, h->hi, h->n, h->d_pad, kNT);
        if (rc != RR_OK) return rc;
        h->tmap_hi_ptr = h->hi;
        h->tmap_hi_rows = h->n;
    }
    if (cta_group == 2 && (h->tmap_hi64_ptr != h->hi || h->tmap_hi64_rows != h->n)) {
        int rc = rr_make_tmap_bf16_rows(&h->tmap_hi64, h->hi, h->n, h->d_pad, kNT / 2);
        if (rc != RR_OK) return rc;
        h->tmap_hi64_ptr = h->hi;
        h->tmap_hi64_rows = h->n;
    }
    GemmParams p;
    p.n_rows = h->n;
    p.n_tiles = (int)((h->n + kNT - 1) / kNT);
    p.num_kc = h->d_pad / RR_TC_KCHUNK;
    p.d_pad = h->d_pad;
    // Where the stationary 128-query tile lives.  Default: 512 dims in TMEM + the rest (<= 512 dims) in shared
    // memory, which leaves 256 TMEM columns = two accumulator stages.
    p.acc_stages = kMaxAcc;
    p.a_tmem_kc = std::min(p.num_kc, kMaxATmemKc);
    p.a_smem_kc = p.num_kc - p.a_tmem_kc;
    const bool one_tile = cta_group == 1 && (nq <= kQTile || max_active <= kQTile);
    constexpr int kTmemKcOneAcc = ((int)kTmemCols - kNT) / 32;  // 12 K-chunks = 768 dims next to ONE accumulator
    if (one_tile && (h->gemm_tmem_all || p.num_kc > kMaxATmemKc + 4)) {
        // one query tile = the HBM-bound regime (intensity <= 128 flop/byte): the tensor pipe has slack, so
        // up to 768 dims go to TMEM with ONE accumulator stage (the MMAs of the next corpus tile wait for the
        // epilogue's tcgen05.ld, a fraction of the tile's HBM time) and the shared memory that frees becomes
        // TMA stages in flight
        p.a_tmem_kc = std::min(p.num_kc, kTmemKcOneAcc);
        p.a_smem_kc = p.num_kc - p.a_tmem_kc;
        p.acc_stages = 1;
    }
    p.stages = 0;
    // sibling units (other query tiles, same slice) re-read each tile from L2 right away
    p.l2_hint = kTmaEvictFirst;
    p.l2_hint_shared = kTmaEvictNormal;
    p.qbf16 = qbf16_dev;
    p.ctl = ctl_dev;
    p.counters = counters_dev;
    p.cand_s = cand_s;
    p.cand_r = cand_r;
    p.cand_stride = cand_stride;
    p.min_active = min_active;
    p.max_active = max_active;
    static const char* lockstep_env = getenv("RR_LOCKSTEP");  // A/B switch for measurements: window in tiles, 0 = off
    const int window = lockstep_env ? atoi(lockstep_env) : h->lockstep;
    const bool lockstep = window > 0 && h->scan_progress != nullptr && h->sm_count <= RR_MAX_SCAN_UNITS;
    p.progress = lockstep ? h->scan_progress : nullptr;
    p.lock_window = window;
    p.progress_base = (unsigned long long)(++h->scan_seq) << 32;
    static const int trace_on = getenv("RR_TRACE") != nullptr;
    p.trace = trace_on;
    p.theta = collect ? collect->theta : nullptr;
    p.col_rows = collect ? collect->rows : nullptr;
    p.col_cnt = collect ? collect->cnt : nullptr;
    p.col_cap = collect ? collect->cap : 0;
    RR_REQUIRE(collect || (int64_t)h->sm_count * kDepth <= cand_stride, "rr_gemm_scan_launch: candidate stride too small");
    // the tail of each query tile is fetched by TMA from the bf16 query buffer
    // (the caller pads the buffer to 256 rows whenever the pair variant may run)
    const int q_align = kQTile * cta_group;
    const int nq_rows = (nq + q_align - 1) / q_align * q_align;
    alignas(64) CUtensorMap tq;
    int rc = rr_make_tmap_bf16_rows(&tq, qbf16_dev, nq_rows, h->d_pad, kQTile);
    if (rc != RR_OK) return rc;
    if (collect)
        return cta_group == 2 ? launch_cg<2, true, kDepth>(h, h->tmap_hi64, tq, p, chain, stream)
                              : launch_cg<1, true, kDepth>(h, h->tmap_hi, tq, p, chain, stream);
    if (rr_gemm_scan_list_depth(h, nq) == 8)
        return cta_group == 2 ? launch_cg<2, false, 8>(h, h->tmap_hi64, tq, p, chain, stream)
                              : launch_cg<1, false, 8>(h, h->tmap_hi, tq, p, chain, stream);
    // 64 KB stages (one barrier round trip and one commit per 16 MMAs instead of 8) for CTA pairs on batches that
    // can hold several query tiles -- the tensor-bound regime -- when at least three such stages fit beside the
    // query tile's tail (d_pad <= 768) and the tile's TMEM / shared-memory split falls on a stage boundary
    static const char* skc_env = getenv("RR_STAGE_KC");  // A/B switch for measurements
    const int skc = skc_env ? atoi(skc_env) : h->gemm_stage_kc;
    static const char* skc_minq_env = getenv("RR_STAGE_KC_MINQ");  // (measurements: smallest batch that takes them)
    const int skc_min_nq = skc_minq_env ? atoi(skc_minq_env) : 2 * kQTile + 1;
    if (cta_group == 2 && skc == 4 && nq >= skc_min_nq && (p.a_smem_kc == 0 || p.a_tmem_kc % 4 == 0)) {
        const size_t fixed = 1024 + (size_t)p.a_smem_kc * kChunkBytes + (size_t)(kDepth + kPend) * kQTile * 8 +
                             (2 * kMaxStages + 2 * kMaxAcc + 2) * 8 + 16;
        if (fixed + 3 * (size_t)(4 * kChunkBytes / 2) <= (size_t)kSmemLimit)
            return launch_cg<2, false, kDepth, 4>(h, h->tmap_hi64, tq, p, chain, stream);
    }
    return cta_group == 2 ? launch_cg<2, false, kDepth>(h, h->tmap_hi64, tq, p, chain, stream)
                          : launch_cg<1, false, kDepth>(h, h->tmap_hi, tq, p, chain, stream);
}
