// Internal layout of an index handle, shared by rr_index.cu and rr_scan_tc.cu.
#pragma once
#include "rr_common.cuh"

// Candidate lists produced by a scan kernel for one chunk of queries:
//   cand_s / cand_r : [n_queries][n_lists][depth]   (score fp32, local row u32)
// Each list is sorted (score desc, row asc); empty slots are (-inf, RR_NO_ROW).
// A list whose last slot is occupied may have dropped rows, all of which score
// <= that last slot.

constexpr int RR_TC_TILE_ROWS = 128;   // corpus rows per MMA tile (UMMA M)
constexpr int RR_TC_KCHUNK = 64;       // bf16 elements per 128-byte swizzle row
constexpr int RR_TC_MAX_QTILE = 64;    // queries per scan pass (UMMA N <= 64)
constexpr int RR_TC_LIST_DEPTH = 16;   // entries per (epilogue warp, query) list
constexpr int RR_TC_EPI_WARPS = 4;
constexpr int RR_EXACT_LIST_DEPTH = 32;
constexpr int RR_EXACT_WARPS = 8;
constexpr int RR_QCHUNK = 256;         // queries handled per internal chunk (64-query scan passes)
constexpr int RR_GEMM_TILE_ROWS = 128; // corpus rows per MMA in the large-batch kernel (UMMA N)
constexpr int RR_GEMM_QTILE = 128;     // queries per CTA in the large-batch kernel (UMMA M, TMEM lanes)
constexpr int RR_GEMM_LIST_DEPTH = 32; // entries per (CTA, query) list of the large-batch kernel
constexpr int RR_GEMM_QCHUNK = 2048;   // queries per chunk on the large-batch path
constexpr int RR_PATH_AUTO = 0, RR_PATH_SMALL = 1, RR_PATH_GEMM = 2, RR_PATH_GEMM_PAIR = 3;  // set_param "scan_path"
constexpr int RR_EXACT_MAX_QSUB = 16;  // queries per exact-scan launch
constexpr int RR_MAX_SCAN_UNITS = 256; // work units (CTAs or CTA pairs) of a queries-in-TMEM scan launch

struct rr_index {
    int d = 0;
    int d_pad = 0;  // d rounded up to a multiple of 64 (zero padded)
    int device = 0;
    int sm_count = 0;
    int64_t n = 0;
    int64_t cap = 0;
    int64_t id_offset = 0;
    uint16_t* hi = nullptr;  // [cap, d_pad] bf16 plane (MMA operand)
    uint16_t* lo = nullptr;  // [cap, d_pad] remainder plane
    uint32_t* max_norm_bits = nullptr;  // device, float bits rounded up: [0] max row L2 norm, [1] max L2 norm of a
                                        // row's bf16 rounding error |x - bf16(x)|
    cudaStream_t host_stream = nullptr;  // rr_index_search_host runs here (non-blocking)
    cudaEvent_t add_event = nullptr;     // recorded after an asynchronous rr_index_add_f32
    bool add_event_pending = false;
    bool dev_dirty = false;              // a device-API search was issued since the last host search

    // parameters
    int mode = RR_MODE_AUTO;
    int kprime = 128;
    int profile = 0;
    int scan_path = RR_PATH_AUTO;
    int gemm_depth = 0;     // 0 = automatic list depth of the queries-in-TMEM scan (8 for small shards, else 32)
    int gemm_tmem_all = 1;  // <= 128 queries: whole query tile in TMEM, one accumulator stage (testing knob)
    int gemm_stage_kc = 4;  // queries-in-TMEM scan on CTA pairs, batches > 256: K-chunks per TMA stage (4 = 64 KB stages; 2)
    int gemm_pair_min = RR_GEMM_QTILE + 1;  // batches > 128: routed queries from which the device picks CTA pairs
    int lockstep = 4;       // queries-in-TMEM scan: units sharing a corpus slice stay within this many tiles of each other (0 = off)
    unsigned long long* scan_progress = nullptr;  // device, [RR_MAX_SCAN_UNITS]: (launch number << 32) | tiles issued
    uint32_t scan_seq = 0;                        // launch number of the last queries-in-TMEM scan

    // per-search scratch (grow-only)
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    float* stage = nullptr;  // host->device staging for add / search_host
    size_t stage_bytes = 0;
    void* pin = nullptr;     // pinned host mirror of search_host's result block
    size_t pin_bytes = 0;

    // TMA descriptor of the hi plane (rebuilt when hi / n change)
    alignas(64) CUtensorMap tmap_hi;
    const void* tmap_hi_ptr = nullptr;
    int64_t tmap_hi_rows = -1;
    alignas(64) CUtensorMap tmap_hi64;  // same plane, 64-row boxes (CTA pairs of the large-batch kernel)
    const void* tmap_hi64_ptr = nullptr;
    int64_t tmap_hi64_rows = -1;

    // stats
    rr_index_stats stats = {};
    int* d_nfail_total = nullptr;  // device counters (32 bytes): [0] certificate failures (int);
                                   // bytes 8..15 scanned bytes (u64); bytes 16..23 scored flops (u64)
    static constexpr int kEvPool = 4096;  // 2048 timed launches between two rr_index_get_stats calls
    cudaEvent_t ev_pool[kEvPool] = {};  // pairs: [2i] before, [2i+1] after a chain of scan launches
    int ev_launches[kEvPool / 2] = {};  // scan launches each pair brackets
    int ev_used = 0;
};

// Control block of one query chunk (device memory, zeroed per chunk).
struct rr_chunk_ctl {
    int n_fail;      // queries needing the exact path
    int any_active;  // at least one query is routed to this source
    int n_active;    // large-batch path: routed queries (they are compacted to rows [0, n_active))
    int n_lists;     // large-batch path: candidate lists per query the scan produced
    int n_fail2;     // queries the second pass could not settle (row list overflow): exact scan
    int pad_[3];
};
// second pass ("collect"): per-query fixed thresholds in, per-query row lists out (all device memory,
// indexed by the compact query order of that pass)
struct rr_collect_args {
    const float* theta;  // rows with bf16 score > theta[q] are reported
    uint32_t* rows;      // [q][cap]
    int* cnt;            // [q], zeroed by the caller; may end above cap (overflow)
    int cap;
};
// profiling counters handed to the scan kernels (or NULL): [0] bytes streamed, [1] flops scored
struct rr_scan_counters {
    unsigned long long bytes;
    unsigned long long flops;
};

// rr_scan_tc.cu: launch the tcgen05 scan for one pass of <= qtile queries: rows
// [q_row0, q_row0 + qtile) of the bf16 query buffer qbf16 [q_rows_total, d_pad]
// (zero padded), of which the first nq_tile are real.  active [nq_tile] uint8:
// 0 = query not routed to this source.  Fills cand_s / cand_r (already offset to
// the pass's first query) and reports the number of lists per query.
int rr_tc_scan_launch(rr_index* h, const uint16_t* qbf16_dev, int64_t q_rows_total, int q_row0, int nq_tile,
                      const uint8_t* active_dev, const int* any_active_dev, float* cand_s, uint32_t* cand_r,
                      int64_t cand_stride, int* n_lists_out, rr_scan_counters* counters_dev, bool chain,
                      cudaStream_t stream);
// number of candidate lists per query the tc scan produces for this index
int rr_tc_scan_num_lists(const rr_index* h);
// queries per pass for this dimension (stationary query tile must fit in smem)
int rr_tc_scan_qtile(const rr_index* h);

// rr_tma_host.cu: 2-D tensor map over a row-major bf16 matrix [rows, d_pad]; box = 64 columns
// (one 128-byte swizzle row) x box_rows rows, SWIZZLE_128B.
int rr_make_tmap_bf16_rows(CUtensorMap* out, const void* ptr, int64_t rows, int d_pad, int box_rows);

// rr_scan_gemm.cu: large-batch scan (queries stationary in TMEM).  qbf16 holds
// ceil(nq / 128) * 128 rows: the ctl->n_active ROUTED queries compacted first, zero rows after
// them.  The kernel sizes its query tiles / corpus slices from ctl->n_active on the device,
// writes one list of RR_GEMM_LIST_DEPTH entries per (corpus slice, compact query) and stores
// the number of slices in ctl->n_lists.
bool rr_gemm_scan_supported(const rr_index* h);
int rr_gemm_scan_max_lists(const rr_index* h);
int rr_gemm_scan_list_depth(const rr_index* h, int nq);  // entries per (corpus slice, query) list of a chunk of nq queries: 8 or 32
// cta_group: 1 = one CTA per 128-query tile; 2 = CTA pairs (cluster of 2, 256 queries per pair),
// for which qbf16 holds ceil(nq / 256) * 256 rows.
int rr_gemm_scan_launch(rr_index* h, const uint16_t* qbf16_dev, int nq, int cta_group, rr_chunk_ctl* ctl_dev,
                        float* cand_s, uint32_t* cand_r, int64_t cand_stride, rr_scan_counters* counters_dev,
                        const rr_collect_args* collect, bool chain, cudaStream_t stream);
// `chain` (both scan launchers): this launch directly follows another scan launch on `stream`; it is
// launched with programmatic stream serialization, i.e. its CTAs may start while the previous scan's last
// CTAs are still running (the scans of one step touch disjoint memory), and it completes after it.
