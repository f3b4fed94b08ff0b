/*
 * ragroute_b200.h -- C ABI of the B200-native RAGRoute retrieval hot path.
 *
 * One shared library (ragroute_b200/csrc/libragroute_b200.so, built by nvcc for
 * sm_100a) sits behind the three arithmetic call sites of sacs-epfl/ragroute:
 *
 *   B1  per-source exact inner-product top-k   faiss IndexFlatIP::search, called at
 *       ragroute/data_source.py:158, :186, :203 (index built by faiss.read_index,
 *       data_source.py:71; faiss.normalize_L2 at :199)
 *   B2  router relevance MLP over (query, source) pairs
 *       ragroute/router.py:37-55 (CorpusRoutingNN) and :241-283
 *       (select_relevant_sources_ragroute)
 *   B3  cross-source top-k merge               ragroute/rerank.py:3-9, :28-34,
 *       called at ragroute/http_server.py:289, :293
 *
 * Conventions
 *   - plain pointers and sizes only; every *_dev pointer is device memory of the
 *     handle's device; `stream` is a cudaStream_t passed as void* (NULL = legacy
 *     default stream).  Calls are asynchronous on `stream` unless stated.
 *   - every function returns 0 on success or a negative rr_status; the message
 *     of the last failure on the calling thread is rr_last_error().
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with RR_ERR_CUDA.
 *   - a handle is not re-entrant (the reference drives each index / router from
 *     one single-threaded asyncio loop: data_source.py:99-132, router.py:207-219);
 *     different handles may be used from different threads or processes.
 */
#ifndef RAGROUTE_B200_H_
#define RAGROUTE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RR_ABI_VERSION 1

typedef enum rr_status {
    RR_OK = 0,
    RR_ERR_INVALID = -1,     /* bad argument (faiss would assert / throw)       */
    RR_ERR_CUDA = -2,        /* CUDA runtime / driver failure, or no device     */
    RR_ERR_OOM = -3,         /* device allocation failed                        */
    RR_ERR_UNSUPPORTED = -4, /* valid request this build does not implement     */
    RR_ERR_ARCH = -5         /* device is not sm_100 (tcgen05 / TMA required)   */
} rr_status;

typedef struct rr_index rr_index_t;   /* opaque: one corpus (or one row shard)  */
typedef struct rr_router rr_router_t; /* opaque: one CorpusRoutingNN + tables   */

/* ------------------------------------------------------------------ misc -- */
int rr_abi_version(void);
const char* rr_last_error(void);
/* number of visible CUDA devices, or a negative rr_status */
int rr_device_count(void);
/* kernels this library has launched in this process (benchmark bookkeeping) */
int64_t rr_kernel_launches(void);

/* ---------------------------------------------------- B1: flat IP index -- */
/* Search modes (rr_index_set_param "mode") */
#define RR_MODE_AUTO 0  /* bf16 tcgen05 scan + fp32-exact rescoring + certificate; a
                           query whose certificate fails is re-done exactly (second
                           tensor-core pass collecting every row above a provable
                           threshold, exact scan if that overflows); results
                           identical to RR_MODE_EXACT                           */
#define RR_MODE_EXACT 1 /* exact scan only (fp32 rows, fp64 accumulation)       */

/* faiss.IndexFlatIP(d) / faiss.read_index (data_source.py:71).
 * capacity_rows is a reservation hint (0 = grow on demand). */
int rr_index_create(int d, int64_t capacity_rows, int device, rr_index_t** out);
int rr_index_destroy(rr_index_t* h);
int rr_index_reserve(rr_index_t* h, int64_t capacity_rows);
/* index.add(x): append n fp32 rows [n, d] row-major.  x_on_device != 0 means x is
 * a device pointer (consumed asynchronously on `stream`); otherwise x is host
 * memory and the call returns after the rows are on the device. */
int rr_index_add_f32(rr_index_t* h, const float* x, int64_t n, int x_on_device, void* stream);
int rr_index_reset(rr_index_t* h);
int64_t rr_index_ntotal(const rr_index_t* h);
int rr_index_dim(const rr_index_t* h);
/* index.reconstruct_n(row0, n): exact fp32 rows back out of the stored planes */
int rr_index_reconstruct(rr_index_t* h, int64_t row0, int64_t n, float* out_dev, void* stream);
/* Column mean of the stored rows (fp64 sums over the exact fp32 values, rounded once): the per-source
 * centroid the router reads from <corpus>_stats.json["centroid"] (router.py:136-151; the reference computes
 * it offline).  out_host [d] fp32; synchronous. */
int rr_index_centroid(rr_index_t* h, float* out_host);
/* ids returned by search are local row numbers + offset (row-sharded corpora) */
int rr_index_set_id_offset(rr_index_t* h, int64_t offset);
/* "mode" (RR_MODE_*), "kprime" (candidates rescored per query, default 128),
 * "profile" (1 = record CUDA events around the scan kernels and count the bytes /
 * flops they cover), "scan_path" (testing: 0 = automatic dispatch, 1 = swap-AB scan in
 * passes of <= 64 queries, 2 = queries-in-TMEM scan with one CTA per 128-query tile,
 * 3 = the same on CTA pairs), "gemm_tmem_all" (testing: 0 = keep the two-accumulator layout of the
 * queries-in-TMEM scan even for <= 128 queries), "gemm_depth" (testing: entries per candidate list of the
 * queries-in-TMEM scan, 8 or 32; 0 = automatic: 8 for shards of <= 16 k rows per list), "lockstep"
 * (default 4: the work units of a queries-in-TMEM scan that stream the same corpus slice for different
 * query tiles stay within that many tiles of each other, so every tile comes from HBM once; 0 = free running),
 * "gemm_stage_kc" (4, default: CTA pairs on batches above 256 queries stream 64 KB TMA stages = 16 MMAs per
 * barrier round trip; 2 = 32 KB stages everywhere), "gemm_pair_min" (129, default: batches above 128 queries
 * take CTA pairs from that many ROUTED queries on, single CTAs below; 2 ... 129) */
int rr_index_set_param(rr_index_t* h, const char* name, double value);

/* index.search(x, k) (data_source.py:158,186,203) on device buffers:
 *   q_dev [nq, d] fp32 row-major; D_dev [nq, k] fp32; I_dev [nq, k] int64.
 * Rows of D are descending; unused slots are (-FLT_MAX, -1) exactly as faiss pads.
 * k <= RR_MAX_K. */
#define RR_MAX_K 32
int rr_index_search(rr_index_t* h, const float* q_dev, int64_t nq, int k, float* D_dev,
                    int64_t* I_dev, void* stream);
/* Routed variant: query b is searched only if mask_dev[b * mask_stride] != 0
 * (column s of the router's [B, S] selection mask with mask_stride = S);
 * unselected queries get all-padding rows.  If no query is selected the corpus
 * is not read at all. */
int rr_index_search_masked(rr_index_t* h, const float* q_dev, int64_t nq, int k,
                           const uint8_t* mask_dev, int64_t mask_stride, float* D_dev,
                           int64_t* I_dev, void* stream);
/* Same, with an explicit output row pitch: query b writes D_dev[b * out_stride + i],
 * I_dev[b * out_stride + i], i < k (out_stride >= k; 0 means k).  Lets S sources
 * write their lists side by side into one [B, S*k] buffer that rr_merge_topk reads. */
int rr_index_search_ex(rr_index_t* h, const float* q_dev, int64_t nq, int k, const uint8_t* mask_dev,
                       int64_t mask_stride, float* D_dev, int64_t* I_dev, int64_t out_stride,
                       void* stream);
/* One routed step over n_sources indexes at once: the fan-out of one query batch to the selected
 * data sources (ragroute/http_server.py:200-209) followed by each source's index.search
 * (data_source.py:186), with the per-source lists written side by side for the merge
 * (http_server.py:280-289).  Same results as n_sources rr_index_search_ex calls; the scans stay one
 * launch per source, every other stage runs as one launch for all sources (fixed cost per step
 * instead of per source).
 *   idx [n_sources] handles on ONE device, each listed once;
 *   q_dev [n_sources] device pointers, q_dev[s] = [nq, dim(idx[s])] fp32 (sources sharing an
 *   encoder pass the same pointer);
 *   mask_dev [nq, mask_stride] uint8 or NULL (all routed): column s gates source s;
 *   D_dev / I_dev [nq, out_stride]: source s writes columns [s*k, (s+1)*k); out_stride 0 = n_sources*k. */
int rr_federation_search(rr_index_t* const* idx, int n_sources, const float* const* q_dev, int64_t nq,
                         int k, const uint8_t* mask_dev, int64_t mask_stride, float* D_dev,
                         int64_t* I_dev, int64_t out_stride, void* stream);
/* Host-buffer convenience used by the faiss-shaped Python object: copies q in,
 * searches, copies D / I out and synchronises (this is the reference's calling
 * convention: numpy in, numpy out). */
int rr_index_search_host(rr_index_t* h, const float* q_host, int64_t nq, int k, float* D_host,
                         int64_t* I_host);
/* The same with flags: RR_SEARCH_NORMALIZE_L2 = faiss.normalize_L2(x) followed by index.search(x, k)
 * (retrieve_docs_wikipedia, data_source.py:199-203) as ONE host round trip: the queries are L2-normalised
 * on the device before the scan.  q_out_host (or NULL) receives the queries as searched (normalised),
 * which is what the reference's in-place normalize_L2 leaves in its array.  Both host calls run on a
 * stream owned by the handle (not the legacy default stream). */
#define RR_SEARCH_NORMALIZE_L2 1
int rr_index_search_host_ex(rr_index_t* h, const float* q_host, int64_t nq, int k, int flags, float* D_host,
                            int64_t* I_host, float* q_out_host);

typedef struct rr_index_stats {
    int64_t searches;          /* search calls since creation                    */
    int64_t queries;           /* queries since creation                         */
    int64_t exact_rescans;     /* queries whose certificate failed (settled exactly by the second,
                                  thresholded tensor-core pass or, failing that, the exact scan) */
    int64_t scan_launches;     /* bf16 scan kernel launches that were timed      */
    double scan_ms_total;      /* their summed CUDA-event time (profile = 1)     */
    int64_t scan_bytes_total;  /* algorithmic bytes they covered: rows * d_pad*2 */
    int64_t scan_flops_total;  /* algorithmic flops: 2 * routed queries * rows * d_pad */
    int64_t exact_scans;       /* of exact_rescans: queries that needed the fp64 exact scan */
} rr_index_stats;
/* Synchronises the recorded events; reset != 0 clears the timing fields and the certificate-failure
 * counters (exact_rescans, exact_scans) after reading them; searches / queries stay cumulative. */
int rr_index_get_stats(rr_index_t* h, rr_index_stats* out, int reset);

/* Testing hook: run ONLY the bf16 tcgen05 scan for nq <= rr_index_tc_qtile(h)
 * queries and copy out its raw candidate lists, cand_s_dev / cand_r_dev
 * [nq, entries_per_query] (score fp32, local row u32; 0xFFFFFFFF = empty slot),
 * entries_per_query = n_lists * depth (written to the three out ints).
 * capacity = entries per query the caller's buffers can hold. */
int rr_index_tc_qtile(const rr_index_t* h);
int rr_index_debug_tc_scan(rr_index_t* h, const float* q_dev, int nq, float* cand_s_dev,
                           uint32_t* cand_r_dev, int64_t capacity, int* n_lists, int* depth,
                           void* stream);

/* faiss.read_index(path) (data_source.py:71) for a flat inner-product index file (fourcc "IxFI" / "IxFl",
 * metric 0): creates the index on `device` and streams the fp32 rows disk -> pinned host -> device with
 * double buffering (the file never has to fit in host memory).  Any other index type: RR_ERR_INVALID.
 * rr_index_write is the inverse (faiss.write_index layout, "IxFI"). */
int rr_index_read(const char* path, int device, rr_index_t** out);
int rr_index_write(rr_index_t* h, const char* path);

/* faiss.normalize_L2(x) (data_source.py:199), in place on device rows [n, d] */
int rr_normalize_l2(float* x_dev, int64_t n, int d, void* stream);

/* ------------------------------------------------------- B2: router MLP -- */
/* CorpusRoutingNN(in_dim) + load_state_dict (router.py:117-119).  All weight
 * pointers are HOST fp32 arrays laid out as torch stores them:
 *   fc1_w [h1, in_dim], fc1_b [h1], ln1_w/ln1_b [h1],
 *   fc2_w [h2, h1],     fc2_b [h2], ln2_w/ln2_b [h2], fc3_w [h2], fc3_b [1]. */
int rr_router_create(int in_dim, int h1, int h2, const float* fc1_w, const float* fc1_b,
                     const float* ln1_w, const float* ln1_b, const float* fc2_w,
                     const float* fc2_b, const float* ln2_w, const float* ln2_b,
                     const float* fc3_w, const float* fc3_b, int device, rr_router_t** out);
int rr_router_destroy(rr_router_t* h);
/* self.router(input_tensor) (router.py:274): x_dev [P, in_dim] fp32 -> logits [P] */
int rr_router_logits(rr_router_t* h, const float* x_dev, int64_t P, float* logits_dev, void* stream);
/* Tables for the fused select (router.py:241-270), all HOST pointers:
 *   n_sources S; d_max = EMBEDDING_MAX_LENGTH; n_slots query encoders;
 *   slot_of_source [S] (which encoder's embedding source s consumes);
 *   centroids [S, d_max] fp32 (zero padded, router.py:149-151);
 *   onehot_id [S] and n_onehot (in_dim == 2*d_max + n_onehot);
 *   scaler_mean / scaler_scale [in_dim] float64 or NULL (feb4rag has no scaler). */
int rr_router_set_tables(rr_router_t* h, int n_sources, int d_max, int n_slots,
                         const int32_t* slot_of_source, const float* centroids,
                         const int32_t* onehot_id, int n_onehot, const double* scaler_mean,
                         const double* scaler_scale);
/* select_relevant_sources_ragroute for B queries at once:
 *   q_dev [B, n_slots, d_max] fp32 (each embedding zero padded to d_max, router.py:248)
 *   mask_dev [B, S] uint8 = sigmoid(logit) > prob_threshold (strict, router.py:278-280)
 *   logits_dev [B, S] fp32 or NULL. */
int rr_router_select(rr_router_t* h, const float* q_dev, int64_t B, float prob_threshold,
                     uint8_t* mask_dev, float* logits_dev, void* stream);

/* ------------------------------------------------------------ B3: merge -- */
/* rerank_medrag / rerank_wikipedia (rerank.py:3-9, :28-34) for B queries:
 *   s_dev [B, L] fp32 scores, id_dev [B, L] int64 ids (id < 0 = empty slot);
 *   out: os [B, k], oid [B, k], opos [B, k] (input position of each winner, may be
 *   NULL).  descending != 0 -> best = largest (medrag); 0 -> smallest first
 *   (wikipedia, as written).  Equal scores: lower input position first.
 *   Unused output slots: (-/+FLT_MAX, -1, -1).  L <= RR_MERGE_MAX_L. */
#define RR_MERGE_MAX_L 4096
int rr_merge_topk(const float* s_dev, const int64_t* id_dev, int64_t B, int L, int k,
                  int descending, float* os_dev, int64_t* oid_dev, int32_t* opos_dev, void* stream);

/* The same merge over lists that arrive as n_seg separate [B, seg_len] blocks -- the all-gathered
 * per-rank lists of a row-sharded federation (one block per rank, seg_stride_* elements apart), read in
 * place -- with the winner's origin decoded on the device.  Entry j of block g has input position
 * g * seg_len + j, so the result equals rr_merge_topk over the concatenation [B, n_seg * seg_len].
 *   list_len > 0: every block holds seg_len / list_len per-source lists of list_len entries (what
 *   rr_federation_search writes); then osrc [B, k] = list (source) index inside the block and
 *   orank [B, k] = block (rank) index of each winner, -1 in unused slots, and EQUAL scores are ordered
 *   by (source index, id, position) instead of position alone -- independent of how rows were sharded
 *   over the blocks, so the sharded merge equals the unsharded one bit for bit even with ties.
 *   Any of opos / osrc / orank may be NULL.  n_seg * seg_len <= RR_MERGE_MAX_L. */
int rr_merge_topk_segments(const float* s_dev, const int64_t* id_dev, int64_t B, int n_seg, int seg_len,
                           int64_t seg_stride_s, int64_t seg_stride_id, int k, int descending,
                           int list_len, float* os_dev, int64_t* oid_dev, int32_t* opos_dev,
                           int32_t* osrc_dev, int32_t* orank_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RAGROUTE_B200_H_ */
