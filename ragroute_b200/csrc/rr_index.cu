// B1: flat inner-product index (drop-in for faiss IndexFlatIP as the reference uses
// it: /root/reference/ragroute/data_source.py:71,158,186,203).
//
// Storage: two 16-bit planes per corpus (see rr_common.cuh) = 4 bytes / element,
// the same footprint as FAISS' fp32 rows, of which a scan streams only 2.
//
// Search pipeline for a chunk of queries against S sources (all on one stream, no host sync; every
// stage but the scans is ONE launch for all sources -- rr_federation_search; index.search is S = 1):
//   begin     zero the control blocks; (large batch) compact the routed queries to the front, so
//             unrouted pairs cost nothing
//   prep      fp32 queries -> bf16 rows (+ L2 norms for the certificate)
//   scan      per source, chained by programmatic dependent launch.  <= 24 queries: rr_scan_tc.cu
//             (swap-AB, per-warp top-16 lists); more: rr_scan_gemm.cu (queries stationary in TMEM,
//             per-query top-32 lists; single CTAs up to 128 ROUTED queries, CTA pairs above --
//             decided on the device)
//   finalize  per query: radix-select the kprime best bf16 candidates, rescore
//             them exactly (fp32 rows rebuilt from the planes, fp64 accumulate),
//             sort, and CERTIFY: every row that is not a candidate has bf16 score
//             <= bound, hence exact score <= bound + eps(|q|, max|x|, d); the
//             result is accepted iff the k-th exact score is above that.
//   collect   queries whose certificate failed go through the tensor-core scan again
//             with a fixed threshold (k-th candidate score - eps) that provably keeps
//             the whole top-k; the rows above it are scored exactly -> exact result.
//   exact     queries whose row list overflowed in `collect` (or every query in
//             RR_MODE_EXACT / for dimensions the tensor-core kernels do not take) are
//             scanned with the exact kernel (fp64 accumulate, lists of depth 32 >= k).
// Every stage after `finalize` is launched unconditionally and exits at once when no
// query needs it.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <vector>

#include "rr_index_internal.cuh"
#include "rr_topm.cuh"

// ------------------------------------------------------------------ kernels --

// index.add: fp32 rows -> hi/lo planes (+ zero padding to d_pad); per corpus the largest row norm
// max|x| and the largest norm of a row's bf16 ROUNDING ERROR max|x - bf16(x)| (both feed the
// certificate's error bound, see rr_finalize_one).  Squares are exact in fp64 and so is x - bf16(x)
// in fp32 (Sterbenz), so both maxima are upper bounds once rounded up.
// one warp per row
__device__ __forceinline__ uint32_t rr_norm_bits_up(double sumsq) {
    // sqrt in double is correctly rounded; the factor covers its half ulp and the summation error
    return __float_as_uint(__double2float_ru(sqrt(sumsq) * (1.0 + 1e-12)));
}

__global__ void __launch_bounds__(256)
rr_split_rows_kernel(const float* __restrict__ x, int64_t n, int d, int d_pad,
                     uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                     uint32_t* __restrict__ max_norm_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    double max_ss = 0.0, max_es = 0.0;
    for (int64_t row = warp; row < n; row += n_warps) {
        const float* xr = x + row * d;
        uint16_t* hr = hi + row * d_pad;
        uint16_t* lr = lo + row * d_pad;
        double ss = 0.0, es = 0.0;
        for (int e = lane; e < d_pad; e += 32) {
            uint16_t h = 0, l = 0x8000u;  // joins back to +0.0f
            if (e < d) {
                const float v = xr[e];
                rr_split_f32(v, h, l);
                const float dv = v - __uint_as_float((uint32_t)h << 16);
                ss = fma((double)v, (double)v, ss);
                es = fma((double)dv, (double)dv, es);
            }
            hr[e] = h;
            lr[e] = l;
        }
        max_ss = fmax(max_ss, rr_warp_sum(ss));
        max_es = fmax(max_es, rr_warp_sum(es));
    }
    if (lane == 0) {  // positive floats order as uints
        if (max_ss > 0.0) atomicMax(max_norm_bits, rr_norm_bits_up(max_ss));
        if (max_es > 0.0) atomicMax(max_norm_bits + 1, rr_norm_bits_up(max_es));
    }
}

__global__ void __launch_bounds__(256)
rr_join_rows_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo,
                    int64_t row0, int64_t n, int d, int d_pad, float* __restrict__ out) {
    const int64_t total = n * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / d;
        const int e = (int)(i - r * d);
        const int64_t o = (row0 + r) * d_pad + e;
        const uint32_t b = (((uint32_t)hi[o] << 16) | lo[o]) - 0x8000u;
        out[i] = __uint_as_float(b);
    }
}

__global__ void __launch_bounds__(128)
rr_normalize_l2_kernel(float* __restrict__ x, int64_t n, int d) {
    rr_pdl_enter();
    // faiss fvec_renorm_L2: x /= sqrt(sum x^2) when the norm is positive
    const int64_t row = blockIdx.x;
    if (row >= n) return;
    float* xr = x + row * d;
    __shared__ float part[4];
    float ss = 0.f;
    for (int e = threadIdx.x; e < d; e += blockDim.x) ss = fmaf(xr[e], xr[e], ss);
    ss = rr_warp_sum(ss);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = ss;
    __syncthreads();
    const float tot = part[0] + part[1] + part[2] + part[3];
    if (tot > 0.f) {
        const float inv = 1.0f / sqrtf(tot);
        for (int e = threadIdx.x; e < d; e += blockDim.x) xr[e] *= inv;
    }
}

// Column sums of the stored rows in fp64 (the router's per-source centroid, router.py:136-151 reads it from
// <corpus>_stats.json["centroid"]): block b sums rows b, b + gridDim.x, ...; thread t the columns t, t + 256, ...
__global__ void __launch_bounds__(256)
rr_column_sums_kernel(const uint16_t* __restrict__ hi, const uint16_t* __restrict__ lo, int64_t n, int d, int d_pad,
                      double* __restrict__ sums) {
    for (int e0 = threadIdx.x * 2; e0 < d; e0 += 512) {
        double a0 = 0.0, a1 = 0.0;
        for (int64_t row = blockIdx.x; row < n; row += gridDim.x) {
            const uint32_t hw = *reinterpret_cast<const uint32_t*>(hi + row * d_pad + e0);
            const uint32_t lw = *reinterpret_cast<const uint32_t*>(lo + row * d_pad + e0);
            const float2 v = rr_join_f32x2(hw, lw);
            a0 += (double)v.x;
            a1 += (double)v.y;
        }
        atomicAdd(sums + e0, a0);
        if (e0 + 1 < d) atomicAdd(sums + e0 + 1, a1);
    }
}

// ---- batched launches ---------------------------------------------------------------------------
// A routed step searches S sources for the same batch (the reference's fan-out,
// /root/reference/ragroute/http_server.py:200-209).  Every stage that is not a scan runs as ONE
// launch for all sources of the step: blockIdx.y (blockIdx.x for `begin`, .z for the exact scan)
// selects the source, whose arguments sit side by side in the kernel parameters.  A plain
// index.search is the S = 1 case of the same code.
constexpr int RR_FED_MAX = 8;  // sources per batched launch (larger federations run in groups)
template <class T>
struct rr_multi {
    T s[RR_FED_MAX];
};

struct rr_prep_args {
    const float* q;  // [nq, d]
    const uint8_t* mask;
    int64_t mask_stride;
    uint16_t* qbf16;
    float* qnorm;  // [2][nq_pad]: |bf16(q)|, then |q - bf16(q)| (both rounded up)
    uint8_t* active;
    uint8_t* fail_flags;
    int* fail_list;
    rr_chunk_ctl* ctl;
    int* cpos;  // large-batch path only
    int nq;
    int nq_pad;
    int d;
    int d_pad;
    int exact_all;
    int compact;  // large-batch path: compact the routed queries to the front
};

// First launch of a chunk, one block per source: zero the chunk's control block and, on the large-batch
// path, compact the ROUTED queries (<= RR_GEMM_QCHUNK) to the front.  cpos[b] = row of query b in the
// bf16 query buffer: routed queries keep their order in [0, n_active), unrouted ones follow in
// [n_active, nq) (their rows are zero).  1024 threads, kCompactPer consecutive queries per thread.
constexpr int kCompactPer = RR_GEMM_QCHUNK / 1024;
static_assert(RR_GEMM_QCHUNK % 1024 == 0 && kCompactPer >= 1 && kCompactPer <= 8, "compaction block covers the chunk");

__global__ void __launch_bounds__(1024)
rr_begin_chunk_kernel(const __grid_constant__ rr_multi<rr_prep_args> m) {
    rr_pdl_enter();
    const rr_prep_args& a = m.s[blockIdx.x];
    __shared__ int warp_cnt[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid < (int)(sizeof(rr_chunk_ctl) / sizeof(int))) reinterpret_cast<int*>(a.ctl)[tid] = 0;
    if (!a.compact) return;
    const uint8_t* __restrict__ mask = a.mask;
    const int nq = a.nq;
    bool act[kCompactPer];
    int mine = 0;
#pragma unroll
    for (int u = 0; u < kCompactPer; ++u) {
        const int b = tid * kCompactPer + u;
        act[u] = (b < nq) && (mask ? (mask[(int64_t)b * a.mask_stride] != 0) : true);
        mine += act[u] ? 1 : 0;
    }
    // exclusive prefix of `mine` over the block
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(RR_FULL, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_cnt[w] = incl;
    __syncthreads();  // also orders the zeroing of ctl before n_active is set
    int before = incl - mine, total = 0;
    for (int i = 0; i < 32; ++i) {
        const int c = warp_cnt[i];
        if (i < w) before += c;
        total += c;
    }
#pragma unroll
    for (int u = 0; u < kCompactPer; ++u) {
        const int b = tid * kCompactPer + u;
        if (b < nq) a.cpos[b] = act[u] ? before : total + (b - before);
        before += act[u] ? 1 : 0;
    }
    if (tid == 0) a.ctl->n_active = total;
}

// fp32 queries -> bf16 tile rows (zero padded to d_pad; rows >= nq zeroed), norms,
// activity flags.  grid = (max nq_pad, sources).
__global__ void __launch_bounds__(128)
rr_prep_queries_kernel(const __grid_constant__ rr_multi<rr_prep_args> m) {
    rr_pdl_enter();
    const rr_prep_args& a = m.s[blockIdx.y];
    const int b = blockIdx.x;
    if (b >= a.nq_pad) return;
    const int nq = a.nq, d = a.d, d_pad = a.d_pad;
    const int* __restrict__ cpos = a.compact ? a.cpos : nullptr;
    // large-batch path: routed queries are compacted to the front of the bf16 buffer
    uint16_t* out = a.qbf16 + (int64_t)((cpos && b < nq) ? cpos[b] : b) * d_pad;
    if (b >= nq) {
        for (int e = threadIdx.x; e < d_pad; e += blockDim.x) out[e] = 0;
        if (threadIdx.x == 0) a.active[b] = 0;
        return;
    }
    const bool act = a.mask ? (a.mask[(int64_t)b * a.mask_stride] != 0) : true;
    const float* qr = a.q + (int64_t)b * d;
    // |bf16(q)| and |q - bf16(q)| for the certificate, accumulated in fp64 (squares are exact there)
    __shared__ double part[2][4];
    double hs = 0.0, es = 0.0;
    for (int e = threadIdx.x; e < d_pad; e += blockDim.x) {
        uint16_t h = 0, l = 0;
        if (e < d && act) {
            const float v = qr[e];
            rr_split_f32(v, h, l);
            const float vh = __uint_as_float((uint32_t)h << 16);
            const float dv = v - vh;
            hs = fma((double)vh, (double)vh, hs);
            es = fma((double)dv, (double)dv, es);
        }
        (void)l;
        out[e] = h;
    }
    hs = rr_warp_sum(hs);
    es = rr_warp_sum(es);
    if ((threadIdx.x & 31) == 0) {
        part[0][threadIdx.x >> 5] = hs;
        part[1][threadIdx.x >> 5] = es;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        a.qnorm[b] = __uint_as_float(rr_norm_bits_up(part[0][0] + part[0][1] + part[0][2] + part[0][3]));
        a.qnorm[a.nq_pad + b] = __uint_as_float(rr_norm_bits_up(part[1][0] + part[1][1] + part[1][2] + part[1][3]));
        a.active[b] = act ? 1 : 0;
        a.fail_flags[b] = 0;
        if (act) {
            a.ctl->any_active = 1;
            if (a.exact_all) {
                a.fail_flags[b] = 1;
                a.fail_list[atomicAdd(&a.ctl->n_fail, 1)] = b;
            }
        }
    }
}

// begin + prep in ONE launch for chunks of <= RR_FUSED_PREP_MAX padded queries (every small-batch step: the two
// launches cost more than their work).  One block of 1024 threads per source: zero the control block, compact the
// routed queries (large-batch path), then one WARP per query converts it (same arithmetic as the kernel above).
constexpr int RR_FUSED_PREP_MAX = 64;  // padded queries: beyond one or two queries per warp the grid-wide kernels win

__global__ void __launch_bounds__(1024)
rr_begin_prep_fused_kernel(const __grid_constant__ rr_multi<rr_prep_args> m) {
    rr_pdl_enter();
    const rr_prep_args& a = m.s[blockIdx.x];
    __shared__ int warp_cnt[32];
    __shared__ int s_cpos[RR_FUSED_PREP_MAX];
    __shared__ int s_any;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int nq = a.nq, nq_pad = a.nq_pad, d = a.d, d_pad = a.d_pad;
    if (nq_pad <= 0) return;
    const uint8_t* __restrict__ mask = a.mask;
    // routed? -- thread b < nq_pad looks at query b
    const bool act_b = (tid < nq) && (mask ? (mask[(int64_t)tid * a.mask_stride] != 0) : true);
    int n_active = 0;
    if (a.compact) {
        const unsigned bal = __ballot_sync(RR_FULL, act_b);
        if (lane == 0) warp_cnt[w] = __popc(bal);
        __syncthreads();
        int before = __popc(bal & ((1u << lane) - 1u));
        for (int i = 0; i < 32; ++i) {
            const int c = warp_cnt[i];
            if (i < w) before += c;
            n_active += c;
        }
        if (tid < nq) {
            const int cp = act_b ? before : n_active + (tid - before);
            a.cpos[tid] = cp;
            s_cpos[tid] = cp;
        }
    }
    if (tid == 0) s_any = 0;
    __syncthreads();
    const bool vec4 = (d & 3) == 0 && (reinterpret_cast<uintptr_t>(a.q) & 15) == 0;
    for (int b = w; b < nq_pad; b += 32) {
        uint16_t* out = a.qbf16 + (int64_t)((a.compact && b < nq) ? s_cpos[b] : b) * d_pad;
        if (b >= nq) {
            for (int e = lane * 8; e < d_pad; e += 256) *reinterpret_cast<uint4*>(out + e) = make_uint4(0, 0, 0, 0);
            if (lane == 0) a.active[b] = 0;
            continue;
        }
        const bool act = mask ? (mask[(int64_t)b * a.mask_stride] != 0) : true;
        const float* qr = a.q + (int64_t)b * d;
        double hs = 0.0, es = 0.0;
        if (vec4) {
            // 8 independent 16-byte loads in flight per lane (1024 dimensions per round): the row is one latency
            for (int e0 = 0; e0 < d_pad; e0 += 1024) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = e0 + (lane + 32 * u) * 4;
                    v[u] = (e < d && act) ? __ldg(reinterpret_cast<const float4*>(qr + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int e = e0 + (lane + 32 * u) * 4;
                    if (e < d_pad) {
                        const float x4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
                        uint16_t h4[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            uint16_t l;
                            rr_split_f32(x4[c], h4[c], l);
                            (void)l;
                            const float vh = __uint_as_float((uint32_t)h4[c] << 16);
                            const float dv = x4[c] - vh;
                            hs = fma((double)vh, (double)vh, hs);
                            es = fma((double)dv, (double)dv, es);
                        }
                        *reinterpret_cast<uint2*>(out + e) =
                            make_uint2((uint32_t)h4[0] | ((uint32_t)h4[1] << 16), (uint32_t)h4[2] | ((uint32_t)h4[3] << 16));
                    }
                }
            }
        } else {
            for (int e = lane; e < d_pad; e += 32) {
                uint16_t h = 0, l = 0;
                if (e < d && act) {
                    const float v = qr[e];
                    rr_split_f32(v, h, l);
                    const float vh = __uint_as_float((uint32_t)h << 16);
                    const float dv = v - vh;
                    hs = fma((double)vh, (double)vh, hs);
                    es = fma((double)dv, (double)dv, es);
                }
                (void)l;
                out[e] = h;
            }
        }
        hs = rr_warp_sum(hs);
        es = rr_warp_sum(es);
        if (lane == 0) {
            a.qnorm[b] = __uint_as_float(rr_norm_bits_up(hs));
            a.qnorm[nq_pad + b] = __uint_as_float(rr_norm_bits_up(es));
            a.active[b] = act ? 1 : 0;
            a.fail_flags[b] = a.exact_all && act ? 1 : 0;
            if (act) s_any = 1;
        }
    }
    __syncthreads();
    if (tid == 0) {  // the whole control block, written once (no other block touches it in this launch)
        rr_chunk_ctl c = {};
        c.any_active = s_any;
        c.n_active = n_active;
        if (a.exact_all) {  // exact-only plans: every routed query goes to the exact scan, in query order
            int nf = 0;
            for (int b = 0; b < nq; ++b)
                if (mask ? (mask[(int64_t)b * a.mask_stride] != 0) : true) a.fail_list[nf++] = b;
            c.n_fail = nf;
        }
        *a.ctl = c;
    }
}

// Exact scan over the queries in fail_list[blockIdx.y*qsub ...): fp32 rows rebuilt from
// the planes, products and sums in fp64 (each product of two fp32 is exact in
// fp64), score rounded once to fp32.  One warp per row; per-warp lists of depth
// 32 per query, merged per block at the end.
// cand layout: [query][block][32]
struct rr_exact_args {
    const uint16_t* hi;
    const uint16_t* lo;
    int64_t n_rows;
    const float* q;
    const int* fail_list;
    const int* n_fail_ptr;
    float* cand_s;
    uint32_t* cand_r;
    int64_t cand_stride;
    int d;
    int d_pad;
    int qsub;
};

template <int QSUB>
__global__ void __launch_bounds__(RR_EXACT_WARPS * 32)
rr_exact_scan_kernel(const __grid_constant__ rr_multi<rr_exact_args> m) {
    rr_pdl_enter();
    const rr_exact_args& a = m.s[blockIdx.z];
    if (a.n_rows <= 0) return;
    const uint16_t* __restrict__ hi = a.hi;
    const uint16_t* __restrict__ lo = a.lo;
    const int64_t n_rows = a.n_rows;
    const int d_pad = a.d_pad, d = a.d, qsub = a.qsub;
    const float* __restrict__ q = a.q;
    const int* __restrict__ fail_list = a.fail_list;
    float* __restrict__ cand_s = a.cand_s;
    uint32_t* __restrict__ cand_r = a.cand_r;
    const int64_t cand_stride = a.cand_stride;
    const int* n_fail_ptr = a.n_fail_ptr;
    const int n_fail = *n_fail_ptr;
    constexpr int M = RR_EXACT_LIST_DEPTH;
    constexpr int W = RR_EXACT_WARPS;
    // sub-batches of failing queries are strided over gridDim.y (a few rows of blocks suffice:
    // normally nothing failed and every block leaves here at once)
    for (int q0 = (int)blockIdx.y * qsub; q0 < n_fail; q0 += (int)gridDim.y * qsub) {
    const int nq = min(qsub, n_fail - q0);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* qs = reinterpret_cast<float*>(smem_raw);                       // [qsub][d_pad]
    float* ls = qs + (size_t)qsub * d_pad;                                // [W][qsub][M]
    uint32_t* lr = reinterpret_cast<uint32_t*>(ls + (size_t)W * qsub * M);  // [W][qsub][M]
    float* thr = reinterpret_cast<float*>(lr + (size_t)W * qsub * M);     // [W][qsub]
    int* qid = reinterpret_cast<int*>(thr + W * qsub);                    // [qsub]

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid < nq) qid[tid] = fail_list[q0 + tid];
    __syncthreads();
    for (int i = tid; i < nq * d_pad; i += blockDim.x) {
        const int j = i / d_pad, e = i - j * d_pad;
        qs[i] = (e < d) ? q[(int64_t)qid[j] * d + e] : 0.f;
    }
    for (int i = tid; i < W * qsub * M; i += blockDim.x) {
        ls[i] = RR_NEG_INF;
        lr[i] = RR_NO_ROW;
    }
    for (int i = tid; i < W * qsub; i += blockDim.x) thr[i] = RR_NEG_INF;
    __syncthreads();

    const int64_t gw = (int64_t)blockIdx.x * W + w;
    const int64_t GW = (int64_t)gridDim.x * W;
    const int n_chunks = (d_pad + 255) >> 8;
    for (int64_t row = gw; row < n_rows; row += GW) {
        double acc[QSUB];
#pragma unroll
        for (int j = 0; j < QSUB; ++j) acc[j] = 0.0;
        const uint4* hrow = reinterpret_cast<const uint4*>(hi + row * d_pad);
        const uint4* lrow = reinterpret_cast<const uint4*>(lo + row * d_pad);
        for (int c = 0; c < n_chunks; ++c) {
            const int e0 = c * 256 + lane * 8;
            if (e0 < d_pad) {
                const uint4 hv = __ldg(hrow + (e0 >> 3));
                const uint4 lv = __ldg(lrow + (e0 >> 3));
                float xv[8];
                float2 t;
                t = rr_join_f32x2(hv.x, lv.x); xv[0] = t.x; xv[1] = t.y;
                t = rr_join_f32x2(hv.y, lv.y); xv[2] = t.x; xv[3] = t.y;
                t = rr_join_f32x2(hv.z, lv.z); xv[4] = t.x; xv[5] = t.y;
                t = rr_join_f32x2(hv.w, lv.w); xv[6] = t.x; xv[7] = t.y;
#pragma unroll
                for (int j = 0; j < QSUB; ++j) {
                    if (j < nq) {
                        const float4 qa = *reinterpret_cast<const float4*>(qs + (size_t)j * d_pad + e0);
                        const float4 qb = *reinterpret_cast<const float4*>(qs + (size_t)j * d_pad + e0 + 4);
                        double a = acc[j];
                        a = fma((double)xv[0], (double)qa.x, a);
                        a = fma((double)xv[1], (double)qa.y, a);
                        a = fma((double)xv[2], (double)qa.z, a);
                        a = fma((double)xv[3], (double)qa.w, a);
                        a = fma((double)xv[4], (double)qb.x, a);
                        a = fma((double)xv[5], (double)qb.y, a);
                        a = fma((double)xv[6], (double)qb.z, a);
                        a = fma((double)xv[7], (double)qb.w, a);
                        acc[j] = a;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < QSUB; ++j) {
            if (j < nq) {
                const float v = (float)rr_warp_sum(acc[j]);
                const int li = (w * qsub + j);
                if (v > thr[li]) {  // warp-uniform
                    float es = ls[(size_t)li * M + lane];
                    uint32_t er = lr[(size_t)li * M + lane];
                    rr_list_insert(es, er, v, (uint32_t)row, M, lane);
                    ls[(size_t)li * M + lane] = es;
                    lr[(size_t)li * M + lane] = er;
                    const float tail = __shfl_sync(RR_FULL, es, M - 1);
                    const uint32_t tail_r = __shfl_sync(RR_FULL, er, M - 1);
                    if (lane == 0) thr[li] = (tail_r != RR_NO_ROW) ? tail : RR_NEG_INF;
                    __syncwarp();
                }
            }
        }
    }
    __syncthreads();
    // block merge: warp (j % W) folds the W lists of query j
    for (int j = w; j < nq; j += W) {
        float es = ls[(size_t)(0 * qsub + j) * M + lane];
        uint32_t er = lr[(size_t)(0 * qsub + j) * M + lane];
        for (int w2 = 1; w2 < W; ++w2) {
            const float cs = ls[(size_t)(w2 * qsub + j) * M + lane];
            const uint32_t cr = lr[(size_t)(w2 * qsub + j) * M + lane];
            rr_list_merge_sorted(es, er, cs, cr, M, lane);
        }
        const int64_t o = (int64_t)qid[j] * cand_stride + (int64_t)blockIdx.x * M + lane;
        cand_s[o] = es;
        cand_r[o] = er;
    }
    __syncthreads();  // shared memory is reused by the next sub-batch
    }
}

// Per-query finalize.  RESCORE = true: candidates carry bf16 scores (tc scan);
// select kprime by score, rescore exactly, certify.  RESCORE = false: candidates
// carry exact scores (exact scan); select the k best by (score, row).
struct rr_finalize_args {
    const float* cand_s;
    const uint32_t* cand_r;
    int64_t cand_stride;
    int n_lists;
    int depth;
    int k;
    int kprime;
    const float* q;  // [nq, d]
    const float* qnorm;  // [2][nq_pad]: |bf16(q)|, |q - bf16(q)|
    int nq_pad;
    const uint16_t* hi;
    const uint16_t* lo;
    int d;
    int d_pad;
    const uint32_t* max_norm_bits;
    int64_t id_offset;
    float* D;    // row qi at D + qi * out_stride
    int64_t* I;
    int64_t out_stride;
    const uint8_t* active;
    uint8_t* fail_flags;
    int* fail_list;
    rr_chunk_ctl* ctl;
    int* nfail_total;
    int pad_inactive;  // exact-only pipeline: the exact finalize also pads unrouted queries
    float* theta_q;    // per query: threshold of the second pass when the certificate fails (or NULL)
    const int* cpos;   // large-batch path: candidate row of query qi (compacted); NULL = qi
    int lists_from_ctl;  // large-batch path: the number of lists is ctl->n_lists (decided on the device)
    int enabled;       // 0: this source takes no part in the launch
    int nq;            // queries of the chunk
    int stage_cap;     // candidate entries the block's dynamic shared memory can stage (more: read in place)
    const int* work_list;   // RESCORE = false after a tensor-core pass: only these queries need the block ...
    const int* work_count;  // ... (a few blocks stride over them; normally the list is empty)
};

constexpr int RR_FIN_THREADS = 512;
constexpr int RR_FIN_MAXSEL = 256;
constexpr int RR_FIN_BINS_LOG2 = 11;
constexpr int RR_FIN_BINS = 1 << RR_FIN_BINS_LOG2;  // histogram bins of the candidate select

// dynamic shared memory of a finalize block: the query (RESCORE) + the staged candidate lists
static size_t rr_finalize_smem(bool rescore, int d_pad, int64_t stage_cap) {
    return (rescore ? (size_t)d_pad * sizeof(float) : 0) + (size_t)stage_cap * 8;
}

// one query of one source; called by every thread of the block (blockDim.x = 256 or 512)
// CH = chunks of 256 dimensions each warp keeps in flight per row while rescoring (registers vs. latency)
template <bool RESCORE, int CH>
__device__ __forceinline__ void rr_finalize_one(const rr_finalize_args& a, const int qi) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int n_threads = blockDim.x;
    float* Dq = a.D + (int64_t)qi * a.out_stride;
    int64_t* Iq = a.I + (int64_t)qi * a.out_stride;
    if (RESCORE) {
        if (!a.active[qi]) {  // not routed to this source: padding row, as faiss pads
            for (int i = tid; i < a.k; i += blockDim.x) {
                Dq[i] = -3.402823466e+38f;
                Iq[i] = -1;
            }
            return;
        }
    } else {
        if (!a.fail_flags[qi]) {
            if (a.pad_inactive && !a.active[qi]) {
                for (int i = tid; i < a.k; i += blockDim.x) {
                    Dq[i] = -3.402823466e+38f;
                    Iq[i] = -1;
                }
            }
            return;
        }
    }

    __shared__ uint32_t hist[RR_FIN_BINS];
    __shared__ unsigned long long s_kmin[RR_FIN_THREADS / 32], s_kmax[RR_FIN_THREADS / 32];
    __shared__ unsigned long long s_lo, s_hi;
    __shared__ int s_bstar, s_cgt, s_ceq, s_cnt_gt, s_cnt_eq, s_nvalid;
    __shared__ uint32_t s_scan[RR_FIN_THREADS / 32];
    __shared__ float s_red[RR_FIN_THREADS / 32];
    __shared__ float sel_s[RR_FIN_MAXSEL];
    __shared__ uint32_t sel_r[RR_FIN_MAXSEL];
    extern __shared__ __align__(16) unsigned char fin_sm[];
    float* q_sm = reinterpret_cast<float*>(fin_sm);                               // [d_pad] (RESCORE only)
    float* cs_sm = reinterpret_cast<float*>(fin_sm) + (RESCORE ? a.d_pad : 0);    // [stage_cap] staged scores
    uint32_t* cr_sm = reinterpret_cast<uint32_t*>(cs_sm + a.stage_cap);           // [stage_cap] staged rows

    const int crow = (RESCORE && a.cpos) ? a.cpos[qi] : qi;
    const float* gs = a.cand_s + (int64_t)crow * a.cand_stride;
    const uint32_t* gr = a.cand_r + (int64_t)crow * a.cand_stride;
    const int n_lists = (RESCORE && a.lists_from_ctl) ? a.ctl->n_lists : a.n_lists;
    const int n_cand = n_lists * a.depth;
    // the select below makes a few passes over the lists: from shared memory when they fit the block's
    // allotment (sized by the host for the batch: everything for small batches, 2048 entries -- and more
    // blocks per SM -- for large ones, whose lists are few), else in place
    const bool staged = n_cand <= a.stage_cap;
    const float* cs = staged ? cs_sm : gs;
    const uint32_t* cr = staged ? cr_sm : gr;
    // ranking key of a candidate: larger = better.  RESCORE: the bf16 score; else (score, lower row first)
    auto key_of = [&](float sc, uint32_t r) -> unsigned long long {
        const unsigned long long k32 = rr_float_key(sc);
        return RESCORE ? k32 : ((k32 << 32) | (unsigned long long)(~r));
    };

    if (tid == 0) {
        s_cnt_gt = 0;
        s_cnt_eq = 0;
        s_nvalid = 0;
    }
    if (RESCORE) {
        const float* qr = a.q + (int64_t)qi * a.d;
        for (int e = tid; e < a.d_pad; e += blockDim.x) q_sm[e] = (e < a.d) ? qr[e] : 0.f;
    }
    // stage the lists; count the occupied slots, take T = largest tail of a full list (every row a
    // list dropped scores <= its tail) and the range of the keys
    float T = RR_NEG_INF;
    int nv = 0;
    unsigned long long kmin = ~0ull, kmax = 0ull;
    {
        constexpr int U = 4;
        const int depth_m1 = a.depth - 1;
        for (int i0 = tid; i0 < n_cand; i0 += n_threads * U) {
            float vs[U];
            uint32_t vr[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * n_threads;
                vs[u] = (i < n_cand) ? __ldg(gs + i) : RR_NEG_INF;
                vr[u] = (i < n_cand) ? __ldg(gr + i) : RR_NO_ROW;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * n_threads;
                if (i < n_cand) {
                    if (staged) {
                        cs_sm[i] = vs[u];
                        cr_sm[i] = vr[u];
                    }
                    if (vr[u] != RR_NO_ROW) {
                        ++nv;
                        if ((i & depth_m1) == depth_m1) T = fmaxf(T, vs[u]);  // list depths are powers of two
                        const unsigned long long key = key_of(vs[u], vr[u]);
                        kmin = min(kmin, key);
                        kmax = max(kmax, key);
                    }
                }
            }
        }
    }
    T = rr_warp_max(T);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nv += __shfl_xor_sync(RR_FULL, nv, o);
        kmin = min(kmin, __shfl_xor_sync(RR_FULL, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(RR_FULL, kmax, o));
    }
    __syncthreads();  // s_nvalid is zero, the staged lists are visible
    if (lane == 0) {
        s_red[w] = T;
        s_kmin[w] = kmin;
        s_kmax[w] = kmax;
        atomicAdd(&s_nvalid, nv);
    }
    __syncthreads();
    T = s_red[0];
    kmin = s_kmin[0];
    kmax = s_kmax[0];
    for (int i = 1; i < (n_threads >> 5); ++i) {
        T = fmaxf(T, s_red[i]);
        kmin = min(kmin, s_kmin[i]);
        kmax = max(kmax, s_kmax[i]);
    }
    const int n_valid = s_nvalid;
    const int want = RESCORE ? min(a.kprime, RR_FIN_MAXSEL) : a.k;
    int kp = min(want, n_valid);

    // Select a SUPERSET of the kp best keys, at most RR_FIN_MAXSEL of them, by linear binning of the key
    // range: one histogram pass over [lo, hi] in up to RR_FIN_BINS bins, a suffix scan finds the bin b* in
    // which the kp-th best falls; everything above b* is taken and, if it fits, all of b* too (rescoring /
    // sorting a few extra candidates costs less than narrowing down to exactly kp).  Only when b* is too
    // crowded is its range binned again (long runs of near-equal scores); a single key value that still
    // does not fit (exact ties) is cut by position.  One pass in the common case -- the 4 (8) passes of an
    // 8-bit radix select with their 12 (24) block barriers were a third of this kernel.
    float t_rest = RR_NEG_INF;
    // |bf16 score - exact score| <= eps0 for every row of the corpus (derivation at the certificate below)
    double eps0 = 0.0;
    if (RESCORE) {
        const double xn = (double)__uint_as_float(a.max_norm_bits[0]);
        const double dxn = (double)__uint_as_float(a.max_norm_bits[1]);
        const double qhn = (double)a.qnorm[qi];
        const double dqn = (double)a.qnorm[a.nq_pad + qi];
        eps0 = (qhn * dxn + dqn * xn + (double)a.d_pad * 2.384185791015625e-7 * qhn * (xn + dxn)) * 1.00001 + 1e-30;
    }
    if (kp > 0) {
        unsigned long long lo = kmin, hi = kmax;
        int need = kp, n_above = 0, take_eq = 0;
        bool first = true;
        const int per = RR_FIN_BINS / n_threads;  // bins per thread in the suffix scan (4 or 8)
        for (;;) {
            const unsigned long long range = hi - lo;
            const int sh = max(0, (64 - __clzll((long long)(range | 1ull))) - RR_FIN_BINS_LOG2);
            for (int i = tid; i < RR_FIN_BINS; i += n_threads) hist[i] = 0;
            __syncthreads();
            for (int i = tid; i < n_cand; i += n_threads) {
                const uint32_t r = cr[i];
                if (r == RR_NO_ROW) continue;
                const unsigned long long key = key_of(cs[i], r);
                if (key >= lo && key <= hi) atomicAdd(&hist[(uint32_t)((key - lo) >> sh)], 1u);
            }
            __syncthreads();
            // thread t owns bins RR_FIN_BINS - 1 - t * per ... downwards; inclusive scan of the thread totals
            uint32_t c[8];
            uint32_t tot = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                c[j] = (j < per) ? hist[RR_FIN_BINS - 1 - (tid * per + j)] : 0u;
                tot += c[j];
            }
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t v = __shfl_up_sync(RR_FULL, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_scan[w] = incl;
            __syncthreads();
            uint32_t before = 0;
            for (int i = 0; i < w; ++i) before += s_scan[i];
            incl += before;
            const uint32_t excl = incl - tot;
            if (RESCORE && first && kp > a.k) {
                // How many candidates are worth rescoring?  With c_k = the k-th best bf16 score, k rows have exact
                // score >= c_k - eps0, so a candidate whose bf16 score is below c_k - 2 eps0 has exact score
                // < c_k - eps0 <= the true k-th score: it cannot be in the top-k.  Only the candidates at or
                // above that line are rescored (at most kprime, at least k) -- typically half of kprime, and
                // the rescoring gathers are the bulk of this kernel.  (Bins make both edges conservative.)
                if (excl < (uint32_t)a.k && (uint32_t)a.k <= incl) {
                    uint32_t run = excl;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j < per && run < (uint32_t)a.k && (uint32_t)a.k <= run + c[j]) s_bstar = RR_FIN_BINS - 1 - (tid * per + j);
                        run += c[j];
                    }
                }
                __syncthreads();
                const float c_k_lo = rr_key_float((uint32_t)(lo + ((unsigned long long)s_bstar << sh)));  // <= c_k
                const float line = __double2float_rd((double)c_k_lo - 2.0 * eps0 - fabs((double)c_k_lo) * 2.4e-7);
                const unsigned long long line_key = rr_float_key(line);
                const int bt = line_key <= lo ? 0 : (int)min((line_key - lo) >> sh, (unsigned long long)(RR_FIN_BINS - 1));
                const int jt = RR_FIN_BINS - 1 - bt - tid * per;  // position of bin bt among this thread's bins
                if (jt >= 0 && jt < per) {
                    uint32_t run = excl;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j <= jt) run += c[j];
                    s_cgt = (int)run;  // candidates in the bins at or above the line
                }
                __syncthreads();
                need = min(kp, max(a.k, s_cgt));
                kp = need;
                __syncthreads();  // s_bstar / s_cgt are written again right below
            }
            first = false;
            if (excl < (uint32_t)need && (uint32_t)need <= incl) {  // exactly one thread
                uint32_t run = excl;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (j < per && run < (uint32_t)need && (uint32_t)need <= run + c[j]) {
                        s_bstar = RR_FIN_BINS - 1 - (tid * per + j);
                        s_cgt = (int)run;
                        s_ceq = (int)c[j];
                    }
                    run += c[j];
                }
            }
            __syncthreads();
            const int bstar = s_bstar, cgt = s_cgt, ceq = s_ceq;
            const unsigned long long wlo = lo + ((unsigned long long)bstar << sh);
            const unsigned long long whi = min(hi, wlo + ((1ull << sh) - 1ull));
            n_above += cgt;
            need -= cgt;
            lo = wlo;
            hi = whi;
            if (RESCORE) {
                // Exactly kp candidates are rescored (a power of two keeps the sort small): `need` members of the
                // boundary bin, whichever arrive first.  Any choice is sound -- t_rest below is the best key really
                // left behind -- and a crowded bin is only refined while that would tighten t_rest noticeably.
                if (ceq - need <= 8 || sh == 0) {
                    take_eq = need;
                    break;
                }
            } else {
                if (n_above + ceq <= RR_FIN_MAXSEL) {  // the whole boundary bin fits: the exact order is sorted out below
                    take_eq = ceq;
                    break;
                }
                if (sh == 0) {  // (keys carry the row: they are unique, a single key value never overflows)
                    take_eq = need;
                    break;
                }
            }
            __syncthreads();  // s_bstar / hist are rewritten by the next level
        }
        // gather: keys above the window all go in, the window's keys fill the remaining slots; the best key
        // left behind bounds every unselected candidate (t_rest)
        unsigned long long rest = 0ull;
        for (int i = tid; i < n_cand; i += n_threads) {
            const uint32_t r = cr[i];
            if (r == RR_NO_ROW) continue;
            const float sc = cs[i];
            const unsigned long long key = key_of(sc, r);
            if (key > hi) {
                const int p = atomicAdd(&s_cnt_gt, 1);
                sel_s[p] = sc;
                sel_r[p] = r;
            } else if (key >= lo) {
                const int p = atomicAdd(&s_cnt_eq, 1);
                if (p < take_eq) {
                    sel_s[n_above + p] = sc;
                    sel_r[n_above + p] = r;
                } else {
                    rest = max(rest, key);
                }
            } else {
                rest = max(rest, key);
            }
        }
        kp = n_above + take_eq;  // what was selected (exact finalize: possibly more than the k asked for)
        if (RESCORE && n_valid > kp) {  // (uniform: both from shared memory)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) rest = max(rest, __shfl_xor_sync(RR_FULL, rest, o));
            __syncthreads();  // s_kmax was read by everyone long ago: reuse it
            if (lane == 0) s_kmax[w] = rest;
            __syncthreads();
            rest = s_kmax[0];
            for (int i = 1; i < (n_threads >> 5); ++i) rest = max(rest, s_kmax[i]);
            t_rest = rest ? rr_key_float((uint32_t)rest) : RR_NEG_INF;
        }
    }
    // pad the selection to a power of two for the sort
    int kp_pad = 1;
    while (kp_pad < kp) kp_pad <<= 1;
    if (kp_pad < 32) kp_pad = 32;
    __syncthreads();
    for (int i = kp + tid; i < kp_pad; i += blockDim.x) {
        sel_s[i] = RR_NEG_INF;
        sel_r[i] = RR_NO_ROW;
    }
    __syncthreads();

    if (RESCORE) {
        // exact rescoring: fp32 row rebuilt from the planes, fp64 accumulation.  One warp per row, RB rows x
        // CH chunks of 256 dimensions per warp in flight at once (the gathers are random 16-byte reads of
        // both planes: latency-bound, so 16 warps x 12 loads cover 128 rows x 768 dims in 4 round trips)
        constexpr int RB = 2;
        const int NW = n_threads >> 5;
        for (int i0 = w * RB; i0 < kp; i0 += NW * RB) {
            double acc[RB];
            int64_t rowo[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                acc[r] = 0.0;
                rowo[r] = (int64_t)sel_r[min(i0 + r, kp - 1)] * a.d_pad;
            }
            for (int e0 = lane * 8; e0 < a.d_pad; e0 += 256 * CH) {
                uint4 hv[RB][CH], lv[RB][CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int e = e0 + 256 * c;
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        if (e < a.d_pad) {
                            hv[r][c] = __ldg(reinterpret_cast<const uint4*>(a.hi + rowo[r]) + (e >> 3));
                            lv[r][c] = __ldg(reinterpret_cast<const uint4*>(a.lo + rowo[r]) + (e >> 3));
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const int e = e0 + 256 * c;
                    if (e < a.d_pad) {
                        const float4 qa = *reinterpret_cast<const float4*>(q_sm + e);
                        const float4 qb = *reinterpret_cast<const float4*>(q_sm + e + 4);
#pragma unroll
                        for (int r = 0; r < RB; ++r) {
                            double s_ = acc[r];
                            float2 t;
                            t = rr_join_f32x2(hv[r][c].x, lv[r][c].x);
                            s_ = fma((double)t.x, (double)qa.x, s_);
                            s_ = fma((double)t.y, (double)qa.y, s_);
                            t = rr_join_f32x2(hv[r][c].y, lv[r][c].y);
                            s_ = fma((double)t.x, (double)qa.z, s_);
                            s_ = fma((double)t.y, (double)qa.w, s_);
                            t = rr_join_f32x2(hv[r][c].z, lv[r][c].z);
                            s_ = fma((double)t.x, (double)qb.x, s_);
                            s_ = fma((double)t.y, (double)qb.y, s_);
                            t = rr_join_f32x2(hv[r][c].w, lv[r][c].w);
                            s_ = fma((double)t.x, (double)qb.z, s_);
                            s_ = fma((double)t.y, (double)qb.w, s_);
                            acc[r] = s_;
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const double tot = rr_warp_sum(acc[r]);
                if (lane == 0 && i0 + r < kp) sel_s[i0 + r] = (float)tot;
            }
        }
        __syncthreads();
    }

    // sort sel[0, kp_pad) into ranking order (score desc, row asc): one warp, entries in registers
    if (w == 0) {
        switch (kp_pad >> 5) {
            case 1: rr_warp_sort_smem<1>(sel_s, sel_r, lane); break;
            case 2: rr_warp_sort_smem<2>(sel_s, sel_r, lane); break;
            case 4: rr_warp_sort_smem<4>(sel_s, sel_r, lane); break;
            default: rr_warp_sort_smem<8>(sel_s, sel_r, lane); break;  // RR_FIN_MAXSEL = 256
        }
    }
    __syncthreads();

    bool pass = true;
    float theta = RR_NEG_INF;  // second-pass threshold if the certificate fails
    if (RESCORE) {
        const float bound = fmaxf(T, t_rest);
        if (bound != RR_NEG_INF) {
            // |bf16 score - exact score| <= eps for EVERY row x of the corpus.  With qh = bf16(q), xh = bf16(x):
            //   qh.xh - q.x = qh.(xh - x) + (qh - q).x, hence by Cauchy-Schwarz
            //   |qh.xh - q.x| <= |qh| max|xh - x| + |qh - q| max|x|
            // with the ACTUAL rounding-error norms (measured at add / prep time, fp64, rounded up) instead of
            // the worst case 2^-8 per operand -- sound for any data (rows sitting on bf16 midpoints simply get
            // the full (2^-7 + 2^-16)|q||x|) and ~2.5x tighter for real embeddings.  Plus
            //   d_pad * 2^-22 |qh||xh|   fp32 accumulation slack in the tensor core (products are exact)
            //   2^-23 |s_k|              the k-th exact score is itself rounded once to fp32
            const double xn = (double)__uint_as_float(a.max_norm_bits[0]);
            const double dxn = (double)__uint_as_float(a.max_norm_bits[1]);
            const double qhn = (double)a.qnorm[qi];
            const double dqn = (double)a.qnorm[a.nq_pad + qi];
            const bool have_k = (kp >= a.k) && sel_r[a.k - 1] != RR_NO_ROW;
            const double sk = have_k ? (double)sel_s[a.k - 1] : 0.0;
            const double eps = (qhn * dxn + dqn * xn + (double)a.d_pad * 2.384185791015625e-7 * qhn * (xn + dxn)) * 1.00001 +
                               fabs(sk) * 1.1920928955078125e-7 + 1e-30;
            pass = have_k && (sk > (double)bound + eps);
            // L = the k-th exact score among the candidates is a lower bound of the true k-th score, and a
            // row with exact score >= L has bf16 score >= L - eps: the second pass collects every row
            // whose bf16 score is > theta = the largest float below L - eps
            if (!pass && have_k) theta = nextafterf(__double2float_rd(sk - eps), RR_NEG_INF);
        }
    }
    if (pass) {
        for (int i = tid; i < a.k; i += blockDim.x) {
            const bool ok = (i < kp) && sel_r[i] != RR_NO_ROW;
            Dq[i] = ok ? sel_s[i] : -3.402823466e+38f;
            Iq[i] = ok ? (int64_t)sel_r[i] + a.id_offset : (int64_t)-1;
        }
    } else if (tid == 0) {
        a.fail_flags[qi] = 1;
        if (RESCORE && a.theta_q) a.theta_q[qi] = theta;
        a.fail_list[atomicAdd(&a.ctl->n_fail, 1)] = qi;
        atomicAdd(a.nfail_total, 1);
    }
}

// grid = (queries, sources), or (a few blocks, sources) striding over a work list.
// BIG = false: small batches, latency-bound: 512 threads, 6 gathers in flight per lane while rescoring.
// BIG = true: large batches, throughput-bound: 256 threads and <= 64 registers, 4 blocks per SM.
template <bool RESCORE, bool BIG>
__global__ void __launch_bounds__(BIG ? RR_FIN_THREADS / 2 : RR_FIN_THREADS, BIG ? 4 : 1)
rr_finalize_kernel(const __grid_constant__ rr_multi<rr_finalize_args> m) {
    rr_pdl_enter();
    const rr_finalize_args& a = m.s[blockIdx.y];
    if (!a.enabled) return;
    constexpr int CH = BIG ? 1 : 3;
    if (!RESCORE && a.work_list) {
        const int n = *a.work_count;
        for (int wi = blockIdx.x; wi < n; wi += gridDim.x) {
            rr_finalize_one<RESCORE, CH>(a, a.work_list[wi]);
            __syncthreads();  // the block's shared state is reused by the next query
        }
        return;
    }
    for (int qi = blockIdx.x; qi < a.nq; qi += gridDim.x) {
        rr_finalize_one<RESCORE, CH>(a, qi);
        __syncthreads();
    }
}

// ---- second pass ("collect") for the queries whose certificate failed --------------------------
// Gather their bf16 rows (in fail_list order) into the second query buffer, hand their thresholds to
// the scan, zero their row counters.  grid = rows of the second buffer, 128 threads.
struct rr_prep_collect_args {
    const uint16_t* qbf16;
    const int* cpos;
    const int* fail_list;
    const float* theta_q;
    const rr_chunk_ctl* ctl;
    uint16_t* qbf2;
    float* theta2;
    int* col_cnt;
    rr_chunk_ctl* ctl2;
    int d_pad;
    int nq;
    int nq_pad2;  // 0: this source has no second pass
};

__global__ void __launch_bounds__(128)
rr_prep_collect_kernel(const __grid_constant__ rr_multi<rr_prep_collect_args> m) {
    rr_pdl_enter();
    const rr_prep_collect_args& a = m.s[blockIdx.y];
    if (a.nq_pad2 == 0) return;
    const int d_pad = a.d_pad;
    const int n_fail = a.ctl->n_fail;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.ctl2->n_active = n_fail;
        a.ctl2->any_active = n_fail > 0;
    }
    if (n_fail == 0) return;  // nothing failed: the scan that follows exits at once, leave the buffer alone
    // the rows of the second buffer its scan can touch: the failing queries, zero padded to whole query tiles
    const int n_rows = min(a.nq_pad2, (n_fail + 255) / 256 * 256);
    for (int i = blockIdx.x; i < n_rows; i += gridDim.x) {
        uint4* out = reinterpret_cast<uint4*>(a.qbf2 + (int64_t)i * d_pad);
        if (i < n_fail) {
            const int qi = a.fail_list[i];
            const uint4* in = reinterpret_cast<const uint4*>(a.qbf16 + (int64_t)(a.cpos ? a.cpos[qi] : qi) * d_pad);
            for (int e = threadIdx.x; e < d_pad / 8; e += blockDim.x) out[e] = in[e];
            if (threadIdx.x == 0) a.theta2[i] = a.theta_q[qi];
        } else {
            for (int e = threadIdx.x; e < d_pad / 8; e += blockDim.x) out[e] = make_uint4(0, 0, 0, 0);
            if (threadIdx.x == 0) a.theta2[i] = __builtin_huge_valf();
        }
        if (threadIdx.x == 0 && i < a.nq) a.col_cnt[i] = 0;
    }
}

constexpr int RR_COLLECT_CAP = 2048;  // rows a failing query may collect before it goes to the exact scan

struct rr_collect_fin_args {
    const uint32_t* col_rows;  // [i][RR_COLLECT_CAP], i = position in fail_list
    const int* col_cnt;
    const int* fail_list;
    rr_chunk_ctl* ctl;
    int* fail2_list;
    uint8_t* fail_flags;
    int* n_exact_total;
    int k;
    const float* q;
    const uint16_t* hi;
    const uint16_t* lo;
    int d;
    int d_pad;
    int64_t id_offset;
    float* D;
    int64_t* I;
    int64_t out_stride;
    int enabled;
};

// One block per failing query: every row that can possibly belong to its top-k was collected; score
// them exactly (fp32 rows rebuilt from the planes, fp64 accumulate) and keep the k best by
// (score desc, row asc) -- no certificate needed.  A query whose list overflowed goes to the exact scan.
__device__ __forceinline__ void rr_finalize_collect_one(const rr_collect_fin_args& a, const int i) {
    const int qi = a.fail_list[i];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int cnt = a.col_cnt[i];
    if (cnt > RR_COLLECT_CAP) {  // too many near-ties: exact scan
        if (tid == 0) {
            a.fail2_list[atomicAdd(&a.ctl->n_fail2, 1)] = qi;
            atomicAdd(a.n_exact_total, 1);
        }
        return;
    }
    __shared__ float s_s[RR_COLLECT_CAP];
    __shared__ uint32_t s_r[RR_COLLECT_CAP];
    extern __shared__ __align__(16) float q_sm[];  // [d_pad]
    const float* qr = a.q + (int64_t)qi * a.d;
    for (int e = tid; e < a.d_pad; e += blockDim.x) q_sm[e] = (e < a.d) ? qr[e] : 0.f;
    int n_pad = 32;
    while (n_pad < cnt) n_pad <<= 1;
    for (int j = tid; j < n_pad; j += blockDim.x) {
        s_r[j] = (j < cnt) ? a.col_rows[(int64_t)i * RR_COLLECT_CAP + j] : RR_NO_ROW;
        s_s[j] = RR_NEG_INF;
    }
    __syncthreads();
    constexpr int RB = 4;
    const int NW = blockDim.x >> 5;
    for (int i0 = w * RB; i0 < cnt; i0 += NW * RB) {
        double acc[RB];
#pragma unroll
        for (int r = 0; r < RB; ++r) acc[r] = 0.0;
        for (int e0 = lane * 8; e0 < a.d_pad; e0 += 256) {
            uint4 hv[RB], lv[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int64_t row = s_r[min(i0 + r, cnt - 1)];
                hv[r] = __ldg(reinterpret_cast<const uint4*>(a.hi + row * a.d_pad) + (e0 >> 3));
                lv[r] = __ldg(reinterpret_cast<const uint4*>(a.lo + row * a.d_pad) + (e0 >> 3));
            }
            const float4 qa = *reinterpret_cast<const float4*>(q_sm + e0);
            const float4 qb = *reinterpret_cast<const float4*>(q_sm + e0 + 4);
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                double s_ = acc[r];
                float2 t;
                t = rr_join_f32x2(hv[r].x, lv[r].x);
                s_ = fma((double)t.x, (double)qa.x, s_);
                s_ = fma((double)t.y, (double)qa.y, s_);
                t = rr_join_f32x2(hv[r].y, lv[r].y);
                s_ = fma((double)t.x, (double)qa.z, s_);
                s_ = fma((double)t.y, (double)qa.w, s_);
                t = rr_join_f32x2(hv[r].z, lv[r].z);
                s_ = fma((double)t.x, (double)qb.x, s_);
                s_ = fma((double)t.y, (double)qb.y, s_);
                t = rr_join_f32x2(hv[r].w, lv[r].w);
                s_ = fma((double)t.x, (double)qb.z, s_);
                s_ = fma((double)t.y, (double)qb.w, s_);
                acc[r] = s_;
            }
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const double tot = rr_warp_sum(acc[r]);
            if (lane == 0 && i0 + r < cnt) s_s[i0 + r] = (float)tot;
        }
    }
    __syncthreads();
    for (int ksz = 2; ksz <= n_pad; ksz <<= 1) {
        for (int j = ksz >> 1; j > 0; j >>= 1) {
            for (int x = tid; x < n_pad; x += blockDim.x) {
                const int y = x ^ j;
                if (y > x) {
                    const bool first_low = (x & ksz) == 0;
                    const float s1 = s_s[x], s2 = s_s[y];
                    const uint32_t r1 = s_r[x], r2 = s_r[y];
                    const bool swap = first_low ? rr_before(s2, r2, s1, r1) : rr_before(s1, r1, s2, r2);
                    if (swap) {
                        s_s[x] = s2;
                        s_r[x] = r2;
                        s_s[y] = s1;
                        s_r[y] = r1;
                    }
                }
            }
            __syncthreads();
        }
    }
    float* Dq = a.D + (int64_t)qi * a.out_stride;
    int64_t* Iq = a.I + (int64_t)qi * a.out_stride;
    for (int x = tid; x < a.k; x += blockDim.x) {
        const bool ok = (x < cnt) && s_r[x] != RR_NO_ROW;
        Dq[x] = ok ? s_s[x] : -3.402823466e+38f;
        Iq[x] = ok ? (int64_t)s_r[x] + a.id_offset : (int64_t)-1;
    }
    if (tid == 0) a.fail_flags[qi] = 0;  // settled: the exact finalize skips it
}

// grid = (a few blocks, sources): the blocks stride over the failing queries (normally none)
__global__ void __launch_bounds__(RR_FIN_THREADS)
rr_finalize_collect_kernel(const __grid_constant__ rr_multi<rr_collect_fin_args> m) {
    rr_pdl_enter();
    const rr_collect_fin_args& a = m.s[blockIdx.y];
    if (!a.enabled) return;
    const int n_fail = a.ctl->n_fail;
    for (int i = blockIdx.x; i < n_fail; i += gridDim.x) {
        rr_finalize_collect_one(a, i);
        __syncthreads();  // the block's shared lists are reused by the next query
    }
}

// ------------------------------------------------------------------- host ---

static int rr_grow(void** p, size_t* have, size_t want) {
    if (*have >= want) return RR_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    size_t sz = want + want / 4;
    cudaError_t e = cudaMalloc(p, sz);
    if (e != cudaSuccess) {
        e = cudaMalloc(p, want);
        sz = want;
    }
    if (e != cudaSuccess) {
        rr_set_error("cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
        cudaGetLastError();
        return RR_ERR_OOM;
    }
    *have = sz;
    return RR_OK;
}

extern "C" int rr_index_create(int d, int64_t capacity_rows, int device, rr_index_t** out) {
    RR_REQUIRE(out != nullptr, "rr_index_create: out is NULL");
    *out = nullptr;
    RR_REQUIRE(d > 0 && d <= 65536, "rr_index_create: invalid dimension d=%d", d);
    RR_REQUIRE(capacity_rows >= 0, "rr_index_create: negative capacity");
    int rc = rr_check_device_arch(device);
    if (rc != RR_OK) return rc;
    RR_REQUIRE(device < RR_MAX_DEVICES, "rr_index_create: device %d >= %d", device, RR_MAX_DEVICES);
    rr_device_guard g(device);
    rr_index* h = new (std::nothrow) rr_index();
    RR_REQUIRE(h != nullptr, "rr_index_create: out of host memory");
    h->d = d;
    h->d_pad = (d + 63) / 64 * 64;
    h->device = device;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) {
        delete h;
        rr_set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e));
        return RR_ERR_CUDA;
    }
    h->sm_count = prop.multiProcessorCount;
    e = cudaMalloc(&h->max_norm_bits, 16);
    if (e == cudaSuccess) e = cudaMemset(h->max_norm_bits, 0, 16);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->add_event, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&h->d_nfail_total, 32);
    if (e == cudaSuccess) e = cudaMemset(h->d_nfail_total, 0, 32);
    if (e == cudaSuccess) e = cudaMalloc(&h->scan_progress, RR_MAX_SCAN_UNITS * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(h->scan_progress, 0, RR_MAX_SCAN_UNITS * sizeof(unsigned long long));
    if (e != cudaSuccess) {
        rr_set_error("rr_index_create: %s", cudaGetErrorString(e));
        rr_index_destroy(h);
        return RR_ERR_CUDA;
    }
    if (capacity_rows > 0) {
        rc = rr_index_reserve(h, capacity_rows);
        if (rc != RR_OK) {
            rr_index_destroy(h);
            return rc;
        }
    }
    *out = h;
    return RR_OK;
}

extern "C" int rr_index_destroy(rr_index_t* h) {
    if (!h) return RR_OK;
    rr_device_guard g(h->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < rr_index::kEvPool; ++i)
        if (h->ev_pool[i]) cudaEventDestroy(h->ev_pool[i]);
    if (h->host_stream) cudaStreamDestroy(h->host_stream);
    if (h->add_event) cudaEventDestroy(h->add_event);
    cudaFree(h->hi);
    cudaFree(h->lo);
    cudaFree(h->max_norm_bits);
    cudaFree(h->d_nfail_total);
    cudaFree(h->scan_progress);
    cudaFree(h->scratch);
    cudaFree(h->stage);
    if (h->pin) cudaFreeHost(h->pin);
    delete h;
    return RR_OK;
}

extern "C" int rr_index_reserve(rr_index_t* h, int64_t capacity_rows) {
    RR_REQUIRE(h != nullptr, "rr_index_reserve: NULL handle");
    if (capacity_rows <= h->cap) return RR_OK;
    rr_device_guard g(h->device);
    const size_t plane = (size_t)capacity_rows * h->d_pad * sizeof(uint16_t);
    uint16_t *nhi = nullptr, *nlo = nullptr;
    cudaError_t e = cudaMalloc(&nhi, plane);
    if (e == cudaSuccess) e = cudaMalloc(&nlo, plane);
    if (e != cudaSuccess) {
        cudaFree(nhi);
        cudaGetLastError();
        rr_set_error("rr_index_reserve: cannot allocate 2 x %zu bytes for %lld rows: %s", plane,
                     (long long)capacity_rows, cudaGetErrorString(e));
        return RR_ERR_OOM;
    }
    // earlier asynchronous adds / searches (any stream) may still touch the old planes
    RR_CUDA_CHECK(cudaDeviceSynchronize());
    if (h->n > 0) {
        const size_t used = (size_t)h->n * h->d_pad * sizeof(uint16_t);
        RR_CUDA_CHECK(cudaMemcpy(nhi, h->hi, used, cudaMemcpyDeviceToDevice));
        RR_CUDA_CHECK(cudaMemcpy(nlo, h->lo, used, cudaMemcpyDeviceToDevice));
    }
    cudaFree(h->hi);
    cudaFree(h->lo);
    h->hi = nhi;
    h->lo = nlo;
    h->cap = capacity_rows;
    h->tmap_hi_ptr = nullptr;
    h->tmap_hi64_ptr = nullptr;
    return RR_OK;
}

extern "C" int rr_index_add_f32(rr_index_t* h, const float* x, int64_t n, int x_on_device,
                                void* stream_) {
    RR_REQUIRE(h != nullptr, "rr_index_add_f32: NULL handle");
    RR_REQUIRE(n >= 0, "rr_index_add_f32: negative n");
    if (n == 0) return RR_OK;
    RR_REQUIRE(x != nullptr, "rr_index_add_f32: x is NULL");
    // local row ids are u32 in the candidate lists and int32 TMA coordinates in the scan kernels
    RR_REQUIRE(h->n + n <= 0x7FFFFF00ll, "rr_index_add_f32: more than 2^31 rows in one shard");
    rr_device_guard g(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    if (h->n + n > h->cap) {
        // growing copies on the legacy stream: make prior async work visible first
        RR_CUDA_CHECK(cudaStreamSynchronize(stream));
        int64_t want = std::max<int64_t>(h->n + n, h->cap + h->cap / 2);
        int rc = rr_index_reserve(h, want);
        if (rc == RR_ERR_OOM && want > h->n + n) rc = rr_index_reserve(h, h->n + n);
        if (rc != RR_OK) return rc;
    }
    const int blocks = h->sm_count * 8;
    if (x_on_device) {
        rr_split_rows_kernel<<<blocks, 256, 0, stream>>>(x, n, h->d, h->d_pad,
                                                         h->hi + (size_t)h->n * h->d_pad,
                                                         h->lo + (size_t)h->n * h->d_pad,
                                                         h->max_norm_bits);
        RR_CUDA_CHECK(cudaGetLastError());
    } else {
        const size_t row_bytes = (size_t)h->d * sizeof(float);
        int64_t rows_per = std::max<int64_t>(1, (int64_t)((64u << 20) / row_bytes));
        rows_per = std::min(rows_per, n);
        int rc = rr_grow((void**)&h->stage, &h->stage_bytes, (size_t)rows_per * row_bytes);
        if (rc != RR_OK) return rc;
        for (int64_t r0 = 0; r0 < n; r0 += rows_per) {
            const int64_t nr = std::min(rows_per, n - r0);
            RR_CUDA_CHECK(cudaMemcpyAsync(h->stage, x + (size_t)r0 * h->d, (size_t)nr * row_bytes,
                                          cudaMemcpyHostToDevice, stream));
            rr_split_rows_kernel<<<blocks, 256, 0, stream>>>(
                h->stage, nr, h->d, h->d_pad, h->hi + (size_t)(h->n + r0) * h->d_pad,
                h->lo + (size_t)(h->n + r0) * h->d_pad, h->max_norm_bits);
            RR_CUDA_CHECK(cudaGetLastError());
            RR_CUDA_CHECK(cudaStreamSynchronize(stream));  // stage is reused
        }
    }
    h->n += n;
    if (x_on_device) {  // consumed asynchronously: rr_index_search_host (own stream) orders itself after it
        RR_CUDA_CHECK(cudaEventRecord(h->add_event, stream));
        h->add_event_pending = true;
    }
    return RR_OK;
}

extern "C" int rr_index_reset(rr_index_t* h) {
    RR_REQUIRE(h != nullptr, "rr_index_reset: NULL handle");
    rr_device_guard g(h->device);
    RR_CUDA_CHECK(cudaDeviceSynchronize());
    h->n = 0;
    RR_CUDA_CHECK(cudaMemset(h->max_norm_bits, 0, 16));
    return RR_OK;
}

extern "C" int64_t rr_index_ntotal(const rr_index_t* h) { return h ? h->n : -1; }
extern "C" int rr_index_dim(const rr_index_t* h) { return h ? h->d : -1; }

extern "C" int rr_index_reconstruct(rr_index_t* h, int64_t row0, int64_t n, float* out_dev,
                                    void* stream) {
    RR_REQUIRE(h != nullptr && out_dev != nullptr, "rr_index_reconstruct: NULL argument");
    RR_REQUIRE(row0 >= 0 && n >= 0 && row0 + n <= h->n, "rr_index_reconstruct: rows [%lld, %lld) out of range (ntotal=%lld)",
               (long long)row0, (long long)(row0 + n), (long long)h->n);
    if (n == 0) return RR_OK;
    rr_device_guard g(h->device);
    rr_join_rows_kernel<<<h->sm_count * 8, 256, 0, (cudaStream_t)stream>>>(h->hi, h->lo, row0, n, h->d,
                                                                          h->d_pad, out_dev);
    RR_CUDA_CHECK(cudaGetLastError());
    return RR_OK;
}

extern "C" int rr_index_centroid(rr_index_t* h, float* out_host) {
    RR_REQUIRE(h != nullptr && out_host != nullptr, "rr_index_centroid: NULL argument");
    RR_REQUIRE(h->n > 0, "rr_index_centroid: empty index");
    rr_device_guard g(h->device);
    double* sums = nullptr;
    RR_CUDA_CHECK(cudaDeviceSynchronize());  // rows added asynchronously on any stream are in place
    RR_CUDA_CHECK(cudaMalloc(&sums, (size_t)h->d * sizeof(double)));
    cudaError_t e = cudaMemset(sums, 0, (size_t)h->d * sizeof(double));
    if (e == cudaSuccess) {
        const int blocks = (int)std::min<int64_t>(h->n, (int64_t)h->sm_count * 16);
        rr_column_sums_kernel<<<blocks, 256>>>(h->hi, h->lo, h->n, h->d, h->d_pad, sums);
        e = cudaGetLastError();
    }
    std::vector<double> host((size_t)h->d);
    if (e == cudaSuccess) e = cudaMemcpy(host.data(), sums, (size_t)h->d * sizeof(double), cudaMemcpyDeviceToHost);
    cudaFree(sums);
    if (e != cudaSuccess) {
        rr_set_error("rr_index_centroid: %s", cudaGetErrorString(e));
        return RR_ERR_CUDA;
    }
    for (int i = 0; i < h->d; ++i) out_host[i] = (float)(host[(size_t)i] / (double)h->n);
    return RR_OK;
}

extern "C" int rr_index_set_id_offset(rr_index_t* h, int64_t offset) {
    RR_REQUIRE(h != nullptr, "rr_index_set_id_offset: NULL handle");
    h->id_offset = offset;
    return RR_OK;
}

extern "C" int rr_index_set_param(rr_index_t* h, const char* name, double value) {
    RR_REQUIRE(h != nullptr && name != nullptr, "rr_index_set_param: NULL argument");
    if (!strcmp(name, "mode")) {
        RR_REQUIRE(value == RR_MODE_AUTO || value == RR_MODE_EXACT, "mode must be 0 (auto) or 1 (exact)");
        h->mode = (int)value;
    } else if (!strcmp(name, "kprime")) {
        RR_REQUIRE(value >= RR_MAX_K && value <= RR_FIN_MAXSEL, "kprime must be in [%d, %d]", RR_MAX_K, RR_FIN_MAXSEL);
        h->kprime = (int)value;
    } else if (!strcmp(name, "profile")) {
        h->profile = value != 0;
    } else if (!strcmp(name, "scan_path")) {
        RR_REQUIRE(value == RR_PATH_AUTO || value == RR_PATH_SMALL || value == RR_PATH_GEMM || value == RR_PATH_GEMM_PAIR,
                   "scan_path must be 0 (auto), 1 (64-query passes), 2 (large-batch kernel) or 3 (large-batch, CTA pairs)");
        h->scan_path = (int)value;
    } else if (!strcmp(name, "gemm_depth")) {
        RR_REQUIRE(value == 0 || value == 8 || value == 32, "gemm_depth must be 0 (automatic), 8 or 32");
        h->gemm_depth = (int)value;
    } else if (!strcmp(name, "gemm_tmem_all")) {
        h->gemm_tmem_all = value != 0;
    } else if (!strcmp(name, "gemm_stage_kc")) {
        RR_REQUIRE(value == 2 || value == 4, "gemm_stage_kc must be 2 or 4");
        h->gemm_stage_kc = (int)value;
    } else if (!strcmp(name, "gemm_pair_min")) {
        RR_REQUIRE(value >= 2 && value <= RR_GEMM_QTILE + 1, "gemm_pair_min must be in [2, %d]", RR_GEMM_QTILE + 1);
        h->gemm_pair_min = (int)value;
    } else if (!strcmp(name, "lockstep")) {
        RR_REQUIRE(value >= 0 && value <= 4096, "lockstep must be 0 (off) or a window of 1 ... 4096 corpus tiles");
        h->lockstep = (int)value;
    } else {
        rr_set_error("rr_index_set_param: unknown parameter '%s'", name);
        return RR_ERR_INVALID;
    }
    return RR_OK;
}

static int rr_flush_events(rr_index* h) {
    for (int i = 0; i + 1 < h->ev_used; i += 2) {
        RR_CUDA_CHECK(cudaEventSynchronize(h->ev_pool[i + 1]));
        float ms = 0.f;
        RR_CUDA_CHECK(cudaEventElapsedTime(&ms, h->ev_pool[i], h->ev_pool[i + 1]));
        h->stats.scan_ms_total += ms;
        h->stats.scan_launches += h->ev_launches[i / 2];
    }
    h->ev_used = 0;
    return RR_OK;
}

extern "C" int rr_index_get_stats(rr_index_t* h, rr_index_stats* out, int reset) {
    RR_REQUIRE(h != nullptr && out != nullptr, "rr_index_get_stats: NULL argument");
    rr_device_guard g(h->device);
    int rc = rr_flush_events(h);
    if (rc != RR_OK) return rc;
    unsigned long long ctr[4] = {0, 0, 0, 0};  // [0] low int = certificate failures, [1] bytes, [2] flops
    RR_CUDA_CHECK(cudaMemcpy(ctr, h->d_nfail_total, 32, cudaMemcpyDeviceToHost));
    h->stats.exact_rescans = (int)(ctr[0] & 0xFFFFFFFFull);
    h->stats.exact_scans = (int)(ctr[0] >> 32);
    h->stats.scan_bytes_total = (int64_t)ctr[1];
    h->stats.scan_flops_total = (int64_t)ctr[2];
    *out = h->stats;
    if (reset) {
        h->stats.scan_ms_total = 0;
        h->stats.scan_launches = 0;
        h->stats.scan_bytes_total = 0;
        h->stats.scan_flops_total = 0;
        RR_CUDA_CHECK(cudaMemset(h->d_nfail_total, 0, 32));  // failure counters too
    }
    return RR_OK;
}

static size_t rr_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Kernel dispatch (measured on B200, full and 1/8-scale MedRAG federation, scripts/gpu_pathcmp*.sh):
//   nq <= 24            swap-AB scan (rr_scan_tc.cu): lowest fixed cost, HBM-bound (7.1-7.3 TB/s)
//   24 < nq <= 128      queries-in-TMEM scan, one CTA per corpus slice: still HBM-bound, but its
//                       per-thread epilogue keeps up when shards are small (many survivors per tile)
//   nq > 128            decided on the device from the number of ROUTED queries (the host only knows
//                       the batch size): <= 128 routed -> as above; more -> CTA pairs (256 queries per
//                       pair; two 128-query single-CTA tiles would pull every corpus tile into two SMs)
// Dimensions whose 128-query tile does not fit TMEM + shared memory (d_pad > 1024) always take the
// swap-AB scan in passes of <= 32 / 16 queries.
static bool rr_use_gemm_path(const rr_index* h, int64_t nq) {
    if (!rr_gemm_scan_supported(h) || h->scan_path == RR_PATH_SMALL) return false;
    if (h->scan_path == RR_PATH_GEMM || h->scan_path == RR_PATH_GEMM_PAIR) return true;
    // (also for small shards: at 64 queries per pass the swap-AB scan's cooperative list warm-up costs more
    // than this kernel's fixed ~40 us -- measured 80 vs 57 us on a 37 k-row shard -- and mixing the two
    // paths in one step would split its batched launches)
    return nq > 24;
}
// 1 = one CTA per 128-query tile, 2 = CTA pairs, 0 = both are launched and the device runs the one that fits
// the number of ROUTED queries (<= 128: single CTAs, more: pairs)
static int rr_gemm_cta_group(const rr_index* h, int nq) {
    if (h->scan_path == RR_PATH_GEMM) return 1;
    if (h->scan_path == RR_PATH_GEMM_PAIR) return 2;
    return nq > RR_GEMM_QTILE ? 0 : 1;
}
static int rr_gemm_qalign(int cta_group) { return RR_GEMM_QTILE * (cta_group == 1 ? 1 : 2); }

// Everything one source needs for one chunk of <= RR_QCHUNK (RR_GEMM_QCHUNK on the large-batch
// path) queries: which scan it takes and where its scratch lives.
struct rr_chunk_plan {
    rr_index* h;
    const float* q;
    const uint8_t* mask;
    int64_t mask_stride;
    float* D;
    int64_t* I;
    int64_t out_stride;
    bool use_gemm, use_tc, use_pass2;
    int gemm_cg, cg2, nq_pad, nq_pad2, qt, g_exact, qsub;
    size_t ex_smem;
    int64_t cand_stride;
    rr_chunk_ctl *ctl, *ctl2;
    uint16_t *qbf, *qbf2;
    float *qnorm, *theta_q, *theta2, *cand_s;
    uint8_t *active, *fail_flags;
    int *fail_list, *cpos, *col_cnt, *fail2_list;
    uint32_t *col_rows, *cand_r;
};

static int rr_plan_chunk(rr_index* h, const float* q_dev, int nq, const uint8_t* mask_dev, int64_t mask_stride,
                         float* D_dev, int64_t* I_dev, int64_t out_stride, cudaStream_t stream, rr_chunk_plan* p) {
    p->h = h;
    p->q = q_dev;
    p->mask = mask_dev;
    p->mask_stride = mask_stride;
    p->D = D_dev;
    p->I = I_dev;
    p->out_stride = out_stride;
    p->qt = rr_tc_scan_qtile(h);  // 0: tc path unavailable for this d
    p->use_gemm = (h->mode == RR_MODE_AUTO) && h->n > 0 && rr_use_gemm_path(h, nq);
    p->use_tc = (h->mode == RR_MODE_AUTO) && p->qt > 0 && h->n > 0 && !p->use_gemm;
    p->gemm_cg = p->use_gemm ? rr_gemm_cta_group(h, nq) : 1;
    p->nq_pad = (int)rr_align_up((size_t)nq, p->use_gemm ? rr_gemm_qalign(p->gemm_cg) : RR_TC_MAX_QTILE);
    p->g_exact = h->sm_count * 2;
    p->cand_stride =
        p->use_gemm ? std::max<int64_t>((int64_t)p->g_exact * RR_EXACT_LIST_DEPTH,
                                        (int64_t)rr_gemm_scan_max_lists(h) * RR_GEMM_LIST_DEPTH)
                    : std::max<int64_t>((int64_t)p->g_exact * RR_EXACT_LIST_DEPTH,
                                        (int64_t)rr_tc_scan_num_lists(h) * RR_TC_LIST_DEPTH);
    // second pass (queries whose certificate failed), only when the first pass was a tensor-core scan
    p->use_pass2 = (p->use_tc || p->use_gemm) && rr_gemm_scan_supported(h);
    p->cg2 = rr_gemm_cta_group(h, nq);
    p->nq_pad2 = p->use_pass2 ? (int)rr_align_up((size_t)nq, (size_t)rr_gemm_qalign(p->cg2)) : 0;
    int qsub = RR_EXACT_MAX_QSUB;
    while (qsub > 1 && (size_t)qsub * h->d_pad * sizeof(float) > (96u << 10)) qsub >>= 1;
    p->qsub = qsub;
    p->ex_smem = (size_t)qsub * h->d_pad * sizeof(float) + (size_t)RR_EXACT_WARPS * qsub * RR_EXACT_LIST_DEPTH * 8 +
                 (size_t)RR_EXACT_WARPS * qsub * sizeof(float) + (size_t)qsub * sizeof(int);
    RR_REQUIRE(p->ex_smem <= (200u << 10), "dimension %d too large for the exact scan kernel", h->d);

    // carve the scratch block
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        size_t o = off;
        off = rr_align_up(off + bytes, 1024);
        return o;
    };
    const int nq_pad = p->nq_pad, nq_pad2 = p->nq_pad2;
    const size_t o_ctl = carve(sizeof(rr_chunk_ctl));
    const size_t o_qbf = carve((size_t)nq_pad * h->d_pad * sizeof(uint16_t));
    const size_t o_qn = carve((size_t)nq_pad * 2 * sizeof(float));
    const size_t o_act = carve((size_t)nq_pad);
    const size_t o_ff = carve((size_t)nq_pad);
    const size_t o_fl = carve((size_t)nq_pad * sizeof(int));
    const size_t o_cp = carve((size_t)nq_pad * sizeof(int));
    const size_t o_ctl2 = carve(sizeof(rr_chunk_ctl));
    const size_t o_qb2 = carve((size_t)nq_pad2 * h->d_pad * sizeof(uint16_t));
    const size_t o_thq = carve((size_t)nq_pad * sizeof(float));
    const size_t o_th2 = carve((size_t)nq_pad2 * sizeof(float));
    const size_t o_cc = carve((size_t)nq_pad2 * sizeof(int));
    const size_t o_f2 = carve((size_t)nq_pad * sizeof(int));
    const size_t o_cr2 = carve(p->use_pass2 ? (size_t)nq * RR_COLLECT_CAP * sizeof(uint32_t) : 0);
    const size_t o_cs = carve((size_t)nq * p->cand_stride * sizeof(float));
    const size_t o_cr = carve((size_t)nq * p->cand_stride * sizeof(uint32_t));
    if (off > h->scratch_bytes) {
        RR_CUDA_CHECK(cudaStreamSynchronize(stream));
        int rc = rr_grow(&h->scratch, &h->scratch_bytes, off);
        if (rc != RR_OK) return rc;
    }
    char* base = (char*)h->scratch;
    p->ctl = (rr_chunk_ctl*)(base + o_ctl);
    p->qbf = (uint16_t*)(base + o_qbf);
    p->qnorm = (float*)(base + o_qn);
    p->active = (uint8_t*)(base + o_act);
    p->fail_flags = (uint8_t*)(base + o_ff);
    p->fail_list = (int*)(base + o_fl);
    p->cpos = (int*)(base + o_cp);
    p->ctl2 = (rr_chunk_ctl*)(base + o_ctl2);
    p->qbf2 = (uint16_t*)(base + o_qb2);
    p->theta_q = (float*)(base + o_thq);
    p->theta2 = (float*)(base + o_th2);
    p->col_cnt = (int*)(base + o_cc);
    p->fail2_list = (int*)(base + o_f2);
    p->col_rows = (uint32_t*)(base + o_cr2);
    p->cand_s = (float*)(base + o_cs);
    p->cand_r = (uint32_t*)(base + o_cr);
    return RR_OK;
}

// never synchronise inside a search: a full event pool just stops recording
static int rr_prof_begin(rr_index* h, cudaStream_t stream, bool* on) {
    *on = h->profile != 0 && h->ev_used + 2 <= rr_index::kEvPool;
    if (!*on) return RR_OK;
    for (int i = 0; i < 2; ++i)
        if (!h->ev_pool[h->ev_used + i]) RR_CUDA_CHECK(cudaEventCreate(&h->ev_pool[h->ev_used + i]));
    RR_CUDA_CHECK(cudaEventRecord(h->ev_pool[h->ev_used], stream));
    return RR_OK;
}
static int rr_prof_end(rr_index* h, cudaStream_t stream, bool on, int launches) {
    if (!on) return RR_OK;
    RR_CUDA_CHECK(cudaEventRecord(h->ev_pool[h->ev_used + 1], stream));
    h->ev_launches[h->ev_used / 2] = launches;
    h->ev_used += 2;
    return RR_OK;
}

// One chunk of queries against S <= RR_FED_MAX sources (all on one device and stream): the scans are one
// launch per source, every other stage is one launch for all of them.
// `phases`: RR_PHASE_FIRST = begin, prep, the tensor-core scans and their finalize (results of every query
// whose certificate held are final after it, and ctl->n_fail counts the others); RR_PHASE_REPAIR = second
// pass, exact scan and their finalizes.  Device callers run both back to back (the repair launches exit at
// once when nothing failed); a host caller that synchronises anyway runs the second only when needed.
constexpr int RR_PHASE_FIRST = 1, RR_PHASE_REPAIR = 2, RR_PHASE_ALL = 3;

static int rr_run_chunk(rr_chunk_plan* plans, int S, int nq, int k, cudaStream_t stream, int phases = RR_PHASE_ALL) {
    const bool do_first = (phases & RR_PHASE_FIRST) != 0, do_repair = (phases & RR_PHASE_REPAIR) != 0;
    int max_nq_pad = 0, max_nq_pad2 = 0, max_dpad = 0;
    size_t max_ex_smem = 0;
    bool any_first = false, any_pass2 = false, any_rows = false;
    for (int s = 0; s < S; ++s) {
        const rr_chunk_plan& p = plans[s];
        max_nq_pad = std::max(max_nq_pad, p.nq_pad);
        max_nq_pad2 = std::max(max_nq_pad2, p.nq_pad2);
        max_dpad = std::max(max_dpad, p.h->d_pad);
        max_ex_smem = std::max(max_ex_smem, p.ex_smem);
        any_first |= p.use_tc || p.use_gemm;
        any_pass2 |= p.use_pass2;
        any_rows |= p.h->n > 0;
    }

    // finalize blocks: small batches are latency-bound (one block per (query, source) of 512 threads, all
    // candidate lists staged in shared memory); large batches are throughput-bound and their lists are few
    // (one per corpus slice): 256 threads and 16 KB of staging, so that 8 blocks share an SM
    const bool fin_small = (int64_t)nq * S <= 4 * (int64_t)plans[0].h->sm_count;
    const int fin_threads = fin_small ? RR_FIN_THREADS : RR_FIN_THREADS / 2;
    int64_t stage_cap = 2048;
    if (fin_small)
        for (int s = 0; s < S; ++s) stage_cap = std::max(stage_cap, plans[s].cand_stride);
    const int fin_blocks_sparse = std::min(nq, 64);  // launches that normally find nothing to do
    size_t fin_smem[2] = {0, 0};                     // [RESCORE]
    for (int s = 0; s < S; ++s)
        for (int r = 0; r < 2; ++r)
            fin_smem[r] = std::max(fin_smem[r], rr_finalize_smem(r != 0, plans[s].h->d_pad, stage_cap));
    RR_REQUIRE(fin_smem[1] <= (200u << 10), "dimension %d too large for the finalize kernel", max_dpad);
    int dev = 0;
    RR_CUDA_CHECK(cudaGetDevice(&dev));
    static std::atomic<bool> attr_done[RR_MAX_DEVICES];  // per device; a double set is harmless
    if (!attr_done[dev].load(std::memory_order_acquire)) {
        RR_CUDA_CHECK(cudaFuncSetAttribute(rr_exact_scan_kernel<RR_EXACT_MAX_QSUB>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
        RR_CUDA_CHECK(cudaFuncSetAttribute(rr_finalize_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
        RR_CUDA_CHECK(cudaFuncSetAttribute(rr_finalize_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
        RR_CUDA_CHECK(cudaFuncSetAttribute(rr_finalize_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
        RR_CUDA_CHECK(cudaFuncSetAttribute(rr_finalize_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 << 10));
        attr_done[dev].store(true, std::memory_order_release);
    }

    rr_multi<rr_prep_args> pa;
    memset(&pa, 0, sizeof(pa));
    for (int s = 0; s < S; ++s) {
        const rr_chunk_plan& p = plans[s];
        rr_prep_args& a = pa.s[s];
        a.q = p.q;
        a.mask = p.mask;
        a.mask_stride = p.mask_stride;
        a.qbf16 = p.qbf;
        a.qnorm = p.qnorm;
        a.active = p.active;
        a.fail_flags = p.fail_flags;
        a.fail_list = p.fail_list;
        a.ctl = p.ctl;
        a.cpos = p.cpos;
        a.nq = nq;
        a.nq_pad = p.nq_pad;
        a.d = p.h->d;
        a.d_pad = p.h->d_pad;
        a.exact_all = (p.use_tc || p.use_gemm) ? 0 : 1;
        a.compact = p.use_gemm ? 1 : 0;
    }
    if (do_first) {
        if (max_nq_pad <= RR_FUSED_PREP_MAX) {
            RR_CUDA_CHECK(rr_launch_pdl(rr_begin_prep_fused_kernel, dim3((unsigned)S), dim3(1024), 0, stream, pa));
            rr_count_launch(1);
        } else {
            RR_CUDA_CHECK(rr_launch_pdl(rr_begin_chunk_kernel, dim3((unsigned)S), dim3(1024), 0, stream, pa));
            RR_CUDA_CHECK(rr_launch_pdl(rr_prep_queries_kernel, dim3((unsigned)max_nq_pad, (unsigned)S), dim3(128), 0, stream, pa));
            rr_count_launch(2);
        }
    }

    rr_multi<rr_finalize_args> fa;
    memset(&fa, 0, sizeof(fa));
    for (int s = 0; s < S; ++s) {
        const rr_chunk_plan& p = plans[s];
        rr_index* h = p.h;
        rr_finalize_args& a = fa.s[s];
        a.cand_s = p.cand_s;
        a.cand_r = p.cand_r;
        a.cand_stride = p.cand_stride;
        a.k = k;
        a.kprime = std::max(h->kprime, k);
        a.q = p.q;
        a.qnorm = p.qnorm;
        a.nq_pad = p.nq_pad;
        a.hi = h->hi;
        a.lo = h->lo;
        a.d = h->d;
        a.d_pad = h->d_pad;
        a.max_norm_bits = h->max_norm_bits;
        a.id_offset = h->id_offset;
        a.D = p.D;
        a.I = p.I;
        a.out_stride = p.out_stride;
        a.active = p.active;
        a.fail_flags = p.fail_flags;
        a.fail_list = p.fail_list;
        a.ctl = p.ctl;
        a.nfail_total = h->d_nfail_total;
        a.pad_inactive = (p.use_tc || p.use_gemm) ? 0 : 1;
        a.theta_q = p.use_pass2 ? p.theta_q : nullptr;
        a.enabled = (p.use_tc || p.use_gemm) ? 1 : 0;
        a.nq = nq;
        a.stage_cap = (int)stage_cap;
    }

    // ---- first pass: bf16 tensor-core scan of every source, then one finalize for all of them.  The scans
    // are chained by programmatic dependent launch (they touch disjoint memory): no gap and no idle SMs
    // between them.  Profiling brackets the whole chain with one event pair.
    if (do_first) {
    rr_index* hprof = nullptr;
    for (int s = 0; s < S && !hprof; ++s)
        if ((plans[s].use_gemm || plans[s].use_tc) && plans[s].h->profile) hprof = plans[s].h;
    bool prof = false;
    if (hprof) {
        int rc = rr_prof_begin(hprof, stream, &prof);
        if (rc != RR_OK) return rc;
    }
    static const bool no_chain = getenv("RR_NO_SCAN_CHAIN") != nullptr;  // measurement knob
    int chain_len = 0;
    for (int s = 0; s < S; ++s) {
        const rr_chunk_plan& p = plans[s];
        rr_index* h = p.h;
        rr_finalize_args& a = fa.s[s];
        rr_scan_counters* ctr = prof ? reinterpret_cast<rr_scan_counters*>(h->d_nfail_total + 2) : nullptr;
        if (p.use_gemm) {
            // one launch covers the whole chunk: queries stationary in TMEM, corpus read once
            int rc = rr_gemm_scan_launch(h, p.qbf, nq, p.gemm_cg, p.ctl, p.cand_s, p.cand_r, p.cand_stride, ctr, nullptr,
                                         chain_len > 0 && !no_chain, stream);
            if (rc != RR_OK) return rc;
            rr_count_launch(p.gemm_cg == 0 ? 2 : 1);
            ++chain_len;  // (the device runs one of the two launches of gemm_cg == 0)
            a.n_lists = 0;
            a.lists_from_ctl = 1;
            a.cpos = p.cpos;
            a.depth = rr_gemm_scan_list_depth(h, nq);
        } else if (p.use_tc) {
            for (int p0 = 0; p0 < nq; p0 += p.qt) {
                const int nqt = std::min(p.qt, nq - p0);
                int n_lists = 0;
                int rc = rr_tc_scan_launch(h, p.qbf, p.nq_pad, p0, nqt, p.active + p0, &p.ctl->any_active,
                                           p.cand_s + (size_t)p0 * p.cand_stride, p.cand_r + (size_t)p0 * p.cand_stride,
                                           p.cand_stride, &n_lists, ctr, chain_len > 0 && !no_chain, stream);
                if (rc != RR_OK) return rc;
                rr_count_launch(1);
                ++chain_len;
                a.n_lists = n_lists;
                a.depth = RR_TC_LIST_DEPTH;
            }
        }
    }
    if (hprof) {
        int rc = rr_prof_end(hprof, stream, prof, chain_len);
        if (rc != RR_OK) return rc;
    }
    if (any_first) {
        const dim3 g1((unsigned)nq, (unsigned)S);
        if (fin_small) RR_CUDA_CHECK(rr_launch_pdl(rr_finalize_kernel<true, false>, g1, dim3(fin_threads), fin_smem[1], stream, fa));
        else RR_CUDA_CHECK(rr_launch_pdl(rr_finalize_kernel<true, true>, g1, dim3(fin_threads), fin_smem[1], stream, fa));
        rr_count_launch(1);
    }
    }  // do_first
    if (!do_repair) return RR_OK;

    // ---- second pass over the queries whose certificate failed: the tensor-core scan again, this time
    // reporting every row above a per-query threshold that provably keeps the whole exact top-k; then
    // exact scores of those rows.  All launches exit at once when nothing failed.
    if (any_pass2) {
        rr_multi<rr_prep_collect_args> pc;
        rr_multi<rr_collect_fin_args> cf;
        memset(&pc, 0, sizeof(pc));
        memset(&cf, 0, sizeof(cf));
        for (int s = 0; s < S; ++s) {
            const rr_chunk_plan& p = plans[s];
            if (!p.use_pass2) continue;
            rr_index* h = p.h;
            rr_prep_collect_args& a = pc.s[s];
            a.qbf16 = p.qbf;
            a.cpos = p.use_gemm ? p.cpos : nullptr;
            a.fail_list = p.fail_list;
            a.theta_q = p.theta_q;
            a.ctl = p.ctl;
            a.qbf2 = p.qbf2;
            a.theta2 = p.theta2;
            a.col_cnt = p.col_cnt;
            a.ctl2 = p.ctl2;
            a.d_pad = h->d_pad;
            a.nq = nq;
            a.nq_pad2 = p.nq_pad2;
            rr_collect_fin_args& c = cf.s[s];
            c.col_rows = p.col_rows;
            c.col_cnt = p.col_cnt;
            c.fail_list = p.fail_list;
            c.ctl = p.ctl;
            c.fail2_list = p.fail2_list;
            c.fail_flags = p.fail_flags;
            c.n_exact_total = h->d_nfail_total + 1;
            c.k = k;
            c.q = p.q;
            c.hi = h->hi;
            c.lo = h->lo;
            c.d = h->d;
            c.d_pad = h->d_pad;
            c.id_offset = h->id_offset;
            c.D = p.D;
            c.I = p.I;
            c.out_stride = p.out_stride;
            c.enabled = 1;
        }
        RR_CUDA_CHECK(rr_launch_pdl(rr_prep_collect_kernel, dim3((unsigned)std::min(max_nq_pad2, 256), (unsigned)S), dim3(128), 0,
                                    stream, pc));
        rr_count_launch(1);
        int n_collect = 0;
        for (int s = 0; s < S; ++s) {
            const rr_chunk_plan& p = plans[s];
            if (!p.use_pass2) continue;
            rr_collect_args ca;
            ca.theta = p.theta2;
            ca.rows = p.col_rows;
            ca.cnt = p.col_cnt;
            ca.cap = RR_COLLECT_CAP;
            int rc = rr_gemm_scan_launch(p.h, p.qbf2, nq, p.cg2, p.ctl2, nullptr, nullptr, 0, nullptr, &ca, n_collect > 0,
                                         stream);
            if (rc != RR_OK) return rc;
            ++n_collect;
            rr_count_launch(p.cg2 == 0 ? 2 : 1);
        }
        RR_CUDA_CHECK(rr_launch_pdl(rr_finalize_collect_kernel, dim3((unsigned)fin_blocks_sparse, (unsigned)S),
                                    dim3(RR_FIN_THREADS), (size_t)max_dpad * sizeof(float), stream, cf));
        rr_count_launch(1);
    }

    // ---- exact path over what is left (everything when there was no tensor-core pass); launches exit
    // at once when empty
    if (any_rows) {
        rr_multi<rr_exact_args> ea;
        memset(&ea, 0, sizeof(ea));
        int min_qsub = RR_EXACT_MAX_QSUB;
        for (int s = 0; s < S; ++s) {
            const rr_chunk_plan& p = plans[s];
            rr_exact_args& a = ea.s[s];
            a.hi = p.h->hi;
            a.lo = p.h->lo;
            a.n_rows = p.h->n;
            a.q = p.q;
            a.fail_list = p.use_pass2 ? p.fail2_list : p.fail_list;
            a.n_fail_ptr = p.use_pass2 ? &p.ctl->n_fail2 : &p.ctl->n_fail;
            a.cand_s = p.cand_s;
            a.cand_r = p.cand_r;
            a.cand_stride = p.cand_stride;
            a.d = p.h->d;
            a.d_pad = p.h->d_pad;
            a.qsub = p.qsub;
            min_qsub = std::min(min_qsub, p.qsub);
        }
        // one launch covers every sub-batch of every source (the kernel strides over them); blocks exit at
        // once when nothing is left for the exact scan -- the normal case after a tensor-core pass, so only
        // exact-only plans get extra rows of blocks
        const int g_exact = plans[0].g_exact;
        const int gy = any_first ? 1 : std::min(4, (nq + min_qsub - 1) / min_qsub);
        const dim3 grid((unsigned)g_exact, (unsigned)gy, (unsigned)S);
        RR_CUDA_CHECK(rr_launch_pdl(rr_exact_scan_kernel<RR_EXACT_MAX_QSUB>, grid, dim3(RR_EXACT_WARPS * 32), max_ex_smem, stream, ea));
        rr_count_launch(1);
    }
    // after a tensor-core pass only the queries the second pass could not settle are left (normally none):
    // a few blocks stride over that list; an exact-only plan visits every query (and pads the unrouted ones)
    bool all_sparse = true;
    for (int s = 0; s < S; ++s) {
        const rr_chunk_plan& p = plans[s];
        rr_finalize_args& a = fa.s[s];
        a.n_lists = (p.h->n > 0) ? p.g_exact : 0;
        a.depth = RR_EXACT_LIST_DEPTH;
        a.lists_from_ctl = 0;
        a.cpos = nullptr;
        a.enabled = 1;
        if (p.use_tc || p.use_gemm) {
            a.work_list = p.use_pass2 ? p.fail2_list : p.fail_list;
            a.work_count = p.use_pass2 ? &p.ctl->n_fail2 : &p.ctl->n_fail;
        } else {
            a.work_list = nullptr;
            a.work_count = nullptr;
            all_sparse = false;
        }
    }
    const dim3 fgrid((unsigned)(all_sparse ? fin_blocks_sparse : nq), (unsigned)S);
    if (fin_small) RR_CUDA_CHECK(rr_launch_pdl(rr_finalize_kernel<false, false>, fgrid, dim3(fin_threads), fin_smem[0], stream, fa));
    else RR_CUDA_CHECK(rr_launch_pdl(rr_finalize_kernel<false, true>, fgrid, dim3(fin_threads), fin_smem[0], stream, fa));
    rr_count_launch(1);
    return RR_OK;
}

// Sources that hold no rows on this device (whole-source placement: a rank passes an empty index for the corpora
// that live elsewhere) answer with padding rows -- one launch for all of them instead of a pipeline each.
struct rr_pad_args {
    float* D;
    int64_t* I;
    int64_t nq;
    int64_t out_stride;
    int k;
    int n_sources;
    unsigned long long empty[4];  // bit s: source s is empty
};
__global__ void __launch_bounds__(256) rr_fill_padding_kernel(const rr_pad_args a) {
    rr_pdl_enter();
    const int cols = a.n_sources * a.k;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.nq * cols; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = i / cols;
        const int c = (int)(i - b * cols), s = c / a.k;
        if ((a.empty[s >> 6] >> (s & 63)) & 1ull) {
            a.D[b * a.out_stride + c] = -3.402823466e+38f;
            a.I[b * a.out_stride + c] = -1;
        }
    }
}

static int64_t rr_chunk_size(const rr_index* h, int64_t nq) {
    return (h->mode == RR_MODE_AUTO && rr_use_gemm_path(h, nq)) ? RR_GEMM_QCHUNK : RR_QCHUNK;
}

extern "C" int rr_federation_search(rr_index_t* const* idx, int n_sources, const float* const* q_dev, int64_t nq,
                                    int k, const uint8_t* mask_dev, int64_t mask_stride, float* D_dev,
                                    int64_t* I_dev, int64_t out_stride, void* stream_) {
    RR_REQUIRE(idx != nullptr && n_sources >= 1, "rr_federation_search: no sources");
    RR_REQUIRE(nq >= 0, "rr_federation_search: negative nq");
    RR_REQUIRE(k > 0, "rr_federation_search: k must be > 0 (got %d)", k);
    if (k > RR_MAX_K) {
        rr_set_error("rr_federation_search: k=%d > RR_MAX_K=%d is not implemented", k, RR_MAX_K);
        return RR_ERR_UNSUPPORTED;
    }
    for (int s = 0; s < n_sources; ++s) {
        RR_REQUIRE(idx[s] != nullptr, "rr_federation_search: source %d is NULL", s);
        RR_REQUIRE(idx[s]->device == idx[0]->device, "rr_federation_search: sources live on different devices");
        for (int t = 0; t < s; ++t)
            RR_REQUIRE(idx[t] != idx[s], "rr_federation_search: source %d listed twice (handles are not re-entrant)", s);
    }
    if (nq == 0) return RR_OK;
    RR_REQUIRE(q_dev && D_dev && I_dev, "rr_federation_search: NULL buffer");
    for (int s = 0; s < n_sources; ++s) RR_REQUIRE(q_dev[s] != nullptr, "rr_federation_search: NULL queries for source %d", s);
    RR_REQUIRE(mask_dev == nullptr || mask_stride >= n_sources, "rr_federation_search: mask stride < n_sources");
    if (out_stride == 0) out_stride = (int64_t)n_sources * k;
    RR_REQUIRE(out_stride >= (int64_t)n_sources * k, "rr_federation_search: out_stride < n_sources * k");
    rr_device_guard g(idx[0]->device);
    cudaStream_t stream = (cudaStream_t)stream_;

    // sources that chunk the batch alike share their launches, RR_FED_MAX at a time
    bool done[256] = {};
    RR_REQUIRE(n_sources <= 256, "rr_federation_search: more than 256 sources");
    {
        rr_pad_args pa = {};
        int n_empty = 0;
        for (int s = 0; s < n_sources; ++s)
            if (idx[s]->n == 0) {
                done[s] = true;
                pa.empty[s >> 6] |= 1ull << (s & 63);
                ++n_empty;
            }
        if (n_empty > 0 && n_sources > 1) {
            pa.D = D_dev;
            pa.I = I_dev;
            pa.nq = nq;
            pa.out_stride = out_stride;
            pa.k = k;
            pa.n_sources = n_sources;
            const int64_t total = nq * (int64_t)n_sources * k;
            const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, (int64_t)idx[0]->sm_count * 8);
            RR_CUDA_CHECK(rr_launch_pdl(rr_fill_padding_kernel, dim3(blocks), dim3(256), 0, stream, pa));
            rr_count_launch(1);
        } else if (n_empty > 0) {
            done[0] = false;  // a lone empty index keeps the plain path (index.search on an empty IndexFlatIP)
        }
    }
    for (int s0 = 0; s0 < n_sources; ++s0) {
        if (done[s0]) continue;
        const int64_t chunk = rr_chunk_size(idx[s0], nq);
        int group[RR_FED_MAX], G = 0;
        for (int s = s0; s < n_sources && G < RR_FED_MAX; ++s)
            if (!done[s] && rr_chunk_size(idx[s], nq) == chunk) {
                group[G++] = s;
                done[s] = true;
            }
        for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
            const int nqc = (int)std::min<int64_t>(chunk, nq - q0);
            rr_chunk_plan plans[RR_FED_MAX];
            for (int i = 0; i < G; ++i) {
                const int s = group[i];
                rr_index* h = idx[s];
                int rc = rr_plan_chunk(h, q_dev[s] + (size_t)q0 * h->d, nqc,
                                       mask_dev ? mask_dev + q0 * mask_stride + s : nullptr, mask_stride,
                                       D_dev + (size_t)q0 * out_stride + (size_t)s * k,
                                       I_dev + (size_t)q0 * out_stride + (size_t)s * k, out_stride, stream, &plans[i]);
                if (rc != RR_OK) return rc;
            }
            int rc = rr_run_chunk(plans, G, nqc, k, stream);
            if (rc != RR_OK) return rc;
        }
    }
    for (int s = 0; s < n_sources; ++s) {
        idx[s]->stats.searches += 1;
        idx[s]->stats.queries += nq;
        idx[s]->dev_dirty = true;
    }
    return RR_OK;
}

extern "C" int rr_index_search_ex(rr_index_t* h, const float* q_dev, int64_t nq, int k,
                                  const uint8_t* mask_dev, int64_t mask_stride, float* D_dev,
                                  int64_t* I_dev, int64_t out_stride, void* stream_) {
    RR_REQUIRE(h != nullptr, "rr_index_search: NULL handle");
    RR_REQUIRE(nq >= 0, "rr_index_search: negative nq");
    RR_REQUIRE(k > 0, "rr_index_search: k must be > 0 (got %d)", k);
    if (k > RR_MAX_K) {
        rr_set_error("rr_index_search: k=%d > RR_MAX_K=%d is not implemented", k, RR_MAX_K);
        return RR_ERR_UNSUPPORTED;
    }
    if (nq == 0) return RR_OK;
    RR_REQUIRE(q_dev && D_dev && I_dev, "rr_index_search: NULL buffer");
    RR_REQUIRE(mask_dev == nullptr || mask_stride >= 1, "rr_index_search: bad mask stride");
    if (out_stride == 0) out_stride = k;
    RR_REQUIRE(out_stride >= k, "rr_index_search: out_stride < k");
    rr_device_guard g(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t chunk = rr_chunk_size(h, nq);
    for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
        const int nqc = (int)std::min<int64_t>(chunk, nq - q0);
        rr_chunk_plan plan;
        int rc = rr_plan_chunk(h, q_dev + (size_t)q0 * h->d, nqc, mask_dev ? mask_dev + q0 * mask_stride : nullptr,
                               mask_stride, D_dev + (size_t)q0 * out_stride, I_dev + (size_t)q0 * out_stride,
                               out_stride, stream, &plan);
        if (rc != RR_OK) return rc;
        rc = rr_run_chunk(&plan, 1, nqc, k, stream);
        if (rc != RR_OK) return rc;
    }
    h->stats.searches += 1;
    h->stats.queries += nq;
    h->dev_dirty = true;
    return RR_OK;
}

extern "C" int rr_index_search_masked(rr_index_t* h, const float* q_dev, int64_t nq, int k,
                                      const uint8_t* mask_dev, int64_t mask_stride, float* D_dev,
                                      int64_t* I_dev, void* stream) {
    return rr_index_search_ex(h, q_dev, nq, k, mask_dev, mask_stride, D_dev, I_dev, 0, stream);
}

extern "C" int rr_index_search(rr_index_t* h, const float* q_dev, int64_t nq, int k, float* D_dev,
                               int64_t* I_dev, void* stream) {
    return rr_index_search_ex(h, q_dev, nq, k, nullptr, 0, D_dev, I_dev, 0, stream);
}

// The reference's calling convention (numpy in, numpy out, one query at a time: data_source.py:113-120,
// 186,199-203).  Runs on the handle's own non-blocking stream, so it does not serialise against the other
// streams of the process; rows added asynchronously on another stream are ordered before it by the event
// rr_index_add_f32 records.
extern "C" int rr_index_search_host_ex(rr_index_t* h, const float* q_host, int64_t nq, int k, int flags,
                                       float* D_host, int64_t* I_host, float* q_out_host) {
    RR_REQUIRE(h != nullptr, "rr_index_search_host: NULL handle");
    RR_REQUIRE(nq >= 0 && k > 0, "rr_index_search_host: bad nq / k");
    RR_REQUIRE((flags & ~RR_SEARCH_NORMALIZE_L2) == 0, "rr_index_search_host: unknown flags 0x%x", flags);
    if (k > RR_MAX_K) {
        rr_set_error("rr_index_search: k=%d > RR_MAX_K=%d is not implemented", k, RR_MAX_K);
        return RR_ERR_UNSUPPORTED;
    }
    if (nq == 0) return RR_OK;
    RR_REQUIRE(q_host && D_host && I_host, "rr_index_search_host: NULL buffer");
    rr_device_guard g(h->device);
    const bool want_q = q_out_host != nullptr;
    // device staging: [ queries | D | I | control block ]; the last three (all four when the normalised
    // queries are wanted back) leave in ONE copy
    const size_t qb = rr_align_up((size_t)nq * h->d * sizeof(float), 256);
    const size_t db = rr_align_up((size_t)nq * k * sizeof(float), 256);
    const size_t ib = rr_align_up((size_t)nq * k * sizeof(int64_t), 256);
    const size_t out_b = db + ib + sizeof(rr_chunk_ctl);
    const size_t pin_b = qb + out_b;
    cudaStream_t stream = h->host_stream;
    if (h->stage_bytes < qb + out_b) RR_CUDA_CHECK(cudaStreamSynchronize(stream));
    int rc = rr_grow((void**)&h->stage, &h->stage_bytes, qb + out_b);
    if (rc != RR_OK) return rc;
    if (h->pin_bytes < pin_b) {
        if (h->pin) cudaFreeHost(h->pin);
        h->pin = nullptr;
        h->pin_bytes = 0;
        RR_CUDA_CHECK(cudaHostAlloc(&h->pin, pin_b + pin_b / 4, cudaHostAllocDefault));
        h->pin_bytes = pin_b + pin_b / 4;
    }
    char* base = (char*)h->stage;
    float* q_dev = (float*)base;
    float* D_dev = (float*)(base + qb);
    int64_t* I_dev = (int64_t*)(base + qb + db);
    rr_chunk_ctl* ctl_dev = (rr_chunk_ctl*)(base + qb + db + ib);
    // what comes back: [queries |] D | I | ctl, landing at the same offsets of the pinned mirror
    const char* back_dev = want_q ? base : base + qb;
    char* back_pin = (char*)h->pin + (want_q ? 0 : qb);
    const size_t back_b = want_q ? qb + out_b : out_b;
    const char* pin_out = (const char*)h->pin + qb;
    if (h->add_event_pending) {
        RR_CUDA_CHECK(cudaStreamWaitEvent(stream, h->add_event, 0));
        h->add_event_pending = false;
    }
    if (h->dev_dirty) {  // an asynchronous device-API search may still use the handle's scratch block
        RR_CUDA_CHECK(cudaDeviceSynchronize());
        h->dev_dirty = false;
    }
    RR_CUDA_CHECK(cudaMemcpyAsync(q_dev, q_host, (size_t)nq * h->d * sizeof(float), cudaMemcpyHostToDevice, stream));
    if (flags & RR_SEARCH_NORMALIZE_L2) {
        rc = rr_normalize_l2(q_dev, nq, h->d, stream);
        if (rc != RR_OK) return rc;
    }
    if (nq <= rr_chunk_size(h, nq)) {
        // One chunk -- the reference's call (nq = 1, data_source.py:186).  This caller waits for its result
        // anyway, so it runs the first pass, looks at the count of failed certificates that comes back with
        // the results, and launches the repair stages only when there is something to repair.
        rr_chunk_plan plan;
        rc = rr_plan_chunk(h, q_dev, (int)nq, nullptr, 0, D_dev, I_dev, k, stream, &plan);
        if (rc != RR_OK) return rc;
        plan.ctl = ctl_dev;
        for (int phase = RR_PHASE_FIRST; phase <= RR_PHASE_REPAIR; ++phase) {
            rc = rr_run_chunk(&plan, 1, (int)nq, k, stream, phase);
            if (rc != RR_OK) return rc;
            RR_CUDA_CHECK(cudaMemcpyAsync(back_pin, back_dev, back_b, cudaMemcpyDeviceToHost, stream));
            RR_CUDA_CHECK(cudaStreamSynchronize(stream));
            const rr_chunk_ctl* ctl_host = (const rr_chunk_ctl*)(pin_out + db + ib);
            if (ctl_host->n_fail == 0) break;  // every certificate held: the results are final
        }
        h->stats.searches += 1;
        h->stats.queries += nq;
    } else {
        rc = rr_index_search(h, q_dev, nq, k, D_dev, I_dev, stream);
        if (rc != RR_OK) return rc;
        RR_CUDA_CHECK(cudaMemcpyAsync(back_pin, back_dev, back_b - sizeof(rr_chunk_ctl), cudaMemcpyDeviceToHost, stream));
        RR_CUDA_CHECK(cudaStreamSynchronize(stream));
    }
    memcpy(D_host, pin_out, (size_t)nq * k * sizeof(float));
    memcpy(I_host, pin_out + db, (size_t)nq * k * sizeof(int64_t));
    if (want_q) memcpy(q_out_host, h->pin, (size_t)nq * h->d * sizeof(float));
    return RR_OK;
}

extern "C" int rr_index_search_host(rr_index_t* h, const float* q_host, int64_t nq, int k,
                                    float* D_host, int64_t* I_host) {
    return rr_index_search_host_ex(h, q_host, nq, k, 0, D_host, I_host, nullptr);
}

extern "C" int rr_index_tc_qtile(const rr_index_t* h) { return h ? rr_tc_scan_qtile(h) : RR_ERR_INVALID; }

extern "C" int rr_index_debug_tc_scan(rr_index_t* h, const float* q_dev, int nq, float* cand_s_dev,
                                      uint32_t* cand_r_dev, int64_t capacity, int* n_lists, int* depth,
                                      void* stream_) {
    RR_REQUIRE(h && q_dev && cand_s_dev && cand_r_dev && n_lists && depth, "rr_index_debug_tc_scan: NULL argument");
    const int qt = rr_tc_scan_qtile(h);
    RR_REQUIRE(qt > 0 && nq >= 1 && nq <= qt, "rr_index_debug_tc_scan: nq must be in [1, %d]", qt);
    RR_REQUIRE(h->n > 0, "rr_index_debug_tc_scan: empty index");
    rr_device_guard g(h->device);
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t stride = (int64_t)rr_tc_scan_num_lists(h) * RR_TC_LIST_DEPTH;
    RR_REQUIRE(capacity >= stride, "rr_index_debug_tc_scan: capacity %lld < %lld entries per query",
               (long long)capacity, (long long)stride);
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        size_t o = off;
        off = rr_align_up(off + bytes, 1024);
        return o;
    };
    const int nq_pad = RR_TC_MAX_QTILE;
    const size_t o_ctl = carve(sizeof(rr_chunk_ctl));
    const size_t o_qbf = carve((size_t)nq_pad * h->d_pad * sizeof(uint16_t));
    const size_t o_qn = carve((size_t)nq_pad * 2 * sizeof(float));
    const size_t o_act = carve((size_t)nq_pad);
    const size_t o_ff = carve((size_t)nq_pad);
    const size_t o_fl = carve((size_t)nq_pad * sizeof(int));
    if (off > h->scratch_bytes) {
        RR_CUDA_CHECK(cudaStreamSynchronize(stream));
        int rc = rr_grow(&h->scratch, &h->scratch_bytes, off);
        if (rc != RR_OK) return rc;
    }
    char* base = (char*)h->scratch;
    rr_chunk_ctl* ctl = (rr_chunk_ctl*)(base + o_ctl);
    RR_CUDA_CHECK(cudaMemsetAsync(ctl, 0, sizeof(rr_chunk_ctl), stream));
    rr_multi<rr_prep_args> pa;
    memset(&pa, 0, sizeof(pa));
    pa.s[0].q = q_dev;
    pa.s[0].qbf16 = (uint16_t*)(base + o_qbf);
    pa.s[0].qnorm = (float*)(base + o_qn);
    pa.s[0].active = (uint8_t*)(base + o_act);
    pa.s[0].fail_flags = (uint8_t*)(base + o_ff);
    pa.s[0].fail_list = (int*)(base + o_fl);
    pa.s[0].ctl = ctl;
    pa.s[0].nq = nq;
    pa.s[0].nq_pad = nq_pad;
    pa.s[0].d = h->d;
    pa.s[0].d_pad = h->d_pad;
    rr_prep_queries_kernel<<<dim3((unsigned)nq_pad, 1), 128, 0, stream>>>(pa);
    RR_CUDA_CHECK(cudaGetLastError());
    RR_CUDA_CHECK(cudaMemsetAsync(cand_s_dev, 0xFF, (size_t)nq * capacity * sizeof(float), stream));
    RR_CUDA_CHECK(cudaMemsetAsync(cand_r_dev, 0xFF, (size_t)nq * capacity * sizeof(uint32_t), stream));
    int rc = rr_tc_scan_launch(h, (uint16_t*)(base + o_qbf), nq_pad, 0, nq, (uint8_t*)(base + o_act),
                               &ctl->any_active, cand_s_dev, cand_r_dev, capacity, n_lists, nullptr, false, stream);
    if (rc != RR_OK) return rc;
    *depth = RR_TC_LIST_DEPTH;
    return RR_OK;
}

extern "C" int rr_normalize_l2(float* x_dev, int64_t n, int d, void* stream) {
    RR_REQUIRE(n >= 0 && d > 0, "rr_normalize_l2: bad shape");
    if (n == 0) return RR_OK;
    RR_REQUIRE(x_dev != nullptr, "rr_normalize_l2: NULL buffer");
    RR_REQUIRE(n <= 0x7FFFFFFF, "rr_normalize_l2: too many rows");
    RR_CUDA_CHECK(rr_launch_pdl(rr_normalize_l2_kernel, dim3((unsigned)n), dim3(128), 0, (cudaStream_t)stream, x_dev, n, d));
    return RR_OK;
}
