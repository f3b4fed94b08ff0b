// B2: router relevance MLP (/root/reference/ragroute/router.py:37-55 CorpusRoutingNN,
// :241-283 select_relevant_sources_ragroute) as ONE kernel: feature scaling, fc1,
// LayerNorm, ReLU, fc2, LayerNorm, ReLU, fc3, sigmoid and threshold for every
// (query, source) pair of a batch, nothing but the mask / logits written to HBM.
//
// The feature row of a pair is [pad(q) | centroid_s | onehot_s] (router.py:253-265);
// only the first d_max columns depend on the query, so fc1 splits into
//   W1[:, :d_max] . xq           shared by all sources that use the same encoder
// + c1_s = b1 + W1[:, d_max:] . xconst_s   folded once per source at load time.
// The scaler is applied exactly as the reference does: (x - mean_) / scale_ in
// float64, then cast to fp32 (router.py:270-271).
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <new>
#include <vector>

#include "rr_common.cuh"

namespace {
constexpr int kH1 = 256;
constexpr int kH2 = 128;
constexpr int kQB = 4;          // queries (or rows) per CTA
constexpr int kThreads = 1024;  // fc1: 4 K-quarters x 256 outputs; fc2: 8 K-slices x 128 outputs
constexpr int kWarps = kThreads / 32;
constexpr int kKQ = kThreads / kH1;   // K-quarters of fc1
constexpr int kKS = kThreads / kH2;   // K-slices of fc2
constexpr int kPG = 16;               // (query, source) pairs per fc2 round
constexpr int kHStride = kPG + 4;     // h1t row pitch: float4-aligned, spreads the transposing stores over banks
constexpr int kInFlight = 16;         // independent weight loads per thread (the loops are L2-latency-bound)
constexpr float kLnEps = 1e-5f;
// fc1 is summed in a FIXED tree whatever the launch shape, so that a query's logits do not depend on the batch
// it arrives in: the input dimensions are cut into kKE x kKQ contiguous slices, each accumulated in ascending
// order; per K-quarter kq the slices e = 0 .. kKE-1 are added in ascending order, then the quarters.
// Small batches run one thread-block CLUSTER of kKE CTAs per kQB queries, CTA e taking slices (e, *): one SM
// pulling fc1's 786 KB through a single L1 was 2/3 of the kernel's 33 us at batch 1 (ncu, round 2).
constexpr int kKE = 8;
constexpr int kClusterMaxBatch = 128;  // above this the grid fills the SMs without clusters
}  // namespace

struct rr_router {
    int device = 0;
    int in_dim = 0;
    // device weights
    float* w1t = nullptr;  // [in_dim][256]  (fc1.weight transposed)
    float* b1 = nullptr;
    float* ln1w = nullptr;
    float* ln1b = nullptr;
    float* w2t = nullptr;  // [256][128]
    float* b2 = nullptr;
    float* ln2w = nullptr;
    float* ln2b = nullptr;
    float* w3 = nullptr;  // [128]
    float* b3 = nullptr;  // [1]
    std::vector<float> h_w1;  // host copy of fc1.weight [256][in_dim] for table folding
    std::vector<float> h_b1;
    // select tables
    int n_sources = 0, d_max = 0, n_slots = 0;
    int* slot_of_source = nullptr;  // device [S]
    float* c1 = nullptr;            // device [S][256]
    double* sc_mean = nullptr;      // device [d_max] or null
    double* sc_scale = nullptr;
};

namespace {

struct RouterArgs {
    const float* w1t;
    const float* ln1w;
    const float* ln1b;
    const float* w2t;
    const float* b2;
    const float* ln2w;
    const float* ln2b;
    const float* w3;
    const float* b3;
    const float* c1;          // [S][256] per-source fc1 constant (logits mode: b1, S = 1)
    const int* slot_of_source;  // [S] or null (all 0)
    const double* sc_mean;    // [d_in] or null
    const double* sc_scale;
    const float* x;           // [B][n_slots][d_in]
    int64_t B;
    int S;
    int n_slots;
    int d_in;                 // columns of x consumed by fc1 (d_max, or in_dim in logits mode)
    float prob_thr;
    uint8_t* mask;            // [B][S] or null
    float* logits;            // [B][S] or null
};

__device__ __forceinline__ uint32_t router_cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void router_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// store into the same shared-memory variable of the cluster's CTA 0
__device__ __forceinline__ void router_store_leader(float* p, float v) {
    uint32_t local = (uint32_t)__cvta_generic_to_shared(p), remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(0));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

// Shared memory: [ xq | part ] is dead once fc1 is done and is reused for the fc2 partial sums.
// CL = 1: one CTA per kQB queries.  CL = kKE: a cluster per kQB queries; CTA e computes the fc1 slices (e, *)
// and ships them to CTA 0 (distributed shared memory), which finishes the network alone.
template <int CL>
__global__ void __launch_bounds__(kThreads, 1) rr_router_kernel(const RouterArgs a) {
    rr_pdl_enter();  // see rr_common.cuh: the step's kernels overlap their launch gaps
    if (CL > 1) router_cluster_sync();  // every CTA of the cluster runs before its shared memory is written remotely
    extern __shared__ __align__(16) float sm[];
    const size_t n_x = ((size_t)kQB * a.d_in + 3) & ~(size_t)3;
    const size_t n_p1 = (CL > 1) ? (size_t)kKE * kKQ * kQB * kH1 : (size_t)kKQ * kQB * kH1;
    const size_t n_u = max(n_x + n_p1, (size_t)kKS * kPG * kH2);
    float* xq = sm;                                   // [kQB][d_in]
    float* part = xq + n_x;                           // [kKQ][kQB][256]  (CL > 1, leader: [kKE][kKQ][kQB][256])
    float* part2 = sm;                                // [kKS][kPG][128]   (aliases xq / part)
    float* g = sm + n_u;                              // [n_slots][kQB][256]
    float* h1t = g + (size_t)a.n_slots * kQB * kH1;   // [256][kHStride]: relu(ln1(fc1)) of kPG pairs, transposed

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int crank = (CL > 1) ? (int)router_cluster_rank() : 0;
    const int64_t b0 = (int64_t)(blockIdx.x / CL) * kQB;
    const int nqb = (int)min((int64_t)kQB, a.B - b0);

    // ---- phase 1: g[slot][qb][:] = W1[:, :d_in] . scaled(x[qb][slot][:]) ----
    const int t = tid & (kH1 - 1);
    const int kq = tid >> 8;
    for (int sl = 0; sl < a.n_slots; ++sl) {
        for (int i = tid; i < kQB * a.d_in; i += kThreads) {
            const int qb = i / a.d_in, j = i - qb * a.d_in;
            float v = 0.f;
            if (qb < nqb) {
                v = a.x[((b0 + qb) * a.n_slots + sl) * (int64_t)a.d_in + j];
                if (a.sc_mean) v = (float)(((double)v - a.sc_mean[j]) / a.sc_scale[j]);
            }
            xq[i] = v;
        }
        __syncthreads();
        float tot[kQB];
#pragma unroll
        for (int qb = 0; qb < kQB; ++qb) tot[qb] = 0.f;
        const float* wp = a.w1t + t;
        for (int e = (CL > 1 ? crank : 0); e < (CL > 1 ? crank + 1 : kKE); ++e) {
            const int slice = e * kKQ + kq;
            const int j_lo = (int)((int64_t)slice * a.d_in / (kKE * kKQ));
            const int j_hi = (int)((int64_t)(slice + 1) * a.d_in / (kKE * kKQ));
            float acc[kQB];
#pragma unroll
            for (int qb = 0; qb < kQB; ++qb) acc[qb] = 0.f;
            int j = j_lo;
            for (; j + kInFlight <= j_hi; j += kInFlight) {
                float wv[kInFlight];
#pragma unroll
                for (int u = 0; u < kInFlight; ++u) wv[u] = __ldg(wp + (size_t)(j + u) * kH1);
#pragma unroll
                for (int qb = 0; qb < kQB; ++qb) {
                    const float* xr = xq + (size_t)qb * a.d_in + j;
#pragma unroll
                    for (int u = 0; u < kInFlight; ++u) acc[qb] = fmaf(wv[u], xr[u], acc[qb]);
                }
            }
            if (j + kInFlight / 2 <= j_hi) {
                float wv[kInFlight / 2];
#pragma unroll
                for (int u = 0; u < kInFlight / 2; ++u) wv[u] = __ldg(wp + (size_t)(j + u) * kH1);
#pragma unroll
                for (int qb = 0; qb < kQB; ++qb) {
                    const float* xr = xq + (size_t)qb * a.d_in + j;
#pragma unroll
                    for (int u = 0; u < kInFlight / 2; ++u) acc[qb] = fmaf(wv[u], xr[u], acc[qb]);
                }
                j += kInFlight / 2;
            }
            for (; j < j_hi; ++j) {
                const float w0 = __ldg(wp + (size_t)j * kH1);
#pragma unroll
                for (int qb = 0; qb < kQB; ++qb) acc[qb] = fmaf(w0, xq[(size_t)qb * a.d_in + j], acc[qb]);
            }
            if (CL > 1) {
#pragma unroll
                for (int qb = 0; qb < kQB; ++qb)
                    router_store_leader(part + ((size_t)(e * kKQ + kq) * kQB + qb) * kH1 + t, acc[qb]);
            } else {
#pragma unroll
                for (int qb = 0; qb < kQB; ++qb) tot[qb] += acc[qb];
            }
        }
        if (CL > 1) {
            router_cluster_sync();  // every CTA's slices have landed in the leader's shared memory
            if (crank == 0) {
                for (int i = tid; i < kQB * kH1; i += kThreads) {
                    float v = 0.f;
#pragma unroll
                    for (int c = 0; c < kKQ; ++c) {
                        float tq = 0.f;
#pragma unroll
                        for (int e = 0; e < kKE; ++e) tq += part[((size_t)(e * kKQ + c) * kQB) * kH1 + i];
                        v = (c == 0) ? tq : v + tq;
                    }
                    g[(size_t)sl * kQB * kH1 + i] = v;
                }
            }
            if (sl + 1 < a.n_slots) router_cluster_sync();  // the leader has read them: the next slot may overwrite
            else if (crank != 0) return;                    // only the leader goes on
            __syncthreads();
        } else {
#pragma unroll
            for (int qb = 0; qb < kQB; ++qb) part[(kq * kQB + qb) * kH1 + t] = tot[qb];
            __syncthreads();
            for (int i = tid; i < kQB * kH1; i += kThreads) {
                float v = part[i];
#pragma unroll
                for (int c = 1; c < kKQ; ++c) v += part[c * kQB * kH1 + i];
                g[(size_t)sl * kQB * kH1 + i] = v;
            }
            __syncthreads();
        }
    }

    // ---- phase 2: (query, source) pairs, kPG at a time ----
    const int n_pairs = nqb * a.S;
    const int o2 = tid & (kH2 - 1);
    const int ks = tid >> 7;
    for (int p0 = 0; p0 < n_pairs; p0 += kPG) {
        const int np = min(kPG, n_pairs - p0);
        // (a) one warp per pair: h = relu(ln1(g + c1_s)), stored transposed for the fc2 threads
        if (w < kPG) {
            const int pi = p0 + w;
            float h[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) h[c] = 0.f;
            if (w < np) {
                const int qb = pi / a.S, s = pi - qb * a.S;
                const int sl = a.slot_of_source ? a.slot_of_source[s] : 0;
                const float* gp = g + ((size_t)sl * kQB + qb) * kH1;
                const float* cp = a.c1 + (size_t)s * kH1;
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int i = lane + 32 * c;
                    h[c] = gp[i] + cp[i];
                    sum += h[c];
                }
                const float mean1 = rr_warp_sum(sum) * (1.0f / kH1);
                float var = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float dlt = h[c] - mean1;
                    var = fmaf(dlt, dlt, var);
                }
                const float rstd1 = 1.0f / sqrtf(rr_warp_sum(var) * (1.0f / kH1) + kLnEps);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int i = lane + 32 * c;
                    h[c] = fmaxf((h[c] - mean1) * rstd1 * a.ln1w[i] + a.ln1b[i], 0.f);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) h1t[(lane + 32 * c) * kHStride + w] = h[c];
        }
        __syncthreads();
        // (b) fc2 partial sums: thread = (output o2, K-slice ks), all kPG pairs at once
        {
            float acc[kPG];
#pragma unroll
            for (int p = 0; p < kPG; ++p) acc[p] = 0.f;
            constexpr int kPer = kH1 / kKS;  // 32 inputs per slice
            const float* wr = a.w2t + (size_t)(ks * kPer) * kH2 + o2;
#pragma unroll
            for (int i0 = 0; i0 < kPer; i0 += kInFlight) {
                float wv[kInFlight];
#pragma unroll
                for (int u = 0; u < kInFlight; ++u) wv[u] = __ldg(wr + (size_t)(i0 + u) * kH2);
#pragma unroll
                for (int u = 0; u < kInFlight; ++u) {
                    const float4* hp = reinterpret_cast<const float4*>(h1t + (size_t)(ks * kPer + i0 + u) * kHStride);
#pragma unroll
                    for (int p4 = 0; p4 < kPG / 4; ++p4) {
                        const float4 hv = hp[p4];
                        acc[4 * p4 + 0] = fmaf(wv[u], hv.x, acc[4 * p4 + 0]);
                        acc[4 * p4 + 1] = fmaf(wv[u], hv.y, acc[4 * p4 + 1]);
                        acc[4 * p4 + 2] = fmaf(wv[u], hv.z, acc[4 * p4 + 2]);
                        acc[4 * p4 + 3] = fmaf(wv[u], hv.w, acc[4 * p4 + 3]);
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < kPG; ++p) part2[((size_t)ks * kPG + p) * kH2 + o2] = acc[p];
        }
        __syncthreads();
        // (c) one warp per pair: fold the K-slices, ln2, relu, fc3, sigmoid, threshold
        if (w < np) {
            const int pi = p0 + w;
            const int qb = pi / a.S, s = pi - qb * a.S;
            float o[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v = a.b2[lane + 32 * c];
#pragma unroll
                for (int k2 = 0; k2 < kKS; ++k2) v += part2[((size_t)k2 * kPG + w) * kH2 + lane + 32 * c];
                o[c] = v;
            }
            const float sum2 = o[0] + o[1] + o[2] + o[3];
            const float mean2 = rr_warp_sum(sum2) * (1.0f / kH2);
            float var2 = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float dlt = o[c] - mean2;
                var2 = fmaf(dlt, dlt, var2);
            }
            const float rstd2 = 1.0f / sqrtf(rr_warp_sum(var2) * (1.0f / kH2) + kLnEps);
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int i = lane + 32 * c;
                const float y = fmaxf((o[c] - mean2) * rstd2 * a.ln2w[i] + a.ln2b[i], 0.f);
                dot = fmaf(a.w3[i], y, dot);
            }
            const float logit = rr_warp_sum(dot) + a.b3[0];
            if (lane == 0) {
                const int64_t oidx = (b0 + qb) * a.S + s;
                if (a.logits) a.logits[oidx] = logit;
                if (a.mask) {
                    const float prob = 1.0f / (1.0f + expf(-logit));  // torch.sigmoid, router.py:276
                    a.mask[oidx] = prob > a.prob_thr ? 1 : 0;          // strict, router.py:278-280
                }
            }
        }
        __syncthreads();  // h1t / part2 are rewritten by the next round
    }
}

template <typename T>
int upload(T** dst, const T* src, size_t n) {
    RR_CUDA_CHECK(cudaMalloc((void**)dst, n * sizeof(T)));
    RR_CUDA_CHECK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return RR_OK;
}

size_t router_smem(int d_in, int n_slots, int cl) {
    const size_t n_x = ((size_t)kQB * d_in + 3) & ~(size_t)3;
    const size_t n_p1 = (cl > 1) ? (size_t)kKE * kKQ * kQB * kH1 : (size_t)kKQ * kQB * kH1;
    const size_t n_u = std::max(n_x + n_p1, (size_t)kKS * kPG * kH2);
    return (n_u + (size_t)n_slots * kQB * kH1 + (size_t)kH1 * kHStride) * sizeof(float);
}

template <int CL>
int router_launch_cl(const RouterArgs& a, size_t smem, cudaStream_t stream) {
    RR_CUDA_CHECK(cudaFuncSetAttribute(rr_router_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((a.B + kQB - 1) / kQB) * CL);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = rr_pdl_enabled() ? 2 : 1;
    RR_CUDA_CHECK(cudaLaunchKernelEx(&cfg, rr_router_kernel<CL>, a));
    return RR_OK;
}

int router_launch(rr_router* h, const RouterArgs& a, cudaStream_t stream) {
    // small batches: a cluster of kKE CTAs per kQB queries shares fc1 (when the leader's staging fits)
    static const bool no_cluster = getenv("RR_ROUTER_NO_CLUSTER") != nullptr;  // measurement knob
    const bool cluster = !no_cluster && a.B <= kClusterMaxBatch && router_smem(a.d_in, a.n_slots, kKE) <= (size_t)227 * 1024;
    const size_t smem = router_smem(a.d_in, a.n_slots, cluster ? kKE : 1);
    RR_REQUIRE(smem <= (size_t)227 * 1024, "router: %zu bytes of shared memory needed (d_in=%d, slots=%d)",
               smem, a.d_in, a.n_slots);
    int rc = cluster ? router_launch_cl<kKE>(a, smem, stream) : router_launch_cl<1>(a, smem, stream);
    if (rc != RR_OK) return rc;
    rr_count_launch(1);
    (void)h;
    return RR_OK;
}

RouterArgs base_args(const rr_router* h) {
    RouterArgs a = {};
    a.w1t = h->w1t;
    a.ln1w = h->ln1w;
    a.ln1b = h->ln1b;
    a.w2t = h->w2t;
    a.b2 = h->b2;
    a.ln2w = h->ln2w;
    a.ln2b = h->ln2b;
    a.w3 = h->w3;
    a.b3 = h->b3;
    return a;
}

}  // namespace

extern "C" int rr_router_create(int in_dim, int h1, int h2, const float* fc1_w, const float* fc1_b,
                                const float* ln1_w, const float* ln1_b, const float* fc2_w,
                                const float* fc2_b, const float* ln2_w, const float* ln2_b,
                                const float* fc3_w, const float* fc3_b, int device, rr_router_t** out) {
    RR_REQUIRE(out != nullptr, "rr_router_create: out is NULL");
    *out = nullptr;
    RR_REQUIRE(in_dim > 0, "rr_router_create: in_dim must be positive");
    if (h1 != kH1 || h2 != kH2) {
        rr_set_error("rr_router_create: hidden sizes %d/%d; CorpusRoutingNN is 256/128 (router.py:40,44)", h1, h2);
        return RR_ERR_UNSUPPORTED;
    }
    RR_REQUIRE(fc1_w && fc1_b && ln1_w && ln1_b && fc2_w && fc2_b && ln2_w && ln2_b && fc3_w && fc3_b,
               "rr_router_create: NULL weight pointer");
    int rc = rr_check_device_arch(device);
    if (rc != RR_OK) return rc;
    rr_device_guard g(device);
    rr_router* h = new (std::nothrow) rr_router();
    RR_REQUIRE(h != nullptr, "rr_router_create: out of host memory");
    h->device = device;
    h->in_dim = in_dim;
    h->h_w1.assign(fc1_w, fc1_w + (size_t)kH1 * in_dim);
    h->h_b1.assign(fc1_b, fc1_b + kH1);
    std::vector<float> w1t((size_t)in_dim * kH1), w2t((size_t)kH1 * kH2);
    for (int o = 0; o < kH1; ++o)
        for (int j = 0; j < in_dim; ++j) w1t[(size_t)j * kH1 + o] = fc1_w[(size_t)o * in_dim + j];
    for (int o = 0; o < kH2; ++o)
        for (int j = 0; j < kH1; ++j) w2t[(size_t)j * kH2 + o] = fc2_w[(size_t)o * kH1 + j];
    rc = upload(&h->w1t, w1t.data(), w1t.size());
    if (rc == RR_OK) rc = upload(&h->b1, fc1_b, kH1);
    if (rc == RR_OK) rc = upload(&h->ln1w, ln1_w, kH1);
    if (rc == RR_OK) rc = upload(&h->ln1b, ln1_b, kH1);
    if (rc == RR_OK) rc = upload(&h->w2t, w2t.data(), w2t.size());
    if (rc == RR_OK) rc = upload(&h->b2, fc2_b, kH2);
    if (rc == RR_OK) rc = upload(&h->ln2w, ln2_w, kH2);
    if (rc == RR_OK) rc = upload(&h->ln2b, ln2_b, kH2);
    if (rc == RR_OK) rc = upload(&h->w3, fc3_w, kH2);
    if (rc == RR_OK) rc = upload(&h->b3, fc3_b, 1);
    if (rc != RR_OK) {
        rr_router_destroy(h);
        return rc;
    }
    *out = h;
    return RR_OK;
}

extern "C" int rr_router_destroy(rr_router_t* h) {
    if (!h) return RR_OK;
    rr_device_guard g(h->device);
    cudaDeviceSynchronize();
    float* fs[] = {h->w1t, h->b1, h->ln1w, h->ln1b, h->w2t, h->b2, h->ln2w, h->ln2b, h->w3, h->b3, h->c1};
    for (float* p : fs) cudaFree(p);
    cudaFree(h->slot_of_source);
    cudaFree(h->sc_mean);
    cudaFree(h->sc_scale);
    delete h;
    return RR_OK;
}

extern "C" int rr_router_logits(rr_router_t* h, const float* x_dev, int64_t P, float* logits_dev,
                                void* stream) {
    RR_REQUIRE(h != nullptr, "rr_router_logits: NULL handle");
    RR_REQUIRE(P >= 0, "rr_router_logits: negative P");
    if (P == 0) return RR_OK;
    RR_REQUIRE(x_dev && logits_dev, "rr_router_logits: NULL buffer");
    rr_device_guard g(h->device);
    RouterArgs a = base_args(h);
    a.c1 = h->b1;
    a.x = x_dev;
    a.B = P;
    a.S = 1;
    a.n_slots = 1;
    a.d_in = h->in_dim;
    a.logits = logits_dev;
    return router_launch(h, a, (cudaStream_t)stream);
}

extern "C" int rr_router_set_tables(rr_router_t* h, int n_sources, int d_max, int n_slots,
                                    const int32_t* slot_of_source, const float* centroids,
                                    const int32_t* onehot_id, int n_onehot, const double* scaler_mean,
                                    const double* scaler_scale) {
    RR_REQUIRE(h != nullptr, "rr_router_set_tables: NULL handle");
    RR_REQUIRE(n_sources > 0 && d_max > 0 && n_slots > 0 && n_onehot >= 0, "rr_router_set_tables: bad sizes");
    RR_REQUIRE(2 * d_max + n_onehot == h->in_dim,
               "rr_router_set_tables: 2*d_max + n_onehot = %d but the network's in_dim is %d",
               2 * d_max + n_onehot, h->in_dim);
    RR_REQUIRE(slot_of_source && centroids && (n_onehot == 0 || onehot_id), "rr_router_set_tables: NULL table");
    RR_REQUIRE((scaler_mean == nullptr) == (scaler_scale == nullptr), "rr_router_set_tables: scaler needs mean and scale");
    for (int s = 0; s < n_sources; ++s) {
        RR_REQUIRE(slot_of_source[s] >= 0 && slot_of_source[s] < n_slots, "rr_router_set_tables: slot out of range");
        RR_REQUIRE(n_onehot == 0 || (onehot_id[s] >= 0 && onehot_id[s] < n_onehot), "rr_router_set_tables: one-hot id out of range");
    }
    rr_device_guard g(h->device);
    RR_CUDA_CHECK(cudaDeviceSynchronize());
    const int in_dim = h->in_dim;
    // c1_s = b1 + W1[:, d_max:] . fp32(scaled [centroid_s | onehot_s])   (float64 accumulate)
    std::vector<float> c1((size_t)n_sources * kH1);
    std::vector<float> xc(in_dim - d_max);
    for (int s = 0; s < n_sources; ++s) {
        for (int j = 0; j < in_dim - d_max; ++j) {
            double f = (j < d_max) ? (double)centroids[(size_t)s * d_max + j]
                                   : ((j - d_max) == onehot_id[s] ? 1.0 : 0.0);
            if (scaler_mean) f = (f - scaler_mean[d_max + j]) / scaler_scale[d_max + j];
            xc[j] = (float)f;  // the reference casts the scaled row to fp32 (router.py:271)
        }
        for (int o = 0; o < kH1; ++o) {
            double acc = (double)h->h_b1[o];
            const float* wr = h->h_w1.data() + (size_t)o * in_dim + d_max;
            for (int j = 0; j < in_dim - d_max; ++j) acc += (double)wr[j] * (double)xc[j];
            c1[(size_t)s * kH1 + o] = (float)acc;
        }
    }
    cudaFree(h->c1);
    cudaFree(h->slot_of_source);
    cudaFree(h->sc_mean);
    cudaFree(h->sc_scale);
    h->c1 = nullptr;
    h->slot_of_source = nullptr;
    h->sc_mean = h->sc_scale = nullptr;
    int rc = upload(&h->c1, c1.data(), c1.size());
    if (rc == RR_OK) rc = upload(&h->slot_of_source, slot_of_source, (size_t)n_sources);
    if (rc == RR_OK && scaler_mean) rc = upload(&h->sc_mean, scaler_mean, (size_t)d_max);
    if (rc == RR_OK && scaler_scale) rc = upload(&h->sc_scale, scaler_scale, (size_t)d_max);
    if (rc != RR_OK) return rc;
    h->n_sources = n_sources;
    h->d_max = d_max;
    h->n_slots = n_slots;
    return RR_OK;
}

extern "C" int rr_router_select(rr_router_t* h, const float* q_dev, int64_t B, float prob_threshold,
                                uint8_t* mask_dev, float* logits_dev, void* stream) {
    RR_REQUIRE(h != nullptr, "rr_router_select: NULL handle");
    RR_REQUIRE(h->n_sources > 0, "rr_router_select: call rr_router_set_tables first");
    RR_REQUIRE(B >= 0, "rr_router_select: negative B");
    if (B == 0) return RR_OK;
    RR_REQUIRE(q_dev && (mask_dev || logits_dev), "rr_router_select: NULL buffer");
    rr_device_guard g(h->device);
    RouterArgs a = base_args(h);
    a.c1 = h->c1;
    a.slot_of_source = h->slot_of_source;
    a.sc_mean = h->sc_mean;
    a.sc_scale = h->sc_scale;
    a.x = q_dev;
    a.B = B;
    a.S = h->n_sources;
    a.n_slots = h->n_slots;
    a.d_in = h->d_max;
    a.prob_thr = prob_threshold;
    a.mask = mask_dev;
    a.logits = logits_dev;
    return router_launch(h, a, (cudaStream_t)stream);
}
