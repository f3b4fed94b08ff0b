/*
 * oracle/flat_ip_ref.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the arithmetic behind the reference's per-source search
 *     res_ = index.search(query_embed, k=k)
 * (/root/reference/ragroute/data_source.py:158, :186, :203), i.e. FAISS
 * IndexFlatIP::search.  The arithmetic lives in the third-party dependency
 * faiss-cpu==1.7.4 (/root/reference/environment.yml:55), whose sources are NOT
 * vendored under /root/reference and which is not installable here (no network,
 * no wheel in /opt/wheelhouse).  This file therefore restates FAISS' published
 * algorithm for nq < 20 (faiss/utils/distances.cpp, exhaustive_inner_product_seq
 * + HeapResultHandler<CMin<float,int64>>; faiss/utils/distances_simd.cpp,
 * AVX2 fvec_inner_product; faiss/utils/Heap.h, minheap_replace_top /
 * heap_reorder):
 *
 *   - one sequential pass per query over the N rows (OpenMP over queries only,
 *     num_threads = min(nq, max_threads));
 *   - the dot product keeps 8 fp32 lane sums (separate multiply and add, no
 *     FMA), folds 8->4 lanes, then two horizontal adds;
 *   - a size-k min-heap of (score, id), initialised to (-FLT_MAX, -1), replaces
 *     its top only on strict "score > heap_top", so at an entry tie the earlier
 *     row stays;
 *   - the heap is finally reordered to descending score; unused slots stay
 *     (-FLT_MAX, -1) at the tail.
 *
 * PARITY UNPINNED: the reference holds no golden vector, known-answer test or
 * fixture for this path and FAISS itself cannot be run here, so this oracle is
 * pinned only by its own hand-computable known-answer tests
 * (tests/test_oracle_flat_ip.py) and by a float64 exhaustive truth pass.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * legs may load this library.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* faiss/utils/distances_simd.cpp (AVX2 build): fvec_inner_product. */
static float ref_fvec_inner_product(const float* x, const float* y, size_t d) {
    float lane[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    while (d >= 8) {
        for (int l = 0; l < 8; ++l) {
            float p = x[l] * y[l]; /* _mm256_mul_ps */
            lane[l] = lane[l] + p; /* _mm256_add_ps */
        }
        x += 8;
        y += 8;
        d -= 8;
    }
    /* msum2 = hi128 + lo128 */
    float m[4];
    for (int l = 0; l < 4; ++l) m[l] = lane[4 + l] + lane[l];
    if (d >= 4) {
        for (int l = 0; l < 4; ++l) {
            float p = x[l] * y[l];
            m[l] = m[l] + p;
        }
        x += 4;
        y += 4;
        d -= 4;
    }
    if (d > 0) { /* masked_read: remaining lanes are zero */
        for (size_t l = 0; l < d; ++l) {
            float p = x[l] * y[l];
            m[l] = m[l] + p;
        }
    }
    /* two _mm_hadd_ps */
    float h0 = m[0] + m[1];
    float h1 = m[2] + m[3];
    return h0 + h1;
}

/* faiss/utils/Heap.h: heap_replace_top<CMin<float,int64_t>> (1-based sift-down). */
static void ref_minheap_replace_top(size_t k, float* bh_val, int64_t* bh_ids,
                                    float val, int64_t id) {
    bh_val--; /* use 1-based indexing */
    bh_ids--;
    size_t i = 1, i1, i2;
    while (1) {
        i1 = i << 1;
        i2 = i1 + 1;
        if (i1 > k) break;
        if (i2 == k + 1 || bh_val[i1] < bh_val[i2]) { /* CMin::cmp(a,b) = a < b */
            if (val < bh_val[i1]) break;
            bh_val[i] = bh_val[i1];
            bh_ids[i] = bh_ids[i1];
            i = i1;
        } else {
            if (val < bh_val[i2]) break;
            bh_val[i] = bh_val[i2];
            bh_ids[i] = bh_ids[i2];
            i = i2;
        }
    }
    bh_val[i] = val;
    bh_ids[i] = id;
}

/* faiss/utils/Heap.h: heap_pop (remove the top = current minimum). */
static void ref_minheap_pop(size_t k, float* bh_val, int64_t* bh_ids) {
    bh_val--;
    bh_ids--;
    float val = bh_val[k];
    size_t i = 1, i1, i2;
    while (1) {
        i1 = i << 1;
        i2 = i1 + 1;
        if (i1 > k) break;
        if (i2 == k + 1 || bh_val[i1] < bh_val[i2]) {
            if (val < bh_val[i1]) break;
            bh_val[i] = bh_val[i1];
            bh_ids[i] = bh_ids[i1];
            i = i1;
        } else {
            if (val < bh_val[i2]) break;
            bh_val[i] = bh_val[i2];
            bh_ids[i] = bh_ids[i2];
            i = i2;
        }
    }
    bh_val[i] = bh_val[k];
    bh_ids[i] = bh_ids[k];
}

/* faiss/utils/Heap.h: heap_reorder -- pops into the tail, compacts, pads with
 * (neutral, -1).  For a min-heap the result is descending by score. */
static void ref_minheap_reorder(size_t k, float* bh_val, int64_t* bh_ids) {
    size_t i, ii;
    for (i = 0, ii = 0; i < k; i++) {
        float val = bh_val[0];
        int64_t id = bh_ids[0];
        ref_minheap_pop(k - i, bh_val, bh_ids);
        bh_val[k - ii - 1] = val;
        bh_ids[k - ii - 1] = id;
        if (id != -1) ii++;
    }
    memmove(bh_val, bh_val + k - ii, ii * sizeof(*bh_val));
    memmove(bh_ids, bh_ids + k - ii, ii * sizeof(*bh_ids));
    for (; ii < k; ii++) {
        bh_val[ii] = -FLT_MAX;
        bh_ids[ii] = -1;
    }
}

/*
 * IndexFlatIP::search for the nq < 20 branch (the only one the reference ever
 * reaches: data_source.py:114 reshapes every query to [1, d]).
 *   x  [nq, d] fp32 row-major queries
 *   xb [nb, d] fp32 row-major corpus
 *   D  [nq, k] fp32 out, descending; I [nq, k] int64 out
 * n_threads <= 0 means "all" (FAISS: omp parallel over queries).
 */
int ref_flat_ip_search_seq(const float* x, const float* xb, int64_t nq, int64_t nb,
                           int64_t d, int64_t k, float* D, int64_t* I, int n_threads) {
    if (k <= 0 || d <= 0 || nq < 0 || nb < 0) return -1;
#ifdef _OPENMP
    int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
    if (nt > nq) nt = (int)(nq > 0 ? nq : 1);
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (int64_t i = 0; i < nq; ++i) {
        const float* xi = x + i * d;
        float* simi = D + i * k;
        int64_t* idxi = I + i * k;
        for (int64_t j = 0; j < k; ++j) { /* heap_heapify: neutral fill */
            simi[j] = -FLT_MAX;
            idxi[j] = -1;
        }
        for (int64_t j = 0; j < nb; ++j) {
            float ip = ref_fvec_inner_product(xi, xb + j * d, (size_t)d);
            if (simi[0] < ip) { /* CMin::cmp(threshold, dis): strict */
                ref_minheap_replace_top((size_t)k, simi, idxi, ip, j);
            }
        }
        ref_minheap_reorder((size_t)k, simi, idxi);
    }
    return 0;
}

/*
 * The nq >= 20 branch (faiss/utils/distances.cpp, exhaustive_inner_product_blas): FAISS multiplies
 * query blocks of 4096 by corpus blocks of 1024 with sgemm_ and hands every block of scores to
 * HeapResultHandler::add_results(j0, j1, ip_block): "#pragma omp parallel for" over the queries of
 * the block, one pass over the row of scores with the same strict "score > heap_top" rule as above.
 * The sgemm itself is the caller's (oracle/flat_ip.py: numpy / OpenBLAS); these two functions are
 * the heap side.  D / I hold the running heaps between blocks (heapified to (-FLT_MAX, -1) by
 * ref_heap_begin, reordered to descending score by ref_heap_end).
 *   ip [nq, ld] fp32 scores of rows j0 .. j0 + nb_blk of the corpus.
 */
int ref_heap_begin(int64_t nq, int64_t k, float* D, int64_t* I) {
    if (k <= 0 || nq < 0) return -1;
    for (int64_t i = 0; i < nq * k; ++i) {
        D[i] = -FLT_MAX;
        I[i] = -1;
    }
    return 0;
}

int ref_heap_add_results(const float* ip, int64_t nq, int64_t nb_blk, int64_t ld, int64_t j0, int64_t k,
                         float* D, int64_t* I, int n_threads) {
    if (k <= 0 || nq < 0 || nb_blk < 0 || ld < nb_blk) return -1;
#ifdef _OPENMP
    int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
#pragma omp parallel for num_threads(nt) schedule(static)
#endif
    for (int64_t i = 0; i < nq; ++i) {
        float* simi = D + i * k;
        int64_t* idxi = I + i * k;
        const float* row = ip + i * ld;
        float thresh = simi[0];
        for (int64_t j = 0; j < nb_blk; ++j) {
            const float dis = row[j];
            if (thresh < dis) { /* CMin::cmp(threshold, dis): strict */
                ref_minheap_replace_top((size_t)k, simi, idxi, dis, j0 + j);
                thresh = simi[0];
            }
        }
    }
    return 0;
}

int ref_heap_end(int64_t nq, int64_t k, float* D, int64_t* I) {
    if (k <= 0 || nq < 0) return -1;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t i = 0; i < nq; ++i) ref_minheap_reorder((size_t)k, D + i * k, I + i * k);
    return 0;
}

/*
 * Row-parallel variant used ONLY as the multi-core CPU baseline in bench.py:
 * the same dot product and the same heap rule, but the N rows of one query are
 * split over threads (FAISS would not do this for nq < 20) and the per-thread
 * heaps are merged.  Reported as kind "port".
 */
int ref_flat_ip_search_rowpar(const float* x, const float* xb, int64_t nq, int64_t nb,
                              int64_t d, int64_t k, float* D, int64_t* I, int n_threads) {
    if (k <= 0 || d <= 0 || nq < 0 || nb < 0) return -1;
    int nt = 1;
#ifdef _OPENMP
    nt = n_threads > 0 ? n_threads : omp_get_max_threads();
#endif
    float* hv = (float*)malloc(sizeof(float) * (size_t)nt * (size_t)k);
    int64_t* hi = (int64_t*)malloc(sizeof(int64_t) * (size_t)nt * (size_t)k);
    if (!hv || !hi) { free(hv); free(hi); return -2; }
    for (int64_t i = 0; i < nq; ++i) {
        const float* xi = x + i * d;
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
        {
            int t = 0, T = 1;
#ifdef _OPENMP
            t = omp_get_thread_num();
            T = omp_get_num_threads();
#endif
            float* simi = hv + (size_t)t * k;
            int64_t* idxi = hi + (size_t)t * k;
            for (int64_t j = 0; j < k; ++j) { simi[j] = -FLT_MAX; idxi[j] = -1; }
            int64_t lo = nb * t / T, up = nb * (t + 1) / T;
            for (int64_t j = lo; j < up; ++j) {
                float ip = ref_fvec_inner_product(xi, xb + j * d, (size_t)d);
                if (simi[0] < ip) ref_minheap_replace_top((size_t)k, simi, idxi, ip, j);
            }
            ref_minheap_reorder((size_t)k, simi, idxi);
#ifdef _OPENMP
#pragma omp barrier
#pragma omp single
#endif
            {
                /* merge the T sorted lists, in thread (= row) order, same heap rule */
                float* so = D + i * k;
                int64_t* io = I + i * k;
                for (int64_t j = 0; j < k; ++j) { so[j] = -FLT_MAX; io[j] = -1; }
                for (int tt = 0; tt < T; ++tt)
                    for (int64_t j = 0; j < k; ++j) {
                        float v = hv[(size_t)tt * k + j];
                        int64_t id = hi[(size_t)tt * k + j];
                        if (id >= 0 && so[0] < v) ref_minheap_replace_top((size_t)k, so, io, v, id);
                    }
                ref_minheap_reorder((size_t)k, so, io);
            }
        }
    }
    free(hv);
    free(hi);
    return 0;
}

/* float64 exhaustive scores for one query (truth pass used to classify ties). */
int ref_flat_ip_scores_f64(const float* x, const float* xb, int64_t nb, int64_t d, double* out) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int64_t j = 0; j < nb; ++j) {
        const float* y = xb + j * d;
        double s = 0.0;
        for (int64_t t = 0; t < d; ++t) s += (double)x[t] * (double)y[t];
        out[j] = s;
    }
    return 0;
}

/* faiss/utils/distances.cpp: fvec_renorm_L2 (faiss.normalize_L2,
 * /root/reference/ragroute/data_source.py:199): x /= sqrt(fvec_norm_L2sqr(x))
 * when the norm is positive.  The squared norm uses the same lane pattern as the
 * inner product in the AVX2 build. */
int ref_normalize_L2(float* x, int64_t n, int64_t d) {
    for (int64_t i = 0; i < n; ++i) {
        float* xi = x + i * d;
        float nr = ref_fvec_inner_product(xi, xi, (size_t)d);
        if (nr > 0) {
            const float inv_nr = 1.0f / sqrtf(nr);
            for (int64_t j = 0; j < d; ++j) xi[j] *= inv_nr;
        }
    }
    return 0;
}
