/* Plain-C caller of the C ABI (include/ragroute_b200.h): what a cgo / JNI / ctypes binding sees.
 * No CUDA headers, no torch types -- only pointers and sizes.
 *
 *   abi_check            load-time checks only (runs on a box without a GPU)
 *   abi_check --compute  also build a tiny hand-computable index and search it (needs a B200)
 *
 * The corpus is the known-answer case of tests/test_oracle_flat_ip.py: row i of 64 x 8 is
 * (i, 1, 0, 0, 0, 0, 0, 0); with q = (1, 0, ...) the scores are 0..63, so the top-k ids are
 * 63, 62, ...; rows 10 and 11 are then made equal to row 63 to exercise the tie rule (lower id
 * first, as faiss' heap keeps the earlier row on ties). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ragroute_b200.h"

#define CHECK(cond, ...)                                  \
    do {                                                  \
        if (!(cond)) {                                    \
            fprintf(stderr, "abi_check FAILED: " __VA_ARGS__); \
            fprintf(stderr, " (last error: %s)\n", rr_last_error()); \
            return 1;                                     \
        }                                                 \
    } while (0)

int main(int argc, char** argv) {
    const int compute = argc > 1 && !strcmp(argv[1], "--compute");
    CHECK(rr_abi_version() == RR_ABI_VERSION, "ABI version %d != header %d", rr_abi_version(), RR_ABI_VERSION);
    CHECK(rr_last_error() != NULL, "rr_last_error returned NULL");
    rr_index_t* h = NULL;
    CHECK(rr_index_create(0, 0, 0, &h) == RR_ERR_INVALID && h == NULL, "d = 0 must be rejected");
    CHECK(strlen(rr_last_error()) > 0, "no message for a rejected call");
    const int n_dev = rr_device_count();
    if (!compute) {
        if (n_dev <= 0) {
            /* no CPU fallback: creating an index must fail loudly, not silently succeed */
            const int rc = rr_index_create(8, 0, 0, &h);
            CHECK(rc == RR_ERR_CUDA || rc == RR_ERR_ARCH, "index creation without a device returned %d", rc);
        }
        printf("abi_check ok (load-time, devices: %d)\n", n_dev);
        return 0;
    }
    CHECK(n_dev > 0, "--compute needs a CUDA device");
    enum { N = 64, D = 8, K = 5, NQ = 2 };
    float x[N * D];
    memset(x, 0, sizeof(x));
    for (int i = 0; i < N; ++i) {
        x[i * D] = (float)i;
        x[i * D + 1] = 1.0f;
    }
    x[10 * D] = 63.0f; /* ties with row 63 */
    x[11 * D] = 63.0f;
    CHECK(rr_index_create(D, N, 0, &h) == RR_OK, "rr_index_create");
    CHECK(rr_index_add_f32(h, x, N, 0, NULL) == RR_OK, "rr_index_add_f32");
    CHECK(rr_index_ntotal(h) == N && rr_index_dim(h) == D, "ntotal / dim");
    const float q[NQ * D] = {1, 0, 0, 0, 0, 0, 0, 0, /* scores = first column */
                             -1, 100, 0, 0, 0, 0, 0, 0 /* scores = 100 - first column */};
    float Dv[NQ * K];
    int64_t Iv[NQ * K];
    for (int mode = 0; mode < 2; ++mode) {
        CHECK(rr_index_set_param(h, "mode", (double)mode) == RR_OK, "set mode %d", mode);
        CHECK(rr_index_search_host(h, q, NQ, K, Dv, Iv) == RR_OK, "rr_index_search_host (mode %d)", mode);
        const int64_t want0[K] = {10, 11, 63, 62, 61};
        const float score0[K] = {63, 63, 63, 62, 61};
        const int64_t want1[K] = {0, 1, 2, 3, 4};
        for (int j = 0; j < K; ++j) {
            CHECK(Iv[j] == want0[j] && Dv[j] == score0[j], "mode %d query 0 slot %d: id %lld score %g", mode, j,
                  (long long)Iv[j], Dv[j]);
            CHECK(Iv[K + j] == want1[j] && Dv[K + j] == 100.0f - (float)j, "mode %d query 1 slot %d: id %lld score %g",
                  mode, j, (long long)Iv[K + j], Dv[K + j]);
        }
    }
    CHECK(rr_index_search_host(h, q, NQ, RR_MAX_K + 1, Dv, Iv) == RR_ERR_UNSUPPORTED, "k > RR_MAX_K must be refused");
    rr_index_stats st;
    CHECK(rr_index_get_stats(h, &st, 0) == RR_OK && st.queries == 2 * NQ, "stats: %lld queries", (long long)st.queries);
    CHECK(rr_index_destroy(h) == RR_OK, "rr_index_destroy");
    printf("abi_check ok (compute, %lld kernel launches)\n", (long long)rr_kernel_launches());
    return 0;
}
