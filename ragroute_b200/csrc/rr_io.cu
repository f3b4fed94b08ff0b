// S5: faiss.read_index / faiss.write_index for flat inner-product indexes
// (/root/reference/ragroute/data_source.py:71 `index = faiss.read_index(self.index_path)`; the files are the
// MedRAG `faiss.index`, FeB4RAG `<name>_<model>.faiss` and Wikipedia `faiss_index_<c>_normalized.index` of
// data_source.py:47,52,56).
//
// File layout (upstream faiss 1.7.4, faiss/impl/index_write.cpp write_index_header + IndexFlat, from
// knowledge -- no FAISS-written file exists in this container to verify against):
//   fourcc "IxFI" (IndexFlatIP) | "IxFl" (IndexFlat with a generic metric) | "IxF2" (IndexFlatL2: rejected)
//   d int32 | ntotal int64 | dummy int64 | dummy int64 | is_trained uint8 | metric_type int32
//   [metric_arg float32 when metric_type > 1]
//   count uint64 (= ntotal * d; 1.7.3+ writes the byte vector `codes` as size / 4, older versions the float
//   vector `xb` -- the same bytes) | ntotal * d little-endian fp32, row-major
//
// The rows never have to fit in host memory: they stream disk -> pinned host buffer -> device staging ->
// split kernel with two buffers of each kind, so the read of chunk i+1 runs while chunk i is copied and
// converted (PubMed's faiss.index is 73 GB).
#include <errno.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "rr_index_internal.cuh"

namespace {

constexpr size_t kChunkBytes = 64u << 20;

struct FlatHeader {
    char fourcc[4];
    int32_t d;
    int64_t ntotal;
    uint8_t is_trained;
    int32_t metric;
    uint64_t count;
};

bool read_exact(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

int read_header(FILE* f, const char* path, FlatHeader* h) {
    int64_t dummy[2];
    float metric_arg;
    bool ok = read_exact(f, h->fourcc, 4) && read_exact(f, &h->d, 4) && read_exact(f, &h->ntotal, 8) &&
              read_exact(f, dummy, 16) && read_exact(f, &h->is_trained, 1) && read_exact(f, &h->metric, 4);
    if (ok && h->metric > 1) ok = read_exact(f, &metric_arg, 4);
    if (ok) ok = read_exact(f, &h->count, 8);
    RR_REQUIRE(ok, "rr_index_read: %s: truncated header", path);
    const bool flat = !memcmp(h->fourcc, "IxFI", 4) || !memcmp(h->fourcc, "IxFl", 4);
    RR_REQUIRE(flat && h->metric == 0,
               "rr_index_read: %s: only flat inner-product indexes are supported (fourcc '%.4s', metric %d); "
               "data_source.py uses IndexFlatIP",
               path, h->fourcc, (int)h->metric);
    RR_REQUIRE(h->d > 0 && h->ntotal >= 0 && h->count == (uint64_t)h->ntotal * (uint64_t)h->d,
               "rr_index_read: %s: corrupt flat index (%llu floats for %lld x %d)", path,
               (unsigned long long)h->count, (long long)h->ntotal, (int)h->d);
    return RR_OK;
}

struct Buffers {
    void* pin[2] = {nullptr, nullptr};
    void* dev[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    cudaStream_t stream = nullptr;
    ~Buffers() {
        if (stream) cudaStreamSynchronize(stream);
        for (int i = 0; i < 2; ++i) {
            if (pin[i]) cudaFreeHost(pin[i]);
            if (dev[i]) cudaFree(dev[i]);
            if (ev[i]) cudaEventDestroy(ev[i]);
        }
        if (stream) cudaStreamDestroy(stream);
    }
    int init(size_t bytes) {
        RR_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            RR_CUDA_CHECK(cudaHostAlloc(&pin[i], bytes, cudaHostAllocDefault));
            RR_CUDA_CHECK(cudaMalloc(&dev[i], bytes));
            RR_CUDA_CHECK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
        }
        return RR_OK;
    }
};

}  // namespace

extern "C" int rr_index_read(const char* path, int device, rr_index_t** out) {
    RR_REQUIRE(path != nullptr && out != nullptr, "rr_index_read: NULL argument");
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    RR_REQUIRE(f != nullptr, "rr_index_read: cannot open %s: %s", path, strerror(errno));
    FlatHeader hd;
    int rc = read_header(f, path, &hd);
    if (rc != RR_OK) {
        fclose(f);
        return rc;
    }
    rr_index_t* h = nullptr;
    rc = rr_index_create(hd.d, hd.ntotal, device, &h);
    if (rc != RR_OK) {
        fclose(f);
        return rc;
    }
    rr_device_guard g(device);
    const size_t row_bytes = (size_t)hd.d * sizeof(float);
    const int64_t chunk_rows = std::max<int64_t>(1, std::min<int64_t>((int64_t)(kChunkBytes / row_bytes), std::max<int64_t>(hd.ntotal, 1)));
    {
        Buffers b;
        rc = b.init((size_t)chunk_rows * row_bytes);
        int i = 0;
        for (int64_t r0 = 0; rc == RR_OK && r0 < hd.ntotal; r0 += chunk_rows, ++i) {
            const int64_t nr = std::min(chunk_rows, hd.ntotal - r0);
            const int s = i & 1;
            // the copy that last used this pinned buffer (two chunks ago) must have left it
            if (i >= 2 && cudaEventSynchronize(b.ev[s]) != cudaSuccess) {
                rr_set_error("rr_index_read: %s: %s", path, cudaGetErrorString(cudaGetLastError()));
                rc = RR_ERR_CUDA;
                break;
            }
            if (!read_exact(f, b.pin[s], (size_t)nr * row_bytes)) {  // the previous chunk is on the GPU meanwhile
                rr_set_error("rr_index_read: %s: truncated at row %lld of %lld", path, (long long)r0, (long long)hd.ntotal);
                rc = RR_ERR_INVALID;
                break;
            }
            cudaError_t e = cudaMemcpyAsync(b.dev[s], b.pin[s], (size_t)nr * row_bytes, cudaMemcpyHostToDevice, b.stream);
            if (e == cudaSuccess) e = cudaEventRecord(b.ev[s], b.stream);
            if (e != cudaSuccess) {
                rr_set_error("rr_index_read: %s: %s", path, cudaGetErrorString(e));
                rc = RR_ERR_CUDA;
                break;
            }
            rc = rr_index_add_f32(h, (const float*)b.dev[s], nr, 1, b.stream);  // asynchronous split into the planes
        }
        if (rc == RR_OK && cudaStreamSynchronize(b.stream) != cudaSuccess) {
            rr_set_error("rr_index_read: %s: %s", path, cudaGetErrorString(cudaGetLastError()));
            rc = RR_ERR_CUDA;
        }
    }
    fclose(f);
    if (rc != RR_OK) {
        rr_index_destroy(h);
        return rc;
    }
    h->add_event_pending = false;  // synchronised above
    *out = h;
    return RR_OK;
}

extern "C" int rr_index_write(rr_index_t* h, const char* path) {
    RR_REQUIRE(h != nullptr && path != nullptr, "rr_index_write: NULL argument");
    rr_device_guard g(h->device);
    RR_CUDA_CHECK(cudaDeviceSynchronize());  // rows added asynchronously on any stream are in place
    FILE* f = fopen(path, "wb");
    RR_REQUIRE(f != nullptr, "rr_index_write: cannot open %s: %s", path, strerror(errno));
    const int32_t d = h->d, metric = 0;
    const int64_t ntotal = h->n, dummy = 1 << 20;
    const uint8_t trained = 1;
    const uint64_t count = (uint64_t)ntotal * (uint64_t)d;
    bool ok = fwrite("IxFI", 1, 4, f) == 4 && fwrite(&d, 4, 1, f) == 1 && fwrite(&ntotal, 8, 1, f) == 1 &&
              fwrite(&dummy, 8, 1, f) == 1 && fwrite(&dummy, 8, 1, f) == 1 && fwrite(&trained, 1, 1, f) == 1 &&
              fwrite(&metric, 4, 1, f) == 1 && fwrite(&count, 8, 1, f) == 1;
    int rc = RR_OK;
    const size_t row_bytes = (size_t)d * sizeof(float);
    const int64_t chunk_rows = std::max<int64_t>(1, std::min<int64_t>((int64_t)(kChunkBytes / row_bytes), std::max<int64_t>(ntotal, 1)));
    if (ok && ntotal > 0) {
        Buffers b;
        rc = b.init((size_t)chunk_rows * row_bytes);
        // chunk i is rebuilt and copied out while chunk i-1 is written to the file
        const int64_t n_chunks = (ntotal + chunk_rows - 1) / chunk_rows;
        for (int64_t i = 0; rc == RR_OK && ok && i <= n_chunks; ++i) {
            const int s = (int)(i & 1);
            if (i < n_chunks) {
                const int64_t r0 = i * chunk_rows, nr = std::min(chunk_rows, ntotal - r0);
                rc = rr_index_reconstruct(h, r0, nr, (float*)b.dev[s], b.stream);
                cudaError_t e = cudaSuccess;
                if (rc == RR_OK) e = cudaMemcpyAsync(b.pin[s], b.dev[s], (size_t)nr * row_bytes, cudaMemcpyDeviceToHost, b.stream);
                if (rc == RR_OK && e == cudaSuccess) e = cudaEventRecord(b.ev[s], b.stream);
                if (rc == RR_OK && e != cudaSuccess) {
                    rr_set_error("rr_index_write: %s", cudaGetErrorString(e));
                    rc = RR_ERR_CUDA;
                }
            }
            if (rc == RR_OK && i >= 1) {
                const int64_t p0 = (i - 1) * chunk_rows, pr = std::min(chunk_rows, ntotal - p0);
                if (cudaEventSynchronize(b.ev[s ^ 1]) != cudaSuccess) {
                    rr_set_error("rr_index_write: %s", cudaGetErrorString(cudaGetLastError()));
                    rc = RR_ERR_CUDA;
                } else {
                    ok = fwrite(b.pin[s ^ 1], row_bytes, (size_t)pr, f) == (size_t)pr;
                }
            }
        }
    }
    if (fclose(f) != 0) ok = false;
    if (rc != RR_OK) return rc;
    RR_REQUIRE(ok, "rr_index_write: %s: write failed: %s", path, strerror(errno));
    return RR_OK;
}
