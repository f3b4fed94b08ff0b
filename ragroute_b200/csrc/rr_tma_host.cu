// Host side of TMA: cuTensorMapEncodeTiled through the runtime's driver entry point
// (no -lcuda link dependency).
#include "rr_index_internal.cuh"

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

int rr_make_tmap_bf16_rows(CUtensorMap* out, const void* ptr, int64_t rows, int d_pad, int box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) {
        rr_set_error("cuTensorMapEncodeTiled is not available from the driver");
        return RR_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)d_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)RR_TC_KCHUNK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        rr_set_error("cuTensorMapEncodeTiled failed (CUresult %d) rows=%lld d_pad=%d box_rows=%d", (int)r,
                     (long long)rows, d_pad, box_rows);
        return RR_ERR_CUDA;
    }
    return RR_OK;
}

