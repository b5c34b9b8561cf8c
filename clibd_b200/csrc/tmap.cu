#include "tmap.h"

#include <mutex>
#include <string>

#include "common.cuh"

namespace clibd {

namespace {
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;
std::once_flag g_once;

void resolve() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = reinterpret_cast<EncodeFn>(fn);
}
}  // namespace

int make_tmap_2d_16bit(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch_elems,
                       uint32_t box_cols, uint32_t box_rows, int fmt_bf16) {
    std::call_once(g_once, resolve);
    if (g_encode == nullptr) {
        set_error("cuTensorMapEncodeTiled is not available from this driver");
        return 2;
    }
    CLIBD_REQUIRE(box_cols * 2 == 128, "TMA box inner extent must be 128 bytes for SWIZZLE_128B");
    CLIBD_REQUIRE(box_rows >= 1 && box_rows <= 256, "TMA box rows must be in [1,256]");
    CLIBD_REQUIRE((pitch_elems * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
    CLIBD_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstride[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estride[2] = {1, 1};
    CUresult r = g_encode(out, fmt_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                          const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
        return 2;
    }
    return 0;
}

}  // namespace clibd
