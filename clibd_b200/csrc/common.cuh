// Shared host/device helpers for the clibd_b200 CUDA library.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdint>
#include <cstdio>
#include <string>

namespace clibd {

// dtype codes of the C ABI (include/clibd_b200.h)
enum : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };
// compute paths
enum : int { PATH_SIMT_F32 = 0, PATH_TC_BF16 = 1, PATH_TC_F16 = 2 };

void set_error(const std::string& msg);
// telemetry (loss_api.cu): number of kernels this library launched, and optional CUDA-event timing
// of the tensor-core kernels on their launching stream (slots below)
void count_launch();
enum : int { PROF_LOSS_FWD_TC = 0, PROF_LOSS_BWD_TC = 1, PROF_KNN_SCREEN_TC = 2, PROF_KNN_RERANK = 3, PROF_LOSS_GRAD_GEMM = 4, PROF_SLOTS = 8 };
struct ProfScope {
    ProfScope(int slot, cudaStream_t s);
    ~ProfScope();
    int slot_;
    cudaStream_t s_;
    cudaEvent_t e0_ = nullptr, e1_ = nullptr;
};

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute: set it once per
// (kernel, device) under a mutex, so that a process that drives several GPUs, or autograd's worker thread racing
// the main thread on first use, never launches with the 48 KB default.  Returns a cudaError_t.
cudaError_t ensure_dynamic_smem(const void* kernel, int bytes);

// NVTX range around an entry point (header-only NVTX3: a no-op unless a profiler injects itself).  The reference has
// no tracing at all; nsys / ncu timelines show the loss forward / backward and the retrieval as
// named ranges.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

#define CLIBD_CHECK_CUDA(expr)                                                                    \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            ::clibd::set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " at " + \
                               __FILE__ + ":" + std::to_string(__LINE__));                        \
            return 2;                                                                             \
        }                                                                                         \
    } while (0)

#define CLIBD_REQUIRE(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            ::clibd::set_error(std::string("invalid argument: ") + (msg)); \
            return 1;                                              \
        }                                                          \
    } while (0)

#define CLIBD_KERNEL_CHECK()                   \
    do {                                       \
        ::clibd::count_launch();               \
        CLIBD_CHECK_CUDA(cudaGetLastError());  \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline int64_t round_up(int64_t a, int64_t b) { return ceil_div(a, b) * b; }

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p, int64_t i);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p, int64_t i) {
    return p[i];
}
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p, int64_t i) {
    return __bfloat162float(p[i]);
}
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p, int64_t i) {
    return __half2float(p[i]);
}

template <typename T>
__device__ __forceinline__ void store_from_float(T* p, int64_t i, float v);
template <>
__device__ __forceinline__ void store_from_float<float>(float* p, int64_t i, float v) {
    p[i] = v;
}
template <>
__device__ __forceinline__ void store_from_float<__nv_bfloat16>(__nv_bfloat16* p, int64_t i, float v) {
    p[i] = __float2bfloat16_rn(v);
}
template <>
__device__ __forceinline__ void store_from_float<__half>(__half* p, int64_t i, float v) {
    p[i] = __float2half_rn(v);
}

// 16-bit operand conversion selected at run time (fmt: 1 = bf16, 0 = f16, matching the
// tcgen05 instruction-descriptor format codes)
__device__ __forceinline__ uint16_t to_operand16(float v, int fmt) {
    return fmt ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
}
__device__ __forceinline__ uint32_t pack2_operand16(float lo, float hi, int fmt) {
    if (fmt) {
        __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&v);
    }
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}


// Device-resident logit scale: when an entry point has one (a tensor scale, see clibd_loss_forward_stats), every
// kernel of that call reads the scale from this pointer instead of its float argument -- no host read per step.
// Host side: thread-local, set for the duration of one extern "C" call.
const float* scale_dev_ptr();
struct ScaleScope {
    explicit ScaleScope(const float* p);
    ~ScaleScope();
    const float* prev_;
};
#ifdef __CUDACC__
__device__ __forceinline__ float eff_scale(float s, const float* dev) { return dev ? __ldg(dev) : s; }
#endif

// Shift of the softmax sums: every kernel forms e_ij = exp(S_ij - shift) ONCE and uses it for the row sums, the
// column sums and the gradient coefficients e_ij (c_i / r_i + c_j / q_j), which do not depend on the shift.  With
// unit rows |S_ij| <= s.  While s <= 43 the shift is s itself: every term is <= 1 and even S_ij = -s stays above the
// float32 underflow (2 s log2(e) < 126).  The reference never clamps its learnable logit_scale (simple_clip.py:32,61),
// so beyond 43 the shift stays at 43 (at s - 70 from s = 113): terms grow up to exp(70), sums of 2^18 of them stay
// finite, and the entries that underflow are those with S_ij < shift - 87, i.e. more than e^-87 below the largest
// representable term -- they cannot matter to a row whose best logit is representable.  A row (or column) whose sum
// underflows entirely (every cosine below (shift - 87) / s) is clamped to 1e-30 instead of producing log(0).
__host__ __device__ inline float softmax_shift(float s) { return s <= 43.f ? s : fmaxf(43.f, s - 70.f); }
constexpr float kMinSoftmaxSum = 1e-30f;

}  // namespace clibd
