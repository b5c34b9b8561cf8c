// Kernels of the two seams next to the hot path (SURVEY.md section 8 f, ranks 2 and 3).  All are HBM-bound
// single-pass kernels: every input byte is read once, every output byte written once.
//
//  * embed_append_kernel   : F.normalize(x, dim=-1) of one batch of encoder outputs written as float32 rows into a
//                            preallocated device store (replaces `F.normalize(out).cpu().tolist()` + `np.array`,
//                            bioscanclip/epoch/inference_epoch.py:96-101,108-119).
//  * softmax_mean_fwd/bwd  : the BarcodeBERT head `logits.softmax(dim=-1).mean(dim=1)` on [n, T, C] logits
//                            (bioscanclip/model/dna_encoder.py:137) and its backward, one warp per (sample, token)
//                            row with the row held in registers; the [n, T, C] probability tensor never exists.
#include "common.cuh"

namespace clibd {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

// ---- 16-byte chunk access: a chunk is 4 floats or 8 16-bit values ------------------------------------
template <typename T>
struct Chunk;
template <>
struct Chunk<float> {
    static constexpr int E = 4;
    __device__ static void load(const float* p, float (&v)[4]) {
        const float4 a = *reinterpret_cast<const float4*>(p);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
    __device__ static void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct Chunk<__nv_bfloat16> {
    static constexpr int E = 8;
    __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 r = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    }
    __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 t = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <>
struct Chunk<__half> {
    static constexpr int E = 8;
    __device__ static void load(const __half* p, float (&v)[8]) {
        const uint4 r = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            v[2 * i] = f.x;
            v[2 * i + 1] = f.y;
        }
    }
    __device__ static void store(__half* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __half2 t = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&t);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Row of C values spread over a warp: lane l owns chunks l, l+32, ... (NCH of them, E values each).
// Returns the row in registers; chunks past the row end hold `fill`.
template <typename T, int NCH>
__device__ __forceinline__ void load_row(const T* row, int64_t C, int lane, float fill, float (&v)[NCH][Chunk<T>::E]) {
    constexpr int E = Chunk<T>::E;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int64_t c = (static_cast<int64_t>(i) * 32 + lane) * E;
        if (c < C) {
            Chunk<T>::load(row + c, v[i]);
        } else {
#pragma unroll
            for (int e = 0; e < E; ++e) v[i][e] = fill;
        }
    }
}

// ---- embedding hand-off ---------------------------------------------------------------------------------
template <typename T, int NCH>
__global__ void embed_append_kernel(const T* __restrict__ x, int64_t n, int64_t d, float* __restrict__ store,
                                    int64_t store_ld, int64_t row0) {
    constexpr int E = Chunk<T>::E;
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    float v[NCH][E];
    load_row<T, NCH>(x + row * d, d, lane, 0.f, v);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int e = 0; e < E; ++e) ss = fmaf(v[i][e], v[i][e], ss);
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(||x||_2, eps)
    float* out = store + (row0 + row) * store_ld;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int64_t c = (static_cast<int64_t>(i) * 32 + lane) * E;
        if (c < d) {
#pragma unroll
            for (int e = 0; e < E; e += 4)
                *reinterpret_cast<float4*>(out + c + e) =
                    make_float4(v[i][e] * inv, v[i][e + 1] * inv, v[i][e + 2] * inv, v[i][e + 3] * inv);
        }
    }
}

// any d / alignment: one warp per row, two passes over the row (the second one hits L1/L2)
template <typename T>
__global__ void embed_append_generic_kernel(const T* __restrict__ x, int64_t n, int64_t d, float* __restrict__ store,
                                            int64_t store_ld, int64_t row0) {
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const T* xr = x + row * d;
    float ss = 0.f;
    for (int64_t k = lane; k < d; k += 32) {
        const float a = load_as_float(xr, k);
        ss = fmaf(a, a, ss);
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
    float* out = store + (row0 + row) * store_ld;
    for (int64_t k = lane; k < d; k += 32) out[k] = load_as_float(xr, k) * inv;
}

// ---- softmax over the last dim, mean over the token dim ---------------------------------------------------
// Block = one sample; its 8 warps take token rows t = w, w+8, ...; each lane accumulates the probabilities of
// its own columns over the warp's rows, then the 8 warps are summed through shared memory in a fixed order.
template <typename T, int NCH>
__global__ void __launch_bounds__(kThreads) softmax_mean_fwd_kernel(const T* __restrict__ logits, int64_t T_tok, int64_t C,
                                                                    T* __restrict__ out) {
    constexpr int E = Chunk<T>::E;
    extern __shared__ float sm_acc[];  // [kWarps][C]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = blockIdx.x;
    const T* base = logits + b * T_tok * C;
    float acc[NCH][E];
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int e = 0; e < E; ++e) acc[i][e] = 0.f;
    for (int64_t t = warp; t < T_tok; t += kWarps) {
        float v[NCH][E];
        load_row<T, NCH>(base + t * C, C, lane, -INFINITY, v);
        float m = -INFINITY;
#pragma unroll
        for (int i = 0; i < NCH; ++i)
#pragma unroll
            for (int e = 0; e < E; ++e) m = fmaxf(m, v[i][e]);
        m = warp_max(m);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < NCH; ++i)
#pragma unroll
            for (int e = 0; e < E; ++e) {
                v[i][e] = __expf(v[i][e] - m);
                s += v[i][e];
            }
        s = warp_sum(s);
        const float r = 1.0f / s;
#pragma unroll
        for (int i = 0; i < NCH; ++i)
#pragma unroll
            for (int e = 0; e < E; ++e) acc[i][e] = fmaf(v[i][e], r, acc[i][e]);
    }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int64_t c = (static_cast<int64_t>(i) * 32 + lane) * E;
        if (c < C) {
#pragma unroll
            for (int e = 0; e < E; ++e) sm_acc[warp * C + c + e] = acc[i][e];
        }
    }
    __syncthreads();
    const float invT = 1.0f / static_cast<float>(T_tok);
    for (int64_t c = threadIdx.x; c < C; c += kThreads) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += sm_acc[w * C + c];
        store_from_float(out, b * C + c, s * invT);
    }
}

// dlogits[b,t,c] = (1/T) p[b,t,c] (g[b,c] - sum_c' g[b,c'] p[b,t,c'])
template <typename T, int NCH>
__global__ void __launch_bounds__(kThreads) softmax_mean_bwd_kernel(const T* __restrict__ logits, const T* __restrict__ gout,
                                                                    int64_t n, int64_t T_tok, int64_t C,
                                                                    T* __restrict__ dlogits) {
    constexpr int E = Chunk<T>::E;
    const int lane = threadIdx.x & 31;
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;  // (b, t) flattened
    if (row >= n * T_tok) return;
    const int64_t b = row / T_tok;
    float v[NCH][E], g[NCH][E];
    load_row<T, NCH>(logits + row * C, C, lane, -INFINITY, v);
    load_row<T, NCH>(gout + b * C, C, lane, 0.f, g);
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int e = 0; e < E; ++e) m = fmaxf(m, v[i][e]);
    m = warp_max(m);
    float s = 0.f, dot = 0.f;
#pragma unroll
    for (int i = 0; i < NCH; ++i)
#pragma unroll
        for (int e = 0; e < E; ++e) {
            v[i][e] = __expf(v[i][e] - m);
            s += v[i][e];
            dot = fmaf(v[i][e], g[i][e], dot);
        }
    s = warp_sum(s);
    dot = warp_sum(dot);
    const float r = 1.0f / s;
    const float gbar = dot * r;  // sum_c g p
    const float k = r / static_cast<float>(T_tok);
    T* out = dlogits + row * C;
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
        const int64_t c = (static_cast<int64_t>(i) * 32 + lane) * E;
        if (c < C) {
            float o[E];
#pragma unroll
            for (int e = 0; e < E; ++e) o[e] = v[i][e] * k * (g[i][e] - gbar);
            Chunk<T>::store(out + c, o);
        }
    }
}

// generic shapes (C not a multiple of the chunk width, unaligned rows, C > 1024): one warp per row, the row is
// re-read from L1/L2 for each pass
template <typename T>
__global__ void softmax_mean_fwd_generic_kernel(const T* __restrict__ logits, int64_t T_tok, int64_t C, T* __restrict__ out,
                                                float* __restrict__ part /* [n, kWarps, C] */) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = blockIdx.x;
    float* acc = part + (b * kWarps + warp) * C;
    for (int64_t c = lane; c < C; c += 32) acc[c] = 0.f;
    for (int64_t t = warp; t < T_tok; t += kWarps) {
        const T* row = logits + (b * T_tok + t) * C;
        float m = -INFINITY;
        for (int64_t c = lane; c < C; c += 32) m = fmaxf(m, load_as_float(row, c));
        m = warp_max(m);
        float s = 0.f;
        for (int64_t c = lane; c < C; c += 32) s += __expf(load_as_float(row, c) - m);
        s = warp_sum(s);
        const float r = 1.0f / s;
        for (int64_t c = lane; c < C; c += 32) acc[c] = fmaf(__expf(load_as_float(row, c) - m), r, acc[c]);
    }
    __syncthreads();
    const float invT = 1.0f / static_cast<float>(T_tok);
    for (int64_t c = threadIdx.x; c < C; c += kThreads) {
        float s = 0.f;
        for (int w = 0; w < kWarps; ++w) s += part[(b * kWarps + w) * C + c];
        store_from_float(out, b * C + c, s * invT);
    }
}

template <typename T>
__global__ void softmax_mean_bwd_generic_kernel(const T* __restrict__ logits, const T* __restrict__ gout, int64_t n,
                                                int64_t T_tok, int64_t C, T* __restrict__ dlogits) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= n * T_tok) return;
    const int64_t b = row / T_tok;
    const T* x = logits + row * C;
    const T* g = gout + b * C;
    float m = -INFINITY;
    for (int64_t c = lane; c < C; c += 32) m = fmaxf(m, load_as_float(x, c));
    m = warp_max(m);
    float s = 0.f, dot = 0.f;
    for (int64_t c = lane; c < C; c += 32) {
        const float e = __expf(load_as_float(x, c) - m);
        s += e;
        dot = fmaf(e, load_as_float(g, c), dot);
    }
    s = warp_sum(s);
    dot = warp_sum(dot);
    const float r = 1.0f / s;
    const float gbar = dot * r;
    const float k = r / static_cast<float>(T_tok);
    for (int64_t c = lane; c < C; c += 32)
        store_from_float(dlogits, row * C + c, __expf(load_as_float(x, c) - m) * k * (load_as_float(g, c) - gbar));
}

template <typename T>
bool vec_ok(const void* p, int64_t width) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && width % Chunk<T>::E == 0;
}

template <typename T>
int nch_for(int64_t width) {  // chunks per lane needed to hold one row in a warp's registers (0: does not fit)
    const int64_t chunks = ceil_div(width, Chunk<T>::E);
    const int64_t per_lane = ceil_div(chunks, 32);
    const int64_t max_nch = 32 / Chunk<T>::E;  // 32 values per lane: rows of up to 1024 values
    return per_lane <= max_nch ? static_cast<int>(per_lane) : 0;
}

}  // namespace

#define DISPATCH_DTYPE(dtype, ...)                                     \
    switch (dtype) {                                                   \
        case DT_F32: { using T = float; __VA_ARGS__; break; }          \
        case DT_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; } \
        case DT_F16: { using T = __half; __VA_ARGS__; break; }         \
        default: set_error("unsupported dtype code"); return 1;        \
    }

// NCH values instantiated: 16-bit rows use 1..4 chunks per lane, float rows 1..8
#define DISPATCH_NCH(nch, ...)                                   \
    switch (nch) {                                               \
        case 1: { constexpr int NCH = 1; __VA_ARGS__; break; }   \
        case 2: { constexpr int NCH = 2; __VA_ARGS__; break; }   \
        case 3: { constexpr int NCH = 3; __VA_ARGS__; break; }   \
        case 4: { constexpr int NCH = 4; __VA_ARGS__; break; }   \
        case 5: { constexpr int NCH = 5; __VA_ARGS__; break; }   \
        case 6: { constexpr int NCH = 6; __VA_ARGS__; break; }   \
        case 7: { constexpr int NCH = 7; __VA_ARGS__; break; }   \
        default: { constexpr int NCH = 8; __VA_ARGS__; break; }  \
    }

int launch_embed_append(const void* x, int dtype, int64_t n, int64_t d, float* store, int64_t store_ld, int64_t row0,
                        cudaStream_t s) {
    if (n == 0) return 0;
    const int64_t blocks = ceil_div(n * 32, kThreads);
    DISPATCH_DTYPE(dtype, {
        const int nch = nch_for<T>(d);
        const bool vec = nch > 0 && vec_ok<T>(x, d) && (reinterpret_cast<uintptr_t>(store) & 15u) == 0 &&
                         store_ld % 4 == 0 && d % 4 == 0;
        if (vec) {
            DISPATCH_NCH(nch, (embed_append_kernel<T, (NCH * Chunk<T>::E <= 32 ? NCH : 1)><<<blocks, kThreads, 0, s>>>(
                                  static_cast<const T*>(x), n, d, store, store_ld, row0)));
        } else {
            embed_append_generic_kernel<T><<<blocks, kThreads, 0, s>>>(static_cast<const T*>(x), n, d, store, store_ld, row0);
        }
    });
    CLIBD_KERNEL_CHECK();
    return 0;
}

int64_t softmax_mean_scratch_bytes(int64_t n, int64_t C, int dtype) {
    // only the generic forward kernel needs scratch (per-warp partial sums)
    const int64_t esize = dtype == DT_F32 ? 4 : 2;
    const int64_t echunk = 16 / esize;
    const bool fits = C % echunk == 0 && ceil_div(ceil_div(C, echunk), 32) <= 32 / echunk && C * kWarps * 4 <= 200 * 1024;
    return fits ? 0 : n * kWarps * C * 4;
}

int launch_softmax_mean_fwd(const void* logits, int dtype, int64_t n, int64_t T_tok, int64_t C, void* out, void* scratch,
                            int64_t scratch_bytes, cudaStream_t s) {
    if (n == 0) return 0;
    DISPATCH_DTYPE(dtype, {
        const int nch = nch_for<T>(C);
        const size_t smem = static_cast<size_t>(kWarps) * C * sizeof(float);
        const bool vec = nch > 0 && vec_ok<T>(logits, C) && smem <= 200 * 1024;
        if (vec) {
            DISPATCH_NCH(nch, {
                auto kern = softmax_mean_fwd_kernel<T, (NCH * Chunk<T>::E <= 32 ? NCH : 1)>;
                if (smem > 48 * 1024)
                    CLIBD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
                kern<<<static_cast<unsigned>(n), kThreads, smem, s>>>(static_cast<const T*>(logits), T_tok, C, static_cast<T*>(out));
            });
        } else {
            CLIBD_REQUIRE(scratch && scratch_bytes >= n * kWarps * C * 4, "softmax_mean: scratch too small");
            softmax_mean_fwd_generic_kernel<T><<<static_cast<unsigned>(n), kThreads, 0, s>>>(
                static_cast<const T*>(logits), T_tok, C, static_cast<T*>(out), static_cast<float*>(scratch));
        }
    });
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_softmax_mean_bwd(const void* logits, const void* gout, int dtype, int64_t n, int64_t T_tok, int64_t C,
                            void* dlogits, cudaStream_t s) {
    if (n == 0) return 0;
    const int64_t blocks = ceil_div(n * T_tok * 32, kThreads);
    DISPATCH_DTYPE(dtype, {
        const int nch = nch_for<T>(C);
        const bool vec = nch > 0 && vec_ok<T>(logits, C) && vec_ok<T>(gout, C) && vec_ok<T>(dlogits, C);
        if (vec) {
            DISPATCH_NCH(nch, (softmax_mean_bwd_kernel<T, (NCH * Chunk<T>::E <= 32 ? NCH : 1)><<<blocks, kThreads, 0, s>>>(
                                  static_cast<const T*>(logits), static_cast<const T*>(gout), n, T_tok, C,
                                  static_cast<T*>(dlogits))));
        } else {
            softmax_mean_bwd_generic_kernel<T><<<blocks, kThreads, 0, s>>>(static_cast<const T*>(logits),
                                                                          static_cast<const T*>(gout), n, T_tok, C,
                                                                          static_cast<T*>(dlogits));
        }
    });
    CLIBD_KERNEL_CHECK();
    return 0;
}

}  // namespace clibd

// ---- C ABI ---------------------------------------------------------------------------------------------------
#include "../../include/clibd_b200.h"

using namespace clibd;

extern "C" {

int clibd_embed_append(const void* x, int dtype, int64_t n, int64_t d, float* store, int64_t store_rows,
                       int64_t store_ld, int64_t row_offset, clibd_stream_t stream) {
    CLIBD_REQUIRE(x && store && n >= 0 && d > 0, "null pointer or bad shape");
    CLIBD_REQUIRE(store_ld >= d && row_offset >= 0 && row_offset + n <= store_rows, "rows do not fit in the store");
    return launch_embed_append(x, dtype, n, d, store, store_ld, row_offset, stream);
}

int64_t clibd_softmax_mean_scratch_bytes(int64_t n, int64_t tokens, int64_t classes, int dtype) {
    if (n < 0 || tokens <= 0 || classes <= 0 || dtype < 0 || dtype > 2) return -1;
    return softmax_mean_scratch_bytes(n, classes, dtype);
}

int clibd_softmax_mean_forward(const void* logits, int dtype, int64_t n, int64_t tokens, int64_t classes, void* out,
                               void* scratch, int64_t scratch_bytes, clibd_stream_t stream) {
    CLIBD_REQUIRE(logits && out && n >= 0 && tokens > 0 && classes > 0, "null pointer or bad shape");
    return launch_softmax_mean_fwd(logits, dtype, n, tokens, classes, out, scratch, scratch_bytes, stream);
}

int clibd_softmax_mean_backward(const void* logits, const void* grad_out, int dtype, int64_t n, int64_t tokens,
                                int64_t classes, void* grad_logits, clibd_stream_t stream) {
    CLIBD_REQUIRE(logits && grad_out && grad_logits && n >= 0 && tokens > 0 && classes > 0, "null pointer or bad shape");
    return launch_softmax_mean_bwd(logits, grad_out, dtype, n, tokens, classes, grad_logits, stream);
}

}  // extern "C"
