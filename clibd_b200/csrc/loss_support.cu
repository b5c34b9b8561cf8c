// HBM-bound support kernels of the contrastive-loss path: row norms, label statistics,
// class sums (the label-matched positive term), 16-bit operand staging, deterministic
// reductions, loss finish, normalise-backward.  Everything here is O(N*d) or a label scan;
// the O(N^2 d) work lives in loss_tc.cu (tcgen05) and loss_simt.cu (fp32 CUDA cores).
//
// Reference semantics: bioscanclip/model/loss_func.py:19-22 (label matrix), :55-56 / :186-187
// (F.normalize), :65-69 / :195-200 (soft-target CE and the mean over the pair list).
#include <climits>

#include "common.cuh"
#include "loss_plan.h"

namespace clibd {

namespace {

constexpr int kThreads = 256;

template <typename T>
__global__ void row_inv_norm_kernel(const T* __restrict__ x, int64_t n, int64_t d, float* __restrict__ inv) {
    int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (row >= n) return;
    const T* xr = x + row * d;
    float ss = 0.f;
    for (int64_t k = lane; k < d; k += 32) {
        float v = load_as_float(xr, k);
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) inv[row] = 1.0f / fmaxf(sqrtf(ss), 1e-12f);
}

__global__ void label_stats_kernel(const int64_t* __restrict__ labels, int64_t N, int32_t* __restrict__ rep,
                                   float* __restrict__ cnt) {
    __shared__ int s_cnt[kThreads / 32];
    __shared__ int s_min[kThreads / 32];
    const int64_t i = blockIdx.x;
    const int64_t lab = labels[i];
    int c = 0;
    int mn = INT_MAX;
    for (int64_t j = threadIdx.x; j < N; j += kThreads) {
        if (labels[j] == lab) {
            ++c;
            mn = min(mn, static_cast<int>(j));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if ((threadIdx.x & 31) == 0) {
        s_cnt[threadIdx.x >> 5] = c;
        s_min[threadIdx.x >> 5] = mn;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int tc = 0, tm = INT_MAX;
        for (int w = 0; w < kThreads / 32; ++w) {
            tc += s_cnt[w];
            tm = min(tm, s_min[w]);
        }
        rep[i] = tm;
        cnt[i] = static_cast<float>(tc);
    }
}

__global__ void gscale_kernel(const float* __restrict__ cnt, int64_t N, int path, float* __restrict__ gscale) {
    __shared__ float s_max[kThreads / 32];
    float m = 1.f;
    for (int64_t j = threadIdx.x; j < N; j += kThreads) m = fmaxf(m, cnt[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w) m = fmaxf(m, s_max[w]);
        float g = 1.f;
        if (path == PATH_TC_F16) {
            // G~ = e*(u+v)*g <= 2*cmax*g must stay below the f16 maximum: 2*cmax*g <= 2^15
            g = exp2f(floorf(log2f(16384.f / m)));
        }
        gscale[0] = g;
        gscale[1] = 1.f / g;
    }
}

// One block per row r.  Only representatives (rep[r] == r) do work: they add, in index
// order, the normalised rows of their class.
template <typename T>
__global__ void class_sums_kernel(const T* __restrict__ x, const float* __restrict__ inv,
                                  const int32_t* __restrict__ rep, const float* __restrict__ cnt, int64_t N,
                                  int64_t d, float* __restrict__ Q) {
    const int64_t r = blockIdx.x;
    if (rep[r] != static_cast<int32_t>(r)) return;
    float* qr = Q + r * d;
    if (cnt[r] == 1.0f) {
        const float iv = inv[r];
        for (int64_t c = threadIdx.x; c < d; c += kThreads) qr[c] = load_as_float(x + r * d, c) * iv;
        return;
    }
    __shared__ int s_list[kThreads];
    __shared__ int s_woff[kThreads / 32 + 1];
    constexpr int MAXQ = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t cbase = 0; cbase < d; cbase += static_cast<int64_t>(kThreads) * MAXQ) {
        float acc[MAXQ];
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) acc[q] = 0.f;
        for (int64_t j0 = r; j0 < N; j0 += kThreads) {  // members have index >= r
            const int64_t j = j0 + threadIdx.x;
            const bool flag = (j < N) && (rep[j] == static_cast<int32_t>(r));
            const int total = __syncthreads_count(flag);
            if (total == 0) continue;
            const unsigned bal = __ballot_sync(0xffffffffu, flag);
            if (lane == 0) s_woff[warp + 1] = __popc(bal);
            __syncthreads();
            if (threadIdx.x == 0) {
                s_woff[0] = 0;
                for (int w = 0; w < kThreads / 32; ++w) s_woff[w + 1] += s_woff[w];
            }
            __syncthreads();
            if (flag) s_list[s_woff[warp] + __popc(bal & ((1u << lane) - 1u))] = static_cast<int>(j);
            __syncthreads();
            for (int m = 0; m < total; ++m) {
                const int64_t jj = s_list[m];
                const float iv = inv[jj];
#pragma unroll
                for (int q = 0; q < MAXQ; ++q) {
                    const int64_t c = cbase + threadIdx.x + static_cast<int64_t>(q) * kThreads;
                    if (c < d) acc[q] += load_as_float(x + jj * d, c) * iv;
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < MAXQ; ++q) {
            const int64_t c = cbase + threadIdx.x + static_cast<int64_t>(q) * kThreads;
            if (c < d) qr[c] = acc[q];
        }
    }
}

// 32x32 tile: xh[row][col] = cvt(x*inv) (zero for col >= d), xhT[col][row] (zero for row >= N)
template <typename T>
__global__ void make_operands_kernel(const T* __restrict__ x, const float* __restrict__ inv, int64_t N, int64_t d,
                                     int64_t dpad, int64_t npad, int fmt_bf16, uint16_t* __restrict__ xh,
                                     uint16_t* __restrict__ xhT) {
    __shared__ uint16_t tile[32][33];
    const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 32, r0 = static_cast<int64_t>(blockIdx.y) * 32;
    for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
        const int64_t row = r0 + dy, col = c0 + threadIdx.x;
        uint16_t h = 0;
        if (row < N && col < d) h = to_operand16(load_as_float(x + row * d, col) * (inv ? inv[row] : 1.f), fmt_bf16);
        tile[dy][threadIdx.x] = h;
        if (row < N && col < dpad) xh[row * dpad + col] = h;
    }
    __syncthreads();
    if (xhT != nullptr) {
        for (int dy = threadIdx.y; dy < 32; dy += blockDim.y) {
            const int64_t col = c0 + dy, row = r0 + threadIdx.x;
            if (col < dpad && row < npad) xhT[col * npad + row] = tile[threadIdx.x][dy];
        }
    }
}

template <typename T>
__global__ void pos_rows_kernel(const T* __restrict__ xa, const float* __restrict__ inv_a,
                                const float* __restrict__ Qb, const int32_t* __restrict__ rep, int64_t d,
                                int64_t row0, int64_t n, float* __restrict__ posrow) {
    int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int64_t gi = row0 + i;
    const T* xr = xa + gi * d;
    const float* q = Qb + static_cast<int64_t>(rep[gi]) * d;
    const float iv = inv_a[gi];
    float acc = 0.f;
    for (int64_t k = lane; k < d; k += 32) acc = fmaf(load_as_float(xr, k) * iv, q[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) posrow[i] = acc;
}

__global__ void reduce_parts_kernel(const float* __restrict__ part, int64_t parts, int64_t stride, int64_t len,
                                    float* __restrict__ out) {
    int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= len) return;
    float acc = 0.f;
    for (int64_t p = 0; p < parts; ++p) acc += part[p * stride + k];
    out[k] = acc;
}

constexpr int kRedBlocks = 256;

__device__ __forceinline__ double block_sum_double(double v, double* s_buf) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < kThreads / 32; ++w) t += s_buf[w];
    __syncthreads();
    return t;  // valid on thread 0
}

__global__ void sum_stage1_kernel(const float* __restrict__ in, int64_t len, double* __restrict__ red) {
    __shared__ double s_buf[kThreads / 32];
    double acc = 0.0;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; k < len;
         k += static_cast<int64_t>(kRedBlocks) * kThreads)
        acc += static_cast<double>(in[k]);
    double t = block_sum_double(acc, s_buf);
    if (threadIdx.x == 0) red[blockIdx.x] = t;
}

__global__ void sum_stage2_kernel(const double* __restrict__ red, double mul, double* __restrict__ out) {
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int b = 0; b < kRedBlocks; ++b) t += red[b];
        out[0] = t * mul;
    }
}

// term_p(k) = cnt[k] * (2*s + ln rowsum_p[k] + ln colsum_p[k]); also u = cnt/rowsum, v = cnt/colsum
__global__ void loss_finish_stage1_kernel(int64_t N, float scale, float w0, float w1, float w2,
                                          const float* __restrict__ cnt, const float* __restrict__ rowsum,
                                          const float* __restrict__ colsum, float* __restrict__ u,
                                          float* __restrict__ v, double* __restrict__ red) {
    __shared__ double s_buf[kThreads / 32];
    const float w[3] = {w0, w1, w2};
    double acc = 0.0;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; k < N;
         k += static_cast<int64_t>(kRedBlocks) * kThreads) {
        const float c = cnt[k];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            if (w[p] == 0.f) continue;
            const float rs = rowsum[p * N + k], cs = colsum[p * N + k];
            u[p * N + k] = c / rs;
            v[p * N + k] = c / cs;
            acc += static_cast<double>(w[p]) * static_cast<double>(c) *
                   (2.0 * static_cast<double>(scale) + log(static_cast<double>(rs)) + log(static_cast<double>(cs)));
        }
    }
    double t = block_sum_double(acc, s_buf);
    if (threadIdx.x == 0) red[blockIdx.x] = t;
}

__global__ void loss_finish_stage2_kernel(int64_t N, float scale, float w0, float w1, float w2,
                                          const double* __restrict__ red, const double* __restrict__ pos,
                                          float* __restrict__ loss_out) {
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int b = 0; b < kRedBlocks; ++b) t += red[b];
        const float w[3] = {w0, w1, w2};
        for (int p = 0; p < 3; ++p)
            if (w[p] != 0.f) t -= 2.0 * static_cast<double>(w[p]) * static_cast<double>(scale) * pos[p];
        loss_out[0] = static_cast<float>(t / static_cast<double>(N));
    }
}

template <typename T>
__global__ void normalize_bwd_kernel(NormBwdArgs a) {
    __shared__ float s_buf[kThreads / 32];
    __shared__ float s_dot;
    const int64_t i = blockIdx.x;
    const int64_t gi = a.row0 + i;
    const T* xr = reinterpret_cast<const T*>(a.x) + gi * a.d;
    const float iv = a.inv_norm[gi];
    const int64_t rp = a.rep[gi];
    const float k1 = a.scale / static_cast<float>(a.N);
    auto dxhat = [&](int64_t c) -> float {
        float g = 0.f;
        for (int s = 0; s < a.jsplit; ++s) g += a.dxh[(static_cast<int64_t>(s) * a.n + i) * a.d + c];
        float t = 0.f;
        if (a.Qp[0]) t = fmaf(a.wp[0], a.Qp[0][rp * a.d + c], t);
        if (a.Qp[1]) t = fmaf(a.wp[1], a.Qp[1][rp * a.d + c], t);
        return k1 * (g - 2.f * t);
    };
    float part = 0.f;
    for (int64_t c = threadIdx.x; c < a.d; c += kThreads) part = fmaf(load_as_float(xr, c) * iv, dxhat(c), part);
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) s_buf[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < kThreads / 32; ++w) t += s_buf[w];
        s_dot = t;
        a.dots[i] = t;
    }
    __syncthreads();
    if (a.dx == nullptr) return;
    const float dot = s_dot;
    T* dxr = reinterpret_cast<T*>(a.dx) + i * a.d;
    for (int64_t c = threadIdx.x; c < a.d; c += kThreads) {
        const float xh = load_as_float(xr, c) * iv;
        store_from_float(dxr, c, a.grad_scale * (dxhat(c) - xh * dot) * iv);
    }
}

}  // namespace

#define DISPATCH_DTYPE(dtype, ...)                                   \
    switch (dtype) {                                                 \
        case DT_F32: { using T = float; __VA_ARGS__; break; }        \
        case DT_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; } \
        case DT_F16: { using T = __half; __VA_ARGS__; break; }       \
        default: set_error("unsupported dtype code"); return 1;      \
    }

int launch_row_inv_norm(const void* x, int dtype, int64_t n, int64_t d, float* inv_norm, cudaStream_t s) {
    if (n == 0) return 0;
    const int64_t blocks = ceil_div(n * 32, kThreads);
    DISPATCH_DTYPE(dtype, (row_inv_norm_kernel<T><<<blocks, kThreads, 0, s>>>(static_cast<const T*>(x), n, d, inv_norm)));
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_label_stats(const int64_t* labels, int64_t N, int32_t* rep, float* cnt, cudaStream_t s) {
    if (N == 0) return 0;
    label_stats_kernel<<<N, kThreads, 0, s>>>(labels, N, rep, cnt);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_gscale(const float* cnt, int64_t N, int path, float* gscale, cudaStream_t s) {
    gscale_kernel<<<1, kThreads, 0, s>>>(cnt, N, path, gscale);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_class_sums(const void* x, int dtype, const float* inv_norm, const int32_t* rep, const float* cnt,
                      int64_t N, int64_t d, float* Q, cudaStream_t s) {
    if (N == 0) return 0;
    DISPATCH_DTYPE(dtype, (class_sums_kernel<T><<<N, kThreads, 0, s>>>(static_cast<const T*>(x), inv_norm, rep, cnt, N, d, Q)));
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_make_operands(const void* x, int dtype, const float* inv_norm, int64_t N, int64_t d, int64_t dpad,
                         int64_t npad, int fmt_bf16, void* xh, void* xhT, cudaStream_t s) {
    const int64_t rows = npad > N ? npad : N;
    dim3 grid(static_cast<unsigned>(ceil_div(dpad, 32)), static_cast<unsigned>(ceil_div(rows, 32)));
    dim3 block(32, 8);
    DISPATCH_DTYPE(dtype, (make_operands_kernel<T><<<grid, block, 0, s>>>(
                              static_cast<const T*>(x), inv_norm, N, d, dpad, npad, fmt_bf16,
                              static_cast<uint16_t*>(xh), static_cast<uint16_t*>(xhT))));
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_pos_rows(const void* xa, int dtype, const float* inv_a, const float* Qb, const int32_t* rep, int64_t d,
                    int64_t row0, int64_t n, float* posrow, cudaStream_t s) {
    if (n == 0) return 0;
    const int64_t blocks = ceil_div(n * 32, kThreads);
    DISPATCH_DTYPE(dtype, (pos_rows_kernel<T><<<blocks, kThreads, 0, s>>>(static_cast<const T*>(xa), inv_a, Qb, rep, d, row0, n, posrow)));
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_reduce_parts(const float* part, int64_t parts, int64_t stride, int64_t len, float* out, cudaStream_t s) {
    if (len == 0) return 0;
    reduce_parts_kernel<<<ceil_div(len, kThreads), kThreads, 0, s>>>(part, parts, stride, len, out);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_sum_to_double(const float* in, int64_t len, double mul, double* red, double* out, cudaStream_t s) {
    sum_stage1_kernel<<<kRedBlocks, kThreads, 0, s>>>(in, len, red);
    sum_stage2_kernel<<<1, 32, 0, s>>>(red, mul, out);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_loss_finish(int64_t N, float scale, const float w[3], const float* cnt, const float* rowsum,
                       const float* colsum, const double* pos, float* u, float* v, double* red, float* loss_out,
                       cudaStream_t s) {
    loss_finish_stage1_kernel<<<kRedBlocks, kThreads, 0, s>>>(N, scale, w[0], w[1], w[2], cnt, rowsum, colsum, u, v, red);
    loss_finish_stage2_kernel<<<1, 32, 0, s>>>(N, scale, w[0], w[1], w[2], red, pos, loss_out);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_normalize_bwd(const NormBwdArgs& a, cudaStream_t s) {
    if (a.n == 0) return 0;
    DISPATCH_DTYPE(a.dtype, (normalize_bwd_kernel<T><<<a.n, kThreads, 0, s>>>(a)));
    CLIBD_KERNEL_CHECK();
    return 0;
}

}  // namespace clibd
