// CUDA-core fp32 path of the contrastive loss (path code 0): exact-arithmetic tiles for
// fp32 inputs (BASELINE config 1: N=256, d=768, tolerance 1e-5, which 16-bit tensor-core
// operands cannot meet) and for any N / d the tcgen05 path does not take.  Same fused
// structure as the tensor-core path: the N x N logits only ever exist as register tiles.
//
//   forward : S = s * Xhat Yhat^T tile -> e = exp(S - s) -> row / column partial sums
//             (loss_func.py:59-66 restated as CE = (1/N) sum_i [c_i LSE_i - sum_j T_ij S_ij])
//   backward: recompute the S tile, G~ = e * (rowcoef_i + colcoef_j), dXhat += G~ Yhat
#include "common.cuh"
#include "loss_plan.h"

namespace clibd {
namespace {

constexpr int T64 = SIMT_T;
constexpr int KT = 32;

// Both kernels are latency-bound at the batches this path serves (N <= 1024; BASELINE config 1 is N = 256: 16 forward
// tiles): every thread keeps the NEXT K chunk of both operands in registers while the current one is consumed from
// shared memory, so the global-load latency of a chunk hides behind the FMAs of the one before.
template <typename T>
__global__ void __launch_bounds__(256)
simt_fwd_kernel(const T* __restrict__ xa, const T* __restrict__ xb, const float* __restrict__ inv_a,
                const float* __restrict__ inv_b, int64_t N, int64_t d, int64_t row0, int64_t n, float scale,
                float* __restrict__ rowpart, float* __restrict__ colpart, int self_mask,
                const float* __restrict__ scale_dev) {
    scale = eff_scale(scale, scale_dev);
    __shared__ float As[KT][T64 + 1];
    __shared__ float Bs[KT][T64 + 1];
    __shared__ float red[T64][17];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // rows ty * 4 + i, columns tx + 16 * j (conflict-free Bs reads)
    const int64_t ct = blockIdx.x, rt = blockIdx.y;
    const int64_t lrow0 = rt * T64;          // local row base
    const int64_t col0 = ct * T64;           // global column base
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    // loader: element q of this thread = (tile row lr8 + 8 q, k = lane); a warp reads 32 consecutive k of one row
    const int lk = tid & 31, lr8 = tid >> 5;
    const T* pa_row[8];
    const T* pb_row[8];
    float ia[8], ib[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int64_t lr = lrow0 + lr8 + 8 * q, gc = col0 + lr8 + 8 * q;
        const bool va = lr < n, vb = gc < N;
        pa_row[q] = va ? xa + (row0 + lr) * d : nullptr;
        pb_row[q] = vb ? xb + gc * d : nullptr;
        ia[q] = va ? inv_a[row0 + lr] : 0.f;
        ib[q] = vb ? inv_b[gc] : 0.f;
    }
    float pa[8], pb[8];
    auto fetch = [&](int64_t k0) {
        const int64_t gk = k0 + lk;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            pa[q] = (pa_row[q] != nullptr && gk < d) ? load_as_float(pa_row[q], gk) * ia[q] : 0.f;
            pb[q] = (pb_row[q] != nullptr && gk < d) ? load_as_float(pb_row[q], gk) * ib[q] : 0.f;
        }
    };
    fetch(0);
    for (int64_t k0 = 0; k0 < d; k0 += KT) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            As[lk][lr8 + 8 * q] = pa[q];
            Bs[lk][lr8 + 8 * q] = pb[q];
        }
        __syncthreads();
        if (k0 + KT < d) fetch(k0 + KT);
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    // e = exp(S - shift); |S| <= s for unit vectors so the fixed shift is safe
    float rsum[4] = {0.f, 0.f, 0.f, 0.f}, csum[4] = {0.f, 0.f, 0.f, 0.f};
    const float shift = softmax_shift(scale);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool valid = (lrow0 + ty * 4 + i < n) && (col0 + tx + 16 * j < N) &&
                               !(self_mask && row0 + lrow0 + ty * 4 + i == col0 + tx + 16 * j);
            const float e = valid ? expf(fmaf(scale, acc[i][j], -shift)) : 0.f;
            rsum[i] += e;
            csum[j] += e;
        }
    // row sums: reduce over the 16 tx lanes in a fixed order
#pragma unroll
    for (int i = 0; i < 4; ++i) red[ty * 4 + i][tx] = rsum[i];
    __syncthreads();
    if (tid < T64) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) t += red[tid][k];
        if (lrow0 + tid < n) rowpart[ct * n + lrow0 + tid] = t;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) red[tx + 16 * j][ty] = csum[j];
    __syncthreads();
    if (tid < T64) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 16; ++k) t += red[tid][k];
        if (col0 + tid < N) colpart[rt * N + col0 + tid] = t;
    }
}

constexpr int BR = SIMT_BR;  // rows per block
constexpr int BJ = 32;       // columns per step
constexpr int DQ = 3;        // 256-wide column groups of dX per pass (768 columns)

template <typename T>
__global__ void __launch_bounds__(256)
simt_bwd_kernel(const T* __restrict__ x, const T* __restrict__ y, const float* __restrict__ inv_x,
                const float* __restrict__ inv_y, int64_t N, int64_t d, int64_t row0, int64_t n, float scale,
                const float* __restrict__ rowcoef, const float* __restrict__ colcoef, float weight, int accumulate,
                float* __restrict__ dxh, int self_mask, const float* __restrict__ scale_dev, int64_t steps_per_split) {
    scale = eff_scale(scale, scale_dev);
    // blockIdx.z = column split: this block sweeps the column steps [z, z + 1) * steps_per_split and writes the partial
    // dxh[z] (summed by normalize_bwd in split order).  Small batches have few 32-row blocks; without the split a
    // 256-row batch ran on 8 of 148 SMs (measured 716 us per sweep at N = 512).
    const int64_t j_begin = static_cast<int64_t>(blockIdx.z) * steps_per_split * BJ;
    const int64_t j_end = min(N, j_begin + steps_per_split * BJ);
    dxh += static_cast<int64_t>(blockIdx.z) * n * d;
    __shared__ float Xs[BR][BJ + 1];
    __shared__ float Ys[BJ][BJ + 1];
    __shared__ float Gs[BR][BJ + 1];
    const int tid = threadIdx.x;
    const int64_t lr0 = static_cast<int64_t>(blockIdx.x) * BR;
    const int64_t dbase = static_cast<int64_t>(blockIdx.y) * 256 * DQ;
    const int srow = tid >> 3, scol = (tid & 7) * 4;  // this thread's 1x4 strip of the S tile
    float acc[BR][DQ];
#pragma unroll
    for (int r = 0; r < BR; ++r)
#pragma unroll
        for (int q = 0; q < DQ; ++q) acc[r][q] = 0.f;
    const float shift = softmax_shift(scale);

    // loader of the S operands: element q of this thread = (tile row lr8 + 8 q, k = lane)
    const int lk = tid & 31, lr8 = tid >> 5;
    const T* px_row[4];
    float ix[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t lr = lr0 + lr8 + 8 * q;
        px_row[q] = lr < n ? x + (row0 + lr) * d : nullptr;
        ix[q] = lr < n ? inv_x[row0 + lr] : 0.f;
    }

    for (int64_t j0 = j_begin; j0 < j_end; j0 += BJ) {
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        const T* py_row[4];
        float iy[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t gj = j0 + lr8 + 8 * q;
            py_row[q] = gj < N ? y + gj * d : nullptr;
            iy[q] = gj < N ? inv_y[gj] : 0.f;
        }
        float pxv[4], pyv[4];
        auto fetch = [&](int64_t k0) {  // the next K chunk waits in registers while the current one is consumed
            const int64_t gk = k0 + lk;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                pxv[q] = (px_row[q] != nullptr && gk < d) ? load_as_float(px_row[q], gk) * ix[q] : 0.f;
                pyv[q] = (py_row[q] != nullptr && gk < d) ? load_as_float(py_row[q], gk) * iy[q] : 0.f;
            }
        };
        fetch(0);
        for (int64_t k0 = 0; k0 < d; k0 += BJ) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                Xs[lr8 + 8 * q][lk] = pxv[q];
                Ys[lr8 + 8 * q][lk] = pyv[q];
            }
            __syncthreads();
            if (k0 + BJ < d) fetch(k0 + BJ);
#pragma unroll
            for (int k = 0; k < BJ; ++k) {
                const float xv = Xs[srow][k];
#pragma unroll
                for (int c = 0; c < 4; ++c) s4[c] = fmaf(xv, Ys[scol + c][k], s4[c]);
            }
            __syncthreads();
        }
        {
            const int64_t lr = lr0 + srow;
            const float rc = (lr < n) ? rowcoef[row0 + lr] : 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int64_t gj = j0 + scol + c;
                float g = 0.f;
                if (lr < n && gj < N && !(self_mask && row0 + lr == gj))
                    g = expf(fmaf(scale, s4[c], -shift)) * (rc + colcoef[gj]);
                Gs[srow][scol + c] = g;
            }
        }
        __syncthreads();
        // dXhat rows += G~ Yhat: four column rows of Yhat in flight per thread (G~ of a column that does not exist is 0)
        const int jmax = static_cast<int>(min(static_cast<int64_t>(BJ), N - j0));
        for (int jj = 0; jj < jmax; jj += 4) {
            float yv[4][DQ];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t gj = j0 + jj + u;
                const bool jv = jj + u < jmax;
                const float iv = jv ? inv_y[gj] : 0.f;
#pragma unroll
                for (int q = 0; q < DQ; ++q) {
                    const int64_t c = dbase + tid + 256 * q;
                    yv[u][q] = (jv && c < d) ? load_as_float(y + gj * d, c) * iv : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int r = 0; r < BR; ++r) {
                    const float g = Gs[r][jj + u];
#pragma unroll
                    for (int q = 0; q < DQ; ++q) acc[r][q] = fmaf(g, yv[u][q], acc[r][q]);
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < BR; ++r) {
        const int64_t lr = lr0 + r;
        if (lr >= n) continue;
#pragma unroll
        for (int q = 0; q < DQ; ++q) {
            const int64_t c = dbase + tid + 256 * q;
            if (c < d) {
                float* p = dxh + lr * d + c;
                *p = (accumulate ? *p : 0.f) + weight * acc[r][q];
            }
        }
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

int simt_forward_pair(const void* xa, const void* xb, int dtype, const float* inv_a, const float* inv_b, int64_t N,
                      int64_t d, int64_t row0, int64_t n, float scale, float* rowpart, float* colpart,
                      cudaStream_t s, int self_mask) {
    if (n == 0 || N == 0) return 0;
    dim3 grid(static_cast<unsigned>(ceil_div(N, T64)), static_cast<unsigned>(ceil_div(n, T64)));
    DISPATCH_DTYPE(dtype, (simt_fwd_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(xa), static_cast<const T*>(xb),
                                                                  inv_a, inv_b, N, d, row0, n, scale, rowpart, colpart,
                                                                  self_mask, scale_dev_ptr())));
    CLIBD_KERNEL_CHECK();
    return 0;
}

int simt_backward_rows(const void* x, const void* y, int dtype, const float* inv_x, const float* inv_y, int64_t N,
                       int64_t d, int64_t row0, int64_t n, float scale, const float* rowcoef, const float* colcoef,
                       float weight, int accumulate, float* dxh, cudaStream_t s, int self_mask, int jsplit) {
    if (n == 0 || N == 0) return 0;
    if (jsplit < 1) jsplit = 1;
    const int64_t steps_per_split = ceil_div(ceil_div(N, BJ), jsplit);  // a split past the last step writes zeros
    dim3 grid(static_cast<unsigned>(ceil_div(n, BR)), static_cast<unsigned>(ceil_div(d, 256 * DQ)),
              static_cast<unsigned>(jsplit));
    DISPATCH_DTYPE(dtype, (simt_bwd_kernel<T><<<grid, 256, 0, s>>>(static_cast<const T*>(x), static_cast<const T*>(y),
                                                                  inv_x, inv_y, N, d, row0, n, scale, rowcoef, colcoef,
                                                                  weight, accumulate, dxh, self_mask, scale_dev_ptr(),
                                                                  steps_per_split)));
    CLIBD_KERNEL_CHECK();
    return 0;
}

}  // namespace clibd
