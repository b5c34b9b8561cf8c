// HBM-bound support kernels of the contrastive-loss path: row norms, label statistics,
// class sums (the label-matched positive term), 16-bit operand staging, deterministic
// reductions, loss finish, normalise-backward.  Everything here is O(N*d) or a label scan;
// the O(N^2 d) work lives in loss_tc.cu (tcgen05) and loss_simt.cu (fp32 CUDA cores).
//
// Reference semantics: bioscanclip/model/loss_func.py:19-22 (label matrix), :55-56 / :186-187
// (F.normalize), :65-69 / :195-200 (soft-target CE and the mean over the pair list).
#include <climits>
#include <string>

#include <cub/device/device_radix_sort.cuh>

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

// ---- 8-wide row access helpers (16-byte vectors when the row layout allows it) ---------------------
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
}
template <>
__device__ __forceinline__ void load8<__half>(const __half* p, float (&v)[8]) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
        v[2 * i] = f.x;
        v[2 * i + 1] = f.y;
    }
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = r;
}
template <>
__device__ __forceinline__ void store8<__half>(__half* p, const float (&v)[8]) {
    uint4 r;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = r;
}
template <typename T>
__host__ __device__ inline bool rows_vec8_ok(const void* base, int64_t d) {
    return (d % 8 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
}

// ---- label statistics: rep[i] = lowest index with the label of i, cnt[i] = class size --------------
// (construct_label_metrix, loss_func.py:19-22, reduced to what the fused loss needs.)  Open-addressing hash
// table keyed by label: a slot is claimed by storing the index of the first row that reaches it (the key is
// then labels[owner], immutable), class minimum / count by atomicMin / atomicAdd -- both order-independent,
// so the result is deterministic.  O(N) instead of the N x N comparison.
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void hash_init_kernel(int32_t* __restrict__ own, int32_t* __restrict__ hmin, int32_t* __restrict__ hcnt,
                                 int64_t H) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < H) {
        own[k] = -1;
        hmin[k] = INT_MAX;
        hcnt[k] = 0;
    }
}

__global__ void hash_insert_kernel(const int64_t* __restrict__ labels, int64_t N, int64_t H, int32_t* own,
                                   int32_t* hmin, int32_t* hcnt, int32_t* __restrict__ slot_of) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int64_t lab = labels[i];
    int64_t slot = static_cast<int64_t>(mix64(static_cast<uint64_t>(lab)) & static_cast<uint64_t>(H - 1));
    while (true) {
        int32_t o = *reinterpret_cast<volatile int32_t*>(own + slot);
        if (o < 0) {
            const int32_t prev = atomicCAS(own + slot, -1, static_cast<int32_t>(i));
            o = prev < 0 ? static_cast<int32_t>(i) : prev;
        }
        if (labels[o] == lab) break;
        slot = (slot + 1) & (H - 1);
    }
    atomicMin(hmin + slot, static_cast<int32_t>(i));
    atomicAdd(hcnt + slot, 1);
    slot_of[i] = static_cast<int32_t>(slot);
}

__global__ void hash_lookup_kernel(int64_t N, const int32_t* __restrict__ hmin, const int32_t* __restrict__ hcnt,
                                   int32_t* __restrict__ rep /* in: slot, out: representative */,
                                   float* __restrict__ cnt, int32_t* __restrict__ iota) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int32_t slot = rep[i];
    rep[i] = hmin[slot];
    cnt[i] = static_cast<float>(hcnt[slot]);
    iota[i] = static_cast<int32_t>(i);
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

// Class sums over the rows sorted by (representative, index): one warp per class segment, members added in
// index order (fixed summation order).  skey = sorted representatives, sidx = the rows in that order.
template <typename T, bool VEC>
__global__ void class_sums_kernel(const T* __restrict__ x, const float* __restrict__ inv,
                                  const int32_t* __restrict__ skey, const int32_t* __restrict__ sidx,
                                  const float* __restrict__ cnt, int64_t N, int64_t d, int64_t row0, int64_t n,
                                  float* __restrict__ Q, const float* __restrict__ lam2) {
    const int64_t p = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= N) return;
    const int32_t r = skey[p];
    if (p > 0 && skey[p - 1] == r) return;  // not the head of its segment
    const int members = static_cast<int>(cnt[r]);
    if (n < N) {  // row-sharded: only classes with a member among the local rows are ever read on this rank
        bool need = false;
        for (int m = lane; m < members; m += 32) {
            const int64_t j = sidx[p + m];
            need |= (j >= row0 && j < row0 + n);
        }
        if (!__any_sync(0xffffffffu, need)) return;
    }
    float* qr = Q + static_cast<int64_t>(r) * d;
    if constexpr (VEC) {
        for (int64_t c = lane * 8; c < d; c += 256) {
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int m = 0; m < members; ++m) {
                const int64_t j = sidx[p + m];
                const float iv = inv[j] * (lam2 ? 1.f - 0.5f * lam2[j] : 1.f);
                float v[8];
                load8(x + j * d + c, v);
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[k] += v[k] * iv;
            }
            store8(qr + c, acc);
        }
    } else {
        for (int64_t c = lane; c < d; c += 32) {
            float acc = 0.f;
            for (int m = 0; m < members; ++m) {
                const int64_t j = sidx[p + m];
                acc += load_as_float(x + j * d, c) * inv[j] * (lam2 ? 1.f - 0.5f * lam2[j] : 1.f);
            }
            qr[c] = acc;
        }
    }
}

template <typename J>
struct JobList {
    J j[MAX_JOBS];
};

// Batched form of the class sums: one warp per (CS_SPAN consecutive sorted positions, 256-column chunk) handles every
// class segment whose HEAD lies in its span (about one class of 8 at the benchmark's label distribution).  What made
// the first versions slow was latency, not bytes (35-42 us per 50 MB modality = 1.4 TB/s): a warp per position leaves
// seven of eight warps with nothing to do, and a serial scan of the span or a one-row-at-a-time member loop is a chain
// of dependent global loads.  Here the span's nine keys are fetched by nine lanes at once (heads by ballot), four
// members' 16-byte loads are in flight per lane, and the second weighted sum (the lam2-weighted class sums of the
// image rows for the pairs (image, dna) and (image, text) share one pass) is a template switch so that the common
// one-sum case keeps its registers -- and its occupancy.  Members are added in index order (fixed summation order).
constexpr int CS_SPAN = 8;
template <typename T, bool TWO>
__global__ void __launch_bounds__(kThreads)
class_sums_jobs_kernel(JobList<ClassSumJob> jobs, const int32_t* __restrict__ skey, const int32_t* __restrict__ sidx,
                       const float* __restrict__ cnt, int64_t N, int64_t d, int64_t row0, int64_t n, int nchunks) {
    const ClassSumJob& jb = jobs.j[blockIdx.y];
    const int64_t item = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t span = item / nchunks;
    const int chunk = static_cast<int>(item - span * nchunks);
    const int64_t p0 = span * CS_SPAN;
    if (p0 >= N) return;
    // lane l in [0, CS_SPAN] holds the key of sorted position p0 - 1 + l (-1 before the first position)
    const int64_t pidx = p0 - 1 + lane;
    int32_t key = -1;
    if (lane <= CS_SPAN && pidx >= 0 && pidx < N) key = skey[pidx];
    const int32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool is_head = lane >= 1 && lane <= CS_SPAN && pidx < N && key != prev;
    unsigned heads = __ballot_sync(0xffffffffu, is_head);
    const int64_t c = static_cast<int64_t>(chunk) * 256 + lane * 8;
    const T* x = static_cast<const T*>(jb.x);
    while (heads != 0u) {
        const int hl = __ffs(heads) - 1;
        heads &= heads - 1;
        const int64_t p = p0 - 1 + hl;
        const int32_t r = __shfl_sync(0xffffffffu, key, hl);
        const int members = static_cast<int>(cnt[r]);
        if (n < N) {  // row-sharded: only classes with a member among the local rows are ever read on this rank
            bool need = false;
            for (int m = lane; m < members; m += 32) {
                const int64_t j = sidx[p + m];
                need |= (j >= row0 && j < row0 + n);
            }
            if (!__any_sync(0xffffffffu, need)) continue;
        }
        if (c >= d) continue;
        float acc0[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float acc1[TWO ? 8 : 1] = {0.f};
        if (TWO) {
#pragma unroll
            for (int k = 0; k < (TWO ? 8 : 1); ++k) acc1[k] = 0.f;
        }
        for (int m0 = 0; m0 < members; m0 += 4) {
            float v[4][8], w0[4], w1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                w0[u] = w1[u] = 0.f;
                if (m0 + u < members) {
                    const int64_t j = sidx[p + m0 + u];
                    const float iv = jb.inv[j];
                    w0[u] = iv * (jb.lam2[0] ? 1.f - 0.5f * jb.lam2[0][j] : 1.f);
                    if (TWO) w1[u] = iv * (jb.lam2[1] ? 1.f - 0.5f * jb.lam2[1][j] : 1.f);
                    load8(x + j * d + c, v[u]);
                } else {
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[u][k] = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    acc0[k] += v[u][k] * w0[u];
                    if (TWO) acc1[k] += v[u][k] * w1[u];
                }
            }
        }
        store8(jb.out[0] + static_cast<int64_t>(r) * d + c, acc0);
        if (TWO) {
            float a1[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) a1[k] = acc1[TWO ? k : 0];
            store8(jb.out[1] + static_cast<int64_t>(r) * d + c, a1);
        }
    }
}

// MO_ROWS x 64 tile: xh[row][col] = cvt(x * inv) (zero for col >= d), xhT[col][row] (zero for row >= N); 16-byte
// global accesses on both outputs (and on the input when rows are 16-byte aligned).  With perm the tile walks the
// rows in permuted order: position k holds input row perm[k]; xh stays in input order, xhS / xhT follow perm.
// (64 rows per tile: a column of the transposed copy receives 128 contiguous bytes per tile.  The tile is chunk-swizzled
//  so that the transposing shared-memory reads are conflict-free: 41 -> 28 us per 32768 x 768 modality.  128-row tiles
//  (256-byte pieces, -DCLIBD_MO_ROWS=128) measured the same with the swizzle and slower without it.)
#ifndef CLIBD_MO_ROWS
#define CLIBD_MO_ROWS 64
#endif
constexpr int MO_ROWS = CLIBD_MO_ROWS;
template <typename T, bool VEC>
__global__ void make_operands_kernel(const T* __restrict__ x, const float* __restrict__ inv, int64_t N, int64_t d,
                                     int64_t dpad, int64_t npad, int fmt_bf16, uint16_t* __restrict__ xh,
                                     uint16_t* __restrict__ xhT, const int32_t* __restrict__ perm,
                                     uint16_t* __restrict__ xhS) {
    __shared__ __align__(16) uint16_t tile[MO_ROWS][72];
    const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 64, r0 = static_cast<int64_t>(blockIdx.y) * MO_ROWS;
    for (int idx = threadIdx.x; idx < MO_ROWS * 8; idx += blockDim.x) {
        const int rr = idx >> 3, ch = idx & 7;
        const int64_t row = r0 + rr, col = c0 + ch * 8;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int64_t src = row;
        if (row < N) {
            if (perm) src = perm[row];
            const float iv = inv ? inv[src] : 1.f;
            if (VEC && col + 8 <= d) {
                load8(x + src * d + col, v);
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] *= iv;
            } else {
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (col + k < d) v[k] = load_as_float(x + src * d, col + k) * iv;
            }
        }
        uint4 pk;
        uint32_t* w = reinterpret_cast<uint32_t*>(&pk);
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = pack2_operand16(v[2 * k], v[2 * k + 1], fmt_bf16);
        // 16-byte chunk ch of row rr sits at chunk position ch ^ (rr / 8 mod 8): the transposing reads below (8 row groups
        // per column) then hit 8 different banks instead of one
        *reinterpret_cast<uint4*>(&tile[rr][(ch ^ ((rr >> 3) & 7)) * 8]) = pk;
        if (row < N && col < dpad) {
            if (xh) *reinterpret_cast<uint4*>(xh + src * dpad + col) = pk;
            if (xhS) *reinterpret_cast<uint4*>(xhS + row * dpad + col) = pk;
        }
    }
    __syncthreads();
    if (xhT != nullptr) {
        for (int idx = threadIdx.x; idx < 64 * (MO_ROWS / 8); idx += blockDim.x) {
            const int cc = idx / (MO_ROWS / 8), rch = idx % (MO_ROWS / 8);
            const int64_t col = c0 + cc, row = r0 + rch * 8;
            if (col < dpad && row < npad) {
                uint4 pk;
                uint16_t* h = reinterpret_cast<uint16_t*>(&pk);
#pragma unroll
                for (int k = 0; k < 8; ++k) h[k] = tile[rch * 8 + k][((((cc >> 3) ^ (rch & 7)) << 3) | (cc & 7))];
                *reinterpret_cast<uint4*>(xhT + col * npad + row) = pk;
            }
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
                                    float* __restrict__ out, const int32_t* __restrict__ scatter) {
    int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= len) return;
    float acc = 0.f;
    for (int64_t p = 0; p < parts; ++p) acc += part[p * stride + k];
    out[scatter ? scatter[k] : k] = acc;
}

template <typename T, bool VEC>
__global__ void pos_rows_jobs_kernel(JobList<PosRowsJob> jobs, const int32_t* __restrict__ rep, int64_t d, int64_t row0,
                                     int64_t n) {
    const PosRowsJob& jb = jobs.j[blockIdx.y];
    const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const int64_t gi = row0 + i;
    const T* xr = static_cast<const T*>(jb.xa) + gi * d;
    const float* q = jb.Qb + static_cast<int64_t>(rep[gi]) * d;
    const float iv = jb.inv_a[gi];
    float acc = 0.f;
    if (VEC) {
        for (int64_t c = lane * 8; c < d; c += 256) {
            float xv[8], qv[8];
            load8(xr + c, xv);
            load8(q + c, qv);
#pragma unroll
            for (int k = 0; k < 8; ++k) acc = fmaf(xv[k] * iv, qv[k], acc);
        }
    } else {
        for (int64_t k = lane; k < d; k += 32) acc = fmaf(load_as_float(xr, k) * iv, q[k], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) jb.posrow[i] = acc;
}

// 32 outputs per block; warp g of the block adds the partials p = g, g + 8, g + 16, ... of those outputs (coalesced
// 128-byte loads, 8 loads in flight per output), then the 8 group sums are combined in a fixed order.
__global__ void reduce_parts_jobs_kernel(JobList<ReduceJob> jobs) {
    __shared__ float s_acc[8][32];
    const ReduceJob& jb = jobs.j[blockIdx.y];
    const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
    const int64_t k = static_cast<int64_t>(blockIdx.x) * 32 + lane;
    if (static_cast<int64_t>(blockIdx.x) * 32 >= jb.len) return;
    float acc = 0.f;
    if (k < jb.len)
        for (int64_t p = g; p < jb.parts; p += 8) acc += jb.part[p * jb.stride + k];
    s_acc[g][lane] = acc;
    __syncthreads();
    if (g == 0 && k < jb.len) {
        float t = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += s_acc[q][lane];
        jb.out[jb.scatter ? jb.scatter[k] : k] = t;
    }
}

// class_start[r] = first sorted position of the class whose representative is r
__global__ void class_start_kernel(const int32_t* __restrict__ skey, int64_t N, int32_t* __restrict__ cstart) {
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= N) return;
    const int32_t r = skey[k];
    if (k == 0 || skey[k - 1] != r) cstart[r] = static_cast<int32_t>(k);
}

__global__ void class_lo_kernel(const int32_t* __restrict__ rep, const int32_t* __restrict__ cstart, int64_t N,
                                int32_t* __restrict__ class_lo) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < N) class_lo[i] = cstart[rep[i]];
}

// ccS[k] = colcoef[sidx[k]]; lam2[i] = 2 min(1, 0.5 e_i (rowcoef_i + colcoef_i)), e_i = exp(s (posrow_i / cnt_i - 1)):
// the value G~ takes on a row's positives when they all sit at the row's mean positive similarity.  ANY lam2 in
// [0, 2] is algebraically exact (the epilogue subtracts lam2 on the positives, the fp32 class-sum term the
// remaining 2 - lam2); this choice makes the 16-bit rounded operand G~ - lam2 small exactly when G~ -> 2 T.
__global__ void sweep_prep_kernel(const float* __restrict__ rowcoef, const float* __restrict__ colcoef,
                                  const int32_t* __restrict__ sidx, const float* __restrict__ cnt,
                                  const float* __restrict__ posrow, int64_t N, int64_t row0, int64_t n, float scale,
                                  float* __restrict__ ccS, float* __restrict__ lam2,
                                  const float* __restrict__ scale_dev) {
    scale = eff_scale(scale, scale_dev);
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < N) ccS[k] = colcoef[sidx ? sidx[k] : k];
    if (lam2 != nullptr && k < n) {
        const int64_t gi = row0 + k;
        const float e = expf(scale * (posrow[k] / cnt[gi]) - softmax_shift(scale));
        lam2[k] = 2.f * fminf(1.f, 0.5f * e * (rowcoef[gi] + colcoef[gi]));
    }
}

__global__ void sweep_prep_jobs_kernel(JobList<SweepPrepJob> jobs, const int32_t* __restrict__ sidx,
                                       const float* __restrict__ cnt, int64_t N, int64_t row0, int64_t n, float scale,
                                       const float* __restrict__ scale_dev) {
    const SweepPrepJob& jb = jobs.j[blockIdx.y];
    scale = eff_scale(scale, scale_dev);
    const int64_t k = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k < N) jb.ccS[k] = jb.colcoef[sidx ? sidx[k] : k];
    if (jb.lam2 != nullptr && k < n) {
        const int64_t gi = row0 + k;
        const float e = expf(scale * (jb.posrow[k] / cnt[gi]) - softmax_shift(scale));
        jb.lam2[k] = 2.f * fminf(1.f, 0.5f * e * (jb.rowcoef[gi] + jb.colcoef[gi]));
    }
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

// sum of the kRedBlocks block partials by ONE warp in a fixed order (lane l adds partials l, l + 32, ...; then the
// shuffle tree): deterministic, and 4 microseconds instead of the 12 a single thread needs
__device__ __forceinline__ double warp_sum_partials(const double* __restrict__ red) {
    double t = 0.0;
    for (int b = threadIdx.x & 31; b < kRedBlocks; b += 32) t += red[b];
    return warp_sum(t);
}

__global__ void sum_stage2_kernel(const double* __restrict__ red, double mul, double* __restrict__ out,
                                  const float* __restrict__ div_dev) {
    const double t = warp_sum_partials(red);
    if (threadIdx.x == 0) out[0] = div_dev ? t * mul / static_cast<double>(div_dev[0]) : t * mul;
}

__global__ void sum_jobs_stage1_kernel(JobList<SumJob> jobs, double* __restrict__ red) {
    __shared__ double s_buf[kThreads / 32];
    const SumJob& jb = jobs.j[blockIdx.y];
    double acc = 0.0;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; k < jb.len;
         k += static_cast<int64_t>(kRedBlocks) * kThreads)
        acc += static_cast<double>(jb.in[k]);
    double t = block_sum_double(acc, s_buf);
    if (threadIdx.x == 0) red[blockIdx.y * kRedBlocks + blockIdx.x] = t;
}

__global__ void sum_jobs_stage2_kernel(JobList<SumJob> jobs, const double* __restrict__ red, double mul,
                                       const float* __restrict__ div_dev) {
    const int job = threadIdx.x >> 5;  // one warp per job
    const double t = warp_sum_partials(red + job * kRedBlocks);
    if ((threadIdx.x & 31) == 0)
        jobs.j[job].out[0] = div_dev ? t * mul / static_cast<double>(div_dev[0]) : t * mul;
}

// term_p(k) = cnt[k] * (2*s + ln rowsum_p[k] + ln colsum_p[k]); also u = cnt/rowsum, v = cnt/colsum
__global__ void loss_finish_stage1_kernel(int64_t N, float scale, float w0, float w1, float w2,
                                          const float* __restrict__ cnt, const float* __restrict__ rowsum,
                                          const float* __restrict__ colsum, float* __restrict__ u,
                                          float* __restrict__ v, double* __restrict__ red,
                                          const float* __restrict__ scale_dev) {
    scale = eff_scale(scale, scale_dev);
    __shared__ double s_buf[kThreads / 32];
    const float w[3] = {w0, w1, w2};
    double acc = 0.0;
    for (int64_t k = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; k < N;
         k += static_cast<int64_t>(kRedBlocks) * kThreads) {
        const float c = cnt[k];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            if (w[p] == 0.f) continue;
            // a sum that underflowed entirely (see softmax_shift) is clamped: finite loss and coefficients, no log(0)
            const float rs = fmaxf(rowsum[p * N + k], kMinSoftmaxSum), cs = fmaxf(colsum[p * N + k], kMinSoftmaxSum);
            u[p * N + k] = c / rs;
            v[p * N + k] = c / cs;
            acc += static_cast<double>(w[p]) * static_cast<double>(c) *
                   (2.0 * static_cast<double>(softmax_shift(scale)) + log(static_cast<double>(rs)) +
                    log(static_cast<double>(cs)));
        }
    }
    double t = block_sum_double(acc, s_buf);
    if (threadIdx.x == 0) red[blockIdx.x] = t;
}

__global__ void loss_finish_stage2_kernel(int64_t N, float scale, float w0, float w1, float w2,
                                          const double* __restrict__ red, const double* __restrict__ pos,
                                          float* __restrict__ loss_out, const float* __restrict__ scale_dev) {
    scale = eff_scale(scale, scale_dev);
    double t = warp_sum_partials(red);
    if (threadIdx.x == 0) {
        const float w[3] = {w0, w1, w2};
        for (int p = 0; p < 3; ++p)
            if (w[p] != 0.f) t -= 2.0 * static_cast<double>(w[p]) * static_cast<double>(scale) * pos[p];
        loss_out[0] = static_cast<float>(t / static_cast<double>(N));
    }
}

// One warp per local row: dxhat = (scale/N) * (sum_splits dxh - 2 * sum_partners w_p Q_partner[rep]),
// dots = xhat . dxhat, dx = grad_scale * (dxhat - xhat * dots) / ||x||.  Rows up to 1024 columns stay in
// registers between the dot product and the projection (one pass over HBM).
struct NormBwdList {
    NormBwdArgs a[3];
};

template <typename T, bool VEC>
__global__ void normalize_bwd_kernel(NormBwdList list) {
    const NormBwdArgs& a = list.a[blockIdx.y];  // blockIdx.y = modality: all of them in one launch
    const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (i >= a.n) return;
    const int64_t gi = a.row0 + i;
    const T* xr = reinterpret_cast<const T*>(a.x) + gi * a.d;
    const float iv = a.inv_norm[gi];
    const int64_t rp = a.rep[gi];
    const float k1 = eff_scale(a.scale, a.scale_dev) / static_cast<float>(a.N);
    float gdev = 1.f;
    if (a.grad_scale_dev) {  // one upstream gradient, or the sum of one per rank (fixed order: same value on every rank)
        gdev = 0.f;
        for (int r = 0; r < a.grad_scale_dev_count; ++r) gdev += a.grad_scale_dev[r];
    }
    const float gsc = a.grad_scale * gdev;
    T* dxr = a.dx ? reinterpret_cast<T*>(a.dx) + i * a.d : nullptr;
    // effective partner weights: the sweep already subtracted lam2 of the 2 on this row's positives
    const float wpe[2] = {a.wp[0] * (a.lam2[0] ? 1.f - 0.5f * a.lam2[0][i] : 1.f),
                          a.wp[1] * (a.lam2[1] ? 1.f - 0.5f * a.lam2[1][i] : 1.f)};
    if constexpr (VEC) {
        auto dxhat8 = [&](int64_t c, float (&g)[8]) {
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] = 0.f;
            for (int s = 0; s < a.jsplit; ++s) {
                float t[8];
                load8(a.dxh + (static_cast<int64_t>(s) * a.n + i) * a.d + c, t);
#pragma unroll
                for (int k = 0; k < 8; ++k) g[k] += t[k];
            }
            // the received slots (one per source rank): four loads in flight, added in slot order
            for (int s0 = 0; s0 < a.extra_slots; s0 += 4) {
                float t[4][8];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (s0 + u < a.extra_slots) {
                        load8(a.extra + (static_cast<int64_t>(s0 + u) * a.n + i) * a.d + c, t[u]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k) t[u][k] = 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int k = 0; k < 8; ++k) g[k] += t[u][k];
            }
            float q[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int pz = 0; pz < 2; ++pz) {
                if (a.Qp[pz]) {
                    float t[8];
                    load8(a.Qp[pz] + rp * a.d + c, t);
#pragma unroll
                    for (int k = 0; k < 8; ++k) q[k] = fmaf(wpe[pz], t[k], q[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] = k1 * (g[k] - 2.f * q[k]);
        };
        if (a.d <= 1024) {
            float xv[4][8], gv[4][8];
            float part = 0.f;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int64_t c = lane * 8 + it * 256;
                if (c < a.d) {
                    load8(xr + c, xv[it]);
                    dxhat8(c, gv[it]);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        xv[it][k] *= iv;
                        part = fmaf(xv[it][k], gv[it][k], part);
                    }
                }
            }
            const float dot = warp_sum(part);
            if (lane == 0) a.dots[i] = dot;
            if (dxr == nullptr) return;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int64_t c = lane * 8 + it * 256;
                if (c < a.d) {
                    float o[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) o[k] = gsc * (gv[it][k] - xv[it][k] * dot) * iv;
                    store8(dxr + c, o);
                }
            }
        } else {
            float part = 0.f;
            for (int64_t c = lane * 8; c < a.d; c += 256) {
                float xv[8], gv[8];
                load8(xr + c, xv);
                dxhat8(c, gv);
#pragma unroll
                for (int k = 0; k < 8; ++k) part = fmaf(xv[k] * iv, gv[k], part);
            }
            const float dot = warp_sum(part);
            if (lane == 0) a.dots[i] = dot;
            if (dxr == nullptr) return;
            for (int64_t c = lane * 8; c < a.d; c += 256) {
                float xv[8], gv[8], o[8];
                load8(xr + c, xv);
                dxhat8(c, gv);
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = gsc * (gv[k] - xv[k] * iv * dot) * iv;
                store8(dxr + c, o);
            }
        }
    } else {
        auto dxhat = [&](int64_t c) -> float {
            float g = 0.f;
            for (int s = 0; s < a.jsplit; ++s) g += a.dxh[(static_cast<int64_t>(s) * a.n + i) * a.d + c];
            for (int s = 0; s < a.extra_slots; ++s) g += a.extra[(static_cast<int64_t>(s) * a.n + i) * a.d + c];
            float t = 0.f;
            if (a.Qp[0]) t = fmaf(wpe[0], a.Qp[0][rp * a.d + c], t);
            if (a.Qp[1]) t = fmaf(wpe[1], a.Qp[1][rp * a.d + c], t);
            return k1 * (g - 2.f * t);
        };
        float part = 0.f;
        for (int64_t c = lane; c < a.d; c += 32) part = fmaf(load_as_float(xr, c) * iv, dxhat(c), part);
        const float dot = warp_sum(part);
        if (lane == 0) a.dots[i] = dot;
        if (dxr == nullptr) return;
        for (int64_t c = lane; c < a.d; c += 32) {
            const float xh = load_as_float(xr, c) * iv;
            store_from_float(dxr, c, gsc * (dxhat(c) - xh * dot) * iv);
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

int launch_row_inv_norm(const void* x, int dtype, int64_t n, int64_t d, float* inv_norm, cudaStream_t s) {
    if (n == 0) return 0;
    const int64_t blocks = ceil_div(n * 32, kThreads);
    DISPATCH_DTYPE(dtype, (row_inv_norm_kernel<T><<<blocks, kThreads, 0, s>>>(static_cast<const T*>(x), n, d, inv_norm)));
    CLIBD_KERNEL_CHECK();
    return 0;
}

int64_t label_hash_slots(int64_t N) {
    int64_t H = 64;
    while (H < 2 * N) H <<= 1;
    return H;
}

size_t class_sort_temp_bytes(int64_t N) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr),
                                    static_cast<const int32_t*>(nullptr), static_cast<int32_t*>(nullptr),
                                    static_cast<int>(N));
    return bytes;
}

int launch_label_stats(const int64_t* labels, int64_t N, int32_t* rep, float* cnt, const LabelScratch& ls,
                       cudaStream_t s) {
    if (N == 0) return 0;
    const int64_t H = label_hash_slots(N);
    hash_init_kernel<<<ceil_div(H, kThreads), kThreads, 0, s>>>(ls.own, ls.hmin, ls.hcnt, H);
    CLIBD_KERNEL_CHECK();
    hash_insert_kernel<<<ceil_div(N, kThreads), kThreads, 0, s>>>(labels, N, H, ls.own, ls.hmin, ls.hcnt, rep);
    CLIBD_KERNEL_CHECK();
    hash_lookup_kernel<<<ceil_div(N, kThreads), kThreads, 0, s>>>(N, ls.hmin, ls.hcnt, rep, cnt, ls.iota);
    CLIBD_KERNEL_CHECK();
    // rows grouped by class, in index order inside a class (stable LSD radix sort on the representative)
    int end_bit = 1;
    while ((int64_t(1) << end_bit) < N) ++end_bit;
    size_t bytes = ls.sort_tmp_bytes;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(ls.sort_tmp, bytes, rep, ls.skey, ls.iota, ls.sidx, static_cast<int>(N),
                                                    0, end_bit, s);
    if (e != cudaSuccess) {
        set_error(std::string("class sort failed: ") + cudaGetErrorString(e));
        return 2;
    }
    count_launch();
    return 0;
}

int launch_gscale(const float* cnt, int64_t N, int path, float* gscale, cudaStream_t s) {
    gscale_kernel<<<1, kThreads, 0, s>>>(cnt, N, path, gscale);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_class_sums(const void* x, int dtype, const float* inv_norm, const int32_t* skey, const int32_t* sidx,
                      const float* cnt, int64_t N, int64_t d, int64_t row0, int64_t n, float* Q, cudaStream_t s,
                      const float* lam2) {
    if (N == 0) return 0;
    const int64_t blocks = ceil_div(N * 32, kThreads);
    const bool vec = rows_vec8_ok<void>(x, d) && rows_vec8_ok<void>(Q, d);
    DISPATCH_DTYPE(dtype, {
        if (vec && (sizeof(T) == 2 || d % 8 == 0))
            class_sums_kernel<T, true><<<blocks, kThreads, 0, s>>>(static_cast<const T*>(x), inv_norm, skey, sidx, cnt, N, d, row0, n, Q, lam2);
        else
            class_sums_kernel<T, false><<<blocks, kThreads, 0, s>>>(static_cast<const T*>(x), inv_norm, skey, sidx, cnt, N, d, row0, n, Q, lam2);
    });
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_make_operands(const void* x, int dtype, const float* inv_norm, int64_t N, int64_t d, int64_t dpad,
                         int64_t npad, int fmt_bf16, void* xh, void* xhT, cudaStream_t s, const int32_t* perm,
                         void* xhS) {
    const int64_t rows = npad > N ? npad : N;
    dim3 grid(static_cast<unsigned>(ceil_div(dpad, 64)), static_cast<unsigned>(ceil_div(rows, MO_ROWS)));
    const bool vec = rows_vec8_ok<void>(x, d);
    DISPATCH_DTYPE(dtype, {
        if (vec)
            make_operands_kernel<T, true><<<grid, kThreads, 0, s>>>(static_cast<const T*>(x), inv_norm, N, d, dpad, npad,
                                                                    fmt_bf16, static_cast<uint16_t*>(xh),
                                                                    static_cast<uint16_t*>(xhT), perm,
                                                                    static_cast<uint16_t*>(xhS));
        else
            make_operands_kernel<T, false><<<grid, kThreads, 0, s>>>(static_cast<const T*>(x), inv_norm, N, d, dpad, npad,
                                                                     fmt_bf16, static_cast<uint16_t*>(xh),
                                                                     static_cast<uint16_t*>(xhT), perm,
                                                                     static_cast<uint16_t*>(xhS));
    });
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

int launch_reduce_parts(const float* part, int64_t parts, int64_t stride, int64_t len, float* out, cudaStream_t s,
                        const int32_t* scatter) {
    if (len == 0) return 0;
    reduce_parts_kernel<<<ceil_div(len, kThreads), kThreads, 0, s>>>(part, parts, stride, len, out, scatter);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_class_ranges(const int32_t* skey, const int32_t* rep, int64_t N, int32_t* cstart, int32_t* class_lo,
                        cudaStream_t s) {
    if (N == 0) return 0;
    class_start_kernel<<<ceil_div(N, kThreads), kThreads, 0, s>>>(skey, N, cstart);
    CLIBD_KERNEL_CHECK();
    class_lo_kernel<<<ceil_div(N, kThreads), kThreads, 0, s>>>(rep, cstart, N, class_lo);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_sweep_prep(const float* rowcoef, const float* colcoef, const int32_t* sidx, const float* cnt,
                      const float* posrow, int64_t N, int64_t row0, int64_t n, float scale, float* ccS, float* lam2,
                      cudaStream_t s) {
    if (N == 0) return 0;
    sweep_prep_kernel<<<ceil_div(N, kThreads), kThreads, 0, s>>>(rowcoef, colcoef, sidx, cnt, posrow, N, row0, n, scale,
                                                               ccS, lam2, scale_dev_ptr());
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_sum_to_double(const float* in, int64_t len, double mul, double* red, double* out, cudaStream_t s,
                         const float* div_dev) {
    sum_stage1_kernel<<<kRedBlocks, kThreads, 0, s>>>(in, len, red);
    sum_stage2_kernel<<<1, 32, 0, s>>>(red, mul, out, div_dev);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_loss_finish(int64_t N, float scale, const float w[3], const float* cnt, const float* rowsum,
                       const float* colsum, const double* pos, float* u, float* v, double* red, float* loss_out,
                       cudaStream_t s) {
    loss_finish_stage1_kernel<<<kRedBlocks, kThreads, 0, s>>>(N, scale, w[0], w[1], w[2], cnt, rowsum, colsum, u, v, red, scale_dev_ptr());
    loss_finish_stage2_kernel<<<1, 32, 0, s>>>(N, scale, w[0], w[1], w[2], red, pos, loss_out, scale_dev_ptr());
    CLIBD_KERNEL_CHECK();
    return 0;
}

template <typename J>
static JobList<J> job_list(const J* jobs, int njobs) {
    JobList<J> l;
    for (int i = 0; i < njobs; ++i) l.j[i] = jobs[i];
    return l;
}

int launch_class_sums_jobs(const ClassSumJob* jobs, int njobs, int dtype, const int32_t* skey, const int32_t* sidx,
                           const float* cnt, int64_t N, int64_t d, int64_t row0, int64_t n, cudaStream_t s) {
    if (N == 0 || njobs == 0) return 0;
    CLIBD_REQUIRE(njobs <= MAX_JOBS, "too many class-sum jobs");
    bool vec = true;
    for (int i = 0; i < njobs; ++i) {
        vec = vec && rows_vec8_ok<void>(jobs[i].x, d) && rows_vec8_ok<void>(jobs[i].out[0], d) &&
              (jobs[i].out[1] == nullptr || rows_vec8_ok<void>(jobs[i].out[1], d));
    }
    if (!vec) {  // general layout: the one-output kernel, job by job
        for (int i = 0; i < njobs; ++i)
            for (int k = 0; k < 2; ++k)
                if (jobs[i].out[k]) {
                    int rc = launch_class_sums(jobs[i].x, dtype, jobs[i].inv, skey, sidx, cnt, N, d, row0, n, jobs[i].out[k], s,
                                               jobs[i].lam2[k]);
                    if (rc) return rc;
                }
        return 0;
    }
    const int nchunks = static_cast<int>(ceil_div(d, 256));
    dim3 grid(static_cast<unsigned>(ceil_div(ceil_div(N, CS_SPAN) * nchunks * 32, kThreads)), static_cast<unsigned>(njobs));
    // a launch forms two sums per job only when some job asks for it; the others then get a second, unused target
    bool two = false;
    for (int i = 0; i < njobs; ++i) two = two || jobs[i].out[1] != nullptr;
    if (two) {  // jobs with one target run in their own one-sum launch (fewer registers, higher occupancy)
        ClassSumJob one[MAX_JOBS], both[MAX_JOBS];
        int n1 = 0, n2 = 0;
        for (int i = 0; i < njobs; ++i) (jobs[i].out[1] ? both[n2++] : one[n1++]) = jobs[i];
        if (n1 > 0) {
            int rc = launch_class_sums_jobs(one, n1, dtype, skey, sidx, cnt, N, d, row0, n, s);
            if (rc) return rc;
        }
        dim3 grid2(grid.x, static_cast<unsigned>(n2));
        const JobList<ClassSumJob> l2 = job_list(both, n2);
        DISPATCH_DTYPE(dtype, (class_sums_jobs_kernel<T, true><<<grid2, kThreads, 0, s>>>(l2, skey, sidx, cnt, N, d, row0, n, nchunks)));
        CLIBD_KERNEL_CHECK();
        return 0;
    }
    const JobList<ClassSumJob> l = job_list(jobs, njobs);
    DISPATCH_DTYPE(dtype, (class_sums_jobs_kernel<T, false><<<grid, kThreads, 0, s>>>(l, skey, sidx, cnt, N, d, row0, n, nchunks)));
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_reduce_parts_jobs(const ReduceJob* jobs, int njobs, cudaStream_t s) {
    if (njobs == 0) return 0;
    CLIBD_REQUIRE(njobs <= MAX_JOBS, "too many reduction jobs");
    int64_t len = 0;
    for (int i = 0; i < njobs; ++i) len = jobs[i].len > len ? jobs[i].len : len;
    if (len == 0) return 0;
    dim3 grid(static_cast<unsigned>(ceil_div(len, 32)), static_cast<unsigned>(njobs));
    reduce_parts_jobs_kernel<<<grid, kThreads, 0, s>>>(job_list(jobs, njobs));
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_pos_rows_jobs(const PosRowsJob* jobs, int njobs, int dtype, const int32_t* rep, int64_t d, int64_t row0,
                         int64_t n, cudaStream_t s) {
    if (n == 0 || njobs == 0) return 0;
    CLIBD_REQUIRE(njobs <= MAX_JOBS, "too many positive-row jobs");
    dim3 grid(static_cast<unsigned>(ceil_div(n * 32, kThreads)), static_cast<unsigned>(njobs));
    const JobList<PosRowsJob> l = job_list(jobs, njobs);
    bool vec = true;
    for (int i = 0; i < njobs; ++i) vec = vec && rows_vec8_ok<void>(jobs[i].xa, d) && rows_vec8_ok<void>(jobs[i].Qb, d);
    DISPATCH_DTYPE(dtype, {
        if (vec) pos_rows_jobs_kernel<T, true><<<grid, kThreads, 0, s>>>(l, rep, d, row0, n);
        else pos_rows_jobs_kernel<T, false><<<grid, kThreads, 0, s>>>(l, rep, d, row0, n);
    });
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_sum_to_double_jobs(const SumJob* jobs, int njobs, double mul, double* red, cudaStream_t s,
                              const float* div_dev) {
    if (njobs == 0) return 0;
    CLIBD_REQUIRE(njobs <= MAX_JOBS, "too many sum jobs");
    const JobList<SumJob> l = job_list(jobs, njobs);
    sum_jobs_stage1_kernel<<<dim3(kRedBlocks, static_cast<unsigned>(njobs)), kThreads, 0, s>>>(l, red);
    sum_jobs_stage2_kernel<<<1, 32 * njobs, 0, s>>>(l, red, mul, div_dev);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_sweep_prep_jobs(const SweepPrepJob* jobs, int njobs, const int32_t* sidx, const float* cnt, int64_t N,
                           int64_t row0, int64_t n, float scale, cudaStream_t s) {
    if (N == 0 || njobs == 0) return 0;
    CLIBD_REQUIRE(njobs <= MAX_JOBS, "too many sweep-prep jobs");
    dim3 grid(static_cast<unsigned>(ceil_div(N, kThreads)), static_cast<unsigned>(njobs));
    sweep_prep_jobs_kernel<<<grid, kThreads, 0, s>>>(job_list(jobs, njobs), sidx, cnt, N, row0, n, scale, scale_dev_ptr());
    CLIBD_KERNEL_CHECK();
    return 0;
}

int launch_normalize_bwd(const NormBwdArgs& a_in, cudaStream_t s) { return launch_normalize_bwd_multi(&a_in, 1, s); }

int launch_normalize_bwd_multi(const NormBwdArgs* args, int count, cudaStream_t s) {
    if (count == 0 || args[0].n == 0) return 0;
    CLIBD_REQUIRE(count <= 3, "at most three modalities");
    NormBwdList list;
    bool vec = true;
    for (int m = 0; m < count; ++m) {
        NormBwdArgs& a = list.a[m];
        a = args[m];
        CLIBD_REQUIRE(a.n == args[0].n && a.dtype == args[0].dtype, "modalities of one launch share shape and dtype");
        a.scale_dev = scale_dev_ptr();
        vec = vec && rows_vec8_ok<void>(a.x, a.d) && (a.jsplit == 0 || rows_vec8_ok<void>(a.dxh, a.d)) &&
              (a.extra_slots == 0 || rows_vec8_ok<void>(a.extra, a.d)) && (a.dx == nullptr || rows_vec8_ok<void>(a.dx, a.d));
        for (int p = 0; p < 2; ++p)
            if (a.Qp[p]) vec = vec && rows_vec8_ok<void>(a.Qp[p], a.d);
    }
    dim3 grid(static_cast<unsigned>(ceil_div(args[0].n * 32, kThreads)), static_cast<unsigned>(count));
    DISPATCH_DTYPE(args[0].dtype, {
        if (vec)
            normalize_bwd_kernel<T, true><<<grid, kThreads, 0, s>>>(list);
        else
            normalize_bwd_kernel<T, false><<<grid, kThreads, 0, s>>>(list);
    });
    CLIBD_KERNEL_CHECK();
    return 0;
}

}  // namespace clibd
