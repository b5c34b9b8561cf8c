// Cosine nearest-neighbour retrieval (replaces faiss IndexFlatIP add+search as used by
// bioscanclip/util/util.py:521-528, 759-766) and top-k accuracy counts (util.py:379-395, 555-599).
//
// Exactness contract (oracle/knn_oracle.py): similarity = float64 sum over d = 0..D-1, in that
// order, of the exact products of the float32 elements; results ordered by (-sim, index).
//
//   tcgen05 path : 16-bit operand copies -> 128x256 similarity tiles in TMEM -> the epilogue keeps,
//                  per query row and per key sub-range, the KP best screened scores (registers,
//                  lowest index first on ties) -> exact float64 re-rank of the candidates ->
//                  a query is accepted only if every sub-range's KP-th screened score plus a
//                  rigorous rounding bound is below its k-th exact score; all other queries are
//                  redone exhaustively in float64.
//   exact path   : 64x64 float64 tiles on CUDA cores + per-row selection.
#include <algorithm>
#include <cfloat>
#include <climits>

#include "../../include/clibd_b200.h"
#include "common.cuh"
#include "loss_plan.h"
#include "ptx.cuh"
#include "tmap.h"

namespace clibd {
namespace {

using namespace ptx;

constexpr int kDT_F64 = 3;

// ------------------------------------------------------------------------------------------
// normalise (float64 arithmetic) -> float32
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void knn_normalize_kernel(const T* __restrict__ x, int64_t n, int64_t d, float* __restrict__ out) {
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const T* xr = x + row * d;
    double ss = 0.0;
    for (int64_t k = lane; k < d; k += 32) {
        const double v = static_cast<double>(xr[k]);
        ss += v * v;
    }
    ss = warp_sum(ss);
    double nrm = sqrt(ss);
    if (nrm == 0.0) nrm = 1.0;  // sklearn.preprocessing.normalize leaves zero rows untouched
    for (int64_t k = lane; k < d; k += 32) out[row * d + k] = static_cast<float>(static_cast<double>(xr[k]) / nrm);
}

// ------------------------------------------------------------------------------------------
// tcgen05 screening kernel
// ------------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, cta_group::2, M = 256): a pair screens 256 queries x 256 keys per tile; each CTA holds
// its 128 query rows in TMEM (2 x 256 columns) and streams 128 query rows + 128 key rows per K block, so that the
// shared-memory traffic (64 B/cycle operand reads + 64 B/cycle TMA writes) fits the 128 B/cycle the SM has --
// the single-CTA 128 x 256 version needed 192 B/cycle and ran at ~70 % of the measured GEMM rate.
constexpr int S_BM = 128, S_BN = 256, S_BK = 64, S_STAGES = 6;
constexpr int S_QB = 2 * S_BM;  // queries per pair
constexpr int S_A_BYTES = S_BM * S_BK * 2, S_B_BYTES = (S_BN / 2) * S_BK * 2, S_STAGE_BYTES = S_A_BYTES + S_B_BYTES;
constexpr int S_THREADS = 384, S_EPI_WARPS = 8;
constexpr int S_SMEM_BARS = S_STAGES * S_STAGE_BYTES;
constexpr int S_NUM_BARS = 2 * S_STAGES + 4;
constexpr int S_SMEM_TMEMPTR = S_SMEM_BARS + S_NUM_BARS * 8;
constexpr int S_SMEM_ALLOC = S_SMEM_TMEMPTR + 16;

__global__ void knn_fill_kernel(float* __restrict__ p, int64_t n, float v) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// atomic max on a float that may be negative (never NaN here): signed-int order for >= 0, reversed unsigned order below
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f)
        atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

template <int KP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(S_THREADS, 1)
knn_screen_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k, int64_t Q,
                     int64_t K, int num_kb, int num_chunks, int64_t tiles_per_chunk, uint32_t idesc, int k,
                     float floor_delta, float* rowfloor, float* __restrict__ cand_score,
                     int32_t* __restrict__ cand_idx) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S_SMEM_BARS);
    uint64_t* full_bar = bars;                      // leader: both CTAs' operands landed
    uint64_t* empty_bar = bars + S_STAGES;          // every CTA
    uint64_t* tfull_bar = bars + 2 * S_STAGES;      // every CTA [2]
    uint64_t* tempty_bar = bars + 2 * S_STAGES + 2; // leader [2]: 16 epilogue warps of the pair
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + S_SMEM_TMEMPTR);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int64_t pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int64_t num_qb = (Q + S_QB - 1) / S_QB;
    const int64_t num_kt = (K + S_BN - 1) / S_BN;
    const int64_t num_units = num_qb * num_chunks;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_q);
        prefetch_tmap(&tm_k);
        for (int i = 0; i < S_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 2 * S_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_cg2(tmem_ptr, 512);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {  // TMA producer (both CTAs)
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t full_l0 = mapa_u32(smem_u32(&full_bar[0]), 0);
        const bool elected = elect_one();
        for (int64_t u = pair; u < num_units; u += num_pairs) {
            const int64_t qb = u % num_qb, kc = u / num_qb;
            const int64_t kt0 = kc * tiles_per_chunk;
            const int64_t kt1 = (kt0 + tiles_per_chunk < num_kt) ? kt0 + tiles_per_chunk : num_kt;
            const int32_t qrow = static_cast<int32_t>(qb * S_QB + rank * S_BM);
            for (int64_t kt = kt0; kt < kt1; ++kt) {
                const int32_t krow = static_cast<int32_t>(kt * S_BN + rank * (S_BN / 2));
#pragma unroll 1
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elected) {
                        uint8_t* sa = smem + stage * S_STAGE_BYTES;
                        if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * S_STAGE_BYTES);
                        tma_load_2d_cg2(&tm_q, full_l0 + stage * 8, sa, kb * S_BK, qrow, kEvictNormal);
                        tma_load_2d_cg2(&tm_k, full_l0 + stage * 8, sa + S_A_BYTES, kb * S_BK, krow, kEvictNormal);
                    }
                    __syncwarp();
                    if (++stage == S_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1 && leader) {  // MMA issuer (leader CTA)
        int stage = 0;
        uint32_t phase = 0, it = 0;
        const uint64_t d0 = make_sw128_kmajor_desc(smem_u32(smem));
        const bool elected = elect_one();
        for (int64_t u = pair; u < num_units; u += num_pairs) {
            const int64_t kc = u / num_qb;
            const int64_t kt0 = kc * tiles_per_chunk;
            const int64_t kt1 = (kt0 + tiles_per_chunk < num_kt) ? kt0 + tiles_per_chunk : num_kt;
            for (int64_t kt = kt0; kt < kt1; ++kt, ++it) {
                const uint32_t as = it & 1, aph = (it >> 1) & 1;
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * S_BN;
#pragma unroll 1
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elected) {
                        const uint64_t da = d0 + stage * (S_STAGE_BYTES >> 4);
                        const uint64_t db = da + (S_A_BYTES >> 4);
                        umma_f16_cg2(d_tmem, da, db, idesc, kb > 0 ? 1u : 0u);
                        umma_f16_cg2(d_tmem, da + 2, db + 2, idesc, 1u);
                        umma_f16_cg2(d_tmem, da + 4, db + 4, idesc, 1u);
                        umma_f16_cg2(d_tmem, da + 6, db + 6, idesc, 1u);
                        umma_commit_cg2(&empty_bar[stage], 3);
                        if (kb == num_kb - 1) umma_commit_cg2(&tfull_bar[as], 3);
                    }
                    __syncwarp();
                    if (++stage == S_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4, q = warp & 3, h = ew >> 2;
        const uint32_t tempty_l0 = mapa_u32(smem_u32(&tempty_bar[0]), 0);
        uint32_t it = 0;
        for (int64_t u = pair; u < num_units; u += num_pairs) {
            const int64_t qb = u % num_qb, kc = u / num_qb;
            const int64_t kt0 = kc * tiles_per_chunk;
            const int64_t kt1 = (kt0 + tiles_per_chunk < num_kt) ? kt0 + tiles_per_chunk : num_kt;
            // Row floor: some earlier unit of this row found k keys with screened score >= T; their exact scores are
            // >= T - eps, so a key screened below T - 2 eps (exact < T - eps) can neither enter nor tie the exact
            // top-k.  Lists start filled with sentinels (index -1) at that floor: after the first chunk of a row
            // almost no column passes the one-compare-per-32-columns test below.
            const int64_t row = qb * S_QB + rank * S_BM + q * 32 + lane;
            const float fl = (row < Q) ? *(volatile float*)(rowfloor + row) : -INFINITY;
            float ls[KP];
            int32_t li[KP];
#pragma unroll
            for (int i = 0; i < KP; ++i) {
                ls[i] = fl;
                li[i] = -1;
            }
            for (int64_t kt = kt0; kt < kt1; ++kt, ++it) {
                const uint32_t as = it & 1, aph = (it >> 1) & 1;
                const int64_t colbase = kt * S_BN + h * 128;
                const bool full_tile = kt * S_BN + S_BN <= K;
                mbar_wait(&tfull_bar[as], aph);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * S_BN + h * 128 + c * 32, v);
                    tmem_ld_wait();
                    if (c == 3) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(tempty_l0 + as * 8);
                    }
                    const int32_t cb = static_cast<int32_t>(colbase) + c * 32;
                    float f[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) f[k] = __uint_as_float(v[k]);
                    if (!full_tile) {
                        const int32_t kvalid = static_cast<int32_t>(K) - cb;  // columns >= kvalid do not exist
#pragma unroll
                        for (int k = 0; k < 32; ++k)
                            if (k >= kvalid) f[k] = -INFINITY;
                    }
                    // one branch per 32 columns: most chunks hold nothing above the list's last entry
                    float m01[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) m01[k] = fmaxf(f[2 * k], f[2 * k + 1]);
#pragma unroll
                    for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
                        for (int k = 0; k < w; ++k) m01[k] = fmaxf(m01[k], m01[k + w]);
                    if (m01[0] > ls[KP - 1]) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            float s = f[k];
                            if (s > ls[KP - 1]) {  // strict: an equal score with a higher index never displaces
                                int32_t id = cb + k;
#pragma unroll
                                for (int i = 0; i < KP; ++i) {
                                    if (s > ls[i]) {
                                        const float ts = ls[i];
                                        const int32_t ti = li[i];
                                        ls[i] = s;
                                        li[i] = id;
                                        s = ts;
                                        id = ti;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            if (row < Q) {
                const int64_t base = (row * (2 * num_chunks) + (kc * 2 + h)) * KP;
                float kth = -INFINITY;
                bool have_k = false;
#pragma unroll
                for (int i = 0; i < KP; ++i) {
                    cand_score[base + i] = ls[i];
                    cand_idx[base + i] = li[i];
                    if (i == k - 1) {
                        kth = ls[i];
                        have_k = li[i] >= 0;
                    }
                }
                if (have_k) atomic_max_float(rowfloor + row, kth - floor_delta);
            }
        }
    }
    tc_fence_before();
    cluster_sync_all();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// exact re-rank of the screened candidates; one block (64 threads) per query
// ------------------------------------------------------------------------------------------
constexpr int R_THREADS = 64;

__global__ void __launch_bounds__(R_THREADS)
knn_rerank_kernel(const float* __restrict__ q32, const float* __restrict__ k32, int64_t Q, int64_t K, int64_t d,
                  int k, int kp, int num_lists, float eps, int64_t key_offset, const float* __restrict__ cand_score,
                  const int32_t* __restrict__ cand_idx, double* __restrict__ out_sims, int64_t* __restrict__ out_idx,
                  int32_t* __restrict__ flagged, int32_t* __restrict__ n_flagged) {
    extern __shared__ uint8_t rr_smem[];
    const int C = num_lists * kp;
    double* sims = reinterpret_cast<double*>(rr_smem);                 // [C]
    int32_t* ids = reinterpret_cast<int32_t*>(sims + C);               // [C] survivors, compacted
    float* scr = reinterpret_cast<float*>(ids + C);                    // [C] screened scores (selection scratch)
    float* qs = scr + C;                                               // [d]
    __shared__ double s_best[2];
    __shared__ int s_besti[2];
    __shared__ int s_bestc[2];
    __shared__ float s_thr;
    __shared__ int s_ns;
    const int64_t qi = blockIdx.x;
    const int tid = threadIdx.x;
    const float* qrow = q32 + qi * d;
    __shared__ int s_nc;
    if (tid == 0) {
        s_ns = 0;
        s_nc = 0;
    }
    for (int64_t t = tid; t < d; t += R_THREADS) qs[t] = qrow[t];
    __syncthreads();
    // compact the real candidates (most lists hold only floor sentinels, index -1); cidx aliases the tail of sims[]
    // until the float64 evaluation starts (sims is only written after the survivors have been chosen)
    int32_t* cidx = reinterpret_cast<int32_t*>(sims);
    for (int c = tid; c < C; c += R_THREADS) {
        const int32_t id = cand_idx[qi * C + c];
        if (id >= 0) {
            const int pos = atomicAdd(&s_nc, 1);
            scr[pos] = cand_score[qi * C + c];
            cidx[pos] = id;
        }
    }
    __syncthreads();
    const int nc = s_nc;
    // Prefilter: with s_k the k-th largest SCREENED score, at least k candidates have an exact score >= s_k - eps,
    // so a candidate whose screened score is below s_k - 2 eps (exact < s_k - eps) cannot be in the exact top-k.
    // Only the survivors are gathered and evaluated in float64 (typically ~k of the candidates).
    if (tid < 32) {
        float kth_scr = -INFINITY;
        float prev = INFINITY;
        int prev_taken = 0;  // how many entries equal to `prev` have been counted already
        for (int r = 0; r < k; ++r) {
            // r-th largest value with multiplicity, without modifying scr[]: largest value < prev, or prev again
            // while copies of it remain
            float bs = -INFINITY;
            int cnt_prev = 0;
            for (int c = tid; c < nc; c += 32) {
                const float v = scr[c];
                if (v == prev) ++cnt_prev;
                else if (v < prev && v > bs) bs = v;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                bs = fmaxf(bs, __shfl_xor_sync(0xffffffffu, bs, o));
                cnt_prev += __shfl_xor_sync(0xffffffffu, cnt_prev, o);
            }
            if (r > 0 && cnt_prev > prev_taken) {
                ++prev_taken;
                kth_scr = prev;
                continue;
            }
            if (bs == -INFINITY) {  // fewer than k candidates: keep everything
                kth_scr = -INFINITY;
                break;
            }
            prev = bs;
            prev_taken = 1;
            kth_scr = bs;
        }
        if (tid == 0) s_thr = (kth_scr == -INFINITY) ? -INFINITY : kth_scr - 2.0f * eps - 1e-6f;
    }
    __syncthreads();
    const float thr = s_thr;
    for (int c = tid; c < nc; c += R_THREADS)
        if (scr[c] >= thr) ids[atomicAdd(&s_ns, 1)] = cidx[c];
    __syncthreads();
    const int ns = s_ns;
    // one survivor per thread: float64 accumulation over d in index order (the oracle's order)
    const bool vec4 = (d % 4 == 0);
    for (int c = tid; c < ns; c += R_THREADS) {
        const int myid = ids[c];
        double acc = 0.0;
        const float* kr = k32 + static_cast<int64_t>(myid) * d;
        if (vec4) {
            const float4* kr4 = reinterpret_cast<const float4*>(kr);
            for (int64_t t = 0; t < d / 4; ++t) {
                const float4 kv = __ldg(kr4 + t);
                acc += static_cast<double>(qs[4 * t]) * static_cast<double>(kv.x);
                acc += static_cast<double>(qs[4 * t + 1]) * static_cast<double>(kv.y);
                acc += static_cast<double>(qs[4 * t + 2]) * static_cast<double>(kv.z);
                acc += static_cast<double>(qs[4 * t + 3]) * static_cast<double>(kv.w);
            }
        } else {
            for (int64_t t = 0; t < d; ++t) acc += static_cast<double>(qs[t]) * static_cast<double>(__ldg(kr + t));
        }
        sims[c] = acc;
    }
    __syncthreads();
    // k rounds of arg-best by (-sim, index)
    double kth = -DBL_MAX;
    for (int r = 0; r < k; ++r) {
        double bs = -DBL_MAX;
        int bi = INT_MAX, bc = -1;
        for (int c = tid; c < ns; c += R_THREADS) {
            const int id = ids[c];
            if (id < 0) continue;
            const double s = sims[c];
            if (s > bs || (s == bs && id < bi)) {
                bs = s;
                bi = id;
                bc = c;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, bs, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
            if (os > bs || (os == bs && oi < bi)) {
                bs = os;
                bi = oi;
                bc = oc;
            }
        }
        if ((tid & 31) == 0) {
            s_best[tid >> 5] = bs;
            s_besti[tid >> 5] = bi;
            s_bestc[tid >> 5] = bc;
        }
        __syncthreads();
        if (tid == 0) {
            int w = 0;
            if (s_best[1] > s_best[0] || (s_best[1] == s_best[0] && s_besti[1] < s_besti[0])) w = 1;
            const int bcw = s_bestc[w];
            if (bcw >= 0) {
                out_sims[qi * k + r] = s_best[w];
                out_idx[qi * k + r] = key_offset + s_besti[w];
                ids[bcw] = -1;  // taken
                sims[bcw] = -DBL_MAX;
            } else {
                out_sims[qi * k + r] = -DBL_MAX;
                out_idx[qi * k + r] = -1;
            }
            s_best[0] = s_best[w];
            s_bestc[0] = bcw;
        }
        __syncthreads();
        kth = (s_bestc[0] >= 0) ? s_best[0] : -DBL_MAX;
        __syncthreads();
    }
    // completeness proof: every key outside the candidate set lies in some sub-range whose list is
    // full; its screened score is <= that list's last score t, hence its exact score <= t + eps.
    int bad = 0;
    for (int l = tid; l < num_lists; l += R_THREADS) {
        const float t = cand_score[(qi * num_lists + l) * kp + kp - 1];
        const int32_t last = cand_idx[(qi * num_lists + l) * kp + kp - 1];
        // list not full: every key of the sub-range is either a candidate or was dropped below a row floor
        // T - 2 eps - 1e-6 (T = k-th best screened score of some earlier list), i.e. strictly below kth
        if (last < 0) continue;
        if (!(static_cast<double>(t) + static_cast<double>(eps) < kth)) bad = 1;
    }
    bad = __syncthreads_or(bad);
    if (tid == 0 && bad) {
        const int slot = atomicAdd(n_flagged, 1);
        flagged[slot] = static_cast<int32_t>(qi);
    }
}

// ------------------------------------------------------------------------------------------
// exhaustive float64 path: 64 x 64 tiles into a [rows][K] buffer, then per-row selection
// ------------------------------------------------------------------------------------------
constexpr int E_T = 64, E_KT = 16;

__global__ void __launch_bounds__(256)
knn_exact_tile_kernel(const float* __restrict__ q32, const int32_t* __restrict__ qsel, int64_t qbase, int64_t nq,
                      const float* __restrict__ k32, int64_t K, int64_t d, double* __restrict__ buf) {
    __shared__ float As[E_KT][E_T + 1];
    __shared__ float Bs[E_KT][E_T + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t col0 = static_cast<int64_t>(blockIdx.x) * E_T, r0 = static_cast<int64_t>(blockIdx.y) * E_T;
    double acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
    for (int64_t k0 = 0; k0 < d; k0 += E_KT) {
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
            const int idx = tid + 256 * qq;
            const int r = idx >> 4, kk = idx & 15;
            const int64_t gk = k0 + kk;
            float va = 0.f, vb = 0.f;
            if (r0 + r < nq && gk < d) {
                const int64_t qrow = qsel ? static_cast<int64_t>(qsel[qbase + r0 + r]) : (qbase + r0 + r);
                va = q32[qrow * d + gk];
            }
            if (col0 + r < K && gk < d) vb = k32[(col0 + r) * d + gk];
            As[kk][r] = va;
            Bs[kk][r] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < E_KT; ++kk) {  // d ascends: sequential float64 accumulation per (q,k) pair
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = static_cast<double>(As[kk][ty * 4 + i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = static_cast<double>(Bs[kk][tx * 4 + j]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t r = r0 + ty * 4 + i, c = col0 + tx * 4 + j;
            if (r < nq && c < K) buf[r * K + c] = acc[i][j];
        }
}

constexpr int SEL_THREADS = 128;
constexpr int SEL_KMAX = 16;

// one block per buffer row: top-k by (-sim, index)
__global__ void __launch_bounds__(SEL_THREADS)
knn_select_kernel(const double* __restrict__ buf, const int32_t* __restrict__ qsel, int64_t qbase, int64_t K, int k,
                  int64_t key_offset, double* __restrict__ out_sims, int64_t* __restrict__ out_idx) {
    extern __shared__ uint8_t sel_smem[];
    double* ls = reinterpret_cast<double*>(sel_smem);             // [SEL_THREADS][k]
    int32_t* li = reinterpret_cast<int32_t*>(ls + SEL_THREADS * k);  // [SEL_THREADS][k]
    __shared__ double s_bs[SEL_THREADS / 32];
    __shared__ int s_bi[SEL_THREADS / 32];
    __shared__ int s_bt[SEL_THREADS / 32];
    __shared__ int s_win;
    const int tid = threadIdx.x;
    const int64_t r = blockIdx.x;
    const double* row = buf + r * K;
    const int64_t qout = qsel ? static_cast<int64_t>(qsel[qbase + r]) : (qbase + r);
    double* my_s = ls + tid * k;
    int32_t* my_i = li + tid * k;
    for (int i = 0; i < k; ++i) {
        my_s[i] = -DBL_MAX;
        my_i[i] = -1;
    }
    for (int64_t c = tid; c < K; c += SEL_THREADS) {
        double s = row[c];
        if (s > my_s[k - 1] || my_i[k - 1] < 0) {
            int32_t id = static_cast<int32_t>(c);
            for (int i = 0; i < k; ++i) {
                if (my_i[i] < 0 || s > my_s[i]) {  // indices ascend per thread, so strict > keeps the lowest on ties
                    const double ts = my_s[i];
                    const int32_t ti = my_i[i];
                    my_s[i] = s;
                    my_i[i] = id;
                    s = ts;
                    id = ti;
                    if (id < 0) break;
                }
            }
        }
    }
    int head = 0;
    __syncthreads();
    for (int rr = 0; rr < k; ++rr) {
        double bs = (head < k && my_i[head] >= 0) ? my_s[head] : -DBL_MAX;
        int bi = (head < k && my_i[head] >= 0) ? my_i[head] : INT_MAX;
        int bt = tid;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, bs, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int ot = __shfl_xor_sync(0xffffffffu, bt, o);
            if (os > bs || (os == bs && oi < bi)) {
                bs = os;
                bi = oi;
                bt = ot;
            }
        }
        if ((tid & 31) == 0) {
            s_bs[tid >> 5] = bs;
            s_bi[tid >> 5] = bi;
            s_bt[tid >> 5] = bt;
        }
        __syncthreads();
        if (tid == 0) {
            int w = 0;
            for (int x = 1; x < SEL_THREADS / 32; ++x)
                if (s_bs[x] > s_bs[w] || (s_bs[x] == s_bs[w] && s_bi[x] < s_bi[w])) w = x;
            if (s_bi[w] != INT_MAX) {
                out_sims[qout * k + rr] = s_bs[w];
                out_idx[qout * k + rr] = key_offset + s_bi[w];
                s_win = s_bt[w];
            } else {
                out_sims[qout * k + rr] = -DBL_MAX;
                out_idx[qout * k + rr] = -1;
                s_win = -1;
            }
        }
        __syncthreads();
        if (s_win == tid) ++head;
        __syncthreads();
    }
}

// merge `parts` sorted lists per query; one warp per query
__global__ void knn_merge_kernel(const double* __restrict__ sims, const int64_t* __restrict__ idx, int parts, int64_t Q,
                                 int k, double* __restrict__ out64, float* __restrict__ out32,
                                 int64_t* __restrict__ out_idx) {
    const int64_t qi = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (qi >= Q) return;
    const int total = parts * k;
    // each lane owns entries lane, lane+32, ...; `taken` bit mask over its (<= 8) entries
    unsigned taken = 0;
    for (int r = 0; r < k; ++r) {
        double bs = -DBL_MAX;
        long long bi = LLONG_MAX;
        int be = -1;
        for (int e = lane, t = 0; e < total; e += 32, ++t) {
            if (taken & (1u << t)) continue;
            const int p = e / k, j = e % k;
            const long long id = idx[(static_cast<int64_t>(p) * Q + qi) * k + j];
            if (id < 0) continue;
            const double s = sims[(static_cast<int64_t>(p) * Q + qi) * k + j];
            if (s > bs || (s == bs && id < bi)) {
                bs = s;
                bi = id;
                be = e;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double os = __shfl_xor_sync(0xffffffffu, bs, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int oe = __shfl_xor_sync(0xffffffffu, be, o);
            if (os > bs || (os == bs && oi < bi)) {
                bs = os;
                bi = oi;
                be = oe;
            }
        }
        if (be >= 0 && (be & 31) == lane) taken |= 1u << (be >> 5);
        if (lane == 0) {
            const bool ok = be >= 0;
            if (out64) out64[qi * k + r] = ok ? bs : -DBL_MAX;
            if (out32) out32[qi * k + r] = ok ? static_cast<float>(bs) : -INFINITY;
            out_idx[qi * k + r] = ok ? bi : -1;
        }
    }
}

// accuracy counts; one thread per query
__global__ void topk_accuracy_kernel(const int64_t* __restrict__ idx, int64_t Q, int kmax,
                                     const int32_t* __restrict__ key_ids, const int32_t* __restrict__ query_ids, int k0,
                                     int k1, int k2, int k3, int nk, int32_t max_class,
                                     unsigned long long* __restrict__ micro_hits, int32_t* __restrict__ class_hit,
                                     int32_t* __restrict__ class_cnt) {
    const int64_t qi = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (qi >= Q) return;
    const int ks[4] = {k0, k1, k2, k3};
    for (int l = 0; l < 4; ++l) {
        const int32_t gt = query_ids[qi * 4 + l];
        int first = INT_MAX;  // first rank at which the ground-truth label appears
        for (int j = 0; j < kmax; ++j) {
            const int64_t id = idx[qi * kmax + j];
            if (id >= 0 && key_ids[id * 4 + l] == gt) {
                first = j;
                break;
            }
        }
        for (int a = 0; a < nk; ++a) {
            const bool hit = first < ks[a];
            if (hit) atomicAdd(&micro_hits[a * 4 + l], 1ull);
            if (gt >= 0 && gt < max_class) {
                atomicAdd(&class_cnt[(a * 4 + l) * max_class + gt], 1);
                if (hit) atomicAdd(&class_hit[(a * 4 + l) * max_class + gt], 1);
            }
        }
    }
}

struct KnnPlan {
    int64_t dpad = 0;
    int kp = 8, num_chunks = 1, num_lists = 2;
    int64_t tiles_per_chunk = 0;
    int64_t exact_rows = 0;
    size_t off_qh = 0, off_kh = 0, off_cs = 0, off_ci = 0, off_floor = 0, off_flag = 0, off_nflag = 0, off_buf = 0, total = 0;
};

size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

KnnPlan make_knn_plan(int64_t Q, int64_t K, int64_t d, int k, int path) {
    KnnPlan p;
    p.dpad = round_up(d, 64);
    p.kp = (k <= 5) ? 8 : 16;
    const int64_t num_qb = ceil_div(Q, S_QB), num_kt = ceil_div(K, S_BN);
    // chunks of the key range: enough units to fill the machine when there are few queries, and -- because every
    // query block sweeps a chunk while the CTA pairs drift apart -- small enough (<= 24 MB of 16-bit keys) that a
    // chunk stays resident in L2 instead of being re-fetched from HBM by every pair
    int64_t nc = ceil_div(2048, num_qb > 0 ? num_qb : 1);
    if (nc > 32) nc = 32;
    const int64_t l2_tiles = std::max<int64_t>(8, (int64_t(24) << 20) / (S_BN * p.dpad * 2));
    const int64_t nc_l2 = std::min<int64_t>(64, ceil_div(num_kt, l2_tiles));
    if (nc < nc_l2) nc = nc_l2;
    if (nc > num_kt) nc = num_kt;
    if (nc < 1) nc = 1;
    p.tiles_per_chunk = ceil_div(num_kt, nc);
    p.num_chunks = static_cast<int>(ceil_div(num_kt, p.tiles_per_chunk));
    p.num_lists = 2 * p.num_chunks;
    // rows of the float64 [rows][K] buffer of the exhaustive path: <= 256 MB
    int64_t rows = (int64_t(256) << 20) / (8 * (K > 0 ? K : 1));
    if (rows < 64) rows = 64;
    if (rows > Q) rows = round_up(Q, 64);
    p.exact_rows = round_up(rows, 64);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes);
        return o;
    };
    if (path != PATH_SIMT_F32) {
        p.off_qh = take(2 * static_cast<size_t>(Q) * p.dpad);
        p.off_kh = take(2 * static_cast<size_t>(K) * p.dpad);
        p.off_cs = take(sizeof(float) * Q * p.num_lists * p.kp);
        p.off_ci = take(sizeof(int32_t) * Q * p.num_lists * p.kp);
        p.off_floor = take(sizeof(float) * Q);
    }
    p.off_flag = take(sizeof(int32_t) * Q);
    p.off_nflag = take(sizeof(int32_t) * 4);
    p.off_buf = take(sizeof(double) * p.exact_rows * K);
    p.total = off;
    return p;
}

int g_sms = 0;
int sm_count() {
    if (!g_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_sms;
}

int run_exhaustive(const float* q32, const int32_t* qsel, int64_t nq_total, const float* k32, int64_t K, int64_t d, int k,
                   int64_t key_offset, const KnnPlan& plan, void* scratch, double* out_sims, int64_t* out_idx,
                   cudaStream_t s) {
    double* buf = at<double>(scratch, plan.off_buf);
    const size_t sel_smem = static_cast<size_t>(SEL_THREADS) * k * (sizeof(double) + sizeof(int32_t));
    for (int64_t base = 0; base < nq_total; base += plan.exact_rows) {
        const int64_t nq = (nq_total - base < plan.exact_rows) ? nq_total - base : plan.exact_rows;
        dim3 grid(static_cast<unsigned>(ceil_div(K, E_T)), static_cast<unsigned>(ceil_div(nq, E_T)));
        knn_exact_tile_kernel<<<grid, 256, 0, s>>>(q32, qsel, base, nq, k32, K, d, buf);
        CLIBD_KERNEL_CHECK();
        knn_select_kernel<<<static_cast<unsigned>(nq), SEL_THREADS, sel_smem, s>>>(buf, qsel, base, K, k, key_offset,
                                                                                  out_sims, out_idx);
        CLIBD_KERNEL_CHECK();
    }
    return 0;
}

}  // namespace

// launch_make_operands (loss_plan.h, loss_support.cu) is reused here with inv == nullptr (rows already normalised)

}  // namespace clibd

using namespace clibd;

extern "C" {

int clibd_knn_normalize(const void* x, int dtype, int64_t n, int64_t d, float* out, clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_knn_normalize");
    CLIBD_REQUIRE(x && out && n >= 0 && d > 0, "null pointer or bad shape");
    CLIBD_REQUIRE(dtype == DT_F32 || dtype == kDT_F64, "knn_normalize takes float32 (0) or float64 (3)");
    if (n == 0) return 0;
    const int64_t blocks = ceil_div(n * 32, 256);
    if (dtype == DT_F32)
        knn_normalize_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(x), n, d, out);
    else
        knn_normalize_kernel<double><<<blocks, 256, 0, stream>>>(static_cast<const double*>(x), n, d, out);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int64_t clibd_knn_scratch_bytes(int64_t n_query, int64_t n_key, int64_t d, int k, int path) {
    if (n_query <= 0 || n_key <= 0 || d <= 0 || k <= 0 || path < 0 || path > 2) return -1;
    return static_cast<int64_t>(make_knn_plan(n_query, n_key, d, k, path).total);
}

int clibd_knn_search(const float* q32, int64_t Q, const float* keys32, int64_t K, int64_t key_offset, int64_t d, int k,
                     int path, void* scratch, int64_t scratch_bytes, double* out_sims64, int64_t* out_idx,
                     int32_t* n_exhaustive, clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_knn_search");
    CLIBD_REQUIRE(q32 && keys32 && out_sims64 && out_idx && n_exhaustive, "null pointer");
    CLIBD_REQUIRE(Q > 0 && K > 0 && d > 0, "bad shape");
    CLIBD_REQUIRE(k >= 1 && k <= SEL_KMAX, "k must be in [1,16]");
    CLIBD_REQUIRE(K < (int64_t(1) << 31) - 512 && Q < (int64_t(1) << 31) - 512, "too many rows for 32-bit indices");
    CLIBD_REQUIRE(path >= 0 && path <= 2, "path must be 0, 1 or 2");
    if (path != PATH_SIMT_F32 && k > 12) path = PATH_SIMT_F32;  // candidate lists hold 16 entries at most
    const KnnPlan plan = make_knn_plan(Q, K, d, k, path);
    CLIBD_REQUIRE(scratch && scratch_bytes >= static_cast<int64_t>(plan.total), "scratch too small");
    int32_t* flagged = at<int32_t>(scratch, plan.off_flag);
    int32_t* nflag = at<int32_t>(scratch, plan.off_nflag);
    if (path == PATH_SIMT_F32) {
        int rc = run_exhaustive(q32, nullptr, Q, keys32, K, d, k, key_offset, plan, scratch, out_sims64, out_idx, stream);
        if (rc) return rc;
        const int32_t all = static_cast<int32_t>(Q);
        CLIBD_CHECK_CUDA(cudaMemcpyAsync(n_exhaustive, &all, sizeof(int32_t), cudaMemcpyHostToDevice, stream));
        CLIBD_CHECK_CUDA(cudaStreamSynchronize(stream));  // `all` lives on this stack frame
        return 0;
    }
    CLIBD_REQUIRE(clibd_device_supported(), "tcgen05 path needs a compute-capability 10.x device");
    const int fmt_bf16 = path == PATH_TC_BF16 ? 1 : 0;
    void* qh = at<void>(scratch, plan.off_qh);
    void* kh = at<void>(scratch, plan.off_kh);
    float* cs = at<float>(scratch, plan.off_cs);
    int32_t* ci = at<int32_t>(scratch, plan.off_ci);
    int rc;
    if ((rc = launch_make_operands(q32, DT_F32, nullptr, Q, d, plan.dpad, 0, fmt_bf16, qh, nullptr, stream))) return rc;
    if ((rc = launch_make_operands(keys32, DT_F32, nullptr, K, d, plan.dpad, 0, fmt_bf16, kh, nullptr, stream))) return rc;
    CUtensorMap tm_q, tm_k;
    if ((rc = make_tmap_2d_16bit(&tm_q, qh, Q, plan.dpad, plan.dpad, S_BK, S_BM, fmt_bf16))) return rc;
    if ((rc = make_tmap_2d_16bit(&tm_k, kh, K, plan.dpad, plan.dpad, S_BK, S_BN / 2, fmt_bf16))) return rc;
    const int64_t units = ceil_div(Q, S_QB) * plan.num_chunks;
    const int64_t max_pairs = sm_count() / 2;
    const int grid = 2 * static_cast<int>(units < max_pairs ? units : max_pairs);
    const uint32_t idesc = make_idesc_f16(S_QB, S_BN, fmt_bf16 ? 1u : 0u);
    // rigorous bound on |screened - exact| for unit-norm rows: operand rounding (2u + u^2) with
    // u = 2^-11 (f16) or 2^-8 (bf16), f16 subnormal flush (<= 2 * sqrt(d) * 2^-25), and fp32
    // accumulation of d exact products inside the tensor core (<= d * 2^-22).
    const double u = fmt_bf16 ? 0.00390625 : 0.00048828125;
    const double eps = (2.0 * u + u * u) * 1.0001 + 2.0 * sqrt(static_cast<double>(d)) * 2.98e-8 +
                       static_cast<double>(d) * 2.384185791015625e-07;
    const float floor_delta = 2.0f * static_cast<float>(eps) + 1e-6f;
    float* rowfloor = at<float>(scratch, plan.off_floor);
    knn_fill_kernel<<<static_cast<unsigned>(ceil_div(Q, 256)), 256, 0, stream>>>(rowfloor, Q, -INFINITY);
    CLIBD_KERNEL_CHECK();
    {
    ProfScope prof(PROF_KNN_SCREEN_TC, stream);
    if (plan.kp == 8) {
        CLIBD_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&knn_screen_tc_kernel<8>), S_SMEM_ALLOC));
        knn_screen_tc_kernel<8><<<grid, S_THREADS, S_SMEM_ALLOC, stream>>>(tm_q, tm_k, Q, K, static_cast<int>(plan.dpad / S_BK),
                                                                           plan.num_chunks, plan.tiles_per_chunk, idesc, k, floor_delta, rowfloor, cs, ci);
    } else {
        CLIBD_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&knn_screen_tc_kernel<16>), S_SMEM_ALLOC));
        knn_screen_tc_kernel<16><<<grid, S_THREADS, S_SMEM_ALLOC, stream>>>(tm_q, tm_k, Q, K, static_cast<int>(plan.dpad / S_BK),
                                                                            plan.num_chunks, plan.tiles_per_chunk, idesc, k, floor_delta, rowfloor, cs, ci);
    }
    }
    CLIBD_KERNEL_CHECK();
    CLIBD_CHECK_CUDA(cudaMemsetAsync(nflag, 0, sizeof(int32_t) * 4, stream));
    const int C = plan.num_lists * plan.kp;
    const size_t rr_smem = static_cast<size_t>(C) * (sizeof(double) + sizeof(int32_t) + sizeof(float)) + sizeof(float) * d;
    CLIBD_REQUIRE(rr_smem <= 48 * 1024, "too many candidate lists for the re-rank kernel");
    {
    ProfScope prof(PROF_KNN_RERANK, stream);
    knn_rerank_kernel<<<static_cast<unsigned>(Q), R_THREADS, rr_smem, stream>>>(
        q32, keys32, Q, K, d, k, plan.kp, plan.num_lists, static_cast<float>(eps), key_offset, cs, ci, out_sims64, out_idx,
        flagged, nflag);
    }
    CLIBD_KERNEL_CHECK();
    int32_t h_nflag = 0;
    CLIBD_CHECK_CUDA(cudaMemcpyAsync(&h_nflag, nflag, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
    CLIBD_CHECK_CUDA(cudaStreamSynchronize(stream));
    CLIBD_CHECK_CUDA(cudaMemcpyAsync(n_exhaustive, nflag, sizeof(int32_t), cudaMemcpyDeviceToDevice, stream));
    if (h_nflag > 0) {
        rc = run_exhaustive(q32, flagged, h_nflag, keys32, K, d, k, key_offset, plan, scratch, out_sims64, out_idx, stream);
        if (rc) return rc;
    }
    return 0;
}

int clibd_knn_merge(const double* sims64, const int64_t* idx, int parts, int64_t Q, int k, double* out_sims64,
                    float* out_sims32, int64_t* out_idx, clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_knn_merge");
    CLIBD_REQUIRE(sims64 && idx && out_idx && parts >= 1 && Q > 0 && k >= 1, "bad arguments");
    CLIBD_REQUIRE(static_cast<int64_t>(parts) * k <= 32 * 32, "parts * k must be <= 1024");
    knn_merge_kernel<<<static_cast<unsigned>(ceil_div(Q * 32, 256)), 256, 0, stream>>>(sims64, idx, parts, Q, k, out_sims64,
                                                                                      out_sims32, out_idx);
    CLIBD_KERNEL_CHECK();
    return 0;
}

int clibd_topk_accuracy(const int64_t* idx, int64_t Q, int kmax, const int32_t* key_ids, int64_t K,
                        const int32_t* query_ids, const int32_t* k_list, int nk, int32_t max_class,
                        int64_t* micro_hits, int32_t* class_hit, int32_t* class_cnt, clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_topk_accuracy");
    CLIBD_REQUIRE(idx && key_ids && query_ids && k_list && micro_hits && class_hit && class_cnt, "null pointer");
    CLIBD_REQUIRE(Q > 0 && K > 0 && kmax >= 1 && nk >= 1 && nk <= 4 && max_class >= 1, "bad arguments (nk <= 4)");
    int ks[4] = {0, 0, 0, 0};
    for (int a = 0; a < nk; ++a) {
        CLIBD_REQUIRE(k_list[a] >= 1 && k_list[a] <= kmax, "k_list entries must be in [1, kmax]");
        ks[a] = k_list[a];
    }
    CLIBD_CHECK_CUDA(cudaMemsetAsync(micro_hits, 0, sizeof(int64_t) * nk * 4, stream));
    CLIBD_CHECK_CUDA(cudaMemsetAsync(class_hit, 0, sizeof(int32_t) * nk * 4 * max_class, stream));
    CLIBD_CHECK_CUDA(cudaMemsetAsync(class_cnt, 0, sizeof(int32_t) * nk * 4 * max_class, stream));
    topk_accuracy_kernel<<<static_cast<unsigned>(ceil_div(Q, 256)), 256, 0, stream>>>(
        idx, Q, kmax, key_ids, query_ids, ks[0], ks[1], ks[2], ks[3], nk, max_class,
        reinterpret_cast<unsigned long long*>(micro_hits), class_hit, class_cnt);
    CLIBD_KERNEL_CHECK();
    return 0;
}

}  // extern "C"
