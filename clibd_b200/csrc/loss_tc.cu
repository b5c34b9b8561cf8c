// tcgen05 / TMEM / TMA kernels of the contrastive loss (path codes 1 and 2), sm_100a only.
//
// Forward (loss_fwd_tc_kernel): persistent CTAs walk 128 x 256 tiles of S = Xhat Yhat^T.
//   warp 0  : TMA producer   (4-stage ring, 64-wide K chunks, 128B swizzle)
//   warp 1  : tcgen05.mma issuer (one thread), accumulators double-buffered in TMEM (2 x 256 cols)
//   warp 2  : TMEM allocator
//   warps 4-11: epilogue. tcgen05.ld 32x32b -> e = exp2(s*log2e*(cos-1)) -> per-row sums in
//             registers, per-column sums by a 31-shuffle transposing butterfly; per-tile partials
//             go to HBM (1.5 KB per 590 KB of operand traffic).  S itself never leaves TMEM.
// The fixed shift exp(S - s) is exact algebra because |cos| <= 1 (loss_func.py:55-60 normalises
// both operands), so one exponential serves the row LSE of CE(S,T) and the column LSE of
// CE(S^T,T) (loss_func.py:65-66).
//
// Backward (loss_bwd_tc_kernel): one CTA owns 128 rows of dXhat x a <=384-wide chunk of d
//   (384 fp32 accumulator columns in TMEM) and sweeps the columns of S in steps of 128:
//   S tile (tcgen05, TMEM cols 384..511) -> epilogue G~ = e * (rowcoef_i + colcoef_j) -> 16-bit,
//   swizzled into shared memory as the A operand -> dXhat += G~ * Yhat (B operand from the
//   transposed 16-bit copy so that every operand is K-major).  The label-matched "-2 T" term
//   is not in this kernel: it is a class-sum of O(N d) done in loss_support.cu.
#include <cstdlib>

#include "common.cuh"
#include "loss_plan.h"
#include "ptx.cuh"
#include "tmap.h"

namespace clibd {
namespace {

using namespace ptx;

constexpr float kLog2e = 1.4426950408889634f;

// ------------------------------------------------------------------------------------------
// Forward
// ------------------------------------------------------------------------------------------
constexpr int F_STAGES = 4;
constexpr int F_BK = 64;
constexpr int F_A_BYTES = FWD_BM * F_BK * 2;  // 16 KB
constexpr int F_B_BYTES = FWD_BN * F_BK * 2;  // 32 KB
constexpr int F_STAGE_BYTES = F_A_BYTES + F_B_BYTES;
constexpr int F_THREADS = 384;
constexpr int F_EPI_WARPS = 8;
constexpr int F_SMEM_COLBUF = F_STAGES * F_STAGE_BYTES;               // float [2][4][256]
constexpr int F_SMEM_ROWBUF = F_SMEM_COLBUF + 2 * 4 * FWD_BN * 4;     // float [2][2][128]
constexpr int F_SMEM_BARS = F_SMEM_ROWBUF + 2 * 2 * FWD_BM * 4;       // uint64 barriers
constexpr int F_NUM_BARS = 2 * F_STAGES + 4;
constexpr int F_SMEM_TMEMPTR = F_SMEM_BARS + F_NUM_BARS * 8;
constexpr int F_SMEM_TOTAL = F_SMEM_TMEMPTR + 16;
constexpr int F_SMEM_ALLOC = F_SMEM_TOTAL + 1024;  // slack for manual 1024 B alignment

__device__ __forceinline__ void tile_coords(int64_t t, int64_t num_mt, int64_t num_nt, int64_t& mt, int64_t& nt) {
    constexpr int64_t GM = 8;
    const int64_t group = GM * num_nt;
    const int64_t g = t / group;
    const int64_t first = g * GM;
    const int64_t gm = (num_mt - first) < GM ? (num_mt - first) : GM;
    const int64_t r = t % group;
    mt = first + r % gm;
    nt = r / gm;
}

__global__ void __launch_bounds__(F_THREADS, 1)
loss_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int64_t N,
                   int64_t row0, int64_t n, int num_kb, float scale, uint32_t idesc, float* __restrict__ rowpart,
                   float* __restrict__ colpart, const float* __restrict__ scale_dev) {
    scale = eff_scale(scale, scale_dev);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* colbuf = reinterpret_cast<float*>(smem + F_SMEM_COLBUF);
    float* rowbuf = reinterpret_cast<float*>(smem + F_SMEM_ROWBUF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F_SMEM_BARS);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + F_STAGES;
    uint64_t* tfull_bar = bars + 2 * F_STAGES;
    uint64_t* tempty_bar = bars + 2 * F_STAGES + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + F_SMEM_TMEMPTR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t num_mt = (n + FWD_BM - 1) / FWD_BM;
    const int64_t num_nt = (N + FWD_BN - 1) / FWD_BN;
    const int64_t num_tiles = num_mt * num_nt;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_a);
        prefetch_tmap(&tm_b);
        for (int i = 0; i < F_STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], F_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // Producer and MMA warps stay converged (all 32 lanes walk the loops, one elected lane issues):
    // loop state then lives in uniform registers, which is what UTMALDG / UTCHMMA take as operands.
    const uint32_t smem_base = smem_u32(smem);
    if (warp == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            int64_t mt, nt;
            tile_coords(t, num_mt, num_nt, mt, nt);
            const int32_t arow = static_cast<int32_t>(row0 + mt * FWD_BM);
            const int32_t brow = static_cast<int32_t>(nt * FWD_BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    uint8_t* sa = smem + stage * F_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], F_STAGE_BYTES);
                    tma_load_2d(&tm_a, &full_bar[stage], sa, kb * F_BK, arow, kEvictNormal);
                    tma_load_2d(&tm_b, &full_bar[stage], sa + F_A_BYTES, kb * F_BK, brow, kEvictNormal);
                }
                __syncwarp();
                if (++stage == F_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1) {
        int stage = 0;
        uint32_t phase = 0;
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            const uint32_t as = it & 1, aph = (it >> 1) & 1;
            mbar_wait(&tempty_bar[as], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * FWD_BN;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_base + stage * F_STAGE_BYTES;
                const uint64_t da = make_sw128_kmajor_desc(sa);
                const uint64_t db = make_sw128_kmajor_desc(sa + F_A_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < F_BK / 16; ++k)
                        umma_f16(d_tmem, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc, (kb | k) ? 1u : 0u);
                    umma_commit(&empty_bar[stage]);
                    if (kb == num_kb - 1) umma_commit(&tfull_bar[as]);
                }
                __syncwarp();
                if (++stage == F_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        const int ew = warp - 4;
        const int q = warp & 3;   // TMEM lane quadrant this warp may access
        const int h = ew >> 2;    // column half of the tile
        const int etid = ew * 32 + lane;
        const float a = scale * kLog2e;
        const float nb = -softmax_shift(scale) * kLog2e;
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            int64_t mt, nt;
            tile_coords(t, num_mt, num_nt, mt, nt);
            const uint32_t as = it & 1, aph = (it >> 1) & 1;
            const int buf = it & 1;
            const int64_t lrow = mt * FWD_BM + q * 32 + lane;
            const bool row_ok = lrow < n;
            const int64_t colbase = nt * FWD_BN + h * 128;
            const bool full_tile = (mt * FWD_BM + FWD_BM <= n) && (nt * FWD_BN + FWD_BN <= N);
            mbar_wait(&tfull_bar[as], aph);
            tc_fence_after();
            float rsum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * FWD_BN + h * 128 + c * 32;
                tmem_ld_32x32b_x32(taddr, v);
                tmem_ld_wait();
                if (c == 3) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tempty_bar[as]);
                }
                float e[32];
                if (full_tile) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) e[k] = ex2_approx(fmaf(__uint_as_float(v[k]), a, nb));
                } else {
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const bool ok = row_ok && (colbase + c * 32 + k < N);
                        e[k] = ok ? ex2_approx(fmaf(__uint_as_float(v[k]), a, nb)) : 0.f;
                    }
                }
#pragma unroll
                for (int k = 0; k < 32; ++k) rsum += e[k];
                // transposing butterfly: afterwards lane l holds the sum over this warp's 32 rows
                // of column (c*32 + l)
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const bool up = (lane & step) != 0;
#pragma unroll
                    for (int k = 0; k < step; ++k) {
                        const float send = up ? e[k] : e[k + step];
                        const float keep = up ? e[k + step] : e[k];
                        e[k] = keep + __shfl_xor_sync(0xffffffffu, send, step);
                    }
                }
                colbuf[(buf * 4 + q) * FWD_BN + h * 128 + c * 32 + lane] = e[0];
            }
            rowbuf[(buf * 2 + h) * FWD_BM + q * 32 + lane] = rsum;
            asm volatile("bar.sync 1, %0;" ::"n"(F_EPI_WARPS * 32) : "memory");
            {
                const int j = etid;  // 0..255: one column of the tile each
                const float cs = colbuf[(buf * 4 + 0) * FWD_BN + j] + colbuf[(buf * 4 + 1) * FWD_BN + j] +
                                 colbuf[(buf * 4 + 2) * FWD_BN + j] + colbuf[(buf * 4 + 3) * FWD_BN + j];
                const int64_t gcol = nt * FWD_BN + j;
                if (gcol < N) colpart[mt * N + gcol] = cs;
                if (etid < FWD_BM) {
                    const float rs = rowbuf[(buf * 2 + 0) * FWD_BM + etid] + rowbuf[(buf * 2 + 1) * FWD_BM + etid];
                    const int64_t lr = mt * FWD_BM + etid;
                    if (lr < n) rowpart[nt * n + lr] = rs;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// Backward
// ------------------------------------------------------------------------------------------
constexpr int B_BK = 64;
constexpr int B_TILE_BYTES = 128 * B_BK * 2;          // 16 KB: a 128-row x 64-column 16-bit operand tile
constexpr int B_STAGE_BYTES = 4 * B_TILE_BYTES;       // 64 KB: two K blocks of X and of Y, or the YhatT tiles of one K block
constexpr int B_STAGES = 3;                           // 192 KB of operands in flight
constexpr int B_G_BYTES = BWD_BM * BWD_BJ * 2;        // 32 KB (two 64-wide K blocks)
constexpr int B_THREADS = 256;
constexpr uint32_t B_TMEM_S_COL = 384;
constexpr int B_SMEM_G = B_STAGES * B_STAGE_BYTES;
constexpr int B_SMEM_CC = B_SMEM_G + B_G_BYTES;       // float [2][128]
constexpr int B_SMEM_BARS = B_SMEM_CC + 2 * BWD_BJ * 4;
constexpr int B_NUM_BARS = 2 * B_STAGES + 5;
constexpr int B_SMEM_TMEMPTR = B_SMEM_BARS + B_NUM_BARS * 8;
constexpr int B_SMEM_TOTAL = B_SMEM_TMEMPTR + 16;
constexpr int B_SMEM_ALLOC = B_SMEM_TOTAL + 1024;
static_assert(B_SMEM_ALLOC <= 232448, "backward kernel shared memory exceeds 227 KB");
static_assert(F_SMEM_ALLOC <= 232448, "forward kernel shared memory exceeds 227 KB");

// One ring of three 64 KB stages serves both GEMMs; the producer warp and the MMA warp walk the SAME
// stage sequence:   for each column tile t:  ceil(num_kb/2) S stages  (X k, X k+1, Y k, Y k+1 : 8 MMAs)
//                   then, for tile t-1:      2 gradient stages        (the YhatT pieces of one K block)
// Few, fat stages matter: issuing is paced by one mbarrier wait + one commit per stage, and with
// N=128 an MMA lasts only 64 cycles, so 4-MMA stages left the tensor pipe waiting on the issuer.
__global__ void __launch_bounds__(B_THREADS, 1)
loss_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                   const __grid_constant__ CUtensorMap tm_yt, int64_t N, int64_t ld, int64_t dvalid, int64_t row0,
                   int64_t n, int num_kb, int chunk_w, int npieces, int64_t tiles_per_split, float scale,
                   uint32_t idesc_s, uint32_t idesc_g, int fmt_bf16, const float* __restrict__ rowcoef,
                   const float* __restrict__ colcoef, const float* __restrict__ gscale, float weight, int accumulate,
                   float* __restrict__ dxh, const float* __restrict__ scale_dev) {
    scale = eff_scale(scale, scale_dev);
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* gbuf = smem + B_SMEM_G;
    float* ccbuf = reinterpret_cast<float*>(smem + B_SMEM_CC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_SMEM_BARS);
    uint64_t* full = bars;
    uint64_t* empty = bars + B_STAGES;
    uint64_t* st_full = bars + 2 * B_STAGES;  // S tile ready in TMEM
    uint64_t* st_empty = st_full + 1;         // S tile drained by the epilogue
    uint64_t* g_full = st_full + 2;           // G~ tile written to smem
    uint64_t* g_empty = st_full + 3;          // G~ tile consumed by the MMA
    uint64_t* acc_full = st_full + 4;         // all accumulation finished
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + B_SMEM_TMEMPTR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t mt = blockIdx.x;
    const int64_t dc = blockIdx.y;
    const int64_t split = blockIdx.z;
    const int64_t num_jt = (N + BWD_BJ - 1) / BWD_BJ;
    const int64_t jt0 = split * tiles_per_split;
    const int64_t jt1 = (jt0 + tiles_per_split < num_jt) ? (jt0 + tiles_per_split) : num_jt;
    const int piece_w = chunk_w / npieces;                       // accumulator columns per gradient MMA
    const uint32_t piece_bytes = static_cast<uint32_t>(piece_w) * B_BK * 2;
    const int num_sst = (num_kb + 1) / 2;                        // S stages per column tile

    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_x);
        prefetch_tmap(&tm_y);
        prefetch_tmap(&tm_yt);
        for (int i = 0; i < B_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(st_full, 1);
        mbar_init(st_empty, 4);
        mbar_init(g_full, 128);
        mbar_init(g_empty, 1);
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t smem_base = smem_u32(smem);

    if (jt0 < jt1) {
        if (warp == 0) {  // producer warp (converged; one elected lane issues)
            int stage = 0;
            uint32_t phase = 0;
            const int32_t xrow = static_cast<int32_t>(row0 + mt * BWD_BM);
            auto advance = [&]() {
                if (++stage == B_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            auto load_grad = [&](int64_t t) {
                for (int kb2 = 0; kb2 < BWD_BJ / B_BK; ++kb2) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (elect_one()) {
                        uint8_t* sb = smem + stage * B_STAGE_BYTES;
                        mbar_arrive_expect_tx(&full[stage], piece_bytes * npieces);
                        for (int pc = 0; pc < npieces; ++pc)
                            tma_load_2d(&tm_yt, &full[stage], sb + pc * piece_bytes,
                                        static_cast<int32_t>(t * BWD_BJ + kb2 * B_BK),
                                        static_cast<int32_t>(dc * BWD_DCH + pc * piece_w), kEvictNormal);
                    }
                    __syncwarp();
                    advance();
                }
            };
            for (int64_t t = jt0; t < jt1; ++t) {
                const int32_t yrow = static_cast<int32_t>(t * BWD_BJ);
                for (int ss = 0; ss < num_sst; ++ss) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (elect_one()) {
                        uint8_t* sb = smem + stage * B_STAGE_BYTES;
                        const int nkk = (num_kb - 2 * ss) < 2 ? (num_kb - 2 * ss) : 2;
                        mbar_arrive_expect_tx(&full[stage], nkk * 2 * B_TILE_BYTES);
                        for (int kk = 0; kk < nkk; ++kk) {
                            tma_load_2d(&tm_x, &full[stage], sb + kk * B_TILE_BYTES, (2 * ss + kk) * B_BK, xrow, kEvictNormal);
                            tma_load_2d(&tm_y, &full[stage], sb + (2 + kk) * B_TILE_BYTES, (2 * ss + kk) * B_BK, yrow,
                                        kEvictNormal);
                        }
                    }
                    __syncwarp();
                    advance();
                }
                if (t > jt0) load_grad(t - 1);
            }
            load_grad(jt1 - 1);
        } else if (warp == 1) {  // MMA warp (converged; one elected lane issues)
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t s_tmem = tmem_base + B_TMEM_S_COL;
            const uint32_t gaddr = smem_u32(gbuf);
            auto advance = [&]() {
                if (++stage == B_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            auto issue_grad = [&](int64_t tl) {  // tl = t - jt0 of the G~ tile to consume
                mbar_wait(g_full, tl & 1);
                tc_fence_after();
                for (int kb2 = 0; kb2 < BWD_BJ / B_BK; ++kb2) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t da = make_sw128_kmajor_desc(gaddr + kb2 * (BWD_BM * 128));
                    const uint32_t sb = smem_base + stage * B_STAGE_BYTES;
                    if (elect_one()) {
                        for (int pc = 0; pc < npieces; ++pc) {
                            const uint64_t db = make_sw128_kmajor_desc(sb + pc * piece_bytes);
#pragma unroll
                            for (int k = 0; k < B_BK / 16; ++k)
                                umma_f16(tmem_base + pc * piece_w, desc_advance(da, k * 32), desc_advance(db, k * 32),
                                         idesc_g, (tl > 0 || kb2 > 0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit(&empty[stage]);
                        if (kb2 == BWD_BJ / B_BK - 1) umma_commit(g_empty);
                    }
                    __syncwarp();
                    advance();
                }
            };
            for (int64_t t = jt0; t < jt1; ++t) {
                const int64_t tl = t - jt0;
                mbar_wait(st_empty, (tl & 1) ^ 1);
                tc_fence_after();
                for (int ss = 0; ss < num_sst; ++ss) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sb = smem_base + stage * B_STAGE_BYTES;
                    const int nkk = (num_kb - 2 * ss) < 2 ? (num_kb - 2 * ss) : 2;
                    if (elect_one()) {
                        for (int kk = 0; kk < nkk; ++kk) {
                            const uint64_t da = make_sw128_kmajor_desc(sb + kk * B_TILE_BYTES);
                            const uint64_t db = make_sw128_kmajor_desc(sb + (2 + kk) * B_TILE_BYTES);
#pragma unroll
                            for (int k = 0; k < B_BK / 16; ++k)
                                umma_f16(s_tmem, desc_advance(da, k * 32), desc_advance(db, k * 32), idesc_s,
                                         (ss | kk | k) ? 1u : 0u);
                        }
                        umma_commit(&empty[stage]);
                        if (ss == num_sst - 1) umma_commit(st_full);
                    }
                    __syncwarp();
                    advance();
                }
                if (tl > 0) issue_grad(tl - 1);
            }
            issue_grad(jt1 - jt0 - 1);
            if (elect_one()) umma_commit(acc_full);
            __syncwarp();
        } else if (warp >= 4) {
            const int q = warp & 3;
            const int etid = (warp - 4) * 32 + lane;  // == row of the tile
            const int64_t lrow = mt * BWD_BM + etid;
            const float gs = gscale[0];
            const float rcg = (lrow < n ? rowcoef[row0 + lrow] : 0.f) * gs;
            const float a = scale * kLog2e;
            const float nb = -softmax_shift(scale) * kLog2e;
            const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
            for (int64_t t = jt0; t < jt1; ++t) {
                const int64_t tl = t - jt0;
                float* cc = ccbuf + (tl & 1) * BWD_BJ;
                {
                    const int64_t gj = t * BWD_BJ + etid;
                    cc[etid] = (gj < N) ? colcoef[gj] * gs : 0.f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                mbar_wait(st_full, tl & 1);
                tc_fence_after();
                uint32_t packed[BWD_BJ / 2];
#pragma unroll
                for (int c = 0; c < BWD_BJ / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_base + lane_base + B_TMEM_S_COL + c * 32, v);
                    tmem_ld_wait();
                    if (c == BWD_BJ / 32 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(st_empty);
                    }
#pragma unroll
                    for (int k = 0; k < 32; k += 2) {
                        const float e0 = ex2_approx(fmaf(__uint_as_float(v[k]), a, nb));
                        const float e1 = ex2_approx(fmaf(__uint_as_float(v[k + 1]), a, nb));
                        const float g0 = e0 * (rcg + cc[c * 32 + k]);
                        const float g1 = e1 * (rcg + cc[c * 32 + k + 1]);
                        packed[c * 16 + k / 2] = pack2_operand16(g0, g1, fmt_bf16);
                    }
                }
                // the previous G~ tile must have been consumed before it is overwritten
                mbar_wait(g_empty, (tl & 1) ^ 1);
                const uint32_t rowaddr = smem_u32(gbuf) + etid * 128;
#pragma unroll
                for (int kb2 = 0; kb2 < BWD_BJ / B_BK; ++kb2) {
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const uint32_t addr = rowaddr + kb2 * (BWD_BM * 128) + ((ch ^ (etid & 7)) << 4);
                        const int p = kb2 * 32 + ch * 4;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[p]),
                                     "r"(packed[p + 1]), "r"(packed[p + 2]), "r"(packed[p + 3])
                                     : "memory");
                    }
                }
                fence_proxy_async_smem();
                mbar_arrive(g_full);
            }
            // drain the accumulators: dxh (+)= weight / gscale * acc
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const float wgt = weight * gscale[1];
            float* out = dxh + (split * n + lrow) * ld + dc * BWD_DCH;
            for (int c = 0; c < chunk_w / 32; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + lane_base + c * 32, v);
                tmem_ld_wait();
                if (lrow < n) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const int64_t col = dc * BWD_DCH + c * 32 + k;
                        if (col < dvalid) {
                            const float val = wgt * __uint_as_float(v[k]);
                            out[c * 32 + k] = accumulate ? out[c * 32 + k] + val : val;
                        }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // empty column range for this split: contribute zeros unless accumulating
        const int etid = (warp - 4) * 32 + lane;
        const int64_t lrow = mt * BWD_BM + etid;
        if (!accumulate && lrow < n) {
            float* out = dxh + (split * n + lrow) * ld;
            for (int c = 0; c < chunk_w; ++c) {
                const int64_t col = dc * BWD_DCH + c;
                if (col < dvalid) out[col] = 0.f;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

int g_num_sms = 0;
int num_sms() {
    if (g_num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_num_sms;
}

}  // namespace

int tc_num_sms() { return num_sms(); }

int tc_forward_pair(const void* xh_a, const void* xh_b, int64_t N, int64_t dpad, int64_t row0, int64_t n, float scale,
                    int fmt_bf16, float* rowpart, float* colpart, cudaStream_t s) {
    if (n == 0 || N == 0) return 0;
    CLIBD_REQUIRE(dpad % F_BK == 0, "padded feature dim must be a multiple of 64");
    if (!std::getenv("CLIBD_FWD_SINGLE"))
        return tc_forward_pair_cg2(xh_a, xh_b, N, dpad, row0, n, scale, fmt_bf16, rowpart, colpart, num_sms(), s);
    CUtensorMap tm_a, tm_b;
    int rc = make_tmap_2d_16bit(&tm_a, xh_a, row0 + n, dpad, dpad, F_BK, FWD_BM, fmt_bf16);  // local rows only
    if (rc) return rc;
    rc = make_tmap_2d_16bit(&tm_b, xh_b, N, dpad, dpad, F_BK, FWD_BN, fmt_bf16);
    if (rc) return rc;
    CLIBD_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&loss_fwd_tc_kernel), F_SMEM_ALLOC));
    const int64_t tiles = ceil_div(n, FWD_BM) * ceil_div(N, FWD_BN);
    const int grid = static_cast<int>(tiles < num_sms() ? tiles : num_sms());
    const uint32_t idesc = make_idesc_f16(FWD_BM, FWD_BN, fmt_bf16 ? 1u : 0u);
    ProfScope prof(PROF_LOSS_FWD_TC, s);
    loss_fwd_tc_kernel<<<grid, F_THREADS, F_SMEM_ALLOC, s>>>(tm_a, tm_b, N, row0, n, static_cast<int>(dpad / F_BK), scale,
                                                            idesc, rowpart, colpart, scale_dev_ptr());
    CLIBD_KERNEL_CHECK();
    return 0;
}

int tc_backward_rows(const void* xh_x, const void* xh_y, const void* xhT_y, int64_t N, int64_t npad, int64_t d,
                     int64_t dpad, int64_t row0, int64_t n, float scale, const float* rowcoef, const float* colcoef,
                     const float* gscale, float weight, int accumulate, int jsplit, int fmt_bf16, float* dxh,
                     cudaStream_t s) {
    if (n == 0 || N == 0) return 0;
    CLIBD_REQUIRE(dpad % B_BK == 0, "padded feature dim must be a multiple of 64");
    CLIBD_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&loss_bwd_tc_kernel), B_SMEM_ALLOC));
    CUtensorMap tm_x, tm_y, tm_yt;
    int rc = make_tmap_2d_16bit(&tm_x, xh_x, row0 + n, dpad, dpad, B_BK, BWD_BM, fmt_bf16);  // local rows only
    if (rc) return rc;
    rc = make_tmap_2d_16bit(&tm_y, xh_y, N, dpad, dpad, B_BK, BWD_BJ, fmt_bf16);
    if (rc) return rc;
    const int64_t num_jt = ceil_div(N, BWD_BJ);
    const int64_t tiles_per_split = ceil_div(num_jt, jsplit);
    const uint32_t idesc_s = make_idesc_f16(BWD_BM, BWD_BJ, fmt_bf16 ? 1u : 0u);
    // d chunks: full BWD_DCH-wide chunks (grid.y), then one narrower remainder chunk (multiple of 64)
    const int64_t full_chunks = dpad / BWD_DCH;
    const int64_t rem = dpad % BWD_DCH;
    for (int pass = 0; pass < 2; ++pass) {
        const int64_t nch = pass == 0 ? full_chunks : (rem ? 1 : 0);
        if (nch == 0) continue;
        const int chunk_w = pass == 0 ? BWD_DCH : static_cast<int>(rem);
        const int npieces = chunk_w > 256 ? 2 : 1;   // gradient MMA N = chunk_w / npieces (<= 256, multiple of 32)
        const int piece_w = chunk_w / npieces;
        const int64_t dc0 = pass == 0 ? 0 : full_chunks;
        dim3 grid(static_cast<unsigned>(ceil_div(n, BWD_BM)), static_cast<unsigned>(nch), static_cast<unsigned>(jsplit));
        // the remainder pass folds its chunk offset into the dxh pointer, the valid-column count and the
        // base row of the transposed-operand map
        rc = make_tmap_2d_16bit(&tm_yt, static_cast<const uint16_t*>(xhT_y) + dc0 * BWD_DCH * npad, dpad - dc0 * BWD_DCH,
                                npad, npad, B_BK, piece_w, fmt_bf16);
        if (rc) return rc;
        const uint32_t idesc_g = make_idesc_f16(BWD_BM, piece_w, fmt_bf16 ? 1u : 0u);
        ProfScope prof(PROF_LOSS_BWD_TC, s);
        loss_bwd_tc_kernel<<<grid, B_THREADS, B_SMEM_ALLOC, s>>>(
            tm_x, tm_y, tm_yt, N, d, d - dc0 * BWD_DCH, row0, n, static_cast<int>(dpad / B_BK), chunk_w, npieces,
            tiles_per_split, scale, idesc_s, idesc_g, fmt_bf16, rowcoef, colcoef, gscale, weight, accumulate,
            dxh + dc0 * BWD_DCH, scale_dev_ptr());
        CLIBD_KERNEL_CHECK();
    }
    return 0;
}

}  // namespace clibd
