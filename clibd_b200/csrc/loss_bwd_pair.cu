// Backward sweep of the contrastive loss on a CTA PAIR (thread-block cluster of 2, tcgen05 cta_group::2).
//
// One pair owns 128 rows of dXhat (64 per CTA) over the whole feature dimension (d <= 768) and sweeps the
// columns of S = Xhat Yhat^T in steps of 256:
//
//   S tile      : tcgen05.mma.cta_group::2, M = 128 (64 rows per CTA), N = 256, K = d.
//                 A = this CTA's 64 rows of Xhat, RESIDENT in shared memory for the whole sweep;
//                 B = 256 rows of Yhat, each CTA streams 128 of them (the tensor cores share them).
//   epilogue    : G~ = exp(S - s) * (rowcoef_i + colcoef_j) -> 16 bit -> this CTA's own shared memory
//   gradient    : dXhat[128 x d] += G~[128 x 256] * Yhat[256 x d], cta_group::2, M = 128, N = 256 pieces;
//                 A = this CTA's 64 rows of G~, B = rows of YhatT (each CTA streams half of every piece).
//
// Why a pair: with M = 128 over two CTAs the accumulator of a CTA is 64 rows x N columns, which TMEM stores
// as 128 lanes x N/2 columns (columns [N/2, N) live in lanes 64..127).  A 64 x 768 fp32 accumulator therefore
// needs 384 TMEM columns, leaving 128 for the 64 x 256 S tile: the whole feature dimension fits and S is
// computed ONCE per sweep (the single-CTA kernel in loss_tc.cu needs two 384-wide chunks, i.e. computes S twice).
// The G~ tile never crosses CTAs: each CTA feeds its own rows as the A operand.
//
// Two additions to the epilogue (8 warps, two per scheduler, 64 S columns per thread):
//   * positives: the column operand is class-sorted, so row i's positives are the columns [pos_lo_i, pos_lo_i +
//     cnt_i); there G~ - lam2_i is formed in fp32 BEFORE the 16-bit rounding (trained regime: G~ -> 2 T);
//   * store_g (single-GPU backward, loss_api.cu backward_shared_s): the G~ tile is also TMA-stored to a strip
//     buffer by the otherwise idle warp 3, and the other side's gradient becomes a plain GEMM over that strip
//     (loss_grad_gemm.cu) instead of a second sweep that recomputes S^T.  Measured at N = 32768: the sweep goes
//     from 2.65 to ~3.05 ms (the store itself: with the same synchronisation but no store it stays at 2.65, with
//     the stores aimed at an L2-resident dummy region at 2.86), the GEMM costs 1.17 ms, a second sweep 2.65 ms.
//     Writing the strip with 16-byte stores from the epilogue's registers instead (no second read of the tile from
//     shared memory, no store -> wait -> release chain through warp 3) was measured and is slower: 18.55 against
//     17.21 ms per step, in-process A/B (profiles/r2o2_ab_strip_store_direct_vs_tma.log) -- the epilogue is the
//     critical path of a tile, and eight more LSU instructions per thread cost more than the TMA's 32 KB of shared-
//     memory reads.
//
// Measured constraints that shaped the code (ncu + clock64 instrumentation, see profiles/):
//  * shared-memory bandwidth is the ceiling: per 64-cycle MMA a CTA's tensor core reads 2 KB of A and 4 KB of
//    B (96 B/cycle) while TMA writes the next operands (64 B/cycle); nothing else may use shared memory in the
//    hot loops (column coefficients travel through registers + shuffles, not LDS);
//  * the instruction cache is shared by three roles running different loops: every loop is kept rolled and
//    small (fully unrolled variants ran into stall_no_inst on 60 % of the epilogue samples);
//  * remote mbarrier arrives use the default (CTA-scope release) form: .release.cluster costs ~2000 cycles;
//  * all CTA pairs of a wave sweep the same columns in lockstep, so every CTA prefetches its own future
//    TMA boxes into L2 two tiles ahead.
//
// Warp roles per CTA: 0 TMA producer, 1 MMA issuer (leader CTA only), 2 TMEM allocator, 3 coefficient store,
// 4-11 epilogue.
#include <cstdlib>

#include "common.cuh"
#include "loss_plan.h"
#include "ptx.cuh"
#include "tmap.h"

namespace clibd {
#ifdef CLIBD_BWD_TIMING
__device__ unsigned long long g_pair_timing[1024 * 16];
#define TWAIT(slot, stmt)                \
    do {                                 \
        const long long _t0 = clock64(); \
        stmt;                            \
        tacc[slot] += clock64() - _t0;   \
    } while (0)
#define TMARK() tmark = clock64()
#define TLAP(slot)                      \
    do {                                \
        const long long _n = clock64(); \
        tacc[slot] += _n - tmark;       \
        tmark = _n;                     \
    } while (0)
#else
#define TWAIT(slot, stmt) stmt
#define TMARK()
#define TLAP(slot)
#endif
namespace {

using namespace ptx;

constexpr float kLog2e = 1.4426950408889634f;

constexpr int P_BK = 64;                          // K block (128 B of 16-bit operands)
constexpr int P_XKB_BYTES = 64 * 128;             // 8 KB : 64 rows x one K block
constexpr int P_STAGE_BYTES = 128 * 128;          // 16 KB: up to 128 rows x one K block = 4 MMAs of 64 cycles
constexpr int P_STAGES = 6;                       // 96 KB of streamed operands in flight
constexpr int P_MAX_KB = PAIR_DCH / P_BK;         // 12
constexpr int P_THREADS = 384;                      // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 epilogue
constexpr uint32_t P_TMEM_S_COL = 384;
constexpr int P_PF_DIST = 2;                      // L2 prefetch distance in column tiles

constexpr int P_SMEM_X = 0;                                         // resident Xhat rows: 96 KB
constexpr int P_SMEM_G = P_SMEM_X + P_MAX_KB * P_XKB_BYTES;         // G~: 4 K blocks x 8 KB
constexpr int P_SMEM_RING = P_SMEM_G + (PAIR_BJ / P_BK) * P_XKB_BYTES;
constexpr int P_SMEM_BARS = P_SMEM_RING + P_STAGES * P_STAGE_BYTES;
constexpr int P_NUM_BARS = 2 * P_STAGES + 8;
constexpr int P_SMEM_TMEMPTR = P_SMEM_BARS + P_NUM_BARS * 8;
constexpr int P_SMEM_TOTAL = P_SMEM_TMEMPTR + 16;
constexpr int P_SMEM_ALLOC = P_SMEM_TOTAL;  // no alignment slack: the base is declared 1024-aligned and checked
static_assert(P_SMEM_ALLOC <= 232448, "pair backward kernel shared memory exceeds 227 KB");

template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    if constexpr (BF16) {
        __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&v);
    } else {
        __half2 v = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&v);
    }
}

template <bool BF16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
loss_bwd_pair_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                     const __grid_constant__ CUtensorMap tm_yt, int64_t N, int64_t ld, int64_t dvalid, int64_t row0,
                     int64_t n, int num_kb, int npieces, int piece_w, int64_t tiles_per_split, float scale,
                     uint32_t idesc_s, uint32_t idesc_g, const float* __restrict__ rowcoef,
                     const float* __restrict__ colcoef, const float* __restrict__ gscale, float weight, int accumulate,
                     float* __restrict__ dxh, int self_mask, const int32_t* __restrict__ pos_lo,
                     const float* __restrict__ pos_cnt, const float* __restrict__ lam2, int64_t jt_lo, int64_t jt_hi,
                     const __grid_constant__ CUtensorMap tm_gs, int store_g, const float* __restrict__ scale_dev) {
    scale = eff_scale(scale, scale_dev);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B operand tiles need 1024-byte alignment
    uint8_t* xres = smem + P_SMEM_X;
    uint8_t* gbuf = smem + P_SMEM_G;
    uint8_t* ring = smem + P_SMEM_RING;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_SMEM_BARS);
    uint64_t* full = bars;                      // leader: TMA bytes of both CTAs landed
    uint64_t* empty = bars + P_STAGES;          // every CTA: the MMAs reading this stage have completed
    uint64_t* xfull = bars + 2 * P_STAGES;      // leader: resident Xhat rows of both CTAs landed
    uint64_t* st_full = xfull + 1;              // every CTA: S tile ready in TMEM
    uint64_t* st_empty = xfull + 2;             // leader: S tile drained by both epilogues (16 warps)
    uint64_t* g_full = xfull + 3;               // leader: G~ written by both epilogues (16 warps)
    uint64_t* g_empty = xfull + 4;              // every CTA: G~ consumed by the gradient MMAs
    uint64_t* acc_full = xfull + 5;             // every CTA: all accumulation finished
    uint64_t* g_lfull = xfull + 6;              // local: G~ written by this CTA's 8 epilogue warps (store_g only)
    uint64_t* g_sdone = xfull + 7;              // local: the TMA store of the G~ tile has finished reading it
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + P_SMEM_TMEMPTR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
#ifdef CLIBD_BWD_TIMING
    long long tacc[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long tmark = 0;
    const long long t_begin = clock64();
#endif
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int64_t mt = blockIdx.x >> 1;
    const int64_t split = blockIdx.z;
    // this launch sweeps the column tiles [jt_lo, jt_hi) (a strip of the columns, or all of them)
    const int64_t jt0 = jt_lo + split * tiles_per_split;
    const int64_t jt1 = (jt0 + tiles_per_split < jt_hi) ? (jt0 + tiles_per_split) : jt_hi;
    const int half_w = piece_w >> 1;  // accumulator columns per piece per CTA

    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_x);
        prefetch_tmap(&tm_y);
        prefetch_tmap(&tm_yt);
        for (int i = 0; i < P_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        mbar_init(xfull, 1);
        mbar_init(st_full, 1);
        mbar_init(st_empty, 16);
        mbar_init(g_full, 16);
        mbar_init(g_empty, 1);
        mbar_init(acc_full, 1);
        mbar_init(g_lfull, 8);
        mbar_init(g_sdone, 1);
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc_cg2(tmem_ptr, 512);
        tmem_relinquish_cg2();
    }
    tc_fence_before();
    cluster_sync_all();  // barrier inits and TMEM allocations of both CTAs are visible cluster-wide
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (jt0 < jt1) {
        if (warp == 0) {  // ---------------- TMA producer (both CTAs; one elected lane issues)
            int slot = 0;
            uint32_t phase = 0;
            const int32_t xrow = static_cast<int32_t>(row0 + mt * PAIR_BM + rank * 64);
            const uint32_t xfull_l = mapa_u32(smem_u32(xfull), 0);
            const uint32_t full_l0 = mapa_u32(smem_u32(&full[0]), 0);  // leader's full[0]; full[s] is 8*s further
            const int32_t ytrow = static_cast<int32_t>(rank * half_w);
            const uint32_t gbytes = 2u * half_w * 128;  // both CTAs' YhatT sub-tiles of one stage
            const bool elected = elect_one();
            if (elected) {
                if (leader) mbar_arrive_expect_tx(xfull, 2u * num_kb * P_XKB_BYTES);
                for (int kb = 0; kb < num_kb; ++kb)
                    tma_load_2d_cg2(&tm_x, xfull_l, xres + kb * P_XKB_BYTES, kb * P_BK, xrow, kEvictNormal);
                for (int64_t t = jt0; t < jt1 && t < jt0 + P_PF_DIST; ++t) {
                    for (int kb = 0; kb < num_kb; ++kb)
                        tma_prefetch_2d(&tm_y, kb * P_BK, static_cast<int32_t>(t * PAIR_BJ + rank * 128));
                    for (int kb2 = 0; kb2 < PAIR_BJ / P_BK; ++kb2)
                        for (int pc = 0; pc < npieces; ++pc)
                            tma_prefetch_2d(&tm_yt, static_cast<int32_t>(t * PAIR_BJ + kb2 * P_BK), pc * piece_w + ytrow);
                }
            }
            __syncwarp();
            auto load_grad = [&](int64_t t) {
                const int32_t jcol = static_cast<int32_t>(t * PAIR_BJ);
                const bool pf = t + P_PF_DIST < jt1;
#pragma unroll 1
                for (int kb2 = 0; kb2 < PAIR_BJ / P_BK; ++kb2) {
#pragma unroll 1
                    for (int pc = 0; pc < npieces; ++pc) {
                        TWAIT(1, mbar_wait(&empty[slot], phase ^ 1));
                        if (elected) {
                            if (leader) mbar_arrive_expect_tx(&full[slot], gbytes);
                            tma_load_2d_cg2(&tm_yt, full_l0 + slot * 8, ring + slot * P_STAGE_BYTES, jcol + kb2 * P_BK,
                                            pc * piece_w + ytrow, kEvictNormal);
                            if (pf) tma_prefetch_2d(&tm_yt, jcol + P_PF_DIST * PAIR_BJ + kb2 * P_BK, pc * piece_w + ytrow);
                        }
                        __syncwarp();
                        if (++slot == P_STAGES) {
                            slot = 0;
                            phase ^= 1;
                        }
                    }
                }
            };
            for (int64_t t = jt0; t < jt1; ++t) {
                const int32_t yrow = static_cast<int32_t>(t * PAIR_BJ + rank * 128);
                const bool pf = t + P_PF_DIST < jt1;
#pragma unroll 1
                for (int kb = 0; kb < num_kb; ++kb) {
                    TWAIT(0, mbar_wait(&empty[slot], phase ^ 1));
                    if (elected) {
                        if (leader) mbar_arrive_expect_tx(&full[slot], 2u * P_STAGE_BYTES);
                        tma_load_2d_cg2(&tm_y, full_l0 + slot * 8, ring + slot * P_STAGE_BYTES, kb * P_BK, yrow,
                                        kEvictNormal);
                        if (pf) tma_prefetch_2d(&tm_y, kb * P_BK, yrow + P_PF_DIST * PAIR_BJ);
                    }
                    __syncwarp();
                    if (++slot == P_STAGES) {
                        slot = 0;
                        phase ^= 1;
                    }
                }
                if (t > jt0) load_grad(t - 1);
            }
            load_grad(jt1 - 1);
        } else if (warp == 1 && leader) {  // ---------------- MMA issuer (leader CTA)
            int slot = 0;
            uint32_t phase = 0;
            const uint32_t s_tmem = tmem_base + P_TMEM_S_COL;
            // descriptors of offset 0 of every operand region; tiles inside a region are reached by adding
            // (byte offset >> 4) to the low word (all regions lie below 256 KB: no carry into other fields)
            const uint64_t dx0 = make_sw128_kmajor_desc(smem_u32(xres));
            const uint64_t dg0 = make_sw128_kmajor_desc(smem_u32(gbuf));
            const uint64_t dr0 = make_sw128_kmajor_desc(smem_u32(ring));
            const bool elected = elect_one();
            auto issue_grad = [&](int64_t tl) {  // tl = t - jt0 of the G~ tile to consume
                TWAIT(2, mbar_wait(g_full, tl & 1));
                tc_fence_after();
                const uint32_t acc_first = tl > 0 ? 1u : 0u;
                uint64_t da = dg0;
#pragma unroll 1
                for (int kb2 = 0; kb2 < PAIR_BJ / P_BK; ++kb2) {
#pragma unroll 1
                    for (int pc = 0; pc < npieces; ++pc) {
                        TWAIT(3, mbar_wait(&full[slot], phase));
                        tc_fence_after();
                        if (elected) {
                            const uint64_t db = dr0 + slot * (P_STAGE_BYTES >> 4);
                            const uint32_t dcol = tmem_base + pc * half_w;
                            umma_f16_cg2(dcol, da, db, idesc_g, kb2 > 0 ? 1u : acc_first);
                            umma_f16_cg2(dcol, da + 2, db + 2, idesc_g, 1u);
                            umma_f16_cg2(dcol, da + 4, db + 4, idesc_g, 1u);
                            umma_f16_cg2(dcol, da + 6, db + 6, idesc_g, 1u);
                            umma_commit_cg2(&empty[slot], 3);
                            if (kb2 == PAIR_BJ / P_BK - 1 && pc == npieces - 1) umma_commit_cg2(g_empty, 3);
                        }
                        __syncwarp();
                        if (++slot == P_STAGES) {
                            slot = 0;
                            phase ^= 1;
                        }
                    }
                    da += P_XKB_BYTES >> 4;
                }
            };
            mbar_wait(xfull, 0);
            tc_fence_after();
            for (int64_t t = jt0; t < jt1; ++t) {
                const int64_t tl = t - jt0;
                TWAIT(4, mbar_wait(st_empty, (tl & 1) ^ 1));
                tc_fence_after();
                uint64_t da = dx0;
#pragma unroll 1
                for (int kb = 0; kb < num_kb; ++kb) {
                    TWAIT(5, mbar_wait(&full[slot], phase));
                    tc_fence_after();
                    if (elected) {
                        const uint64_t db = dr0 + slot * (P_STAGE_BYTES >> 4);
                        umma_f16_cg2(s_tmem, da, db, idesc_s, kb > 0 ? 1u : 0u);
                        umma_f16_cg2(s_tmem, da + 2, db + 2, idesc_s, 1u);
                        umma_f16_cg2(s_tmem, da + 4, db + 4, idesc_s, 1u);
                        umma_f16_cg2(s_tmem, da + 6, db + 6, idesc_s, 1u);
                        umma_commit_cg2(&empty[slot], 3);
                        if (kb == num_kb - 1) umma_commit_cg2(st_full, 3);
                    }
                    __syncwarp();
                    da += P_XKB_BYTES >> 4;
                    if (++slot == P_STAGES) {
                        slot = 0;
                        phase ^= 1;
                    }
                }
                if (tl > 0) issue_grad(tl - 1);
            }
            issue_grad(jt1 - jt0 - 1);
            if (elected) umma_commit_cg2(acc_full, 3);
            __syncwarp();
        } else if (warp == 3 && store_g) {  // ---------------- coefficient store (both CTAs)
            // The G~ tile the epilogue wrote for the gradient MMAs (4 K blocks of 64 rows x 128 B, SWIZZLE_128B) is
            // also sent to global memory as it is -- rows = this CTA's 64 rows, columns = the tile's 256 columns of
            // the strip -- for the other side's gradient GEMM (loss_grad_gemm.cu reads it as an M-major operand).
            // No registers, no LSU instructions: one TMA store per K block.
            const bool elected = elect_one();
            const int32_t grow = static_cast<int32_t>(mt * PAIR_BM + rank * 64);  // LOCAL row: the strip has n rows
            for (int64_t t = jt0; t < jt1; ++t) {
                const int64_t tl = t - jt0;
                mbar_wait(g_lfull, tl & 1);
                if (elected) {
                    const int32_t gcol = static_cast<int32_t>((t - jt_lo) * PAIR_BJ);
#pragma unroll
                    for (int kb2 = 0; kb2 < PAIR_BJ / P_BK; ++kb2)
                        tma_store_2d(&tm_gs, gbuf + kb2 * P_XKB_BYTES, gcol + kb2 * P_BK, grow, kEvictFirst);
                    bulk_commit_group();
                    bulk_wait_group_read0();
                    mbar_arrive(g_sdone);
                }
                __syncwarp();
            }
            if (elected) bulk_wait_group0();  // the stores are complete before the kernel ends
            __syncwarp();
        } else if (warp >= 4) {  // ---------------- epilogue (both CTAs), 8 warps: two per scheduler
            // TMEM lane quadrant q = warp & 3 (hardware rule); the two warps of a quadrant split the 128 S columns a
            // lane holds into halves ch = 0 / 1 of 64.  One epilogue warp per scheduler left the ~1400 dependent
            // instructions per tile exposed (measured: +32 % instructions = +32 % kernel time); two interleave.
            const int q = warp & 3;
            const int ch = (warp - 4) >> 2;
            const int tl_lane = q * 32 + lane;          // TMEM lane of this thread
            const int rloc = tl_lane & 63;              // row inside this CTA's 64-row slab
            const int h = tl_lane >> 6;                 // which half of the tile / piece columns this lane holds
            const int cw = h * 128 + ch * 64;           // first tile column of this thread's 64-column window
            const int64_t lrow = mt * PAIR_BM + rank * 64 + rloc;
            const float gs = gscale[0];
            const float rcg = (lrow < n ? rowcoef[row0 + lrow] : 0.f) * gs;
            const float a = scale * kLog2e;
            const float nb = -softmax_shift(scale) * kLog2e;
            const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
            const uint32_t st_empty_l = mapa_u32(smem_u32(st_empty), 0);
            const uint32_t g_full_l = mapa_u32(smem_u32(g_full), 0);
            const uint32_t rowaddr = smem_u32(gbuf) + (2 * h + ch) * P_XKB_BYTES + rloc * 128;
            // positives of this row: columns [plo, plo + plen) (the column operand is class-sorted); there the
            // epilogue emits G~ - lam2 so that the 16-bit rounding acts on the small difference when G~ -> 2 T
            // (the fp32 class-sum term of normalize_bwd carries the remaining 2 - lam2)
            int plo = 0, plen = 0;
            float l2g = 0.f;
            if (lam2 != nullptr && lrow < n) {
                plo = pos_lo[row0 + lrow];
                plen = static_cast<int>(pos_cnt[row0 + lrow]);
                l2g = lam2[lrow] * gs;
            }
            // Column coefficients: lane l keeps those of columns l and 32+l of its window in registers (fetched one
            // tile ahead) and the warp broadcasts them with shuffles -- no shared memory.
            auto load_cc = [&](int64_t t, float (&dst)[2]) {
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const int64_t gj = t * PAIR_BJ + cw + m * 32 + lane;
                    dst[m] = (gj < N) ? colcoef[gj] * gs : 0.f;
                }
            };
            float ccr[2], ccn[2] = {0.f, 0.f};
            load_cc(jt0, ccr);
            for (int64_t t = jt0; t < jt1; ++t) {
                const int64_t tl = t - jt0;
                const int64_t nvalid = N - t * PAIR_BJ - cw;  // columns of this thread's window that exist
                if (t + 1 < jt1) load_cc(t + 1, ccn);
                TWAIT(6, mbar_wait(st_full, tl & 1));
                tc_fence_after();
                TMARK();
                // pull this thread's S window into registers at once and hand the TMEM columns straight back to
                // the MMA warp: S(t+1) can then start as soon as the gradient MMAs of tile t-1 retire
                uint32_t v[64];
                tmem_ld_32x32b_x32(tmem_base + lane_base + P_TMEM_S_COL + ch * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                tmem_ld_32x32b_x32(tmem_base + lane_base + P_TMEM_S_COL + ch * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(st_empty_l);
                TLAP(10);
                uint32_t packed[32];
                // first column of this thread's window relative to the row's positive range
                const int prel = static_cast<int>(t * PAIR_BJ + cw) - plo;
                if (__any_sync(0xffffffffu, prel > -64 && prel < plen)) {  // rare: some row of the warp has positives here
#pragma unroll
                    for (int k = 0; k < 64; k += 2) {
                        const float c0 = __shfl_sync(0xffffffffu, ccr[k >> 5], k & 31);
                        const float c1 = __shfl_sync(0xffffffffu, ccr[k >> 5], (k & 31) + 1);
                        const float e0 = ex2_approx(fmaf(__uint_as_float(v[k]), a, nb));
                        const float e1 = ex2_approx(fmaf(__uint_as_float(v[k + 1]), a, nb));
                        const float s0 = static_cast<unsigned>(prel + k) < static_cast<unsigned>(plen) ? l2g : 0.f;
                        const float s1 = static_cast<unsigned>(prel + k + 1) < static_cast<unsigned>(plen) ? l2g : 0.f;
                        packed[k / 2] = pack2<BF16>(fmaf(e0, rcg + c0, -s0), fmaf(e1, rcg + c1, -s1));
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 64; k += 2) {
                        const float c0 = __shfl_sync(0xffffffffu, ccr[k >> 5], k & 31);
                        const float c1 = __shfl_sync(0xffffffffu, ccr[k >> 5], (k & 31) + 1);
                        const float e0 = ex2_approx(fmaf(__uint_as_float(v[k]), a, nb));
                        const float e1 = ex2_approx(fmaf(__uint_as_float(v[k + 1]), a, nb));
                        packed[k / 2] = pack2<BF16>(e0 * (rcg + c0), e1 * (rcg + c1));
                    }
                }
                if (nvalid < 64) {  // ragged last tile (rare): columns that do not exist contribute nothing
#pragma unroll
                    for (int p = 0; p < 32; ++p) {
                        if (2 * p >= nvalid) packed[p] = 0u;
                        else if (2 * p + 1 >= nvalid) packed[p] &= 0xFFFFu;
                    }
                }
                if (self_mask) {  // info-NCE on one feature set: the entry (row, row) does not exist
                    const int64_t kd = row0 + lrow - (t * PAIR_BJ + cw);
                    if (kd >= 0 && kd < 64) {
#pragma unroll
                        for (int p = 0; p < 32; ++p) {
                            if (2 * p == kd) packed[p] &= 0xFFFF0000u;
                            else if (2 * p + 1 == kd) packed[p] &= 0xFFFFu;
                        }
                    }
                }
                TLAP(11);
                // the previous G~ tile must have been consumed before it is overwritten
                TWAIT(7, mbar_wait(g_empty, (tl & 1) ^ 1));
                if (store_g) mbar_wait(g_sdone, (tl & 1) ^ 1);
                TMARK();
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    const uint32_t addr = rowaddr + ((c8 ^ (rloc & 7)) << 4);
                    const int p = c8 * 4;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[p]),
                                 "r"(packed[p + 1]), "r"(packed[p + 2]), "r"(packed[p + 3])
                                 : "memory");
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_cluster(g_full_l);
                    if (store_g) mbar_arrive(g_lfull);
                }
                TLAP(12);
#pragma unroll
                for (int m = 0; m < 2; ++m) ccr[m] = ccn[m];
            }
            // drain the accumulators: dxh (+)= weight / gscale * acc (the two warps of a quadrant alternate chunks)
            TWAIT(8, mbar_wait(acc_full, 0));
            tc_fence_after();
            const float wgt = weight * gscale[1];
            float* out = dxh + (split * n + lrow) * ld;
            const bool vec_ok = (ld & 3) == 0;  // rows of dxh are then 16-byte aligned
#pragma unroll 1
            for (int pc = 0; pc < npieces; ++pc) {
#pragma unroll 1
                for (int c0 = ch * 32; c0 < half_w; c0 += 64) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_base + lane_base + pc * half_w + c0, v);
                    tmem_ld_wait();
                    if (lrow < n) {
                        const int colbase = pc * piece_w + h * half_w + c0;
                        if (vec_ok && c0 + 32 <= half_w && colbase + 32 <= dvalid) {
                            float4* o4 = reinterpret_cast<float4*>(out + colbase);
                            float4 old[8];
                            if (accumulate) {
#pragma unroll
                                for (int i = 0; i < 8; ++i) old[i] = o4[i];
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float4 r;
                                r.x = wgt * __uint_as_float(v[4 * i]);
                                r.y = wgt * __uint_as_float(v[4 * i + 1]);
                                r.z = wgt * __uint_as_float(v[4 * i + 2]);
                                r.w = wgt * __uint_as_float(v[4 * i + 3]);
                                if (accumulate) {
                                    r.x += old[i].x;
                                    r.y += old[i].y;
                                    r.z += old[i].z;
                                    r.w += old[i].w;
                                }
                                o4[i] = r;
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 32; ++k) {
                                const int col = colbase + k;
                                if (c0 + k < half_w && col < dvalid) {
                                    const float val = wgt * __uint_as_float(v[k]);
                                    out[col] = accumulate ? out[col] + val : val;
                                }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp >= 4) {
        // empty column range for this split: contribute zeros unless accumulating
        const int etid = (warp - 4) * 32 + lane;
        if (etid < 64) {
            const int64_t lrow = mt * PAIR_BM + rank * 64 + etid;
            if (!accumulate && lrow < n) {
                float* out = dxh + (split * n + lrow) * ld;
                for (int64_t col = 0; col < dvalid; ++col) out[col] = 0.f;
            }
        }
    }
#ifdef CLIBD_BWD_TIMING
    if (lane == 0 && blockIdx.x < 1024 && blockIdx.z == 0) {
        tacc[9] = clock64() - t_begin;
        unsigned long long* o = g_pair_timing + blockIdx.x * 16;
        if (warp == 0) { o[0] = tacc[0]; o[1] = tacc[1]; o[9] = tacc[9]; }
        if (warp == 1) { o[2] = tacc[2]; o[3] = tacc[3]; o[4] = tacc[4]; o[5] = tacc[5]; o[10] = tacc[9]; }
        if (warp == 4) { o[6] = tacc[6]; o[7] = tacc[7]; o[8] = tacc[8]; o[11] = tacc[9]; o[12] = tacc[10]; o[13] = tacc[11]; o[14] = tacc[12]; o[15] = tacc[13]; }
    }
#endif
    tc_fence_before();
    cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still signal it or use its operands
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, 512);
    }
}

}  // namespace

#ifdef CLIBD_BWD_TIMING
extern "C" int clibd_debug_pair_timing(unsigned long long* host_out, int n) {
    return cudaMemcpyFromSymbol(host_out, g_pair_timing, sizeof(unsigned long long) * n) == cudaSuccess ? 0 : 2;
}
#endif

bool pair_backward_supported(int64_t dpad) { return dpad <= PAIR_DCH; }

int tc_backward_rows_pair(const void* xh_x, const void* xh_y, const void* xhT_y, int64_t N, int64_t npad, int64_t d,
                          int64_t dpad, int64_t row0, int64_t n, float scale, const float* rowcoef,
                          const float* colcoef, const float* gscale, float weight, int accumulate, int jsplit,
                          int fmt_bf16, float* dxh, cudaStream_t s, int self_mask, const int32_t* pos_lo,
                          const float* pos_cnt, const float* lam2, int64_t col_begin, int64_t col_end, void* gt,
                          int64_t gt_ld) {
    if (n == 0 || N == 0) return 0;
    CLIBD_REQUIRE(dpad % P_BK == 0 && dpad <= PAIR_DCH, "pair backward needs a padded feature dim <= 768");
    CLIBD_REQUIRE(lam2 == nullptr || (pos_lo != nullptr && pos_cnt != nullptr), "lam2 needs the positive ranges");
    CLIBD_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&loss_bwd_pair_kernel<true>), P_SMEM_ALLOC));
    CLIBD_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&loss_bwd_pair_kernel<false>), P_SMEM_ALLOC));
    // gradient pieces: npieces equal pieces of width piece_w <= 256, piece_w a multiple of 16 (8 rows of YhatT per
    // swizzle atom and per CTA)
    int npieces = 1;
    while (dpad % npieces != 0 || dpad / npieces > 256 || (dpad / npieces) % 16 != 0) ++npieces;
    const int piece_w = static_cast<int>(dpad / npieces);
    CUtensorMap tm_x, tm_y, tm_yt, tm_gs;
    // the row operand is staged for the local rows only: rows past row0 + n read as zeros (TMA fill)
    int rc = make_tmap_2d_16bit(&tm_x, xh_x, row0 + n, dpad, dpad, P_BK, 64, fmt_bf16);
    if (rc) return rc;
    rc = make_tmap_2d_16bit(&tm_y, xh_y, N, dpad, dpad, P_BK, 128, fmt_bf16);
    if (rc) return rc;
    rc = make_tmap_2d_16bit(&tm_yt, xhT_y, dpad, npad, npad, P_BK, piece_w / 2, fmt_bf16);
    if (rc) return rc;
    if (col_end < 0 || col_end > N) col_end = N;
    CLIBD_REQUIRE(col_begin >= 0 && col_begin % PAIR_BJ == 0 && col_begin < col_end, "bad column strip");
    const int64_t jt_lo = col_begin / PAIR_BJ, jt_hi = ceil_div(col_end, PAIR_BJ);
    const int64_t num_jt = jt_hi - jt_lo;
    if (gt != nullptr) {  // coefficient strip [n local rows, columns of this strip], pitch gt_ld
        rc = make_tmap_2d_16bit(&tm_gs, gt, n, col_end - col_begin, gt_ld, P_BK, 64, fmt_bf16);
        if (rc) return rc;
    } else {
        tm_gs = tm_x;  // unused
    }
    const int64_t tiles_per_split = ceil_div(num_jt, jsplit);
    const uint32_t idesc_s = make_idesc_f16(PAIR_BM, PAIR_BJ, fmt_bf16 ? 1u : 0u);
    const uint32_t idesc_g = make_idesc_f16(PAIR_BM, piece_w, fmt_bf16 ? 1u : 0u);
    dim3 grid(static_cast<unsigned>(2 * ceil_div(n, PAIR_BM)), 1, static_cast<unsigned>(jsplit));
    ProfScope prof(PROF_LOSS_BWD_TC, s);
    auto kern = fmt_bf16 ? loss_bwd_pair_kernel<true> : loss_bwd_pair_kernel<false>;
    kern<<<grid, P_THREADS, P_SMEM_ALLOC, s>>>(tm_x, tm_y, tm_yt, N, d, d, row0, n, static_cast<int>(dpad / P_BK), npieces,
                                               piece_w, tiles_per_split, scale, idesc_s, idesc_g, rowcoef, colcoef,
                                               gscale, weight, accumulate, dxh, self_mask, pos_lo, pos_cnt,
                                               lam2, jt_lo, jt_hi, tm_gs, gt != nullptr ? 1 : 0, scale_dev_ptr());
    CLIBD_KERNEL_CHECK();
    return 0;
}

}  // namespace clibd
