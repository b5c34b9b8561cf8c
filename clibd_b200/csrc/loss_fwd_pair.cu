// Forward statistics of the contrastive loss on CTA PAIRS (cluster of 2, tcgen05 cta_group::2, M = 256).
//
// A pair walks 256 x 256 tiles of S = Xhat Yhat^T; each CTA holds 128 rows of the tile in TMEM (2 x 256 columns,
// double buffered) and streams 128 rows of Xhat and 128 rows of Yhat per 64-wide K block -- the tensor cores of
// the two SMs share the Yhat halves.  Per 128-cycle MMA a CTA reads 4 KB of A and 4 KB of B from shared memory
// and TMA writes 8 KB: 128 B/cycle in total, exactly the shared-memory bandwidth, where the single-CTA 128 x 256
// kernel (loss_tc.cu) needs 192 B/cycle and stalls at ~68 % of the tensor peak (measured).
//
// Epilogue (8 warps per CTA, same as loss_tc.cu): e = exp2(s*log2e*(cos-1)), row sums in registers, column sums
// by a transposing shuffle butterfly, per-tile partials to HBM.  S never leaves TMEM.
//
// Warp roles per CTA: 0 TMA producer, 1 MMA issuer (leader CTA only), 2 TMEM allocator, 4-11 epilogue.
#include "common.cuh"
#include "loss_plan.h"
#include "ptx.cuh"
#include "tmap.h"

namespace clibd {
namespace {

using namespace ptx;

constexpr float kLog2e = 1.4426950408889634f;

constexpr int Q_BK = 64;
constexpr int Q_A_BYTES = 128 * Q_BK * 2;   // 16 KB: this CTA's 128 rows of Xhat, one K block
constexpr int Q_B_BYTES = 128 * Q_BK * 2;   // 16 KB: this CTA's 128 rows of Yhat (half of the tile's columns)
constexpr int Q_STAGE_BYTES = Q_A_BYTES + Q_B_BYTES;
constexpr int Q_STAGES = 6;
constexpr int Q_THREADS = 384;
constexpr int Q_EPI_WARPS = 8;
constexpr int Q_TN = 256;                   // tile columns
constexpr int Q_SMEM_COLBUF = Q_STAGES * Q_STAGE_BYTES;            // float [2][4][256]
constexpr int Q_SMEM_ROWBUF = Q_SMEM_COLBUF + 2 * 4 * Q_TN * 4;    // float [2][2][128]
constexpr int Q_SMEM_BARS = Q_SMEM_ROWBUF + 2 * 2 * 128 * 4;
constexpr int Q_NUM_BARS = 2 * Q_STAGES + 4;
constexpr int Q_SMEM_TMEMPTR = Q_SMEM_BARS + Q_NUM_BARS * 8;
constexpr int Q_SMEM_TOTAL = Q_SMEM_TMEMPTR + 16;
static_assert(Q_SMEM_TOTAL <= 232448, "pair forward kernel shared memory exceeds 227 KB");

__device__ __forceinline__ void pair_tile_coords(int64_t t, int64_t num_mt, int64_t num_nt, int64_t& mt, int64_t& nt) {
    constexpr int64_t GM = 8;  // 8 row tiles share a column tile back to back: operands stay hot in L2
    const int64_t group = GM * num_nt;
    const int64_t g = t / group;
    const int64_t first = g * GM;
    const int64_t gm = (num_mt - first) < GM ? (num_mt - first) : GM;
    const int64_t r = t % group;
    mt = first + r % gm;
    nt = r / gm;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Q_THREADS, 1)
loss_fwd_pair_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int64_t N,
                     int64_t row0, int64_t n, int num_kb, float scale, uint32_t idesc, float* __restrict__ rowpart,
                     float* __restrict__ colpart, int self_mask, const float* __restrict__ scale_dev) {
    scale = eff_scale(scale, scale_dev);
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    float* colbuf = reinterpret_cast<float*>(smem + Q_SMEM_COLBUF);
    float* rowbuf = reinterpret_cast<float*>(smem + Q_SMEM_ROWBUF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Q_SMEM_BARS);
    uint64_t* full = bars;                        // leader: both CTAs' operands of the stage landed
    uint64_t* empty = bars + Q_STAGES;            // every CTA: MMAs reading the stage completed
    uint64_t* tfull = bars + 2 * Q_STAGES;        // every CTA [2]: accumulator buffer ready
    uint64_t* tempty = bars + 2 * Q_STAGES + 2;   // leader [2]: accumulator drained by both epilogues (16 warps)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Q_SMEM_TMEMPTR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int64_t num_mt = (n + 255) / 256;
    const int64_t num_nt = (N + Q_TN - 1) / Q_TN;
    const int64_t num_tiles = num_mt * num_nt;
    const int64_t pair = blockIdx.x >> 1;
    const int64_t num_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_a);
        prefetch_tmap(&tm_b);
        for (int i = 0; i < Q_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 2 * Q_EPI_WARPS);
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

    if (warp == 0) {  // ---------------- TMA producer (both CTAs)
        int slot = 0;
        uint32_t phase = 0;
        const uint32_t full_l0 = mapa_u32(smem_u32(&full[0]), 0);
        const bool elected = elect_one();
        for (int64_t t = pair; t < num_tiles; t += num_pairs) {
            int64_t mt, nt;
            pair_tile_coords(t, num_mt, num_nt, mt, nt);
            const int32_t arow = static_cast<int32_t>(row0 + mt * 256 + rank * 128);
            const int32_t brow = static_cast<int32_t>(nt * Q_TN + rank * 128);
#pragma unroll 1
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&empty[slot], phase ^ 1);
                if (elected) {
                    uint8_t* sa = smem + slot * Q_STAGE_BYTES;
                    if (leader) mbar_arrive_expect_tx(&full[slot], 2u * Q_STAGE_BYTES);
                    tma_load_2d_cg2(&tm_a, full_l0 + slot * 8, sa, kb * Q_BK, arow, kEvictNormal);
                    tma_load_2d_cg2(&tm_b, full_l0 + slot * 8, sa + Q_A_BYTES, kb * Q_BK, brow, kEvictNormal);
                }
                __syncwarp();
                if (++slot == Q_STAGES) {
                    slot = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && leader) {  // ---------------- MMA issuer (leader CTA)
        int slot = 0;
        uint32_t phase = 0;
        uint32_t it = 0;
        const uint64_t d0 = make_sw128_kmajor_desc(smem_u32(smem));
        const bool elected = elect_one();
        for (int64_t t = pair; t < num_tiles; t += num_pairs, ++it) {
            const uint32_t as = it & 1, aph = (it >> 1) & 1;
            mbar_wait(&tempty[as], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * Q_TN;
#pragma unroll 1
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[slot], phase);
                tc_fence_after();
                if (elected) {
                    const uint64_t da = d0 + slot * (Q_STAGE_BYTES >> 4);
                    const uint64_t db = da + (Q_A_BYTES >> 4);
                    umma_f16_cg2(d_tmem, da, db, idesc, kb > 0 ? 1u : 0u);
                    umma_f16_cg2(d_tmem, da + 2, db + 2, idesc, 1u);
                    umma_f16_cg2(d_tmem, da + 4, db + 4, idesc, 1u);
                    umma_f16_cg2(d_tmem, da + 6, db + 6, idesc, 1u);
                    umma_commit_cg2(&empty[slot], 3);
                    if (kb == num_kb - 1) umma_commit_cg2(&tfull[as], 3);
                }
                __syncwarp();
                if (++slot == Q_STAGES) {
                    slot = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {  // ---------------- epilogue (both CTAs)
        const int ew = warp - 4;
        const int q = warp & 3;   // TMEM lane quadrant this warp may access
        const int h = ew >> 2;    // column half of the tile
        const int etid = ew * 32 + lane;
        const float a = scale * kLog2e;
        const float nb = -softmax_shift(scale) * kLog2e;
        const uint32_t tempty_l0 = mapa_u32(smem_u32(&tempty[0]), 0);
        uint32_t it = 0;
        for (int64_t t = pair; t < num_tiles; t += num_pairs, ++it) {
            int64_t mt, nt;
            pair_tile_coords(t, num_mt, num_nt, mt, nt);
            const uint32_t as = it & 1, aph = (it >> 1) & 1;
            const int buf = it & 1;
            const int64_t rbase = mt * 256 + rank * 128;       // first local row of this CTA's slab
            const int64_t lrow = rbase + q * 32 + lane;
            const bool row_ok = lrow < n;
            const int64_t colbase = nt * Q_TN + h * 128;
            // self_mask (info-NCE on one feature set, simclr.py:76-79): entries whose global row equals their column are
            // excluded; only the tiles the diagonal crosses take the predicated branch
            const int64_t grow = row0 + lrow;
            const bool on_diag = self_mask && (row0 + rbase < colbase + 128) && (colbase < row0 + rbase + 128);
            const bool full_tile = (rbase + 128 <= n) && (nt * Q_TN + Q_TN <= N) && !on_diag;
            mbar_wait(&tfull[as], aph);
            tc_fence_after();
            float rsum = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Q_TN + h * 128 + c * 32;
                tmem_ld_32x32b_x32(taddr, v);
                tmem_ld_wait();
                if (c == 3) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tempty_l0 + as * 8);
                }
                float e[32];
                if (full_tile) {
#pragma unroll
                    for (int k = 0; k < 32; ++k) e[k] = ex2_approx(fmaf(__uint_as_float(v[k]), a, nb));
                } else {
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        const int64_t gc = colbase + c * 32 + k;
                        const bool ok = row_ok && (gc < N) && !(on_diag && gc == grow);
                        e[k] = ok ? ex2_approx(fmaf(__uint_as_float(v[k]), a, nb)) : 0.f;
                    }
                }
#pragma unroll
                for (int k = 0; k < 32; ++k) rsum += e[k];
                // transposing butterfly: afterwards lane l holds the sum over this warp's 32 rows of column c*32+l
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
                colbuf[(buf * 4 + q) * Q_TN + h * 128 + c * 32 + lane] = e[0];
            }
            rowbuf[(buf * 2 + h) * 128 + q * 32 + lane] = rsum;
            asm volatile("bar.sync 1, %0;" ::"n"(Q_EPI_WARPS * 32) : "memory");
            {
                const int j = etid;  // 0..255: one column of the tile each
                const float cs = colbuf[(buf * 4 + 0) * Q_TN + j] + colbuf[(buf * 4 + 1) * Q_TN + j] +
                                 colbuf[(buf * 4 + 2) * Q_TN + j] + colbuf[(buf * 4 + 3) * Q_TN + j];
                const int64_t gcol = nt * Q_TN + j;
                const int64_t cpart = 2 * mt + rank;  // one partial per 128-row slab
                if (gcol < N && rbase < n) colpart[cpart * N + gcol] = cs;
                if (etid < 128) {
                    const float rs = rowbuf[(buf * 2 + 0) * 128 + etid] + rowbuf[(buf * 2 + 1) * 128 + etid];
                    const int64_t lr = rbase + etid;
                    if (lr < n) rowpart[nt * n + lr] = rs;
                }
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

}  // namespace

int tc_forward_pair_cg2(const void* xh_a, const void* xh_b, int64_t N, int64_t dpad, int64_t row0, int64_t n, float scale,
                        int fmt_bf16, float* rowpart, float* colpart, int num_sms, cudaStream_t s, int self_mask) {
    if (n == 0 || N == 0) return 0;
    CLIBD_REQUIRE(dpad % Q_BK == 0, "padded feature dim must be a multiple of 64");
    CUtensorMap tm_a, tm_b;
    // the row operand is staged for the local rows only: rows past row0 + n read as zeros (TMA fill)
    int rc = make_tmap_2d_16bit(&tm_a, xh_a, row0 + n, dpad, dpad, Q_BK, 128, fmt_bf16);
    if (rc) return rc;
    rc = make_tmap_2d_16bit(&tm_b, xh_b, N, dpad, dpad, Q_BK, 128, fmt_bf16);
    if (rc) return rc;
    CLIBD_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&loss_fwd_pair_kernel), Q_SMEM_TOTAL));
    const int64_t tiles = ceil_div(n, 256) * ceil_div(N, Q_TN);
    const int64_t max_pairs = num_sms / 2;
    const int pairs = static_cast<int>(tiles < max_pairs ? tiles : max_pairs);
    const uint32_t idesc = make_idesc_f16(256, Q_TN, fmt_bf16 ? 1u : 0u);
    ProfScope prof(PROF_LOSS_FWD_TC, s);
    loss_fwd_pair_kernel<<<2 * pairs, Q_THREADS, Q_SMEM_TOTAL, s>>>(tm_a, tm_b, N, row0, n, static_cast<int>(dpad / Q_BK),
                                                                  scale, idesc, rowpart, colpart, self_mask, scale_dev_ptr());
    CLIBD_KERNEL_CHECK();
    return 0;
}

}  // namespace clibd
