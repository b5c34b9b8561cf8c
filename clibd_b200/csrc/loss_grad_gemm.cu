// Second gradient of a modality pair from the STORED coefficient strip (single-GPU backward, see loss_api.cu).
//
// The row sweep of a pair (loss_bwd_pair.cu) computes S = Xhat Yhat^T once, turns every tile into the 16-bit
// coefficients G~ and accumulates dXhat += G~ Yhat.  The same coefficients give the other side's gradient,
//
//     dYhat[j, :] = sum_i G~[i, j] Xhat[i, :],
//
// so instead of recomputing S^T in a second sweep (2 n N d tensor flops per pair) the row sweep also sends its G~
// tiles to a bounded strip buffer Gs[i, strip column] (TMA stores of the tiles it built for its own gradient MMAs)
// and THIS kernel runs the plain GEMM
//
//     dYhat[strip columns, d] = Gs^T [M = strip columns, K = N rows] * XhatT[d, K = N]^T
//
// with Gs^T taken as an M-MAJOR A operand: the TMA box is 64 strip columns (128 B) x 64 rows, which is exactly
// the canonical MN-major SWIZZLE_128B layout of tcgen05 (8-row atoms 1024 B apart along K) -- no transposition
// anywhere.  The GEMM runs
// on CTA pairs: 256 x 256 output tiles, tcgen05 cta_group::2 M = 256, the K range split into `ksplit` partial
// outputs so that the (row tile, feature tile, K part) work items fill whole waves of 74 pairs.  Per 128-cycle
// MMA a CTA reads 4 KB of A and 4 KB of B and TMA writes 8 KB: 128 B/cycle, the shared-memory bandwidth
// (the geometry of loss_fwd_pair.cu).  Rows of the strip are class-sorted column positions of the row sweep;
// the epilogue scatters them to the input order through sidx.
//
// Two pairs that share their column modality -- (image, text) and (dna, text) -- are ONE launch: the K range is the
// concatenation of both pairs' strips and row operands ([Gs_1; Gs_2]^T [Xhat_image; Xhat_dna]), so text's gradient is
// accumulated in TMEM instead of a second read-modify-write pass over the output (or a second set of slots and a
// second trip over NVLink in the sharded step).
//
// Row-sharded step (exchange mode, loss_api.cu): K = this rank's n local rows, the output is the rank's PARTIAL
// gradient of all N rows of the column modality, and the epilogue sends every row straight to its owner: global
// row g belongs to rank g / n, whose slot array is peer-mapped memory (GradDest) -- the reduce-scatter of the
// reference's all_gather backward (loss_func.py:97) happens tile by tile over NVLink while the GEMM is running,
// and the owner adds the W slots in rank order (deterministic) in its normalise-backward pass.
//
// Warp roles per CTA: 0 TMA producer, 1 MMA issuer (leader CTA only), 2 TMEM allocator, 4-11 epilogue.
#include "common.cuh"
#include "loss_plan.h"
#include "ptx.cuh"
#include "tmap.h"

namespace clibd {
namespace {

using namespace ptx;

constexpr int G_BK = 64;
constexpr int G_A_BYTES = 128 * G_BK * 2;   // 16 KB: this CTA's 128 strip columns of Gs x 64 rows, as two 64-column boxes
constexpr int G_B_BYTES = 128 * G_BK * 2;   // 16 KB: this CTA's 128 feature rows of XhatT (half of the tile's columns)
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_STAGES = 7;
constexpr int G_THREADS = 384;
constexpr int G_EPI_WARPS = 8;
constexpr int G_TN = 256;
constexpr int G_SMEM_BARS = G_STAGES * G_STAGE_BYTES;
constexpr int G_NUM_BARS = 2 * G_STAGES + 4;
constexpr int G_SMEM_TMEMPTR = G_SMEM_BARS + G_NUM_BARS * 8;
constexpr int G_SMEM_ROWPTR = G_SMEM_TMEMPTR + 16;  // float* [8 epilogue warps][32 rows]: destination row of every lane
constexpr int G_SMEM_TOTAL = G_SMEM_ROWPTR + G_EPI_WARPS * 32 * 8;
static_assert(G_SMEM_TOTAL <= 232448, "gradient GEMM kernel shared memory exceeds 227 KB");

// In-register transpose of a 32 x 32 block held as v[k] = element (row = lane, column = k) of every lane: afterwards
// v[k] = element (row = k, column = lane).  Five butterfly steps of 16 shuffles.
__device__ __forceinline__ void transpose32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if ((k & s) == 0) {
                const float send = up ? v[k] : v[k + s];
                const float recv = __shfl_xor_sync(0xffffffffu, send, s);
                if (up) v[k] = recv;
                else v[k + s] = recv;
            }
        }
    }
}

struct GItem {
    int64_t mt, nt, ks;
};
// feature tiles fastest: the pairs working on one (row tile, K part) stream the same Gt blocks at the same time
__device__ __forceinline__ GItem g_item(int64_t t, int64_t num_nt, int ksplit) {
    GItem it;
    it.nt = t % num_nt;
    const int64_t r = t / num_nt;
    it.ks = r % ksplit;
    it.mt = r / ksplit;
    return it;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G_THREADS, 1)
loss_grad_gemm_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_xt,
                      const __grid_constant__ CUtensorMap tm_g1, const __grid_constant__ CUtensorMap tm_xt1, int num_kb0,
                      int64_t Ms, int64_t strip0, int64_t Ntot, int64_t d, int64_t ld, int num_kb, int ksplit,
                      int kb_per_split, uint32_t idesc, const int32_t* __restrict__ sidx,
                      const float* __restrict__ gscale, float weight, int accumulate, const GradDest dest) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_SMEM_BARS);
    uint64_t* full = bars;
    uint64_t* empty = bars + G_STAGES;
    uint64_t* tfull = bars + 2 * G_STAGES;
    uint64_t* tempty = bars + 2 * G_STAGES + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + G_SMEM_TMEMPTR);
    float** rowptr = reinterpret_cast<float**>(smem + G_SMEM_ROWPTR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int64_t num_mt = (Ms + 255) / 256;
    const int64_t num_nt = (d + G_TN - 1) / G_TN;
    const int64_t num_items = num_mt * num_nt * ksplit;
    const int64_t pair = blockIdx.x >> 1;
    const int64_t num_pairs = gridDim.x >> 1;

    if (threadIdx.x == 0) {
        prefetch_tmap(&tm_g);
        prefetch_tmap(&tm_xt);
        if (num_kb > num_kb0) {
            prefetch_tmap(&tm_g1);
            prefetch_tmap(&tm_xt1);
        }
        for (int i = 0; i < G_STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 2 * G_EPI_WARPS);
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
        for (int64_t t = pair; t < num_items; t += num_pairs) {
            const GItem it = g_item(t, num_nt, ksplit);
            const int32_t arow = static_cast<int32_t>(it.mt * 256 + rank * 128);
            const int32_t brow = static_cast<int32_t>(it.nt * G_TN + rank * 128);
            const int kb0 = static_cast<int>(it.ks) * kb_per_split;
            const int kb1 = kb0 + kb_per_split < num_kb ? kb0 + kb_per_split : num_kb;
#pragma unroll 1
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&empty[slot], phase ^ 1);
                if (elected) {
                    uint8_t* sa = smem + slot * G_STAGE_BYTES;
                    // K blocks [0, num_kb0) come from the first (strip, row operand) part, the rest from the second
                    const bool second = kb >= num_kb0;
                    const CUtensorMap* tg = second ? &tm_g1 : &tm_g;
                    const CUtensorMap* tx = second ? &tm_xt1 : &tm_xt;
                    const int32_t kcrd = (second ? kb - num_kb0 : kb) * G_BK;
                    if (leader) mbar_arrive_expect_tx(&full[slot], 2u * G_STAGE_BYTES);
                    tma_load_2d_cg2(tg, full_l0 + slot * 8, sa, arow, kcrd, kEvictNormal);
                    tma_load_2d_cg2(tg, full_l0 + slot * 8, sa + G_A_BYTES / 2, arow + 64, kcrd, kEvictNormal);
                    tma_load_2d_cg2(tx, full_l0 + slot * 8, sa + G_A_BYTES, kcrd, brow, kEvictLast);
                }
                __syncwarp();
                if (++slot == G_STAGES) {
                    slot = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp == 1 && leader) {  // ---------------- MMA issuer (leader CTA)
        int slot = 0;
        uint32_t phase = 0;
        uint32_t itc = 0;
        // A: M-major (two 64-column groups 8 KB apart, K = 16 step = 2 atoms = 2048 B); B: K-major
        const uint64_t da0 = make_sw128_mnmajor_desc(smem_u32(smem), G_A_BYTES / 2);
        const uint64_t db0 = make_sw128_kmajor_desc(smem_u32(smem + G_A_BYTES));
        const bool elected = elect_one();
        for (int64_t t = pair; t < num_items; t += num_pairs, ++itc) {
            const GItem it = g_item(t, num_nt, ksplit);
            const int kb0 = static_cast<int>(it.ks) * kb_per_split;
            const int kb1 = kb0 + kb_per_split < num_kb ? kb0 + kb_per_split : num_kb;
            const uint32_t as = itc & 1, aph = (itc >> 1) & 1;
            mbar_wait(&tempty[as], aph ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * G_TN;
#pragma unroll 1
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&full[slot], phase);
                tc_fence_after();
                if (elected) {
                    const uint64_t da = da0 + slot * (G_STAGE_BYTES >> 4);
                    const uint64_t db = db0 + slot * (G_STAGE_BYTES >> 4);
                    umma_f16_cg2(d_tmem, da, db, idesc, kb > kb0 ? 1u : 0u);
                    umma_f16_cg2(d_tmem, da + 128, db + 2, idesc, 1u);
                    umma_f16_cg2(d_tmem, da + 256, db + 4, idesc, 1u);
                    umma_f16_cg2(d_tmem, da + 384, db + 6, idesc, 1u);
                    umma_commit_cg2(&empty[slot], 3);
                    if (kb == kb1 - 1) umma_commit_cg2(&tfull[as], 3);
                }
                __syncwarp();
                if (++slot == G_STAGES) {
                    slot = 0;
                    phase ^= 1;
                }
            }
        }
    } else if (warp >= 4) {  // ---------------- epilogue (both CTAs): accumulator -> out rows (input order)
        const int ew = warp - 4;
        const int q = warp & 3;   // TMEM lane quadrant this warp may access
        const int h = ew >> 2;    // column half of the tile
        const float wgt = weight * gscale[1];
        const uint32_t tempty_l0 = mapa_u32(smem_u32(&tempty[0]), 0);
        const bool vec_ok = (ld & 3) == 0;
        uint32_t itc = 0;
        for (int64_t t = pair; t < num_items; t += num_pairs, ++itc) {
            const GItem it = g_item(t, num_nt, ksplit);
            const uint32_t as = itc & 1, aph = (itc >> 1) & 1;
            const int64_t srow = it.mt * 256 + rank * 128 + q * 32 + lane;  // row of the strip
            const bool row_ok = srow < Ms && strip0 + srow < Ntot;
            const int64_t orow = row_ok ? static_cast<int64_t>(sidx[strip0 + srow]) : 0;
            const int64_t oq = orow / dest.rows_per_dest;       // owner of the row (0 unless peer form)
            const int64_t olr = orow - oq * dest.rows_per_dest;
            float* orp = dest.base[oq] + ((dest.slot0 + it.ks) * dest.slot_rows + olr) * ld;
            // Stores go out ROW-CONTIGUOUS: a TMEM load leaves lane = row, register = column, i.e. a warp store would
            // touch 32 rows with 16 bytes each -- 32 partial sectors, and over NVLink 32 minimum-size packets (measured
            // with 7/8 of the rows remote: the GEMM took 2.6x its local time).  Each 32 x 32 block is transposed in
            // registers first, so that one store instruction writes the 128 contiguous bytes of ONE row; the lanes'
            // destination rows are published through shared memory.
            __syncwarp();
            rowptr[ew * 32 + lane] = row_ok ? orp : nullptr;
            __syncwarp();
            mbar_wait(&tfull[as], aph);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * G_TN + h * 128 + c * 32;
                tmem_ld_32x32b_x32(taddr, v);
                tmem_ld_wait();
                if (c == 3) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tempty_l0 + as * 8);
                }
                const int64_t col0 = it.nt * G_TN + h * 128 + c * 32;
                if (col0 >= d) continue;
                if (!accumulate && col0 + 32 <= d) {
                    float f[32];
#pragma unroll
                    for (int k = 0; k < 32; ++k) f[k] = wgt * __uint_as_float(v[k]);
                    transpose32(f, lane);
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        float* rp = rowptr[ew * 32 + k];  // same address for the whole warp: one broadcast read
                        if (rp != nullptr) rp[col0 + lane] = f[k];
                    }
                    continue;
                }
                if (!row_ok) continue;
                if (vec_ok && col0 + 32 <= d) {
                    float4* o4 = reinterpret_cast<float4*>(orp + col0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 r;
                        r.x = wgt * __uint_as_float(v[4 * i]);
                        r.y = wgt * __uint_as_float(v[4 * i + 1]);
                        r.z = wgt * __uint_as_float(v[4 * i + 2]);
                        r.w = wgt * __uint_as_float(v[4 * i + 3]);
                        if (accumulate) {
                            const float4 o = o4[i];
                            r.x += o.x;
                            r.y += o.y;
                            r.z += o.z;
                            r.w += o.w;
                        }
                        o4[i] = r;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 32; ++k) {
                        if (col0 + k < d) {
                            const float val = wgt * __uint_as_float(v[k]);
                            orp[col0 + k] = accumulate ? orp[col0 + k] + val : val;
                        }
                    }
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

int tc_grad_from_strip(const GradPart* parts, int nparts, int64_t gs_ld, int64_t Ms, int64_t strip0, int64_t K,
                       int64_t npad, int64_t Ntot, int64_t d, int64_t dpad, const int32_t* sidx, const float* gscale,
                       float weight, int accumulate, int ksplit, int fmt_bf16, const GradDest& dest, int num_sms,
                       cudaStream_t s) {
    if (Ms == 0 || K == 0 || nparts == 0) return 0;
    CLIBD_REQUIRE(nparts == 1 || nparts == 2, "the gradient GEMM concatenates at most two K parts");
    CLIBD_REQUIRE(ksplit >= 1 && gs_ld % 8 == 0 && gs_ld >= Ms, "bad strip geometry");
    CLIBD_REQUIRE(dest.rows_per_dest > 0 && ceil_div(Ntot, dest.rows_per_dest) <= MAX_PEERS, "bad gradient destination");
    CUtensorMap tm_g[2], tm_xt[2];
    for (int i = 0; i < 2; ++i) {
        const GradPart& pt = parts[i < nparts ? i : 0];
        // Gs [K rows, Ms strip columns (pitch gs_ld)]: rows beyond K / columns beyond Ms read as zero (TMA fill)
        int rc = make_tmap_2d_16bit(&tm_g[i], pt.gs, K, Ms, gs_ld, 64, G_BK, fmt_bf16);
        if (rc) return rc;
        rc = make_tmap_2d_16bit(&tm_xt[i], pt.xhT_x, dpad, K, npad, G_BK, 128, fmt_bf16);
        if (rc) return rc;
    }
    CLIBD_CHECK_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(&loss_grad_gemm_kernel), G_SMEM_TOTAL));
    const int num_kb0 = static_cast<int>(ceil_div(K, G_BK));  // K blocks per part (a ragged tail block reads zeros)
    const int num_kb = num_kb0 * nparts;
    const int slots = ksplit;  // the caller sums `slots` partial outputs: parts that get no K blocks are zeroed
    if (ksplit > num_kb) ksplit = num_kb;
    const int kb_per_split = static_cast<int>(ceil_div(num_kb, ksplit));
    ksplit = static_cast<int>(ceil_div(num_kb, kb_per_split));  // no empty K part
    if (!accumulate && ksplit < slots) {  // only the local forms split K
        CLIBD_REQUIRE(dest.rows_per_dest >= Ntot, "the peer form of the gradient GEMM does not split K");
        CLIBD_CHECK_CUDA(cudaMemsetAsync(dest.base[0] + static_cast<int64_t>(dest.slot0 + ksplit) * dest.slot_rows * d, 0,
                                         sizeof(float) * static_cast<size_t>(slots - ksplit) * dest.slot_rows * d, s));
    }
    const int64_t items = ceil_div(Ms, 256) * ceil_div(d, G_TN) * ksplit;
    // a multiple of the number of feature tiles: the pairs that work on the feature tiles of one (row tile, K part)
    // then always run in the same wave and share the Gs blocks through L2 (74 -> 72 pairs for d = 768)
    const int64_t num_nt = ceil_div(d, G_TN);
    int64_t max_pairs = num_sms / 2;
    if (max_pairs > num_nt) max_pairs -= max_pairs % num_nt;
    const int pairs = static_cast<int>(items < max_pairs ? items : max_pairs);
    const uint32_t idesc = make_idesc_f16(256, G_TN, fmt_bf16 ? 1u : 0u, /*a_mn_major=*/1u);
    ProfScope prof(PROF_LOSS_GRAD_GEMM, s);
    loss_grad_gemm_kernel<<<2 * pairs, G_THREADS, G_SMEM_TOTAL, s>>>(tm_g[0], tm_xt[0], tm_g[1], tm_xt[1], num_kb0, Ms, strip0,
                                                                   Ntot, d, d, num_kb, ksplit, kb_per_split, idesc, sidx,
                                                                   gscale, weight, accumulate, dest);
    CLIBD_KERNEL_CHECK();
    return 0;
}

}  // namespace clibd
