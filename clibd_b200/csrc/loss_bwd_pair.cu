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
// Warp roles per CTA: 0 TMA producer, 1 MMA issuer (leader CTA only), 2 TMEM allocator, 4-7 epilogue.
#include "common.cuh"
#include "loss_plan.h"
#include "ptx.cuh"
#include "tmap.h"

namespace clibd {
namespace {

using namespace ptx;

constexpr float kLog2e = 1.4426950408889634f;

constexpr int P_BK = 64;                          // K block (128 B of 16-bit operands)
constexpr int P_XKB_BYTES = 64 * 128;             // 8 KB : 64 rows x one K block
constexpr int P_SUB_BYTES = 128 * 128;            // 16 KB: up to 128 rows x one K block
constexpr int P_STAGE_BYTES = 2 * P_SUB_BYTES;    // a stage carries two sub-tiles (8 MMAs of 64 cycles)
constexpr int P_STAGES = 3;
constexpr int P_MAX_KB = PAIR_DCH / P_BK;         // 12
constexpr int P_THREADS = 256;
constexpr uint32_t P_TMEM_S_COL = 384;
constexpr int P_PF_DIST = 2;                      // L2 prefetch distance in column tiles

constexpr int P_SMEM_X = 0;                                         // resident Xhat rows: 96 KB
constexpr int P_SMEM_G = P_SMEM_X + P_MAX_KB * P_XKB_BYTES;         // G~: 4 K blocks x 8 KB
constexpr int P_SMEM_RING = P_SMEM_G + (PAIR_BJ / P_BK) * P_XKB_BYTES;
constexpr int P_SMEM_CC = P_SMEM_RING + P_STAGES * P_STAGE_BYTES;   // float [256] column coefficients
constexpr int P_SMEM_BARS = P_SMEM_CC + PAIR_BJ * 4;
constexpr int P_NUM_BARS = 2 * P_STAGES + 6;
constexpr int P_SMEM_TMEMPTR = P_SMEM_BARS + P_NUM_BARS * 8;
constexpr int P_SMEM_TOTAL = P_SMEM_TMEMPTR + 16;
constexpr int P_SMEM_ALLOC = P_SMEM_TOTAL + 1024;
static_assert(P_SMEM_ALLOC <= 232448, "pair backward kernel shared memory exceeds 227 KB");

template <bool ST>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
loss_bwd_pair_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
                     const __grid_constant__ CUtensorMap tm_yt, int64_t N, int64_t ld, int64_t dvalid, int64_t row0,
                     int64_t n, int num_kb, int npieces, int piece_w, int64_t tiles_per_split, float scale,
                     uint32_t idesc_s, uint32_t idesc_g, int fmt_bf16, const float* __restrict__ rowcoef,
                     const float* __restrict__ colcoef, const float* __restrict__ gscale, float weight, int accumulate,
                     float* __restrict__ dxh) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* xres = smem + P_SMEM_X;
    uint8_t* gbuf = smem + P_SMEM_G;
    uint8_t* ring = smem + P_SMEM_RING;
    float* ccbuf = reinterpret_cast<float*>(smem + P_SMEM_CC);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_SMEM_BARS);
    uint64_t* full = bars;                      // leader: TMA bytes of both CTAs landed
    uint64_t* empty = bars + P_STAGES;          // every CTA: the MMAs reading this stage have completed
    uint64_t* xfull = bars + 2 * P_STAGES;      // leader: resident Xhat rows of both CTAs landed
    uint64_t* st_full = xfull + 1;              // every CTA: S tile ready in TMEM
    uint64_t* st_empty = xfull + 2;             // leader: S tile drained by both epilogues (8 warps)
    uint64_t* g_full = xfull + 3;               // leader: G~ written by both epilogues (8 warps)
    uint64_t* g_empty = xfull + 4;              // every CTA: G~ consumed by the gradient MMAs
    uint64_t* acc_full = xfull + 5;             // every CTA: all accumulation finished
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + P_SMEM_TMEMPTR);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int64_t mt = blockIdx.x >> 1;
    const int64_t split = blockIdx.z;
    const int64_t num_jt = (N + PAIR_BJ - 1) / PAIR_BJ;
    const int64_t jt0 = split * tiles_per_split;
    const int64_t jt1 = (jt0 + tiles_per_split < num_jt) ? (jt0 + tiles_per_split) : num_jt;
    const int half_w = piece_w >> 1;                                  // accumulator columns per piece per CTA
    const int num_sst = (num_kb + 1) / 2;                             // S stages per column tile
    const int units = (PAIR_BJ / P_BK) * npieces;                     // (K block of the tile, piece) pairs
    const int num_gst = (units + 1) / 2;                              // gradient stages per column tile

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
        mbar_init(st_empty, 8);
        mbar_init(g_full, 8);
        mbar_init(g_empty, 1);
        mbar_init(acc_full, 1);
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
    const uint32_t smem_base = smem_u32(smem);

    if (jt0 < jt1) {
        // In the ST (static) instantiation d = 768: 12 K blocks, 3 gradient pieces of 256 columns.  A column tile
        // is then 6 S stages + 6 gradient stages = 4 full turns of the 3-stage ring, so the ring slot and the
        // mbarrier parity of every stage are compile-time constants and the loops below unroll into straight-line
        // code (the generic loops cost ~530 issue cycles per 512-cycle stage: the tensor pipe starved on the issuer).
        const int nkb = ST ? 12 : num_kb;
        const int npc = ST ? 3 : npieces;
        const int pw = ST ? 256 : piece_w;
        const int hw = pw >> 1;
        const int n_sst = ST ? 6 : num_sst;
        const int n_gst = ST ? 6 : num_gst;
        const int n_units = ST ? 12 : units;
        if (warp == 0) {  // ---------------- TMA producer (both CTAs; one elected lane issues)
            int stage = 0;
            uint32_t phase = 0;
            const int32_t xrow = static_cast<int32_t>(row0 + mt * PAIR_BM + rank * 64);
            const uint32_t xfull_l = mapa_u32(smem_u32(xfull), 0);
            uint32_t full_l[P_STAGES];
#pragma unroll
            for (int i = 0; i < P_STAGES; ++i) full_l[i] = mapa_u32(smem_u32(&full[i]), 0);
            const int32_t ytrow = static_cast<int32_t>(rank * hw);
            auto advance = [&]() {
                if (++stage == P_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            if (elect_one()) {
                if (leader) mbar_arrive_expect_tx(xfull, 2u * nkb * P_XKB_BYTES);
                for (int kb = 0; kb < nkb; ++kb)
                    tma_load_2d_cg2(&tm_x, xfull_l, xres + kb * P_XKB_BYTES, kb * P_BK, xrow, kEvictNormal);
            }
            // all CTA pairs of a wave sweep the same columns in lockstep, so nobody warms L2 for anybody else:
            // every CTA pulls its own future boxes into L2 P_PF_DIST tiles ahead (DRAM latency would otherwise
            // be exposed on every ring refill; the ring itself only covers an L2 hit).
            if (elect_one()) {
                for (int64_t t = jt0; t < jt1 && t < jt0 + P_PF_DIST; ++t) {
                    for (int kb = 0; kb < nkb; ++kb)
                        tma_prefetch_2d(&tm_y, kb * P_BK, static_cast<int32_t>(t * PAIR_BJ + rank * 128));
                    for (int kb2 = 0; kb2 < PAIR_BJ / P_BK; ++kb2)
                        for (int pc = 0; pc < npc; ++pc)
                            tma_prefetch_2d(&tm_yt, static_cast<int32_t>(t * PAIR_BJ + kb2 * P_BK), pc * pw + ytrow);
                }
            }
            __syncwarp();
            auto load_grad = [&](int64_t t) {
                const int32_t jcol = static_cast<int32_t>(t * PAIR_BJ);
                const bool pf = t + P_PF_DIST < jt1;
#pragma unroll
                for (int gs = 0; gs < n_gst; ++gs) {
                    const int st = ST ? (gs % P_STAGES) : stage;
                    const uint32_t ph = ST ? static_cast<uint32_t>((gs / P_STAGES) & 1) : phase;
                    mbar_wait(&empty[st], ph ^ 1);
                    if (elect_one()) {
                        uint8_t* sb = ring + st * P_STAGE_BYTES;
                        const int nu = (n_units - 2 * gs) < 2 ? (n_units - 2 * gs) : 2;
                        if (leader) mbar_arrive_expect_tx(&full[st], 2u * nu * hw * 128);
#pragma unroll
                        for (int uu = 0; uu < 2; ++uu) {
                            if (uu < nu) {
                                const int u = 2 * gs + uu;
                                const int kb2 = u / npc, pc = u - kb2 * npc;
                                tma_load_2d_cg2(&tm_yt, full_l[st], sb + uu * P_SUB_BYTES, jcol + kb2 * P_BK,
                                                pc * pw + ytrow, kEvictNormal);
                                if (pf) tma_prefetch_2d(&tm_yt, jcol + P_PF_DIST * PAIR_BJ + kb2 * P_BK, pc * pw + ytrow);
                            }
                        }
                    }
                    __syncwarp();
                    advance();
                }
            };
            for (int64_t t = jt0; t < jt1; ++t) {
                const int32_t yrow = static_cast<int32_t>(t * PAIR_BJ + rank * 128);
                const bool pf = t + P_PF_DIST < jt1;
#pragma unroll
                for (int ss = 0; ss < n_sst; ++ss) {
                    const int st = ST ? (ss % P_STAGES) : stage;
                    const uint32_t ph = ST ? static_cast<uint32_t>((ss / P_STAGES) & 1) : phase;
                    mbar_wait(&empty[st], ph ^ 1);
                    if (elect_one()) {
                        uint8_t* sb = ring + st * P_STAGE_BYTES;
                        const int nkk = (nkb - 2 * ss) < 2 ? (nkb - 2 * ss) : 2;
                        if (leader) mbar_arrive_expect_tx(&full[st], 2u * nkk * P_SUB_BYTES);
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            if (kk < nkk) {
                                tma_load_2d_cg2(&tm_y, full_l[st], sb + kk * P_SUB_BYTES, (2 * ss + kk) * P_BK, yrow,
                                                kEvictNormal);
                                if (pf) tma_prefetch_2d(&tm_y, (2 * ss + kk) * P_BK, yrow + P_PF_DIST * PAIR_BJ);
                            }
                        }
                    }
                    __syncwarp();
                    advance();
                }
                if (t > jt0) load_grad(t - 1);
            }
            load_grad(jt1 - 1);
        } else if (warp == 1 && leader) {  // ---------------- MMA issuer (leader CTA)
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t s_tmem = tmem_base + P_TMEM_S_COL;
            // descriptors of offset 0 of every operand region; tiles inside a region are reached by adding
            // (byte offset >> 4) to the low word (all regions lie below 256 KB, no carry into other fields)
            const uint64_t dx0 = make_sw128_kmajor_desc(smem_u32(xres));
            const uint64_t dg0 = make_sw128_kmajor_desc(smem_u32(gbuf));
            const uint64_t dr0 = make_sw128_kmajor_desc(smem_u32(ring));
            auto advance = [&]() {
                if (++stage == P_STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            };
            auto issue_grad = [&](int64_t tl) {  // tl = t - jt0 of the G~ tile to consume
                mbar_wait(g_full, tl & 1);
                tc_fence_after();
                const uint32_t acc_first = tl > 0 ? 1u : 0u;
#pragma unroll
                for (int gs = 0; gs < n_gst; ++gs) {
                    const int st = ST ? (gs % P_STAGES) : stage;
                    const uint32_t ph = ST ? static_cast<uint32_t>((gs / P_STAGES) & 1) : phase;
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
                    const int nu = (n_units - 2 * gs) < 2 ? (n_units - 2 * gs) : 2;
                    if (elect_one()) {
#pragma unroll
                        for (int uu = 0; uu < 2; ++uu) {
                            if (uu < nu) {
                                const int u = 2 * gs + uu;
                                const int kb2 = u / npc, pc = u - kb2 * npc;
                                const uint64_t da = dg0 + ((kb2 * P_XKB_BYTES) >> 4);
                                const uint64_t db = dr0 + ((st * P_STAGE_BYTES + uu * P_SUB_BYTES) >> 4);
#pragma unroll
                                for (int k = 0; k < P_BK / 16; ++k)
                                    umma_f16_cg2(tmem_base + pc * hw, da + 2 * k, db + 2 * k, idesc_g,
                                                 (kb2 > 0 || k > 0) ? 1u : acc_first);
                            }
                        }
                        umma_commit_cg2(&empty[st], 3);
                        if (gs == n_gst - 1) umma_commit_cg2(g_empty, 3);
                    }
                    __syncwarp();
                    advance();
                }
            };
            mbar_wait(xfull, 0);
            tc_fence_after();
            for (int64_t t = jt0; t < jt1; ++t) {
                const int64_t tl = t - jt0;
                mbar_wait(st_empty, (tl & 1) ^ 1);
                tc_fence_after();
#pragma unroll
                for (int ss = 0; ss < n_sst; ++ss) {
                    const int st = ST ? (ss % P_STAGES) : stage;
                    const uint32_t ph = ST ? static_cast<uint32_t>((ss / P_STAGES) & 1) : phase;
                    mbar_wait(&full[st], ph);
                    tc_fence_after();
                    const int nkk = (nkb - 2 * ss) < 2 ? (nkb - 2 * ss) : 2;
                    if (elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < 2; ++kk) {
                            if (kk < nkk) {
                                const uint64_t da = dx0 + (((2 * ss + kk) * P_XKB_BYTES) >> 4);
                                const uint64_t db = dr0 + ((st * P_STAGE_BYTES + kk * P_SUB_BYTES) >> 4);
#pragma unroll
                                for (int k = 0; k < P_BK / 16; ++k)
                                    umma_f16_cg2(s_tmem, da + 2 * k, db + 2 * k, idesc_s, (ss | kk | k) ? 1u : 0u);
                            }
                        }
                        umma_commit_cg2(&empty[st], 3);
                        if (ss == n_sst - 1) umma_commit_cg2(st_full, 3);
                    }
                    __syncwarp();
                    advance();
                }
                if (tl > 0) issue_grad(tl - 1);
            }
            issue_grad(jt1 - jt0 - 1);
            if (elect_one()) umma_commit_cg2(acc_full, 3);
            __syncwarp();
        } else if (warp >= 4) {  // ---------------- epilogue (both CTAs)
            const int q = warp & 3;
            const int tl_lane = q * 32 + lane;          // TMEM lane of this thread
            const int rloc = tl_lane & 63;              // row inside this CTA's 64-row slab
            const int h = tl_lane >> 6;                 // which half of the tile / piece columns this lane holds
            const int etid = (warp - 4) * 32 + lane;    // 0..127
            const int64_t lrow = mt * PAIR_BM + rank * 64 + rloc;
            const float gs = gscale[0];
            const float rcg = (lrow < n ? rowcoef[row0 + lrow] : 0.f) * gs;
            const float a = scale * kLog2e;
            const float nb = -scale * kLog2e;
            const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
            const uint32_t st_empty_l = mapa_u32(smem_u32(st_empty), 0);
            const uint32_t g_full_l = mapa_u32(smem_u32(g_full), 0);
            for (int64_t t = jt0; t < jt1; ++t) {
                const int64_t tl = t - jt0;
                asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone finished reading the previous coefficients
                {
                    const int64_t gj = t * PAIR_BJ + etid;
                    ccbuf[etid] = (gj < N) ? colcoef[gj] * gs : 0.f;
                    ccbuf[etid + 128] = (gj + 128 < N) ? colcoef[gj + 128] * gs : 0.f;
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const int64_t nvalid = N - t * PAIR_BJ;  // columns of this tile that exist (>= 256: all)
                const float* cc = ccbuf + h * 128;
                mbar_wait(st_full, tl & 1);
                tc_fence_after();
                uint32_t packed[64];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_base + lane_base + P_TMEM_S_COL + c * 32, v);
                    tmem_ld_wait();
                    if (c == 3) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(st_empty_l);
                    }
                    if (nvalid >= PAIR_BJ) {
#pragma unroll
                        for (int k = 0; k < 32; k += 4) {
                            const float4 c4 = *reinterpret_cast<const float4*>(cc + c * 32 + k);
                            const float e0 = ex2_approx(fmaf(__uint_as_float(v[k]), a, nb));
                            const float e1 = ex2_approx(fmaf(__uint_as_float(v[k + 1]), a, nb));
                            const float e2 = ex2_approx(fmaf(__uint_as_float(v[k + 2]), a, nb));
                            const float e3 = ex2_approx(fmaf(__uint_as_float(v[k + 3]), a, nb));
                            packed[c * 16 + k / 2] = pack2_operand16(e0 * (rcg + c4.x), e1 * (rcg + c4.y), fmt_bf16);
                            packed[c * 16 + k / 2 + 1] = pack2_operand16(e2 * (rcg + c4.z), e3 * (rcg + c4.w), fmt_bf16);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 32; k += 2) {
                            const int jl = h * 128 + c * 32 + k;
                            const float e0 = ex2_approx(fmaf(__uint_as_float(v[k]), a, nb));
                            const float e1 = ex2_approx(fmaf(__uint_as_float(v[k + 1]), a, nb));
                            const float g0 = (jl < nvalid) ? e0 * (rcg + cc[c * 32 + k]) : 0.f;
                            const float g1 = (jl + 1 < nvalid) ? e1 * (rcg + cc[c * 32 + k + 1]) : 0.f;
                            packed[c * 16 + k / 2] = pack2_operand16(g0, g1, fmt_bf16);
                        }
                    }
                }
                // the previous G~ tile must have been consumed before it is overwritten
                mbar_wait(g_empty, (tl & 1) ^ 1);
                const uint32_t rowaddr = smem_u32(gbuf) + (2 * h) * P_XKB_BYTES + rloc * 128;
#pragma unroll
                for (int kbh = 0; kbh < 2; ++kbh) {
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const uint32_t addr = rowaddr + kbh * P_XKB_BYTES + ((ch ^ (rloc & 7)) << 4);
                        const int p = kbh * 32 + ch * 4;
                        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(packed[p]),
                                     "r"(packed[p + 1]), "r"(packed[p + 2]), "r"(packed[p + 3])
                                     : "memory");
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(g_full_l);
            }
            // drain the accumulators: dxh (+)= weight / gscale * acc
            mbar_wait(acc_full, 0);
            tc_fence_after();
            const float wgt = weight * gscale[1];
            float* out = dxh + (split * n + lrow) * ld;
            const bool vec_ok = (ld & 3) == 0;  // rows of dxh are then 16-byte aligned
            for (int pc = 0; pc < npieces; ++pc) {
                for (int c0 = 0; c0 < half_w; c0 += 32) {
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
    tc_fence_before();
    cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still signal it or use its operands
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_cg2(tmem_base, 512);
    }
}

}  // namespace

bool pair_backward_supported(int64_t dpad) { return dpad <= PAIR_DCH; }

int tc_backward_rows_pair(const void* xh_x, const void* xh_y, const void* xhT_y, int64_t N, int64_t npad, int64_t d,
                          int64_t dpad, int64_t row0, int64_t n, float scale, const float* rowcoef,
                          const float* colcoef, const float* gscale, float weight, int accumulate, int jsplit,
                          int fmt_bf16, float* dxh, cudaStream_t s) {
    if (n == 0 || N == 0) return 0;
    CLIBD_REQUIRE(dpad % P_BK == 0 && dpad <= PAIR_DCH, "pair backward needs a padded feature dim <= 768");
    static bool attr_set = false;
    if (!attr_set) {
        CLIBD_CHECK_CUDA(cudaFuncSetAttribute(loss_bwd_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_ALLOC));
        CLIBD_CHECK_CUDA(cudaFuncSetAttribute(loss_bwd_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_ALLOC));
        attr_set = true;
    }
    // gradient pieces: npieces equal pieces of width piece_w <= 256, piece_w a multiple of 16 (8 rows of YhatT per
    // swizzle atom and per CTA)
    int npieces = 1;
    while (dpad % npieces != 0 || dpad / npieces > 256 || (dpad / npieces) % 16 != 0) ++npieces;
    const int piece_w = static_cast<int>(dpad / npieces);
    CUtensorMap tm_x, tm_y, tm_yt;
    int rc = make_tmap_2d_16bit(&tm_x, xh_x, N, dpad, dpad, P_BK, 64, fmt_bf16);
    if (rc) return rc;
    rc = make_tmap_2d_16bit(&tm_y, xh_y, N, dpad, dpad, P_BK, 128, fmt_bf16);
    if (rc) return rc;
    rc = make_tmap_2d_16bit(&tm_yt, xhT_y, dpad, npad, npad, P_BK, piece_w / 2, fmt_bf16);
    if (rc) return rc;
    const int64_t num_jt = ceil_div(N, PAIR_BJ);
    const int64_t tiles_per_split = ceil_div(num_jt, jsplit);
    const uint32_t idesc_s = make_idesc_f16(PAIR_BM, PAIR_BJ, fmt_bf16 ? 1u : 0u);
    const uint32_t idesc_g = make_idesc_f16(PAIR_BM, piece_w, fmt_bf16 ? 1u : 0u);
    dim3 grid(static_cast<unsigned>(2 * ceil_div(n, PAIR_BM)), 1, static_cast<unsigned>(jsplit));
    ProfScope prof(PROF_LOSS_BWD_TC, s);
    auto kern = (dpad == PAIR_DCH && npieces == 3) ? loss_bwd_pair_kernel<true> : loss_bwd_pair_kernel<false>;
    kern<<<grid, P_THREADS, P_SMEM_ALLOC, s>>>(tm_x, tm_y, tm_yt, N, d, d, row0, n, static_cast<int>(dpad / P_BK), npieces,
                                               piece_w, tiles_per_split, scale, idesc_s, idesc_g, fmt_bf16, rowcoef,
                                               colcoef, gscale, weight, accumulate, dxh);
    CLIBD_KERNEL_CHECK();
    return 0;
}

}  // namespace clibd
