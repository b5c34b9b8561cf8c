// Scratch layout and internal launch interfaces of the contrastive-loss path.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace clibd {

// tcgen05 tile geometry (loss_tc.cu)
constexpr int FWD_BM = 128;   // rows of S per CTA tile
constexpr int FWD_BN = 256;   // columns of S per CTA tile
constexpr int BWD_BM = 128;   // rows of S (= rows of dX) per CTA
constexpr int BWD_BJ = 128;   // columns of S per step
constexpr int BWD_DCH = 384;  // columns of dX accumulated in TMEM per CTA
// CTA-pair backward (loss_bwd_pair.cu, cta_group::2)
constexpr int PAIR_BM = 128;   // rows of dX per CTA pair (64 per CTA)
constexpr int PAIR_BJ = 256;   // columns of S per step
constexpr int PAIR_DCH = 768;  // feature columns a pair accumulates (384 TMEM columns per CTA)
constexpr int PAIR_SLOTS = 74; // CTA pairs resident at once on 148 SMs
// CUDA-core fp32 tile geometry (loss_simt.cu)
constexpr int SIMT_T = 64;    // forward tile (rows == cols)
constexpr int SIMT_BR = 32;   // backward rows per block

struct LossPlan {
    int64_t N = 0, n = 0, d = 0, dpad = 0, npad = 0;
    int path = 0;
    bool bwd_single = false;  // development switch: single-CTA sweep instead of the CTA-pair kernel
    int jsplit = 1;          // backward: column range split across CTAs (partials summed later)
    int64_t row_parts = 0;   // forward: number of row-sum partials per row
    int64_t col_parts = 0;   // forward: number of col-sum partials per column
    size_t rowpart_elems = 0, colpart_elems = 0;  // floats of one set of forward partials
    int part_sets = 1;       // 3: one set per pair (reduced together in one launch), 1: shared
    size_t off_rep = 0, off_cnt = 0, off_gscale = 0, off_scale = 0;  // off_scale: the call's logit scale on the device
    // xh: 16-bit unit rows in INPUT order (row operand); xhS / xhT: the same rows in CLASS-SORTED order and their
    // transpose (column operands: every row's positives are then one contiguous column range)
    size_t off_xh[3] = {0, 0, 0}, off_xhS[3] = {0, 0, 0}, off_xhT[3] = {0, 0, 0}, off_Q[3] = {0, 0, 0}, off_dxh[3] = {0, 0, 0};
    // class_lo[i]: first sorted position of the class of row i; ccS: the current sweep's column coefficients in
    // sorted order; posrow2[2p+dir][n]: xhat_i . Q_partner[rep_i] per pair and direction; lam2[2p+dir][n]: the part
    // of the "- 2 T_ij" target term that the tensor-core epilogue subtracts (see loss_bwd_pair.cu)
    size_t off_cstart = 0, off_class_lo = 0, off_ccS = 0, off_posrow2 = 0, off_lam2 = 0;
    // backward with S computed once per pair (loss_api.cu: backward_shared_s): xhTo = transposed operand of the LOCAL
    // rows in input order [dpad, npad_loc], Qw = class sums weighted by 1 - lam2/2 (one per pair), gt = strip of
    // coefficients [n local rows, strip columns].  One GPU: n == N.  Row-sharded (mode = LOSS_MODE_EXCHANGE): the
    // other side's gradient leaves the rank as partial rows (reduce-scatter, or peer-memory slots) and lam2 covers
    // all N rows of the pair's row modality.
    bool shared_s = false;
    bool exchange = false;
    int64_t strip_rows = 0, gt_ld = 0, npad_loc = 0;
    size_t off_xhTo[3] = {0, 0, 0}, off_Qw[3] = {0, 0, 0}, off_gt = 0, off_gt2 = 0, off_gt3 = 0;  // gt2, gt3: further strips (sharded step)
    size_t off_u = 0, off_v = 0, off_rowpart = 0, off_colpart = 0, off_posrow = 0, off_dots = 0;
    size_t off_red = 0;  // small double buffer for block reductions
    // label hash table (own/min/count per slot), rows sorted by class (keys, indices), sort input and CUB scratch
    size_t off_hown = 0, off_hmin = 0, off_hcnt = 0, off_skey = 0, off_sidx = 0, off_iota = 0, off_sorttmp = 0;
    size_t sort_tmp_bytes = 0;
    size_t total = 0;
};

// mode (include/clibd_b200.h): 0 = every rank produces both gradients of its rows itself (two sweeps per pair when the
// rows are sharded), 1 = exchange (S once per pair on every rank; column-side gradient partials are exchanged)
constexpr int LOSS_MODE_LOCAL = 0;
constexpr int LOSS_MODE_EXCHANGE = 1;
// Development / test switches read from the environment (CLIBD_SHARED_S_MIN_N, CLIBD_GT_STRIP_MB, CLIBD_JSPLIT_MAX,
// CLIBD_BWD_TWO_SWEEPS, CLIBD_BWD_SINGLE).  They shape the scratch layout, so the forward snapshots them per scratch
// pointer and the later phases of the same step (finish, backward) reuse that snapshot: an environment change between
// a forward and its backward cannot shift the offsets under the backward's feet.
struct PlanKnobs {
    int64_t shared_min_n = 4096;
    double strip_mb = 2304.0;
    int64_t js_cap = 8;
    bool two_sweeps = false;
    bool bwd_single = false;
};
PlanKnobs plan_knobs_from_env();
void remember_plan_knobs(const void* scratch, const PlanKnobs& k);
PlanKnobs plan_knobs_for(const void* scratch);  // the forward's snapshot, else the environment
LossPlan make_loss_plan(int64_t N, int64_t n_local, int64_t d, int path, bool allow_shared_s = true, int mode = 0,
                        const PlanKnobs* knobs = nullptr);

template <typename T>
inline T* at(void* base, size_t off) {
    return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + off);
}

// ---- support kernels (loss_support.cu); all return 0 / error code ------------------
int launch_row_inv_norm(const void* x, int dtype, int64_t n, int64_t d, float* inv_norm, cudaStream_t s);
// rep[i] = lowest index with the same label, cnt[i] = number of rows sharing label i; also the rows sorted by
// (rep, index): ls.skey / ls.sidx
struct LabelScratch {
    int32_t *own, *hmin, *hcnt, *skey, *sidx, *iota;
    void* sort_tmp;
    size_t sort_tmp_bytes;
};
int64_t label_hash_slots(int64_t N);
size_t class_sort_temp_bytes(int64_t N);
int launch_label_stats(const int64_t* labels, int64_t N, int32_t* rep, float* cnt, const LabelScratch& ls,
                       cudaStream_t s);
// gscale[0] = power-of-two scale applied to 16-bit G operands (1 for bf16), gscale[1] = 1/gscale[0]
int launch_gscale(const float* cnt, int64_t N, int path, float* gscale, cudaStream_t s);
// Q[r,:] = sum over rows j with rep[j]==r of xhat_j (fp32), for every representative r whose class has a member
// among the local rows [row0, row0 + n) (the only class sums a rank reads)
int launch_class_sums(const void* x, int dtype, const float* inv_norm, const int32_t* skey, const int32_t* sidx,
                      const float* cnt, int64_t N, int64_t d, int64_t row0, int64_t n, float* Q, cudaStream_t s,
                      const float* lam2 = nullptr);
// (lam2 != null, indexed by row: member j enters with weight 1 - lam2[j] / 2 -- the share of the target term
//  "- 2 T" that the tensor-core sweep has NOT already subtracted on row j's positives)
// 16-bit normalised operand copies: xh [N,dpad] in input order; with perm != null also xhS [N,dpad] whose row k is
// input row perm[k], and the transpose xhT [dpad,npad] (zero padded) follows that order (perm == null: input order)
int launch_make_operands(const void* x, int dtype, const float* inv_norm, int64_t N, int64_t d, int64_t dpad,
                         int64_t npad, int fmt_bf16, void* xh, void* xhT, cudaStream_t s,
                         const int32_t* perm = nullptr, void* xhS = nullptr);
// ---- batched forms: several independent jobs of the same kind in ONE launch (blockIdx.y = job).  A sharded step at
// 8 GPUs is ~2.5 ms long; six 9-microsecond launches of a reduction are 2 % of it.
constexpr int MAX_JOBS = 6;
struct ClassSumJob {      // up to two weighted sums of the same rows in one pass: out[k][r] = sum_j w_k(j) xhat_j
    const void* x = nullptr;
    const float* inv = nullptr;
    const float* lam2[2] = {nullptr, nullptr};  // weight 1 - lam2[j] / 2 (null: 1)
    float* out[2] = {nullptr, nullptr};         // out[1] may be null
};
int launch_class_sums_jobs(const ClassSumJob* jobs, int njobs, int dtype, const int32_t* skey, const int32_t* sidx,
                           const float* cnt, int64_t N, int64_t d, int64_t row0, int64_t n, cudaStream_t s);
struct ReduceJob {
    const float* part = nullptr;
    int64_t parts = 0, stride = 0, len = 0;
    float* out = nullptr;
    const int32_t* scatter = nullptr;
};
int launch_reduce_parts_jobs(const ReduceJob* jobs, int njobs, cudaStream_t s);
struct PosRowsJob {
    const void* xa = nullptr;
    const float* inv_a = nullptr;
    const float* Qb = nullptr;
    float* posrow = nullptr;
};
int launch_pos_rows_jobs(const PosRowsJob* jobs, int njobs, int dtype, const int32_t* rep, int64_t d, int64_t row0,
                         int64_t n, cudaStream_t s);
struct SumJob {
    const float* in = nullptr;
    int64_t len = 0;
    double* out = nullptr;
};
// red: njobs * 256 doubles
int launch_sum_to_double_jobs(const SumJob* jobs, int njobs, double mul, double* red, cudaStream_t s,
                              const float* div_dev = nullptr);
struct SweepPrepJob {
    const float *rowcoef = nullptr, *colcoef = nullptr, *posrow = nullptr;
    float *ccS = nullptr, *lam2 = nullptr;
};
int launch_sweep_prep_jobs(const SweepPrepJob* jobs, int njobs, const int32_t* sidx, const float* cnt, int64_t N,
                           int64_t row0, int64_t n, float scale, cudaStream_t s);

// class_lo[i] = first position of the class of row i in the class-sorted order (skey = sorted representatives)
int launch_class_ranges(const int32_t* skey, const int32_t* rep, int64_t N, int32_t* cstart, int32_t* class_lo,
                        cudaStream_t s);
// Per backward sweep: ccS[k] = colcoef[sidx[k]] (column coefficients in sorted order; sidx == null: copy) and, when
// lam2 != null, lam2[i] = 2 min(1, 0.5 exp(s posrow_i / cnt_i - s) (rowcoef_i + colcoef_i)) for the local rows
int launch_sweep_prep(const float* rowcoef, const float* colcoef, const int32_t* sidx, const float* cnt,
                      const float* posrow, int64_t N, int64_t row0, int64_t n, float scale, float* ccS, float* lam2,
                      cudaStream_t s);
// posrow[i] = xhat_a[row0+i] . Qb[rep[row0+i]]
int launch_pos_rows(const void* xa, int dtype, const float* inv_a, const float* Qb, const int32_t* rep, int64_t d,
                    int64_t row0, int64_t n, float* posrow, cudaStream_t s);
// out[k] = sum_{p < parts} part[p*stride + k] for k < len (fixed order)
// (scatter != null: out[scatter[k]] instead of out[k])
int launch_reduce_parts(const float* part, int64_t parts, int64_t stride, int64_t len, float* out, cudaStream_t s,
                        const int32_t* scatter = nullptr);
// out[0] = mul * sum_k in[k] in double, fixed order
// (div_dev != null: additionally divided by the device scalar div_dev[0])
int launch_sum_to_double(const float* in, int64_t len, double mul, double* red, double* out, cudaStream_t s,
                         const float* div_dev = nullptr);
// loss + backward coefficients u = cnt/rowsum, v = cnt/colsum
int launch_loss_finish(int64_t N, float scale, const float w[3], const float* cnt, const float* rowsum,
                       const float* colsum, const double* pos, float* u, float* v, double* red, float* loss_out,
                       cudaStream_t s);
// dx = grad_scale * normalize_bwd( (scale/N) * (sum_splits dxh - sum_partners w_p (2 - lam2_p,i) Q_partner[rep]) );
// dots[i] = xhat_i . dxhat_i for unit grad
struct NormBwdArgs {
    const void* x;
    int dtype;
    const float* inv_norm;
    const int32_t* rep;
    const float* dxh;       // [jsplit][n][d]
    int jsplit;             // 0: no row-side partials
    const float* extra = nullptr;  // [extra_slots][n][d] further partials (column-side gradients received from the
    int extra_slots = 0;           //  exchange: one slot per (pair, source rank), summed in slot order)
    const float* Qp[2];     // partner class sums (may be null)
    float wp[2];            // their pair weights
    const float* lam2[2] = {nullptr, nullptr};  // [n] part of the target term already subtracted by the sweep (null: 0)
    int64_t N, d, row0, n;
    float scale, grad_scale;
    const float* scale_dev = nullptr;  // filled by launch_normalize_bwd from the call's ScaleScope
    const float* grad_scale_dev;  // optional device scalar(s) multiplied into grad_scale (may be null)
    int grad_scale_dev_count = 1;  // > 1: the SUM of that many device scalars (one grad_output per rank)
    void* dx;               // [n,d] in dtype, may be null
    float* dots;            // [n]
};
int launch_normalize_bwd(const NormBwdArgs& a, cudaStream_t s);
// up to three modalities (same n, dtype) in one launch
int launch_normalize_bwd_multi(const NormBwdArgs* args, int count, cudaStream_t s);

// ---- CUDA-core fp32 path (loss_simt.cu) ---------------------------------------------
int simt_forward_pair(const void* xa, const void* xb, int dtype, const float* inv_a, const float* inv_b, int64_t N,
                      int64_t d, int64_t row0, int64_t n, float scale, float* rowpart, float* colpart,
                      cudaStream_t s, int self_mask = 0);
// dxh[n,d] (+)= weight * sum_j e_ij (rowcoef_i + colcoef_j) yhat_j   for local rows i of x
int simt_backward_rows(const void* x, const void* y, int dtype, const float* inv_x, const float* inv_y, int64_t N,
                       int64_t d, int64_t row0, int64_t n, float scale, const float* rowcoef, const float* colcoef,
                       float weight, int accumulate, float* dxh, cudaStream_t s, int self_mask = 0, int jsplit = 1);

// ---- tcgen05 path (loss_tc.cu) --------------------------------------------------------
int tc_forward_pair(const void* xh_a, const void* xh_b, int64_t N, int64_t dpad, int64_t row0, int64_t n, float scale,
                    int fmt_bf16, float* rowpart, float* colpart, cudaStream_t s);
int tc_backward_rows(const void* xh_x, const void* xh_y, const void* xhT_y, int64_t N, int64_t npad, int64_t d,
                     int64_t dpad, int64_t row0, int64_t n, float scale, const float* rowcoef, const float* colcoef,
                     const float* gscale, float weight, int accumulate, int jsplit, int fmt_bf16, float* dxh,
                     cudaStream_t s);

// CTA-pair forward (loss_fwd_pair.cu): 256 x 256 tiles, cta_group::2
int tc_forward_pair_cg2(const void* xh_a, const void* xh_b, int64_t N, int64_t dpad, int64_t row0, int64_t n, float scale,
                        int fmt_bf16, float* rowpart, float* colpart, int num_sms, cudaStream_t s, int self_mask = 0);
// CTA-pair variant (loss_bwd_pair.cu): whole feature dimension per pair, S computed once per sweep; dpad <= 768
bool pair_backward_supported(int64_t dpad);
int tc_backward_rows_pair(const void* xh_x, const void* xh_y, const void* xhT_y, int64_t N, int64_t npad, int64_t d,
                          int64_t dpad, int64_t row0, int64_t n, float scale, const float* rowcoef,
                          const float* colcoef, const float* gscale, float weight, int accumulate, int jsplit,
                          int fmt_bf16, float* dxh, cudaStream_t s, int self_mask = 0,
                          const int32_t* pos_lo = nullptr, const float* pos_cnt = nullptr, const float* lam2 = nullptr,
                          int64_t col_begin = 0, int64_t col_end = -1, void* gt = nullptr, int64_t gt_ld = 0);
// col_begin / col_end: sweep only the columns [col_begin, col_end) (col_begin a multiple of 256; -1 = N);
// gt != null: also store the 16-bit coefficient tiles, gt[LOCAL row * gt_ld + (column - col_begin)] (TMA stores;
// gt has n rows)
// Other side's gradient from such a strip (loss_grad_gemm.cu): with g = sidx[strip0 + r] the global row of strip
// column r,  dest(g)[:] (+)= weight / gscale * sum_{i < K} gs[i][r] * xhat_x[i][:]  for the Ms strip columns r; the K
// range (the rows of gs = the rows of xhT_x's K extent) is split over `ksplit` partial outputs.
// Destination of global row g: q = g / rows_per_dest, lr = g % rows_per_dest,
//     base[q] + ((slot0 + ks) * slot_rows + lr) * d
// One GPU / reduce-scatter form: rows_per_dest = N, base[0] = the local buffer.  Peer form: rows_per_dest = n,
// base[q] = rank q's slot array (peer-mapped memory: the epilogue stores travel over NVLink), slot0 = this rank's slot.
constexpr int MAX_PEERS = 16;
struct GradDest {
    float* base[MAX_PEERS];
    int64_t rows_per_dest;
    int64_t slot_rows;
    int slot0;
};
// parts: one or two (strip, transposed row operand) pairs whose K ranges are concatenated -- two modality pairs with the
// same column modality and the same weight become one GEMM (both strips have K rows and pitch gs_ld)
struct GradPart {
    const void* gs = nullptr;
    const void* xhT_x = nullptr;
};
int tc_grad_from_strip(const GradPart* parts, int nparts, int64_t gs_ld, int64_t Ms, int64_t strip0, int64_t K,
                       int64_t npad, int64_t Ntot, int64_t d, int64_t dpad, const int32_t* sidx, const float* gscale,
                       float weight, int accumulate, int ksplit, int fmt_bf16, const GradDest& dest, int num_sms,
                       cudaStream_t s);
// pos_lo / pos_cnt [N] (by global row), lam2 [n] (by local row): columns [pos_lo, pos_lo + pos_cnt) of row i are its
// positives (T_ij = 1) and the epilogue subtracts lam2_i from G~ there before the 16-bit rounding
// self_mask = 1 (all four launchers above): the operands are the SAME feature set and the entries whose global row
// equals their column are excluded from the sums (SimCLR info-NCE, bioscanclip/util/simclr.py:76-79)
int tc_num_sms();

}  // namespace clibd
