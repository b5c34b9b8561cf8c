// extern "C" entry points of the contrastive-loss path (see include/clibd_b200.h).
#include <atomic>
#include <cstdlib>
#include <mutex>
#include <set>
#include <unordered_map>
#include <string>
#include <utility>
#include <vector>

#include "../../include/clibd_b200.h"
#include "common.cuh"
#include "graph_cache.h"
#include "loss_plan.h"

namespace clibd {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error() { return g_last_error; }

static thread_local const float* g_scale_dev = nullptr;
const float* scale_dev_ptr() { return g_scale_dev; }
ScaleScope::ScaleScope(const float* p) : prev_(g_scale_dev) { g_scale_dev = p; }
ScaleScope::~ScaleScope() { g_scale_dev = prev_; }

// cell[0] = the call's logit scale (from the device scalar when given), NaN when it is not a positive finite number:
// a tensor scale cannot raise on the host without a device -> host read, so an invalid one poisons the loss instead.
// There is no upper limit: the reference never clamps its learnable logit_scale (simple_clip.py:32,61) and the kernels
// pick their softmax shift from the scale (common.cuh: softmax_shift).
__global__ void scale_set_kernel(float value, const float* __restrict__ dev, float* __restrict__ cell) {
    const float s = dev ? dev[0] : value;
    cell[0] = (s > 0.f && s < 1e30f) ? s : __int_as_float(0x7fc00000);
}

static std::atomic<int64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int64_t launch_count_now() { return g_launches.load(std::memory_order_relaxed); }
void add_launches(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::mutex g_prof_mu;
static bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events[PROF_SLOTS];

bool profiling_enabled() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    return g_prof_on;
}

ProfScope::ProfScope(int slot, cudaStream_t s) : slot_(slot), s_(s) {
    if (!g_prof_on) return;
    cudaEventCreate(&e0_);
    cudaEventCreate(&e1_);
    cudaEventRecord(e0_, s_);
}
ProfScope::~ProfScope() {
    if (e0_ == nullptr) return;
    cudaEventRecord(e1_, s_);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_events[slot_].emplace_back(e0_, e1_);
}

cudaError_t ensure_dynamic_smem(const void* kernel, int bytes) {
    static std::mutex mu;
    static std::set<std::pair<const void*, int>> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(mu);
    if (done.count({kernel, dev})) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done.insert({kernel, dev});
    return e;
}

// A second stream per device (and per role: forward / backward) on which HBM-bound staging work runs NEXT TO the tensor
// kernels: the forward and gradient-GEMM kernels leave registers (104 / 102 per thread x 384 threads) and issue slots for
// a 256-thread block without shared memory, and a sweep of a sharded step leaves whole SMs idle.  Fork / join by events,
// which is also the legal pattern inside a stream capture (graph_cache.cu).  CLIBD_SIDE_STREAM=0 runs everything in line.
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
};
static SideStream* side_stream(int role) {
    static std::mutex mu;
    static SideStream table[64][2];
    const char* e = std::getenv("CLIBD_SIDE_STREAM");  // read per call: tools/ab_step.py flips it inside one process
    if (e != nullptr && std::atoi(e) == 0) return nullptr;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(mu);
    SideStream& st = table[dev][role];
    if (st.stream == nullptr) {
        cudaStream_t s = nullptr;
        cudaEvent_t a = nullptr, b = nullptr;
        if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&a, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&b, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        st.stream = s;
        st.fork = a;
        st.join = b;
    }
    return &st;
}
// fork: everything enqueued on `main` so far happens before the side work; join: `main` waits for the side work
static int side_fork(SideStream* ss, cudaStream_t main) {
    CLIBD_CHECK_CUDA(cudaEventRecord(ss->fork, main));
    CLIBD_CHECK_CUDA(cudaStreamWaitEvent(ss->stream, ss->fork, 0));
    return 0;
}
static int side_join(SideStream* ss, cudaStream_t main) {
    CLIBD_CHECK_CUDA(cudaEventRecord(ss->join, ss->stream));
    CLIBD_CHECK_CUDA(cudaStreamWaitEvent(main, ss->join, 0));
    return 0;
}

static size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

PlanKnobs plan_knobs_from_env() {
    PlanKnobs k;
    if (const char* e = std::getenv("CLIBD_SHARED_S_MIN_N")) k.shared_min_n = std::atoll(e);  // tests force the path
    if (const char* e = std::getenv("CLIBD_GT_STRIP_MB")) k.strip_mb = std::atof(e);
    if (const char* e = std::getenv("CLIBD_JSPLIT_MAX")) k.js_cap = std::atoll(e);  // cap (negative: force) the split
    k.two_sweeps = std::getenv("CLIBD_BWD_TWO_SWEEPS") != nullptr;
    k.bwd_single = std::getenv("CLIBD_BWD_SINGLE") != nullptr;
    return k;
}

static std::mutex g_knob_mu;
static std::unordered_map<const void*, PlanKnobs> g_knobs;

void remember_plan_knobs(const void* scratch, const PlanKnobs& k) {
    std::lock_guard<std::mutex> lk(g_knob_mu);
    if (g_knobs.size() > 1024) g_knobs.clear();
    g_knobs[scratch] = k;
}

PlanKnobs plan_knobs_for(const void* scratch) {
    {
        std::lock_guard<std::mutex> lk(g_knob_mu);
        auto it = g_knobs.find(scratch);
        if (it != g_knobs.end()) return it->second;
    }
    return plan_knobs_from_env();
}

LossPlan make_loss_plan(int64_t N, int64_t n, int64_t d, int path, bool allow_shared_s, int mode,
                        const PlanKnobs* knobs_in) {
    const PlanKnobs knobs = knobs_in ? *knobs_in : plan_knobs_from_env();
    LossPlan p;
    p.bwd_single = knobs.bwd_single;
    p.N = N;
    p.n = n;
    p.d = d;
    p.path = path;
    p.dpad = round_up(d, 64);
    p.npad = round_up(N, 8);
    const bool tc = path != PATH_SIMT_F32;
    if (tc) {
        p.row_parts = ceil_div(N, FWD_BN);
        p.col_parts = ceil_div(n, FWD_BM);
        // split the column sweep so that the work items fill whole waves (1 CTA / SM; the pair kernel runs
        // 74 CTA pairs at once, the single-CTA kernel 148 CTAs of 384-wide chunks)
        const bool pair = p.dpad <= PAIR_DCH;
        const int64_t slots = pair ? PAIR_SLOTS : 148;
        const int64_t ctas = pair ? ceil_div(n, PAIR_BM) : ceil_div(n, BWD_BM) * ceil_div(p.dpad, BWD_DCH);
        const int64_t num_jt = pair ? ceil_div(N, PAIR_BJ) : ceil_div(N, BWD_BJ);
        int64_t js = 1;
        double best = 0.0;
        const int64_t js_cap = knobs.js_cap;
        for (int64_t c = 1; c <= (js_cap > 0 ? js_cap : 1) && c <= num_jt; ++c) {
            const int64_t total = ctas * c;
            const double eff = static_cast<double>(total) / static_cast<double>(ceil_div(total, slots) * slots);
            if (eff > best + 0.02) {
                best = eff;
                js = c;
            }
        }
        p.jsplit = static_cast<int>(js_cap < 0 ? (-js_cap < num_jt ? -js_cap : num_jt) : js);
    } else {
        p.row_parts = ceil_div(N, SIMT_T);
        p.col_parts = ceil_div(n, SIMT_T);
        // CUDA-core sweep: 32-row blocks; split the columns until about two waves of blocks exist (small batches)
        const int64_t blocks = ceil_div(n, SIMT_BR) * ceil_div(d, 768);
        int64_t js = ceil_div(2 * 148, blocks > 0 ? blocks : 1);
        const int64_t steps = ceil_div(N, 32);
        if (js > 8) js = 8;
        if (js > steps) js = steps;
        p.jsplit = static_cast<int>(js < 1 ? 1 : js);
    }
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes);
        return o;
    };
    p.off_rep = take(sizeof(int32_t) * N);
    p.off_cnt = take(sizeof(float) * N);
    p.off_gscale = take(sizeof(float) * 4);
    p.off_scale = take(sizeof(float) * 4);
    for (int m = 0; m < 3; ++m) {
        p.off_Q[m] = take(sizeof(float) * N * d);
        p.off_dxh[m] = take(sizeof(float) * p.jsplit * n * d);
        if (tc) {
            p.off_xh[m] = take(2 * static_cast<size_t>(N) * p.dpad);
            p.off_xhS[m] = take(2 * static_cast<size_t>(N) * p.dpad);
            p.off_xhT[m] = take(2 * static_cast<size_t>(p.dpad) * p.npad);
        }
    }
    p.off_u = take(sizeof(float) * 3 * N);
    p.off_v = take(sizeof(float) * 3 * N);
    // per-tile partial sums of the forward: one set per pair (so that ONE launch reduces all of them after the three
    // forward kernels) while the three sets stay below 256 MB, a single shared set beyond that (N >= ~100k)
    p.rowpart_elems = static_cast<size_t>(p.row_parts) * n;
    p.colpart_elems = static_cast<size_t>(p.col_parts) * N;
    p.part_sets = (3 * sizeof(float) * (p.rowpart_elems + p.colpart_elems) <= (size_t(256) << 20)) ? 3 : 1;
    p.off_rowpart = take(sizeof(float) * p.rowpart_elems * p.part_sets);
    p.off_colpart = take(sizeof(float) * p.colpart_elems * p.part_sets);
    p.off_posrow = take(sizeof(float) * n);
    p.off_cstart = take(sizeof(int32_t) * N);
    p.off_class_lo = take(sizeof(int32_t) * N);
    p.off_ccS = take(sizeof(float) * 3 * N);  // one per pair (S-once backward: all pairs prepared in one launch)
    p.off_posrow2 = take(sizeof(float) * 6 * n);
    // exchange mode: lam2 of ALL rows of every pair's row modality (the weighted class sums Qw need them)
    const bool want_exchange = mode == LOSS_MODE_EXCHANGE && n < N;
    p.off_lam2 = take(sizeof(float) * 6 * (want_exchange ? N : n));
    // Single-GPU tcgen05 pair path: S is computed once per pair; the row sweep stores its 16-bit coefficients
    // transposed in a strip buffer (CLIBD_GT_STRIP_MB, default 2304 MB = the whole column range at
    // N = 32768, and never fewer than 18944 columns) and the other side's gradient is a plain GEMM over that strip.
    // (used from N = 4096 up: measured 0.53 vs 0.61 ms per step at N = 4096, three modalities, tools/small_batch_probe.py;
    //  at N = 2048 the step is bound by the host's launch rate either way)
    const int64_t shared_min_n = knobs.shared_min_n;
    // Row-sharded exchange mode: the same S-once backward on the local rows of every pair's row modality; the caller
    // asked for it explicitly, so only the shape limits of the pair kernels apply.
    p.exchange = allow_shared_s && tc && want_exchange && p.dpad <= PAIR_DCH;
    p.shared_s = p.exchange || (allow_shared_s && tc && n == N && N >= shared_min_n && p.dpad <= PAIR_DCH &&
                                !knobs.two_sweeps && !knobs.bwd_single);
    if (p.shared_s) {
        const double budget = knobs.strip_mb * 1048576.0;
        int64_t rows = static_cast<int64_t>(budget / (2.0 * static_cast<double>(n))) / 256 * 256;
        if (rows < 74 * 256) rows = 74 * 256;  // at least one 256-row tile per CTA pair and feature tile
        const int64_t all = round_up(N, 256);
        p.strip_rows = rows < all ? rows : all;
        p.gt_ld = p.strip_rows;  // Gs[n local rows][strip columns]
        p.off_gt = take(2 * static_cast<size_t>(n) * p.gt_ld);
        // sharded step: (image, text) and (dna, text) feed ONE gradient GEMM, each from its own strip
        p.off_gt2 = p.exchange ? take(2 * static_cast<size_t>(n) * p.gt_ld) : p.off_gt;
        // ... and a third one, so that the gradient GEMM of (image, dna) can still read its strip while the sweeps of
        // the two text pairs fill theirs (shared_s_sweeps: GEMM on the side stream next to the following sweeps)
        p.off_gt3 = p.exchange ? take(2 * static_cast<size_t>(n) * p.gt_ld) : p.off_gt;
        p.npad_loc = round_up(n, 8);
        for (int m = 0; m < 3; ++m) {
            p.off_xhTo[m] = take(2 * static_cast<size_t>(p.dpad) * p.npad_loc);
            p.off_Qw[m] = take(sizeof(float) * N * d);
        }
    }
    p.off_dots = take(sizeof(float) * 3 * n);
    p.off_red = take(sizeof(double) * 256 * 8);  // 256 block partials per reduction job
    const int64_t H = label_hash_slots(N);
    p.off_hown = take(sizeof(int32_t) * H);
    p.off_hmin = take(sizeof(int32_t) * H);
    p.off_hcnt = take(sizeof(int32_t) * H);
    p.off_skey = take(sizeof(int32_t) * N);
    p.off_sidx = take(sizeof(int32_t) * N);
    p.off_iota = take(sizeof(int32_t) * N);
    p.sort_tmp_bytes = class_sort_temp_bytes(N);
    p.off_sorttmp = take(p.sort_tmp_bytes);
    p.total = off;
    return p;
}

static const int kPairA[3] = {0, 0, 1};
static const int kPairB[3] = {1, 2, 2};

static int check_common(const void* const x[3], const float* const inv_norm[3], const float pair_weight[3],
                        int64_t N, int64_t d, int64_t row0, int64_t n, int dtype, int path, void* scratch,
                        int64_t scratch_bytes, const LossPlan& plan) {
    CLIBD_REQUIRE(N > 0 && d > 0 && n >= 0 && row0 >= 0 && row0 + n <= N, "bad row range");
    CLIBD_REQUIRE(N < (int64_t(1) << 31) - 512, "n_global too large for 32-bit TMA coordinates");
    CLIBD_REQUIRE(dtype == DT_F32 || dtype == DT_BF16 || dtype == DT_F16, "dtype must be 0, 1 or 2");
    CLIBD_REQUIRE(path >= 0 && path <= 2, "path must be 0, 1 or 2");
    CLIBD_REQUIRE(scratch != nullptr && scratch_bytes >= static_cast<int64_t>(plan.total), "scratch too small");
    for (int p = 0; p < 3; ++p) {
        if (pair_weight[p] != 0.f) {
            CLIBD_REQUIRE(x[kPairA[p]] && x[kPairB[p]] && inv_norm[kPairA[p]] && inv_norm[kPairB[p]],
                          "a weighted pair references an absent modality");
        }
    }
    return 0;
}


// Backward with S computed ONCE per modality pair (plan.shared_s).  For pair p = (a, b), on the LOCAL rows of a
// (all rows on one GPU):
//   1. row sweep (loss_bwd_pair.cu): S tile -> G~ (minus lam2_i on the positives) -> dxh[a] += G~ Yhat_b, and the
//      same 16-bit G~ tiles stored into the strip buffer [n local rows, strip columns];
//   2. the column side's gradient Gs^T * Xhat_a as a plain GEMM over the strip (loss_grad_gemm.cu) -- no second S^T
//      sweep.  One GPU: it lands in dxh[b].  Row-sharded (exchange mode): it is this rank's PARTIAL gradient of all N
//      rows of b and leaves the rank -- into part[b] (the caller reduce-scatters it) or straight into the owners'
//      peer-mapped slot arrays (ExchangeArgs);
//   3. the fp32 target term of b's rows uses class sums of a weighted by 1 - lam2_i / 2 (the share the sweep
//      has not subtracted), the one of a's rows the usual (2 - lam2_i) Q_b[rep_i].
// Tensor work per pair: 2 (S) + 2 + 2 = 6 n N d flops instead of 8 for two sweeps.
struct ExchangeArgs {
    const float* posrow = nullptr;     // [3][N] complete per-row positive dot products (exchange mode)
    float* const* part = nullptr;      // [3]: local [N, d] partial-gradient buffers, or null
    float* const* peer_red = nullptr;  // [world * 3]: entry [q * 3 + p] = rank q's slot array of pair p, or null
    int rank = 0, world = 1;
};

// What normalize_bwd adds up for modality m (fixed by the pair weights alone, so the two halves of a split backward
// derive the same bookkeeping): per pair containing m the partner's class sums, their weight and the lam2 the sweep
// already subtracted on the row's positives; whether row sweeps wrote into dxh[m].
struct NormPlan {
    const float* Qp[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    const float* lam[3][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
    float wp[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    int nparts[3] = {0, 0, 0};
    bool wrote_dxh[3] = {false, false, false};
};

static float* lam2_of_pair(void* scratch, const LossPlan& plan, int p) {
    // one GPU: [n = N] per (pair, direction) slot; exchange mode: all N rows (the slots are N wide there too)
    return at<float>(scratch, plan.off_lam2) + static_cast<int64_t>(2 * p) * plan.N;
}

static NormPlan shared_s_norm_plan(void* scratch, const LossPlan& plan, const float pair_weight[3], int64_t row0) {
    NormPlan np;
    for (int p = 0; p < 3; ++p) {
        if (pair_weight[p] == 0.f) continue;
        const int a = kPairA[p], b = kPairB[p];
        np.Qp[a][np.nparts[a]] = at<float>(scratch, plan.off_Q[b]);
        np.lam[a][np.nparts[a]] = lam2_of_pair(scratch, plan, p) + row0;  // indexed by local row
        np.wp[a][np.nparts[a]] = pair_weight[p];
        ++np.nparts[a];
        np.wrote_dxh[a] = true;
        np.Qp[b][np.nparts[b]] = at<float>(scratch, plan.off_Qw[p]);
        np.lam[b][np.nparts[b]] = nullptr;
        np.wp[b][np.nparts[b]] = pair_weight[p];
        ++np.nparts[b];
        if (!plan.exchange) np.wrote_dxh[b] = true;  // one GPU: the gradient GEMM writes into dxh[b] as well
    }
    return np;
}

static int shared_s_sweeps(const void* const x[3], int dtype, const float* const inv_norm[3], int64_t N, int64_t d,
                           int64_t row0, int64_t n, float logit_scale, const float pair_weight[3], int path,
                           void* scratch, const LossPlan& plan, const ExchangeArgs& ex, cudaStream_t stream) {
    const int fmt_bf16 = path == PATH_TC_BF16 ? 1 : 0;
    const float* gscale = at<float>(scratch, plan.off_gscale);
    const float* u = at<float>(scratch, plan.off_u);
    const float* v = at<float>(scratch, plan.off_v);
    const float* cnt = at<float>(scratch, plan.off_cnt);
    const int32_t* skey = at<int32_t>(scratch, plan.off_skey);
    const int32_t* sidx = at<int32_t>(scratch, plan.off_sidx);
    const int32_t* class_lo = at<int32_t>(scratch, plan.off_class_lo);
    void* gt = at<void>(scratch, plan.off_gt);
    int rc = 0;
    SideStream* side = nullptr;
    bool wrote_dxh[3] = {false, false, false};
    bool wrote_part[3] = {false, false, false};
    // Per pair: the column coefficients in class-sorted order (ccS) and lam2 of every row the weighted class sums Qw
    // touch -- all N rows (exchange mode: from the exchanged posrow) -- then Qw itself: rows of b see the lam2-weighted
    // class sums of a (classes with a local member only).  One launch each for all pairs; the lam2-weighted sums of
    // the image rows for (image, dna) and (image, text) share one pass over those rows.
    {
        SweepPrepJob sp[3];
        ClassSumJob cj[3];
        int nsp = 0, ncj = 0;
        int job_of_mod[3] = {-1, -1, -1};
        for (int p = 0; p < 3; ++p) {
            if (pair_weight[p] == 0.f) continue;
            const int a = kPairA[p];
            float* lam2 = lam2_of_pair(scratch, plan, p);
            sp[nsp].rowcoef = u + p * N;
            sp[nsp].colcoef = v + p * N;
            sp[nsp].posrow = plan.exchange ? ex.posrow + static_cast<int64_t>(p) * N
                                           : at<float>(scratch, plan.off_posrow2) + static_cast<int64_t>(2 * p) * N;
            sp[nsp].ccS = at<float>(scratch, plan.off_ccS) + static_cast<int64_t>(p) * N;
            sp[nsp].lam2 = lam2;
            ++nsp;
            int j = job_of_mod[a];
            if (j < 0 || cj[j].out[1] != nullptr) {
                j = ncj++;
                job_of_mod[a] = j;
                cj[j].x = x[a];
                cj[j].inv = inv_norm[a];
                cj[j].lam2[0] = lam2;
                cj[j].out[0] = at<float>(scratch, plan.off_Qw[p]);
            } else {
                cj[j].lam2[1] = lam2;
                cj[j].out[1] = at<float>(scratch, plan.off_Qw[p]);
            }
        }
        if ((rc = launch_sweep_prep_jobs(sp, nsp, sidx, cnt, N, 0, N, logit_scale, stream))) return rc;
        // the weighted class sums are read by the normalise-backward pass only: they run on the side stream next to the
        // sweeps (a sharded sweep leaves SMs idle) and the gradient GEMMs, and are joined at the end of this function
        side = side_stream(1);
        if (side != nullptr && (rc = side_fork(side, stream))) return rc;
        if ((rc = launch_class_sums_jobs(cj, ncj, dtype, skey, sidx, cnt, N, d, row0, n, side ? side->stream : stream)))
            return rc;
    }
    // Pairs are processed in groups that share their column modality b: (image, dna) -> dna; (image, text) and
    // (dna, text) -> text.  Every pair of a group sweeps its rows into its own coefficient strip, then ONE gradient
    // GEMM contracts the concatenated strips (equal pair weights; unequal ones fall back to one GEMM per pair).
    int group_pairs[3][2];
    int group_size[3] = {0, 0, 0};
    int ngroups = 0;
    for (int p = 0; p < 3; ++p) {
        if (pair_weight[p] == 0.f) continue;
        // Merged only in the sharded step, where K = n local rows per pair is short and the merge saves a trip over
        // NVLink.  On one GPU a K range of 2 N rows lets the three pairs of CTAs that share a strip block (one per
        // 256-column feature tile) drift apart until L2 no longer serves the re-reads: ncu showed 12.1 GB of DRAM reads
        // for the merged launch against 2 x 3.7 GB for two launches, and 2.70 ms against 2 x 1.21 ms.
        int g = -1;
        for (int k = 0; k < ngroups && plan.exchange; ++k)
            if (group_size[k] == 1 && kPairB[group_pairs[k][0]] == kPairB[p] && pair_weight[group_pairs[k][0]] == pair_weight[p])
                g = k;
        if (g < 0) g = ngroups++;
        group_pairs[g][group_size[g]++] = p;
    }
    // Sharded step: a rank's sweep runs 2 x n / 128 x jsplit CTAs -- 128 of 148 SMs at n = 4096 -- and the gradient GEMM
    // of a group only needs that group's strips.  With one strip per pair (the whole column range fits) the GEMM of every
    // group but the last therefore goes to the side stream, where its CTA pairs fill the SMs the following sweeps leave
    // idle and its epilogue's NVLink stores spread over a longer time; the groups then use disjoint strip buffers.
    // CLIBD_OVERLAP_GEMM=0 keeps everything in line.
    void* bufs[3] = {gt, at<void>(scratch, plan.off_gt2), at<void>(scratch, plan.off_gt3)};
    const char* ov_env = std::getenv("CLIBD_OVERLAP_GEMM");
    const bool overlap_gemm = plan.exchange && side != nullptr && ngroups > 1 && plan.strip_rows >= N &&
                              !(ov_env != nullptr && std::atoi(ov_env) == 0);
    int buf0 = 0;
    for (int g = 0; g < ngroups; ++g) {
        void* gts[2] = {bufs[overlap_gemm ? buf0 : 0], bufs[overlap_gemm ? (buf0 + 1 < 3 ? buf0 + 1 : 2) : 1]};  // [1]: merged groups only
        buf0 += group_size[g];
        cudaStream_t gemm_stream = stream;
        const int p0 = group_pairs[g][0];
        const int b = kPairB[p0];
        GradDest dest;
        for (int q = 0; q < MAX_PEERS; ++q) dest.base[q] = nullptr;
        int ksplit = 1, acc_grad = 0;
        if (!plan.exchange) {  // one GPU: into dxh[b] (K split over the jsplit partial outputs)
            dest.base[0] = at<float>(scratch, plan.off_dxh[b]);
            dest.rows_per_dest = N;
            dest.slot_rows = N;
            dest.slot0 = 0;
            ksplit = plan.jsplit;
            acc_grad = wrote_dxh[b] ? 1 : 0;
        } else if (ex.peer_red != nullptr) {  // owners' slot arrays over NVLink, slot = this rank
            CLIBD_REQUIRE(!wrote_part[b], "the peer form needs equal weights for the pairs that share a column modality");
            for (int q = 0; q < ex.world; ++q) dest.base[q] = ex.peer_red[q * 3 + p0];
            dest.rows_per_dest = n;
            dest.slot_rows = n;
            dest.slot0 = ex.rank;
        } else {  // local partial buffer, reduce-scattered by the caller
            dest.base[0] = ex.part[b];
            dest.rows_per_dest = N;
            dest.slot_rows = N;
            dest.slot0 = 0;
            acc_grad = wrote_part[b] ? 1 : 0;
        }
        int strip = 0;
        for (int64_t c0 = 0; c0 < N; c0 += plan.strip_rows, ++strip) {
            const int64_t c1 = c0 + plan.strip_rows < N ? c0 + plan.strip_rows : N;
            GradPart parts[2];
            for (int k = 0; k < group_size[g]; ++k) {
                const int p = group_pairs[g][k];
                const int a = kPairA[p];
                const float* lam2 = lam2_of_pair(scratch, plan, p);
                const float* ccS = at<float>(scratch, plan.off_ccS) + static_cast<int64_t>(p) * N;
                if ((rc = tc_backward_rows_pair(at<void>(scratch, plan.off_xh[a]), at<void>(scratch, plan.off_xhS[b]),
                                                at<void>(scratch, plan.off_xhT[b]), N, plan.npad, d, plan.dpad, row0, n,
                                                logit_scale, u + p * N, ccS, gscale, pair_weight[p], wrote_dxh[a] || strip > 0,
                                                plan.jsplit, fmt_bf16, at<float>(scratch, plan.off_dxh[a]), stream,
                                                /*self_mask=*/0, class_lo, cnt, lam2 + row0, c0, c1, gts[k], plan.gt_ld)))
                    return rc;
                parts[k].gs = gts[k];
                parts[k].xhT_x = at<void>(scratch, plan.off_xhTo[a]);
            }
            if (overlap_gemm && g + 1 < ngroups) {
                if ((rc = side_fork(side, stream))) return rc;  // after this group's sweeps; joined at the end
                gemm_stream = side->stream;
            } else if (overlap_gemm && acc_grad) {
                // accumulates onto what an earlier group's GEMM (possibly still running on the side stream) wrote
                if ((rc = side_join(side, stream))) return rc;
            }
            if ((rc = tc_grad_from_strip(parts, group_size[g], plan.gt_ld, c1 - c0, c0, n, plan.npad_loc, N, d, plan.dpad, sidx,
                                         gscale, pair_weight[p0], acc_grad, ksplit, fmt_bf16, dest, tc_num_sms(),
                                         gemm_stream)))
                return rc;
        }
        for (int k = 0; k < group_size[g]; ++k) wrote_dxh[kPairA[group_pairs[g][k]]] = true;
        if (!plan.exchange) wrote_dxh[b] = true;
        else wrote_part[b] = true;
    }
    if (side != nullptr && (rc = side_join(side, stream))) return rc;
    return 0;
}

static int shared_s_finish(const void* const x[3], int dtype, const float* const inv_norm[3], int64_t N, int64_t d,
                           int64_t row0, int64_t n, float logit_scale, const float pair_weight[3], void* scratch,
                           const LossPlan& plan, const float* const reduced[3], const int reduced_slots[3],
                           float grad_feat_scale, const float* grad_feat_scale_dev, int grad_count, void* const dx[3],
                           double* dscale_partial, cudaStream_t stream) {
    const NormPlan np = shared_s_norm_plan(scratch, plan, pair_weight, row0);
    const int32_t* rep = at<int32_t>(scratch, plan.off_rep);
    float* dots = at<float>(scratch, plan.off_dots);
    double* red = at<double>(scratch, plan.off_red);
    int rc = 0;
    int n_mod_used = 0;
    NormBwdArgs all[3];
    for (int m = 0; m < 3; ++m) {
        if (np.nparts[m] == 0) continue;
        NormBwdArgs& a = all[n_mod_used];
        a.x = x[m];
        a.dtype = dtype;
        a.inv_norm = inv_norm[m];
        a.rep = rep;
        a.dxh = at<float>(scratch, plan.off_dxh[m]);
        a.jsplit = np.wrote_dxh[m] ? plan.jsplit : 0;
        if (reduced != nullptr && reduced[m] != nullptr && reduced_slots[m] > 0) {
            a.extra = reduced[m];
            a.extra_slots = reduced_slots[m];
        }
        for (int k = 0; k < 2; ++k) {
            a.Qp[k] = np.Qp[m][k];
            a.wp[k] = np.wp[m][k];
            a.lam2[k] = np.lam[m][k];
        }
        a.N = N;
        a.d = d;
        a.row0 = row0;
        a.n = n;
        a.scale = logit_scale;
        a.grad_scale = grad_feat_scale;
        a.grad_scale_dev = grad_feat_scale_dev;
        a.grad_scale_dev_count = grad_count;
        a.dx = dx ? dx[m] : nullptr;
        a.dots = dots + n_mod_used * n;
        ++n_mod_used;
    }
    if ((rc = launch_normalize_bwd_multi(all, n_mod_used, stream))) return rc;
    return launch_sum_to_double(dots, static_cast<int64_t>(n_mod_used) * n, 0.5, red, dscale_partial, stream,
                                scale_dev_ptr());
}

}  // namespace clibd

using namespace clibd;

extern "C" {

int clibd_abi_version(void) { return CLIBD_ABI_VERSION; }

const char* clibd_last_error(void) { return last_error().c_str(); }

int64_t clibd_kernel_launch_count(void) { return g_launches.load(); }

int clibd_graphs_active(int64_t n_global, int64_t n_local) {
    const char* e = std::getenv("CLIBD_GRAPHS");
    return (e == nullptr || std::atoi(e) != 0) && graph_worthwhile(n_global, n_local) ? 1 : 0;
}

int clibd_profile_enable(int enable) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof_on = enable != 0;
    return 0;
}

int clibd_profile_read(double* total_ms, int64_t* launches) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (int sl = 0; sl < PROF_SLOTS; ++sl) {
        double ms = 0.0;
        for (auto& ev : g_prof_events[sl]) {
            cudaEventSynchronize(ev.second);
            float t = 0.f;
            cudaEventElapsedTime(&t, ev.first, ev.second);
            ms += t;
            cudaEventDestroy(ev.first);
            cudaEventDestroy(ev.second);
        }
        total_ms[sl] = ms;
        launches[sl] = static_cast<int64_t>(g_prof_events[sl].size());
        g_prof_events[sl].clear();
    }
    return 0;
}

int clibd_device_supported(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    return major == 10 ? 1 : 0;
}

int clibd_row_inv_norm(const void* x, int dtype, int64_t n, int64_t d, float* inv_norm, clibd_stream_t stream) {
    CLIBD_REQUIRE(x && inv_norm && n >= 0 && d > 0, "null pointer or bad shape");
    return launch_row_inv_norm(x, dtype, n, d, inv_norm, stream);
}

int64_t clibd_loss_scratch_bytes(int64_t n_global, int64_t n_local, int64_t d, int path, int mode) {
    if (n_global <= 0 || n_local < 0 || d <= 0 || path < 0 || path > 2 || mode < 0 || mode > 1) return -1;
    return static_cast<int64_t>(make_loss_plan(n_global, n_local, d, path, true, mode).total);
}

static int loss_forward_stats_impl(const void* const x[3], int dtype, const float* const inv_norm[3],
                             const int64_t* labels, int64_t N, int64_t d, int64_t row0, int64_t n,
                             float logit_scale, const float* logit_scale_dev, const float pair_weight[3], int path,
                             int mode, void* scratch, int64_t scratch_bytes, float* rowsum, float* colsum,
                             float* posrow_out, double* pos, clibd_stream_t stream) {
    CLIBD_REQUIRE(mode == LOSS_MODE_LOCAL || mode == LOSS_MODE_EXCHANGE, "mode must be 0 or 1");
    const PlanKnobs knobs = labels != nullptr ? plan_knobs_from_env() : plan_knobs_for(scratch);
    remember_plan_knobs(scratch, knobs);  // finish / backward of this step lay the scratch out the same way
    const LossPlan plan = make_loss_plan(N, n, d, path, true, mode, &knobs);
    int rc = check_common(x, inv_norm, pair_weight, N, d, row0, n, dtype, path, scratch, scratch_bytes, plan);
    if (rc) return rc;
    CLIBD_REQUIRE(rowsum && colsum && pos, "null output pointer");  // labels == null: clibd_loss_label_stage ran already
    if (mode == LOSS_MODE_EXCHANGE && n < N) {
        CLIBD_REQUIRE(plan.exchange, "exchange mode needs a tensor-core path and a feature dim <= 768");
        CLIBD_REQUIRE(posrow_out != nullptr, "exchange mode needs the posrow output");
    }
    const bool tc = path != PATH_SIMT_F32;
    if (tc) CLIBD_REQUIRE(clibd_device_supported(), "tcgen05 path needs a compute-capability 10.x device");
    // A host value is checked here; a device scalar is checked by scale_set_kernel (NaN loss when not positive).
    if (logit_scale_dev == nullptr)
        CLIBD_REQUIRE(logit_scale > 0.f && logit_scale < 1e30f, "logit_scale must be a positive finite number");
    float* scale_cell = at<float>(scratch, plan.off_scale);
    scale_set_kernel<<<1, 1, 0, stream>>>(logit_scale, logit_scale_dev, scale_cell);
    CLIBD_KERNEL_CHECK();
    ScaleScope scale_scope(scale_cell);  // every kernel below (and in finish / backward) reads the scale from there
    const int fmt_bf16 = path == PATH_TC_BF16 ? 1 : 0;
    int32_t* rep = at<int32_t>(scratch, plan.off_rep);
    float* cnt = at<float>(scratch, plan.off_cnt);
    float* gscale = at<float>(scratch, plan.off_gscale);
    LabelScratch ls;
    ls.own = at<int32_t>(scratch, plan.off_hown);
    ls.hmin = at<int32_t>(scratch, plan.off_hmin);
    ls.hcnt = at<int32_t>(scratch, plan.off_hcnt);
    ls.skey = at<int32_t>(scratch, plan.off_skey);
    ls.sidx = at<int32_t>(scratch, plan.off_sidx);
    ls.iota = at<int32_t>(scratch, plan.off_iota);
    ls.sort_tmp = at<void>(scratch, plan.off_sorttmp);
    ls.sort_tmp_bytes = plan.sort_tmp_bytes;
    if (labels != nullptr) {
        if ((rc = launch_label_stats(labels, N, rep, cnt, ls, stream))) return rc;
        if ((rc = launch_gscale(cnt, N, path, gscale, stream))) return rc;
        if ((rc = launch_class_ranges(ls.skey, rep, N, at<int32_t>(scratch, plan.off_cstart),
                                      at<int32_t>(scratch, plan.off_class_lo), stream)))
            return rc;
    }
    bool used[3] = {false, false, false};
    for (int p = 0; p < 3; ++p)
        if (pair_weight[p] != 0.f) used[kPairA[p]] = used[kPairB[p]] = true;
    // Which 16-bit operand copies a modality needs.  Row operand (xh, input order): only the LOCAL rows are ever read
    // (forward tiles and row sweeps of this rank's block).  Column operands (xhS class-sorted, xhT its transpose): all N
    // rows, for modalities that appear as the column side -- every used modality when both directions are swept, only
    // the pairs' second modality when S is computed once per pair (then image is never a column operand and text never
    // a row operand: 400 MB of staging writes instead of 600 MB at N = 32768).  xhTo (transposed local rows in input
    // order) is the K operand of the stored-coefficient GEMM.
    bool row_op[3] = {false, false, false}, col_op[3] = {false, false, false};
    for (int p = 0; p < 3; ++p) {
        if (pair_weight[p] == 0.f) continue;
        row_op[kPairA[p]] = col_op[kPairB[p]] = true;
        if (!plan.shared_s) row_op[kPairB[p]] = col_op[kPairA[p]] = true;
    }
    // class sums Q[m] feed the positive term of the partner's rows: every pair's second modality, and the first one too
    // when both directions are swept (S once per pair: the column side's target term uses the lam2-weighted Qw)
    ClassSumJob cj[3];
    int ncj = 0;
    for (int m = 0; m < 3; ++m) {
        if (!used[m] || !col_op[m]) continue;
        cj[ncj].x = x[m];
        cj[ncj].inv = inv_norm[m];
        cj[ncj].out[0] = at<float>(scratch, plan.off_Q[m]);
        ++ncj;
    }
    // staging tasks: (modality, column operands) and (modality, row operand of the local rows)
    auto stage_col = [&](int m, cudaStream_t st) -> int {
        return launch_make_operands(x[m], dtype, inv_norm[m], N, d, plan.dpad, plan.npad, fmt_bf16, nullptr,
                                    at<void>(scratch, plan.off_xhT[m]), st, ls.sidx, at<void>(scratch, plan.off_xhS[m]));
    };
    auto stage_row = [&](int m, cudaStream_t st) -> int {
        const size_t esize = dtype == DT_F32 ? 4 : 2;
        const void* xl = static_cast<const char*>(x[m]) + static_cast<size_t>(row0) * d * esize;
        void* xh_loc = at<char>(scratch, plan.off_xh[m]) + static_cast<size_t>(row0) * plan.dpad * 2;
        return launch_make_operands(xl, dtype, inv_norm[m] + row0, n, d, plan.dpad, plan.npad_loc, fmt_bf16, xh_loc,
                                    plan.shared_s ? at<void>(scratch, plan.off_xhTo[m]) : nullptr, st);
    };
    // The first weighted pair's operands are staged on the caller's stream and its forward kernel starts; everything
    // else -- the class sums and the other modalities' operand copies -- runs on the side stream next to that kernel
    // and is joined before the second pair.
    int first_pair = -1;
    for (int p = 0; p < 3 && first_pair < 0; ++p)
        if (pair_weight[p] != 0.f) first_pair = p;
    SideStream* side = (tc && first_pair >= 0) ? side_stream(0) : nullptr;
    bool side_pending = false;
    if (tc) {
        bool done_row[3] = {false, false, false}, done_col[3] = {false, false, false};
        if (side != nullptr) {
            const int a0 = kPairA[first_pair], b0 = kPairB[first_pair];
            if ((rc = side_fork(side, stream))) return rc;
            if ((rc = stage_row(a0, stream))) return rc;
            if ((rc = stage_col(b0, stream))) return rc;
            done_row[a0] = done_col[b0] = true;
            side_pending = true;
        }
        cudaStream_t st2 = side ? side->stream : stream;
        if ((rc = launch_class_sums_jobs(cj, ncj, dtype, ls.skey, ls.sidx, cnt, N, d, row0, n, st2))) return rc;
        for (int m = 0; m < 3; ++m) {
            if (!used[m]) continue;
            if (col_op[m] && !done_col[m] && (rc = stage_col(m, st2))) return rc;
            if (row_op[m] && !done_row[m] && (rc = stage_row(m, st2))) return rc;
        }
    } else {
        if ((rc = launch_class_sums_jobs(cj, ncj, dtype, ls.skey, ls.sidx, cnt, N, d, row0, n, stream))) return rc;
    }
    float* posrow2 = at<float>(scratch, plan.off_posrow2);
    double* red = at<double>(scratch, plan.off_red);
    // The three forward kernels run back to back; their per-tile partial sums are reduced, the per-row positive dot
    // products formed and summed afterwards, each in ONE launch for all pairs (one set of partial buffers per pair).
    ReduceJob rj[6];
    PosRowsJob pj[6];
    SumJob sj[3];
    int nrj = 0, npj = 0, nsj = 0;
    auto flush_reduce = [&]() -> int {
        const int r = launch_reduce_parts_jobs(rj, nrj, stream);
        nrj = 0;
        return r;
    };
    for (int p = 0; p < 3; ++p) {
        if (pair_weight[p] == 0.f) {
            CLIBD_CHECK_CUDA(cudaMemsetAsync(pos + p, 0, sizeof(double), stream));
            continue;
        }
        const int a = kPairA[p], b = kPairB[p];
        const int set = plan.part_sets == 3 ? p : 0;
        float* rowpart = at<float>(scratch, plan.off_rowpart) + set * plan.rowpart_elems;
        float* colpart = at<float>(scratch, plan.off_colpart) + set * plan.colpart_elems;
        if (tc) {
            rc = tc_forward_pair(at<void>(scratch, plan.off_xh[a]), at<void>(scratch, plan.off_xhS[b]), N, plan.dpad, row0,
                                 n, logit_scale, fmt_bf16, rowpart, colpart, stream);
        } else {
            rc = simt_forward_pair(x[a], x[b], dtype, inv_norm[a], inv_norm[b], N, d, row0, n, logit_scale, rowpart,
                                   colpart, stream);
        }
        if (rc) return rc;
        if (side_pending) {  // the other pairs' operands and the class sums must be there from here on
            if ((rc = side_join(side, stream))) return rc;
            side_pending = false;
        }
        rj[nrj].part = rowpart;
        rj[nrj].parts = plan.row_parts;
        rj[nrj].stride = n;
        rj[nrj].len = n;
        rj[nrj].out = rowsum + p * N + row0;
        ++nrj;
        // the tcgen05 column sums come out in class-sorted column order: scatter them back to input order
        rj[nrj].part = colpart;
        rj[nrj].parts = plan.col_parts;
        rj[nrj].stride = N;
        rj[nrj].len = N;
        rj[nrj].out = colsum + p * N;
        rj[nrj].scatter = tc ? ls.sidx : nullptr;
        ++nrj;
        if (plan.part_sets == 1 && (rc = flush_reduce())) return rc;  // the next pair reuses the partial buffers
        // per-row positive dot products of both directions (direction 0 also gives the loss's positive term).
        // Exchange mode: direction 0 goes to the caller's [3, N] buffer at the local rows (the backward needs it for
        // all rows).  S once per pair: direction 1 is not used (no transposed sweep).
        float* posrow = plan.exchange ? posrow_out + static_cast<int64_t>(p) * N + row0 : posrow2 + (2 * p) * n;
        pj[npj].xa = x[a];
        pj[npj].inv_a = inv_norm[a];
        pj[npj].Qb = at<float>(scratch, plan.off_Q[b]);
        pj[npj].posrow = posrow;
        ++npj;
        if (!plan.shared_s) {
            pj[npj].xa = x[b];
            pj[npj].inv_a = inv_norm[b];
            pj[npj].Qb = at<float>(scratch, plan.off_Q[a]);
            pj[npj].posrow = posrow + n;
            ++npj;
        }
        sj[nsj].in = posrow;
        sj[nsj].len = n;
        sj[nsj].out = pos + p;
        ++nsj;
    }
    if ((rc = flush_reduce())) return rc;
    if ((rc = launch_pos_rows_jobs(pj, npj, dtype, rep, d, row0, n, stream))) return rc;
    return launch_sum_to_double_jobs(sj, nsj, 1.0, red, stream);
}

static int loss_forward_finish_impl(int64_t N, int64_t n, int64_t d, float logit_scale, const float pair_weight[3],
                              int path, int mode, void* scratch, int64_t scratch_bytes, const float* rowsum,
                              const float* colsum, const double* pos, float* loss_out, clibd_stream_t stream) {
    CLIBD_REQUIRE(N > 0 && d > 0 && path >= 0 && path <= 2 && mode >= 0 && mode <= 1, "bad shape");
    const PlanKnobs knobs = plan_knobs_for(scratch);
    const LossPlan plan = make_loss_plan(N, n, d, path, true, mode, &knobs);
    CLIBD_REQUIRE(scratch && scratch_bytes >= static_cast<int64_t>(plan.total), "scratch too small");
    CLIBD_REQUIRE(rowsum && colsum && pos && loss_out, "null pointer");
    ScaleScope scale_scope(at<float>(scratch, plan.off_scale));  // written by clibd_loss_forward_stats
    return launch_loss_finish(N, logit_scale, pair_weight, at<float>(scratch, plan.off_cnt), rowsum, colsum, pos,
                              at<float>(scratch, plan.off_u), at<float>(scratch, plan.off_v),
                              at<double>(scratch, plan.off_red), loss_out, stream);
}

static int loss_backward_impl(const void* const x[3], int dtype, const float* const inv_norm[3], int64_t N, int64_t d,
                        int64_t row0, int64_t n, float logit_scale, const float pair_weight[3], int path,
                        void* scratch, int64_t scratch_bytes, float grad_feat_scale, const float* grad_feat_scale_dev,
                        void* const dx[3],
                        double* dscale_partial, clibd_stream_t stream) {
    const PlanKnobs knobs = plan_knobs_for(scratch);
    const LossPlan plan = make_loss_plan(N, n, d, path, true, LOSS_MODE_LOCAL, &knobs);
    int rc = check_common(x, inv_norm, pair_weight, N, d, row0, n, dtype, path, scratch, scratch_bytes, plan);
    if (rc) return rc;
    CLIBD_REQUIRE(dscale_partial != nullptr, "null dscale_partial");
    ScaleScope scale_scope(at<float>(scratch, plan.off_scale));  // written by clibd_loss_forward_stats
    if (plan.shared_s) {  // one GPU (n == N): S once per pair, both gradients stay in the scratch
        if ((rc = shared_s_sweeps(x, dtype, inv_norm, N, d, row0, n, logit_scale, pair_weight, path, scratch, plan,
                                  ExchangeArgs(), stream)))
            return rc;
        return shared_s_finish(x, dtype, inv_norm, N, d, row0, n, logit_scale, pair_weight, scratch, plan, nullptr, nullptr,
                               grad_feat_scale, grad_feat_scale_dev, 1, dx, dscale_partial, stream);
    }
    const bool tc = path != PATH_SIMT_F32;
    const int fmt_bf16 = path == PATH_TC_BF16 ? 1 : 0;
    const int32_t* rep = at<int32_t>(scratch, plan.off_rep);
    const float* gscale = at<float>(scratch, plan.off_gscale);
    const float* u = at<float>(scratch, plan.off_u);
    const float* v = at<float>(scratch, plan.off_v);
    const float* cnt = at<float>(scratch, plan.off_cnt);
    const int32_t* sidx = at<int32_t>(scratch, plan.off_sidx);
    const int32_t* class_lo = at<int32_t>(scratch, plan.off_class_lo);
    float* ccS = at<float>(scratch, plan.off_ccS);
    float* dots = at<float>(scratch, plan.off_dots);
    const bool use_pair = tc && pair_backward_supported(plan.dpad) && !plan.bwd_single;
    double* red = at<double>(scratch, plan.off_red);
    int n_mod_used = 0;
    for (int m = 0; m < 3; ++m) {
        // ordered sweeps that write rows of modality m: for every weighted pair containing m
        int npart = 0;
        const float* Qp[2] = {nullptr, nullptr};
        const float* lam[2] = {nullptr, nullptr};
        float wp[2] = {0.f, 0.f};
        float* dxh = at<float>(scratch, plan.off_dxh[m]);
        for (int p = 0; p < 3; ++p) {
            if (pair_weight[p] == 0.f) continue;
            int other, dir;
            const float *rowcoef, *colcoef;
            if (kPairA[p] == m) {          // m indexes the rows of S_p
                other = kPairB[p];
                dir = 0;
                rowcoef = u + p * N;
                colcoef = v + p * N;
            } else if (kPairB[p] == m) {   // m indexes the columns of S_p: sweep S_p^T
                other = kPairA[p];
                dir = 1;
                rowcoef = v + p * N;
                colcoef = u + p * N;
            } else {
                continue;
            }
            float* lam2 = use_pair ? at<float>(scratch, plan.off_lam2) + (2 * p + dir) * n : nullptr;
            if (tc) {
                // column coefficients in the class-sorted column order; lam2 of this sweep's rows
                if ((rc = launch_sweep_prep(rowcoef, colcoef, sidx, cnt, at<float>(scratch, plan.off_posrow2) + (2 * p + dir) * n,
                                            N, row0, n, logit_scale, ccS, lam2, stream)))
                    return rc;
            }
            if (use_pair) {
                rc = tc_backward_rows_pair(at<void>(scratch, plan.off_xh[m]), at<void>(scratch, plan.off_xhS[other]),
                                           at<void>(scratch, plan.off_xhT[other]), N, plan.npad, d, plan.dpad, row0, n,
                                           logit_scale, rowcoef, ccS, gscale, pair_weight[p], npart > 0,
                                           plan.jsplit, fmt_bf16, dxh, stream, /*self_mask=*/0, class_lo, cnt, lam2);
            } else if (tc) {
                rc = tc_backward_rows(at<void>(scratch, plan.off_xh[m]), at<void>(scratch, plan.off_xhS[other]),
                                      at<void>(scratch, plan.off_xhT[other]), N, plan.npad, d, plan.dpad, row0, n,
                                      logit_scale, rowcoef, ccS, gscale, pair_weight[p], npart > 0, plan.jsplit,
                                      fmt_bf16, dxh, stream);
            } else {
                rc = simt_backward_rows(x[m], x[other], dtype, inv_norm[m], inv_norm[other], N, d, row0, n, logit_scale,
                                        rowcoef, colcoef, pair_weight[p], npart > 0, dxh, stream, /*self_mask=*/0,
                                        plan.jsplit);
            }
            if (rc) return rc;
            Qp[npart] = at<float>(scratch, plan.off_Q[other]);
            lam[npart] = lam2;
            wp[npart] = pair_weight[p];
            ++npart;
        }
        if (npart == 0) continue;
        NormBwdArgs a;
        a.x = x[m];
        a.dtype = dtype;
        a.inv_norm = inv_norm[m];
        a.rep = rep;
        a.dxh = dxh;
        a.jsplit = plan.jsplit;
        a.Qp[0] = Qp[0];
        a.Qp[1] = Qp[1];
        a.wp[0] = wp[0];
        a.wp[1] = wp[1];
        a.lam2[0] = lam[0];
        a.lam2[1] = lam[1];
        a.N = N;
        a.d = d;
        a.row0 = row0;
        a.n = n;
        a.scale = logit_scale;
        a.grad_scale = grad_feat_scale;
        a.grad_scale_dev = grad_feat_scale_dev;
        a.dx = dx ? dx[m] : nullptr;
        a.dots = dots + n_mod_used * n;
        if ((rc = launch_normalize_bwd(a, stream))) return rc;
        ++n_mod_used;
    }
    // dL/ds = (1 / (2 s)) * sum over modalities and rows of xhat_i . dxhat_i  (unit upstream grad)
    if ((rc = launch_sum_to_double(dots, static_cast<int64_t>(n_mod_used) * n, 0.5, red, dscale_partial, stream,
                                   scale_dev_ptr()))) return rc;
    return 0;
}

static int loss_backward_sweeps_impl(const void* const x[3], int dtype, const float* const inv_norm[3], int64_t N, int64_t d,
                               int64_t row0, int64_t n, float logit_scale, const float pair_weight[3], int path,
                               void* scratch, int64_t scratch_bytes, const float* posrow, float* const part[3],
                               float* const peer_red[], int rank, int world, clibd_stream_t stream) {
    const PlanKnobs knobs = plan_knobs_for(scratch);
    const LossPlan plan = make_loss_plan(N, n, d, path, true, LOSS_MODE_EXCHANGE, &knobs);
    int rc = check_common(x, inv_norm, pair_weight, N, d, row0, n, dtype, path, scratch, scratch_bytes, plan);
    if (rc) return rc;
    CLIBD_REQUIRE(plan.exchange, "clibd_loss_backward_sweeps needs a row-sharded tensor-core plan (n_local < n_global, d <= 768)");
    CLIBD_REQUIRE(posrow != nullptr, "null posrow");
    CLIBD_REQUIRE(world >= 1 && world <= MAX_PEERS && rank >= 0 && rank < world && n * world == N && row0 == rank * n,
                  "exchange mode needs equal row blocks, row0 = rank * n_local, world <= 16");
    for (int p = 0; p < 3; ++p) {
        if (pair_weight[p] == 0.f) continue;
        if (peer_red != nullptr) {
            for (int q = 0; q < world; ++q)
                CLIBD_REQUIRE(peer_red[q * 3 + p] != nullptr, "missing peer slot array of a weighted pair");
        } else {
            CLIBD_REQUIRE(part != nullptr && part[kPairB[p]] != nullptr, "missing partial-gradient buffer of a column modality");
        }
    }
    ScaleScope scale_scope(at<float>(scratch, plan.off_scale));
    ExchangeArgs ex;
    ex.posrow = posrow;
    ex.part = part;
    ex.peer_red = peer_red;
    ex.rank = rank;
    ex.world = world;
    return shared_s_sweeps(x, dtype, inv_norm, N, d, row0, n, logit_scale, pair_weight, path, scratch, plan, ex, stream);
}

static int loss_backward_finish_impl(const void* const x[3], int dtype, const float* const inv_norm[3], int64_t N, int64_t d,
                               int64_t row0, int64_t n, float logit_scale, const float pair_weight[3], int path,
                               void* scratch, int64_t scratch_bytes, const float* const reduced[3],
                               const int reduced_slots[3], float grad_feat_scale, const float* grad_feat_scale_dev,
                               int grad_count, void* const dx[3], double* dscale_partial, clibd_stream_t stream) {
    const PlanKnobs knobs = plan_knobs_for(scratch);
    const LossPlan plan = make_loss_plan(N, n, d, path, true, LOSS_MODE_EXCHANGE, &knobs);
    int rc = check_common(x, inv_norm, pair_weight, N, d, row0, n, dtype, path, scratch, scratch_bytes, plan);
    if (rc) return rc;
    CLIBD_REQUIRE(plan.exchange, "clibd_loss_backward_finish needs a row-sharded tensor-core plan");
    CLIBD_REQUIRE(dscale_partial != nullptr && reduced != nullptr && reduced_slots != nullptr && grad_count >= 1,
                  "null pointer");
    for (int p = 0; p < 3; ++p)
        if (pair_weight[p] != 0.f)
            CLIBD_REQUIRE(reduced[kPairB[p]] != nullptr && reduced_slots[kPairB[p]] > 0,
                          "missing received gradient partials of a column modality");
    ScaleScope scale_scope(at<float>(scratch, plan.off_scale));
    return shared_s_finish(x, dtype, inv_norm, N, d, row0, n, logit_scale, pair_weight, scratch, plan, reduced,
                           reduced_slots, grad_feat_scale, grad_feat_scale_dev, grad_count, dx, dscale_partial, stream);
}

// Label statistics alone (representatives, class sizes, class sort, ranges): everything the forward derives from the
// labels and nothing else.  A sharded step runs it as soon as the labels have arrived, next to the push of the feature
// rows, and then calls clibd_loss_forward_stats with labels == NULL.
static int loss_label_stage_impl(const int64_t* labels, int64_t N, int64_t n, int64_t d, int path, int mode, void* scratch,
                                 int64_t scratch_bytes, cudaStream_t stream) {
    CLIBD_REQUIRE(labels != nullptr && N > 0 && n >= 0 && d > 0 && path >= 0 && path <= 2 && mode >= 0 && mode <= 1,
                  "bad arguments");
    const PlanKnobs knobs = plan_knobs_from_env();
    remember_plan_knobs(scratch, knobs);
    const LossPlan plan = make_loss_plan(N, n, d, path, true, mode, &knobs);
    CLIBD_REQUIRE(scratch && scratch_bytes >= static_cast<int64_t>(plan.total), "scratch too small");
    int32_t* rep = at<int32_t>(scratch, plan.off_rep);
    float* cnt = at<float>(scratch, plan.off_cnt);
    LabelScratch ls;
    ls.own = at<int32_t>(scratch, plan.off_hown);
    ls.hmin = at<int32_t>(scratch, plan.off_hmin);
    ls.hcnt = at<int32_t>(scratch, plan.off_hcnt);
    ls.skey = at<int32_t>(scratch, plan.off_skey);
    ls.sidx = at<int32_t>(scratch, plan.off_sidx);
    ls.iota = at<int32_t>(scratch, plan.off_iota);
    ls.sort_tmp = at<void>(scratch, plan.off_sorttmp);
    ls.sort_tmp_bytes = plan.sort_tmp_bytes;
    int rc = 0;
    if ((rc = launch_label_stats(labels, N, rep, cnt, ls, stream))) return rc;
    if ((rc = launch_gscale(cnt, N, path, at<float>(scratch, plan.off_gscale), stream))) return rc;
    return launch_class_ranges(ls.skey, rep, N, at<int32_t>(scratch, plan.off_cstart),
                               at<int32_t>(scratch, plan.off_class_lo), stream);
}

// ---- the exported entry points: argument tuple -> CUDA-graph cache (graph_cache.h) -> the bodies above ---------------
static GraphKey base_key(int tag, const void* const x[3], int dtype, const float* const inv_norm[3], int64_t N, int64_t d,
                         int64_t row0, int64_t n, float logit_scale, const float pair_weight[3], int path, int mode,
                         const void* scratch, int64_t scratch_bytes) {
    GraphKey k;
    k.add(tag).add_array(x, 3).add(dtype).add_array(inv_norm, 3).add(N).add(d).add(row0).add(n).add(logit_scale);
    k.add_array(pair_weight, 3).add(path).add(mode).add(scratch).add(scratch_bytes);
    // environment knobs that change the plan (development switches) are part of the key through the plan itself
    const PlanKnobs knobs = tag == 1 ? plan_knobs_from_env() : plan_knobs_for(scratch);
    const LossPlan plan = make_loss_plan(N, n, d, path, true, mode, &knobs);
    k.add(plan.total).add(plan.jsplit).add(plan.shared_s).add(plan.strip_rows).add(plan.bwd_single);
    const char* side_env = std::getenv("CLIBD_SIDE_STREAM");
    k.add(side_env != nullptr && std::atoi(side_env) == 0);  // the captured sequence forks a side stream or does not
    const char* ov_env = std::getenv("CLIBD_OVERLAP_GEMM");
    k.add(ov_env != nullptr && std::atoi(ov_env) == 0);
    return k;
}

int clibd_loss_forward_stats(const void* const x[3], int dtype, const float* const inv_norm[3],
                             const int64_t* labels, int64_t N, int64_t d, int64_t row0, int64_t n,
                             float logit_scale, const float* logit_scale_dev, const float pair_weight[3], int path,
                             int mode, void* scratch, int64_t scratch_bytes, float* rowsum, float* colsum,
                             float* posrow_out, double* pos, clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_loss_forward_stats");
    CLIBD_REQUIRE(x && inv_norm && pair_weight && N > 0 && d > 0 && n >= 0, "null pointer or bad shape");
    GraphKey k = base_key(1, x, dtype, inv_norm, N, d, row0, n, logit_scale, pair_weight, path, mode, scratch, scratch_bytes);
    k.add(labels).add(logit_scale_dev).add(rowsum).add(colsum).add(posrow_out).add(pos);
    return run_graphed(k, graph_worthwhile(N, n), stream, [&](cudaStream_t s) {
        return loss_forward_stats_impl(x, dtype, inv_norm, labels, N, d, row0, n, logit_scale, logit_scale_dev, pair_weight,
                                       path, mode, scratch, scratch_bytes, rowsum, colsum, posrow_out, pos, s);
    });
}

int clibd_loss_label_stage(const int64_t* labels, int64_t N, int64_t n, int64_t d, int path, int mode, void* scratch,
                           int64_t scratch_bytes, clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_loss_label_stage");
    GraphKey k;
    k.add(6).add(labels).add(N).add(n).add(d).add(path).add(mode).add(scratch).add(scratch_bytes);
    return run_graphed(k, graph_worthwhile(N, n), stream, [&](cudaStream_t s) {
        return loss_label_stage_impl(labels, N, n, d, path, mode, scratch, scratch_bytes, s);
    });
}

int clibd_loss_forward_finish(int64_t N, int64_t n, int64_t d, float logit_scale, const float pair_weight[3],
                              int path, int mode, void* scratch, int64_t scratch_bytes, const float* rowsum,
                              const float* colsum, const double* pos, float* loss_out, clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_loss_forward_finish");
    CLIBD_REQUIRE(pair_weight && N > 0 && d > 0 && n >= 0, "null pointer or bad shape");
    GraphKey k;
    k.add(2).add(N).add(n).add(d).add(logit_scale).add_array(pair_weight, 3).add(path).add(mode).add(scratch);
    k.add(scratch_bytes).add(rowsum).add(colsum).add(pos).add(loss_out);
    return run_graphed(k, graph_worthwhile(N, n), stream, [&](cudaStream_t s) {
        return loss_forward_finish_impl(N, n, d, logit_scale, pair_weight, path, mode, scratch, scratch_bytes, rowsum, colsum,
                                        pos, loss_out, s);
    });
}

int clibd_loss_backward(const void* const x[3], int dtype, const float* const inv_norm[3], int64_t N, int64_t d,
                        int64_t row0, int64_t n, float logit_scale, const float pair_weight[3], int path,
                        void* scratch, int64_t scratch_bytes, float grad_feat_scale, const float* grad_feat_scale_dev,
                        void* const dx[3], double* dscale_partial, clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_loss_backward");
    CLIBD_REQUIRE(x && inv_norm && pair_weight && N > 0 && d > 0 && n >= 0, "null pointer or bad shape");
    GraphKey k = base_key(3, x, dtype, inv_norm, N, d, row0, n, logit_scale, pair_weight, path, 0, scratch, scratch_bytes);
    k.add(grad_feat_scale).add(grad_feat_scale_dev).add_array(dx, dx ? 3 : 0).add(dscale_partial);
    return run_graphed(k, graph_worthwhile(N, n), stream, [&](cudaStream_t s) {
        return loss_backward_impl(x, dtype, inv_norm, N, d, row0, n, logit_scale, pair_weight, path, scratch, scratch_bytes,
                                  grad_feat_scale, grad_feat_scale_dev, dx, dscale_partial, s);
    });
}

int clibd_loss_backward_sweeps(const void* const x[3], int dtype, const float* const inv_norm[3], int64_t N, int64_t d,
                               int64_t row0, int64_t n, float logit_scale, const float pair_weight[3], int path,
                               void* scratch, int64_t scratch_bytes, const float* posrow, float* const part[3],
                               float* const peer_red[], int rank, int world, clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_loss_backward_sweeps");
    CLIBD_REQUIRE(x && inv_norm && pair_weight && N > 0 && d > 0 && n >= 0, "null pointer or bad shape");
    CLIBD_REQUIRE(world >= 1 && world <= MAX_PEERS, "world must be in [1, 16]");
    GraphKey k = base_key(4, x, dtype, inv_norm, N, d, row0, n, logit_scale, pair_weight, path, 1, scratch, scratch_bytes);
    k.add(posrow).add_array(part, part ? 3 : 0).add_array(peer_red, peer_red ? world * 3 : 0).add(rank).add(world);
    return run_graphed(k, graph_worthwhile(N, n), stream, [&](cudaStream_t s) {
        return loss_backward_sweeps_impl(x, dtype, inv_norm, N, d, row0, n, logit_scale, pair_weight, path, scratch,
                                         scratch_bytes, posrow, part, peer_red, rank, world, s);
    });
}

int clibd_loss_backward_finish(const void* const x[3], int dtype, const float* const inv_norm[3], int64_t N, int64_t d,
                               int64_t row0, int64_t n, float logit_scale, const float pair_weight[3], int path,
                               void* scratch, int64_t scratch_bytes, const float* const reduced[3],
                               const int reduced_slots[3], float grad_feat_scale, const float* grad_feat_scale_dev,
                               int grad_count, void* const dx[3], double* dscale_partial, clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_loss_backward_finish");
    CLIBD_REQUIRE(x && inv_norm && pair_weight && N > 0 && d > 0 && n >= 0, "null pointer or bad shape");
    GraphKey k = base_key(5, x, dtype, inv_norm, N, d, row0, n, logit_scale, pair_weight, path, 1, scratch, scratch_bytes);
    k.add_array(reduced, reduced ? 3 : 0).add_array(reduced_slots, reduced_slots ? 3 : 0).add(grad_feat_scale);
    k.add(grad_feat_scale_dev).add(grad_count).add_array(dx, dx ? 3 : 0).add(dscale_partial);
    return run_graphed(k, graph_worthwhile(N, n), stream, [&](cudaStream_t s) {
        return loss_backward_finish_impl(x, dtype, inv_norm, N, d, row0, n, logit_scale, pair_weight, path, scratch,
                                         scratch_bytes, reduced, reduced_slots, grad_feat_scale, grad_feat_scale_dev,
                                         grad_count, dx, dscale_partial, s);
    });
}

}  // extern "C"
