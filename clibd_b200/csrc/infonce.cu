// SimCLR info-NCE loss on one feature set (SURVEY.md section 8 f, rank 4).
//
// Replaces SimCLR.info_nce_loss followed by nn.CrossEntropyLoss (bioscanclip/util/simclr.py:64-92, 118-119) for
// n_views = 2: features [M = 2 B, d], row i and row (i + B) mod M are the two views of one image.
//
//   reference : S = Zhat Zhat^T / tau, main diagonal removed, row i's logits = [S_i,p(i) | all other S_ij],
//               loss = mean_i CE(logits_i, 0) = (1/M) sum_i [ LSE_{j != i} S_ij - S_i,p(i) ].
//   here      : with s = 1/tau, e_ij = exp(S_ij - s), r_i = sum_{j != i} e_ij,
//               loss = (1/M) [ sum_i (s + ln r_i) - s sum_i zhat_i . zhat_p(i) ]
//               dL/dzhat_k = (s/M) [ sum_{j != k} e_kj (1/r_k + 1/r_j) zhat_j - 2 zhat_p(k) ]
//               (S is symmetric in ONE variable: row and column roles of zhat_k add up).
//
// It is the contrastive pair machinery with both operands the same feature set and self_mask = 1 (the tcgen05
// / CUDA-core kernels drop the entries whose row equals their column): class size 1, representative = the row
// itself, "class sum of the partner modality" = the other view's row.  No M x M matrix exists anywhere.
#include "../../include/clibd_b200.h"
#include "common.cuh"
#include "loss_plan.h"

namespace clibd {
namespace {

constexpr int kThreads = 256;

__global__ void infonce_prepare_kernel(int64_t M, int64_t B, int32_t* __restrict__ rep, float* __restrict__ cnt,
                                       int32_t* __restrict__ partner) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < M) {
        rep[i] = static_cast<int32_t>(i);
        cnt[i] = 1.f;
        partner[i] = static_cast<int32_t>((i + B) % M);  // the row's one positive column
    }
}

// P[k,:] = zhat[(k + B) mod M, :] in float32 (one warp per row)
template <typename T>
__global__ void infonce_partner_kernel(const T* __restrict__ z, const float* __restrict__ inv, int64_t M, int64_t B,
                                       int64_t d, float* __restrict__ P) {
    const int64_t k = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= M) return;
    const int64_t j = (k + B) % M;
    const float iv = inv[j];
    const T* zr = z + j * d;
    float* pr = P + k * d;
    for (int64_t c = lane; c < d; c += 32) pr[c] = load_as_float(zr, c) * iv;
}

struct Views {
    int32_t *rep, *partner;
    float *cnt, *gscale, *P, *u, *v, *rowpart, *colpart, *posrow, *dots, *dxh, *ccS, *lam2;
    double* red;
    void *xh, *xhT;
};

Views carve(void* scratch, const LossPlan& plan) {
    Views w;
    w.rep = at<int32_t>(scratch, plan.off_rep);
    w.cnt = at<float>(scratch, plan.off_cnt);
    w.partner = at<int32_t>(scratch, plan.off_class_lo);
    w.ccS = at<float>(scratch, plan.off_ccS);
    w.lam2 = at<float>(scratch, plan.off_lam2);
    w.gscale = at<float>(scratch, plan.off_gscale);
    w.P = at<float>(scratch, plan.off_Q[0]);
    w.u = at<float>(scratch, plan.off_u);
    w.v = at<float>(scratch, plan.off_v);
    w.rowpart = at<float>(scratch, plan.off_rowpart);
    w.colpart = at<float>(scratch, plan.off_colpart);
    w.posrow = at<float>(scratch, plan.off_posrow);
    w.dots = at<float>(scratch, plan.off_dots);
    w.dxh = at<float>(scratch, plan.off_dxh[0]);
    w.red = at<double>(scratch, plan.off_red);
    w.xh = plan.path != PATH_SIMT_F32 ? at<void>(scratch, plan.off_xh[0]) : nullptr;
    w.xhT = plan.path != PATH_SIMT_F32 ? at<void>(scratch, plan.off_xhT[0]) : nullptr;
    return w;
}

int check_args(const void* z, const float* inv_norm, int dtype, int64_t M, int64_t d, float scale, int path,
               void* scratch, int64_t scratch_bytes, const LossPlan& plan) {
    CLIBD_REQUIRE(z && inv_norm, "null pointer");
    CLIBD_REQUIRE(M >= 2 && M % 2 == 0 && d > 0, "info-NCE needs an even number of rows (two views per image)");
    CLIBD_REQUIRE(M < (int64_t(1) << 31) - 512, "too many rows for 32-bit TMA coordinates");
    CLIBD_REQUIRE(dtype == DT_F32 || dtype == DT_BF16 || dtype == DT_F16, "dtype must be 0, 1 or 2");
    CLIBD_REQUIRE(path >= 0 && path <= 2, "path must be 0, 1 or 2");
    // fixed-shift softmax exp(S - s): |S| <= s for unit rows, 2 s log2(e) must stay below 126
    CLIBD_REQUIRE(scale > 0.f && scale < 1e30f, "1 / temperature must be a positive finite number");
    CLIBD_REQUIRE(scratch != nullptr && scratch_bytes >= static_cast<int64_t>(plan.total), "scratch too small");
    if (path != PATH_SIMT_F32) {
        CLIBD_REQUIRE(clibd_device_supported(), "tcgen05 path needs a compute-capability 10.x device");
        CLIBD_REQUIRE(pair_backward_supported(plan.dpad), "tcgen05 info-NCE supports d <= 768; use path 0");
    }
    return 0;
}

}  // namespace
}  // namespace clibd

using namespace clibd;

extern "C" {

int clibd_infonce_forward(const void* z, int dtype, const float* inv_norm, int64_t M, int64_t d,
                          float inv_temperature, int path, void* scratch, int64_t scratch_bytes, float* rowsum,
                          float* loss_out, clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_infonce_forward");
    CLIBD_REQUIRE(M > 0 && d > 0 && path >= 0 && path <= 2, "bad shape");
    const LossPlan plan = make_loss_plan(M, M, d, path, /*allow_shared_s=*/false);
    int rc = check_args(z, inv_norm, dtype, M, d, inv_temperature, path, scratch, scratch_bytes, plan);
    if (rc) return rc;
    CLIBD_REQUIRE(rowsum && loss_out, "null output pointer");
    const Views w = carve(scratch, plan);
    const int64_t B = M / 2;
    const float scale = inv_temperature;
    infonce_prepare_kernel<<<ceil_div(M, kThreads), kThreads, 0, stream>>>(M, B, w.rep, w.cnt, w.partner);
    CLIBD_KERNEL_CHECK();
    if ((rc = launch_gscale(w.cnt, M, path, w.gscale, stream))) return rc;
    {
        const int64_t blocks = ceil_div(M * 32, kThreads);
        switch (dtype) {
            case DT_F32:
                infonce_partner_kernel<float><<<blocks, kThreads, 0, stream>>>(static_cast<const float*>(z), inv_norm, M, B, d, w.P);
                break;
            case DT_BF16:
                infonce_partner_kernel<__nv_bfloat16><<<blocks, kThreads, 0, stream>>>(static_cast<const __nv_bfloat16*>(z), inv_norm, M, B, d, w.P);
                break;
            default:
                infonce_partner_kernel<__half><<<blocks, kThreads, 0, stream>>>(static_cast<const __half*>(z), inv_norm, M, B, d, w.P);
                break;
        }
        CLIBD_KERNEL_CHECK();
    }
    if (path != PATH_SIMT_F32) {
        const int fmt_bf16 = path == PATH_TC_BF16 ? 1 : 0;
        if ((rc = launch_make_operands(z, dtype, inv_norm, M, d, plan.dpad, plan.npad, fmt_bf16, w.xh, w.xhT, stream)))
            return rc;
        rc = tc_forward_pair_cg2(w.xh, w.xh, M, plan.dpad, 0, M, scale, fmt_bf16, w.rowpart, w.colpart, tc_num_sms(),
                                 stream, /*self_mask=*/1);
    } else {
        rc = simt_forward_pair(z, z, dtype, inv_norm, inv_norm, M, d, 0, M, scale, w.rowpart, w.colpart, stream,
                               /*self_mask=*/1);
    }
    if (rc) return rc;
    if ((rc = launch_reduce_parts(w.rowpart, plan.row_parts, M, M, rowsum, stream))) return rc;
    // sum_i zhat_i . zhat_p(i)
    if ((rc = launch_pos_rows(z, dtype, inv_norm, w.P, w.rep, d, 0, M, w.posrow, stream))) return rc;
    double* pos = w.red + 300;  // beyond the 256 block partials the reductions use
    if ((rc = launch_sum_to_double(w.posrow, M, 1.0, w.red, pos, stream))) return rc;
    // S is symmetric: column sums equal row sums, and half of [CE(S) + CE(S^T)] is the info-NCE loss
    const float wts[3] = {0.5f, 0.f, 0.f};
    return launch_loss_finish(M, scale, wts, w.cnt, rowsum, rowsum, pos, w.u, w.v, w.red, loss_out, stream);
}

int clibd_infonce_backward(const void* z, int dtype, const float* inv_norm, int64_t M, int64_t d,
                           float inv_temperature, int path, void* scratch, int64_t scratch_bytes,
                           float grad_scale, const float* grad_scale_dev, void* dz, clibd_stream_t stream) {
    NvtxRange nvtx_range("clibd_infonce_backward");
    CLIBD_REQUIRE(M > 0 && d > 0 && path >= 0 && path <= 2, "bad shape");
    const LossPlan plan = make_loss_plan(M, M, d, path, /*allow_shared_s=*/false);
    int rc = check_args(z, inv_norm, dtype, M, d, inv_temperature, path, scratch, scratch_bytes, plan);
    if (rc) return rc;
    CLIBD_REQUIRE(dz != nullptr, "null output pointer");
    const Views w = carve(scratch, plan);
    const float scale = inv_temperature;
    const bool tc = path != PATH_SIMT_F32;
    if (tc) {
        const int fmt_bf16 = path == PATH_TC_BF16 ? 1 : 0;
        // lam2_i ~ G~ at the partner column: the epilogue emits G~ - lam2 there (columns stay in input order)
        if ((rc = launch_sweep_prep(w.u, w.v, nullptr, w.cnt, w.posrow, M, 0, M, scale, w.ccS, w.lam2, stream))) return rc;
        rc = tc_backward_rows_pair(w.xh, w.xh, w.xhT, M, plan.npad, d, plan.dpad, 0, M, scale, w.u, w.ccS, w.gscale, 1.0f,
                                   /*accumulate=*/0, plan.jsplit, fmt_bf16, w.dxh, stream, /*self_mask=*/1, w.partner,
                                   w.cnt, w.lam2);
    } else {
        rc = simt_backward_rows(z, z, dtype, inv_norm, inv_norm, M, d, 0, M, scale, w.u, w.v, 1.0f, /*accumulate=*/0,
                                w.dxh, stream, /*self_mask=*/1, plan.jsplit);
    }
    if (rc) return rc;
    NormBwdArgs a;
    a.x = z;
    a.dtype = dtype;
    a.inv_norm = inv_norm;
    a.rep = w.rep;
    a.dxh = w.dxh;
    a.jsplit = plan.jsplit;
    a.Qp[0] = w.P;
    a.Qp[1] = nullptr;
    a.wp[0] = 1.f;
    a.wp[1] = 0.f;
    a.lam2[0] = tc ? w.lam2 : nullptr;
    a.N = M;
    a.d = d;
    a.row0 = 0;
    a.n = M;
    a.scale = scale;
    a.grad_scale = grad_scale;
    a.grad_scale_dev = grad_scale_dev;
    a.dx = dz;
    a.dots = w.dots;
    return launch_normalize_bwd(a, stream);
}

}  // extern "C"
