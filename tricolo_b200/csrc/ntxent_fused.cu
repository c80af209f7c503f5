// Single-call forms of the whole loss forward / backward on one GPU: the host side of
// TriCoLoNet._calculate_losses (tricolo/model/tricolo_net.py:56-65) + NTXentLoss.forward
// (tricolo/loss/nt_xent.py:24-74) and of their autograd, as two C-ABI entry points.  They only
// sequence the kernels of l2norm.cu / ntxent_fwd.cu / ntxent_bwd.cu over one caller-allocated
// "state" buffer (kept between forward and backward) and one scratch workspace, so that a training
// step costs two library calls instead of seven.
#include "ntxent_bwd.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// layout of the state buffer (all offsets 256-byte aligned)
struct LossState {
  size_t z, inv, row_sum, col_sum, diag2, lse_row, lse_col, parts, total;
};
static LossState loss_state_layout(int n_tensors, int n_pairs, int64_t batch, int64_t dim) {
  LossState L;
  size_t off = 0;
  L.z = off;       off = align_up(off + static_cast<size_t>(n_tensors) * batch * dim * 2, 256);
  L.inv = off;     off = align_up(off + static_cast<size_t>(n_tensors) * batch * 4, 256);
  L.row_sum = off; off = align_up(off + static_cast<size_t>(n_pairs) * batch * 4, 256);
  L.col_sum = off; off = align_up(off + static_cast<size_t>(n_pairs) * batch * 4, 256);
  L.diag2 = off;   off = align_up(off + static_cast<size_t>(n_pairs) * batch * 4, 256);
  L.lse_row = off; off = align_up(off + static_cast<size_t>(n_pairs) * batch * 4, 256);
  L.lse_col = off; off = align_up(off + static_cast<size_t>(n_pairs) * batch * 4, 256);
  L.parts = off;   off = align_up(off + static_cast<size_t>(n_pairs) * 2 * 4, 256);
  L.total = off;
  return L;
}

// out[p] = d/d loss[p] + d/d (sum of the losses); either input may be null
__global__ void combine_grads_kernel(const float* grad_losses, const float* grad_total, int n, float* out) {
  griddep_wait();
  const int p = threadIdx.x;
  if (p < n) out[p] = (grad_losses ? grad_losses[p] : 0.f) + (grad_total ? *grad_total : 0.f);
}

}  // namespace tcl

using namespace tcl;

extern "C" size_t tcl_ntxent_loss_state_bytes(int n_tensors, int n_pairs, int64_t batch, int64_t dim) {
  if (n_tensors < 2 || n_tensors > TCL_MAX_TENSORS || n_pairs < 1 || n_pairs > TCL_MAX_PAIRS || batch < 1 || dim < 1) return 0;
  return loss_state_layout(n_tensors, n_pairs, batch, dim).total;
}

extern "C" size_t tcl_ntxent_loss_workspace_bytes(int n_tensors, int n_pairs, int64_t batch, int64_t dim) {
  if (n_tensors < 2 || n_tensors > TCL_MAX_TENSORS || n_pairs < 1 || n_pairs > TCL_MAX_PAIRS || batch < 1 || dim < 1) return 0;
  const size_t fwd = tcl_ntxent_fwd_workspace_bytes(n_pairs, batch, batch);
  const int64_t ld_t = (batch + 7) / 8 * 8;
  size_t bwd = align_up(static_cast<size_t>(n_tensors) * dim * ld_t * 2, 256) +
               align_up(tcl_ntxent_bwd_workspace_bytes(n_tensors, batch, dim), 1024);
  if (bwd_sharedg_enabled(n_pairs, batch, dim)) bwd += bwd_sharedg_workspace_bytes(n_pairs, batch);  // G per pair
  size_t need = fwd > bwd ? fwd : bwd;
  // small-batch single-launch form (ntxent_small.cu); sized independently of the device so that the answer is the
  // same with and without a GPU
  if (batch <= 4096 && dim <= 512) {
    const size_t sf = small_fwd_workspace_bytes(n_pairs, batch), sb = small_bwd_workspace_bytes(n_tensors, batch, dim);
    if (batch <= 1024) need = need > sb ? need : sb;
    need = need > sf ? need : sf;
  }
  return need + 256 + 256;  // the last 256 bytes: combined upstream gradients of tcl_ntxent_loss_bwd_total
}

static int loss_fwd_impl(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim,
                         int64_t x_row_stride, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                         int op_format, float inv_tau, float alpha, float eps, void* state, size_t state_bytes,
                         void* workspace, size_t workspace_bytes, float* loss, bool want_total, void* stream) {
  TCL_REQUIRE(n_tensors >= 2 && n_tensors <= TCL_MAX_TENSORS && n_pairs >= 1 && n_pairs <= TCL_MAX_PAIRS, TCL_ERR_BAD_ARG,
              "loss_fwd: %d tensors / %d pairs", n_tensors, n_pairs);
  TCL_REQUIRE(x && pair_row && pair_col && state && workspace && loss, TCL_ERR_BAD_ARG, "loss_fwd: null pointer");
  const LossState L = loss_state_layout(n_tensors, n_pairs, batch, dim);
  TCL_REQUIRE(state_bytes >= L.total, TCL_ERR_WORKSPACE, "loss_fwd: state buffer too small");
  TCL_REQUIRE(workspace_bytes >= tcl_ntxent_fwd_workspace_bytes(n_pairs, batch, batch), TCL_ERR_WORKSPACE, "loss_fwd: workspace too small");
  TCL_REQUIRE(aligned_to(state, 256), TCL_ERR_BAD_ALIGN, "loss_fwd: state buffer must be 256-byte aligned");
  char* st8 = static_cast<char*>(state);
  void* z[TCL_MAX_TENSORS];
  float* inv[TCL_MAX_TENSORS];
  for (int m = 0; m < n_tensors; ++m) {
    z[m] = st8 + L.z + static_cast<size_t>(m) * batch * dim * 2;
    inv[m] = reinterpret_cast<float*>(st8 + L.inv) + static_cast<size_t>(m) * batch;
  }
  if (small_enabled(n_tensors, n_pairs, batch, dim, op_format) && workspace_bytes >= small_fwd_workspace_bytes(n_pairs, batch))
    return launch_small_fwd(n_tensors, x, x_dtype, batch, dim, x_row_stride, n_pairs, pair_row, pair_col, op_format, inv_tau,
                            alpha, eps, z, inv, reinterpret_cast<float*>(st8 + L.diag2),
                            reinterpret_cast<float*>(st8 + L.lse_row), reinterpret_cast<float*>(st8 + L.lse_col),
                            reinterpret_cast<float*>(st8 + L.parts), loss, want_total, workspace, workspace_bytes,
                            static_cast<cudaStream_t>(stream));
  if (int e = tcl_l2norm_fwd(n_tensors, x, x_dtype, batch, dim, x_row_stride, z, 0, op_format, inv, eps, stream)) return e;
  const void* zrow[TCL_MAX_PAIRS];
  const void* zcol[TCL_MAX_PAIRS];
  for (int p = 0; p < n_pairs; ++p) {
    TCL_REQUIRE(pair_row[p] >= 0 && pair_row[p] < n_tensors && pair_col[p] >= 0 && pair_col[p] < n_tensors, TCL_ERR_BAD_ARG,
                "loss_fwd: pair %d out of range", p);
    zrow[p] = z[pair_row[p]];
    zcol[p] = z[pair_col[p]];
  }
  float* row_sum = reinterpret_cast<float*>(st8 + L.row_sum);
  float* col_sum = reinterpret_cast<float*>(st8 + L.col_sum);
  float* diag2 = reinterpret_cast<float*>(st8 + L.diag2);
  // the sum of the pair losses (want_total) is added by the last pair's finalise cluster
  return ntxent_fwd_finalize_fused(n_pairs, zrow, zcol, batch, dim, op_format, inv_tau, alpha, row_sum, col_sum, diag2,
                                   reinterpret_cast<float*>(st8 + L.lse_row), reinterpret_cast<float*>(st8 + L.lse_col),
                                   reinterpret_cast<float*>(st8 + L.parts), loss, workspace, workspace_bytes, stream,
                                   want_total);
}

extern "C" int tcl_ntxent_loss_fwd(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim,
                                   int64_t x_row_stride, int n_pairs, const int32_t* pair_row,
                                   const int32_t* pair_col, int op_format, float inv_tau, float alpha, float eps,
                                   void* state, size_t state_bytes, void* workspace, size_t workspace_bytes,
                                   float* loss, void* stream) {
  return loss_fwd_impl(n_tensors, x, x_dtype, batch, dim, x_row_stride, n_pairs, pair_row, pair_col, op_format, inv_tau,
                       alpha, eps, state, state_bytes, workspace, workspace_bytes, loss, false, stream);
}
extern "C" int tcl_ntxent_loss_fwd_total(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim,
                                         int64_t x_row_stride, int n_pairs, const int32_t* pair_row,
                                         const int32_t* pair_col, int op_format, float inv_tau, float alpha, float eps,
                                         void* state, size_t state_bytes, void* workspace, size_t workspace_bytes,
                                         float* loss, void* stream) {
  return loss_fwd_impl(n_tensors, x, x_dtype, batch, dim, x_row_stride, n_pairs, pair_row, pair_col, op_format, inv_tau,
                       alpha, eps, state, state_bytes, workspace, workspace_bytes, loss, true, stream);
}

static int loss_bwd_impl(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim,
                         int64_t x_row_stride, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                         int op_format, float inv_tau, float alpha, float eps, const void* state,
                         const float* grad_losses, const float* grad_total, const uint8_t* need_grad, void* const* dx,
                         void* workspace, size_t workspace_bytes, void* stream) {
  TCL_REQUIRE(n_tensors >= 2 && n_tensors <= TCL_MAX_TENSORS && n_pairs >= 1 && n_pairs <= TCL_MAX_PAIRS, TCL_ERR_BAD_ARG,
              "loss_bwd: %d tensors / %d pairs", n_tensors, n_pairs);
  TCL_REQUIRE(x && pair_row && pair_col && state && workspace && (grad_losses || grad_total) && need_grad && dx, TCL_ERR_BAD_ARG,
              "loss_bwd: null pointer");
  TCL_REQUIRE(workspace_bytes >= tcl_ntxent_loss_workspace_bytes(n_tensors, n_pairs, batch, dim), TCL_ERR_WORKSPACE, "loss_bwd: workspace too small");
  TCL_REQUIRE(aligned_to(workspace, 256), TCL_ERR_BAD_ALIGN, "loss_bwd: workspace must be 256-byte aligned");
  const LossState L = loss_state_layout(n_tensors, n_pairs, batch, dim);
  const char* st8 = static_cast<const char*>(state);
  char* ws8 = static_cast<char*>(workspace);
  const int64_t ld_t = (batch + 7) / 8 * 8;
  const void* z[TCL_MAX_TENSORS];
  void* zt[TCL_MAX_TENSORS];
  for (int m = 0; m < n_tensors; ++m) {
    z[m] = st8 + L.z + static_cast<size_t>(m) * batch * dim * 2;
    zt[m] = ws8 + static_cast<size_t>(m) * dim * ld_t * 2;
  }
  if (small_enabled(n_tensors, n_pairs, batch, dim, op_format) && workspace_bytes >= small_bwd_workspace_bytes(n_tensors, batch, dim))
    return launch_small_bwd(n_tensors, x, x_dtype, batch, dim, x_row_stride, n_pairs, pair_row, pair_col, op_format, inv_tau,
                            alpha, eps, z, reinterpret_cast<const float*>(st8 + L.inv),
                            reinterpret_cast<const float*>(st8 + L.lse_row), reinterpret_cast<const float*>(st8 + L.lse_col),
                            grad_losses, grad_total, need_grad, dx, workspace, workspace_bytes,
                            static_cast<cudaStream_t>(stream));
  if (grad_total != nullptr) {
    // the multi-kernel forms read ONE array of upstream gradients: combine into the reserved tail of the workspace
    float* comb = reinterpret_cast<float*>(ws8 + tcl_ntxent_loss_workspace_bytes(n_tensors, n_pairs, batch, dim) - 256);
    LaunchCfg lc(dim3(1), dim3(32), 0, static_cast<cudaStream_t>(stream));
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&lc.cfg, combine_grads_kernel, grad_losses, grad_total, n_pairs, comb));
    grad_losses = comb;
  }
  const size_t zt_bytes = align_up(static_cast<size_t>(n_tensors) * dim * ld_t * 2, 256);
  const bool need_t = tcl_ntxent_bwd_needs_transpose(dim) != 0 && !bwd_sharedg_enabled(n_pairs, batch, dim);
  if (need_t) {
    if (int e = tcl_transpose_16bit(n_tensors, z, batch, dim, 0, zt, ld_t, stream)) return e;
  }
  const float* lse_row = reinterpret_cast<const float*>(st8 + L.lse_row);
  const float* lse_col = reinterpret_cast<const float*>(st8 + L.lse_col);
  if (bwd_sharedg_enabled(n_pairs, batch, dim)) {
    // one GPU, whole batch: the pair's gradient matrix G is formed once and shared by both of its tensors
    for (int p = 0; p < n_pairs; ++p)
      TCL_REQUIRE(pair_row[p] >= 0 && pair_row[p] < n_tensors && pair_col[p] >= 0 && pair_col[p] < n_tensors &&
                      pair_row[p] != pair_col[p], TCL_ERR_BAD_ARG, "loss_bwd: pair %d out of range", p);
    TCL_REQUIRE(op_format == TCL_OP_F16 || op_format == TCL_OP_BF16, TCL_ERR_BAD_ARG, "op_format %d", op_format);
    TCL_REQUIRE(dim % 64 == 0 && dim <= 512, TCL_ERR_BAD_SHAPE, "loss_bwd: dim %lld", (long long)dim);
    if (int e = require_sm100()) return e;
    BwdSharedGArgs a;
    a.n_tensors = n_tensors; a.n_pairs = n_pairs;
    a.pair_row = pair_row; a.pair_col = pair_col; a.need_grad = need_grad;
    a.z = z; a.x = x; a.dx = dx;
    a.inv_norm = reinterpret_cast<const float*>(st8 + L.inv);
    a.lse_row = lse_row; a.lse_col = lse_col; a.grad_losses = grad_losses;
    a.batch = batch; a.dim = dim; a.x_row_stride = x_row_stride;
    a.x_dtype = x_dtype; a.op_format = op_format;
    a.inv_tau = inv_tau; a.alpha = alpha; a.eps = eps;
    a.workspace = ws8 + zt_bytes;
    a.partials_bytes = align_up(tcl_ntxent_bwd_workspace_bytes(n_tensors, batch, dim), 1024);
    for (int m = 0; m < n_tensors; ++m)
      TCL_REQUIRE(!need_grad[m] || dx[m] != nullptr, TCL_ERR_BAD_ARG, "loss_bwd: dx[%d] is null", m);
    return launch_bwd_sharedg(a, static_cast<cudaStream_t>(stream));
  }
  tcl_bwd_job jobs[TCL_MAX_TENSORS];
  memset(jobs, 0, sizeof(jobs));
  int n_jobs = 0;
  for (int m = 0; m < n_tensors; ++m) {
    if (!need_grad[m]) continue;
    tcl_bwd_job& J = jobs[n_jobs];
    J.z_self = z[m];
    J.x_self = x[m];
    J.inv_norm = reinterpret_cast<const float*>(st8 + L.inv) + static_cast<size_t>(m) * batch;
    J.dx = dx[m];
    TCL_REQUIRE(J.dx != nullptr, TCL_ERR_BAD_ARG, "loss_bwd: dx[%d] is null", m);
    for (int p = 0; p < n_pairs; ++p) {
      if (pair_row[p] != m && pair_col[p] != m) continue;
      TCL_REQUIRE(J.n_segments < 2, TCL_ERR_BAD_ARG, "loss_bwd: tensor %d takes part in more than two pairs", m);
      tcl_bwd_segment& S = J.seg[J.n_segments++];
      const bool is_row = pair_row[p] == m;  // self is the pair's first argument: row softmax weight alpha
      const int o = is_row ? pair_col[p] : pair_row[p];
      S.z_other = z[o];
      S.z_other_t = need_t ? zt[o] : nullptr;
      S.lse2_self = (is_row ? lse_row : lse_col) + static_cast<size_t>(p) * batch;
      S.lse2_other = (is_row ? lse_col : lse_row) + static_cast<size_t>(p) * batch;
      S.grad_scale = grad_losses + p;
      S.w_self = is_row ? alpha : 1.f - alpha;
      S.w_other = is_row ? 1.f - alpha : alpha;
    }
    if (J.n_segments > 0) ++n_jobs;
  }
  if (n_jobs == 0) return TCL_OK;
  return tcl_ntxent_bwd(n_jobs, jobs, batch, batch, dim, 0, 0, ld_t, x_dtype, x_row_stride, op_format, inv_tau, eps,
                        ws8 + zt_bytes, workspace_bytes - zt_bytes, stream);
}

extern "C" int tcl_ntxent_loss_bwd(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim,
                                   int64_t x_row_stride, int n_pairs, const int32_t* pair_row,
                                   const int32_t* pair_col, int op_format, float inv_tau, float alpha, float eps,
                                   const void* state, const float* grad_losses, const uint8_t* need_grad,
                                   void* const* dx, void* workspace, size_t workspace_bytes, void* stream) {
  TCL_REQUIRE(grad_losses != nullptr, TCL_ERR_BAD_ARG, "loss_bwd: null pointer");
  return loss_bwd_impl(n_tensors, x, x_dtype, batch, dim, x_row_stride, n_pairs, pair_row, pair_col, op_format, inv_tau,
                       alpha, eps, state, grad_losses, nullptr, need_grad, dx, workspace, workspace_bytes, stream);
}
extern "C" int tcl_ntxent_loss_bwd_total(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim,
                                         int64_t x_row_stride, int n_pairs, const int32_t* pair_row,
                                         const int32_t* pair_col, int op_format, float inv_tau, float alpha, float eps,
                                         const void* state, const float* grad_losses, const float* grad_total,
                                         const uint8_t* need_grad, void* const* dx, void* workspace,
                                         size_t workspace_bytes, void* stream) {
  return loss_bwd_impl(n_tensors, x, x_dtype, batch, dim, x_row_stride, n_pairs, pair_row, pair_col, op_format, inv_tau,
                       alpha, eps, state, grad_losses, grad_total, need_grad, dx, workspace, workspace_bytes, stream);
}
