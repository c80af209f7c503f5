// Shared declarations of the NT-Xent backward kernels (ntxent_bwd.cu: one CTA per dim half; ntxent_bwd_pc.cu:
// producer/consumer SM pairs; ntxent_bwd_g.cu: shared-G form, one GPU or sharded over ranks).
#pragma once
#include "host_common.h"

namespace tcl {

static constexpr int BW_BM = 128, BW_BN = 128, BW_BK = 64;
static constexpr int BW_KB_BYTES = BW_BM * BW_BK * 2;  // 16 KB
static constexpr int BW_STAGES = 4;
static constexpr int BW_EPI_WARPS = 8;  // two per TMEM lane quarter
static constexpr int BW_EPI_THREADS = BW_EPI_WARPS * 32;
static constexpr int BW_THREADS = 64 + BW_EPI_THREADS;
static constexpr int BW_DH = 256;  // dim columns per CTA
static constexpr int kBwdMaxSplit = 8;
// The 16-bit gradient weights are stored as G * kGScale / max|grad_scale| (a power of two: exact), and the factor is
// taken out again in fp32 by the normalise backward.  Off-diagonal entries are ~ weight / B: at B = 8192 that is 3e-5,
// below fp16's smallest normal number (6.1e-5), where only 9 instead of 11 significant bits survive; scaled by 4096
// every entry of interest is a normal number and the largest (|G| <= 1 on the diagonal) is far from fp16's maximum.
static constexpr float kGScale = 4096.f;  // fp32 gradient partials per 128-row unit the workspace provides for


struct BwdSegDev {
  CUtensorMap tm_other;    // [n_other, dim]   box {64, 128}
  CUtensorMap tm_other_t;  // [dim, n_other]   box {64, 128} (producer/consumer kernel: {64, 256})
  const float* lse2_self;
  const float* lse2_other;
  const float* grad_scale;
  float w_self, w_other;
};
struct BwdJobDev {
  CUtensorMap tm_self;  // [n_self, dim] box {64, 128}
  BwdSegDev seg[2];
  const uint16_t* z_self;  // raw pointer of the self operand (producer/consumer kernel: loaded into TMEM)
  float* gpart;      // [n_split][n_self][dim]
  float* scale_out;  // device scalar consumed by the normalise backward
  int n_seg;
};
struct BwdParams {
  BwdJobDev job[TCL_MAX_TENSORS];
  int n_self, n_other, self_offset, dim;
  int num_kb, n_jtiles, n_split, n_dhalf;
  int64_t z_row_stride;  // elements
  float c1;         // log2(e)/tau
  float out_scale;  // 1/(tau*n_other)
  uint32_t idesc;   // M=128, N=128
  uint32_t idesc_n256;      // M=128, N=256 (producer/consumer kernel: gradient MMA)
  // persistent producer/consumer kernel: the flat tile sequence (job-major, then 128-row unit, then tile) is cut
  // into n_clusters equal contiguous ranges; a unit cut by a range boundary leaves one partial per piece
  CUtensorMap tm_gpart[TCL_MAX_TENSORS];  // [kBwdMaxSplit * n_self_pad, dim] f32, box {32 cols, 32 rows}, 128-B swizzle
  int64_t job_tile_base[TCL_MAX_TENSORS + 1];
  int unit_tiles[TCL_MAX_TENSORS];  // n_seg * n_jtiles
  int n_iblocks, n_self_pad;
};

struct BwdSmem {
  static constexpr uint32_t x_off = 0;  // num_kb * 16 KB
  static constexpr uint32_t g_off(int num_kb) { return num_kb * BW_KB_BYTES; }  // 2 * 16 KB
  static constexpr uint32_t ring_off(int num_kb) { return g_off(num_kb) + 2 * BW_KB_BYTES; }
  static constexpr uint32_t bar_off(int num_kb) { return ring_off(num_kb) + BW_STAGES * BW_KB_BYTES; }
  static constexpr uint32_t bj_off(int num_kb) { return bar_off(num_kb) + 256; }  // 2 x 128 floats
  static constexpr uint32_t total(int num_kb) { return bj_off(num_kb) + 1024 + 1024; }
};

template <int kOp>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if (kOp == TCL_OP_F16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

// ntxent_bwd_pc.cu: persistent; *n_clusters_out = the number of tile ranges it used (<= pc_max_pieces() per unit)
int launch_bwd_pc(const BwdParams& P, int n_jobs, int op_format, int* n_clusters_out, cudaStream_t st);
// ntxent_bwd_g.cu: shared-G form of the single-GPU whole-loss backward (G of a pair formed once, two GEMM passes)
struct BwdSharedGArgs {
  int n_tensors, n_pairs;
  const int32_t* pair_row;
  const int32_t* pair_col;
  const uint8_t* need_grad;
  const void* const* z;      // [n_tensors] normalised 16-bit operands [batch, dim], contiguous
  const void* const* x;      // [n_tensors] original inputs
  void* const* dx;           // [n_tensors]
  const float* inv_norm;     // [n_tensors][batch]
  const float* lse_row;      // [n_pairs][batch]
  const float* lse_col;
  const float* grad_losses;  // [n_pairs] device
  int64_t batch, dim, x_row_stride;
  int x_dtype, op_format;
  float inv_tau, alpha, eps;
  void* workspace;           // [partials (tcl_ntxent_bwd_workspace_bytes)] [G matrices]
  size_t partials_bytes;
};
size_t bwd_sharedg_workspace_bytes(int n_pairs, int64_t batch);
bool bwd_sharedg_enabled(int n_pairs, int64_t batch, int64_t dim);
// ntxent_small.cu: whole loss in one cooperative launch per direction when every 64 x 64 tile gets its own SM
bool small_enabled(int n_tensors, int n_pairs, int64_t batch, int64_t dim, int op_format);
size_t small_fwd_workspace_bytes(int n_pairs, int64_t batch);
size_t small_bwd_workspace_bytes(int n_tensors, int64_t batch, int64_t dim);
int launch_small_fwd(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim, int64_t x_row_stride,
                     int n_pairs, const int32_t* pair_row, const int32_t* pair_col, int op_format, float inv_tau,
                     float alpha, float eps, void* const* z, float* const* inv, float* diag2, float* lse_row,
                     float* lse_col, float* loss_parts, float* loss, bool want_total, void* workspace,
                     size_t workspace_bytes, cudaStream_t st);
int launch_small_bwd(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim, int64_t x_row_stride,
                     int n_pairs, const int32_t* pair_row, const int32_t* pair_col, int op_format, float inv_tau,
                     float alpha, float eps, const void* const* z, const float* inv_base, const float* lse_row,
                     const float* lse_col, const float* grad_losses, const float* grad_total, const uint8_t* need_grad,
                     void* const* dx, void* workspace, size_t workspace_bytes, cudaStream_t st);
bool fold_enabled();  // normalise backward inside the gradient kernels' read-out (norm_fold.cuh)
int launch_bwd_sharedg(const BwdSharedGArgs& a, cudaStream_t st);

}  // namespace tcl
