// Host-side helpers shared by the C-ABI entry points: error plumbing, device
// check, TMA tensor-map encoding through the driver entry point (no link-time
// dependency on libcuda).
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "common.cuh"

namespace tcl {

// thread-local last error text, read back through tcl_last_error_string()
char* last_error_buf();
int set_error(int code, const char* fmt, ...);

#define TCL_CHECK_CUDA(expr)                                                             \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::tcl::set_error(TCL_ERR_CUDA_BASE + static_cast<int>(_e), "%s: %s", \
                              #expr, cudaGetErrorString(_e));                            \
  } while (0)

#define TCL_REQUIRE(cond, code, ...)                           \
  do {                                                         \
    if (!(cond)) return ::tcl::set_error((code), __VA_ARGS__); \
  } while (0)

// 0 when the current device is sm_100 (B200); error otherwise.  Cached per device.
int require_sm100();

// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device.  The attribute is per device, so the
// cache is keyed by (kernel, device ordinal) and guarded by a mutex (a process may drive several GPUs / host threads).
int ensure_dyn_smem_impl(const void* func, int bytes);
template <typename F>
inline int ensure_dyn_smem(F* kernel, int bytes) {
  return ensure_dyn_smem_impl(reinterpret_cast<const void*>(kernel), bytes);
}

// Launch configuration with the optional attributes this library uses: a thread-block cluster along x and programmatic
// dependent launch (TRICOLO_B200_PDL=0 switches the latter off).
bool pdl_enabled();
struct LaunchCfg {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[2];
  LaunchCfg(dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x = 1, bool pdl = true) {
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    int n = 0;
    if (cluster_x > 1) {
      attr[n].id = cudaLaunchAttributeClusterDimension;
      attr[n].val.clusterDim.x = cluster_x;
      attr[n].val.clusterDim.y = 1;
      attr[n].val.clusterDim.z = 1;
      ++n;
    }
    if (pdl && pdl_enabled()) {
      attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[n].val.programmaticStreamSerializationAllowed = 1;
      ++n;
    }
    cfg.attrs = attr;
    cfg.numAttrs = n;
  }
};

// Encode a 2D row-major [rows, cols] 16-bit tensor for TMA tiled loads with a
// {box_cols, box_rows} box and the 128-byte swizzle.  Out-of-bounds elements
// are zero-filled (ragged edge tiles rely on this).
int make_tmap_2d_16bit(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols,
                       uint64_t row_stride_elems, uint32_t box_rows, uint32_t box_cols);
// A row-major [rows, n_chunks * 64] 16-bit matrix viewed as (64, rows, n_chunks): box {64, box_rows, box_chunks} lands
// as box_chunks consecutive 128-byte-swizzled {64, box_rows} tiles (tma_load_3d).  Out-of-range rows / chunks are
// zero-filled.
int make_tmap_3d_16bit(CUtensorMap* out, const void* base, uint64_t rows, uint64_t n_chunks, uint64_t row_stride_elems,
                       uint32_t box_rows, uint32_t box_chunks);
// The same for a dense fp32 [rows, cols] tensor (TMA stores of accumulator tiles): 32-element inner box.
int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                     uint32_t box_cols);

// Persistent gradient kernel (ntxent_bwd_pc.cu): tile range [lo, hi) of range c out of n over `total` tiles, and the
// range a tile belongs to.
__host__ __device__ inline int64_t pc_range_lo(int64_t total, int c, int n) { return total * c / n; }
__host__ __device__ inline int pc_range_of(int64_t total, int64_t t, int n) {
  // 32-bit division when everything fits (the normalise backward evaluates this twice per row; the emulated 64-bit
  // division is ~10x the instructions)
  if (((static_cast<uint64_t>(total) * static_cast<uint64_t>(n)) >> 31) == 0)
    return static_cast<int>(((static_cast<uint32_t>(t) + 1u) * static_cast<uint32_t>(n) - 1u) / static_cast<uint32_t>(total));
  return static_cast<int>(((t + 1) * n - 1) / total);
}

// Word layout (uint32) of the "sync pad" of the sharded loss: one zero-initialised, peer-mapped 4 KB buffer per rank.
// Epoch counters are written by the rank itself; flags are written by the peers (release stores over NVLink) and
// polled locally.  include/tricolo_b200.h: tcl_shard_sync_bytes.
struct ShardSync {
  static constexpr int kFwdEpoch = 0;    // forward steps completed by this rank's K1
  static constexpr int kBwdEpoch = 1;    // backward steps completed by this rank's gradient GEMM
  static constexpr int kK1Done = 2;      // block counters (reset to 0 by the block that completes them)
  static constexpr int kStatsDone = 3;
  static constexpr int kGemmDone = 4;
  static constexpr int kReady = 16;      // [src]: rank src no longer reads its operand buffer of the previous step
  static constexpr int kStats = 32;      // [src]: rank src's sum-exp statistics have landed here
  static constexpr int kGrads = 48;      // [src]: rank src's gradient partials have landed here
  static constexpr int kArrived = 64;    // [src][chunk]: 128-row chunk of rank src's normalised rows has landed here
  static constexpr int kMaxChunks = 64;  // rows per rank <= 8192
  static constexpr int kChunkCnt = kArrived + TCL_MAX_PEERS * kMaxChunks;  // [chunk]: local block counters of K1
  static constexpr int kWords = 1024;
};
static_assert(ShardSync::kChunkCnt + ShardSync::kMaxChunks <= ShardSync::kWords, "sync pad layout");

// normalise backward (l2norm.cu), launched by tcl_ntxent_bwd after the gradient GEMM
struct NormBwdJob {
  const void* x;
  const float* inv_norm;
  const float* gpart;  // [n_split][rows][dim]
  const float* scale;  // device scalar written by the gradient kernel
  void* dx;
};
struct NormBwdParams {
  NormBwdJob job[3];
  // partial layout [slot][split_rows][dim]; split_rows = 0 means `rows`.  n_clusters > 0: the partial count of a
  // 128-row unit follows from the persistent gradient kernel's tile ranges (ntxent_bwd.h: pc_range_of).
  int64_t split_rows;
  int64_t total_tiles;
  int64_t job_tile_base[3];
  int unit_tiles[3];
  int n_clusters;
  int unit_shift;  // log2 of the rows of one unit: 7 (one 128-row block; 0 means 7) or 8 (CTA-pair gradient GEMM)
};
// normalise backward of the sharded shared-G form (ntxent_bwd_g.cu): a tensor's gradient is the sum of its row-side
// partials (local slots, one per tile range that touched the row's unit) and its column-side partials (receive
// buffer: [source rank][slot][b_loc][dim], written by every rank's GEMM kernel over NVLink), each source scaled by
// that source's scale.  Piece counts follow from the shared tile table (pc_range_of), identical on every rank.
struct NormShJob {
  const void* x;
  const float* inv_norm;
  void* dx;
  const float* row_part;  // [kBwdMaxSplit][b_loc][dim] or unused (row_job < 0)
  const void* col_part;   // [world][n_slots_col][b_loc][dim] fp32 or fp16 (NormShParams::col16), or unused (col_job < 0)
  int row_job, col_job;   // indices into the tile table
};
struct NormShParams {
  NormShJob job[3];
  int64_t job_tile_base[6];
  int unit_tiles[6];
  int64_t total_tiles;
  int64_t row_slot_stride;  // elements between row-side slots
  const float* scales;      // [world] scale of every source rank (receive-buffer header)
  const uint32_t* sync;     // own sync pad: wait for every source's kGrads flag (NULL: the caller ran a barrier)
  int n_ranges, n_dsplit, world, rank, n_slots_col;
  int unit_shift;  // log2 rows of a unit: 7, or 8 with the CTA-pair gradient GEMM
  int col16;       // the column-side partials are fp16
};
int launch_l2norm_bwd_sharded(const NormShParams& pr, int n_jobs, int x_dtype, int64_t rows, int dim, int64_t x_stride,
                              float eps, cudaStream_t st);

// ntxent_fwd.cu: partial buffers of the forward tile kernel, for the finalise kernel that reduces them itself
struct FwdPartials {
  const float* row_part;  // [pairs][n_row_slots][n_rows]
  const float* col_part;  // [pairs][n_iblocks][n_cols]
  int n_row_slots, n_iblocks;
};
int ntxent_fwd_finalize_fused(int n_pairs, const void* const* zrow, const void* const* zcol, int64_t batch, int64_t dim,
                              int op_format, float inv_tau, float alpha, float* row_sumexp, float* col_sumexp,
                              float* diag2, float* lse2_row, float* lse2_col, float* loss_parts, float* loss,
                              void* workspace, size_t workspace_bytes, void* stream, bool want_total = false);
// sim_gemm_resident.cu: retrieval GEMM with the query block resident (dim % 64 == 0, dim <= 512)
int launch_sim_gemm_resident(const void* q, const void* g, int64_t n_q, int64_t n_g, int64_t dim, int op_format,
                             float* s, int64_t ld_s, cudaStream_t st);

int launch_l2norm_bwd(const NormBwdParams& pr, int n_jobs, int x_dtype, int64_t rows, int dim,
                      int64_t x_stride, int n_split, float eps, cudaStream_t st);

// launch accounting + optional event timing (api.cu)
void prof_begin(int kernel_id, cudaStream_t st);
void prof_end(int kernel_id, cudaStream_t st);
struct ProfScope {
  int id;
  cudaStream_t st;
  ProfScope(int i, cudaStream_t s) : id(i), st(s) { prof_begin(id, st); }
  ~ProfScope() { prof_end(id, st); }
};

static inline bool aligned_to(const void* p, size_t a) {
  return (reinterpret_cast<uintptr_t>(p) % a) == 0;
}

}  // namespace tcl
