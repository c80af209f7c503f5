// K2'+K4 fused — retrieval without materialising the similarity matrix.
//
// Same tcgen05 main loop as sim_gemm_resident.cu (128 query rows resident in shared memory, gallery
// tiles streamed through a TMA ring, two TMEM accumulators), but the epilogue consumes each 128x128
// tile in registers: every thread owns one query row and 64 of the tile's columns, counts the entries
// that rank before the ground truth and maintains a sorted top-k list (strictly-greater insertion in
// ascending column order = lowest index wins ties).  The Q x G fp32 similarities never reach HBM, so
// the path is bound by the tensor pipe instead of by 2 x Q x G x 4 bytes of traffic.
//
// Order and rank are the same as K4's (topk.cu): similarity descending, gallery index ascending;
//   n_before = #{ s_j > s_gt } + #{ j < gt : s_j == s_gt }.
// The ground-truth similarity must be the very number the MMA produces for (q, gt): it comes from a
// pre-pass (tcl_gt_sim_mma) that gathers the labelled gallery rows and runs the SAME instruction
// sequence (32 MMAs of K=16, same shapes, same order) on the 128x128 tile whose diagonal holds
// q . g[label[q]]; tests check it bit-for-bit against the materialised matrix of tcl_sim_gemm.
#include <float.h>
#include <limits.h>
#include <math.h>

#include "host_common.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

static constexpr int FT_BM = 128, FT_BN = 128, FT_BK = 64;
static constexpr int FT_KB_BYTES = FT_BM * FT_BK * 2;  // 16 KB
static constexpr int FT_STAGES = 6;
static constexpr int FT_EPI_WARPS = 8;
static constexpr int FT_ACC = 4;  // TMEM accumulator buffers (4 x 128 columns = all of TMEM): the epilogue may hold one while
                                  // it re-reads it, without stalling the MMA stream
static constexpr int FT_THREADS = 64 + FT_EPI_WARPS * 32;

enum { FT_MODE_TOPK = 0, FT_MODE_DIAG = 1 };

struct FusedParams {
  CUtensorMap tm_q;  // [n_q, dim] box {64, 128}
  CUtensorMap tm_g;  // TOPK: gallery [n_g, dim]; DIAG: gathered rows [n_q, dim]; box {64, 128}
  const int64_t* labels;
  const float* gt_sim;
  float* out_val;    // [n_split][n_q][k]
  int32_t* out_idx;  // [n_split][n_q][k]   global gallery index, -1 = none
  int32_t* out_nb;   // [n_split][n_q]
  float* gt_out;     // DIAG mode: [n_q]
  int64_t idx_base;
  int n_q, n_g, num_kb, n_gtiles, n_split, k;
  uint32_t idesc;
};

struct FusedSmem {
  static constexpr uint32_t ring_off(int num_kb) { return num_kb * FT_KB_BYTES; }
  static constexpr uint32_t bar_off(int num_kb) { return ring_off(num_kb) + FT_STAGES * FT_KB_BYTES; }
  static constexpr uint32_t total(int num_kb) { return bar_off(num_kb) + 256 + 1024; }
};

__device__ __forceinline__ bool ranks_before(float a, int ia, float b, int ib) {
  return a > b || (a == b && ia < ib);
}

template <int K>
struct RowTop {
  float v[K];
  int i[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int t = 0; t < K; ++t) { v[t] = -FLT_MAX; i[t] = INT_MAX; }
  }
  // candidates arrive in ascending column order: strictly-greater insertion keeps the lowest index on ties
  __device__ __forceinline__ void push_ascending(float x, int idx) {
    if (!(x > v[K - 1])) return;
    v[K - 1] = x; i[K - 1] = idx;
#pragma unroll
    for (int t = K - 1; t > 0; --t)
      if (v[t] > v[t - 1]) {
        float tv = v[t]; v[t] = v[t - 1]; v[t - 1] = tv;
        int ti = i[t]; i[t] = i[t - 1]; i[t - 1] = ti;
      }
  }
  // arbitrary order (merging two lists): full (value, index) comparison
  __device__ __forceinline__ void push_any(float x, int idx) {
    if (!ranks_before(x, idx, v[K - 1], i[K - 1])) return;
    v[K - 1] = x; i[K - 1] = idx;
#pragma unroll
    for (int t = K - 1; t > 0; --t)
      if (ranks_before(v[t], i[t], v[t - 1], i[t - 1])) {
        float tv = v[t]; v[t] = v[t - 1]; v[t - 1] = tv;
        int ti = i[t]; i[t] = i[t - 1]; i[t - 1] = ti;
      }
  }
};

template <int K, int kMode>
__global__ void __launch_bounds__(FT_THREADS, 1) sim_topk_fused_kernel(const __grid_constant__ FusedParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int num_kb = P.num_kb;
  const uint32_t q_smem = base;
  const uint32_t ring = base + FusedSmem::ring_off(num_kb);
  const uint32_t bars = base + FusedSmem::bar_off(num_kb);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (FT_STAGES + s); };
  const uint32_t q_full_bar = bars + 8u * (2 * FT_STAGES);
  auto tmem_full_bar = [&](int b) { return bars + 8u * (2 * FT_STAGES + 1 + b); };
  auto tmem_empty_bar = [&](int b) { return bars + 8u * (2 * FT_STAGES + 1 + FT_ACC + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * FT_STAGES + 1 + 2 * FT_ACC);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + FusedSmem::bar_off(num_kb) + 8u * (2 * FT_STAGES + 1 + 2 * FT_ACC));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * FT_BM;
  const int sp = blockIdx.y;
  int t_begin, t_end;
  if (kMode == FT_MODE_DIAG) {  // the one tile whose diagonal pairs query row r with gathered row r
    t_begin = blockIdx.x;
    t_end = t_begin + 1;
  } else {
    t_begin = static_cast<int>((static_cast<int64_t>(P.n_gtiles) * sp) / P.n_split);
    t_end = static_cast<int>((static_cast<int64_t>(P.n_gtiles) * (sp + 1)) / P.n_split);
  }
  const int n_tiles = t_end - t_begin;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&P.tm_q);
    tma_prefetch_desc(&P.tm_g);
    for (int s = 0; s < FT_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(q_full_bar, 1);
    for (int b = 0; b < FT_ACC; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), FT_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 0) {
    if (elect_one() && n_tiles > 0) {
      mbar_arrive_expect_tx(q_full_bar, num_kb * FT_KB_BYTES);
      for (int kb = 0; kb < num_kb; ++kb)
        tma_load_2d(q_smem + kb * FT_KB_BYTES, &P.tm_q, q_full_bar, kb * FT_BK, m0);
      int it = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int n0 = (t_begin + t) * FT_BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % FT_STAGES;
          const uint32_t ph = (it / FT_STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_arrive_expect_tx(full_bar(s), FT_KB_BYTES);
          tma_load_2d(ring + s * FT_KB_BYTES, &P.tm_g, full_bar(s), kb * FT_BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one() && n_tiles > 0) {
      mbar_wait(q_full_bar, 0);
      int it = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int b = t % FT_ACC;
        mbar_wait(tmem_empty_bar(b), ((t / FT_ACC) & 1) ^ 1);
        tc_fence_after();
        // NOTE: the K order (kb ascending, kk ascending) is part of the contract between the two modes
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % FT_STAGES;
          const uint32_t ph = (it / FT_STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint64_t ad = umma_desc_k_sw128(q_smem + kb * FT_KB_BYTES);
          const uint64_t bd = umma_desc_k_sw128(ring + s * FT_KB_BYTES);
#pragma unroll
          for (int kk = 0; kk < FT_BK / 16; ++kk)
            tc_mma_f16(tmem + b * FT_BN, ad + 2 * kk, bd + 2 * kk, P.idesc, (kb | kk) != 0);
          tc_commit(empty_bar(s));
        }
        tc_commit(tmem_full_bar(b));
      }
    }
  } else {
    const int q = warp & 3;          // TMEM lane quarter
    const int ch = (warp - 2) >> 2;  // column half of every tile (64 columns)
    const int r = q * 32 + lane;     // tile-local query row == TMEM lane
    const int grow = m0 + r;
    const bool row_ok = grow < P.n_q;

    if (kMode == FT_MODE_DIAG) {
      // diagonal element (r, r) lives in column half r/64 = q/2, 32-column chunk q%2, register `lane`
      if (n_tiles > 0) {
        mbar_wait(tmem_full_bar(0), 0);
        tc_fence_after();
        if (ch == (q >> 1)) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_addr(tmem, q * 32, q * 32), v);
          tc_wait_ld();
          float d = 0.f;
#pragma unroll
          for (int e = 0; e < 32; ++e) d = (e == lane) ? __uint_as_float(v[e]) : d;
          if (row_ok) P.gt_out[grow] = d;
        }
      }
    } else {
      RowTop<K> top;
      top.init();
      int cnt = 0;
      // rank bookkeeping: columns below the ground truth count when >= s_gt, columns above when > s_gt.
      // x >= s  <=>  x > pred(s) for finite floats, so both cases are one strict compare.
      const int64_t gt64 = row_ok ? P.labels[grow] - P.idx_base : 0;
      const int gtcol = gt64 < 0 ? -1 : (gt64 >= P.n_g ? P.n_g : static_cast<int>(gt64));  // clamp into [-1, n_g]
      const float s_gt = row_ok ? P.gt_sim[grow] : 0.f;
      const float s_gt_pred = nextafterf(s_gt, -INFINITY);

      for (int t = 0; t < n_tiles; ++t) {
        const int b = t % FT_ACC;
        const int c0 = (t_begin + t) * FT_BN + ch * 64;  // first gallery column (shard-local) of this thread
        const uint32_t tcol = tmem + b * FT_BN + ch * 64;
        mbar_wait(tmem_full_bar(b), (t / FT_ACC) & 1);
        tc_fence_after();
        uint32_t v[2][32];
        tmem_ld_32x32b_x32(tmem_addr(tcol, q * 32, 0), v[0]);
        tmem_ld_32x32b_x32(tmem_addr(tcol, q * 32, 32), v[1]);
        tc_wait_ld();
        // rows whose 64 columns hold zero-filled padding or the ground truth need the exact per-element rule
        const bool exact = (c0 + 64 > P.n_g) || (gtcol >= c0 && gtcol < c0 + 64);
        // fast scan: one compare per element for the rank (four independent counters), one 3-input max per
        // element pair for the top-k threshold, one flag bit per group of 8 columns
        const float thr_rank = (c0 + 64 <= gtcol) ? s_gt_pred : s_gt;
        const float thr_top = top.v[K - 1];
        int c4[4] = {0, 0, 0, 0};
        unsigned gmask = exact ? 0xFFu : 0u;
#pragma unroll
        for (int g8 = 0; g8 < 8; ++g8) {
          float m = -FLT_MAX;
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float x = __uint_as_float(v[g8 >> 2][(g8 & 3) * 8 + e]);
            c4[e & 3] += (x > thr_rank) ? 1 : 0;
            m = fmaxf(m, x);
          }
          gmask |= (m > thr_top) ? (1u << g8) : 0u;
        }
        if (!exact) cnt += (c4[0] + c4[1]) + (c4[2] + c4[3]);
        // Rare exact pass over the flagged 8-column groups only (new top-k candidates, the ground-truth tile,
        // padding).  The group is re-read from TMEM in a rolled, warp-uniform loop so the insertion code exists
        // once (fully unrolled it is ~36 KB of SASS and thrashes the instruction cache).
        unsigned wmask = __reduce_or_sync(0xffffffffu, gmask);
        while (wmask) {
          const int g8 = __ffs(wmask) - 1;
          wmask &= wmask - 1;
          uint32_t w[8];
          tmem_ld_32x32b_x8(tmem_addr(tcol, q * 32, g8 * 8), w);
          tc_wait_ld();
          if ((gmask >> g8) & 1u) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const int col = c0 + g8 * 8 + e;
              const float x = __uint_as_float(w[e]);
              if (col < P.n_g) {
                if (exact) cnt += (x > s_gt || (x == s_gt && col < gtcol)) ? 1 : 0;
                top.push_ascending(x, col);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tmem_empty_bar(b));
      }
      // merge the two column halves of each row through shared memory (the operand ring is dead by now)
      float* mv = reinterpret_cast<float*>(base_ptr + FusedSmem::ring_off(num_kb));  // [128][K]
      int* mi = reinterpret_cast<int*>(mv + 128 * K);                                 // [128][K]
      int* mc = mi + 128 * K;                                                          // [128]
      if (ch == 1) {
#pragma unroll
        for (int t = 0; t < K; ++t) { mv[r * K + t] = top.v[t]; mi[r * K + t] = top.i[t]; }
        mc[r] = cnt;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (ch == 0 && row_ok) {
#pragma unroll
        for (int t = 0; t < K; ++t) top.push_any(mv[r * K + t], mi[r * K + t]);
        cnt += mc[r];
        const int64_t o = (static_cast<int64_t>(sp) * P.n_q + grow);
#pragma unroll
        for (int t = 0; t < K; ++t) {  // constant indices: the list must stay in registers
          if (t < P.k) {
            P.out_val[o * P.k + t] = top.v[t];
            P.out_idx[o * P.k + t] = top.i[t] == INT_MAX ? -1 : static_cast<int32_t>(P.idx_base + top.i[t]);
          }
        }
        P.out_nb[o] = cnt;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// gathered[q, :] = g[label[q] - idx_base, :] (zeros when the label is not in this shard); one warp per row
__global__ void __launch_bounds__(256) gather_rows16_kernel(const uint16_t* __restrict__ g, int64_t n_g, int dim,
                                                            const int64_t* __restrict__ labels, int64_t idx_base,
                                                            int64_t n_q, uint16_t* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (q >= n_q) return;
  const int64_t l = labels[q] - idx_base;
  const bool ok = l >= 0 && l < n_g;
  const uint4* src = reinterpret_cast<const uint4*>(g + (ok ? l : 0) * dim);
  uint4* dst = reinterpret_cast<uint4*>(out + q * dim);
  for (int c = lane; c < dim / 8; c += 32) dst[c] = ok ? __ldg(src + c) : make_uint4(0u, 0u, 0u, 0u);
}

// sum the per-split rank counts (n_split > 1)
__global__ void sum_splits_kernel(const int32_t* __restrict__ part, int n_split, int64_t n_q, int32_t* __restrict__ out) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n_q) return;
  int a = 0;
  for (int s = 0; s < n_split; ++s) a += part[static_cast<int64_t>(s) * n_q + q];
  out[q] = a;
}

static int fused_split(int n_qblocks, int n_gtiles) {
  // one CTA per SM: split the gallery sweep only when there are too few query blocks to fill the GPU
  if (n_qblocks >= kNumSMsB200) return 1;
  int s = kNumSMsB200 / n_qblocks;
  if (s > 16) s = 16;  // tcl_topk_merge handles up to 16 candidate lists
  if (s > n_gtiles) s = n_gtiles;
  return s < 1 ? 1 : s;
}

template <int kMode>
static int launch_fused(const FusedParams& P, int n_qblocks, int n_split, cudaStream_t st) {
  const int smem = static_cast<int>(FusedSmem::total(P.num_kb));
  dim3 grid(n_qblocks, n_split);
  if (P.k <= 5) {
    if (int e = ensure_dyn_smem(sim_topk_fused_kernel<5, kMode>, smem)) return e;
    sim_topk_fused_kernel<5, kMode><<<grid, FT_THREADS, smem, st>>>(P);
  } else {
    if (int e = ensure_dyn_smem(sim_topk_fused_kernel<16, kMode>, smem)) return e;
    sim_topk_fused_kernel<16, kMode><<<grid, FT_THREADS, smem, st>>>(P);
  }
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

}  // namespace tcl

using namespace tcl;

static int check_fused_args(const void* q, const void* g, int64_t n_q, int64_t n_g, int64_t dim, int op_format) {
  TCL_REQUIRE(q && g, TCL_ERR_BAD_ARG, "fused retrieval: null operand");
  TCL_REQUIRE(n_q >= 0 && n_g >= 1 && n_q < (1LL << 31) - 256 && n_g < (1LL << 31) - 256, TCL_ERR_BAD_SHAPE,
              "fused retrieval: bad sizes");
  TCL_REQUIRE(dim >= 64 && dim % 64 == 0 && dim <= 512, TCL_ERR_BAD_SHAPE,
              "fused retrieval: dim must be a multiple of 64 in [64, 512] (got %lld); use tcl_sim_gemm + tcl_topk_rank",
              (long long)dim);
  TCL_REQUIRE(op_format == TCL_OP_F16 || op_format == TCL_OP_BF16, TCL_ERR_BAD_ARG, "op_format %d", op_format);
  return require_sm100();
}

extern "C" int tcl_gt_sim_mma(const void* q, const void* g, int64_t n_q, int64_t n_g, int64_t dim, int op_format,
                              const int64_t* labels, int64_t idx_base, float* gt_sim, void* workspace,
                              size_t workspace_bytes, void* stream) {
  if (int e = check_fused_args(q, g, n_q, n_g, dim, op_format)) return e;
  TCL_REQUIRE(labels && gt_sim && workspace, TCL_ERR_BAD_ARG, "gt_sim_mma: null pointer");
  TCL_REQUIRE(workspace_bytes >= static_cast<size_t>(n_q) * dim * 2, TCL_ERR_WORKSPACE, "gt_sim_mma: workspace too small");
  TCL_REQUIRE(aligned_to(workspace, 16) && aligned_to(g, 16), TCL_ERR_BAD_ALIGN, "gt_sim_mma: alignment");
  if (n_q == 0) return TCL_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(TCL_K_GATHER_GT, st);
  gather_rows16_kernel<<<static_cast<unsigned>((n_q + 7) / 8), 256, 0, st>>>(
      static_cast<const uint16_t*>(g), n_g, static_cast<int>(dim), labels, idx_base, n_q, static_cast<uint16_t*>(workspace));
  TCL_CHECK_CUDA(cudaGetLastError());
  FusedParams P;
  memset(&P, 0, sizeof(P));
  if (int e = make_tmap_2d_16bit(&P.tm_q, q, n_q, dim, dim, FT_BM, FT_BK)) return e;
  if (int e = make_tmap_2d_16bit(&P.tm_g, workspace, n_q, dim, dim, FT_BN, FT_BK)) return e;
  P.gt_out = gt_sim;
  P.n_q = static_cast<int>(n_q);
  P.n_g = static_cast<int>(n_q);
  P.num_kb = static_cast<int>(dim / 64);
  P.n_gtiles = static_cast<int>((n_q + FT_BN - 1) / FT_BN);
  P.n_split = 1;
  P.k = 1;
  P.idesc = umma_idesc_f16(FT_BM, FT_BN, op_format);
  return launch_fused<FT_MODE_DIAG>(P, P.n_gtiles, 1, st);
}

extern "C" size_t tcl_sim_topk_fused_workspace_bytes(int64_t n_q, int64_t n_g, int k) {
  if (n_q < 1 || n_g < 1 || k < 1) return 0;
  const int n_qblocks = static_cast<int>((n_q + FT_BM - 1) / FT_BM);
  const int n_split = fused_split(n_qblocks, static_cast<int>((n_g + FT_BN - 1) / FT_BN));
  if (n_split == 1) return 256;
  return static_cast<size_t>(n_split) * n_q * (static_cast<size_t>(k) * 8 + 4) + 256;
}

extern "C" int tcl_sim_topk_fused(const void* q, const void* g, int64_t n_q, int64_t n_g, int64_t dim, int op_format,
                                  int k, const int64_t* labels, int64_t idx_base, const float* gt_sim,
                                  float* topk_val, int32_t* topk_idx, int32_t* n_before, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  if (int e = check_fused_args(q, g, n_q, n_g, dim, op_format)) return e;
  TCL_REQUIRE(k >= 1 && k <= 16, TCL_ERR_BAD_ARG, "fused retrieval: k must be in [1,16] (got %d)", k);
  TCL_REQUIRE(labels && gt_sim && topk_val && topk_idx && n_before, TCL_ERR_BAD_ARG, "fused retrieval: null pointer");
  if (n_q == 0) return TCL_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FusedParams P;
  memset(&P, 0, sizeof(P));
  if (int e = make_tmap_2d_16bit(&P.tm_q, q, n_q, dim, dim, FT_BM, FT_BK)) return e;
  if (int e = make_tmap_2d_16bit(&P.tm_g, g, n_g, dim, dim, FT_BN, FT_BK)) return e;
  P.labels = labels;
  P.gt_sim = gt_sim;
  P.idx_base = idx_base;
  P.n_q = static_cast<int>(n_q);
  P.n_g = static_cast<int>(n_g);
  P.num_kb = static_cast<int>(dim / 64);
  P.n_gtiles = static_cast<int>((n_g + FT_BN - 1) / FT_BN);
  P.k = k;
  P.idesc = umma_idesc_f16(FT_BM, FT_BN, op_format);
  const int n_qblocks = static_cast<int>((n_q + FT_BM - 1) / FT_BM);
  P.n_split = fused_split(n_qblocks, P.n_gtiles);
  if (P.n_split == 1) {
    P.out_val = topk_val;
    P.out_idx = topk_idx;
    P.out_nb = n_before;
    ProfScope prof(TCL_K_SIM_TOPK_FUSED, st);
    return launch_fused<FT_MODE_TOPK>(P, n_qblocks, 1, st);
  }
  TCL_REQUIRE(workspace && workspace_bytes >= tcl_sim_topk_fused_workspace_bytes(n_q, n_g, k), TCL_ERR_WORKSPACE,
              "fused retrieval: workspace too small");
  char* ws = static_cast<char*>(workspace);
  const size_t n_cand = static_cast<size_t>(P.n_split) * n_q * k;
  P.out_val = reinterpret_cast<float*>(ws);
  P.out_idx = reinterpret_cast<int32_t*>(ws + n_cand * 4);
  P.out_nb = reinterpret_cast<int32_t*>(ws + n_cand * 8);
  {
    ProfScope prof(TCL_K_SIM_TOPK_FUSED, st);
    if (int e = launch_fused<FT_MODE_TOPK>(P, n_qblocks, P.n_split, st)) return e;
  }
  if (int e = tcl_topk_merge(P.out_val, P.out_idx, P.n_split, n_q, k, topk_val, topk_idx, stream)) return e;
  sum_splits_kernel<<<static_cast<unsigned>((n_q + 255) / 256), 256, 0, st>>>(P.out_nb, P.n_split, n_q, n_before);
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}
