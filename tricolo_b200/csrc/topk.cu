// K4 — per-query top-k and rank of the ground-truth shape over a materialised
// fp32 similarity matrix, plus the shard merge.  HBM-bound: each similarity is
// read exactly once (16-byte loads, one warp per query row), nothing but the
// k winners and two scalars per query is written.
//
// Replaces the two full sorts and the Python rank search of
// tricolo/evaluation/eval_retrieval.py:75-82 and :184-186 / :217-219.
// Order (the stated tie-break): similarity descending, gallery index ascending.
#include <float.h>
#include <limits.h>

#include "host_common.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

// (a, ia) ranks before (b, ib)?
__device__ __forceinline__ bool before(float a, int ia, float b, int ib) {
  return a > b || (a == b && ia < ib);
}

// warp-wide fp32 maximum in one instruction (sm_100a: redux.sync on f32 -> CREDUX.MAX.F32)
__device__ __forceinline__ float warp_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

template <int K>
struct TopList {
  float v[K];
  int i[K];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int t = 0; t < K; ++t) { v[t] = -FLT_MAX; i[t] = INT_MAX; }
  }
  // Insert a candidate whose index is larger than every index already seen by
  // this lane: ties therefore never displace an earlier entry.
  __device__ __forceinline__ void push(float x, int idx) {
    if (!(x > v[K - 1])) return;
    v[K - 1] = x; i[K - 1] = idx;
#pragma unroll
    for (int t = K - 1; t > 0; --t) {
      if (v[t] > v[t - 1]) {
        float tv = v[t]; v[t] = v[t - 1]; v[t - 1] = tv;
        int ti = i[t]; i[t] = i[t - 1]; i[t - 1] = ti;
      }
    }
  }
  __device__ __forceinline__ void pop() {
#pragma unroll
    for (int t = 0; t < K - 1; ++t) { v[t] = v[t + 1]; i[t] = i[t + 1]; }
    v[K - 1] = -FLT_MAX; i[K - 1] = INT_MAX;
  }
};

template <int K>
__global__ void __launch_bounds__(256) topk_rank_kernel(
    const float* __restrict__ s, int64_t ld, int64_t n_q, int n_g, int k,
    const int64_t* __restrict__ labels, int64_t idx_base, const float* __restrict__ gt_sim_in,
    float* __restrict__ topk_val, int32_t* __restrict__ topk_idx, float* __restrict__ gt_sim_out,
    int32_t* __restrict__ n_before) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (q >= n_q) return;
  const float* row = s + q * ld;

  const int64_t gt_global = labels[q];
  const int64_t gt_local64 = gt_global - idx_base;
  float s_gt;
  if (gt_sim_in != nullptr) {
    s_gt = gt_sim_in[q];
  } else {
    // the label must fall inside this shard when no external value is given
    s_gt = (gt_local64 >= 0 && gt_local64 < n_g) ? row[gt_local64] : -FLT_MAX;
  }
  // columns strictly below gt_cut tie-break ahead of the ground truth
  const int gt_cut = gt_local64 < 0 ? 0 : (gt_local64 > n_g ? n_g : static_cast<int>(gt_local64));

  TopList<K> top;
  top.init();
  int cnt = 0;

  const int n_vec = n_g >> 2;  // full float4 groups
  const float4* row4 = reinterpret_cast<const float4*>(row);
  int g4 = lane;
  // software-pipelined loads (the next 4 x 16 bytes per lane are requested before the current 16 values are consumed)
  // and a warp-wide entry threshold (see refresh() below): a value <= thr can no longer enter the row's top-k, so the
  // per-lane sorted lists are only touched by the few values that can.
  float thr = -FLT_MAX;
  int cnt_eq = 0;
  auto consume = [&](const float4& a, const float4& b, const float4& c, const float4& d, int base4) {
    const float vals[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
    // common case in ~3 instructions per value: one compare-and-add for the rank, one equality folded into a
    // predicate, one max; ties with the ground truth and candidates above the threshold take the slow paths below
    bool eq = false;
    float m = vals[0];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const float x = vals[u];
      cnt += x > s_gt;
      eq |= x == s_gt;
      m = fmaxf(m, x);
    }
    if (eq) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int col = (base4 + (u >> 2) * 32) * 4 + (u & 3);
        cnt_eq += (vals[u] == s_gt) && (col < gt_cut);
      }
    }
    if (m > thr) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int col = (base4 + (u >> 2) * 32) * 4 + (u & 3);
        if (vals[u] > thr) top.push(vals[u], col);
      }
    }
  };
  // EXACT k-th largest value of everything the warp has consumed so far, without touching the lists: k rounds of a
  // warp-wide max over per-lane cursors into the (sorted) lane lists.  Called between chunks only, when every consumed
  // column is smaller than every column still to come, so a later value <= thr loses against k earlier candidates
  // (ties included) and can be skipped.  The max over lanes of the lanes' own k-th value used before is a far weaker
  // bar (the k-th best of 1/32 of the row): ncu showed 28 warp instructions per value on 25k-column rows, i.e. the
  // kernel was issue-bound by divergent insertions (sm 84 %, dram 54 %).
  auto refresh = [&]() {
    int p = 0;
    float t = -FLT_MAX;
    for (int r = 0; r < k; ++r) {
      float c = -FLT_MAX;
#pragma unroll
      for (int j = 0; j < K; ++j) c = (p == j) ? top.v[j] : c;
      const float mx = warp_max_f32(c);
      const unsigned who = __ballot_sync(0xffffffffu, c == mx);
      if (lane == __ffs(who) - 1) ++p;
      t = mx;
    }
    thr = t;
  };
  // warp-uniform trip count (the refresh shuffles over the full warp): whole 512-column chunks only
  const int n_full = n_vec >> 7;  // chunks of 128 float4 in which every lane has all four loads
  if (n_full > 0) {
    float4 a = __ldcs(row4 + g4), b = __ldcs(row4 + g4 + 32), c = __ldcs(row4 + g4 + 64), d = __ldcs(row4 + g4 + 96);
    for (int ch = 0; ch < n_full; ++ch, g4 += 128) {
      float4 na = a, nb = b, nc = c, nd = d;
      if (ch + 1 < n_full) {
        na = __ldcs(row4 + g4 + 128); nb = __ldcs(row4 + g4 + 160); nc = __ldcs(row4 + g4 + 192); nd = __ldcs(row4 + g4 + 224);
      }
      // refresh after 1, 2, 4, 8 chunks, then every 8: the expected number of later values above the exact k-th of
      // c chunks is k * (chunks until the next refresh) / c
      if (ch != 0 && ((ch & (ch - 1)) == 0 || (ch & 7) == 0)) refresh();
      consume(a, b, c, d, g4);
      a = na; b = nb; c = nc; d = nd;
    }
    refresh();
  }
  for (; g4 < n_vec; g4 += 32) {
    float4 a = __ldcs(row4 + g4);
    const float vals[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int col = g4 * 4 + u;
      const float x = vals[u];
      cnt += (x > s_gt) || (x == s_gt && col < gt_cut);
      if (x > thr) top.push(x, col);
    }
  }
  {  // scalar tail (n_g % 4 columns); still ascending per lane: only lane 31 .. no: use lane order
    const int col = (n_vec << 2) + lane;
    if (col < n_g) {
      const float x = __ldg(row + col);
      cnt += (x > s_gt) || (x == s_gt && col < gt_cut);
      // tail columns are larger than any column this lane has seen
      top.push(x, col);
    }
  }

  // rank: total count
  cnt += cnt_eq;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);

  // k rounds of warp arg-best over the lane heads
  for (int r = 0; r < k; ++r) {
    float bv = top.v[0];
    int bi = top.i[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (before(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    if (top.i[0] == bi && bi != INT_MAX) top.pop();  // indices are unique across lanes
    if (lane == 0) {
      topk_val[q * k + r] = bv;
      topk_idx[q * k + r] = bi == INT_MAX ? -1 : static_cast<int32_t>(idx_base + bi);
    }
  }
  if (lane == 0) {
    n_before[q] = cnt;
    if (gt_sim_out != nullptr && gt_sim_in == nullptr) gt_sim_out[q] = s_gt;
  }
}

__global__ void gather_gt_kernel(const float* __restrict__ s, int64_t ld, int64_t n_q, int64_t n_g,
                                 const int64_t* __restrict__ labels, int64_t idx_base,
                                 float* __restrict__ gt_sim) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n_q) return;
  const int64_t l = labels[q] - idx_base;
  gt_sim[q] = (l >= 0 && l < n_g) ? s[q * ld + l] : 0.f;
}

// one thread per query: k rounds of picking the best head of n_shards sorted lists
__global__ void topk_merge_kernel(const float* __restrict__ cv, const int32_t* __restrict__ ci,
                                  int n_shards, int64_t n_q, int k, float* __restrict__ ov,
                                  int32_t* __restrict__ oi) {
  const int64_t q = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= n_q) return;
  int head[16];
#pragma unroll
  for (int s = 0; s < 16; ++s) head[s] = 0;
  for (int r = 0; r < k; ++r) {
    float bv = -FLT_MAX;
    int bi = INT_MAX, bs = -1;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      if (s < n_shards && head[s] < k) {
        const int64_t off = (static_cast<int64_t>(s) * n_q + q) * k + head[s];
        const float v = cv[off];
        const int i = ci[off];
        if (i >= 0 && (bs < 0 || before(v, i, bv, bi))) { bv = v; bi = i; bs = s; }
      }
    }
#pragma unroll
    for (int s = 0; s < 16; ++s)
      if (s == bs) head[s]++;
    ov[q * k + r] = bv;
    oi[q * k + r] = bs < 0 ? -1 : bi;
  }
}

// K5: metric reduction from the ranks of the ground truth (with one relevant gallery item per query all
// of RR@k / NDCG@k / precision / recall / MRR are functions of the rank): count[j] = #{rank == j+1}, j < k, and
// sum 1/rank in fp64.  Two levels in one launch, both in a fixed order (per-block partials, then the block that
// finishes last adds them in block order): deterministic.
static constexpr int RM_MAX_K = 16;
__global__ void __launch_bounds__(256) rank_metrics_kernel(const int32_t* __restrict__ rank, int64_t n_q, int k,
                                                           double* __restrict__ partial, unsigned int* __restrict__ counter,
                                                           double* __restrict__ out) {
  __shared__ double sh[8][RM_MAX_K + 1];
  __shared__ bool last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned int cnt[RM_MAX_K];
#pragma unroll
  for (int j = 0; j < RM_MAX_K; ++j) cnt[j] = 0;
  double rr = 0.0;
  // contiguous chunk per thread so that the fp64 sum order does not depend on the launch geometry's interleaving
  const int64_t per_block = (n_q + gridDim.x - 1) / gridDim.x;
  const int64_t b0 = per_block * blockIdx.x, b1 = b0 + per_block < n_q ? b0 + per_block : n_q;
  for (int64_t i = b0 + threadIdx.x; i < b1; i += blockDim.x) {
    const int r = rank[i];
    rr += 1.0 / static_cast<double>(r);
#pragma unroll
    for (int j = 0; j < RM_MAX_K; ++j) cnt[j] += (j < k && r == j + 1) ? 1u : 0u;
  }
  double v[RM_MAX_K + 1];
#pragma unroll
  for (int j = 0; j < RM_MAX_K; ++j) v[j] = static_cast<double>(cnt[j]);
  v[RM_MAX_K] = rr;
#pragma unroll
  for (int j = 0; j <= RM_MAX_K; ++j) {
    double x = v[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) sh[warp][j] = x;
  }
  __syncthreads();
  if (threadIdx.x <= RM_MAX_K) {
    double x = 0.0;
    for (int w = 0; w < 8; ++w) x += sh[w][threadIdx.x];
    partial[static_cast<int64_t>(blockIdx.x) * (RM_MAX_K + 1) + threadIdx.x] = x;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    last = atomicAdd(counter, 1u) + 1u == gridDim.x;
    if (last) *counter = 0u;
  }
  __syncthreads();
  if (last && threadIdx.x <= RM_MAX_K) {
    __threadfence();
    double x = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) x += __ldcg(partial + static_cast<int64_t>(b) * (RM_MAX_K + 1) + threadIdx.x);
    if (threadIdx.x < k) out[threadIdx.x] = x;
    if (threadIdx.x == RM_MAX_K) out[k] = x;
  }
}

}  // namespace tcl

using namespace tcl;

static constexpr int kRankMetricsBlocks = 592;  // 4 per SM

extern "C" size_t tcl_rank_metrics_workspace_bytes(void) {
  return sizeof(double) * kRankMetricsBlocks * (RM_MAX_K + 1) + 256;
}

extern "C" int tcl_rank_metrics(const int32_t* rank, int64_t n_q, int k, double* out, void* workspace,
                                size_t workspace_bytes, void* stream) {
  TCL_REQUIRE(rank && out && workspace, TCL_ERR_BAD_ARG, "rank_metrics: null pointer");
  TCL_REQUIRE(k >= 1 && k <= RM_MAX_K && n_q >= 1, TCL_ERR_BAD_ARG, "rank_metrics: k %d, n_q %lld", k, (long long)n_q);
  TCL_REQUIRE(workspace_bytes >= tcl_rank_metrics_workspace_bytes() && aligned_to(workspace, 256), TCL_ERR_WORKSPACE,
              "rank_metrics: workspace (zero-initialised once by the caller, 256-byte aligned)");
  if (int e = require_sm100()) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int64_t blocks = (n_q + 2047) / 2048;
  if (blocks > kRankMetricsBlocks) blocks = kRankMetricsBlocks;
  unsigned int* counter = static_cast<unsigned int*>(workspace);
  double* partial = reinterpret_cast<double*>(static_cast<char*>(workspace) + 256);
  ProfScope prof(TCL_K_RANK_METRICS, st);
  rank_metrics_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(rank, n_q, k, partial, counter, out);
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

extern "C" int tcl_topk_rank(const float* s, int64_t ld_s, int64_t n_q, int64_t n_g, int k,
                             const int64_t* labels, int64_t idx_base, const float* gt_sim_in,
                             float* topk_val, int32_t* topk_idx, float* gt_sim_out,
                             int32_t* n_before, void* stream) {
  TCL_REQUIRE(k >= 1 && k <= 16, TCL_ERR_BAD_ARG, "topk: k must be in [1,16] (got %d)", k);
  TCL_REQUIRE(n_q >= 0 && n_g >= 1 && n_g < (1LL << 31) - 256, TCL_ERR_BAD_SHAPE, "topk: n_g %lld", (long long)n_g);
  TCL_REQUIRE(ld_s >= n_g && ld_s % 4 == 0 && aligned_to(s, 16), TCL_ERR_BAD_ALIGN,
              "topk: similarity rows must be 16-byte aligned (ld_s %% 4 == 0)");
  TCL_REQUIRE(s && labels && topk_val && topk_idx && n_before, TCL_ERR_BAD_ARG, "topk: null pointer");
  if (int e = require_sm100()) return e;
  if (n_q == 0) return TCL_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = static_cast<unsigned>((n_q + 7) / 8);
  ProfScope prof(TCL_K_TOPK_RANK, st);
  if (k <= 5)
    topk_rank_kernel<5><<<grid, 256, 0, st>>>(s, ld_s, n_q, (int)n_g, k, labels, idx_base, gt_sim_in,
                                              topk_val, topk_idx, gt_sim_out, n_before);
  else
    topk_rank_kernel<16><<<grid, 256, 0, st>>>(s, ld_s, n_q, (int)n_g, k, labels, idx_base, gt_sim_in,
                                               topk_val, topk_idx, gt_sim_out, n_before);
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

extern "C" int tcl_gather_gt_sim(const float* s, int64_t ld_s, int64_t n_q, int64_t n_g,
                                 const int64_t* labels, int64_t idx_base, float* gt_sim,
                                 void* stream) {
  TCL_REQUIRE(s && labels && gt_sim, TCL_ERR_BAD_ARG, "gather_gt: null pointer");
  TCL_REQUIRE(n_q >= 0 && n_g >= 1 && ld_s >= n_g, TCL_ERR_BAD_SHAPE, "gather_gt: shape");
  if (int e = require_sm100()) return e;
  if (n_q == 0) return TCL_OK;
  ProfScope prof(TCL_K_GATHER_GT, static_cast<cudaStream_t>(stream));
  gather_gt_kernel<<<static_cast<unsigned>((n_q + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      s, ld_s, n_q, n_g, labels, idx_base, gt_sim);
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

extern "C" int tcl_topk_merge(const float* cand_val, const int32_t* cand_idx, int n_shards,
                              int64_t n_q, int k, float* topk_val, int32_t* topk_idx,
                              void* stream) {
  TCL_REQUIRE(n_shards >= 1 && n_shards <= 16, TCL_ERR_BAD_ARG, "merge: n_shards must be in [1,16]");
  TCL_REQUIRE(k >= 1 && k <= 16, TCL_ERR_BAD_ARG, "merge: k must be in [1,16]");
  TCL_REQUIRE(cand_val && cand_idx && topk_val && topk_idx, TCL_ERR_BAD_ARG, "merge: null pointer");
  if (int e = require_sm100()) return e;
  if (n_q == 0) return TCL_OK;
  ProfScope prof(TCL_K_TOPK_MERGE, static_cast<cudaStream_t>(stream));
  topk_merge_kernel<<<static_cast<unsigned>((n_q + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      cand_val, cand_idx, n_shards, n_q, k, topk_val, topk_idx);
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}
