// K3 — NT-Xent backward by recomputation.  Replaces the autograd graph of
// tricolo/loss/nt_xent.py:55-74 (softmax backward x2, four matmuls, normalise
// backward x2) with one tcgen05 kernel + one streaming kernel.
//
// "Directional" formulation: a job computes the gradient of ONE tensor (self)
// from the 1-2 pairs it takes part in (segments).  For a segment
//     S   = Zself · Zother^T                      (re-formed in TMEM, never stored)
//     G'  = r * [ w_self 2^(c1 S - lse_self_i) + w_other 2^(c1 S - lse_other_j) - [i==j] ]
//     acc += G' · Zother                           (second MMA, accumulator in TMEM)
// with r = grad_scale_seg / max_seg |grad_scale|.  Then g = acc * gmax / (tau n_other) and
// the normalise backward (l2norm.cu) turns g into dx.  Both directions of a pair
// use the same kernel with the roles of the operands swapped, so in the sharded
// (global-negative) setting no gradient reduce-scatter is needed.
//
// CTA = (i-block of 128 self rows, half of dim, split of the tile range, job).
// The full-dim fp32 accumulator of 128 rows would need all 512 TMEM columns, so
// each CTA owns 256 columns of dim (TMEM: 2 x 128 logit buffers + 256 accumulator).
// Warp roles: 0 TMA producer, 1 MMA issuer, 2..9 epilogue (thread = TMEM lane = row; two warps
// per lane quarter split the 128 logit columns).
#include <stdlib.h>

#include "ntxent_bwd.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

template <int kOp>
__global__ void __launch_bounds__(BW_THREADS, 1) ntxent_bwd_kernel(const __grid_constant__ BwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int num_kb = P.num_kb;
  const uint32_t x_smem = base + BwdSmem::x_off;
  const uint32_t g_smem = base + BwdSmem::g_off(num_kb);
  const uint32_t ring = base + BwdSmem::ring_off(num_kb);
  const uint32_t bars = base + BwdSmem::bar_off(num_kb);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (BW_STAGES + s); };
  const uint32_t x_full_bar = bars + 8u * (2 * BW_STAGES);
  auto s_full_bar = [&](int b) { return bars + 8u * (2 * BW_STAGES + 1 + b); };
  auto s_empty_bar = [&](int b) { return bars + 8u * (2 * BW_STAGES + 3 + b); };
  const uint32_t g_full_bar = bars + 8u * (2 * BW_STAGES + 5);
  const uint32_t g_empty_bar = bars + 8u * (2 * BW_STAGES + 6);
  const uint32_t acc_full_bar = bars + 8u * (2 * BW_STAGES + 7);
  const uint32_t tmem_slot = bars + 8u * (2 * BW_STAGES + 8);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      base_ptr + BwdSmem::bar_off(num_kb) + 8u * (2 * BW_STAGES + 8));
  float* bj = reinterpret_cast<float*>(base_ptr + BwdSmem::bj_off(num_kb));  // [2][128]
  uint8_t* g_ptr = base_ptr + BwdSmem::g_off(num_kb);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ib = blockIdx.x;
  const int dh = blockIdx.y % P.n_dhalf;
  const int split = blockIdx.y / P.n_dhalf;
  const BwdJobDev& J = P.job[blockIdx.z];
  const int i0 = ib * BW_BM;
  const int d0 = dh * BW_DH;
  // d chunks of 128 columns this CTA accumulates (2 unless dim < 512)
  const int n_dc = (P.dim - d0) >= BW_DH ? 2 : ((P.dim - d0) + 127) / 128;
  const int total_tiles = J.n_seg * P.n_jtiles;
  const int t_begin = static_cast<int>((static_cast<int64_t>(total_tiles) * split) / P.n_split);
  const int t_end = static_cast<int>((static_cast<int64_t>(total_tiles) * (split + 1)) / P.n_split);
  const int n_tiles = t_end - t_begin;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&J.tm_self);
    for (int s = 0; s < J.n_seg; ++s) {
      tma_prefetch_desc(&J.seg[s].tm_other);
      tma_prefetch_desc(&J.seg[s].tm_other_t);
    }
    for (int s = 0; s < BW_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(x_full_bar, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full_bar(b), 1);
      mbar_init(s_empty_bar(b), BW_EPI_THREADS);
    }
    mbar_init(g_full_bar, BW_EPI_THREADS);
    mbar_init(g_empty_bar, 1);
    mbar_init(acc_full_bar, 1);
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
  const uint32_t tmem_acc = tmem + 2 * BW_BN;  // columns 256..511

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one() && n_tiles > 0) {
      mbar_arrive_expect_tx(x_full_bar, num_kb * BW_KB_BYTES);
      for (int kb = 0; kb < num_kb; ++kb)
        tma_load_2d(x_smem + kb * BW_KB_BYTES, &J.tm_self, x_full_bar, kb * BW_BK, i0);
      int it = 0;
      auto push = [&](const CUtensorMap* tm, int c0, int c1) {
        const int s = it % BW_STAGES;
        const uint32_t ph = (it / BW_STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(s), BW_KB_BYTES);
        tma_load_2d(ring + s * BW_KB_BYTES, tm, full_bar(s), c0, c1);
        ++it;
      };
      auto load_s = [&](int t) {  // operands of the logit tile t
        const int tt = t_begin + t;
        const BwdSegDev& sg = J.seg[tt / P.n_jtiles];
        const int j0 = (tt % P.n_jtiles) * BW_BN;
        for (int kb = 0; kb < num_kb; ++kb) push(&sg.tm_other, kb * BW_BK, j0);
      };
      auto load_a = [&](int t) {  // operands of the gradient MMAs of tile t
        const int tt = t_begin + t;
        const BwdSegDev& sg = J.seg[tt / P.n_jtiles];
        const int j0 = (tt % P.n_jtiles) * BW_BN;
        for (int kb2 = 0; kb2 < 2; ++kb2)
          for (int dc = 0; dc < n_dc; ++dc) push(&sg.tm_other_t, j0 + kb2 * BW_BK, d0 + dc * 128);
      };
      // same order as the MMA warp consumes: S(0), [S(t+1), A(t)] ...
      load_s(0);
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) load_s(t + 1);
        load_a(t);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one() && n_tiles > 0) {
      mbar_wait(x_full_bar, 0);
      int it = 0;
      auto issue_s = [&](int t) {
        const int b = t & 1;
        mbar_wait(s_empty_bar(b), ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % BW_STAGES;
          const uint32_t ph = (it / BW_STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint64_t ad = umma_desc_k_sw128(x_smem + kb * BW_KB_BYTES);
          const uint64_t bd = umma_desc_k_sw128(ring + s * BW_KB_BYTES);
#pragma unroll
          for (int kk = 0; kk < BW_BK / 16; ++kk)
            tc_mma_f16(tmem + b * BW_BN, ad + 2 * kk, bd + 2 * kk, P.idesc, (kb | kk) != 0);
          tc_commit(empty_bar(s));
        }
        tc_commit(s_full_bar(b));
      };
      issue_s(0);
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) issue_s(t + 1);
        mbar_wait(g_full_bar, t & 1);
        tc_fence_after();
        for (int kb2 = 0; kb2 < 2; ++kb2) {
          for (int dc = 0; dc < n_dc; ++dc, ++it) {
            const int s = it % BW_STAGES;
            const uint32_t ph = (it / BW_STAGES) & 1;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint64_t ad = umma_desc_k_sw128(g_smem + kb2 * BW_KB_BYTES);
            const uint64_t bd = umma_desc_k_sw128(ring + s * BW_KB_BYTES);
#pragma unroll
            for (int kk = 0; kk < BW_BK / 16; ++kk)
              tc_mma_f16(tmem_acc + dc * 128, ad + 2 * kk, bd + 2 * kk, P.idesc,
                         (t | kb2 | kk) != 0);
            tc_commit(empty_bar(s));
          }
        }
        tc_commit(g_empty_bar);
      }
      tc_commit(acc_full_bar);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;          // TMEM lane quarter
    const int ch = (warp - 2) >> 2;  // column half of the logit tile == K-block of the G operand
    const int r = q * 32 + lane;     // tile-local row == TMEM lane
    const int et = threadIdx.x - 64;  // 0..255
    const int grow = i0 + r;          // local self row
    // gradient scales: ratio to the largest magnitude keeps G' inside [-1, 1]
    float gs[2] = {0.f, 0.f};
    float gmax = 0.f;
    for (int s = 0; s < J.n_seg; ++s) {
      gs[s] = J.seg[s].grad_scale ? *J.seg[s].grad_scale : 1.f;
      gmax = fmaxf(gmax, fabsf(gs[s]));
    }
    const float inv_gmax = gmax > 0.f ? kGScale / gmax : 0.f;  // G in [-kGScale, kGScale]: see ntxent_bwd.h
    if (blockIdx.x == 0 && blockIdx.y == 0 && et == 0) *J.scale_out = gmax * P.out_scale * (1.f / kGScale);

    // per-column factors 2^(c1 - lse_other_j) of tile t, fetched one tile ahead
    auto load_bj = [&](int t) -> float {
      if (et >= 128 || t >= n_tiles) return 0.f;
      const int tt = t_begin + t;
      const BwdSegDev& sg = J.seg[tt / P.n_jtiles];
      const int j = (tt % P.n_jtiles) * BW_BN + et;
      return j < P.n_other ? ex2_approx(P.c1 - sg.lse2_other[j]) : 0.f;
    };
    if (et < 128 && n_tiles > 0) bj[et] = load_bj(0);
    int cur_seg = -1;
    float lse_i = 0.f, ws = 0.f, wo_i = 0.f, rr = 0.f;

    for (int t = 0; t < n_tiles; ++t) {
      const int tt = t_begin + t;
      const int si = tt / P.n_jtiles;
      const int j0 = (tt % P.n_jtiles) * BW_BN;
      const int b = t & 1;
      if (si != cur_seg) {  // per-row constants of this segment
        const BwdSegDev& sg = J.seg[si];
        cur_seg = si;
        rr = gs[si] * inv_gmax;
        lse_i = grow < P.n_self ? sg.lse2_self[grow] : 0.f;
        ws = rr * sg.w_self;
        // p_other = p_self * 2^(lse_self_i - c1) * 2^(c1 - lse_other_j)
        wo_i = rr * sg.w_other * ex2_approx(lse_i - P.c1);
      }
      const float bj_next = load_bj(t + 1);  // in flight while this tile is processed
      asm volatile("bar.sync 1, 256;" ::: "memory");  // bj[t&1] visible to all epilogue threads
      const float* bjt = bj + (t & 1) * 128;
      // diagonal of this segment: global self index == other index
      const int dcol = P.self_offset + grow - j0;  // tile-local column of the positive

      mbar_wait(s_full_bar(b), (t >> 1) & 1);
      tc_fence_after();
      uint32_t v[2][32];
      tmem_ld_32x32b_x32(tmem_addr(tmem + b * BW_BN, q * 32, (2 * ch) * 32), v[0]);
      tmem_ld_32x32b_x32(tmem_addr(tmem + b * BW_BN, q * 32, (2 * ch + 1) * 32), v[1]);
      tc_wait_ld();
      // logits are in registers: the TMEM buffer can be refilled
      tc_fence_before();
      mbar_arrive(s_empty_bar(b));
      uint32_t pk[2][16];
#pragma unroll
      for (int cl = 0; cl < 2; ++cl) {
        const int cc = 2 * ch + cl;
        const int dl = dcol - cc * 32;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(v[cl][e]), P.c1, -lse_i));
          const float p1 = ex2_approx(fmaf(__uint_as_float(v[cl][e + 1]), P.c1, -lse_i));
          const float g0 = fmaf(p0, fmaf(wo_i, bjt[cc * 32 + e], ws), (e == dl) ? -rr : 0.f);
          const float g1 = fmaf(p1, fmaf(wo_i, bjt[cc * 32 + e + 1], ws), (e + 1 == dl) ? -rr : 0.f);
          pk[cl][e >> 1] = pack2<kOp>(g0, g1);
        }
      }
      mbar_wait(g_empty_bar, (t & 1) ^ 1);  // previous gradient MMAs finished reading G
      // K-major, 128-byte-swizzled operand tile: row r, 16-byte chunk c16 -> c16 ^ (r & 7)
      uint8_t* gk = g_ptr + ch * BW_KB_BYTES + r * 128;
#pragma unroll
      for (int cl = 0; cl < 2; ++cl) {
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const int c16 = cl * 4 + c4;
          *reinterpret_cast<uint4*>(gk + ((c16 ^ (r & 7)) << 4)) =
              make_uint4(pk[cl][4 * c4], pk[cl][4 * c4 + 1], pk[cl][4 * c4 + 2], pk[cl][4 * c4 + 3]);
        }
      }
      // G visible to the tensor core (async proxy)
      fence_proxy_async_smem();
      mbar_arrive(g_full_bar);
      if (et < 128 && t + 1 < n_tiles) bj[((t + 1) & 1) * 128 + et] = bj_next;
    }

    if (n_tiles > 0) {
      mbar_wait(acc_full_bar, 0);
      tc_fence_after();
    }
    // accumulator read-out: column half ch of this CTA's 256 dim columns
    float* gout = J.gpart + (static_cast<int64_t>(split) * P.n_self + grow) * P.dim + d0;
    if (ch < n_dc) {
#pragma unroll 1
      for (int cc = ch * 4; cc < ch * 4 + 4; ++cc) {
        uint32_t v[32];
        if (n_tiles > 0) {
          tmem_ld_32x32b_x32(tmem_addr(tmem_acc, q * 32, cc * 32), v);
          tc_wait_ld();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0u;
        }
        if (grow < P.n_self && d0 + cc * 32 < P.dim) {  // dim % 128 != 0: the last chunk is partly outside dim
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            *reinterpret_cast<float4*>(gout + cc * 32 + e) =
                make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                            __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

enum class BwdMode { Pc, Indep };
static BwdMode bwd_mode() {  // read once per process: TRICOLO_B200_BWD=indep selects the one-CTA-per-dim-half kernel for dim > 256
  static const BwdMode m = [] {
    const char* e = getenv("TRICOLO_B200_BWD");
    return (e && !strcmp(e, "indep")) ? BwdMode::Indep : BwdMode::Pc;
  }();
  return m;
}

static int bwd_split(int n_jobs, int n_iblocks, int n_dhalf, int min_tiles) {
  // One CTA per SM (224 KB shared memory): pick the split of the tile range that wastes the least
  // of the last wave; every extra split costs one more fp32 partial of the gradient, so a larger
  // split must buy at least 3 % of wave efficiency.
  const int ctas = n_jobs * n_iblocks * n_dhalf;
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= kBwdMaxSplit && s <= min_tiles; ++s) {
    const int total = ctas * s;
    const int waves = (total + kNumSMsB200 - 1) / kNumSMsB200;
    const double eff = static_cast<double>(total) / (static_cast<double>(waves) * kNumSMsB200);
    if (eff > best_eff + 0.03) {
      best_eff = eff;
      best = s;
    }
  }
  return best;
}

}  // namespace tcl

using namespace tcl;

extern "C" size_t tcl_ntxent_bwd_workspace_bytes(int n_jobs, int64_t n_self, int64_t dim) {
  if (n_jobs < 1 || n_self < 1 || dim < 1) return 0;
  const int64_t n_self_pad = (n_self + BW_BM - 1) / BW_BM * BW_BM;  // the persistent kernel stores whole 128-row units
  // [64 floats of scales][n_jobs x kBwdMaxSplit partial slots][block counters of the folded normalise backward]
  return static_cast<size_t>(n_jobs) * kBwdMaxSplit * n_self_pad * dim * sizeof(float) + 256 +
         sizeof(uint32_t) * 8 * static_cast<size_t>(n_self_pad / BW_BM) + 256;
}

extern "C" int tcl_ntxent_bwd_needs_transpose(int64_t dim) {
  // only the producer/consumer kernel (dim > 256, default) reads the row-major operand directly
  return (dim > BW_DH && bwd_mode() == BwdMode::Pc) ? 0 : 1;
}

extern "C" int tcl_ntxent_bwd(int n_jobs, const tcl_bwd_job* jobs, int64_t n_self, int64_t n_other,
                              int64_t dim, int64_t z_row_stride, int64_t self_offset, int64_t ld_t, int x_dtype,
                              int64_t x_row_stride, int op_format, float inv_tau, float eps,
                              void* workspace, size_t workspace_bytes, void* stream) {
  TCL_REQUIRE(n_jobs >= 1 && n_jobs <= TCL_MAX_TENSORS && jobs, TCL_ERR_BAD_ARG, "ntxent_bwd: n_jobs %d", n_jobs);
  TCL_REQUIRE(n_self >= 1 && n_other >= 1 && n_self < (1 << 24) && n_other < (1 << 24), TCL_ERR_BAD_SHAPE, "ntxent_bwd: sizes");
  TCL_REQUIRE(dim >= 64 && dim % 64 == 0 && dim <= 512, TCL_ERR_BAD_SHAPE,
              "ntxent_bwd: dim must be a multiple of 64 in [64, 512] (got %lld)", (long long)dim);
  TCL_REQUIRE(self_offset >= 0 && self_offset + n_self <= n_other, TCL_ERR_BAD_SHAPE, "ntxent_bwd: self rows outside the global batch");
  const bool need_t = tcl_ntxent_bwd_needs_transpose(dim) != 0;
  TCL_REQUIRE(!need_t || (ld_t >= n_other && ld_t % 8 == 0), TCL_ERR_BAD_ALIGN, "ntxent_bwd: ld_t must be >= n_other and a multiple of 8");
  TCL_REQUIRE(op_format == TCL_OP_F16 || op_format == TCL_OP_BF16, TCL_ERR_BAD_ARG, "op_format %d", op_format);
  TCL_REQUIRE(workspace && workspace_bytes >= tcl_ntxent_bwd_workspace_bytes(n_jobs, n_self, dim), TCL_ERR_WORKSPACE, "ntxent_bwd: workspace too small");
  const float c1 = inv_tau * 1.4426950408889634f;
  TCL_REQUIRE(inv_tau > 0.f && 2.f * c1 < 120.f, TCL_ERR_BAD_ARG, "ntxent_bwd: temperature too small (need tau >= 0.025)");
  if (z_row_stride == 0) z_row_stride = dim;
  TCL_REQUIRE(z_row_stride >= dim && z_row_stride % 8 == 0, TCL_ERR_BAD_ALIGN, "ntxent_bwd: z_row_stride");
  if (int e = require_sm100()) return e;

  BwdParams P;
  memset(&P, 0, sizeof(P));
  NormBwdParams N;
  memset(&N, 0, sizeof(N));
  P.n_self = (int)n_self; P.n_other = (int)n_other; P.self_offset = (int)self_offset; P.dim = (int)dim;
  P.num_kb = (int)(dim / 64);
  P.z_row_stride = z_row_stride;
  P.n_jtiles = (int)((n_other + BW_BN - 1) / BW_BN);
  P.n_dhalf = (int)((dim + BW_DH - 1) / BW_DH);
  const int n_iblocks = (int)((n_self + BW_BM - 1) / BW_BM);
  P.n_iblocks = n_iblocks;
  P.n_self_pad = n_iblocks * BW_BM;
  int min_seg = 2;
  for (int j = 0; j < n_jobs; ++j) min_seg = jobs[j].n_segments < min_seg ? jobs[j].n_segments : min_seg;
  if (min_seg < 1) min_seg = 1;
  P.c1 = c1;
  P.out_scale = inv_tau / static_cast<float>(n_other);
  P.idesc = umma_idesc_f16(BW_BM, BW_BN, op_format);
  // dim > 256 has two implementations behind this entry (TRICOLO_B200_BWD=pc|indep):
  //  pc      (default) producer/consumer 2-CTA cluster, ntxent_bwd_pc.cu: logit recompute on one SM, gradient GEMM
  //          with a full-dim accumulator on the other; 8 B^2 D executed per pair.
  //  indep   the kernel above (the only one for dim <= 256): one CTA per dim half, logits recomputed per half
  //          (12 B^2 D executed).
  const bool use_pc = P.n_dhalf == 2 && bwd_mode() != BwdMode::Indep;
  P.idesc_n256 = umma_idesc_f16(BW_BM, 256, op_format) | (1u << 16);  // B operand MN-major
  P.n_split = bwd_split(n_jobs, n_iblocks, P.n_dhalf, min_seg * P.n_jtiles);
  float* ws = static_cast<float*>(workspace);
  float* scales = ws;  // 64 floats reserved
  float* gbase = ws + 64;
  for (int j = 0; j < n_jobs; ++j) {
    const tcl_bwd_job& src = jobs[j];
    BwdJobDev& J = P.job[j];
    TCL_REQUIRE(src.n_segments >= 1 && src.n_segments <= 2, TCL_ERR_BAD_ARG, "ntxent_bwd: job %d has %d segments", j, src.n_segments);
    TCL_REQUIRE(src.z_self && src.x_self && src.inv_norm && src.dx, TCL_ERR_BAD_ARG, "ntxent_bwd: null pointer in job %d", j);
    if (int e = make_tmap_2d_16bit(&J.tm_self, src.z_self, n_self, dim, z_row_stride, BW_BM, BW_BK)) return e;
    J.n_seg = src.n_segments;
    J.z_self = static_cast<const uint16_t*>(src.z_self);
    for (int s = 0; s < src.n_segments; ++s) {
      const tcl_bwd_segment& sg = src.seg[s];
      TCL_REQUIRE(sg.z_other && (sg.z_other_t || !need_t) && sg.lse2_self && sg.lse2_other, TCL_ERR_BAD_ARG, "ntxent_bwd: null pointer in job %d segment %d", j, s);
      if (int e = make_tmap_2d_16bit(&J.seg[s].tm_other, sg.z_other, n_other, dim, z_row_stride, BW_BN, BW_BK)) return e;
      if (use_pc) {  // the gradient GEMM reads the row-major operand itself (MN-major B): boxes {64 dim, 64 rows}
        if (int e = make_tmap_2d_16bit(&J.seg[s].tm_other_t, sg.z_other, n_other, dim, z_row_stride, 64, 64)) return e;
      } else {
        if (int e = make_tmap_2d_16bit(&J.seg[s].tm_other_t, sg.z_other_t, dim, n_other, ld_t, 128, BW_BK)) return e;
      }
      J.seg[s].lse2_self = sg.lse2_self;
      J.seg[s].lse2_other = sg.lse2_other;
      J.seg[s].grad_scale = sg.grad_scale;
      J.seg[s].w_self = sg.w_self;
      J.seg[s].w_other = sg.w_other;
    }
    J.gpart = gbase + static_cast<size_t>(j) * kBwdMaxSplit * P.n_self_pad * dim;
    J.scale_out = scales + j;
    P.unit_tiles[j] = src.n_segments * P.n_jtiles;
    P.job_tile_base[j + 1] = P.job_tile_base[j] + static_cast<int64_t>(n_iblocks) * P.unit_tiles[j];
    if (use_pc)
      if (int e = make_tmap_2d_f32(&P.tm_gpart[j], J.gpart, static_cast<uint64_t>(kBwdMaxSplit) * P.n_self_pad, dim, 32, 32)) return e;
    N.job[j].x = src.x_self;
    N.job[j].inv_norm = src.inv_norm;
    N.job[j].gpart = J.gpart;
    N.job[j].scale = J.scale_out;
    N.job[j].dx = src.dx;
  }
  for (int j = n_jobs; j < TCL_MAX_TENSORS; ++j) P.job_tile_base[j + 1] = P.job_tile_base[n_jobs];
  const int smem = (int)BwdSmem::total(P.num_kb);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(n_iblocks, P.n_dhalf * P.n_split, n_jobs);
  prof_begin(TCL_K_NTXENT_BWD, st);
  if (use_pc) {
    int n_clusters = 0;
    if (int e = launch_bwd_pc(P, n_jobs, op_format, &n_clusters, st)) return e;
    // partial count per 128-row unit follows from the tile ranges (ntxent_bwd.h: pc_range_of)
    N.n_clusters = n_clusters;
    N.split_rows = P.n_self_pad;
    N.total_tiles = P.job_tile_base[TCL_MAX_TENSORS];
    for (int j = 0; j < TCL_MAX_TENSORS; ++j) {
      N.job_tile_base[j] = P.job_tile_base[j];
      N.unit_tiles[j] = P.unit_tiles[j];
    }
    P.n_split = kBwdMaxSplit;
  } else if (op_format == TCL_OP_F16) {
    if (int e = ensure_dyn_smem(ntxent_bwd_kernel<TCL_OP_F16>, smem)) return e;
    ntxent_bwd_kernel<TCL_OP_F16><<<grid, BW_THREADS, smem, st>>>(P);
  } else {
    if (int e = ensure_dyn_smem(ntxent_bwd_kernel<TCL_OP_BF16>, smem)) return e;
    ntxent_bwd_kernel<TCL_OP_BF16><<<grid, BW_THREADS, smem, st>>>(P);
  }
  prof_end(TCL_K_NTXENT_BWD, st);
  TCL_CHECK_CUDA(cudaGetLastError());
  return launch_l2norm_bwd(N, n_jobs, x_dtype, n_self, (int)dim, x_row_stride, P.n_split, eps, st);
}
