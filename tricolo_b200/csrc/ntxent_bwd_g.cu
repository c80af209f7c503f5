// K3, shared-G form (one GPU, dim > 256): the backward of tricolo/loss/nt_xent.py:55-74 with the softmax-gradient
// matrix of a pair formed ONCE and used for both of the pair's tensors.
//
// The producer/consumer kernel (ntxent_bwd_pc.cu) recomputes the logits once per direction: 8 B^2 D executed flop
// per pair for 4 B^2 D algorithmic.  On one GPU both directions of a pair need the SAME matrix
//     G = r [ alpha softmax_rows(Z) + (1 - alpha) softmax_cols(Z) - I ]          (dRow = G Zcol, dCol = G^T Zrow)
// so here it is written once, 16-bit, to global memory (B x B x 2 bytes per pair; at B = 8192 it mostly stays in the
// 126 MB L2 between the two kernels) and read back by plain gradient GEMMs: 6 B^2 D executed.
//   kernel A  ntxent_g_kernel     every SM is a "producer" of ntxent_bwd_pc.cu: self rows in TMEM, logit tile on
//                                 tcgen05, the same epilogue arithmetic, G tile staged in the swizzled operand layout
//                                 and written with TMA stores (two 16 KB boxes per tile).
//   kernel B  ntxent_ggemm_kernel every SM is a "consumer": acc[128 x dim] += A[128 x 128] * Zother[128 x dim], A a G
//                                 tile loaded by TMA - K-major when the tensor is the pair's row side, MN-major (the
//                                 transposed read of the same row-major G) when it is the column side - B the other
//                                 operand MN-major straight from the row-major tensor; persistent tile ranges and the
//                                 TMA-store drain of the producer/consumer kernel.
// A sharded run cannot use this form (a rank's two directions need different blocks of G): it keeps ntxent_bwd_pc.cu.
#include <stdlib.h>

#include "ntxent_bwd.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

static constexpr int GA_PSTAGES = 5;     // kernel A ring: slots of two 16 KB K-blocks
static constexpr int GA_SLOT = 32768;
static constexpr int GA_EPI_WARPS = 16;  // two groups of 8, one staging slot each
static constexpr int GA_THREADS = 64 + GA_EPI_WARPS * 32;
static constexpr int GA_XCOL = 256;
struct GASmem {
  static constexpr uint32_t stage_off = 0;                 // 2 x 32 KB
  static constexpr uint32_t ring_off = 2 * GA_SLOT;        // 5 x 32 KB
  static constexpr uint32_t bar_off = (2 + GA_PSTAGES) * GA_SLOT;
  static constexpr uint32_t bj_off = bar_off + 256;        // [2 groups][128] floats
  static constexpr uint32_t total = bj_off + 1024 + 1024;
};
static_assert(GASmem::total <= 232448, "shared-G backward, kernel A: shared memory budget");

struct GPairDev {
  CUtensorMap tm_row;  // row operand [B, dim], box {64, 128}: the unit's self rows (staged through the ring)
  CUtensorMap tm_col;  // column operand [B, dim], box {64, 128}
  CUtensorMap tm_g;    // G [B, ld_g] 16-bit, box {64, 128}
  const float* lse_row;
  const float* lse_col;
  const float* grad_scale;
};
struct GAParams {
  GPairDev pair[TCL_MAX_PAIRS];
  float* scale_out;  // device scalar: max|grad_scale| / (tau B), consumed by the normalise backward
  int n_pairs, batch, num_kb, n_jtiles, n_iblocks;
  float c1, alpha, out_scale;
  uint32_t idesc;
};

struct GWalk {  // equal contiguous tile ranges; unit = (pair, 128-row block), T tiles each
  int64_t cursor, end;
  int T;
  __device__ GWalk(int64_t total, int T_) : T(T_) {
    cursor = pc_range_lo(total, blockIdx.x, gridDim.x);
    end = pc_range_lo(total, blockIdx.x + 1, gridDim.x);
  }
  __device__ bool next(int& unit, int& ta, int& tb) {
    if (cursor >= end) return false;
    unit = static_cast<int>(cursor / T);
    ta = static_cast<int>(cursor % T);
    const int64_t left = end - cursor;
    tb = left < T - ta ? ta + static_cast<int>(left) : T;
    cursor += tb - ta;
    return true;
  }
};

template <int kOp>
__global__ void __launch_bounds__(GA_THREADS, 1) ntxent_g_kernel(const __grid_constant__ GAParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int num_kb = P.num_kb;
  const uint32_t bars = base + GASmem::bar_off;
  auto p_full = [&](int s) { return bars + 8u * s; };
  auto p_empty = [&](int s) { return bars + 8u * (GA_PSTAGES + s); };
  const uint32_t x_full_bar = bars + 8u * (2 * GA_PSTAGES);
  auto s_full = [&](int b) { return bars + 8u * (2 * GA_PSTAGES + 1 + b); };
  auto s_empty = [&](int b) { return bars + 8u * (2 * GA_PSTAGES + 3 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * GA_PSTAGES + 5);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + GASmem::bar_off + 8u * (2 * GA_PSTAGES + 5));
  float* bj_all = reinterpret_cast<float*>(base_ptr + GASmem::bj_off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x_slots = (num_kb + 1) / 2;
  const uint32_t ring = base + GASmem::ring_off;
  const uint32_t stage = base + GASmem::stage_off;

  if (warp == 0 && elect_one()) {
    for (int p = 0; p < P.n_pairs; ++p) {
      tma_prefetch_desc(&P.pair[p].tm_row);
      tma_prefetch_desc(&P.pair[p].tm_col);
      tma_prefetch_desc(&P.pair[p].tm_g);
    }
    for (int s = 0; s < GA_PSTAGES; ++s) {
      mbar_init(p_full(s), 1);
      mbar_init(p_empty(s), 1);
    }
    mbar_init(x_full_bar, GA_EPI_WARPS);
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full(b), 1);
      mbar_init(s_empty(b), GA_EPI_WARPS / 2);
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
  const uint32_t tmem_x = tmem + GA_XCOL;
  const int64_t total = static_cast<int64_t>(P.n_pairs) * P.n_iblocks * P.n_jtiles;
  GWalk walk(total, P.n_jtiles);
  int unit, ta, tb;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA warp
    if (elect_one()) {
      uint32_t it = 0;
      while (walk.next(unit, ta, tb)) {
        const GPairDev& G = P.pair[unit / P.n_iblocks];
        const int i0 = (unit % P.n_iblocks) * BW_BM;
        for (int kb = 0; kb < num_kb; kb += 2, ++it) {  // the piece's self rows: staging for the copy into TMEM
          const int nk = kb + 1 < num_kb ? 2 : 1;
          const int s = it % GA_PSTAGES;
          mbar_wait(p_empty(s), ((it / GA_PSTAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(p_full(s), nk * BW_KB_BYTES);
          for (int k2 = 0; k2 < nk; ++k2)
            tma_load_2d(ring + s * GA_SLOT + k2 * BW_KB_BYTES, &G.tm_row, p_full(s), (kb + k2) * BW_BK, i0);
        }
        for (int t = ta; t < tb; ++t) {
          const int j0 = t * BW_BN;
          for (int kb = 0; kb < num_kb; kb += 2, ++it) {
            const int nk = kb + 1 < num_kb ? 2 : 1;
            const int s = it % GA_PSTAGES;
            mbar_wait(p_empty(s), ((it / GA_PSTAGES) & 1) ^ 1);
            mbar_arrive_expect_tx(p_full(s), nk * BW_KB_BYTES);
            for (int k2 = 0; k2 < nk; ++k2)
              tma_load_2d(ring + s * GA_SLOT + k2 * BW_KB_BYTES, &G.tm_col, p_full(s), (kb + k2) * BW_BK, j0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- logit MMAs (A from TMEM)
    if (elect_one()) {
      uint32_t it = 0, tg = 0, piece = 0;
      while (walk.next(unit, ta, tb)) {
        mbar_wait(x_full_bar, piece & 1);
        tc_fence_after();
        it += static_cast<uint32_t>(x_slots);
        for (int t = ta; t < tb; ++t, ++tg) {
          const int b = tg & 1;
          mbar_wait(s_empty(b), ((tg >> 1) & 1) ^ 1);
          tc_fence_after();
          for (int kb = 0; kb < num_kb; kb += 2, ++it) {
            const int nk = kb + 1 < num_kb ? 2 : 1;
            const int s = it % GA_PSTAGES;
            mbar_wait(p_full(s), (it / GA_PSTAGES) & 1);
            tc_fence_after();
            for (int k2 = 0; k2 < nk; ++k2) {
              const uint32_t ax = tmem_x + (kb + k2) * (BW_BK / 2);
              const uint64_t bd = umma_desc_k_sw128(ring + s * GA_SLOT + k2 * BW_KB_BYTES);
#pragma unroll
              for (int kk = 0; kk < BW_BK / 16; ++kk)
                tc_mma_f16_ts(tmem + b * BW_BN, ax + 8 * kk, bd + 2 * kk, P.idesc, (kb | k2 | kk) != 0);
            }
            tc_commit(p_empty(s));
          }
          tc_commit(s_full(b));
        }
        ++piece;
      }
    }
  } else {
    // ---------------------------------------------------------------- G epilogue (2 groups x 8 warps)
    const int ew = warp - 2;
    const int gi = ew >> 3;
    const int q = warp & 3;
    const int ch = (ew >> 2) & 1;    // column half of the logit tile == 64-column box of the G tile
    const int r = q * 32 + lane;
    const int gt = (ew & 7) * 32 + lane;
    const int bar_grp = 1 + gi;
    const int bar_kb = 3 + 2 * gi + ch;
    float* bj = bj_all + gi * 128;
    const bool issuer = q == 0 && lane == 0;  // this thread stores the (group, half) box of every tile of its group
    const uint32_t row_off = static_cast<uint32_t>(gi * GA_SLOT + ch * BW_KB_BYTES + r * 128);
    uint8_t* stage_ptr = base_ptr + GASmem::stage_off;
    const uint32_t s_addr = tmem_addr(tmem + gi * BW_BN, q * 32, ch * 64);
    uint32_t tg0 = 0, s_par = 0, piece = 0, it_ring = 0;
    // one global scale for every pair: G is shared by both tensors of a pair
    float gs[TCL_MAX_PAIRS], gmax = 0.f;
    for (int p = 0; p < P.n_pairs; ++p) {
      gs[p] = P.pair[p].grad_scale ? *P.pair[p].grad_scale : 1.f;
      gmax = fmaxf(gmax, fabsf(gs[p]));
    }
    const float inv_gmax = gmax > 0.f ? kGScale / gmax : 0.f;  // G in [-kGScale, kGScale]: see ntxent_bwd.h
    if (blockIdx.x == 0 && ew == 0 && lane == 0) *P.scale_out = gmax * P.out_scale * (1.f / kGScale);

    while (walk.next(unit, ta, tb)) {
      const int pi = unit / P.n_iblocks;
      const GPairDev& G = P.pair[pi];
      const int i0 = (unit % P.n_iblocks) * BW_BM;
      const int grow = i0 + r;
      if (piece > 0) asm volatile("bar.sync 7, 512;" ::: "memory");  // the previous piece's logit MMAs are complete
      {
        const int c0 = (gi * 2 + ch) * 2;
        if (c0 < num_kb) {
          const uint32_t itx = it_ring + static_cast<uint32_t>(c0 >> 1);
          const int sx = itx % GA_PSTAGES;
          mbar_wait(p_full(sx), (itx / GA_PSTAGES) & 1);
          const uint8_t* slot = base_ptr + GASmem::ring_off + sx * GA_SLOT;
#pragma unroll 1
          for (int c32 = c0; c32 < c0 + 2; ++c32) {
            if (c32 >= num_kb) break;
            const uint8_t* rowp = slot + (c32 - c0) * BW_KB_BYTES + r * 128;
            uint32_t xv[32];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint4 u = *reinterpret_cast<const uint4*>(rowp + ((e ^ (r & 7)) << 4));
              xv[4 * e] = u.x; xv[4 * e + 1] = u.y; xv[4 * e + 2] = u.z; xv[4 * e + 3] = u.w;
            }
            tmem_st_32x32b_x32(tmem_addr(tmem_x, q * 32, c32 * 32), xv);
          }
          tc_wait_st();
        }
        tc_fence_before();
        asm volatile("bar.sync 8, 512;" ::: "memory");
        if (ew == 0 && lane == 0)
          for (int x = 0; x < x_slots; ++x) mbar_arrive(p_empty((it_ring + x) % GA_PSTAGES));
        if (lane == 0) mbar_arrive(x_full_bar);
        it_ring += static_cast<uint32_t>(x_slots + (tb - ta) * x_slots);
      }
      const float rr = gs[pi] * inv_gmax;
      const float lse_i = grow < P.batch ? G.lse_row[grow] : 0.f;
      const float ws = rr * P.alpha;
      const float wo_i = rr * (1.f - P.alpha) * ex2_approx(lse_i - P.c1);

      int t = ta + static_cast<int>((static_cast<uint32_t>(gi) - tg0) & 1u);
      auto load_lse = [&](int jtile, bool valid) -> float {
        if (gt >= 128 || !valid) return 1e30f;
        const int j = jtile * BW_BN + gt;
        return j < P.batch ? G.lse_col[j] : 1e30f;
      };
      float lse_col = load_lse(t, t < tb);
      for (; t < tb; t += 2) {
        const int j0 = t * BW_BN;
        if (issuer) bulk_wait_read_all();  // the group's previous G box has left the staging slot
        if (gt < 128) bj[gt] = ex2_approx(P.c1 - lse_col);
        lse_col = load_lse(t + 2, t + 2 < tb);
        asm volatile("bar.sync %0, 256;" ::"r"(bar_grp) : "memory");
        const int dcol = grow - j0 - ch * 64;
        const bool has_diag = (i0 < j0 + BW_BN) && (i0 + BW_BM > j0);

        mbar_wait(s_full(gi), s_par);
        s_par ^= 1;
        tc_fence_after();
        uint32_t pk[2][16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(s_addr + h * 32, v);
          tc_wait_ld();
          if (h == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_empty(gi));
          }
          const float4* bj4 = reinterpret_cast<const float4*>(bj + ch * 64 + h * 32);
          if (!has_diag) {
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 bb = bj4[e >> 2];
              const float p0 = ex2_approx(fmaf(__uint_as_float(v[e]), P.c1, -lse_i));
              const float p1 = ex2_approx(fmaf(__uint_as_float(v[e + 1]), P.c1, -lse_i));
              const float p2 = ex2_approx(fmaf(__uint_as_float(v[e + 2]), P.c1, -lse_i));
              const float p3 = ex2_approx(fmaf(__uint_as_float(v[e + 3]), P.c1, -lse_i));
              pk[h][e >> 1] = pack2<kOp>(p0 * fmaf(wo_i, bb.x, ws), p1 * fmaf(wo_i, bb.y, ws));
              pk[h][(e >> 1) + 1] = pack2<kOp>(p2 * fmaf(wo_i, bb.z, ws), p3 * fmaf(wo_i, bb.w, ws));
            }
          } else {
            const int dl = dcol - h * 32;
#pragma unroll
            for (int e = 0; e < 32; e += 4) {
              const float4 bb = bj4[e >> 2];
              const float p0 = ex2_approx(fmaf(__uint_as_float(v[e]), P.c1, -lse_i));
              const float p1 = ex2_approx(fmaf(__uint_as_float(v[e + 1]), P.c1, -lse_i));
              const float p2 = ex2_approx(fmaf(__uint_as_float(v[e + 2]), P.c1, -lse_i));
              const float p3 = ex2_approx(fmaf(__uint_as_float(v[e + 3]), P.c1, -lse_i));
              const float g0 = fmaf(p0, fmaf(wo_i, bb.x, ws), (e == dl) ? -rr : 0.f);
              const float g1 = fmaf(p1, fmaf(wo_i, bb.y, ws), (e + 1 == dl) ? -rr : 0.f);
              const float g2 = fmaf(p2, fmaf(wo_i, bb.z, ws), (e + 2 == dl) ? -rr : 0.f);
              const float g3 = fmaf(p3, fmaf(wo_i, bb.w, ws), (e + 3 == dl) ? -rr : 0.f);
              pk[h][e >> 1] = pack2<kOp>(g0, g1);
              pk[h][(e >> 1) + 1] = pack2<kOp>(g2, g3);
            }
          }
        }
        // swizzled [128 rows][64 cols] 16-bit box: row r, 16-byte chunk c16 -> c16 ^ (r & 7)
        uint8_t* gk = stage_ptr + row_off;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const int c16 = h * 4 + c4;
            *reinterpret_cast<uint4*>(gk + ((c16 ^ (r & 7)) << 4)) =
                make_uint4(pk[h][4 * c4], pk[h][4 * c4 + 1], pk[h][4 * c4 + 2], pk[h][4 * c4 + 3]);
          }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(bar_kb) : "memory");
        if (issuer) {
          tma_store_2d(&G.tm_g, stage + gi * GA_SLOT + ch * BW_KB_BYTES, j0 + ch * 64, i0);
          bulk_commit_group();
        }
        asm volatile("bar.sync %0, 256;" ::"r"(bar_grp) : "memory");  // bj may be rewritten
      }
      tg0 += static_cast<uint32_t>(tb - ta);
      ++piece;
    }
    if (issuer) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// =================================================================================================================
// kernel B: gradient GEMMs over the stored G
// =================================================================================================================
static constexpr int GB_GSLOTS = 2;
static constexpr int GB_CSTAGES = 4;
static constexpr int GB_SLOT = 32768;
static constexpr int GB_DRAIN_WARPS = 8;
static constexpr int GB_DRAIN_BYTES = 4096;
static constexpr int GB_THREADS = 64 + GB_DRAIN_WARPS * 32;
struct GBSmem {
  static constexpr uint32_t g_off = 0;
  static constexpr uint32_t ring_off = GB_GSLOTS * GB_SLOT;
  static constexpr uint32_t drain_off = (GB_GSLOTS + GB_CSTAGES) * GB_SLOT;
  static constexpr uint32_t bar_off = drain_off + GB_DRAIN_WARPS * GB_DRAIN_BYTES;
  static constexpr uint32_t total = bar_off + 256 + 1024;
};
static_assert(GBSmem::total <= 232448, "shared-G backward, kernel B: shared memory budget");

struct GBSegDev {
  CUtensorMap tm_g;      // row side: G [B, ld_g] box {64 k, 128 m};  column side: box {64 m, 64 k} (transposed read)
  CUtensorMap tm_other;  // other operand [B, dim], box {64 dims, 64 rows} (MN-major B)
  int col_side;
};
struct GBJobDev {
  GBSegDev seg[2];
  CUtensorMap tm_gpart;  // [kBwdMaxSplit * n_self_pad, dim] f32, box {32, 32}
  int n_seg;
};
struct GBParams {
  GBJobDev job[TCL_MAX_TENSORS];
  int64_t job_tile_base[TCL_MAX_TENSORS + 1];
  int unit_tiles[TCL_MAX_TENSORS];
  int n_jtiles, n_self_pad, dim;
  uint32_t idesc_row, idesc_col;  // M=128, N=256, B MN-major; A K-major / MN-major
};

struct GBPiece {
  int job, ib, ta, tb, slot;
};
struct GBWalk {
  int64_t cursor, end;
  int c, n;
  __device__ explicit GBWalk(const GBParams& P) {
    n = static_cast<int>(gridDim.x);
    c = static_cast<int>(blockIdx.x);
    const int64_t total = P.job_tile_base[TCL_MAX_TENSORS];
    cursor = pc_range_lo(total, c, n);
    end = pc_range_lo(total, c + 1, n);
  }
  __device__ bool next(const GBParams& P, GBPiece& pc) {
    if (cursor >= end) return false;
    int j = 0;
    while (j + 1 < TCL_MAX_TENSORS && cursor >= P.job_tile_base[j + 1]) ++j;
    const int T = P.unit_tiles[j];
    const int64_t local = cursor - P.job_tile_base[j];
    pc.job = j;
    pc.ib = static_cast<int>(local / T);
    pc.ta = static_cast<int>(local % T);
    const int64_t left = end - cursor;
    pc.tb = left < T - pc.ta ? pc.ta + static_cast<int>(left) : T;
    pc.slot = c - pc_range_of(P.job_tile_base[TCL_MAX_TENSORS], cursor - pc.ta, n);
    cursor += pc.tb - pc.ta;
    return true;
  }
};

__global__ void __launch_bounds__(GB_THREADS, 1) ntxent_ggemm_kernel(const __grid_constant__ GBParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + GBSmem::bar_off;
  auto g_full = [&](int s) { return bars + 8u * s; };
  auto g_empty = [&](int s) { return bars + 8u * (GB_GSLOTS + s); };
  auto c_full = [&](int s) { return bars + 8u * (2 * GB_GSLOTS + s); };
  auto c_empty = [&](int s) { return bars + 8u * (2 * GB_GSLOTS + GB_CSTAGES + s); };
  const uint32_t acc_full_bar = bars + 8u * (2 * GB_GSLOTS + 2 * GB_CSTAGES);
  const uint32_t acc_empty_bar = acc_full_bar + 8u;
  const uint32_t tmem_slot = acc_full_bar + 16u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      base_ptr + GBSmem::bar_off + 8u * (2 * GB_GSLOTS + 2 * GB_CSTAGES) + 16u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunk = (P.dim + 255) / 256;
  const uint32_t g_smem = base + GBSmem::g_off;
  const uint32_t ring = base + GBSmem::ring_off;

  if (warp == 0 && elect_one()) {
    for (int j = 0; j < TCL_MAX_TENSORS; ++j) {
      if (P.unit_tiles[j] == 0) continue;
      tma_prefetch_desc(&P.job[j].tm_gpart);
      for (int s = 0; s < P.job[j].n_seg; ++s) {
        tma_prefetch_desc(&P.job[j].seg[s].tm_g);
        tma_prefetch_desc(&P.job[j].seg[s].tm_other);
      }
    }
    for (int s = 0; s < GB_GSLOTS; ++s) {
      mbar_init(g_full(s), 1);
      mbar_init(g_empty(s), 1);
    }
    for (int s = 0; s < GB_CSTAGES; ++s) {
      mbar_init(c_full(s), 1);
      mbar_init(c_empty(s), 1);
    }
    mbar_init(acc_full_bar, 1);
    mbar_init(acc_empty_bar, GB_DRAIN_WARPS);
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
  GBWalk walk(P);
  GBPiece pc;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA warp: G tiles and operand tiles
    if (elect_one()) {
      uint32_t it = 0, tg = 0;
      while (walk.next(P, pc)) {
        const GBJobDev& J = P.job[pc.job];
        const int i0 = pc.ib * BW_BM;
        for (int t = pc.ta; t < pc.tb; ++t, ++tg) {
          const GBSegDev& sg = J.seg[t / P.n_jtiles];
          const int k0 = (t % P.n_jtiles) * BW_BN;  // first row of the other operand = first K index
          const int gsl = tg % GB_GSLOTS;
          mbar_wait(g_empty(gsl), ((tg / GB_GSLOTS) & 1) ^ 1);
          mbar_arrive_expect_tx(g_full(gsl), GB_SLOT);
          if (!sg.col_side) {  // A[m = self row, k = other row] = G[self row][other row]: K-major, two K-blocks
            for (int h = 0; h < 2; ++h)
              tma_load_2d(g_smem + gsl * GB_SLOT + h * BW_KB_BYTES, &sg.tm_g, g_full(gsl), k0 + h * BW_BK, i0);
          } else {  // A[m, k] = G[other row k][self row m]: MN-major, per K half two groups of 64 self rows
            for (int kh = 0; kh < 2; ++kh)
              for (int mg = 0; mg < 2; ++mg)
                tma_load_2d(g_smem + gsl * GB_SLOT + kh * BW_KB_BYTES + mg * 8192, &sg.tm_g, g_full(gsl), i0 + mg * 64,
                            k0 + kh * BW_BK);
          }
          for (int kb2 = 0; kb2 < 2; ++kb2)
            for (int c = 0; c < n_chunk; ++c, ++it) {
              const int s = it % GB_CSTAGES;
              mbar_wait(c_empty(s), ((it / GB_CSTAGES) & 1) ^ 1);
              mbar_arrive_expect_tx(c_full(s), GB_SLOT);
              for (int a = 0; a < 4; ++a)
                tma_load_2d(ring + s * GB_SLOT + a * 8192, &sg.tm_other, c_full(s), c * 256 + a * 64, k0 + kb2 * BW_BK);
            }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- gradient MMAs
    if (elect_one()) {
      uint32_t it = 0, tg = 0, piece = 0;
      while (walk.next(P, pc)) {
        const GBJobDev& J = P.job[pc.job];
        mbar_wait(acc_empty_bar, (piece & 1) ^ 1);
        tc_fence_after();
        for (int t = pc.ta; t < pc.tb; ++t, ++tg) {
          const bool col_side = J.seg[t / P.n_jtiles].col_side != 0;
          const uint32_t idesc = col_side ? P.idesc_col : P.idesc_row;
          const int gsl = tg % GB_GSLOTS;
          mbar_wait(g_full(gsl), (tg / GB_GSLOTS) & 1);
          tc_fence_after();
          for (int kb2 = 0; kb2 < 2; ++kb2)
            for (int c = 0; c < n_chunk; ++c, ++it) {
              const int s = it % GB_CSTAGES;
              mbar_wait(c_full(s), (it / GB_CSTAGES) & 1);
              tc_fence_after();
              const uint32_t a_addr = g_smem + gsl * GB_SLOT + kb2 * BW_KB_BYTES;
              const uint64_t ad = col_side ? umma_desc_mn_sw128(a_addr, 8192) : umma_desc_k_sw128(a_addr);
              const uint32_t a_step = col_side ? 128u : 2u;  // one UMMA_K: 16 K rows of 128 bytes / 32 bytes along K
              const uint64_t bd = umma_desc_mn_sw128(ring + s * GB_SLOT, 8192);
#pragma unroll
              for (int kk = 0; kk < BW_BK / 16; ++kk)
                tc_mma_f16(tmem + c * 256, ad + a_step * kk, bd + 128 * kk, idesc, ((t - pc.ta) | kb2 | kk) != 0);
              tc_commit(c_empty(s));
            }
          tc_commit(g_empty(gsl));
        }
        tc_commit(acc_full_bar);
        ++piece;
      }
    }
  } else {
    // ---------------------------------------------------------------- accumulator read-out (8 warps), as ntxent_bwd_pc.cu
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t stg = base + GBSmem::drain_off + static_cast<uint32_t>(warp - 2) * GB_DRAIN_BYTES;
    uint8_t* stg_ptr = base_ptr + GBSmem::drain_off + (warp - 2) * GB_DRAIN_BYTES;
    uint32_t piece = 0;
    while (walk.next(P, pc)) {
      mbar_wait(acc_full_bar, piece & 1);
      tc_fence_after();
      const int row0 = pc.slot * P.n_self_pad + pc.ib * BW_BM + q * 32;
#pragma unroll 1
      for (int cc = 0; cc < 8; ++cc) {
        const int col = half * 256 + cc * 32;
        if (col >= P.dim) break;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_addr(tmem, q * 32, col), v);
        tc_wait_ld();
        if (lane == 0) bulk_wait_read_all();
        __syncwarp();
        uint8_t* rowp = stg_ptr + lane * 128;
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16)
          *reinterpret_cast<uint4*>(rowp + ((c16 ^ (lane & 7)) << 4)) =
              make_uint4(v[4 * c16], v[4 * c16 + 1], v[4 * c16 + 2], v[4 * c16 + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&P.job[pc.job].tm_gpart, stg, col, row0);
          bulk_commit_group();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty_bar);
      ++piece;
    }
    if (lane == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// -----------------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------------
size_t bwd_sharedg_workspace_bytes(int n_pairs, int64_t batch) {
  const int64_t ld_g = (batch + 63) / 64 * 64;
  return static_cast<size_t>(n_pairs) * batch * ld_g * 2 + 1024;
}

// The shared-G form applies to the single-GPU whole-loss entry (ntxent_fused.cu) for dim > 256.  TRICOLO_B200_BWD=sharedg
// forces it, any other explicit value keeps the per-direction kernels; by default it is used from 2048 rows on
// (below, the producer/consumer kernel's two launches fewer win) while G fits 8 GB.
bool bwd_sharedg_enabled(int n_pairs, int64_t batch, int64_t dim) {
  if (dim <= BW_DH) return false;
  const char* e = getenv("TRICOLO_B200_BWD");
  if (e && *e) return !strcmp(e, "sharedg");
  return batch >= 2048 && bwd_sharedg_workspace_bytes(n_pairs, batch) <= (8ull << 30);
}

template <int kOp>
static int launch_g_kernel(const GAParams& A, int n_ctas, cudaStream_t st) {
  const int smem = static_cast<int>(GASmem::total);
  if (int e = ensure_dyn_smem(ntxent_g_kernel<kOp>, smem)) return e;
  ntxent_g_kernel<kOp><<<n_ctas, GA_THREADS, smem, st>>>(A);
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

int launch_bwd_sharedg(const BwdSharedGArgs& a, cudaStream_t st) {
  const int64_t batch = a.batch, dim = a.dim;
  const int64_t ld_g = (batch + 63) / 64 * 64;
  const int n_iblocks = static_cast<int>((batch + BW_BM - 1) / BW_BM);
  const int n_jtiles = static_cast<int>((batch + BW_BN - 1) / BW_BN);
  const int n_self_pad = n_iblocks * BW_BM;
  int n_sm = 0, dev = 0;
  TCL_CHECK_CUDA(cudaGetDevice(&dev));
  TCL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  char* ws = static_cast<char*>(a.workspace);
  float* scale = reinterpret_cast<float*>(ws);          // 64 floats reserved
  float* gbase = reinterpret_cast<float*>(ws) + 64;     // gradient partial slots, as tcl_ntxent_bwd
  uint16_t* g_mat = reinterpret_cast<uint16_t*>(ws + a.partials_bytes);

  // ---- kernel A: one G per pair that has at least one tensor needing a gradient
  GAParams A;
  memset(&A, 0, sizeof(A));
  int pair_slot[TCL_MAX_PAIRS];
  for (int p = 0; p < a.n_pairs; ++p) {
    pair_slot[p] = -1;
    if (!a.need_grad[a.pair_row[p]] && !a.need_grad[a.pair_col[p]]) continue;
    const int k = A.n_pairs++;
    pair_slot[p] = k;
    GPairDev& G = A.pair[k];
    if (int e = make_tmap_2d_16bit(&G.tm_row, a.z[a.pair_row[p]], batch, dim, dim, BW_BM, BW_BK)) return e;
    if (int e = make_tmap_2d_16bit(&G.tm_col, a.z[a.pair_col[p]], batch, dim, dim, BW_BN, BW_BK)) return e;
    if (int e = make_tmap_2d_16bit(&G.tm_g, g_mat + static_cast<size_t>(k) * batch * ld_g, batch, batch, ld_g, BW_BM, 64)) return e;
    G.lse_row = a.lse_row + static_cast<size_t>(p) * batch;
    G.lse_col = a.lse_col + static_cast<size_t>(p) * batch;
    G.grad_scale = a.grad_losses + p;
  }
  if (A.n_pairs == 0) return TCL_OK;
  A.scale_out = scale;
  A.batch = static_cast<int>(batch);
  A.num_kb = static_cast<int>(dim / 64);
  A.n_jtiles = n_jtiles;
  A.n_iblocks = n_iblocks;
  A.c1 = a.inv_tau * 1.4426950408889634f;
  A.alpha = a.alpha;
  A.out_scale = a.inv_tau / static_cast<float>(batch);
  A.idesc = umma_idesc_f16(BW_BM, BW_BN, a.op_format);
  const int64_t total_a = static_cast<int64_t>(A.n_pairs) * n_iblocks * n_jtiles;
  const int ctas_a = static_cast<int>(total_a < n_sm ? total_a : n_sm);
  prof_begin(TCL_K_NTXENT_G, st);
  if (int e = (a.op_format == TCL_OP_F16 ? launch_g_kernel<TCL_OP_F16>(A, ctas_a, st) : launch_g_kernel<TCL_OP_BF16>(A, ctas_a, st)))
    return e;
  prof_end(TCL_K_NTXENT_G, st);

  // ---- kernel B: per tensor, acc += A * Zother over its pairs
  GBParams B;
  memset(&B, 0, sizeof(B));
  NormBwdParams N;
  memset(&N, 0, sizeof(N));
  int n_jobs = 0;
  for (int m = 0; m < a.n_tensors; ++m) {
    if (!a.need_grad[m]) continue;
    GBJobDev& J = B.job[n_jobs];
    for (int p = 0; p < a.n_pairs; ++p) {
      if (a.pair_row[p] != m && a.pair_col[p] != m) continue;
      TCL_REQUIRE(J.n_seg < 2, TCL_ERR_BAD_ARG, "loss_bwd: tensor %d takes part in more than two pairs", m);
      GBSegDev& S = J.seg[J.n_seg++];
      S.col_side = a.pair_col[p] == m;
      const int o = S.col_side ? a.pair_row[p] : a.pair_col[p];
      uint16_t* gp = g_mat + static_cast<size_t>(pair_slot[p]) * batch * ld_g;
      if (int e = make_tmap_2d_16bit(&S.tm_g, gp, batch, batch, ld_g, S.col_side ? 64 : BW_BM, 64)) return e;
      if (int e = make_tmap_2d_16bit(&S.tm_other, a.z[o], batch, dim, dim, 64, 64)) return e;
    }
    if (J.n_seg == 0) continue;
    float* gpart = gbase + static_cast<size_t>(n_jobs) * kBwdMaxSplit * n_self_pad * dim;
    if (int e = make_tmap_2d_f32(&J.tm_gpart, gpart, static_cast<uint64_t>(kBwdMaxSplit) * n_self_pad, dim, 32, 32)) return e;
    B.unit_tiles[n_jobs] = J.n_seg * n_jtiles;
    B.job_tile_base[n_jobs + 1] = B.job_tile_base[n_jobs] + static_cast<int64_t>(n_iblocks) * B.unit_tiles[n_jobs];
    N.job[n_jobs].x = a.x[m];
    N.job[n_jobs].inv_norm = a.inv_norm + static_cast<size_t>(m) * batch;
    N.job[n_jobs].gpart = gpart;
    N.job[n_jobs].scale = scale;
    N.job[n_jobs].dx = a.dx[m];
    ++n_jobs;
  }
  for (int j = n_jobs; j < TCL_MAX_TENSORS; ++j) B.job_tile_base[j + 1] = B.job_tile_base[n_jobs];
  B.n_jtiles = n_jtiles;
  B.n_self_pad = n_self_pad;
  B.dim = static_cast<int>(dim);
  B.idesc_row = umma_idesc_f16(BW_BM, 256, a.op_format) | (1u << 16);
  B.idesc_col = B.idesc_row | (1u << 15);
  const int64_t total_b = B.job_tile_base[TCL_MAX_TENSORS];
  int t_max = 1;
  for (int j = 0; j < n_jobs; ++j) t_max = B.unit_tiles[j] > t_max ? B.unit_tiles[j] : t_max;
  int64_t ctas_b = n_sm;
  if (ctas_b > total_b) ctas_b = total_b;
  const int64_t cap = (kBwdMaxSplit - 1) * total_b / t_max;
  if (ctas_b > cap) ctas_b = cap;
  if (ctas_b < 1) ctas_b = 1;
  {
    const int smem = static_cast<int>(GBSmem::total);
    if (int e = ensure_dyn_smem(ntxent_ggemm_kernel, smem)) return e;
    prof_begin(TCL_K_NTXENT_BWD, st);
    ntxent_ggemm_kernel<<<static_cast<unsigned>(ctas_b), GB_THREADS, smem, st>>>(B);
    prof_end(TCL_K_NTXENT_BWD, st);
    TCL_CHECK_CUDA(cudaGetLastError());
  }
  N.n_clusters = static_cast<int>(ctas_b);
  N.split_rows = n_self_pad;
  N.total_tiles = total_b;
  for (int j = 0; j < TCL_MAX_TENSORS; ++j) {
    N.job_tile_base[j] = B.job_tile_base[j];
    N.unit_tiles[j] = B.unit_tiles[j];
  }
  return launch_l2norm_bwd(N, n_jobs, a.x_dtype, batch, static_cast<int>(dim), a.x_row_stride, kBwdMaxSplit, a.eps, st);
}

}  // namespace tcl
