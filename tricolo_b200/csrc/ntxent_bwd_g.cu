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
// Sharded over ranks (tricolo_b200/distributed.py): a rank owns the row block G[b_loc x B] of every pair.
// Its row-side gradients dRow = G Zcol are complete locally; the column-side products dCol = G^T Zrow_local are partial
// sums for ALL B rows of the column tensor, and kernel B's drain TMA-stores each 128-row piece straight into the owner
// rank's receive buffer over NVLink peer memory (slot = source rank): the gradient reduce-scatter is the GEMM kernel's
// own store traffic, 6 b B D executed per pair instead of the 8 of ntxent_bwd_pc.cu.  The owner's normalise backward
// adds the W x pieces partials in a fixed order (bit-reproducible, no atomics).
#include <stdlib.h>

#include "ntxent_bwd.h"
#include "norm_fold.cuh"
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
  GPairDev pair[TCL_MAX_PAIRS];  // lse_row is indexed by the LOCAL row (the host passes the rank's slice)
  // device scalars max|grad_scale| / (tau B), consumed by the normalise backward: one local copy, or (sharded) this
  // rank's entry of every rank's receive-buffer header
  float* scale_out[TCL_MAX_PEERS];
  int n_scale_out;
  uint32_t* zero_words;  // block counters of the folded normalise backward (norm_fold.cuh), cleared here for kernel B
  int n_zero_words;
  int n_pairs, n_rows, n_cols, row_offset, num_kb, n_jtiles, n_iblocks;
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
  griddep_launch();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tmem_x = tmem + GA_XCOL;
  griddep_wait();
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
    if (blockIdx.x == 0 && ew == 0 && lane < P.n_scale_out) *P.scale_out[lane] = gmax * P.out_scale * (1.f / kGScale);
    if (blockIdx.x == 0 && ew == 1)
      for (int i = lane; i < P.n_zero_words; i += 32) P.zero_words[i] = 0u;

    while (walk.next(unit, ta, tb)) {
      const int pi = unit / P.n_iblocks;
      const GPairDev& G = P.pair[pi];
      const int i0 = (unit % P.n_iblocks) * BW_BM;
      const int lrow = i0 + r;                  // local row (row of G)
      const int grow = P.row_offset + lrow;     // its index in the global batch (column of the positive)
      const int i0g = P.row_offset + i0;
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
      const float lse_i = lrow < P.n_rows ? G.lse_row[lrow] : 0.f;
      const float ws = rr * P.alpha;
      const float wo_i = rr * (1.f - P.alpha) * ex2_approx(lse_i - P.c1);

      int t = ta + static_cast<int>((static_cast<uint32_t>(gi) - tg0) & 1u);
      auto load_lse = [&](int jtile, bool valid) -> float {
        if (gt >= 128 || !valid) return 1e30f;
        const int j = jtile * BW_BN + gt;
        return j < P.n_cols ? G.lse_col[j] : 1e30f;
      };
      float lse_col = load_lse(t, t < tb);
      for (; t < tb; t += 2) {
        const int j0 = t * BW_BN;
        if (issuer) bulk_wait_read_all();  // the group's previous G box has left the staging slot
        if (gt < 128) bj[gt] = ex2_approx(P.c1 - lse_col);
        lse_col = load_lse(t + 2, t + 2 < tb);
        asm volatile("bar.sync %0, 256;" ::"r"(bar_grp) : "memory");
        const int dcol = grow - j0 - ch * 64;
        const bool has_diag = (i0g < j0 + BW_BN) && (i0g + BW_BM > j0);

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
// A job = one output matrix: the gradient of one tensor from the pairs in which it is the row side (A = G tiles
// K-major, K runs over the B columns of G, output rows = the rank's local rows, local partial slots) or from the pairs
// in which it is the column side (A = the transposed read of G, K runs over the rank's G rows, output rows = all B
// rows of the tensor; sharded: stored into the owner rank's receive buffer).  On one GPU both kinds of a tensor share
// one accumulator (one job with a row-side and a column-side segment).
// A unit = (job, 128-row block[, 256-column half of dim]) = `unit_tiles` consecutive tiles of one accumulator.  With
// n_dsplit = 2 the accumulator is one half of TMEM and consecutive pieces alternate halves, so the read-out of a piece
// overlaps the MMAs of the next (short units: sharded runs); with n_dsplit = 1 it is all of TMEM (dim columns).
static constexpr int GB_GSLOTS = 2;
static constexpr int GB_CSTAGES = 4;
static constexpr int GB_SLOT = 32768;
static constexpr int GB_DRAIN_WARPS = 8;
static constexpr int GB_DRAIN_BYTES = 4096;
static constexpr int GB_THREADS = 64 + GB_DRAIN_WARPS * 32;
static constexpr int GB_MAX_JOBS = 2 * TCL_MAX_TENSORS;
static constexpr int GB_MAX_DST = TCL_MAX_TENSORS + TCL_MAX_TENSORS * TCL_MAX_PEERS;
struct GBSmem {
  static constexpr uint32_t g_off = 0;
  static constexpr uint32_t ring_off = GB_GSLOTS * GB_SLOT;
  static constexpr uint32_t drain_off = (GB_GSLOTS + GB_CSTAGES) * GB_SLOT;
  static constexpr uint32_t bar_off = drain_off + GB_DRAIN_WARPS * GB_DRAIN_BYTES;
  static constexpr uint32_t total = bar_off + 256 + 1024;
};
static_assert(GBSmem::total <= 232448, "shared-G backward, kernel B: shared memory budget");

struct GBSegDev {
  // G [n_rows, ld_g] as (64, n_rows, ld_g / 64), box {64, 128, 2}.  Row side: two K-blocks {64 k, 128 m} (K-major A);
  // column side (transposed read): two 64-wide m groups of 128 k rows each (MN-major A), one instruction either way
  CUtensorMap tm_g;
  CUtensorMap tm_other;  // other operand [k rows, dim] as (64, rows, dim / 64), box {64, 64, 4}: MN-major B, 256 columns
  int col_side;
};
struct GBJobDev {
  GBSegDev seg[2];
  int n_seg;
  int n_rowblocks;   // 128-row blocks of the output (the CTA-pair kernel pads an odd count)
  int k_tiles;       // 128-wide K tiles per segment
  int dst_first;     // first tensor map of this job in GBParams::tm_dst
  int owner_rows;    // > 0: output rows are owned in blocks of owner_rows by successive ranks (tm_dst[dst_first + owner])
  int slot_rows;     // rows of one partial slot in the destination
  int src_row_base;  // first destination row of this rank's slots (remote: rank * n_slots * slot_rows)
  int dst16;         // the destination holds fp16 (partials that cross NVLink): tm_dst boxes are {64 cols, 32 rows}
};
struct GBParams {
  GBJobDev job[GB_MAX_JOBS];
  CUtensorMap tm_dst[GB_MAX_DST];  // f32 [slots * rows, dim], box {32, 32}
  int64_t job_tile_base[GB_MAX_JOBS + 1];
  int unit_tiles[GB_MAX_JOBS];  // n_seg * k_tiles
  int dim, n_dsplit;
  uint32_t idesc_row, idesc_col;  // M=128, N=256, B MN-major; A K-major / MN-major
  // sharded: when the LAST CTA has finished (all partials stored, also those that crossed NVLink) it signals
  // kGrads[rank] = epoch to every rank and publishes the backward epoch (host_common.h: ShardSync); world = 0: off
  uint32_t* sync[TCL_MAX_PEERS];
  int rank, world;
  FoldParams fold;  // normalise backward by the read-out warps that complete a row block (norm_fold.cuh); jobs indexed alike
};

struct GBPiece {
  int job, ib, dh, ta, tb, slot, n_pieces;  // n_pieces: pieces the whole unit is cut into
};

// wait-time accounting of the first and the last CTA (`make trace`): [0..31] CTA 0, [32..63] the last CTA
__device__ unsigned long long g_gb_trace[64];
#ifdef TCL_PAIR_TRACE
#define GT_DECL unsigned long long gt_t0 = 0; const bool gt_on = (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && (threadIdx.x & 31) == 0; const int gt_base = blockIdx.x == 0 ? 0 : 32; (void)gt_t0; (void)gt_base;
#define GT_BEGIN() do { if (gt_on) gt_t0 = clock64(); } while (0)
#define GT_END(slot) do { if (gt_on) { const unsigned long long gt_t1 = clock64(); atomicAdd(&g_gb_trace[gt_base + (slot)], gt_t1 - gt_t0); gt_t0 = gt_t1; } } while (0)
#define GT_ADD(slot, v) do { if (gt_on) atomicAdd(&g_gb_trace[gt_base + (slot)], (unsigned long long)(v)); } while (0)
#else
#define GT_DECL
#define GT_BEGIN() do {} while (0)
#define GT_END(slot) do {} while (0)
#define GT_ADD(slot, v) do {} while (0)
#endif
struct GBWalk {
  int64_t cursor, end;
  int c, n;
  __device__ explicit GBWalk(const GBParams& P) {
    n = static_cast<int>(gridDim.x);
    c = static_cast<int>(blockIdx.x);
    const int64_t total = P.job_tile_base[GB_MAX_JOBS];
    cursor = pc_range_lo(total, c, n);
    end = pc_range_lo(total, c + 1, n);
  }
  __device__ bool next(const GBParams& P, GBPiece& pc) {
    if (cursor >= end) return false;
    int j = 0;
    while (j + 1 < GB_MAX_JOBS && cursor >= P.job_tile_base[j + 1]) ++j;
    const int T = P.unit_tiles[j];
    const int64_t local = cursor - P.job_tile_base[j];
    const int u = static_cast<int>(local / T);
    pc.job = j;
    pc.ib = P.n_dsplit == 2 ? u >> 1 : u;
    pc.dh = P.n_dsplit == 2 ? u & 1 : 0;
    pc.ta = static_cast<int>(local % T);
    const int64_t left = end - cursor;
    pc.tb = left < T - pc.ta ? pc.ta + static_cast<int>(left) : T;
    {
      const int64_t total = P.job_tile_base[GB_MAX_JOBS], first = cursor - pc.ta;
      const int r0 = pc_range_of(total, first, n);
      pc.slot = c - r0;
      pc.n_pieces = pc_range_of(total, first + T - 1, n) - r0 + 1;
    }
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
  auto acc_full = [&](int b) { return bars + 8u * (2 * GB_GSLOTS + 2 * GB_CSTAGES + b); };
  auto acc_empty = [&](int b) { return bars + 8u * (2 * GB_GSLOTS + 2 * GB_CSTAGES + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * GB_GSLOTS + 2 * GB_CSTAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      base_ptr + GBSmem::bar_off + 8u * (2 * GB_GSLOTS + 2 * GB_CSTAGES + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool split = P.n_dsplit == 2;
  const int n_chunk = split ? 1 : (P.dim + 255) / 256;  // 256-column MMAs per K step of one unit
  const int n_buf = split ? 2 : 1;
  const uint32_t g_smem = base + GBSmem::g_off;
  const uint32_t ring = base + GBSmem::ring_off;

  if (warp == 0 && elect_one()) {
    for (int j = 0; j < GB_MAX_JOBS; ++j) {
      if (P.unit_tiles[j] == 0) continue;
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
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), GB_DRAIN_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  griddep_launch();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  griddep_wait();
  GBWalk walk(P);
  GBPiece pc;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA warp: G tiles and operand tiles
    if (elect_one()) {
      uint32_t it = 0, tg = 0;
      GT_DECL
#ifdef TCL_PAIR_TRACE
      const unsigned long long gt_w0 = clock64();
#endif
      while (walk.next(P, pc)) {
        const GBJobDev& J = P.job[pc.job];
        const int i0 = pc.ib * BW_BM;
        for (int t = pc.ta; t < pc.tb; ++t, ++tg) {
          const GBSegDev& sg = J.seg[t / J.k_tiles];
          const int k0 = (t % J.k_tiles) * BW_BN;  // first row of the other operand = first K index
          const int gsl = tg % GB_GSLOTS;
          GT_BEGIN();
          mbar_wait(g_empty(gsl), ((tg / GB_GSLOTS) & 1) ^ 1);
          GT_END(0);
          mbar_arrive_expect_tx(g_full(gsl), GB_SLOT);
          // row side: A[m = self row, k = other row] = G[self row][other row], K-major: [k block][128 m][64 k];
          // column side: A[m, k] = G[other row k][self row m], MN-major: [m group][128 k][64 m]
          if (!sg.col_side) tma_load_3d(g_smem + gsl * GB_SLOT, &sg.tm_g, g_full(gsl), 0, i0, k0 >> 6);
          else tma_load_3d(g_smem + gsl * GB_SLOT, &sg.tm_g, g_full(gsl), 0, k0, i0 >> 6);
          for (int kb2 = 0; kb2 < 2; ++kb2)
            for (int c = 0; c < n_chunk; ++c, ++it) {
              const int s = it % GB_CSTAGES;
              const int col0 = (split ? pc.dh : c) * 256;
              GT_BEGIN();
              mbar_wait(c_empty(s), ((it / GB_CSTAGES) & 1) ^ 1);
              GT_END(1);
              mbar_arrive_expect_tx(c_full(s), GB_SLOT);
              tma_load_3d(ring + s * GB_SLOT, &sg.tm_other, c_full(s), 0, k0 + kb2 * BW_BK, col0 >> 6);
            }
        }
      }
#ifdef TCL_PAIR_TRACE
      GT_ADD(2, clock64() - gt_w0);
#endif
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- gradient MMAs
    if (elect_one()) {
      uint32_t it = 0, tg = 0, piece = 0;
      GT_DECL
#ifdef TCL_PAIR_TRACE
      const unsigned long long gt_w0 = clock64();
#endif
      while (walk.next(P, pc)) {
        const GBJobDev& J = P.job[pc.job];
        const int buf = static_cast<int>(piece % n_buf);
        GT_BEGIN();
        mbar_wait(acc_empty(buf), ((piece / n_buf) & 1) ^ 1);
        GT_END(3);
        tc_fence_after();
        for (int t = pc.ta; t < pc.tb; ++t, ++tg) {
          const bool col_side = J.seg[t / J.k_tiles].col_side != 0;
          const uint32_t idesc = col_side ? P.idesc_col : P.idesc_row;
          const int gsl = tg % GB_GSLOTS;
          GT_BEGIN();
          mbar_wait(g_full(gsl), (tg / GB_GSLOTS) & 1);
          GT_END(4);
          tc_fence_after();
          for (int kb2 = 0; kb2 < 2; ++kb2)
            for (int c = 0; c < n_chunk; ++c, ++it) {
              const int s = it % GB_CSTAGES;
              GT_BEGIN();
              mbar_wait(c_full(s), (it / GB_CSTAGES) & 1);
              GT_END(5);
              tc_fence_after();
              // K half kb2: the second K-block (row side) / the second 64 k rows of both m groups (column side)
              const uint32_t a_addr = g_smem + gsl * GB_SLOT + kb2 * (col_side ? 8192 : BW_KB_BYTES);
              const uint64_t ad = col_side ? umma_desc_mn_sw128(a_addr, BW_KB_BYTES) : umma_desc_k_sw128(a_addr);
              const uint32_t a_step = col_side ? 128u : 2u;  // one UMMA_K: 16 K rows of 128 bytes / 32 bytes along K
              const uint64_t bd = umma_desc_mn_sw128(ring + s * GB_SLOT, 8192);
#pragma unroll
              for (int kk = 0; kk < BW_BK / 16; ++kk)
                tc_mma_f16(tmem + (buf + c) * 256, ad + a_step * kk, bd + 128 * kk, idesc, ((t - pc.ta) | kb2 | kk) != 0);
              tc_commit(c_empty(s));
            }
          tc_commit(g_empty(gsl));
        }
        tc_commit(acc_full(buf));
        ++piece;
      }
#ifdef TCL_PAIR_TRACE
      GT_ADD(6, clock64() - gt_w0);
      GT_ADD(10, tg);
      GT_ADD(11, piece);
#endif
    }
  } else {
    // ---------------------------------------------------------------- accumulator read-out (8 warps), as ntxent_bwd_pc.cu
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t stg = base + GBSmem::drain_off + static_cast<uint32_t>(warp - 2) * GB_DRAIN_BYTES;
    uint8_t* stg_ptr = base_ptr + GBSmem::drain_off + (warp - 2) * GB_DRAIN_BYTES;
    const int ucols = split ? 256 : P.dim;  // columns of one unit
    const int hw = split ? 128 : 256;       // columns read out by one group of four warps
    uint32_t piece = 0;
#ifdef TCL_PAIR_TRACE
    unsigned long long gt_t0 = 0;
    const bool gt_on = (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && threadIdx.x == 64;
    const int gt_base = blockIdx.x == 0 ? 0 : 32;
    const unsigned long long gt_w0 = clock64();
#endif
    while (walk.next(P, pc)) {
      const GBJobDev& J = P.job[pc.job];
      const int buf = static_cast<int>(piece % n_buf);
      GT_BEGIN();
      mbar_wait(acc_full(buf), (piece / n_buf) & 1);
      GT_END(7);
      tc_fence_after();
      int owner = 0, rin = pc.ib * BW_BM;
      if (J.owner_rows > 0) {
        owner = rin / J.owner_rows;
        rin -= owner * J.owner_rows;
      }
      const CUtensorMap* tm = &P.tm_dst[J.dst_first + owner];
      const int row0 = J.src_row_base + pc.slot * J.slot_rows + rin + q * 32;
      if (J.dst16) {
        // partials that cross NVLink travel as fp16 (half the bytes; |acc| <= ~kGScale fits, 11 significant bits):
        // 64 columns = one 128-byte row of the staging tile per step
#pragma unroll 1
        for (int cc = 0; cc < hw / 64; ++cc) {
          const int ucol = half * hw + cc * 64;
          if (ucol >= ucols) break;
          uint32_t v[32], w[32];
          tmem_ld_32x32b_x32(tmem_addr(tmem, q * 32, buf * 256 + ucol), v);
          tmem_ld_32x32b_x32(tmem_addr(tmem, q * 32, buf * 256 + ucol + 32), w);
          tc_wait_ld();
          if (lane == 0) bulk_wait_read_all();
          __syncwarp();
          uint8_t* rowp = stg_ptr + lane * 128;
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16) {
            const uint32_t* src = c16 < 4 ? v + 8 * c16 : w + 8 * (c16 - 4);
            *reinterpret_cast<uint4*>(rowp + ((c16 ^ (lane & 7)) << 4)) =
                make_uint4(pack2<TCL_OP_F16>(__uint_as_float(src[0]), __uint_as_float(src[1])),
                           pack2<TCL_OP_F16>(__uint_as_float(src[2]), __uint_as_float(src[3])),
                           pack2<TCL_OP_F16>(__uint_as_float(src[4]), __uint_as_float(src[5])),
                           pack2<TCL_OP_F16>(__uint_as_float(src[6]), __uint_as_float(src[7])));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(tm, stg, pc.dh * 256 + ucol, row0);
            bulk_commit_group();
          }
        }
      } else {
#pragma unroll 1
      for (int cc = 0; cc < hw / 32; ++cc) {
        const int ucol = half * hw + cc * 32;
        if (ucol >= ucols) break;
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_addr(tmem, q * 32, buf * 256 + ucol), v);
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
          tma_store_2d(tm, stg, pc.dh * 256 + ucol, row0);
          bulk_commit_group();
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
      if (P.fold.enabled && J.owner_rows == 0)
        fold_piece_done(P.fold, pc.job, pc.ib, pc.n_pieces, warp - 2, lane,
                        reinterpret_cast<volatile uint32_t*>(base_ptr + GBSmem::bar_off + 240), 1);
      GT_END(8);
      ++piece;
    }
    if (lane == 0) {
      bulk_wait_all();               // every TMA store of this warp is complete ...
      fence_proxy_async_generic();   // ... and ordered before the generic-proxy signalling below
      __threadfence_system();
    }
#ifdef TCL_PAIR_TRACE
    GT_ADD(9, clock64() - gt_w0);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
  if (P.world > 0 && threadIdx.x == 0) {
    uint32_t* sy = P.sync[P.rank];
    const uint32_t e = ld_relaxed_u32(sy + ShardSync::kBwdEpoch) + 1u;  // written only by the last CTA, below
    __threadfence_system();
    if (atomicAdd(sy + ShardSync::kGemmDone, 1u) + 1u == gridDim.x) {
      sy[ShardSync::kGemmDone] = 0u;
      __threadfence_system();
      for (int p = 0; p < P.world; ++p) st_release_sys_u32(P.sync[p] + ShardSync::kGrads + P.rank, e);
      *reinterpret_cast<volatile uint32_t*>(sy + ShardSync::kBwdEpoch) = e;
    }
  }
}

// =================================================================================================================
// kernel B, CTA-pair form (cta_group::2): two adjacent 128-row blocks of a job as ONE M = 256 MMA over two SMs
// =================================================================================================================
// The one-SM kernel above ingests 160 KB of operands per 16.8 MFLOP tile (A 32 KB + B 128 KB): 105 flop per byte, and
// the L2 -> SM fabric delivers ~12 TB/s over the whole chip, i.e. 1.29 PFLOP/s - exactly what it measures.  As a CTA
// pair the two row blocks share the B operand: tcgen05.mma.cta_group::2 (M = 256, N = 256) takes each CTA's own 128 A
// rows from that CTA's shared memory and HALF of the N columns of B from each: CTA r loads dims [256 c + 128 r, +128)
// of the other operand, 64 KB instead of 128 KB per tile - 175 flop per byte, above the tensor roof.
//   * only the leader (cluster rank 0) issues MMAs; its `full` barriers count the TMA bytes of BOTH CTAs (the peer's
//     loads name the leader's barrier: cp.async.bulk.tensor ... cta_group::2); tcgen05.commit multicasts to both CTAs'
//     `empty` / `acc_full` barriers; the read-out warps of both CTAs arrive on the leader's `acc_empty`;
//   * a unit = (job, pair of row blocks), T tiles; the flat tile sequence is cut into one range per CLUSTER; each CTA
//     drains its own 128 rows of the accumulator as above.
static constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // same shared-memory offset in the even (leader) CTA of the pair
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t smem_result_addr, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_result_addr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}
// TMA load into THIS CTA's shared memory whose completion bytes are counted on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_3d_2sm(uint32_t smem_dst, const CUtensorMap* m, uint32_t leader_bar, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

struct GB2Walk {  // as GBWalk with one range per cluster; pc.ib = first row block of the pair
  int64_t cursor, end;
  int c, n;
  __device__ explicit GB2Walk(const GBParams& P) {
    n = static_cast<int>(gridDim.x >> 1);
    c = static_cast<int>(blockIdx.x >> 1);
    const int64_t total = P.job_tile_base[GB_MAX_JOBS];
    cursor = pc_range_lo(total, c, n);
    end = pc_range_lo(total, c + 1, n);
  }
  __device__ bool next(const GBParams& P, GBPiece& pc) {
    if (cursor >= end) return false;
    int j = 0;
    while (j + 1 < GB_MAX_JOBS && cursor >= P.job_tile_base[j + 1]) ++j;
    const int T = P.unit_tiles[j];
    const int64_t local = cursor - P.job_tile_base[j];
    pc.job = j;
    pc.ib = 2 * static_cast<int>(local / T);
    pc.dh = 0;
    pc.ta = static_cast<int>(local % T);
    const int64_t left = end - cursor;
    pc.tb = left < T - pc.ta ? pc.ta + static_cast<int>(left) : T;
    {
      const int64_t total = P.job_tile_base[GB_MAX_JOBS], first = cursor - pc.ta;
      const int r0 = pc_range_of(total, first, n);
      pc.slot = c - r0;
      pc.n_pieces = pc_range_of(total, first + T - 1, n) - r0 + 1;
    }
    cursor += pc.tb - pc.ta;
    return true;
  }
};

__global__ void __launch_bounds__(GB_THREADS, 1) ntxent_ggemm2_kernel(const __grid_constant__ GBParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + GBSmem::bar_off;
  auto g_full = [&](int s) { return bars + 8u * s; };                                 // leader's
  auto g_empty = [&](int s) { return bars + 8u * (GB_GSLOTS + s); };                  // per CTA (multicast commit)
  auto c_full = [&](int s) { return bars + 8u * (2 * GB_GSLOTS + s); };               // leader's
  auto c_empty = [&](int s) { return bars + 8u * (2 * GB_GSLOTS + GB_CSTAGES + s); };  // per CTA
  const uint32_t acc_full_bar = bars + 8u * (2 * GB_GSLOTS + 2 * GB_CSTAGES);         // per CTA (multicast commit)
  const uint32_t acc_empty_bar = acc_full_bar + 8u;                                   // leader's: read-out warps of both CTAs
  const uint32_t tmem_slot = acc_full_bar + 32u;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      base_ptr + GBSmem::bar_off + 8u * (2 * GB_GSLOTS + 2 * GB_CSTAGES) + 32u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int n_chunk = (P.dim + 255) / 256;
  const uint32_t g_smem = base + GBSmem::g_off;
  const uint32_t ring = base + GBSmem::ring_off;

  if (warp == 0 && elect_one()) {
    for (int j = 0; j < GB_MAX_JOBS; ++j) {
      if (P.unit_tiles[j] == 0) continue;
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
    mbar_init(acc_empty_bar, 2 * GB_DRAIN_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  griddep_launch();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before any TMA byte or commit may reach them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  griddep_wait();
  GB2Walk walk(P);
  GBPiece pc;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA warp (both CTAs): own G tile, own half of B
    if (elect_one()) {
      uint32_t it = 0, tg = 0;
      GT_DECL
#ifdef TCL_PAIR_TRACE
      const unsigned long long gt_w0 = clock64();
#endif
      while (walk.next(P, pc)) {
        const GBJobDev& J = P.job[pc.job];
        const int i0 = (pc.ib + static_cast<int>(crank)) * BW_BM;  // this CTA's row block
        for (int t = pc.ta; t < pc.tb; ++t, ++tg) {
          const GBSegDev& sg = J.seg[t / J.k_tiles];
          const int k0 = (t % J.k_tiles) * BW_BN;
          const int gsl = tg % GB_GSLOTS;
          GT_BEGIN();
          mbar_wait(g_empty(gsl), ((tg / GB_GSLOTS) & 1) ^ 1);
          GT_END(0);
          if (leader) mbar_arrive_expect_tx(g_full(gsl), 2 * GB_SLOT);  // both CTAs' G tiles
          if (!sg.col_side) tma_load_3d_2sm(g_smem + gsl * GB_SLOT, &sg.tm_g, g_full(gsl) & kPeerBitMask, 0, i0, k0 >> 6);
          else tma_load_3d_2sm(g_smem + gsl * GB_SLOT, &sg.tm_g, g_full(gsl) & kPeerBitMask, 0, k0, i0 >> 6);
          for (int c = 0; c < n_chunk; ++c, ++it) {
            const int s = it % GB_CSTAGES;
            GT_BEGIN();
            mbar_wait(c_empty(s), ((it / GB_CSTAGES) & 1) ^ 1);
            GT_END(1);
            if (leader) mbar_arrive_expect_tx(c_full(s), 2 * GB_SLOT);
            // [2 groups of 64 dims][128 k rows][128 B]: dims [256 c + 128 rank, +128) of the tile's 128 other rows
            tma_load_3d_2sm(ring + s * GB_SLOT, &sg.tm_other, c_full(s) & kPeerBitMask, 0, k0, (c * 256 + static_cast<int>(crank) * 128) >> 6);
          }
        }
      }
#ifdef TCL_PAIR_TRACE
      GT_ADD(2, clock64() - gt_w0);
#endif
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- gradient MMAs (leader only)
    if (leader && elect_one()) {
      uint32_t it = 0, tg = 0, piece = 0;
      GT_DECL
#ifdef TCL_PAIR_TRACE
      const unsigned long long gt_w0 = clock64();
#endif
      while (walk.next(P, pc)) {
        const GBJobDev& J = P.job[pc.job];
        GT_BEGIN();
        mbar_wait_cluster(acc_empty_bar, (piece & 1) ^ 1);  // both CTAs' read-outs of the previous piece are done
        GT_END(3);
        tc_fence_after();
        for (int t = pc.ta; t < pc.tb; ++t, ++tg) {
          const bool col_side = J.seg[t / J.k_tiles].col_side != 0;
          const uint32_t idesc = col_side ? P.idesc_col : P.idesc_row;
          const int gsl = tg % GB_GSLOTS;
          GT_BEGIN();
          mbar_wait(g_full(gsl), (tg / GB_GSLOTS) & 1);
          GT_END(4);
          tc_fence_after();
          for (int c = 0; c < n_chunk; ++c, ++it) {
            const int s = it % GB_CSTAGES;
            GT_BEGIN();
            mbar_wait(c_full(s), (it / GB_CSTAGES) & 1);
            GT_END(5);
            tc_fence_after();
#pragma unroll
            for (int kb2 = 0; kb2 < 2; ++kb2) {
              const uint32_t a_addr = g_smem + gsl * GB_SLOT + kb2 * (col_side ? 8192 : BW_KB_BYTES);
              const uint64_t ad = col_side ? umma_desc_mn_sw128(a_addr, BW_KB_BYTES) : umma_desc_k_sw128(a_addr);
              const uint32_t a_step = col_side ? 128u : 2u;
              const uint64_t bd = umma_desc_mn_sw128(ring + s * GB_SLOT + kb2 * 8192, BW_KB_BYTES);
#pragma unroll
              for (int kk = 0; kk < BW_BK / 16; ++kk)
                tc_mma_f16_2sm(tmem + c * 256, ad + a_step * kk, bd + 128 * kk, idesc, ((t - pc.ta) | kb2 | kk) != 0);
            }
            tc_commit_2sm(c_empty(s), 0x3);
          }
          tc_commit_2sm(g_empty(gsl), 0x3);
        }
        tc_commit_2sm(acc_full_bar, 0x3);
        ++piece;
      }
#ifdef TCL_PAIR_TRACE
      GT_ADD(6, clock64() - gt_w0);
      GT_ADD(10, tg);
      GT_ADD(11, piece);
#endif
    }
  } else {
    // ---------------------------------------------------------------- accumulator read-out (8 warps per CTA)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const uint32_t stg = base + GBSmem::drain_off + static_cast<uint32_t>(warp - 2) * GB_DRAIN_BYTES;
    uint8_t* stg_ptr = base_ptr + GBSmem::drain_off + (warp - 2) * GB_DRAIN_BYTES;
    const uint32_t acc_empty_leader = map_to_peer(acc_empty_bar, 0u);
    uint32_t piece = 0;
    while (walk.next(P, pc)) {
      const GBJobDev& J = P.job[pc.job];
      mbar_wait(acc_full_bar, piece & 1);
      tc_fence_after();
      const int ib = pc.ib + static_cast<int>(crank);
      const bool live = ib < J.n_rowblocks;  // an odd block count leaves the last pair's second CTA without rows
      int owner = 0, rin = ib * BW_BM;
      if (J.owner_rows > 0) {
        owner = rin / J.owner_rows;
        rin -= owner * J.owner_rows;
      }
      const CUtensorMap* tm = &P.tm_dst[J.dst_first + (live ? owner : 0)];
      const int row0 = J.src_row_base + pc.slot * J.slot_rows + rin + q * 32;
      if (J.dst16) {  // fp16 partials for the rows that cross NVLink (see the one-SM kernel)
#pragma unroll 1
        for (int cc = 0; cc < 4 && live; ++cc) {
          const int col = half * 256 + cc * 64;
          if (col >= P.dim) break;
          uint32_t v[32], w[32];
          tmem_ld_32x32b_x32(tmem_addr(tmem, q * 32, col), v);
          tmem_ld_32x32b_x32(tmem_addr(tmem, q * 32, col + 32), w);
          tc_wait_ld();
          if (lane == 0) bulk_wait_read_all();
          __syncwarp();
          uint8_t* rowp = stg_ptr + lane * 128;
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16) {
            const uint32_t* src = c16 < 4 ? v + 8 * c16 : w + 8 * (c16 - 4);
            *reinterpret_cast<uint4*>(rowp + ((c16 ^ (lane & 7)) << 4)) =
                make_uint4(pack2<TCL_OP_F16>(__uint_as_float(src[0]), __uint_as_float(src[1])),
                           pack2<TCL_OP_F16>(__uint_as_float(src[2]), __uint_as_float(src[3])),
                           pack2<TCL_OP_F16>(__uint_as_float(src[4]), __uint_as_float(src[5])),
                           pack2<TCL_OP_F16>(__uint_as_float(src[6]), __uint_as_float(src[7])));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(tm, stg, col, row0);
            bulk_commit_group();
          }
        }
      } else {
#pragma unroll 1
      for (int cc = 0; cc < 8 && live; ++cc) {
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
          tma_store_2d(tm, stg, col, row0);
          bulk_commit_group();
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(acc_empty_bar);
        else mbar_arrive_remote(acc_empty_leader);
      }
      if (P.fold.enabled && J.owner_rows == 0 && live)
        fold_piece_done(P.fold, pc.job, ib, pc.n_pieces, warp - 2, lane,
                        reinterpret_cast<volatile uint32_t*>(base_ptr + GBSmem::bar_off + 240), 1);
      ++piece;
    }
    if (lane == 0) {
      bulk_wait_all();
      fence_proxy_async_generic();
      __threadfence_system();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves while the other may still signal its barriers / read its shared memory
  if (warp == 1) tmem_dealloc_2sm(tmem, 512);
  if (P.world > 0 && threadIdx.x == 0) {
    uint32_t* sy = P.sync[P.rank];
    const uint32_t e = ld_relaxed_u32(sy + ShardSync::kBwdEpoch) + 1u;
    __threadfence_system();
    if (atomicAdd(sy + ShardSync::kGemmDone, 1u) + 1u == gridDim.x) {
      sy[ShardSync::kGemmDone] = 0u;
      __threadfence_system();
      for (int p = 0; p < P.world; ++p) st_release_sys_u32(P.sync[p] + ShardSync::kGrads + P.rank, e);
      *reinterpret_cast<volatile uint32_t*>(sy + ShardSync::kBwdEpoch) = e;
    }
  }
}

// -----------------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------------
}  // namespace tcl
extern "C" int tcl_debug_gb_trace(unsigned long long* out64, int reset) {
  using namespace tcl;
  TCL_CHECK_CUDA(cudaDeviceSynchronize());
  if (out64) TCL_CHECK_CUDA(cudaMemcpyFromSymbol(out64, g_gb_trace, sizeof(unsigned long long) * 64));
  if (reset) {
    unsigned long long z[64] = {0};
    TCL_CHECK_CUDA(cudaMemcpyToSymbol(g_gb_trace, z, sizeof(z)));
  }
  return TCL_OK;
}
namespace tcl {

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
  LaunchCfg L(dim3(n_ctas), dim3(GA_THREADS), smem, st);
  TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, ntxent_g_kernel<kOp>, A));
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

static int launch_ggemm(const GBParams& B, int n_ctas, cudaStream_t st) {
  const int smem = static_cast<int>(GBSmem::total);
  if (int e = ensure_dyn_smem(ntxent_ggemm_kernel, smem)) return e;
  prof_begin(TCL_K_NTXENT_BWD, st);
  LaunchCfg L(dim3(static_cast<unsigned>(n_ctas)), dim3(GB_THREADS), smem, st);
  TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, ntxent_ggemm_kernel, B));
  prof_end(TCL_K_NTXENT_BWD, st);
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

// TRICOLO_B200_GGEMM=1sm keeps the one-SM gradient GEMM kernel; default: the CTA-pair (cta_group::2) kernel
static bool ggemm_2sm_enabled() {
  const char* e = getenv("TRICOLO_B200_GGEMM");
  return !(e && !strcmp(e, "1sm"));
}

static int launch_ggemm2(const GBParams& B, int n_clusters, cudaStream_t st) {
  const int smem = static_cast<int>(GBSmem::total);
  if (int e = ensure_dyn_smem(ntxent_ggemm2_kernel, smem)) return e;
  LaunchCfg L(dim3(2 * static_cast<unsigned>(n_clusters), 1, 1), dim3(GB_THREADS), smem, st, 2);
  prof_begin(TCL_K_NTXENT_BWD, st);
  TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, ntxent_ggemm2_kernel, B));
  prof_end(TCL_K_NTXENT_BWD, st);
  return TCL_OK;
}

// TRICOLO_B200_FOLD=1 folds the normalise backward into kernel B's read-out (norm_fold.cuh).  Off by default: measured
// at B = 8192 on one GPU the read-out warps are on the critical path (one 512-column accumulator), kernel B grows by
// 58 us while the separate kernel it replaces takes 42 us.  Read per call, not cached.
bool fold_enabled() {
  const char* e = getenv("TRICOLO_B200_FOLD");
  return e && e[0] == '1';
}

static int device_sm_count(int* n_sm) {
  int dev = 0;
  TCL_CHECK_CUDA(cudaGetDevice(&dev));
  TCL_CHECK_CUDA(cudaDeviceGetAttribute(n_sm, cudaDevAttrMultiProcessorCount, dev));
  return TCL_OK;
}

int launch_bwd_sharedg(const BwdSharedGArgs& a, cudaStream_t st) {
  const int64_t batch = a.batch, dim = a.dim;
  const int64_t ld_g = (batch + 63) / 64 * 64;
  const int n_iblocks = static_cast<int>((batch + BW_BM - 1) / BW_BM);
  const int n_jtiles = static_cast<int>((batch + BW_BN - 1) / BW_BN);
  const int n_self_pad = n_iblocks * BW_BM;
  int n_sm = 0;
  if (int e = device_sm_count(&n_sm)) return e;
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
    // all ld_g columns are stored (finite values in the padding), so that kernel B may read whole 64-column chunks
    if (int e = make_tmap_2d_16bit(&G.tm_g, g_mat + static_cast<size_t>(k) * batch * ld_g, batch, ld_g, ld_g, BW_BM, 64)) return e;
    G.lse_row = a.lse_row + static_cast<size_t>(p) * batch;
    G.lse_col = a.lse_col + static_cast<size_t>(p) * batch;
    G.grad_scale = a.grad_losses + p;
  }
  if (A.n_pairs == 0) return TCL_OK;
  A.scale_out[0] = scale;
  A.n_scale_out = 1;
  // normalise backward folded into kernel B's read-out (norm_fold.cuh): block counters behind the partial slots
  const bool fold = fold_enabled() && dim % 4 == 0;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(gbase + static_cast<size_t>(a.n_tensors) * kBwdMaxSplit * n_self_pad * dim);
  if (fold) {
    A.zero_words = cnt;
    A.n_zero_words = FOLD_MAX_JOBS * n_iblocks;
  }
  A.n_rows = A.n_cols = static_cast<int>(batch);
  A.row_offset = 0;
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
  const bool two_sm = ggemm_2sm_enabled();  // CTA-pair kernel: a unit is a PAIR of 128-row blocks
  const int n_units = two_sm ? (n_iblocks + 1) / 2 : n_iblocks;
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
      if (int e = make_tmap_3d_16bit(&S.tm_g, gp, batch, ld_g / 64, ld_g, BW_BM, 2)) return e;
      // one-SM kernel: stage = 64 k rows x 256 dims; CTA pair: stage = 128 k rows x this CTA's 128 dims
      if (int e = make_tmap_3d_16bit(&S.tm_other, a.z[o], batch, dim / 64, dim, ggemm_2sm_enabled() ? 128 : 64,
                                     ggemm_2sm_enabled() ? 2 : 4)) return e;
    }
    if (J.n_seg == 0) continue;
    float* gpart = gbase + static_cast<size_t>(n_jobs) * kBwdMaxSplit * n_self_pad * dim;
    if (int e = make_tmap_2d_f32(&B.tm_dst[n_jobs], gpart, static_cast<uint64_t>(kBwdMaxSplit) * n_self_pad, dim, 32, 32)) return e;
    J.k_tiles = n_jtiles;
    J.n_rowblocks = n_iblocks;
    J.dst_first = n_jobs;
    J.owner_rows = 0;
    J.slot_rows = n_self_pad;
    J.src_row_base = 0;
    B.unit_tiles[n_jobs] = J.n_seg * n_jtiles;
    B.job_tile_base[n_jobs + 1] = B.job_tile_base[n_jobs] + static_cast<int64_t>(n_units) * B.unit_tiles[n_jobs];
    N.job[n_jobs].x = a.x[m];
    N.job[n_jobs].inv_norm = a.inv_norm + static_cast<size_t>(m) * batch;
    N.job[n_jobs].gpart = gpart;
    N.job[n_jobs].scale = scale;
    N.job[n_jobs].dx = a.dx[m];
    ++n_jobs;
  }
  for (int j = n_jobs; j < GB_MAX_JOBS; ++j) B.job_tile_base[j + 1] = B.job_tile_base[n_jobs];
  B.dim = static_cast<int>(dim);
  B.n_dsplit = 1;  // one GPU: long units, G streams from HBM - reading it once per 256-column half would double that
  B.idesc_row = umma_idesc_f16(two_sm ? 256 : BW_BM, 256, a.op_format) | (1u << 16);
  B.idesc_col = B.idesc_row | (1u << 15);
  const int64_t total_b = B.job_tile_base[GB_MAX_JOBS];
  int t_max = 1;
  for (int j = 0; j < n_jobs; ++j) t_max = B.unit_tiles[j] > t_max ? B.unit_tiles[j] : t_max;
  int64_t ctas_b = two_sm ? n_sm / 2 : n_sm;  // tile ranges: one per CTA, or one per CTA pair
  if (ctas_b > total_b) ctas_b = total_b;
  const int64_t cap = (kBwdMaxSplit - 1) * total_b / t_max;
  if (ctas_b > cap) ctas_b = cap;
  if (ctas_b < 1) ctas_b = 1;
  if (fold) {
    B.fold.enabled = 1;
    B.fold.counters = cnt;
    B.fold.x_stride = a.x_row_stride;
    B.fold.slot_stride = static_cast<int64_t>(n_self_pad) * dim;
    B.fold.rows = static_cast<int>(batch);
    B.fold.dim = static_cast<int>(dim);
    B.fold.x_dtype = a.x_dtype;
    B.fold.n_rowblocks = n_iblocks;
    B.fold.eps = a.eps;
    for (int j = 0; j < n_jobs; ++j) {
      B.fold.job[j].x = N.job[j].x;
      B.fold.job[j].dx = N.job[j].dx;
      B.fold.job[j].inv_norm = N.job[j].inv_norm;
      B.fold.job[j].gpart = N.job[j].gpart;
      B.fold.job[j].scale = N.job[j].scale;
    }
  }
  if (int e = two_sm ? launch_ggemm2(B, static_cast<int>(ctas_b), st) : launch_ggemm(B, static_cast<int>(ctas_b), st)) return e;
  if (fold) return TCL_OK;  // the read-out warps have written dx
  N.n_clusters = static_cast<int>(ctas_b);
  N.unit_shift = two_sm ? 8 : 7;
  N.split_rows = n_self_pad;
  N.total_tiles = total_b;
  for (int j = 0; j < TCL_MAX_TENSORS; ++j) {
    N.job_tile_base[j] = B.job_tile_base[j];
    N.unit_tiles[j] = B.unit_tiles[j];
  }
  return launch_l2norm_bwd(N, n_jobs, a.x_dtype, batch, static_cast<int>(dim), a.x_row_stride, kBwdMaxSplit, a.eps, st);
}

// =================================================================================================================
// sharded form: rank r owns rows [r b_loc, (r+1) b_loc) of every tensor
// =================================================================================================================
// The plan (jobs, tile table, slot counts, buffer layout) is a pure function of the sizes, so every rank - and the
// size queries, the GEMM launch and the finishing launch of one rank - derive the same one.
struct ShardPlan {
  int n_row_jobs, n_col_jobs;
  int row_tensor[TCL_MAX_TENSORS], col_tensor[TCL_MAX_TENSORS];  // tensor index of each job
  int n_dsplit, n_ctas, n_slots_col;
  int two_sm;  // CTA-pair gradient GEMM: a unit is a PAIR of 128-row blocks, n_ctas counts clusters (tile ranges)
  int n_iblocks, n_jtiles, n_self_pad;
  int64_t ld_g;
  int64_t job_tile_base[GB_MAX_JOBS + 1];
  int unit_tiles[GB_MAX_JOBS];  // row jobs first, then column jobs: the short column-side units (whose stores cross
                                // NVLink) are NOT first, so that... see make_shard_plan
  int job_is_col[GB_MAX_JOBS], job_tensor[GB_MAX_JOBS];
  int n_jobs;
  size_t ws_scale, ws_part, ws_g, ws_total;  // local workspace offsets / size
  size_t recv_hdr, recv_job_bytes, recv_total;
  int rs16;  // column-side partials travel and are stored as fp16 (TRICOLO_B200_RS16=0: fp32)
  int pair_slot[TCL_MAX_PAIRS], n_gpairs;
};

static int max_pieces(int64_t total, int64_t first, int T, int n_units, int n) {
  int worst = 1;
  for (int u = 0; u < n_units; ++u) {
    const int64_t u0 = first + static_cast<int64_t>(u) * T;
    const int k = pc_range_of(total, u0 + T - 1, n) - pc_range_of(total, u0, n) + 1;
    worst = k > worst ? k : worst;
  }
  return worst;
}

static int make_shard_plan(ShardPlan* out, int n_tensors, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                           const uint8_t* need_grad, int64_t b_loc, int64_t b_glob, int64_t dim, int world) {
  ShardPlan& S = *out;
  memset(&S, 0, sizeof(S));
  TCL_REQUIRE(n_tensors >= 2 && n_tensors <= TCL_MAX_TENSORS && n_pairs >= 1 && n_pairs <= TCL_MAX_PAIRS, TCL_ERR_BAD_ARG,
              "bwd_sharded: %d tensors / %d pairs", n_tensors, n_pairs);
  TCL_REQUIRE(world >= 1 && world <= TCL_MAX_PEERS, TCL_ERR_BAD_ARG, "bwd_sharded: world %d", world);
  TCL_REQUIRE(dim > BW_DH && dim % 64 == 0 && dim <= 512, TCL_ERR_BAD_SHAPE, "bwd_sharded: dim %lld (needs 256 < dim <= 512)", (long long)dim);
  TCL_REQUIRE(b_loc >= BW_BM && b_loc % BW_BM == 0 && b_glob == b_loc * world && b_glob < (1 << 24), TCL_ERR_BAD_SHAPE,
              "bwd_sharded: rows per rank (%lld) must be a multiple of 128 and b_glob = world * b_loc", (long long)b_loc);
  S.n_iblocks = static_cast<int>(b_loc / BW_BM);
  S.n_jtiles = static_cast<int>(b_glob / BW_BN);
  S.n_self_pad = static_cast<int>(b_loc);
  S.ld_g = b_glob;
  for (int p = 0; p < n_pairs; ++p) {
    TCL_REQUIRE(pair_row[p] >= 0 && pair_row[p] < n_tensors && pair_col[p] >= 0 && pair_col[p] < n_tensors &&
                    pair_row[p] != pair_col[p], TCL_ERR_BAD_ARG, "bwd_sharded: pair %d out of range", p);
    S.pair_slot[p] = (need_grad[pair_row[p]] || need_grad[pair_col[p]]) ? S.n_gpairs++ : -1;
  }
  // G of all pairs in L2 (126 MB): halve the accumulator so that read-outs overlap MMAs; otherwise G would stream from
  // HBM once per half
  const size_t g_bytes = static_cast<size_t>(S.n_gpairs) * b_loc * S.ld_g * 2;
  S.n_dsplit = g_bytes <= (64ull << 20) ? 2 : 1;
  if (const char* e = getenv("TRICOLO_B200_GSPLIT")) S.n_dsplit = atoi(e) == 2 ? 2 : 1;
  S.two_sm = ggemm_2sm_enabled() && S.n_dsplit == 1 && b_loc % (2 * BW_BM) == 0;
  const int ush = S.two_sm ? 2 : 1;  // row blocks per unit
  // jobs: column-side first - their drains cross NVLink and should be in flight while the row-side tiles still compute
  for (int pass = 0; pass < 2; ++pass) {
    for (int m = 0; m < n_tensors; ++m) {
      if (!need_grad[m]) continue;
      int n_seg = 0;
      for (int p = 0; p < n_pairs; ++p) n_seg += (pass == 0 ? pair_col[p] : pair_row[p]) == m;
      if (n_seg == 0) continue;
      TCL_REQUIRE(n_seg <= 2, TCL_ERR_BAD_ARG, "bwd_sharded: tensor %d is on one side of more than two pairs", m);
      const int j = S.n_jobs++;
      S.job_is_col[j] = pass == 0;
      S.job_tensor[j] = m;
      const int k_tiles = pass == 0 ? S.n_iblocks : S.n_jtiles;
      const int n_units = (pass == 0 ? S.n_jtiles : S.n_iblocks) * S.n_dsplit / ush;
      S.unit_tiles[j] = n_seg * k_tiles;
      S.job_tile_base[j + 1] = S.job_tile_base[j] + static_cast<int64_t>(n_units) * S.unit_tiles[j];
      if (pass == 0) S.col_tensor[S.n_col_jobs++] = m; else S.row_tensor[S.n_row_jobs++] = m;
    }
  }
  for (int j = S.n_jobs; j < GB_MAX_JOBS; ++j) S.job_tile_base[j + 1] = S.job_tile_base[S.n_jobs];
  const int64_t total = S.job_tile_base[GB_MAX_JOBS];
  // tile ranges: one per SM of a B200 (every rank must derive the same number: it fixes the partial slots a unit uses)
  int64_t n = S.two_sm ? kNumSMsB200 / 2 : kNumSMsB200;
  if (const char* e = getenv("TRICOLO_B200_GCTAS")) { const int v = atoi(e); if (v >= 1 && v < n) n = v; }
  if (n > total) n = total;
  int t_max = 1;
  for (int j = 0; j < S.n_jobs; ++j) t_max = S.unit_tiles[j] > t_max ? S.unit_tiles[j] : t_max;
  const int64_t cap = (kBwdMaxSplit - 1) * total / t_max;
  if (n > cap) n = cap;
  if (n < 1) n = 1;
  S.n_ctas = static_cast<int>(n);
  S.n_slots_col = 1;
  for (int j = 0; j < S.n_jobs; ++j) {
    if (!S.job_is_col[j]) continue;
    const int k = max_pieces(total, S.job_tile_base[j], S.unit_tiles[j], S.n_jtiles * S.n_dsplit / ush, S.n_ctas);
    S.n_slots_col = k > S.n_slots_col ? k : S.n_slots_col;
  }
  TCL_REQUIRE(S.n_slots_col <= kBwdMaxSplit, TCL_ERR_BAD_SHAPE, "bwd_sharded: %d pieces per column unit", S.n_slots_col);
  auto up = [](size_t x) { return (x + 1023) / 1024 * 1024; };
  S.ws_scale = 0;
  S.ws_part = 1024;
  S.ws_g = S.ws_part + up(static_cast<size_t>(S.n_row_jobs) * kBwdMaxSplit * S.n_self_pad * dim * 4);
  S.ws_total = S.ws_g + up(g_bytes) + 1024;
  S.recv_hdr = 1024;  // [world] floats: every source rank's scale
  {
    const char* e = getenv("TRICOLO_B200_RS16");
    S.rs16 = !(e && e[0] == '0');
  }
  S.recv_job_bytes = static_cast<size_t>(world) * S.n_slots_col * b_loc * dim * (S.rs16 ? 2 : 4);
  S.recv_total = S.recv_hdr + static_cast<size_t>(S.n_col_jobs) * S.recv_job_bytes;
  return TCL_OK;
}

}  // namespace tcl

using namespace tcl;

extern "C" size_t tcl_ntxent_bwd_sharded_workspace_bytes(int n_tensors, int n_pairs, const int32_t* pair_row,
                                                         const int32_t* pair_col, const uint8_t* need_grad,
                                                         int64_t b_loc, int64_t b_glob, int64_t dim, int world) {
  ShardPlan S;
  if (!pair_row || !pair_col || !need_grad) return 0;
  if (make_shard_plan(&S, n_tensors, n_pairs, pair_row, pair_col, need_grad, b_loc, b_glob, dim, world)) return 0;
  return S.ws_total;
}

extern "C" size_t tcl_ntxent_bwd_sharded_recv_bytes(int n_tensors, int n_pairs, const int32_t* pair_row,
                                                    const int32_t* pair_col, const uint8_t* need_grad, int64_t b_loc,
                                                    int64_t b_glob, int64_t dim, int world) {
  ShardPlan S;
  if (!pair_row || !pair_col || !need_grad) return 0;
  if (make_shard_plan(&S, n_tensors, n_pairs, pair_row, pair_col, need_grad, b_loc, b_glob, dim, world)) return 0;
  return S.recv_total;
}

extern "C" int tcl_ntxent_bwd_sharded_gemm(int n_tensors, const void* const* z_all, int64_t b_loc, int64_t b_glob,
                                           int64_t dim, int64_t z_row_stride, int rank, int world, int n_pairs,
                                           const int32_t* pair_row, const int32_t* pair_col, int op_format,
                                           float inv_tau, float alpha, const float* lse_row, const float* lse_col,
                                           const float* grad_losses, const uint8_t* need_grad, void* workspace,
                                           size_t workspace_bytes, void* const* recv_ptrs, size_t recv_bytes,
                                           void* const* sync_ptrs, void* stream) {
  TCL_REQUIRE(z_all && pair_row && pair_col && lse_row && lse_col && grad_losses && need_grad && workspace && recv_ptrs,
              TCL_ERR_BAD_ARG, "bwd_sharded_gemm: null pointer");
  TCL_REQUIRE(op_format == TCL_OP_F16 || op_format == TCL_OP_BF16, TCL_ERR_BAD_ARG, "op_format %d", op_format);
  ShardPlan S;
  if (int e = make_shard_plan(&S, n_tensors, n_pairs, pair_row, pair_col, need_grad, b_loc, b_glob, dim, world)) return e;
  TCL_REQUIRE(rank >= 0 && rank < world, TCL_ERR_BAD_ARG, "bwd_sharded_gemm: rank %d of %d", rank, world);
  TCL_REQUIRE(workspace_bytes >= S.ws_total && aligned_to(workspace, 256), TCL_ERR_WORKSPACE, "bwd_sharded_gemm: workspace");
  TCL_REQUIRE(recv_bytes >= S.recv_total, TCL_ERR_WORKSPACE, "bwd_sharded_gemm: receive buffers too small");
  const float c1 = inv_tau * 1.4426950408889634f;
  TCL_REQUIRE(inv_tau > 0.f && 2.f * c1 < 120.f, TCL_ERR_BAD_ARG, "bwd_sharded_gemm: temperature too small (need tau >= 0.025)");
  if (z_row_stride == 0) z_row_stride = dim;
  TCL_REQUIRE(z_row_stride >= dim && z_row_stride % 8 == 0, TCL_ERR_BAD_ALIGN, "bwd_sharded_gemm: z_row_stride");
  for (int r = 0; r < world; ++r)
    TCL_REQUIRE(recv_ptrs[r] && aligned_to(recv_ptrs[r], 256), TCL_ERR_BAD_ALIGN, "bwd_sharded_gemm: receive buffer %d", r);
  if (int e = require_sm100()) return e;
  if (S.n_jobs == 0) return TCL_OK;
  int n_sm = 0;
  if (int e = device_sm_count(&n_sm)) return e;
  TCL_REQUIRE(n_sm >= S.n_ctas * (S.two_sm ? 2 : 1), TCL_ERR_BAD_ARCH, "bwd_sharded_gemm: %d SMs, the plan assumes %d", n_sm,
              S.n_ctas * (S.two_sm ? 2 : 1));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  uint16_t* g_mat = reinterpret_cast<uint16_t*>(ws + S.ws_g);
  const size_t g_pair = static_cast<size_t>(b_loc) * S.ld_g;
  const int64_t row_offset = static_cast<int64_t>(rank) * b_loc;
  auto z_loc = [&](int m) { return static_cast<const char*>(z_all[m]) + static_cast<size_t>(row_offset) * z_row_stride * 2; };

  // ---- kernel A: the rank's row block of every pair's G
  GAParams A;
  memset(&A, 0, sizeof(A));
  for (int p = 0; p < n_pairs; ++p) {
    if (S.pair_slot[p] < 0) continue;
    GPairDev& G = A.pair[A.n_pairs++];
    if (int e = make_tmap_2d_16bit(&G.tm_row, z_loc(pair_row[p]), b_loc, dim, z_row_stride, BW_BM, BW_BK)) return e;
    if (int e = make_tmap_2d_16bit(&G.tm_col, z_all[pair_col[p]], b_glob, dim, z_row_stride, BW_BN, BW_BK)) return e;
    if (int e = make_tmap_2d_16bit(&G.tm_g, g_mat + S.pair_slot[p] * g_pair, b_loc, b_glob, S.ld_g, BW_BM, 64)) return e;
    G.lse_row = lse_row + static_cast<size_t>(p) * b_glob + row_offset;
    G.lse_col = lse_col + static_cast<size_t>(p) * b_glob;
    G.grad_scale = grad_losses + p;
  }
  A.n_scale_out = 0;
  // this rank's scale goes to entry `rank` of every rank's receive-buffer header (the owner multiplies each source's
  // partials by that source's scale); the own buffer is one of them
  for (int r = 0; r < world; ++r) A.scale_out[A.n_scale_out++] = reinterpret_cast<float*>(recv_ptrs[r]) + rank;
  A.n_rows = static_cast<int>(b_loc);
  A.n_cols = static_cast<int>(b_glob);
  A.row_offset = static_cast<int>(row_offset);
  A.num_kb = static_cast<int>(dim / 64);
  A.n_jtiles = S.n_jtiles;
  A.n_iblocks = S.n_iblocks;
  A.c1 = c1;
  A.alpha = alpha;
  A.out_scale = inv_tau / static_cast<float>(b_glob);
  A.idesc = umma_idesc_f16(BW_BM, BW_BN, op_format);
  const int64_t total_a = static_cast<int64_t>(A.n_pairs) * S.n_iblocks * S.n_jtiles;
  const int ctas_a = static_cast<int>(total_a < n_sm ? total_a : n_sm);
  prof_begin(TCL_K_NTXENT_G, st);
  if (int e = (op_format == TCL_OP_F16 ? launch_g_kernel<TCL_OP_F16>(A, ctas_a, st) : launch_g_kernel<TCL_OP_BF16>(A, ctas_a, st)))
    return e;
  prof_end(TCL_K_NTXENT_G, st);

  // ---- kernel B
  GBParams B;
  memset(&B, 0, sizeof(B));
  int n_dst = 0, i_row = 0, i_col = 0;
  for (int j = 0; j < S.n_jobs; ++j) {
    GBJobDev& J = B.job[j];
    const int m = S.job_tensor[j];
    const bool col = S.job_is_col[j] != 0;
    for (int p = 0; p < n_pairs; ++p) {
      if ((col ? pair_col[p] : pair_row[p]) != m) continue;
      GBSegDev& sg = J.seg[J.n_seg++];
      sg.col_side = col;
      uint16_t* gp = g_mat + S.pair_slot[p] * g_pair;
      if (int e = make_tmap_3d_16bit(&sg.tm_g, gp, b_loc, S.ld_g / 64, S.ld_g, BW_BM, 2)) return e;
      const uint32_t brows = S.two_sm ? 128 : 64, bchunks = S.two_sm ? 2 : 4;  // stage shape: see launch_bwd_sharedg
      if (col) {  // B operand: the pair's row tensor, this rank's rows (K = local rows)
        if (int e = make_tmap_3d_16bit(&sg.tm_other, z_loc(pair_row[p]), b_loc, dim / 64, z_row_stride, brows, bchunks)) return e;
      } else {    // B operand: the pair's column tensor, all rows (K = global columns of G)
        if (int e = make_tmap_3d_16bit(&sg.tm_other, z_all[pair_col[p]], b_glob, dim / 64, z_row_stride, brows, bchunks)) return e;
      }
    }
    J.k_tiles = col ? S.n_iblocks : S.n_jtiles;
    J.n_rowblocks = col ? S.n_jtiles : S.n_iblocks;
    J.dst_first = n_dst;
    if (col) {
      for (int r = 0; r < world; ++r) {
        char* dst = static_cast<char*>(recv_ptrs[r]) + S.recv_hdr + static_cast<size_t>(i_col) * S.recv_job_bytes;
        if (S.rs16) {
          if (int e = make_tmap_2d_16bit(&B.tm_dst[n_dst++], dst, static_cast<uint64_t>(world) * S.n_slots_col * b_loc, dim, dim, 32, 64)) return e;
        } else {
          if (int e = make_tmap_2d_f32(&B.tm_dst[n_dst++], dst, static_cast<uint64_t>(world) * S.n_slots_col * b_loc, dim, 32, 32)) return e;
        }
      }
      J.dst16 = S.rs16;
      J.owner_rows = static_cast<int>(b_loc);
      J.slot_rows = static_cast<int>(b_loc);
      J.src_row_base = rank * S.n_slots_col * static_cast<int>(b_loc);
      ++i_col;
    } else {
      float* gpart = reinterpret_cast<float*>(ws + S.ws_part) + static_cast<size_t>(i_row) * kBwdMaxSplit * S.n_self_pad * dim;
      if (int e = make_tmap_2d_f32(&B.tm_dst[n_dst++], gpart, static_cast<uint64_t>(kBwdMaxSplit) * S.n_self_pad, dim, 32, 32)) return e;
      J.owner_rows = 0;
      J.slot_rows = S.n_self_pad;
      J.src_row_base = 0;
      ++i_row;
    }
    B.unit_tiles[j] = S.unit_tiles[j];
  }
  for (int j = 0; j <= GB_MAX_JOBS; ++j) B.job_tile_base[j] = S.job_tile_base[j];
  B.dim = static_cast<int>(dim);
  B.n_dsplit = S.n_dsplit;
  B.idesc_row = umma_idesc_f16(S.two_sm ? 256 : BW_BM, 256, op_format) | (1u << 16);
  B.idesc_col = B.idesc_row | (1u << 15);
  if (sync_ptrs != nullptr) {
    for (int r = 0; r < world; ++r) {
      TCL_REQUIRE(sync_ptrs[r] && aligned_to(sync_ptrs[r], 16), TCL_ERR_BAD_ALIGN, "bwd_sharded_gemm: sync pad %d", r);
      B.sync[r] = static_cast<uint32_t*>(sync_ptrs[r]);
    }
    B.rank = rank;
    B.world = world;
  }
  return S.two_sm ? launch_ggemm2(B, S.n_ctas, st) : launch_ggemm(B, S.n_ctas, st);
}

extern "C" int tcl_ntxent_bwd_sharded_finish(int n_tensors, const void* const* x, int x_dtype, int64_t b_loc,
                                             int64_t b_glob, int64_t dim, int64_t x_row_stride, int rank, int world,
                                             int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                                             const float* inv_norm, const uint8_t* need_grad, float eps,
                                             const void* workspace, const void* recv_own, const void* sync_own,
                                             void* const* dx, void* stream) {
  TCL_REQUIRE(x && pair_row && pair_col && inv_norm && need_grad && workspace && recv_own && dx, TCL_ERR_BAD_ARG,
              "bwd_sharded_finish: null pointer");
  ShardPlan S;
  if (int e = make_shard_plan(&S, n_tensors, n_pairs, pair_row, pair_col, need_grad, b_loc, b_glob, dim, world)) return e;
  TCL_REQUIRE(rank >= 0 && rank < world, TCL_ERR_BAD_ARG, "bwd_sharded_finish: rank %d of %d", rank, world);
  if (int e = require_sm100()) return e;
  NormShParams N;
  memset(&N, 0, sizeof(N));
  const char* ws = static_cast<const char*>(workspace);
  const char* rv = static_cast<const char*>(recv_own);
  int n_out = 0;
  for (int m = 0; m < n_tensors; ++m) {
    if (!need_grad[m]) continue;
    NormShJob& J = N.job[n_out];
    J.x = x[m];
    J.inv_norm = inv_norm + static_cast<size_t>(m) * b_loc;
    J.dx = dx[m];
    TCL_REQUIRE(J.dx != nullptr, TCL_ERR_BAD_ARG, "bwd_sharded_finish: dx[%d] is null", m);
    J.row_job = J.col_job = -1;
    int i_row = 0, i_col = 0;
    for (int j = 0; j < S.n_jobs; ++j) {
      if (S.job_tensor[j] == m && !S.job_is_col[j]) {
        J.row_job = j;
        J.row_part = reinterpret_cast<const float*>(ws + S.ws_part) + static_cast<size_t>(i_row) * kBwdMaxSplit * S.n_self_pad * dim;
      }
      if (S.job_tensor[j] == m && S.job_is_col[j]) {
        J.col_job = j;
        J.col_part = rv + S.recv_hdr + static_cast<size_t>(i_col) * S.recv_job_bytes;
      }
      i_row += !S.job_is_col[j];
      i_col += S.job_is_col[j] != 0;
    }
    if (J.row_job < 0 && J.col_job < 0) continue;
    ++n_out;
  }
  if (n_out == 0) return TCL_OK;
  for (int j = 0; j < GB_MAX_JOBS; ++j) {
    N.job_tile_base[j] = S.job_tile_base[j];
    N.unit_tiles[j] = S.unit_tiles[j];
  }
  N.total_tiles = S.job_tile_base[GB_MAX_JOBS];
  N.n_ranges = S.n_ctas;
  N.unit_shift = S.two_sm ? 8 : 7;
  N.n_dsplit = S.n_dsplit;
  N.world = world;
  N.rank = rank;
  N.n_slots_col = S.n_slots_col;
  N.row_slot_stride = static_cast<int64_t>(S.n_self_pad) * dim;
  N.col16 = S.rs16;
  N.scales = reinterpret_cast<const float*>(rv);
  N.sync = static_cast<const uint32_t*>(sync_own);
  return launch_l2norm_bwd_sharded(N, n_out, x_dtype, b_loc, static_cast<int>(dim), x_row_stride, eps,
                                   static_cast<cudaStream_t>(stream));
}
