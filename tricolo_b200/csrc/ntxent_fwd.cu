// K2 — NT-Xent forward: similarity GEMM on tcgen05 with a fused sum-exp
// epilogue.  Replaces tricolo/loss/nt_xent.py:62-72 (torch.eye, the two
// matmuls, the two log_softmax passes); the B x B logits exist only in TMEM.
//
// Work decomposition: CTA (i-block, j-split, pair) keeps its 128 rows of the row
// operand resident in shared memory (dim/64 swizzled 16 KB K-blocks) and sweeps a
// range of 128-column tiles of the column operand, streamed through a TMA ring.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+TMEM owner), 2..9 = epilogue (two warps per
// TMEM lane quarter, each owning 64 of the tile's 128 columns).
// Two 128-column TMEM accumulators alternate so the epilogue of tile t overlaps
// the MMAs of tile t+1.
//
// Epilogue math (c1 = log2(e)/tau, fixed shift c1 because |S| <= 1 after K1):
//   e_ij = 2^(c1*S_ij - c1); row sums stay in registers across the sweep, column
//   sums of the tile are reduced over the 128 rows with a register butterfly
//   (quad TMEM load layout) and written to a per-(i-block) partial buffer, so the
//   final sums are formed in a fixed order (deterministic, no atomics).
#include <atomic>
#include "host_common.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

static constexpr int FW_BM = 128, FW_BN = 128, FW_BK = 64;
static constexpr int FW_KB_BYTES = FW_BM * FW_BK * 2;  // 16 KB
static constexpr int FW_MAX_KB = 8;                     // dim <= 512
static constexpr int FW_STAGES = 5;
static constexpr int FW_EPI_WARPS = 8;  // two per TMEM lane quarter, 64 tile columns each
static constexpr int FW_EPI_THREADS = FW_EPI_WARPS * 32;
static constexpr int FW_THREADS = 64 + FW_EPI_THREADS;

struct FwdParams {
  CUtensorMap tm_row[TCL_MAX_PAIRS];
  CUtensorMap tm_col[TCL_MAX_PAIRS];
  float* row_part;  // [pairs][2 * n_jsplit][n_rows]  (two column halves per split)
  float* col_part;  // [pairs][n_iblocks][n_cols]
  float* diag2;     // [pairs][n_rows]
  int n_rows, n_cols, row_offset;
  int num_kb, n_jtiles, n_jsplit, n_iblocks;
  float c1;
  uint32_t idesc;
  // CTA-pair kernel: M = 256 instruction descriptor, raw row-operand pointers (the row block goes to TMEM)
  uint32_t idesc_pair;
  const void* z_row[TCL_MAX_PAIRS];
  int64_t z_row_stride;
  // sharded form (tcl_ntxent_fwd_sharded): the column operand is the gathered buffer that the ranks' K1 kernels are
  // still filling over NVLink.  A CTA takes every n_jsplit-th tile of the ARRIVAL order (own rows first, then chunk c
  // of every peer, c = 0, 1, ...) and loads a tile once its chunk's kArrived flag shows this step's epoch.
  const uint32_t* sync;  // own sync pad (host_common.h: ShardSync); nullptr = not sharded: contiguous tile ranges
  int rank, world, chunks_per_rank;
  // fused all-gather (n_push > 0): two extra warps per CTA copy this rank's normalised rows of the pushed modalities
  // from the own gathered buffer into every peer's (16-byte stores over NVLink) while the MMAs run, chunk by chunk in
  // row order, and flag each completed 128-row chunk - the transfer overlaps the tile sweep of the same kernel
  uint16_t* z_peer[TCL_MAX_PEERS];     // every rank's gathered buffer [b_glob, z_row_stride] as mapped here
  uint32_t* sync_peer[TCL_MAX_PEERS];  // every rank's sync pad
  int n_push, dim;
  int n_push_ctas;                     // CTAs (lowest linear ids = first wave, co-resident) that carry push warps' work:
                                       // a later wave could not start while the first one waits for a peer's rows,
                                       // and that peer waits for OUR rows
  int push_off[TCL_MAX_TENSORS];       // element offset of a pushed modality inside a gathered row
};

// number of column tiles of split js, and the t-th of them
__device__ __forceinline__ int fwd_n_tiles(const FwdParams& P, int js) {
  if (P.sync == nullptr)
    return static_cast<int>((static_cast<int64_t>(P.n_jtiles) * (js + 1)) / P.n_jsplit) -
           static_cast<int>((static_cast<int64_t>(P.n_jtiles) * js) / P.n_jsplit);
  return (P.n_jtiles - js + P.n_jsplit - 1) / P.n_jsplit;
}
__device__ __forceinline__ int fwd_tile(const FwdParams& P, int js, int t) {
  if (P.sync == nullptr) return static_cast<int>((static_cast<int64_t>(P.n_jtiles) * js) / P.n_jsplit) + t;
  int k = js + t * P.n_jsplit;  // position in the arrival order
  const int nct = P.chunks_per_rank;
  if (k < nct || P.world == 1) return P.rank * nct + k;
  k -= nct;
  const int c = k / (P.world - 1);
  int peer = P.rank + 1 + (k - c * (P.world - 1));
  if (peer >= P.world) peer -= P.world;
  return peer * nct + c;
}
// push warps: units (row, pushed modality) in row order, strided over all push warps of the grid, so that every warp
// works on chunk 0 first: chunks complete - and are flagged - progressively
__device__ __forceinline__ void fwd_push_flush(const FwdParams& P, uint32_t* sy, int chunk, uint32_t cnt, uint32_t e, int lane) {
  if (chunk < 0 || cnt == 0) return;
  __threadfence_system();  // this lane's stores are performed system-wide
  __syncwarp();
  if (lane == 0) {
    const int left = P.n_rows - chunk * 128;
    const uint32_t per_chunk = static_cast<uint32_t>(left < 128 ? left : 128) * P.n_push;
    if (atomicAdd(sy + ShardSync::kChunkCnt + chunk, cnt) + cnt == per_chunk) {  // this warp completed the chunk
      sy[ShardSync::kChunkCnt + chunk] = 0u;
      __threadfence_system();
      for (int p = 0; p < P.world; ++p)
        if (p != P.rank) st_release_sys_u32(P.sync_peer[p] + ShardSync::kArrived + P.rank * ShardSync::kMaxChunks + chunk, e);
    }
  }
}
__device__ __forceinline__ void fwd_push_rows(const FwdParams& P, int pusher, int n_pushers, int lane) {
  uint32_t* sy = P.sync_peer[P.rank];
  const uint32_t e = ld_acquire_sys_u32(sy + ShardSync::kFwdEpoch);
  const int total = P.n_rows * P.n_push;
  uint32_t ready = 0, cnt = 0;
  int cur = -1;
  for (int u = pusher; u < total; u += n_pushers) {
    const int row = u / P.n_push, mi = u - row * P.n_push;
    if ((row >> 7) != cur) {
      fwd_push_flush(P, sy, cur, cnt, e, lane);
      cur = row >> 7;
      cnt = 0;
    }
    const size_t off = static_cast<size_t>(P.row_offset + row) * P.z_row_stride + P.push_off[mi];
    const uint4* src = reinterpret_cast<const uint4*>(P.z_peer[P.rank] + off);
    const bool h0 = lane * 8 < P.dim, h1 = (lane + 32) * 8 < P.dim;
    const uint4 v0 = h0 ? src[lane] : make_uint4(0, 0, 0, 0);
    const uint4 v1 = h1 ? src[lane + 32] : make_uint4(0, 0, 0, 0);
    for (int d = 1; d < P.world; ++d) {
      int p = P.rank + d;
      if (p >= P.world) p -= P.world;
      if (!((ready >> p) & 1u)) {  // the peer no longer reads its buffer of the previous step
        if (lane == 0) flag_wait_ge(sy + ShardSync::kReady + p, e);
        __syncwarp();
        ready |= 1u << p;
      }
      uint4* dst = reinterpret_cast<uint4*>(P.z_peer[p] + off);
      if (h0) dst[lane] = v0;
      if (h1) dst[lane + 32] = v1;
    }
    ++cnt;
  }
  fwd_push_flush(P, sy, cur, cnt, e, lane);
}

// producer side: tile jt's rows have landed (rank-local tiles were written by this rank's own K1)
__device__ __forceinline__ void fwd_wait_tile(const FwdParams& P, int jt, uint32_t epoch) {
  if (P.sync == nullptr) return;
  const int owner = jt / P.chunks_per_rank;
  if (owner == P.rank) return;
  flag_wait_ge(P.sync + ShardSync::kArrived + owner * ShardSync::kMaxChunks + (jt - owner * P.chunks_per_rank), epoch);
  fence_proxy_async_generic();
}

// shared memory map (offsets from the 1024-aligned base)
struct FwdSmem {
  static constexpr uint32_t x_off = 0;                                      // num_kb * 16 KB
  static constexpr uint32_t ring_off(int num_kb) { return num_kb * FW_KB_BYTES; }
  static constexpr uint32_t bar_off(int num_kb) { return ring_off(num_kb) + FW_STAGES * FW_KB_BYTES; }
  // barriers: full[S], empty[S], x_full, tmem_full[2], tmem_empty[2]  (8 B each), tmem slot (4 B)
  static constexpr uint32_t colbuf_off(int num_kb) { return bar_off(num_kb) + 256; }  // 2*4*128 floats
  static constexpr uint32_t total(int num_kb) { return colbuf_off(num_kb) + 2 * 4 * 128 * 4 + 1024; }
};

// One epilogue warp: TMEM lane quarter q (32 rows), column half ch (64 of the tile's 128 columns).
template <bool kMasked>
__device__ __forceinline__ void fwd_tile_epilogue(uint32_t tmem_acc, int q, int ch, int lane, float c1,
                                                  float (&rs)[4], float (&cs)[16], int row_base,
                                                  int col_base, int n_rows, int n_cols,
                                                  int diag_delta, bool has_diag, float* diag_out) {
  // quad layout: ql = lane/4 -> rows, p = lane%4 -> column pairs
  const int ql = lane >> 2, p = lane & 3;
#pragma unroll
  for (int c = 0; c < 16; ++c) cs[c] = 0.f;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int r_lo = q * 32 + h * 16 + ql;  // tile-local rows r_lo and r_lo + 8
    uint32_t v[2][16];
    tmem_ld_16x256b_x4(tmem_addr(tmem_acc, q * 32 + h * 16, (2 * ch) * 32), v[0]);
    tmem_ld_16x256b_x4(tmem_addr(tmem_acc, q * 32 + h * 16, (2 * ch + 1) * 32), v[1]);
    tc_wait_ld();
#pragma unroll
    for (int cl = 0; cl < 2; ++cl) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = (2 * ch + cl) * 32 + g * 8 + 2 * p + e;  // tile-local column
          const float s0 = __uint_as_float(v[cl][4 * g + e]);      // row r_lo
          const float s1 = __uint_as_float(v[cl][4 * g + 2 + e]);  // row r_lo + 8
          float e0 = ex2_approx(fmaf(s0, c1, -c1));
          float e1 = ex2_approx(fmaf(s1, c1, -c1));
          if (has_diag) {
            // global row index == global column index  <=>  local col == local row + diag_delta
            if (col == r_lo + diag_delta && (!kMasked || row_base + r_lo < n_rows))
              diag_out[r_lo] = s0 * c1;
            if (col == r_lo + 8 + diag_delta && (!kMasked || row_base + r_lo + 8 < n_rows))
              diag_out[r_lo + 8] = s1 * c1;
          }
          if (kMasked) {
            const bool cv = col_base + col < n_cols;
            const bool r0v = row_base + r_lo < n_rows;
            const bool r1v = row_base + r_lo + 8 < n_rows;
            e0 = (cv && r0v) ? e0 : 0.f;
            e1 = (cv && r1v) ? e1 : 0.f;
          }
          rs[2 * h + 0] += e0;
          rs[2 * h + 1] += e1;
          cs[cl * 8 + g * 2 + e] += e0 + e1;
        }
      }
    }
  }
}

// column sums of one tile: butterfly over the 8 row groups of the quad layout, then over the four lane quarters
// through shared memory; one partial per (row block, column)
__device__ __forceinline__ void fwd_col_sums(const float (&cs)[16], float* colbuf, int buf, int bar_id, int q, int ch,
                                             int lane, int p, int et, int j0, int pair, int ib, const FwdParams& P) {
  // column sums: butterfly over the 8 row groups (lane bits 4,3,2): 16 -> 8 -> 4 -> 2 values
  float a8[8], a4[4], a2[2];
  {
    const bool hi = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = hi ? cs[i] : cs[i + 8];
      const float keep = hi ? cs[i + 8] : cs[i];
      a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = (lane & 8) != 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = hi ? a8[i] : a8[i + 4];
      const float keep = hi ? a8[i + 4] : a8[i];
      a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = (lane & 4) != 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = hi ? a4[i] : a4[i + 2];
      const float keep = hi ? a4[i + 2] : a4[i];
      a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  // this thread now owns cidx = bit4*8 + bit3*4 + bit2*2 + i  (cl = cidx>>3, g = (cidx>>1)&3, e = cidx&1)
  float* cb = colbuf + buf * 512 + q * 128;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int cidx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + i;
    const int cl = cidx >> 3, g = (cidx >> 1) & 3, e = cidx & 1;
    cb[(2 * ch + cl) * 32 + g * 8 + 2 * p + e] = a2[i];
  }
  asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
  if (et < 128) {
    const float* c0 = colbuf + buf * 512;
    const float tot = c0[et] + c0[128 + et] + c0[256 + et] + c0[384 + et];
    if (j0 + et < P.n_cols && ib < P.n_iblocks)  // (the pair kernel pads an odd block count with an empty block)
      P.col_part[(static_cast<int64_t>(pair) * P.n_iblocks + ib) * P.n_cols + j0 + et] = tot;
  }
}

// row sums: reduce over the 4 column-pair lanes, then one store per row and column half
__device__ __forceinline__ void fwd_row_sums(float (&rs)[4], int q, int ch, int ql, int p, int i0, int js, int pair,
                                             const FwdParams& P) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
    rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
  }
  if (p == 0) {
    float* rp = P.row_part + (static_cast<int64_t>(pair) * (2 * P.n_jsplit) + 2 * js + ch) * P.n_rows;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int row = i0 + q * 32 + (r >> 1) * 16 + ql + (r & 1) * 8;
      if (row < P.n_rows) rp[row] = rs[r];
    }
  }
}

static constexpr int FW_PUSH_WARPS = 2;
__global__ void __launch_bounds__(FW_THREADS + FW_PUSH_WARPS * 32, 1) ntxent_fwd_kernel(const __grid_constant__ FwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int num_kb = P.num_kb;
  const uint32_t x_smem = base + FwdSmem::x_off;
  const uint32_t ring = base + FwdSmem::ring_off(num_kb);
  const uint32_t bars = base + FwdSmem::bar_off(num_kb);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (FW_STAGES + s); };
  const uint32_t x_full_bar = bars + 8u * (2 * FW_STAGES);
  auto tmem_full_bar = [&](int b) { return bars + 8u * (2 * FW_STAGES + 1 + b); };
  auto tmem_empty_bar = [&](int b) { return bars + 8u * (2 * FW_STAGES + 3 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * FW_STAGES + 5);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + FwdSmem::bar_off(num_kb) + 8u * (2 * FW_STAGES + 5));
  float* colbuf = reinterpret_cast<float*>(base_ptr + FwdSmem::colbuf_off(num_kb));  // [2][4][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ib = blockIdx.x, js = blockIdx.y, pair = blockIdx.z;
  const int i0 = ib * FW_BM;
  const int n_tiles = fwd_n_tiles(P, js);  // contiguous range of column tiles (sharded: strided arrival order)

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&P.tm_row[pair]);
    tma_prefetch_desc(&P.tm_col[pair]);
    for (int s = 0; s < FW_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(x_full_bar, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), FW_EPI_THREADS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  griddep_launch();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  griddep_wait();  // the normalised operands of the preceding kernel are complete

  if (warp == 0) {
    if (elect_one()) {
      // resident row block
      mbar_arrive_expect_tx(x_full_bar, num_kb * FW_KB_BYTES);
      for (int kb = 0; kb < num_kb; ++kb)
        tma_load_2d(x_smem + kb * FW_KB_BYTES, &P.tm_row[pair], x_full_bar, kb * FW_BK, i0);
      int it = 0;
      const uint32_t epoch = P.sync ? ld_acquire_sys_u32(P.sync + ShardSync::kFwdEpoch) : 0u;
      for (int t = 0; t < n_tiles; ++t) {
        const int jt = fwd_tile(P, js, t);
        const int j0 = jt * FW_BN;
        fwd_wait_tile(P, jt, epoch);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % FW_STAGES;
          const uint32_t ph = (it / FW_STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_arrive_expect_tx(full_bar(s), FW_KB_BYTES);
          tma_load_2d(ring + s * FW_KB_BYTES, &P.tm_col[pair], full_bar(s), kb * FW_BK, j0);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      mbar_wait(x_full_bar, 0);
      int it = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int b = t & 1;
        mbar_wait(tmem_empty_bar(b), ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem + b * FW_BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % FW_STAGES;
          const uint32_t ph = (it / FW_STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint64_t ad = umma_desc_k_sw128(x_smem + kb * FW_KB_BYTES);
          const uint64_t bd = umma_desc_k_sw128(ring + s * FW_KB_BYTES);
#pragma unroll
          for (int kk = 0; kk < FW_BK / 16; ++kk)
            tc_mma_f16(acc, ad + 2 * kk, bd + 2 * kk, P.idesc, (kb | kk) != 0);
          tc_commit(empty_bar(s));
        }
        tc_commit(tmem_full_bar(b));
      }
    }
  } else if (warp >= 2 + FW_EPI_WARPS) {
    // push warps (only launched for the sharded form with n_push > 0)
    const int cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    if (cta < P.n_push_ctas)
      fwd_push_rows(P, cta * FW_PUSH_WARPS + (warp - 2 - FW_EPI_WARPS), P.n_push_ctas * FW_PUSH_WARPS, lane);
  } else {
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int ch = (warp - 2) >> 2;  // column half of the tile (64 columns)
    const int et = threadIdx.x - 64;  // 0..255 among epilogue threads
    const int ql = lane >> 2, p = lane & 3;
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    float cs[16];
    float* diag_out = P.diag2 + static_cast<int64_t>(pair) * P.n_rows + i0;
    const bool row_edge = i0 + FW_BM > P.n_rows;
    for (int t = 0; t < n_tiles; ++t) {
      const int b = t & 1;
      const int j0 = fwd_tile(P, js, t) * FW_BN;
      mbar_wait(tmem_full_bar(b), (t >> 1) & 1);
      tc_fence_after();
      // diagonal: global row = row_offset + i0 + r, global col = j0 + c  ->  c = r + delta
      const int diag_delta = P.row_offset + i0 - j0;
      const bool has_diag = diag_delta > -FW_BN && diag_delta < FW_BM;
      const bool masked = row_edge || (j0 + FW_BN > P.n_cols);
      if (masked)
        fwd_tile_epilogue<true>(tmem + b * FW_BN, q, ch, lane, P.c1, rs, cs, i0, j0, P.n_rows, P.n_cols,
                                diag_delta, has_diag, diag_out);
      else
        fwd_tile_epilogue<false>(tmem + b * FW_BN, q, ch, lane, P.c1, rs, cs, i0, j0, P.n_rows, P.n_cols,
                                 diag_delta, has_diag, diag_out);
      // TMEM buffer fully read -> hand it back to the MMA warp
      tc_fence_before();
      mbar_arrive(tmem_empty_bar(b));

      fwd_col_sums(cs, colbuf, t & 1, 1, q, ch, lane, p, et, j0, pair, ib, P);
    }
    fwd_row_sums(rs, q, ch, ql, p, i0, js, pair, P);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair form (TRICOLO_B200_FWD=pair|single): two adjacent row blocks share one column sweep through TMA MULTICAST.
// Why: with one CTA per row block every SM ingests the whole 128 KB column tile per 2048 MMA cycles; the two SMs of
// a TPC then sit on a shared ingest limit of ~85 B/clk (measured on the backward: ~3100 cycles per tile whatever the
// number of active SMs, profiles/r1c_*).  Here each CTA of a 2-CTA cluster loads HALF of every K-block (64 of the 128
// tile rows) and multicasts it into both CTAs' ring slots, so a TPC fetches each tile once.  The MMAs stay 1-SM
// (cta_group::1, M = N = 128: a 2-SM M=256/N=128 version issued at ~111 cycles per MMA instead of 64 and was slower,
// profiles/r1c_fwd_pair_trace.log).  The row block is the A operand and lives in TMEM (128 lanes x dim/2 columns,
// written once with tcgen05.st), so shared memory holds only the ring (6 slots of two 16 KB K-blocks) and its read
// traffic is the B operand alone.  A ring slot is refilled once BOTH CTAs' MMAs have consumed it: every MMA thread
// commits (multicast) onto the `empty` barrier of both CTAs (count 2).
static constexpr int F2_STAGES = 6;
static constexpr int F2_SLOT = 32768;   // two K-blocks of 128 rows x 128 bytes
static constexpr int F2_XCOL = 256;     // first TMEM column of the resident row block
static constexpr int F2_EPI_WARPS = 16; // two groups of 8: group g owns logit buffer g and the tiles t = g (mod 2)
static constexpr int F2_THREADS = 64 + F2_EPI_WARPS * 32;

struct Fwd2Smem {
  static constexpr uint32_t ring_off = 0;
  static constexpr uint32_t bar_off = F2_STAGES * F2_SLOT;
  static constexpr uint32_t colbuf_off = bar_off + 256;                     // [4][4 quarters][128] floats
  static constexpr uint32_t rowbuf_off = colbuf_off + 4 * 4 * 128 * 4;      // [2 column halves][128] floats
  static constexpr uint32_t total = rowbuf_off + 2 * 128 * 4 + 1024;
};
static_assert(Fwd2Smem::total <= 232448, "pair forward: shared memory budget");

// Optional wait-time accounting of the first cluster (make trace; profiles/fwd_trace.py).  Slots: MMA thread of CTA 0
// 0 x_full, 1 tmem_empty, 2 full, 3 total, 4 tiles | TMA thread 5 empty (CTA 0), 6 empty (CTA 1) | epilogue warp 2:
// 7 tmem_full (CTA 0), 8 loads+math (both), 9 column sums (both), 10 total (CTA 0), 11 tmem_full (CTA 1),
// 12 prologue (both), 13 total (CTA 1)
__device__ unsigned long long g_f2_trace[32];
#ifdef TCL_PAIR_TRACE
#define FT_DECL unsigned long long ft_t0 = 0; const bool ft_on = blockIdx.x < 2 && blockIdx.y == 0 && blockIdx.z == 0 && (threadIdx.x & 31) == 0 && ((threadIdx.x >> 5) <= 2); (void)ft_t0;
#define FT_BEGIN() do { if (ft_on) ft_t0 = clock64(); } while (0)
#define FT_END(slot) do { if (ft_on) { const unsigned long long ft_t1 = clock64(); atomicAdd(&g_f2_trace[slot], ft_t1 - ft_t0); ft_t0 = ft_t1; } } while (0)
#else
#define FT_DECL
#define FT_BEGIN() do {} while (0)
#define FT_END(slot) do {} while (0)
#endif

__global__ void __launch_bounds__(F2_THREADS, 1) ntxent_fwd_pair_kernel(const __grid_constant__ FwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int num_kb = P.num_kb;
  const uint32_t ring = base + Fwd2Smem::ring_off;
  const uint32_t bars = base + Fwd2Smem::bar_off;
  auto full_bar = [&](int s) { return bars + 8u * s; };                   // own TMA thread arms it; bytes from both CTAs
  auto empty_bar = [&](int s) { return bars + 8u * (F2_STAGES + s); };    // count 2: the MMA threads of both CTAs
  const uint32_t x_full_bar = bars + 8u * (2 * F2_STAGES);
  auto tmem_full_bar = [&](int b) { return bars + 8u * (2 * F2_STAGES + 1 + b); };
  auto tmem_empty_bar = [&](int b) { return bars + 8u * (2 * F2_STAGES + 3 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * F2_STAGES + 5);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + Fwd2Smem::bar_off + 8u * (2 * F2_STAGES + 5));
  float* colbuf = reinterpret_cast<float*>(base_ptr + Fwd2Smem::colbuf_off);  // [4][4][128]
  float* rowbuf = reinterpret_cast<float*>(base_ptr + Fwd2Smem::rowbuf_off);  // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  FT_DECL
  const int ib = blockIdx.x, js = blockIdx.y, pair = blockIdx.z;  // cluster = two CTAs adjacent in x
  const int i0 = ib * FW_BM;
  const int n_tiles = fwd_n_tiles(P, js);

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&P.tm_col[pair]);
    tma_prefetch_desc(&P.tm_row[pair]);
    for (int s = 0; s < F2_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 2);
    }
    mbar_init(x_full_bar, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), F2_EPI_WARPS / 2);  // the eight warps of the group that owns the buffer
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
  const uint32_t tmem_x = tmem + F2_XCOL;
  griddep_wait();

  // Prologue: the row block travels global -> ring slots (TMA, 16 KB K-blocks) -> registers -> TMEM.  (Per-thread row
  // loads straight from global took ~12k cycles per CTA: 32 distinct lines per load instruction.)  The ring proper
  // starts after the cluster barrier: until then the peer must not multicast into the slots used as staging.
  FT_BEGIN();
  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(x_full_bar, num_kb * FW_KB_BYTES);
    for (int kb = 0; kb < num_kb; ++kb)
      tma_load_2d(ring + kb * FW_KB_BYTES, &P.tm_row[pair], x_full_bar, kb * FW_BK, i0);
  }
  if (warp >= 2) {
    const int ew = warp - 2;
    const int r = (warp & 3) * 32 + lane;           // row == TMEM lane
    const int c0 = ((ew >> 3) * 2 + ((ew >> 2) & 1)) * 2;  // this thread: K-blocks c0, c0 + 1 of its row
    mbar_wait(x_full_bar, 0);
#pragma unroll 1
    for (int c32 = c0; c32 < c0 + 2; ++c32) {
      if (c32 >= num_kb) break;
      const uint8_t* rowp = base_ptr + Fwd2Smem::ring_off + c32 * FW_KB_BYTES + r * 128;
      uint32_t xv[32];
#pragma unroll
      for (int e = 0; e < 8; ++e) {  // 16-byte chunk e of the row sits at chunk position e ^ (r & 7)
        const uint4 u = *reinterpret_cast<const uint4*>(rowp + ((e ^ (r & 7)) << 4));
        xv[4 * e] = u.x; xv[4 * e + 1] = u.y; xv[4 * e + 2] = u.z; xv[4 * e + 3] = u.w;
      }
      tmem_st_32x32b_x32(tmem_addr(tmem_x, (warp & 3) * 32, c32 * 32), xv);
    }
    tc_wait_st();
  }
  tc_fence_before();
  cluster_sync_all();  // row blocks are in TMEM, the staging slots are free in both CTAs, all barriers exist
  tc_fence_after();
  FT_END(12);

  if (warp == 0) {
    if (elect_one()) {
      // this CTA's half of every K-block (tile rows 64 * rank .. + 64), multicast into both CTAs' slots
      uint32_t it = 0;
      const uint32_t epoch = P.sync ? ld_acquire_sys_u32(P.sync + ShardSync::kFwdEpoch) : 0u;
      for (int t = 0; t < n_tiles; ++t) {
        const int jt = fwd_tile(P, js, t);
        const int j0 = jt * FW_BN + static_cast<int>(crank) * 64;
        fwd_wait_tile(P, jt, epoch);
        for (int kb = 0; kb < num_kb; kb += 2, ++it) {
          const int nk = kb + 1 < num_kb ? 2 : 1;
          const int s = it % F2_STAGES;
          FT_BEGIN();
          mbar_wait_cluster(empty_bar(s), ((it / F2_STAGES) & 1) ^ 1);  // slot s is free in BOTH CTAs
          FT_END(crank == 0 ? 5 : 6);
          mbar_arrive_expect_tx(full_bar(s), static_cast<uint32_t>(nk) * FW_KB_BYTES);  // both halves of nk K-blocks
          for (int k2 = 0; k2 < nk; ++k2)
            tma_load_2d_multicast(ring + s * F2_SLOT + k2 * FW_KB_BYTES + crank * 8192u, &P.tm_col[pair], full_bar(s),
                                  (kb + k2) * FW_BK, j0, 0x3);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one() && n_tiles > 0) {
#ifdef TCL_PAIR_TRACE
      const unsigned long long ft_m0 = clock64();
#endif
      uint32_t it = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int b = t & 1;
        FT_BEGIN();
        mbar_wait(tmem_empty_bar(b), ((t >> 1) & 1) ^ 1);
        FT_END(1);
        tc_fence_after();
        const uint32_t acc = tmem + b * FW_BN;
        for (int kb = 0; kb < num_kb; kb += 2, ++it) {
          const int nk = kb + 1 < num_kb ? 2 : 1;
          const int s = it % F2_STAGES;
          FT_BEGIN();
          mbar_wait(full_bar(s), (it / F2_STAGES) & 1);
          FT_END(2);
          tc_fence_after();
          for (int k2 = 0; k2 < nk; ++k2) {
            const uint32_t ax = tmem_x + (kb + k2) * (FW_BK / 2);  // 32 columns per K-block of 64
            const uint64_t bd = umma_desc_k_sw128(ring + s * F2_SLOT + k2 * FW_KB_BYTES);
#pragma unroll
            for (int kk = 0; kk < FW_BK / 16; ++kk)
              tc_mma_f16_ts(acc, ax + 8 * kk, bd + 2 * kk, P.idesc, (kb | k2 | kk) != 0);
          }
          tc_commit_multicast(empty_bar(s), 0x3);  // this CTA is done with slot s: tell both TMA threads
        }
        tc_commit(tmem_full_bar(b));
      }
#ifdef TCL_PAIR_TRACE
      if (ft_on) { atomicAdd(&g_f2_trace[3], clock64() - ft_m0); atomicAdd(&g_f2_trace[4], (unsigned long long)n_tiles); }
#endif
    }
  } else {
    const int ew = warp - 2;
    const int gi = ew >> 3;           // group: logit buffer gi, tiles t = gi (mod 2)
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int ch = (ew >> 2) & 1;     // column half of the tile (64 columns)
    const int et = (ew & 7) * 32 + lane;  // 0..255 inside the group
    const int ql = lane >> 2, p = lane & 3;
#ifdef TCL_PAIR_TRACE
    const unsigned long long ft_e0 = clock64();
#endif
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    float cs[16];
    float* diag_out = P.diag2 + static_cast<int64_t>(pair) * P.n_rows + i0;
    const bool row_edge = i0 + FW_BM > P.n_rows;
    for (int t = gi; t < n_tiles; t += 2) {
      const int j0 = fwd_tile(P, js, t) * FW_BN;
      FT_BEGIN();
      mbar_wait(tmem_full_bar(gi), (t >> 1) & 1);
      FT_END(crank == 0 ? 7 : 11);
      tc_fence_after();
      const int diag_delta = P.row_offset + i0 - j0;
      const bool has_diag = diag_delta > -FW_BN && diag_delta < FW_BM;
      const bool masked = row_edge || (j0 + FW_BN > P.n_cols);
      if (masked)
        fwd_tile_epilogue<true>(tmem + gi * FW_BN, q, ch, lane, P.c1, rs, cs, i0, j0, P.n_rows, P.n_cols,
                                diag_delta, has_diag, diag_out);
      else
        fwd_tile_epilogue<false>(tmem + gi * FW_BN, q, ch, lane, P.c1, rs, cs, i0, j0, P.n_rows, P.n_cols,
                                 diag_delta, has_diag, diag_out);
      FT_END(8);
      // this warp's part of the logit buffer is in registers -> release it to the MMA thread
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(gi));
      // the group's column-sum buffers alternate between its consecutive tiles
      fwd_col_sums(cs, colbuf, gi * 2 + ((t >> 1) & 1), 1 + gi, q, ch, lane, p, et, j0, pair, ib, P);
      FT_END(9);
    }
#ifdef TCL_PAIR_TRACE
    if (ft_on) atomicAdd(&g_f2_trace[crank == 0 ? 10 : 13], clock64() - ft_e0);
#endif
    // row sums: over the 4 column-pair lanes, then group 1 hands its sums to group 0 through shared memory
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
    }
    if (gi == 1 && p == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r) rowbuf[ch * 128 + q * 32 + (r >> 1) * 16 + ql + (r & 1) * 8] = rs[r];
    }
    asm volatile("bar.sync 3, 512;" ::: "memory");
    if (gi == 0 && p == 0) {
      float* rp = P.row_part + (static_cast<int64_t>(pair) * (2 * P.n_jsplit) + 2 * js + ch) * P.n_rows;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int rl = q * 32 + (r >> 1) * 16 + ql + (r & 1) * 8;
        if (i0 + rl < P.n_rows) rp[i0 + rl] = rs[r] + rowbuf[ch * 128 + rl];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves while the other may still multicast into its memory / signal its barriers
  if (warp == 1) tmem_dealloc(tmem, 512);
}
}  // namespace tcl
extern "C" int tcl_debug_fwd_trace(unsigned long long* out32, int reset) {
  using namespace tcl;
  TCL_CHECK_CUDA(cudaDeviceSynchronize());
  if (out32) TCL_CHECK_CUDA(cudaMemcpyFromSymbol(out32, g_f2_trace, sizeof(unsigned long long) * 32));
  if (reset) {
    unsigned long long z[32] = {0};
    TCL_CHECK_CUDA(cudaMemcpyToSymbol(g_f2_trace, z, sizeof(z)));
  }
  return TCL_OK;
}
namespace tcl {

// fixed-order reduction of the partial buffers
__global__ void fwd_reduce_kernel(const float* __restrict__ row_part, const float* __restrict__ col_part,
                                  float* __restrict__ row_sum, float* __restrict__ col_sum, int n_pairs,
                                  int n_rows, int n_cols, int n_jsplit, int n_iblocks) {
  griddep_launch();
  griddep_wait();
  const int pair = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_rows) {
    float a = 0.f;
    for (int s = 0; s < n_jsplit; ++s) a += row_part[(static_cast<int64_t>(pair) * n_jsplit + s) * n_rows + i];
    row_sum[static_cast<int64_t>(pair) * n_rows + i] = a;
  }
  if (i < n_cols) {
    float a = 0.f;
    for (int b = 0; b < n_iblocks; ++b) a += col_part[(static_cast<int64_t>(pair) * n_iblocks + b) * n_cols + i];
    col_sum[static_cast<int64_t>(pair) * n_cols + i] = a;
  }
}

// lse + loss; one cluster of FIN_CTAS blocks per pair, fixed-order trees (inside a block, then over the blocks in
// rank order through distributed shared memory) so the result is reproducible and needs no global scratch
static constexpr int FIN_CTAS = 8;
static constexpr int FIN_LANES = 8;
// pairs whose loss has been written, per launch lane (host counter): the last one adds the pair losses; self-resetting
__device__ unsigned int g_fin_done[FIN_LANES];
__global__ void __launch_bounds__(1024) fwd_finalize_kernel(
    int n_rows, int n_cols, int row_offset, float c1, float alpha, float* __restrict__ row_sum,
    float* __restrict__ col_sum, const float* __restrict__ diag2, float* __restrict__ lse2_row,
    float* __restrict__ lse2_col, float* __restrict__ loss_parts, float* __restrict__ loss,
    const float* __restrict__ row_part, const float* __restrict__ col_part, int n_row_slots, int n_iblocks,
    int total_lane) {  // >= 0: loss[n_pairs] = sum of the pair losses (fp32, pair order), by the last pair to finish
  griddep_launch();
  griddep_wait();
  const int pair = blockIdx.y;
  const int crank = static_cast<int>(cluster_ctarank());
  float* rs = row_sum + static_cast<int64_t>(pair) * n_rows;
  float* cs = col_sum + static_cast<int64_t>(pair) * n_cols;
  const float* dg = diag2 + static_cast<int64_t>(pair) * n_rows;
  float* lr = lse2_row + static_cast<int64_t>(pair) * n_rows;
  float* lc = lse2_col + static_cast<int64_t>(pair) * n_cols;
  const int tid = crank * blockDim.x + threadIdx.x, nth = FIN_CTAS * blockDim.x;
  if (row_part != nullptr) {
    // fused with the reduction of the tile kernel's partials (single-GPU whole-loss entry: n_rows == n_cols,
    // row_offset == 0): the same fixed summation order as fwd_reduce_kernel
    for (int i = tid; i < n_rows; i += nth) {
      float a = 0.f;
      for (int s = 0; s < n_row_slots; ++s) a += row_part[(static_cast<int64_t>(pair) * n_row_slots + s) * n_rows + i];
      rs[i] = a;
    }
    for (int j = tid; j < n_cols; j += nth) {
      float a = 0.f;
      for (int b = 0; b < n_iblocks; ++b) a += col_part[(static_cast<int64_t>(pair) * n_iblocks + b) * n_cols + j];
      cs[j] = a;
    }
    // a thread reads back only what it wrote itself (same index sets: tid + k nth), so no barrier is needed as long
    // as n_rows == n_cols and row_offset == 0 (checked on the host)
  }
  for (int j = tid; j < n_cols; j += nth) lc[j] = log2f(cs[j]) + c1;
  double a = 0.0, b = 0.0;
  for (int i = tid; i < n_rows; i += nth) {
    const float l = log2f(rs[i]) + c1;
    lr[i] = l;
    a += static_cast<double>(l - dg[i]);
    const int j = row_offset + i;
    if (j < n_cols) b += static_cast<double>((log2f(cs[j]) + c1) - dg[i]);
  }
  __shared__ double sa[1024], sb[1024];
  __shared__ double part[2 * FIN_CTAS];  // rank 0's copy collects the block sums
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sa[threadIdx.x] += sa[threadIdx.x + o];
      sb[threadIdx.x] += sb[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const uint32_t dst = map_to_peer(smem_u32(part), 0u) + 16u * crank;
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(dst), "d"(sa[0]) : "memory");
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(dst + 8u), "d"(sb[0]) : "memory");
  }
  cluster_sync_all();
  if (crank == 0 && threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int r = 0; r < FIN_CTAS; ++r) {
      ta += part[2 * r];
      tb += part[2 * r + 1];
    }
    const double ln2 = 0.69314718055994530942;
    const double pa = ta * ln2, pb = tb * ln2;
    loss_parts[pair * 2 + 0] = static_cast<float>(pa);
    loss_parts[pair * 2 + 1] = static_cast<float>(pb);
    if (loss != nullptr) {
      __stcg(loss + pair, static_cast<float>((alpha * pa + (1.0 - alpha) * pb) / n_cols));
      if (total_lane >= 0) {
        const int n_pairs = gridDim.y;
        __threadfence();
        if (atomicAdd(g_fin_done + total_lane, 1u) + 1u == static_cast<unsigned int>(n_pairs)) {
          g_fin_done[total_lane] = 0u;
          __threadfence();
          float t = 0.f;
          for (int p = 0; p < n_pairs; ++p) t += __ldcg(loss + p);  // sum(loss_dict.values()), tricolo_net.py:64
          loss[n_pairs] = t;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Sharded statistics exchange (replaces fwd_reduce + zero/copy kernels + barrier + peer_sum of the first version):
//   fwd_reduce_push   reduces the tile kernel's partials and stores this rank's column sum-exp partials [P][B], row
//                     sum-exp [P][b_loc] and positives [P][b_loc] into slot `rank` of EVERY rank's statistics buffer
//                     (plain stores over NVLink); the last block signals kStats[rank] = epoch to every rank;
//   fwd_finalize_sharded  waits for the W flags, adds the column partials in rank order (bit-identical on every
//                     rank), and finalises ALL rows: lse2_row [P][B], lse2_col [P][B], loss [P].
// Slot layout (floats): col [P][B] | row [P][b_loc] | diag [P][b_loc], padded to a multiple of 4.
// ---------------------------------------------------------------------------------------------------------------
static inline int64_t shard_stats_slot_floats(int n_pairs, int64_t b_loc, int64_t b_glob) {
  return (static_cast<int64_t>(n_pairs) * (b_glob + 2 * b_loc) + 3) / 4 * 4;
}
struct StatsPushParams {
  float* dst[TCL_MAX_PEERS];
  uint32_t* sync[TCL_MAX_PEERS];
  const float* row_part;
  const float* col_part;
  const float* diag2;
  int64_t slot_floats;
  int n_pairs, n_rows, n_cols, n_row_slots, n_iblocks, rank, world;
  int local_only;  // 1: slot `rank` of the OWN buffer only, no flag: the ranks pull after a barrier (tcl_ntxent_finalize_sharded)
};
__global__ void __launch_bounds__(256) fwd_reduce_push_kernel(const __grid_constant__ StatsPushParams P) {
  griddep_launch();
  griddep_wait();
  const int pair = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t* sy = P.sync[P.rank];
  const uint32_t e = P.local_only ? 0u : ld_relaxed_u32(sy + ShardSync::kFwdEpoch);  // published by this step's K1
  const int64_t slot = static_cast<int64_t>(P.rank) * P.slot_floats;
  const int n_dst = P.local_only ? 1 : P.world;
  if (i < P.n_cols) {
    float a = 0.f;
    for (int b = 0; b < P.n_iblocks; ++b) a += P.col_part[(static_cast<int64_t>(pair) * P.n_iblocks + b) * P.n_cols + i];
    for (int d = 0; d < n_dst; ++d) {
      const int p = (P.rank + d) % P.world;
      P.dst[p][slot + static_cast<int64_t>(pair) * P.n_cols + i] = a;
    }
  }
  if (i < P.n_rows) {
    float a = 0.f;
    for (int s = 0; s < P.n_row_slots; ++s) a += P.row_part[(static_cast<int64_t>(pair) * P.n_row_slots + s) * P.n_rows + i];
    const float dg = P.diag2[static_cast<int64_t>(pair) * P.n_rows + i];
    const int64_t ro = slot + static_cast<int64_t>(P.n_pairs) * P.n_cols + static_cast<int64_t>(pair) * P.n_rows + i;
    for (int d = 0; d < n_dst; ++d) {
      const int p = (P.rank + d) % P.world;
      P.dst[p][ro] = a;
      P.dst[p][ro + static_cast<int64_t>(P.n_pairs) * P.n_rows] = dg;
    }
  }
  if (P.local_only) return;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(sy + ShardSync::kStatsDone, 1u) + 1u == gridDim.x * gridDim.y) {
    sy[ShardSync::kStatsDone] = 0u;
    __threadfence_system();
    for (int p = 0; p < P.world; ++p) st_release_sys_u32(P.sync[p] + ShardSync::kStats + P.rank, e);
  }
}

struct StatsSrc {
  const float* slot[TCL_MAX_PEERS];  // slot s = rank s's statistics: in the own buffer (flags) or in rank s's (pull)
};
__global__ void __launch_bounds__(1024) fwd_finalize_sharded_kernel(
    int b_loc, int b_glob, int n_pairs, int world, float c1, float alpha, const __grid_constant__ StatsSrc src,
    const uint32_t* __restrict__ sync, float* __restrict__ lse2_row, float* __restrict__ lse2_col,
    float* __restrict__ loss) {
  griddep_launch();
  griddep_wait();
  const int pair = blockIdx.y;
  const int crank = static_cast<int>(cluster_ctarank());
  if (sync != nullptr) {
    const uint32_t e = ld_relaxed_u32(sync + ShardSync::kFwdEpoch);
    if (static_cast<int>(threadIdx.x) < world) flag_wait_ge(sync + ShardSync::kStats + threadIdx.x, e);
    __syncthreads();
  }
  float* lr = lse2_row + static_cast<int64_t>(pair) * b_glob;
  float* lc = lse2_col + static_cast<int64_t>(pair) * b_glob;
  const int tid = crank * blockDim.x + threadIdx.x, nth = FIN_CTAS * blockDim.x;
  const int64_t row_base = static_cast<int64_t>(n_pairs) * b_glob + static_cast<int64_t>(pair) * b_loc;
  const int64_t diag_base = row_base + static_cast<int64_t>(n_pairs) * b_loc;
  double a = 0.0, b = 0.0;
  for (int j = tid; j < b_glob; j += nth) {
    float part[TCL_MAX_PEERS];  // column j: partials of all ranks (all loads in flight: they may cross NVLink)
#pragma unroll
    for (int s = 0; s < TCL_MAX_PEERS; ++s)
      part[s] = s < world ? __ldcv(src.slot[s] + static_cast<int64_t>(pair) * b_glob + j) : 0.f;
    const int so = j / b_loc, i = j - so * b_loc;  // row j belongs to rank so
    const float rsum = __ldcv(src.slot[so] + row_base + i);
    const float dg = __ldcv(src.slot[so] + diag_base + i);
    float cs = 0.f;  // rank order: bit-identical on every rank
#pragma unroll
    for (int s = 0; s < TCL_MAX_PEERS; ++s)
      if (s < world) cs += part[s];
    const float lcj = log2f(cs) + c1;
    lc[j] = lcj;
    const float l = log2f(rsum) + c1;
    lr[j] = l;
    a += static_cast<double>(l - dg);
    b += static_cast<double>(lcj - dg);
  }
  __shared__ double sa[1024], sb[1024];
  __shared__ double part[2 * FIN_CTAS];
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sa[threadIdx.x] += sa[threadIdx.x + o];
      sb[threadIdx.x] += sb[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const uint32_t dst = map_to_peer(smem_u32(part), 0u) + 16u * crank;
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(dst), "d"(sa[0]) : "memory");
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(dst + 8u), "d"(sb[0]) : "memory");
  }
  cluster_sync_all();
  if (crank == 0 && threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int r = 0; r < FIN_CTAS; ++r) {
      ta += part[2 * r];
      tb += part[2 * r + 1];
    }
    const double ln2 = 0.69314718055994530942;
    loss[pair] = static_cast<float>((alpha * ta * ln2 + (1.0 - alpha) * tb * ln2) / b_glob);
  }
}

// Large batches on one GPU: the wide reduction of the tile kernel's partials AND the finalise in ONE launch (the
// reduce -> finalise pair cost 10 + 12 us at B = 8192 for ~7 MB of partials).  Fixed-order sums of the partials (the
// column partials as four quarter sums, combined pairwise), lse2 = log2(sum) + c1, block sums of (lse2_row - diag, lse2_col - diag) in fp64; the
// block that finishes last for a pair adds the block sums in block order and writes the pair's loss, the last pair
// adds the pair losses in pair order.  Requires n_rows == n_cols, row_offset == 0 (the whole-loss entry).  Counters
// and block sums live in per-launch-lane device arrays (self-resetting), so concurrent streams do not collide.
static constexpr int RF_MAX_BLOCKS = 1024;  // 64 indices per block: batch <= 65536
__device__ unsigned int g_rf_done[FIN_LANES][TCL_MAX_PAIRS + 1];
__device__ double g_rf_part[FIN_LANES][TCL_MAX_PAIRS][RF_MAX_BLOCKS][2];
__global__ void __launch_bounds__(256) fwd_reduce_finalize_kernel(
    int n, float c1, float alpha, float* __restrict__ row_sum, float* __restrict__ col_sum,
    const float* __restrict__ diag2, float* __restrict__ lse2_row, float* __restrict__ lse2_col,
    float* __restrict__ loss_parts, float* __restrict__ loss, const float* __restrict__ row_part,
    const float* __restrict__ col_part, int n_row_slots, int n_iblocks, int lane_id, int want_total) {
  griddep_launch();
  griddep_wait();
  __shared__ float cpart[4][64];
  __shared__ double sa[64], sb[64];
  __shared__ bool last;
  const int pair = blockIdx.y;
  // 64 indices per block, four quarter-sums of the column partials per index (a warp reads 32 consecutive floats
  // of one partial row per load; 16 loads in flight per thread at 64 row blocks); quarter 0 also adds the row slots
  const int q = threadIdx.x >> 6, li = threadIdx.x & 63;
  const int i = blockIdx.x * 64 + li;
  float r = 0.f;
  {
    const int per = (n_iblocks + 3) >> 2;
    const int b0 = q * per, b1 = (b0 + per < n_iblocks) ? b0 + per : n_iblocks;
    float c = 0.f;
    if (i < n) {
      const float* cp = col_part + static_cast<int64_t>(pair) * n_iblocks * n + i;
      int bk = b0;
      for (; bk + 16 <= b1; bk += 16) {
        float v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = __ldcs(cp + static_cast<int64_t>(bk + u) * n);
#pragma unroll
        for (int u = 0; u < 16; ++u) c += v[u];
      }
      for (; bk < b1; ++bk) c += __ldcs(cp + static_cast<int64_t>(bk) * n);
      if (q == 0)
        for (int s = 0; s < n_row_slots; ++s) r += row_part[(static_cast<int64_t>(pair) * n_row_slots + s) * n + i];
    }
    cpart[q][li] = c;
  }
  __syncthreads();
  double a = 0.0, b = 0.0;
  if (q == 0 && i < n) {
    const float c = (cpart[0][li] + cpart[1][li]) + (cpart[2][li] + cpart[3][li]);  // fixed order
    const int64_t o = static_cast<int64_t>(pair) * n + i;
    row_sum[o] = r;
    col_sum[o] = c;
    const float lr = log2f(r) + c1, lc = log2f(c) + c1, dg = diag2[o];
    lse2_row[o] = lr;
    lse2_col[o] = lc;
    a = static_cast<double>(lr - dg);
    b = static_cast<double>(lc - dg);
  }
  if (q == 0) {
    sa[li] = a;
    sb[li] = b;
  }
  __syncthreads();
  for (int o = 32; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sa[threadIdx.x] += sa[threadIdx.x + o];
      sb[threadIdx.x] += sb[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    g_rf_part[lane_id][pair][blockIdx.x][0] = sa[0];
    g_rf_part[lane_id][pair][blockIdx.x][1] = sb[0];
    __threadfence();
    last = atomicAdd(&g_rf_done[lane_id][pair], 1u) + 1u == gridDim.x;
  }
  __syncthreads();
  if (!last) return;
  // the pair's last block: block sums in block order (a fixed tree over 64 threads, then thread 0)
  __threadfence();
  double ta = 0.0, tb = 0.0;
  if (threadIdx.x < 64) {
    for (unsigned blk = threadIdx.x; blk < gridDim.x; blk += 64) {
      ta += __ldcg(&g_rf_part[lane_id][pair][blk][0]);
      tb += __ldcg(&g_rf_part[lane_id][pair][blk][1]);
    }
    sa[threadIdx.x] = ta;
    sb[threadIdx.x] = tb;
  }
  __syncthreads();
  for (int o = 32; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sa[threadIdx.x] += sa[threadIdx.x + o];
      sb[threadIdx.x] += sb[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    g_rf_done[lane_id][pair] = 0u;
    const double ln2 = 0.69314718055994530942;
    const double pa = sa[0] * ln2, pb = sb[0] * ln2;
    loss_parts[pair * 2 + 0] = static_cast<float>(pa);
    loss_parts[pair * 2 + 1] = static_cast<float>(pb);
    if (loss != nullptr) {
      __stcg(loss + pair, static_cast<float>((alpha * pa + (1.0 - alpha) * pb) / n));
      if (want_total) {
        const int n_pairs = gridDim.y;
        __threadfence();
        if (atomicAdd(&g_rf_done[lane_id][TCL_MAX_PAIRS], 1u) + 1u == static_cast<unsigned int>(n_pairs)) {
          g_rf_done[lane_id][TCL_MAX_PAIRS] = 0u;
          __threadfence();
          float t = 0.f;
          for (int p = 0; p < n_pairs; ++p) t += __ldcg(loss + p);  // sum(loss_dict.values()), tricolo_net.py:64
          loss[n_pairs] = t;
        }
      }
    }
  }
}

static int fwd_split(int n_pairs, int n_iblocks, int n_jtiles) {
  // One CTA per SM.  Cost model of a split s of the column sweep: waves(s) x (prologue + tiles per CTA), the
  // prologue (row block into shared memory / TMEM, pipeline fill) being worth about two tiles.  Small problems end up
  // with one tile per CTA (latency), large ones with the split that fills whole waves of 148 SMs.
  const int ctas = n_pairs * n_iblocks;
  const double prologue = 2.0;
  int best = 1;
  double best_cost = 0.0;
  const int max_split = n_jtiles < 64 ? n_jtiles : 64;
  for (int s = 1; s <= max_split; ++s) {
    const int total = ctas * s;
    const int waves = (total + kNumSMsB200 - 1) / kNumSMsB200;
    const double cost = waves * (prologue + static_cast<double>((n_jtiles + s - 1) / s));
    if (s == 1 || cost < best_cost * 0.98) {
      best_cost = cost;
      best = s;
    }
  }
  return best;
}

}  // namespace tcl

using namespace tcl;

extern "C" size_t tcl_ntxent_fwd_workspace_bytes(int n_pairs, int64_t n_rows, int64_t n_cols) {
  if (n_pairs < 1 || n_rows < 1 || n_cols < 1) return 0;
  const int n_iblocks = static_cast<int>((n_rows + FW_BM - 1) / FW_BM);
  const int n_jtiles = static_cast<int>((n_cols + FW_BN - 1) / FW_BN);
  const int n_jsplit = fwd_split(n_pairs, n_iblocks, n_jtiles);
  return sizeof(float) * static_cast<size_t>(n_pairs) *
         (static_cast<size_t>(2 * n_jsplit) * n_rows + static_cast<size_t>(n_iblocks) * n_cols);
}

struct FwdShard {
  const uint32_t* sync;  // own sync pad
  int rank, world;
  // fused all-gather (n_push > 0)
  void* const* z_base;     // [world] gathered buffers
  void* const* sync_all;   // [world] sync pads
  int n_push;
  const int64_t* push_off;
};
static int ntxent_fwd_impl(int n_pairs, const void* const* zrow, const void* const* zcol,
                           int64_t n_rows, int64_t n_cols, int64_t dim, int64_t z_row_stride, int64_t row_offset,
                           int op_format, float inv_tau, float* row_sumexp, float* col_sumexp,
                           float* diag2, void* workspace, size_t workspace_bytes, void* stream, FwdPartials* parts,
                           const FwdShard* shard = nullptr) {
  TCL_REQUIRE(n_pairs >= 1 && n_pairs <= TCL_MAX_PAIRS, TCL_ERR_BAD_ARG, "ntxent_fwd: n_pairs %d", n_pairs);
  TCL_REQUIRE(n_rows >= 1 && n_cols >= 1 && n_rows < (1 << 24) && n_cols < (1 << 24), TCL_ERR_BAD_SHAPE,
              "ntxent_fwd: batch sizes out of range (%lld x %lld)", (long long)n_rows, (long long)n_cols);
  TCL_REQUIRE(dim >= 64 && dim % 64 == 0 && dim <= 64 * FW_MAX_KB, TCL_ERR_BAD_SHAPE,
              "ntxent_fwd: dim must be a multiple of 64 in [64, 512] (got %lld)", (long long)dim);
  TCL_REQUIRE(row_offset >= 0 && row_offset + n_rows <= n_cols, TCL_ERR_BAD_SHAPE,
              "ntxent_fwd: local rows [%lld, %lld) must lie inside the global batch %lld",
              (long long)row_offset, (long long)(row_offset + n_rows), (long long)n_cols);
  TCL_REQUIRE(op_format == TCL_OP_F16 || op_format == TCL_OP_BF16, TCL_ERR_BAD_ARG, "op_format %d", op_format);
  const float c1 = inv_tau * 1.4426950408889634f;
  TCL_REQUIRE(inv_tau > 0.f && 2.f * c1 < 120.f, TCL_ERR_BAD_ARG,
              "ntxent_fwd: temperature %g too small for the fixed-shift sum-exp (need tau >= 0.025)", 1.0 / inv_tau);
  TCL_REQUIRE((parts || (row_sumexp && col_sumexp)) && diag2 && workspace, TCL_ERR_BAD_ARG, "ntxent_fwd: null pointer");
  TCL_REQUIRE(workspace_bytes >= tcl_ntxent_fwd_workspace_bytes(n_pairs, n_rows, n_cols), TCL_ERR_WORKSPACE,
              "ntxent_fwd: workspace too small");
  if (z_row_stride == 0) z_row_stride = dim;
  TCL_REQUIRE(z_row_stride >= dim && z_row_stride % 8 == 0, TCL_ERR_BAD_ALIGN, "ntxent_fwd: z_row_stride");
  if (int e = require_sm100()) return e;

  FwdParams P;
  memset(&P, 0, sizeof(P));
  // The multicast CTA-pair kernel is the default for large batches (>= 2048 rows: 0.188 vs 0.195 ms at 8192 x 3
  // pairs); below that the one-CTA-per-row-block kernel's shorter prologue wins.  TRICOLO_B200_FWD=single|pair forces one.
  static const int fwd_mode = [] {
    const char* e = getenv("TRICOLO_B200_FWD");
    return e && !strcmp(e, "single") ? 1 : (e && !strcmp(e, "pair") ? 2 : 0);
  }();
  // (the push warps of the fused all-gather live in the one-CTA-per-row-block kernel)
  const bool use_pair = (fwd_mode == 2 || (fwd_mode == 0 && n_rows >= 2048)) && !(shard != nullptr && shard->n_push > 0);
  for (int p = 0; p < n_pairs; ++p) {
    TCL_REQUIRE(zrow[p] && zcol[p], TCL_ERR_BAD_ARG, "ntxent_fwd: null operand (pair %d)", p);
    TCL_REQUIRE(aligned_to(zrow[p], 16), TCL_ERR_BAD_ALIGN, "ntxent_fwd: operands must be 16-byte aligned");
    if (int e = make_tmap_2d_16bit(&P.tm_row[p], zrow[p], n_rows, dim, z_row_stride, FW_BM, FW_BK)) return e;
    // pair kernel: each CTA loads its 64-row half of a column tile
    if (int e = make_tmap_2d_16bit(&P.tm_col[p], zcol[p], n_cols, dim, z_row_stride, use_pair ? 64 : FW_BN, FW_BK)) return e;
    P.z_row[p] = zrow[p];
  }
  P.z_row_stride = z_row_stride;
  P.idesc_pair = umma_idesc_f16(256, FW_BN, op_format);
  P.n_rows = (int)n_rows; P.n_cols = (int)n_cols; P.row_offset = (int)row_offset;
  P.num_kb = (int)(dim / 64);
  P.n_iblocks = (int)((n_rows + FW_BM - 1) / FW_BM);
  P.n_jtiles = (int)((n_cols + FW_BN - 1) / FW_BN);
  P.n_jsplit = fwd_split(n_pairs, P.n_iblocks, P.n_jtiles);
  P.c1 = c1;
  P.idesc = umma_idesc_f16(FW_BM, FW_BN, op_format);
  P.row_part = static_cast<float*>(workspace);
  P.col_part = P.row_part + static_cast<size_t>(n_pairs) * 2 * P.n_jsplit * n_rows;
  P.diag2 = diag2;
  if (shard != nullptr) {
    TCL_REQUIRE(n_rows % FW_BM == 0 && n_cols == n_rows * shard->world && row_offset == n_rows * shard->rank, TCL_ERR_BAD_SHAPE,
                "ntxent_fwd_sharded: rows per rank must be a multiple of 128 and n_cols = world * n_rows");
    P.sync = shard->sync;
    P.rank = shard->rank;
    P.world = shard->world;
    P.chunks_per_rank = static_cast<int>(n_rows / FW_BM);
    P.n_push = shard->n_push;
    P.dim = static_cast<int>(dim);
    for (int r = 0; r < shard->world && shard->n_push > 0; ++r) {
      P.z_peer[r] = static_cast<uint16_t*>(shard->z_base[r]);
      P.sync_peer[r] = static_cast<uint32_t*>(shard->sync_all[r]);
    }
    for (int m = 0; m < shard->n_push; ++m) P.push_off[m] = static_cast<int>(shard->push_off[m]);
  }

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (use_pair) {
    const int smem = (int)Fwd2Smem::total;
    if (int e = ensure_dyn_smem(ntxent_fwd_pair_kernel, smem)) return e;
    LaunchCfg L(dim3(2 * ((P.n_iblocks + 1) / 2), P.n_jsplit, n_pairs), dim3(F2_THREADS), smem, st, 2);
    ProfScope prof(TCL_K_NTXENT_FWD, st);
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, ntxent_fwd_pair_kernel, P));
  } else {
    const int smem = (int)FwdSmem::total(P.num_kb);
    if (int e = ensure_dyn_smem(ntxent_fwd_kernel, smem)) return e;
    dim3 grid(P.n_iblocks, P.n_jsplit, n_pairs);
    if (P.n_push > 0) {
      int dev = 0, n_sm = 0;
      TCL_CHECK_CUDA(cudaGetDevice(&dev));
      TCL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
      const int64_t total = static_cast<int64_t>(grid.x) * grid.y * grid.z;
      P.n_push_ctas = static_cast<int>(total < n_sm ? total : n_sm);
    }
    ProfScope prof(TCL_K_NTXENT_FWD, st);
    LaunchCfg L(grid, dim3(FW_THREADS + (P.n_push > 0 ? FW_PUSH_WARPS * 32 : 0)), smem, st);
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, ntxent_fwd_kernel, P));
  }
  TCL_CHECK_CUDA(cudaGetLastError());
  if (parts != nullptr) {  // the caller reduces the partials itself (fused into the finalise kernel)
    parts->row_part = P.row_part;
    parts->col_part = P.col_part;
    parts->n_row_slots = 2 * P.n_jsplit;
    parts->n_iblocks = P.n_iblocks;
    return TCL_OK;
  }
  const int nmax = (int)(n_rows > n_cols ? n_rows : n_cols);
  {
    ProfScope prof(TCL_K_FWD_REDUCE, st);
    LaunchCfg L(dim3((nmax + 255) / 256, n_pairs), dim3(256), 0, st);
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, fwd_reduce_kernel, (const float*)P.row_part, (const float*)P.col_part, row_sumexp,
                                      col_sumexp, n_pairs, P.n_rows, P.n_cols, 2 * P.n_jsplit, P.n_iblocks));
  }
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

extern "C" int tcl_ntxent_fwd(int n_pairs, const void* const* zrow, const void* const* zcol,
                              int64_t n_rows, int64_t n_cols, int64_t dim, int64_t z_row_stride, int64_t row_offset,
                              int op_format, float inv_tau, float* row_sumexp, float* col_sumexp,
                              float* diag2, void* workspace, size_t workspace_bytes, void* stream) {
  return ntxent_fwd_impl(n_pairs, zrow, zcol, n_rows, n_cols, dim, z_row_stride, row_offset, op_format, inv_tau,
                         row_sumexp, col_sumexp, diag2, workspace, workspace_bytes, stream, nullptr);
}

static int ntxent_finalize_impl(int n_pairs, int64_t n_rows, int64_t n_cols, int64_t row_offset,
                                float inv_tau, float alpha, float* row_sumexp, float* col_sumexp,
                                const float* diag2, float* lse2_row, float* lse2_col, float* loss_parts, float* loss,
                                void* stream, const FwdPartials* parts, bool want_total = false) {
  static std::atomic<unsigned> next_lane{0};
  const int total_lane = (want_total && loss) ? static_cast<int>(next_lane.fetch_add(1u) % FIN_LANES) : -1;
  TCL_REQUIRE(n_pairs >= 1 && n_pairs <= TCL_MAX_PAIRS, TCL_ERR_BAD_ARG, "finalize: n_pairs %d", n_pairs);
  TCL_REQUIRE(n_rows >= 1 && n_cols >= 1, TCL_ERR_BAD_SHAPE, "finalize: sizes");
  TCL_REQUIRE(row_sumexp && col_sumexp && diag2 && lse2_row && lse2_col && loss_parts, TCL_ERR_BAD_ARG, "finalize: null pointer");
  if (int e = require_sm100()) return e;
  const float c1 = inv_tau * 1.4426950408889634f;
  {
    ProfScope prof(TCL_K_FWD_FINALIZE, static_cast<cudaStream_t>(stream));
    const int nmax = (int)(n_rows > n_cols ? n_rows : n_cols);
    int threads = 32;
    while (threads < 1024 && threads * FIN_CTAS < nmax) threads <<= 1;
    LaunchCfg L(dim3(FIN_CTAS, n_pairs, 1), dim3(threads), 0, static_cast<cudaStream_t>(stream), FIN_CTAS);
    const float* rp = parts ? parts->row_part : nullptr;
    const float* cp = parts ? parts->col_part : nullptr;
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, fwd_finalize_kernel, (int)n_rows, (int)n_cols, (int)row_offset, c1, alpha,
                                      row_sumexp, col_sumexp, diag2, lse2_row, lse2_col, loss_parts, loss, rp, cp,
                                      parts ? parts->n_row_slots : 0, parts ? parts->n_iblocks : 0, total_lane));
  }
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

extern "C" int tcl_ntxent_finalize(int n_pairs, int64_t n_rows, int64_t n_cols, int64_t row_offset,
                                   float inv_tau, float alpha, const float* row_sumexp,
                                   const float* col_sumexp, const float* diag2, float* lse2_row,
                                   float* lse2_col, float* loss_parts, float* loss, void* stream) {
  return ntxent_finalize_impl(n_pairs, n_rows, n_cols, row_offset, inv_tau, alpha, const_cast<float*>(row_sumexp),
                              const_cast<float*>(col_sumexp), diag2, lse2_row, lse2_col, loss_parts, loss, stream, nullptr);
}

namespace tcl {
// Whole-loss forward on one GPU (ntxent_fused.cu): tile kernel, then ONE cluster kernel that reduces the tile
// kernel's partials and finalises (two launches instead of three) - for small batches.
int ntxent_fwd_finalize_fused(int n_pairs, const void* const* zrow, const void* const* zcol, int64_t batch, int64_t dim,
                              int op_format, float inv_tau, float alpha, float* row_sumexp, float* col_sumexp,
                              float* diag2, float* lse2_row, float* lse2_col, float* loss_parts, float* loss,
                              void* workspace, size_t workspace_bytes, void* stream, bool want_total) {
  // Small batches are launch-bound: the finalise cluster reduces the partials itself (B = 256: 15 -> 9 us).  At
  // large batches the partials are megabytes and the wide reduce kernel in front of the finalise is faster.
  const bool fuse = batch <= 2048;
  // ... and from there on ONE wide kernel reduces and finalises (TRICOLO_B200_FINALIZE=split keeps the two kernels)
  static const bool wide_ok = [] { const char* e = getenv("TRICOLO_B200_FINALIZE"); return !(e && strcmp(e, "split") == 0); }();
  const bool wide = !fuse && wide_ok && (batch + 63) / 64 <= RF_MAX_BLOCKS;
  FwdPartials parts;
  if (int e = ntxent_fwd_impl(n_pairs, zrow, zcol, batch, batch, dim, 0, 0, op_format, inv_tau, row_sumexp, col_sumexp,
                              diag2, workspace, workspace_bytes, stream, (fuse || wide) ? &parts : nullptr))
    return e;
  if (wide) {
    static std::atomic<unsigned> next_lane{0};
    const int lane_id = static_cast<int>(next_lane.fetch_add(1u) % FIN_LANES);
    TCL_REQUIRE(row_sumexp && col_sumexp && diag2 && lse2_row && lse2_col && loss_parts, TCL_ERR_BAD_ARG, "finalize: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ProfScope prof(TCL_K_FWD_FINALIZE, st);
    LaunchCfg L(dim3(static_cast<unsigned>((batch + 63) / 64), n_pairs), dim3(256), 0, st);
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, fwd_reduce_finalize_kernel, (int)batch, inv_tau * 1.4426950408889634f, alpha,
                                      row_sumexp, col_sumexp, (const float*)diag2, lse2_row, lse2_col, loss_parts, loss,
                                      parts.row_part, parts.col_part, parts.n_row_slots, parts.n_iblocks, lane_id,
                                      (want_total && loss) ? 1 : 0));
    TCL_CHECK_CUDA(cudaGetLastError());
    return TCL_OK;
  }
  return ntxent_finalize_impl(n_pairs, batch, batch, 0, inv_tau, alpha, row_sumexp, col_sumexp, diag2, lse2_row, lse2_col,
                              loss_parts, loss, stream, fuse ? &parts : nullptr, want_total);
}
}  // namespace tcl

// ---------------------------------------------------------------------------------------------------------------
// sharded forward: tile kernel gated on the arrival flags + statistics push, then the waiting finalise
// ---------------------------------------------------------------------------------------------------------------
extern "C" size_t tcl_shard_stats_bytes(int n_pairs, int64_t b_loc, int world) {
  if (n_pairs < 1 || n_pairs > TCL_MAX_PAIRS || b_loc < 1 || world < 1 || world > TCL_MAX_PEERS) return 0;
  return static_cast<size_t>(world) * tcl::shard_stats_slot_floats(n_pairs, b_loc, b_loc * world) * sizeof(float);
}

extern "C" int tcl_ntxent_fwd_sharded(int n_pairs, const void* const* zrow, const void* const* zcol, int64_t b_loc,
                                      int64_t b_glob, int64_t dim, int64_t z_row_stride, int rank, int world,
                                      int op_format, float inv_tau, float* diag2, void* workspace,
                                      size_t workspace_bytes, void* const* stats_ptrs, void* const* sync_ptrs,
                                      void* const* z_base_ptrs, int n_push, const int64_t* push_offsets, void* stream) {
  TCL_REQUIRE(world >= 1 && world <= TCL_MAX_PEERS && rank >= 0 && rank < world, TCL_ERR_BAD_ARG, "fwd_sharded: rank %d of %d", rank, world);
  TCL_REQUIRE(stats_ptrs && diag2, TCL_ERR_BAD_ARG, "fwd_sharded: null pointer");
  TCL_REQUIRE(b_loc >= 1 && b_loc <= 128 * ShardSync::kMaxChunks && b_glob == b_loc * world, TCL_ERR_BAD_SHAPE, "fwd_sharded: sizes");
  const bool flags = sync_ptrs != nullptr;  // else: the caller brackets the calls with cross-rank barriers
  TCL_REQUIRE(flags || n_push == 0, TCL_ERR_BAD_ARG, "fwd_sharded: the fused all-gather needs the sync pads");
  for (int r = 0; r < world; ++r)
    TCL_REQUIRE(stats_ptrs[r] && aligned_to(stats_ptrs[r], 16) && (!flags || (sync_ptrs[r] && aligned_to(sync_ptrs[r], 16))),
                TCL_ERR_BAD_ALIGN, "fwd_sharded: buffers of rank %d", r);
  TCL_REQUIRE(n_push >= 0 && n_push <= TCL_MAX_TENSORS && (n_push == 0 || (z_base_ptrs && push_offsets)), TCL_ERR_BAD_ARG,
              "fwd_sharded: n_push %d", n_push);
  for (int r = 0; r < world && n_push > 0; ++r)
    TCL_REQUIRE(z_base_ptrs[r] && aligned_to(z_base_ptrs[r], 16), TCL_ERR_BAD_ALIGN, "fwd_sharded: gathered buffer of rank %d", r);
  for (int m = 0; m < n_push; ++m)
    TCL_REQUIRE(push_offsets[m] >= 0 && push_offsets[m] % 8 == 0, TCL_ERR_BAD_ALIGN, "fwd_sharded: push offset %d", m);
  FwdShard sh{flags ? static_cast<const uint32_t*>(sync_ptrs[rank]) : nullptr, rank, world, z_base_ptrs, sync_ptrs, n_push,
              push_offsets};
  FwdPartials parts;
  if (int e = ntxent_fwd_impl(n_pairs, zrow, zcol, b_loc, b_glob, dim, z_row_stride, b_loc * rank, op_format, inv_tau, nullptr,
                              nullptr, diag2, workspace, workspace_bytes, stream, &parts, flags ? &sh : nullptr))
    return e;
  StatsPushParams S;
  memset(&S, 0, sizeof(S));
  for (int r = 0; r < world; ++r) {
    S.dst[r] = static_cast<float*>(stats_ptrs[r]);
    S.sync[r] = flags ? static_cast<uint32_t*>(sync_ptrs[r]) : nullptr;
  }
  S.local_only = flags ? 0 : 1;
  S.row_part = parts.row_part;
  S.col_part = parts.col_part;
  S.diag2 = diag2;
  S.slot_floats = shard_stats_slot_floats(n_pairs, b_loc, b_glob);
  S.n_pairs = n_pairs;
  S.n_rows = static_cast<int>(b_loc);
  S.n_cols = static_cast<int>(b_glob);
  S.n_row_slots = parts.n_row_slots;
  S.n_iblocks = parts.n_iblocks;
  S.rank = rank;
  S.world = world;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ProfScope prof(TCL_K_FWD_REDUCE, st);
    LaunchCfg L(dim3(static_cast<unsigned>((b_glob + 255) / 256), n_pairs), dim3(256), 0, st);
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, fwd_reduce_push_kernel, S));
  }
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

extern "C" int tcl_ntxent_finalize_sharded(int n_pairs, int64_t b_loc, int64_t b_glob, int rank, int world, float inv_tau,
                                           float alpha, void* const* stats_ptrs, const void* sync_own, float* lse2_row,
                                           float* lse2_col, float* loss, void* stream) {
  TCL_REQUIRE(n_pairs >= 1 && n_pairs <= TCL_MAX_PAIRS && world >= 1 && world <= TCL_MAX_PEERS && b_glob == b_loc * world &&
                  rank >= 0 && rank < world, TCL_ERR_BAD_ARG, "finalize_sharded: sizes");
  TCL_REQUIRE(stats_ptrs && lse2_row && lse2_col && loss, TCL_ERR_BAD_ARG, "finalize_sharded: null pointer");
  // flags: every rank pushed its statistics into slot s of THIS rank's buffer; pull: slot s is read from rank s's buffer
  StatsSrc src;
  memset(&src, 0, sizeof(src));
  const int64_t slot_floats = shard_stats_slot_floats(n_pairs, b_loc, b_glob);
  for (int s2 = 0; s2 < world; ++s2) {
    TCL_REQUIRE(stats_ptrs[s2] != nullptr, TCL_ERR_BAD_ARG, "finalize_sharded: statistics buffer %d", s2);
    src.slot[s2] = static_cast<const float*>(stats_ptrs[sync_own ? rank : s2]) + s2 * slot_floats;
  }
  if (int e = require_sm100()) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ProfScope prof(TCL_K_FWD_FINALIZE, st);
  int threads = 32;
  while (threads < 1024 && threads * FIN_CTAS < b_glob) threads <<= 1;
  LaunchCfg L(dim3(FIN_CTAS, n_pairs, 1), dim3(threads), 0, st, FIN_CTAS);
  TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, fwd_finalize_sharded_kernel, (int)b_loc, (int)b_glob, n_pairs, world,
                                    inv_tau * 1.4426950408889634f, alpha, src, static_cast<const uint32_t*>(sync_own), lse2_row,
                                    lse2_col, loss));
  return TCL_OK;
}
