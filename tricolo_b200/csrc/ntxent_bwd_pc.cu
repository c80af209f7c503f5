// K3, producer/consumer form (dim > 256): the backward of tricolo/loss/nt_xent.py:55-74 with the logit
// recompute and the gradient GEMM on DIFFERENT SMs of a 2-CTA cluster.
//
// Why: the fp32 gradient accumulator of 128 self rows x dim 512 fills all 512 TMEM columns of an SM, so a single
// CTA cannot hold it next to the logit buffers (ntxent_bwd.cu splits dim and recomputes the logits twice; the
// pair kernel exchanges G both ways on the per-tile critical path).  Here the two roles get one SM each:
//   producer CTA (cluster rank 0)   S = Zself[128 rows] · Zother[tile]^T into two 128-column TMEM buffers.  The
//                                   self block is the A operand and stays resident IN TMEM (128 lanes x dim/2
//                                   columns, written once by tcgen05.st), so shared memory holds only the
//                                   operand ring and the G staging slots.  The epilogue warps turn S into the
//                                   16-bit gradient weights G' (same formula as ntxent_bwd.cu), store them in
//                                   the K-major swizzled operand layout into a local staging slot, and one
//                                   thread per 16 KB K-block pushes it into the consumer's shared memory with
//                                   a bulk DSMEM copy (cp.async.bulk.shared::cluster) that completes on the
//                                   consumer's mbarrier.  (Direct st.shared::cluster stores were measured at
//                                   ~6 B/clk for this access pattern: 5600 cycles per 32 KB tile.)
//   consumer CTA (cluster rank 1)   acc[128 rows x dim] += G'[128 x 128] · Zother[tile]  with the whole 512-column
//                                   TMEM as accumulator: M=128, N=256 MMAs, A = a G slot, B = the other operand
//                                   in MN-major form (TMA boxes {64 dim, 64 rows} of the row-major tensor, so
//                                   no transposed copy is needed) streamed by its own TMA ring.
// Both SMs run 2·128·128·dim flop per tile: the work is balanced, nothing is recomputed twice (8 B^2 D executed
// per pair, the count of a recompute backward) and the 32 KB/tile G hand-over is one-way and three slots deep,
// i.e. off the critical path.  Hand-shakes: producer -> consumer MMA thread: the copies' complete_tx on g_full[slot]
// (armed with arrive.expect_tx by the consumer itself: a remote release-arrive cost ~1800 cycles per tile); consumer -> producer: tcgen05.commit multicast onto the producer's g_empty[slot]
// once the MMAs that read the slot have completed (which also frees the staging slot of the same index).
//
// Persistent form: the launch is one cluster per SM pair.  The flat tile sequence (job-major, then 128-row unit,
// then (segment, column tile)) is cut into gridDim.x/2 equal contiguous ranges; a cluster walks its range unit piece
// by unit piece.  At the end of a piece the consumer drains its accumulator through shared memory with TMA stores
// (full 128-byte lines; the old per-thread row stores took ~22k cycles per 256 KB) into partial slot
// `cluster - cluster_of(first tile of the unit)`; the producer already works on the next piece meanwhile (its
// hand-over is three G slots deep).  No wave quantisation, 3-4 drains per SM pair instead of 8 at B = 8192, and the
// normalise backward reads 1-2 partials per row instead of 3.
#include <stdlib.h>

#include "ntxent_bwd.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

#ifndef TCL_PC_PSTAGES
#define TCL_PC_PSTAGES 4
#endif
// Timing experiments (never in the product build): TCL_PC_EXP 1 = the producer's ring is never loaded or waited for
// (logit MMAs on stale shared memory), 2 = the same for the consumer's ring as well.
#ifndef TCL_PC_EXP
#define TCL_PC_EXP 0
#endif
static constexpr int PC_PSTAGES = TCL_PC_PSTAGES;  // producer ring: slots of two 16 KB K-blocks of the other operand
static constexpr int PC_CSTAGES = 3;     // consumer ring: slots of 4 boxes {64 dim x 64 other rows} (N = 256, K = 64)
static constexpr int PC_SLOT = 32768;
static constexpr int PC_GSLOTS = 3;      // G tiles in flight (128 rows x 128 K, two K-blocks of 16 KB)
static constexpr int PC_SBUFS = 2;       // logit buffers in the producer's TMEM (columns 0..255; the self block: 256..)
static constexpr int PC_XCOL = 256;      // first TMEM column of the resident self block
static constexpr int PC_EPI_WARPS = 16;  // two groups of 8: group g handles the tiles t = g (mod 2) (logit buffer g)
static constexpr int PC_DRAIN_WARPS = 8; // consumer warps that read the accumulator out (TMEM read rate is the limit)
static constexpr int PC_DRAIN_BYTES = 4096;  // per drain warp: 32 rows x 32 fp32 columns, 128-byte swizzle
static constexpr int PC_THREADS = 64 + PC_EPI_WARPS * 32;

struct PcSmem {
  // producer: [G staging 3 x 32 KB][ring 4 x 32 KB];  consumer: [G 3 x 32 KB][ring 3 x 32 KB][drain staging 8 x 4 KB]
  static constexpr uint32_t p_stage_off = 0;
  static constexpr uint32_t p_ring_off = PC_GSLOTS * PC_SLOT;
  static constexpr uint32_t c_g_off = 0;
  static constexpr uint32_t c_ring_off = PC_GSLOTS * PC_SLOT;
  static constexpr uint32_t c_drain_off = (PC_GSLOTS + PC_CSTAGES) * PC_SLOT;
  static constexpr uint32_t buf_bytes() {
    return (PC_GSLOTS + PC_PSTAGES) * PC_SLOT > c_drain_off + PC_DRAIN_WARPS * PC_DRAIN_BYTES
               ? (PC_GSLOTS + PC_PSTAGES) * PC_SLOT
               : c_drain_off + PC_DRAIN_WARPS * PC_DRAIN_BYTES;
  }
  static constexpr uint32_t bar_off() { return buf_bytes(); }   // same offset in both CTAs
  static constexpr uint32_t bj_off() { return bar_off() + 512; }  // [2 groups][128] floats
  static constexpr uint32_t total() { return bj_off() + 1024 + 1024; }
};
static_assert(PcSmem::total() <= 232448, "producer/consumer backward: shared memory budget");

// Optional wait-time accounting of the first cluster (make trace; profiles/pc_trace.py): cycles per role and wait.
// slots: 0 P-tma p_empty | 1 P-mma s_empty, 2 P-mma p_full, 3 P-mma total, 12 P-mma x_full | 4 P-epi s_full,
// 6 P-epi math, 7 P-epi g_empty, 8 P-epi stores, 9 P-epi fence+arrive, 10 P-epi total, 11 P-epi bar.sync,
// 13 P-epi piece prologue | 16 C-tma c_empty | 17 C-mma g_full, 18 C-mma c_full, 19 C-mma total, 21 C-mma acc_empty |
// 20 C-epi read-out, 22 C-epi acc_full | 30 pieces, 31 tiles
__device__ unsigned long long g_pc_trace[32];
#ifdef TCL_PAIR_TRACE
#define PT_DECL unsigned long long pt_t0 = 0; const bool pt_on = blockIdx.x < 2 && (threadIdx.x & 31) == 0 && ((threadIdx.x >> 5) <= 2); (void)pt_t0;
#define PT_BEGIN() do { if (pt_on) pt_t0 = clock64(); } while (0)
#define PT_END(slot) do { if (pt_on) { const unsigned long long pt_t1 = clock64(); atomicAdd(&g_pc_trace[slot], pt_t1 - pt_t0); pt_t0 = pt_t1; } } while (0)
#else
#define PT_DECL
#define PT_BEGIN() do {} while (0)
#define PT_END(slot) do {} while (0)
#endif

// One piece = a contiguous tile range [ta, tb) of one 128-row unit (job, ib); its accumulator goes to partial `slot`.
struct PcPiece {
  int job, ib, ta, tb, slot;
};
// Every thread of the cluster walks the same piece sequence.
struct PcWalk {
  int64_t cursor, end;
  int c, n;
  __device__ explicit PcWalk(const BwdParams& P) {
    n = static_cast<int>(gridDim.x >> 1);
    c = static_cast<int>(blockIdx.x >> 1);
    const int64_t total = P.job_tile_base[TCL_MAX_TENSORS];
    cursor = pc_range_lo(total, c, n);
    end = pc_range_lo(total, c + 1, n);
  }
  __device__ bool next(const BwdParams& P, PcPiece& pc) {
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

template <int kOp>
__global__ void __launch_bounds__(PC_THREADS, 1) ntxent_bwd_pc_kernel(const __grid_constant__ BwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int num_kb = P.num_kb;
  const uint32_t bars = base + PcSmem::bar_off();
  auto p_full = [&](int s) { return bars + 8u * (40 + s); };
  auto p_empty = [&](int s) { return bars + 8u * (32 + s); };
  const uint32_t x_full_bar = bars + 8u * 6;  // the self block is in TMEM (one arrive per epilogue warp and piece)
  auto s_full = [&](int b) { return bars + 8u * (7 + b); };
  auto s_empty = [&](int b) { return bars + 8u * (11 + b); };
  auto g_empty = [&](int g) { return bars + 8u * (15 + g); };  // lives in the PRODUCER's shared memory
  auto c_full = [&](int s) { return bars + 8u * (18 + s); };
  auto c_empty = [&](int s) { return bars + 8u * (22 + s); };
  auto g_full = [&](int g) { return bars + 8u * (26 + g); };   // lives in the CONSUMER's shared memory
  const uint32_t acc_full_bar = bars + 8u * 29;
  const uint32_t tmem_slot = bars + 8u * 30;
  const uint32_t acc_empty_bar = bars + 8u * 31;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + PcSmem::bar_off() + 8u * 30);
  float* bj_all = reinterpret_cast<float*>(base_ptr + PcSmem::bj_off());  // [2 groups][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool is_producer = crank == 0;
  PT_DECL
  const int n_chunk = (P.dim + 255) / 256;  // 256-column accumulator chunks of the consumer
  const int x_slots = (num_kb + 1) / 2;     // producer ring slots per tile, and per staged self block

  if (warp == 0 && elect_one()) {
    for (int j = 0; j < TCL_MAX_TENSORS; ++j) {
      if (P.unit_tiles[j] == 0) continue;
      tma_prefetch_desc(&P.job[j].tm_self);
      tma_prefetch_desc(&P.tm_gpart[j]);
      for (int s = 0; s < P.job[j].n_seg; ++s) {
        tma_prefetch_desc(&P.job[j].seg[s].tm_other);
        tma_prefetch_desc(&P.job[j].seg[s].tm_other_t);
      }
    }
    for (int s = 0; s < PC_PSTAGES; ++s) {
      mbar_init(p_full(s), 1);
      mbar_init(p_empty(s), 1);
    }
    mbar_init(x_full_bar, PC_EPI_WARPS);
    for (int b = 0; b < PC_SBUFS; ++b) {
      mbar_init(s_full(b), 1);
      mbar_init(s_empty(b), PC_EPI_WARPS / 2);  // the eight warps of the group that owns the buffer
    }
    for (int g = 0; g < PC_GSLOTS; ++g) {
      mbar_init(g_empty(g), 1);           // one multicast commit from the consumer's MMA thread
      mbar_init(g_full(g), 1);             // armed (arrive.expect_tx 32 KB) by the consumer's MMA thread; the bytes
                                           // come from the producer's two bulk copies per tile
    }
    for (int s = 0; s < PC_CSTAGES; ++s) {
      mbar_init(c_full(s), 1);
      mbar_init(c_empty(s), 1);
    }
    mbar_init(acc_full_bar, 1);
    mbar_init(acc_empty_bar, PC_DRAIN_WARPS);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  griddep_launch();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  griddep_wait();

  PcWalk walk(P);
  PcPiece pc;

  if (is_producer) {
    const uint32_t stage = base + PcSmem::p_stage_off;
    const uint32_t ring = base + PcSmem::p_ring_off;
    const uint32_t tmem_x = tmem + PC_XCOL;
    if (warp == 0) {
      // -------------------------------------------------------------- producer: TMA warp
      if (TCL_PC_EXP < 1 && elect_one()) {
        uint32_t it = 0;
        while (walk.next(P, pc)) {
          const BwdJobDev& J = P.job[pc.job];
          // the piece's self rows travel through the ring first (staging for the copy into TMEM): they are fetched
          // while the previous piece is still being computed
          for (int kb = 0; kb < num_kb; kb += 2, ++it) {
            const int nk = kb + 1 < num_kb ? 2 : 1;
            const int s = it % PC_PSTAGES;
            mbar_wait(p_empty(s), ((it / PC_PSTAGES) & 1) ^ 1);
            mbar_arrive_expect_tx(p_full(s), nk * BW_KB_BYTES);
            for (int k2 = 0; k2 < nk; ++k2)
              tma_load_2d(ring + s * PC_SLOT + k2 * BW_KB_BYTES, &J.tm_self, p_full(s), (kb + k2) * BW_BK, pc.ib * BW_BM);
          }
          for (int t = pc.ta; t < pc.tb; ++t) {
            const BwdSegDev& sg = J.seg[t / P.n_jtiles];
            const int j0 = (t % P.n_jtiles) * BW_BN;
            for (int kb = 0; kb < num_kb; kb += 2, ++it) {
              const int nk = kb + 1 < num_kb ? 2 : 1;
              const int s = it % PC_PSTAGES;
              PT_BEGIN();
              mbar_wait(p_empty(s), ((it / PC_PSTAGES) & 1) ^ 1);
              PT_END(0);
              mbar_arrive_expect_tx(p_full(s), nk * BW_KB_BYTES);
              for (int k2 = 0; k2 < nk; ++k2)
                tma_load_2d(ring + s * PC_SLOT + k2 * BW_KB_BYTES, &sg.tm_other, p_full(s), (kb + k2) * BW_BK, j0);
            }
          }
        }
      }
    } else if (warp == 1) {
      // -------------------------------------------------------------- producer: logit MMAs (A from TMEM)
      if (elect_one()) {
        uint32_t it = 0, tg = 0, piece = 0;
#ifdef TCL_PAIR_TRACE
        const unsigned long long pt_m0 = clock64();
#endif
        while (walk.next(P, pc)) {
          PT_BEGIN();
          mbar_wait(x_full_bar, piece & 1);  // this piece's self rows are in TMEM
          PT_END(12);
          tc_fence_after();
          it += static_cast<uint32_t>(x_slots);  // ring slots that staged the self rows (released by the epilogue warps)
          for (int t = pc.ta; t < pc.tb; ++t, ++tg) {
            const int b = tg % PC_SBUFS;
            PT_BEGIN();
            mbar_wait(s_empty(b), ((tg / PC_SBUFS) & 1) ^ 1);
            PT_END(1);
            tc_fence_after();
            for (int kb = 0; kb < num_kb; kb += 2, ++it) {
              const int nk = kb + 1 < num_kb ? 2 : 1;
              const int s = it % PC_PSTAGES;
              PT_BEGIN();
              if (TCL_PC_EXP < 1) mbar_wait(p_full(s), (it / PC_PSTAGES) & 1);
              PT_END(2);
              tc_fence_after();
              for (int k2 = 0; k2 < nk; ++k2) {
                const uint32_t ax = tmem_x + (kb + k2) * (BW_BK / 2);  // 32 columns per K-block of 64
                const uint64_t bd = umma_desc_k_sw128(ring + s * PC_SLOT + k2 * BW_KB_BYTES);
#pragma unroll
                for (int kk = 0; kk < BW_BK / 16; ++kk)
                  tc_mma_f16_ts(tmem + b * BW_BN, ax + 8 * kk, bd + 2 * kk, P.idesc, (kb | k2 | kk) != 0);
              }
              if (TCL_PC_EXP < 1) tc_commit(p_empty(s));
            }
            tc_commit(s_full(b));
          }
          ++piece;
        }
#ifdef TCL_PAIR_TRACE
        if (pt_on) { atomicAdd(&g_pc_trace[3], clock64() - pt_m0); atomicAdd(&g_pc_trace[31], (unsigned long long)tg); atomicAdd(&g_pc_trace[30], (unsigned long long)piece); }
#endif
      }
    } else {
      // -------------------------------------------------------------- producer: G epilogue (2 groups x 8 warps)
      // Group gi owns logit buffer gi and the tiles whose running index is gi (mod 2): each group has two tile
      // periods for its serial chain (wait, TMEM load, math, staging stores, hand-over).
      const int ew = warp - 2;
      const int gi = ew >> 3;          // group
      const int q = warp & 3;          // TMEM lane quarter
      const int ch = (ew >> 2) & 1;    // column half of the logit tile == K-block of the G operand
      const int r = q * 32 + lane;     // tile-local row == TMEM lane
      const int gt = (ew & 7) * 32 + lane;  // thread index inside the group, 0..255
      const int bar_grp = 1 + gi;           // named barriers: 1,2 = group; 3..6 = (group, K-block); 7 = all 16 warps
      const int bar_kb = 3 + 2 * gi + ch;
      float* bj = bj_all + gi * 128;
      const uint32_t g_peer = map_to_peer(base + PcSmem::c_g_off, 1u);  // G slots in the consumer's shared memory
      const uint32_t g_full_peer0 = map_to_peer(g_full(0), 1u);
      const uint32_t row_off = static_cast<uint32_t>(ch * BW_KB_BYTES + r * 128);
      uint8_t* stage_ptr = base_ptr + PcSmem::p_stage_off;
      const uint32_t s_addr = tmem_addr(tmem + gi * BW_BN, q * 32, ch * 64);
      uint32_t tg0 = 0, s_par = 0, piece = 0, it_ring = 0;
#ifdef TCL_PAIR_TRACE
      const unsigned long long pt_e0 = clock64();
#endif
      while (walk.next(P, pc)) {
        const BwdJobDev& J = P.job[pc.job];
        const int i0 = pc.ib * BW_BM;
        const int grow = i0 + r;
        PT_BEGIN();
        // every logit MMA of the previous piece has completed (each group has seen the s_full of its last tile)
        // before the self block in TMEM is overwritten
        if (piece > 0) asm volatile("bar.sync 7, 512;" ::: "memory");
        {
          // self block ring slots -> registers -> TMEM: row = lane, 16-bit element k -> column k/2; this thread: the two
          // K-blocks (= one ring slot) c0, c0+1 of its row.  16-byte chunk e of a row sits at chunk position e ^ (r & 7).
          const int c0 = (gi * 2 + ch) * 2;
          if (c0 < num_kb) {
            const uint32_t itx = it_ring + static_cast<uint32_t>(c0 >> 1);
            const int sx = itx % PC_PSTAGES;
            mbar_wait(p_full(sx), (itx / PC_PSTAGES) & 1);
            const uint8_t* slot = base_ptr + PcSmem::p_ring_off + sx * PC_SLOT;
#pragma unroll 1
            for (int c32 = c0; c32 < c0 + 2; ++c32) {  // 32 TMEM columns = 64 elements = one K-block
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
          // everyone has read its staging slot: hand the slots back to the TMA warp, then publish the TMEM block
          asm volatile("bar.sync 8, 512;" ::: "memory");
          if (ew == 0 && lane == 0)
            for (int x = 0; x < x_slots; ++x) mbar_arrive(p_empty((it_ring + x) % PC_PSTAGES));
          if (lane == 0) mbar_arrive(x_full_bar);
          it_ring += static_cast<uint32_t>(x_slots + (pc.tb - pc.ta) * x_slots);  // ring slots of this piece
        }
        PT_END(13);
        float gs[2] = {0.f, 0.f};
        float gmax = 0.f;
        for (int s = 0; s < J.n_seg; ++s) {
          gs[s] = J.seg[s].grad_scale ? *J.seg[s].grad_scale : 1.f;
          gmax = fmaxf(gmax, fabsf(gs[s]));
        }
        const float inv_gmax = gmax > 0.f ? kGScale / gmax : 0.f;  // G in [-kGScale, kGScale]: see ntxent_bwd.h
        if (pc.ib == 0 && pc.ta == 0 && ew == 0 && lane == 0) *J.scale_out = gmax * P.out_scale * (1.f / kGScale);

        // tiles of this group in the piece: running index tg0 + (t - ta) == gi (mod 2)
        int t = pc.ta + static_cast<int>((static_cast<uint32_t>(gi) - tg0) & 1u);
        int si = t / P.n_jtiles, jt = t % P.n_jtiles;
        // per-column factors 2^(c1 - lse_other_j): the global load is issued one iteration ahead and consumed
        // (ex2 + shared-memory store) at the top of the next one, so its latency is off the loop's path
        auto load_lse = [&](int seg, int jtile, bool valid) -> float {
          if (gt >= 128 || !valid) return 1e30f;
          const int j = jtile * BW_BN + gt;
          return j < P.n_other ? J.seg[seg].lse2_other[j] : 1e30f;  // 2^(c1 - 1e30) = 0
        };
        float lse_col = load_lse(si, jt, t < pc.tb);
        int cur_seg = -1;
        float lse_i = 0.f, ws = 0.f, wo_i = 0.f, rr = 0.f;
        for (; t < pc.tb; t += 2) {
          const uint32_t tg = tg0 + static_cast<uint32_t>(t - pc.ta);
          const int g = tg % PC_GSLOTS;
          if (si != cur_seg) {
            const BwdSegDev& sg = J.seg[si];
            cur_seg = si;
            rr = gs[si] * inv_gmax;
            lse_i = grow < P.n_self ? sg.lse2_self[grow] : 0.f;
            ws = rr * sg.w_self;
            wo_i = rr * sg.w_other * ex2_approx(lse_i - P.c1);
          }
          const int j0 = jt * BW_BN;
          if (gt < 128) bj[gt] = ex2_approx(P.c1 - lse_col);
          // next tile of this group: prefetch its column constants
          int si_n = si, jt_n = jt + 2;
          if (jt_n >= P.n_jtiles) { jt_n -= P.n_jtiles; ++si_n; }
          lse_col = load_lse(si_n, jt_n, t + 2 < pc.tb);
          PT_BEGIN();
          asm volatile("bar.sync %0, 256;" ::"r"(bar_grp) : "memory");  // bj visible to the group
          PT_END(11);
          const int dcol = P.self_offset + grow - j0 - ch * 64;  // column of the positive inside this thread's 64
          const bool has_diag = (P.self_offset + i0 < j0 + BW_BN) && (P.self_offset + i0 + BW_BM > j0);  // CTA-uniform

          mbar_wait(s_full(gi), s_par);
          s_par ^= 1;
          PT_END(4);
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
              if (lane == 0) mbar_arrive(s_empty(gi));  // logits are in registers: the TMEM buffer can be refilled
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
          PT_END(6);
          // staging slot g (and the consumer's slot g) is free once the consumer's MMAs of tile tg - PC_GSLOTS completed
          mbar_wait(g_empty(g), ((tg / PC_GSLOTS) & 1u) ^ 1u);
          PT_END(7);
          // K-major, 128-byte-swizzled operand tile: row r, 16-byte chunk c16 -> c16 ^ (r & 7)
          uint8_t* gk = stage_ptr + g * PC_SLOT + row_off;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const int c16 = h * 4 + c4;
              *reinterpret_cast<uint4*>(gk + ((c16 ^ (r & 7)) << 4)) =
                  make_uint4(pk[h][4 * c4], pk[h][4 * c4 + 1], pk[h][4 * c4 + 2], pk[h][4 * c4 + 3]);
            }
          }
          fence_proxy_async_smem();  // generic-proxy stores -> async proxy (the bulk copy reads them)
          PT_END(8);
          // the four warps of this K-block are done (also: everyone has read bj): one thread ships 16 KB to the consumer
          asm volatile("bar.sync %0, 128;" ::"r"(bar_kb) : "memory");
          if (q == 0 && lane == 0) {
            const uint32_t off = static_cast<uint32_t>(g * PC_SLOT + ch * BW_KB_BYTES);
            bulk_copy_to_cluster(g_peer + off, stage + off, BW_KB_BYTES, g_full_peer0 + 8u * g);
          }
          PT_END(9);
          // both K-block halves of the group have passed their barrier before bj is rewritten: the writers (gt < 128,
          // i.e. K-block half 0 of the group) only need the readers of half 1 -> one more group barrier
          asm volatile("bar.sync %0, 256;" ::"r"(bar_grp) : "memory");
          jt += 2;
          if (jt >= P.n_jtiles) { jt -= P.n_jtiles; ++si; }
        }
        tg0 += static_cast<uint32_t>(pc.tb - pc.ta);
        ++piece;
      }
#ifdef TCL_PAIR_TRACE
      if (pt_on) atomicAdd(&g_pc_trace[10], clock64() - pt_e0);
#endif
    }
  } else {
    const uint32_t g_smem = base + PcSmem::c_g_off;
    const uint32_t ring = base + PcSmem::c_ring_off;
    if (warp == 0) {
      // -------------------------------------------------------------- consumer: TMA warp
      if (TCL_PC_EXP < 2 && elect_one()) {
        uint32_t it = 0;
        while (walk.next(P, pc)) {
          const BwdJobDev& J = P.job[pc.job];
          for (int t = pc.ta; t < pc.tb; ++t) {
            const BwdSegDev& sg = J.seg[t / P.n_jtiles];
            const int j0 = (t % P.n_jtiles) * BW_BN;
            for (int kb2 = 0; kb2 < 2; ++kb2)
              for (int c = 0; c < n_chunk; ++c, ++it) {
                const int s = it % PC_CSTAGES;
                PT_BEGIN();
                mbar_wait(c_empty(s), ((it / PC_CSTAGES) & 1) ^ 1);
                PT_END(16);
                mbar_arrive_expect_tx(c_full(s), PC_SLOT);
                // four boxes {64 dim columns, 64 other rows}: the B operand in MN-major form (N = dim, K = other
                // rows), read straight from the row-major operand: no transposed copy
                for (int a = 0; a < 4; ++a)
                  tma_load_2d(ring + s * PC_SLOT + a * 8192, &sg.tm_other_t, c_full(s), c * 256 + a * 64,
                              j0 + kb2 * BW_BK);
              }
          }
        }
      }
    } else if (warp == 1) {
      // -------------------------------------------------------------- consumer: gradient MMAs
      if (elect_one()) {
        uint32_t it = 0, tg = 0, piece = 0;
        for (int g = 0; g < PC_GSLOTS; ++g) mbar_arrive_expect_tx(g_full(g), PC_SLOT);  // arm the first round
#ifdef TCL_PAIR_TRACE
        const unsigned long long pt_m0 = clock64();
#endif
        while (walk.next(P, pc)) {
          PT_BEGIN();
          mbar_wait(acc_empty_bar, (piece & 1) ^ 1);  // the previous piece's accumulator has been read out
          PT_END(21);
          tc_fence_after();
          for (int t = pc.ta; t < pc.tb; ++t, ++tg) {
            const int g = tg % PC_GSLOTS;
            PT_BEGIN();
            mbar_wait(g_full(g), (tg / PC_GSLOTS) & 1);  // both bulk copies of the producer have landed
            PT_END(17);
            // arm the next round of this slot: its bytes cannot be sent before the g_empty commit below
            mbar_arrive_expect_tx(g_full(g), PC_SLOT);
            tc_fence_after();
            for (int kb2 = 0; kb2 < 2; ++kb2)
              for (int c = 0; c < n_chunk; ++c, ++it) {
                const int s = it % PC_CSTAGES;
                PT_BEGIN();
                if (TCL_PC_EXP < 2) mbar_wait(c_full(s), (it / PC_CSTAGES) & 1);
                PT_END(18);
                tc_fence_after();
                const uint64_t ad = umma_desc_k_sw128(g_smem + g * PC_SLOT + kb2 * BW_KB_BYTES);
                const uint64_t bd = umma_desc_mn_sw128(ring + s * PC_SLOT, 8192);
#pragma unroll
                for (int kk = 0; kk < BW_BK / 16; ++kk)
                  tc_mma_f16(tmem + c * 256, ad + 2 * kk, bd + 128 * kk, P.idesc_n256, ((t - pc.ta) | kb2 | kk) != 0);
                if (TCL_PC_EXP < 2) tc_commit(c_empty(s));
              }
            tc_commit_multicast(g_empty(g), 0x1);  // slot g consumed: tell the producer (cluster rank 0)
          }
          tc_commit(acc_full_bar);
          ++piece;
        }
#ifdef TCL_PAIR_TRACE
        if (pt_on) atomicAdd(&g_pc_trace[19], clock64() - pt_m0);
#endif
      }
    } else if (warp < 2 + PC_DRAIN_WARPS) {
      // -------------------------------------------------------------- consumer: accumulator read-out (8 warps)
      // warp = (lane quarter q, column half): 32 rows x 256 columns in chunks of 32 columns: TMEM -> registers ->
      // swizzled 4 KB staging tile -> TMA store {32 cols, 32 rows} into the partial of this piece
      const int q = warp & 3;
      const int half = (warp - 2) >> 2;
      const uint32_t stg = base + PcSmem::c_drain_off + static_cast<uint32_t>(warp - 2) * PC_DRAIN_BYTES;
      uint8_t* stg_ptr = base_ptr + PcSmem::c_drain_off + (warp - 2) * PC_DRAIN_BYTES;
      uint32_t piece = 0;
      while (walk.next(P, pc)) {
        PT_BEGIN();
        mbar_wait(acc_full_bar, piece & 1);
        PT_END(22);
        tc_fence_after();
        const int row0 = pc.slot * P.n_self_pad + pc.ib * BW_BM + q * 32;
#pragma unroll 1
        for (int cc = 0; cc < 8; ++cc) {
          const int col = half * 256 + cc * 32;
          if (col >= P.dim) break;
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_addr(tmem, q * 32, col), v);
          tc_wait_ld();
          if (lane == 0) bulk_wait_read_all();  // the previous store has read the staging tile
          __syncwarp();
          uint8_t* rowp = stg_ptr + lane * 128;
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16)
            *reinterpret_cast<uint4*>(rowp + ((c16 ^ (lane & 7)) << 4)) =
                make_uint4(v[4 * c16], v[4 * c16 + 1], v[4 * c16 + 2], v[4 * c16 + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&P.tm_gpart[pc.job], stg, col, row0);
            bulk_commit_group();
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty_bar);  // this warp's part of the accumulator is in flight to HBM
        PT_END(20);
        ++piece;
      }
      if (lane == 0) bulk_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves while the other may still signal its barriers / write its memory
  if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace tcl

extern "C" int tcl_debug_pc_trace(unsigned long long* out32, int reset) {
  using namespace tcl;
  TCL_CHECK_CUDA(cudaDeviceSynchronize());
  if (out32) TCL_CHECK_CUDA(cudaMemcpyFromSymbol(out32, g_pc_trace, sizeof(unsigned long long) * 32));
  if (reset) {
    unsigned long long z[32] = {0};
    TCL_CHECK_CUDA(cudaMemcpyToSymbol(g_pc_trace, z, sizeof(z)));
  }
  return TCL_OK;
}

namespace tcl {

template <int kOp>
static int pc_max_clusters(int cluster_size, int* out) {
  const int smem = static_cast<int>(PcSmem::total());
  TCL_CHECK_CUDA(cudaFuncSetAttribute(ntxent_bwd_pc_kernel<kOp>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (cluster_size > 8)
    TCL_CHECK_CUDA(cudaFuncSetAttribute(ntxent_bwd_pc_kernel<kOp>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cluster_size * 64, 1, 1);
  cfg.blockDim = dim3(PC_THREADS);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster_size;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TCL_CHECK_CUDA(cudaOccupancyMaxActiveClusters(out, ntxent_bwd_pc_kernel<kOp>, &cfg));
  return TCL_OK;
}

}  // namespace tcl

// how many clusters of `cluster_size` CTAs of this kernel's footprint the device can hold at once
extern "C" int tcl_debug_max_clusters(int cluster_size, int* out) { return tcl::pc_max_clusters<TCL_OP_F16>(cluster_size, out); }

namespace tcl {

template <int kOp>
static int launch_bwd_pc_t(const BwdParams& P, int n_jobs, int* n_clusters_out, cudaStream_t st) {
  // co-resident 2-CTA clusters of this footprint (74 on a 148-SM B200), cached per device
  static int resident[64];
  int dev = 0;
  TCL_CHECK_CUDA(cudaGetDevice(&dev));
  TCL_REQUIRE(dev >= 0 && dev < 64, TCL_ERR_BAD_ARG, "device index %d", dev);
  if (resident[dev] == 0) {
    int n = 0;
    if (int e = pc_max_clusters<kOp>(2, &n)) return e;
    TCL_REQUIRE(n >= 1, TCL_ERR_CUDA_BASE, "producer/consumer backward: no resident cluster");
    resident[dev] = n;
  }
  const int64_t total = P.job_tile_base[TCL_MAX_TENSORS];
  int t_max = 1;
  for (int j = 0; j < n_jobs; ++j) t_max = P.unit_tiles[j] > t_max ? P.unit_tiles[j] : t_max;
  // a unit of T tiles is cut into at most ceil(T / range) + 1 pieces; the workspace holds kBwdMaxSplit partials
  int64_t n = resident[dev];
  if (const char* e = getenv("TRICOLO_B200_PC_CLUSTERS")) {  // experiments: fewer tile ranges than resident clusters
    const int v = atoi(e);
    if (v >= 1 && v < n) n = v;
  }
  if (n > total) n = total;
  const int64_t cap = (kBwdMaxSplit - 1) * total / t_max;  // range >= T / (kBwdMaxSplit - 1)
  if (n > cap) n = cap;
  if (n < 1) n = 1;
  *n_clusters_out = static_cast<int>(n);
  const int smem = static_cast<int>(PcSmem::total());
  LaunchCfg L(dim3(2 * static_cast<unsigned>(n), 1, 1), dim3(PC_THREADS), smem, st, 2);
  TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, ntxent_bwd_pc_kernel<kOp>, P));
  return TCL_OK;
}

// number of tile ranges the launch below will use (the normalise backward needs it before the launch is built)
int launch_bwd_pc(const BwdParams& P, int n_jobs, int op_format, int* n_clusters_out, cudaStream_t st) {
  return op_format == TCL_OP_F16 ? launch_bwd_pc_t<TCL_OP_F16>(P, n_jobs, n_clusters_out, st)
                                 : launch_bwd_pc_t<TCL_OP_BF16>(P, n_jobs, n_clusters_out, st);
}

}  // namespace tcl
