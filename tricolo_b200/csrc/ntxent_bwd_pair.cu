// K3, CTA-pair form (dim > 256): both GEMMs of the backward as full-rate 2-SM tcgen05 MMAs (cta_group::2,
// M=256), the logit tile recomputed ONCE per (row block, column tile) instead of once per dim half.
//
// A cluster of two CTAs (adjacent in x) owns 128 "self" rows and ALL of dim.  One step = two column tiles:
//   S phase   S^T[256 j x 128 i] = Zother[2 tiles] · Zself^T      2-SM MMA, M=256, N=128, K=dim.
//             A = the other-operand tile of each CTA (CTA h streams tile 2u+h), B = the self block, N-split:
//             each CTA keeps 64 of the 128 self rows resident.  Per CTA the accumulator is S^T of ITS tile:
//             128 lanes (j) x 128 TMEM columns (i).
//             (M=128 over two SMs — 64 rows per SM — runs the tensor cores at half rate, measured: 1.02 ms for
//             the whole backward, no better than the independent-CTA kernel; mapping j to M avoids it.)
//   epilogue  thread = column j of the logits.  It turns its S values into the 16-bit gradient weights G[i, j]
//             and stores them as the B operand of the gradient MMA: rows = K index j, 128 bytes = 64 self rows i
//             (MN-major, 128-byte swizzle).  That operand is N-split by self-row halves across the pair, so the
//             half with i < 64 goes to CTA 0's shared memory and the half with i >= 64 to CTA 1's: of the two
//             warps per lane quarter one writes locally, the other through DSMEM (16 KB per step and direction).
//   A phase   acc^T[d, i] += Zother^T[d, j] · G[i, j]     2-SM MMA, M=256 (128 dim rows per CTA), N=128, K=256
//             (the step's two tiles).  Two accumulators per CTA: dim rows [256c + 128 rank, +128), c = 0,1.
// Executed flop per pair and direction: 2 B^2 D (recompute) + 2 B^2 D (gradient) = the algorithmic count, against
// 3x the gradient term in the independent-CTA kernel (ntxent_bwd.cu).
//
// Only the leader CTA (cluster rank 0) issues MMAs.  TMA loads of both CTAs complete on the leader's `full`
// barriers; tcgen05.commit multicasts to both CTAs' `empty`, `s_full`, `g_empty`, `acc_full`; the epilogue warps
// of both CTAs arrive on the leader's `s_empty` / `g_full` (remote arrives from the peer).
#include "ntxent_bwd.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

static constexpr int BP_STAGES = 3;    // ring slots
static constexpr int BP_BLK = 16384;   // one TMA box: 128 rows x 64 K elements
static constexpr int BP_SLOT = 2 * BP_BLK;
static constexpr int BP_GBUF = 32768;  // one step of G: 256 K rows x 128 bytes
static constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // same offset in the even (leader) CTA of the pair

struct PairSmem {
  static constexpr uint32_t x_off = 0;                                                   // num_kb * 8 KB: 64 self rows
  static constexpr uint32_t g_off(int num_kb) { return num_kb * 8192; }                  // 2 buffers x 32 KB
  static constexpr uint32_t ring_off(int num_kb) { return g_off(num_kb) + 2 * BP_GBUF; }
  static constexpr uint32_t bar_off(int num_kb) { return ring_off(num_kb) + BP_STAGES * BP_SLOT; }
  static constexpr uint32_t col_off(int num_kb) { return bar_off(num_kb) + 256; }        // [lse_i | wo_i][128] floats
  static constexpr uint32_t total(int num_kb) { return col_off(num_kb) + 1024 + 1024; }
};

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
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t smem_dst, const CUtensorMap* m, uint32_t leader_bar, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}

// Optional wait-time accounting of the first cluster (make trace): cycles each role spends in each wait.
__device__ unsigned long long g_pair_trace[32];
#ifdef TCL_PAIR_TRACE
#define TR_DECL unsigned long long tr_t0 = 0; const bool tr_on = blockIdx.x < 2 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0; (void)tr_t0;
#endif
#if defined(TCL_PAIR_TRACE) && TCL_PAIR_TRACE >= 2
#define TR_BEGIN() do { if (tr_on) tr_t0 = clock64(); } while (0)
#define TR_END(slot) do { if (tr_on) atomicAdd(&g_pair_trace[(slot) + 16 * (blockIdx.x & 1)], clock64() - tr_t0); } while (0)
#else
#ifndef TCL_PAIR_TRACE
#define TR_DECL
#endif
#define TR_BEGIN() do {} while (0)
#define TR_END(slot) do {} while (0)
#endif
// Timing experiments (never in the product build): TCL_PAIR_EXP 1 = MMA thread ignores the `full` barriers,
// 2 = also no ring traffic at all (no TMA loads, no commits to `empty`), 3 = also no epilogue handshakes.
#ifndef TCL_PAIR_EXP
#define TCL_PAIR_EXP 0
#endif
// slots (+16 for the peer CTA): 0 producer empty-wait, 1 mma s_empty, 2 mma full (S), 3 mma g_full, 4 mma full (A),
// 5 mma total, 12 mma issue (S), 13 commits, 14 mma issue (A), 6 epi s_full wait, 7 epi g_empty wait, 8 epi tmem-ld, 9 epi compute+stores, 10 epi fence+arrive, 11 epi total

template <int kOp>
__global__ void __launch_bounds__(BW_THREADS, 1) ntxent_bwd_pair_kernel(const __grid_constant__ BwdParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int num_kb = P.num_kb;
  const uint32_t x_smem = base + PairSmem::x_off;
  const uint32_t g_smem = base + PairSmem::g_off(num_kb);
  const uint32_t ring = base + PairSmem::ring_off(num_kb);
  const uint32_t bars = base + PairSmem::bar_off(num_kb);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (BP_STAGES + s); };
  const uint32_t x_full_bar = bars + 8u * (2 * BP_STAGES);
  auto s_full_bar = [&](int b) { return bars + 8u * (2 * BP_STAGES + 1 + b); };
  auto s_empty_bar = [&](int b) { return bars + 8u * (2 * BP_STAGES + 3 + b); };
  auto g_full_bar = [&](int b) { return bars + 8u * (2 * BP_STAGES + 5 + b); };
  auto g_empty_bar = [&](int b) { return bars + 8u * (2 * BP_STAGES + 7 + b); };
  const uint32_t acc_full_bar = bars + 8u * (2 * BP_STAGES + 9);
  const uint32_t tmem_slot = bars + 8u * (2 * BP_STAGES + 10);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      base_ptr + PairSmem::bar_off(num_kb) + 8u * (2 * BP_STAGES + 10));
  float* colc = reinterpret_cast<float*>(base_ptr + PairSmem::col_off(num_kb));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TR_DECL
  const uint32_t crank = cluster_ctarank();  // 0 = leader
  const bool leader = crank == 0;
  const int ib = blockIdx.x >> 1;  // the pair = two CTAs adjacent in x (cluster dims (2,1,1))
  const int split = blockIdx.y;
  const BwdJobDev& J = P.job[blockIdx.z];
  const int i0 = ib * BW_BM;                             // first self row of the pair
  const int i0_own = i0 + static_cast<int>(crank) * 64;  // first of the 64 self rows resident in this CTA
  const int n_chunk = (P.dim + 255) / 256;               // accumulators per CTA (dim rows [256c + 128 rank, +128))
  const int total_tiles = J.n_seg * P.n_jtiles;
  const int total_steps = (total_tiles + 1) / 2;
  const int u_begin = static_cast<int>((static_cast<int64_t>(total_steps) * split) / P.n_split);
  const int u_end = static_cast<int>((static_cast<int64_t>(total_steps) * (split + 1)) / P.n_split);
  const int n_steps = u_end - u_begin;
  // tile of CTA h in step u: 2u + h (one past the end in the last step of an odd tile count: contributes nothing)
  auto tile_of = [&](int u, int h) { return 2 * (u_begin + u) + h; };

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&J.tm_self);
    for (int s = 0; s < J.n_seg; ++s) {
      tma_prefetch_desc(&J.seg[s].tm_other);
      tma_prefetch_desc(&J.seg[s].tm_other_t);
    }
    for (int s = 0; s < BP_STAGES; ++s) {
      mbar_init(full_bar(s), 1);   // leader's: one arrive.expect_tx by the leader's producer (bytes of both CTAs)
      mbar_init(empty_bar(s), 1);  // one multicast commit
    }
    mbar_init(x_full_bar, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(s_full_bar(b), 1);
      mbar_init(s_empty_bar(b), 2 * BW_EPI_WARPS);  // leader's: epilogue warps of both CTAs
      mbar_init(g_full_bar(b), 2 * BW_EPI_WARPS);   // leader's
      mbar_init(g_empty_bar(b), 1);
    }
    mbar_init(acc_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tmem_acc = tmem + 256;  // logit buffers: columns [0,128) and [128,256); accumulators [256,512)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    // Ring slots hold TWO 16 KB operand blocks (8 MMAs per full/empty handshake): the per-slot wait + commit
    // latency of the single MMA thread is what keeps 64-cycle MMAs off the tensor-core floor.
    if (elect_one() && n_steps > 0) {
      const uint32_t x_full_leader = x_full_bar & kPeerBitMask;
      if (leader) mbar_arrive_expect_tx(x_full_bar, 2 * num_kb * 8192);
      for (int kb = 0; kb < num_kb; ++kb)
        tma_load_2d_2sm(x_smem + kb * 8192, &J.tm_self, x_full_leader, kb * BW_BK, i0_own);
      int it = 0;
      auto acquire = [&](uint32_t pair_bytes) -> int {
        const int s = it % BP_STAGES;
        const uint32_t ph = (it / BP_STAGES) & 1;
        TR_BEGIN();
        mbar_wait(empty_bar(s), ph ^ 1);
        TR_END(0);
        if (leader) mbar_arrive_expect_tx(full_bar(s), pair_bytes);
        ++it;
        return s;
      };
      auto load_s = [&](int u) {  // A operand of the logit MMA: this CTA's tile of the other operand, all of K
        const int t_own = tile_of(u, static_cast<int>(crank));
        const bool own_ok = t_own < total_tiles;
        const bool peer_ok = tile_of(u, static_cast<int>(crank) ^ 1) < total_tiles;
        for (int kb = 0; kb < num_kb; kb += 2) {
          const int nk = kb + 1 < num_kb ? 2 : 1;
          const int s = acquire(static_cast<uint32_t>(nk) * ((own_ok ? BP_BLK : 0) + (peer_ok ? BP_BLK : 0)));
          if (own_ok) {
            const BwdSegDev& sg = J.seg[t_own / P.n_jtiles];
            for (int k2 = 0; k2 < nk; ++k2)
              tma_load_2d_2sm(ring + s * BP_SLOT + k2 * BP_BLK, &sg.tm_other, full_bar(s) & kPeerBitMask,
                              (kb + k2) * BW_BK, (t_own % P.n_jtiles) * BW_BN);
          }
        }
      };
      auto load_a = [&](int u) {  // A operand of the gradient MMA: transposed other operand, this CTA's 128 dim rows
        for (int h = 0; h < 2; ++h) {
          const int tt = tile_of(u, h);
          if (tt >= total_tiles) break;  // both CTAs and the MMA issuer skip the missing tile identically
          const BwdSegDev& sg = J.seg[tt / P.n_jtiles];
          const int j0 = (tt % P.n_jtiles) * BW_BN;
          for (int kb2 = 0; kb2 < 2; ++kb2) {
            const int s = acquire(static_cast<uint32_t>(n_chunk) * 2 * BP_BLK);
            for (int c = 0; c < n_chunk; ++c)
              tma_load_2d_2sm(ring + s * BP_SLOT + c * BP_BLK, &sg.tm_other_t, full_bar(s) & kPeerBitMask,
                              j0 + kb2 * BW_BK, c * 256 + static_cast<int>(crank) * 128);
          }
        }
      };
#if TCL_PAIR_EXP == 4
      // second issuing thread: all gradient-shaped MMAs, no synchronisation at all (issue-rate experiment)
      if (leader) {
        const unsigned long long t0 = clock64();
        for (int u = 0; u < n_steps; ++u)
          for (int g8 = 0; g8 < 4; ++g8)
            for (int c = 0; c < n_chunk; ++c) {
              const uint64_t ad = umma_desc_k_sw128(ring + c * BP_BLK);
              const uint64_t bd = umma_desc_k_sw128(g_smem + g8 * 8192);
#pragma unroll
              for (int kk = 0; kk < BW_BK / 16; ++kk)
                tc_mma_f16_2sm(tmem_acc + c * 128, ad + 2 * kk, bd + 128 * kk, P.idesc_m256_bmn, 1);
            }
        tc_commit_2sm(acc_full_bar, 0x3);
        if (tr_on) atomicAdd(&g_pair_trace[4], clock64() - t0);
      }
#elif TCL_PAIR_EXP < 2
      load_s(0);
      for (int u = 0; u < n_steps; ++u) {
        if (u + 1 < n_steps) load_s(u + 1);
        load_a(u);
      }
#endif
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && elect_one() && n_steps > 0) {
      mbar_wait(x_full_bar, 0);
#ifdef TCL_PAIR_TRACE
      const unsigned long long tr_mma0 = clock64();
#endif
      int it = 0;
      auto issue_s = [&](int u) {
        const int b = u & 1;
        TR_BEGIN();
        if (TCL_PAIR_EXP < 3) mbar_wait(s_empty_bar(b), ((u >> 1) & 1) ^ 1);
        TR_END(1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; kb += 2, ++it) {
          const int s = it % BP_STAGES;
          const uint32_t ph = (it / BP_STAGES) & 1;
          const int nk = kb + 1 < num_kb ? 2 : 1;
          TR_BEGIN();
          if (TCL_PAIR_EXP < 1) mbar_wait(full_bar(s), ph);
          TR_END(2);
          tc_fence_after();
          TR_BEGIN();
          for (int k2 = 0; k2 < nk; ++k2) {
            const uint64_t ad = umma_desc_k_sw128(ring + s * BP_SLOT + k2 * BP_BLK);  // 128 other rows per CTA (M = 256)
            const uint64_t bd = umma_desc_k_sw128(x_smem + (kb + k2) * 8192);         // 64 self rows per CTA (N = 128)
#pragma unroll
            for (int kk = 0; kk < BW_BK / 16; ++kk)
              tc_mma_f16_2sm(tmem + b * 128, ad + 2 * kk, bd + 2 * kk, P.idesc_m256, (kb | k2 | kk) != 0);
          }
          TR_END(12);
          TR_BEGIN();
          if (TCL_PAIR_EXP < 2) tc_commit_2sm(empty_bar(s), 0x3);
          TR_END(13);
        }
        if (TCL_PAIR_EXP < 3) tc_commit_2sm(s_full_bar(b), 0x3);
      };
      issue_s(0);
      for (int u = 0; u < n_steps; ++u) {
        if (u + 1 < n_steps) issue_s(u + 1);
        if (TCL_PAIR_EXP == 4) continue;
        const int gb = u & 1;
        TR_BEGIN();
        if (TCL_PAIR_EXP < 3) mbar_wait_cluster(g_full_bar(gb), (u >> 1) & 1);  // acquires the peer's generic-proxy (DSMEM) stores
        TR_END(3);
        tc_fence_after();
        for (int h = 0; h < 2; ++h) {
          if (tile_of(u, h) >= total_tiles) break;
          for (int kb2 = 0; kb2 < 2; ++kb2, ++it) {
            // B = K rows h*128 + kb2*64 .. +63 of this step's G buffer: MN-major, 128 bytes per K row, so one
            // UMMA_K (16 K rows) advances the start address by 2048 bytes
            const uint32_t gk = g_smem + gb * BP_GBUF + (h * 128 + kb2 * 64) * 128;
            const int s = it % BP_STAGES;
            const uint32_t ph = (it / BP_STAGES) & 1;
            TR_BEGIN();
            if (TCL_PAIR_EXP < 1) mbar_wait(full_bar(s), ph);
            TR_END(4);
            tc_fence_after();
            TR_BEGIN();
            for (int c = 0; c < n_chunk; ++c) {
              const uint64_t ad = umma_desc_k_sw128(ring + s * BP_SLOT + c * BP_BLK);  // 128 dim rows per CTA (M = 256)
              const uint64_t bd = umma_desc_k_sw128(gk);  // same descriptor fields; B is MN-major in the idesc
#pragma unroll
              for (int kk = 0; kk < BW_BK / 16; ++kk)
                tc_mma_f16_2sm(tmem_acc + c * 128, ad + 2 * kk, bd + 128 * kk, P.idesc_m256_bmn,
                               (u | h | kb2 | kk) != 0);
            }
            TR_END(14);
            TR_BEGIN();
            if (TCL_PAIR_EXP < 2) tc_commit_2sm(empty_bar(s), 0x3);
            TR_END(13);
          }
        }
        if (TCL_PAIR_EXP < 3) tc_commit_2sm(g_empty_bar(gb), 0x3);
      }
      if (TCL_PAIR_EXP != 4) tc_commit_2sm(acc_full_bar, 0x3);
#ifdef TCL_PAIR_TRACE
      if (tr_on) {
        atomicAdd(&g_pair_trace[5], clock64() - tr_mma0);
        atomicAdd(&g_pair_trace[15], static_cast<unsigned long long>(n_steps));
      }
#endif
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    // Accumulator of this CTA: S^T of its own tile, TMEM lane = column j of the logits, TMEM column = self row i.
    // Two warps per lane quarter; each handles 32 self rows of BOTH row halves: first the half that belongs to the
    // peer's G buffer (its DSMEM stores then drain while the second half is computed), then the local half.
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;
    const int jl = q * 32 + lane;     // tile-local column j == TMEM lane
    const int et = threadIdx.x - 64;  // 0..255
    float gs[2] = {0.f, 0.f};
    float gmax = 0.f;
    for (int s = 0; s < J.n_seg; ++s) {
      gs[s] = J.seg[s].grad_scale ? *J.seg[s].grad_scale : 1.f;
      gmax = fmaxf(gmax, fabsf(gs[s]));
    }
    const float inv_gmax = gmax > 0.f ? kGScale / gmax : 0.f;  // G in [-kGScale, kGScale]: see ntxent_bwd.h
    if (blockIdx.x == 0 && blockIdx.y == 0 && et == 0) *J.scale_out = gmax * P.out_scale * (1.f / kGScale);  // one CTA per job

    int cur_seg = -1;
    uint8_t* g_local_ptr = base_ptr + PairSmem::g_off(num_kb);
    const uint32_t g_peer = map_to_peer(g_smem, crank ^ 1u);
    const uint32_t s_empty_leader0 = map_to_peer(s_empty_bar(0), 0);  // barriers are 8 bytes apart
    const uint32_t g_full_leader0 = map_to_peer(g_full_bar(0), 0);
    const int krow = static_cast<int>(crank) * 128 + jl;  // K row of this thread inside a step's G buffer
    const int col_far = static_cast<int>(crank ^ 1u) * 64 + ch * 32;  // first self row (pair-local) of the far part
    const int col_near = static_cast<int>(crank) * 64 + ch * 32;

    for (int u = 0; u < (TCL_PAIR_EXP < 3 ? n_steps : 0); ++u) {
      const int tt = tile_of(u, static_cast<int>(crank));
      const bool tile_ok = tt < total_tiles;
      const int si = tile_ok ? tt / P.n_jtiles : (cur_seg < 0 ? 0 : cur_seg);
      const BwdSegDev& sg = J.seg[si];
      const int j0 = tile_ok ? (tt % P.n_jtiles) * BW_BN : 0;
      const int b = u & 1;
      const uint32_t bpar = (u >> 1) & 1;
      const float rr = si == 0 ? gs[0] * inv_gmax : gs[1] * inv_gmax;
      const float ws = rr * sg.w_self;
      if (si != cur_seg) {  // uniform over the CTA: per-self-row constants of the segment, one row per thread
        cur_seg = si;
        asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone is done with the previous segment's constants
        if (et < 128) {
          const int gi = i0 + et;
          const float l = gi < P.n_self ? sg.lse2_self[gi] : 0.f;
          colc[et] = l;
          colc[128 + et] = rr * sg.w_other * ex2_approx(l - P.c1);
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      const int j = j0 + jl;
      const float bj = (tile_ok && j < P.n_other) ? ex2_approx(P.c1 - sg.lse2_other[j]) : 0.f;
      const int dpos = j - P.self_offset - i0;  // pair-local self row of this column's positive (if in [0,128))
      const uint32_t off_row = static_cast<uint32_t>(b * BP_GBUF + krow * 128);

      TR_BEGIN();
      mbar_wait(s_full_bar(b), bpar);
      TR_END(6);
      TR_BEGIN();
      tc_fence_after();
      uint32_t vf[32], vn[32];
      tmem_ld_32x32b_x32(tmem_addr(tmem + b * 128, q * 32, col_far), vf);
      tmem_ld_32x32b_x32(tmem_addr(tmem + b * 128, q * 32, col_near), vn);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // logit buffer drained: tell the leader's MMA thread (orders TMEM reads only: relaxed)
        if (leader) mbar_arrive(s_empty_bar(b)); else mbar_arrive_remote_relaxed(s_empty_leader0 + 8u * b);
      }
      TR_END(8);
      // G buffer (u & 1) is free once the gradient MMAs of step u-2 have completed (they read both CTAs' halves)
      TR_BEGIN();
      mbar_wait(g_empty_bar(b), bpar ^ 1);
      TR_END(7);
      TR_BEGIN();
      // B operand of the gradient MMA, MN-major: 16-byte chunk c of K row k lives at chunk position (c ^ (k & 7))
#pragma unroll
      for (int part = 0; part < 2; ++part) {
        const int col0 = part == 0 ? col_far : col_near;
        const float* lse_i = colc + col0;
        const float* wo_i = colc + 128 + col0;
        const int dl = dpos - col0;
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const float s0 = __uint_as_float(part == 0 ? vf[e] : vn[e]);
          const float s1 = __uint_as_float(part == 0 ? vf[e + 1] : vn[e + 1]);
          const float p0 = ex2_approx(fmaf(s0, P.c1, -lse_i[e]));
          const float p1 = ex2_approx(fmaf(s1, P.c1, -lse_i[e + 1]));
          const float g0 = fmaf(p0, fmaf(wo_i[e], bj, ws), (e == dl) ? -rr : 0.f);
          const float g1 = fmaf(p1, fmaf(wo_i[e + 1], bj, ws), (e + 1 == dl) ? -rr : 0.f);
          pk[e >> 1] = tile_ok ? pack2<kOp>(g0, g1) : 0u;
        }
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const uint32_t off = off_row + (((ch * 4 + c4) ^ (krow & 7)) << 4);
          if (part == 0)
            st_cluster_v4(g_peer + off, pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
          else
            *reinterpret_cast<uint4*>(g_local_ptr + off) = make_uint4(pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
        }
      }
      TR_END(9);
      TR_BEGIN();
      fence_proxy_async_all();  // generic-proxy stores (local and DSMEM) -> async proxy (tensor cores of both SMs)
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(g_full_bar(b)); else mbar_arrive_remote(g_full_leader0 + 8u * b);
      }
      TR_END(10);
    }

    if (n_steps > 0) {
      mbar_wait(acc_full_bar, 0);
      tc_fence_after();
    }
    // accumulator read-out: the accumulator is transposed (TMEM lane = dim row within this CTA's 128 of chunk c,
    // TMEM column = self row), so for a fixed register index the 32 lanes of a warp hold 32 consecutive dim
    // entries of ONE gradient row: every store instruction writes 128 contiguous bytes, no staging needed.
    for (int c = 0; c < n_chunk; ++c) {
      const int d = c * 256 + static_cast<int>(crank) * 128 + q * 32 + lane;
#pragma unroll 1
      for (int cb = ch * 2; cb < ch * 2 + 2; ++cb) {  // this warp's two 32-row blocks of the 128 self rows
        uint32_t v[32];
        if (n_steps > 0) {
          tmem_ld_32x32b_x32(tmem_addr(tmem_acc + c * 128, q * 32, cb * 32), v);
          tc_wait_ld();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0u;
        }
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int srow = i0 + cb * 32 + e;
          if (srow < P.n_self && d < P.dim)
            J.gpart[(static_cast<int64_t>(split) * P.n_self + srow) * P.dim + d] = __uint_as_float(v[e]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem, 512);
}

}  // namespace tcl

extern "C" int tcl_debug_pair_trace(unsigned long long* out32, int reset) {
  using namespace tcl;
  TCL_CHECK_CUDA(cudaDeviceSynchronize());
  if (out32) TCL_CHECK_CUDA(cudaMemcpyFromSymbol(out32, g_pair_trace, sizeof(unsigned long long) * 32));
  if (reset) {
    unsigned long long z[32] = {0};
    TCL_CHECK_CUDA(cudaMemcpyToSymbol(g_pair_trace, z, sizeof(z)));
  }
  return TCL_OK;
}

namespace tcl {

int launch_bwd_pair(const BwdParams& P, int n_iblocks, int n_jobs, int op_format, cudaStream_t st) {
  const int smem = static_cast<int>(PairSmem::total(P.num_kb));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * n_iblocks, P.n_split, n_jobs);
  cfg.blockDim = dim3(BW_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;  // the CTA pair of a 2-SM MMA must be adjacent in x
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (op_format == TCL_OP_F16) {
    static int set = 0;
    if (set < smem) {
      TCL_CHECK_CUDA(cudaFuncSetAttribute(ntxent_bwd_pair_kernel<TCL_OP_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      set = smem;
    }
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ntxent_bwd_pair_kernel<TCL_OP_F16>, P));
  } else {
    static int set = 0;
    if (set < smem) {
      TCL_CHECK_CUDA(cudaFuncSetAttribute(ntxent_bwd_pair_kernel<TCL_OP_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      set = smem;
    }
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ntxent_bwd_pair_kernel<TCL_OP_BF16>, P));
  }
  return TCL_OK;
}

}  // namespace tcl
