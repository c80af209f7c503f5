// K3, CTA-pair form (dim > 256): both GEMMs of the backward as 2-SM tcgen05 MMAs (cta_group::2), so the
// logit recompute is done ONCE per (row block, column tile) instead of once per dim half.
//
// A cluster of two CTAs owns 128 "self" rows (64 each) and ALL of dim:
//   S phase   S[128 x 128] = Zself · Zother_tile^T   2-SM MMA, M=128 (64 rows per CTA), N=128, K=dim.
//             A = each CTA's 64 resident self rows, B = the other-tile, N-split: each CTA TMA-loads 64 of
//             its 128 rows.  Per CTA the accumulator is 64 rows x 128 columns, stored as 128 lanes x 64
//             TMEM columns (tile columns 64..127 sit in lanes 64..127).
//   epilogue  each CTA turns ITS 64 rows into the 16-bit G tile [64 x 128] and stores it K-major/swizzled
//             in its own shared memory.
//   A phase   acc^T[d, i] += Zother^T[d, j] · G[i, j]      2-SM MMA, M=256 (128 dim rows per CTA), N=128
//             (= the 128 self rows), K=128.  B = G, N-split across the pair: every CTA contributes exactly
//             the 64 rows it has just produced — the tensor core reads the peer's half from the peer's shared
//             memory, so NOTHING is exchanged in software.  A = transposed other operand, each CTA loads its
//             own dim rows.  Two such accumulators per CTA (dim rows [256c + 128 rank, +128), c = 0,1).
// Executed flop per pair and direction: 2 B^2 D (recompute) + 2 B^2 D (gradient) — the algorithmic count —
// against 3x that in the independent-CTA kernel (ntxent_bwd.cu), which recomputes per dim half.
//
// Only the leader CTA (cluster rank 0) issues MMAs.  Barriers: TMA loads of both CTAs complete on the
// leader's `full` barriers; tcgen05.commit multicasts to both CTAs' `empty`, `s_full`, `g_empty`, `acc_full`;
// epilogue warps of both CTAs arrive on the leader's `s_empty` / `g_full` (remote arrives from the peer).
#include "ntxent_bwd.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

static constexpr int BP_STAGES = 6;           // 16 KB ring slots
static constexpr int BP_NBUF = 3;             // logit (TMEM) and G (smem) buffers: the logit MMAs run two tiles ahead of the
                                              // gradient MMAs, which gives the epilogue two tile times to turn S into G
static constexpr int BP_SLOT = 16384;
static constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address of the same offset in the even (leader) CTA

struct PairSmem {
  static constexpr uint32_t x_off = 0;  // num_kb * 8 KB  (64 self rows)
  static constexpr uint32_t g_off(int num_kb) { return num_kb * 8192; }                 // BP_NBUF buffers x 16 KB
  static constexpr uint32_t ring_off(int num_kb) { return g_off(num_kb) + BP_NBUF * 16384; }
  static constexpr uint32_t bar_off(int num_kb) { return ring_off(num_kb) + BP_STAGES * BP_SLOT; }
  static constexpr uint32_t bj_off(int num_kb) { return bar_off(num_kb) + 256; }  // 2 x 128 floats
  static constexpr uint32_t total(int num_kb) { return bj_off(num_kb) + 1024 + 1024; }
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
  auto s_empty_bar = [&](int b) { return bars + 8u * (2 * BP_STAGES + 1 + BP_NBUF + b); };
  auto g_full_bar = [&](int b) { return bars + 8u * (2 * BP_STAGES + 1 + 2 * BP_NBUF + b); };
  auto g_empty_bar = [&](int b) { return bars + 8u * (2 * BP_STAGES + 1 + 3 * BP_NBUF + b); };
  const uint32_t acc_full_bar = bars + 8u * (2 * BP_STAGES + 1 + 4 * BP_NBUF);
  const uint32_t tmem_slot = bars + 8u * (2 * BP_STAGES + 2 + 4 * BP_NBUF);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      base_ptr + PairSmem::bar_off(num_kb) + 8u * (2 * BP_STAGES + 2 + 4 * BP_NBUF));
  float* bj = reinterpret_cast<float*>(base_ptr + PairSmem::bj_off(num_kb));  // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();  // 0 = leader
  const bool leader = crank == 0;
  const int ib = blockIdx.x >> 1;  // the pair = two CTAs adjacent in x (cluster dims (2,1,1))
  const int split = blockIdx.y;
  const BwdJobDev& J = P.job[blockIdx.z];
  const int i0 = ib * BW_BM;                          // first self row of the pair
  const int i0_own = i0 + static_cast<int>(crank) * 64;  // first self row of this CTA
  const int n_chunk = (P.dim + 255) / 256;            // accumulators per CTA (dim rows [256c + 128 rank, +128))
  const int total_tiles = J.n_seg * P.n_jtiles;
  const int t_begin = static_cast<int>((static_cast<int64_t>(total_tiles) * split) / P.n_split);
  const int t_end = static_cast<int>((static_cast<int64_t>(total_tiles) * (split + 1)) / P.n_split);
  const int n_tiles = t_end - t_begin;
  const int n_sslot = (num_kb + 1) / 2;  // ring slots per logit tile: two 8 KB K-blocks per 16 KB slot

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
    for (int b = 0; b < BP_NBUF; ++b) {
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
  const uint32_t tmem_acc = tmem + BP_NBUF * 64;  // logit buffers: 3 x 64 columns; accumulators behind them (2 x 128)

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one() && n_tiles > 0) {
      const uint32_t x_full_leader = x_full_bar & kPeerBitMask;
      if (leader) mbar_arrive_expect_tx(x_full_bar, 2 * num_kb * 8192);
      for (int kb = 0; kb < num_kb; ++kb)
        tma_load_2d_2sm(x_smem + kb * 8192, &J.tm_self, x_full_leader, kb * BW_BK, i0_own);
      int it = 0;
      auto acquire = [&](uint32_t pair_bytes) -> int {
        const int s = it % BP_STAGES;
        const uint32_t ph = (it / BP_STAGES) & 1;
        mbar_wait_cluster(empty_bar(s), ph ^ 1);
        if (leader) mbar_arrive_expect_tx(full_bar(s), pair_bytes);
        ++it;
        return s;
      };
      auto load_s = [&](int t) {  // this CTA's 64 rows of the other-tile (N half), two K-blocks per slot
        const int tt = t_begin + t;
        const BwdSegDev& sg = J.seg[tt / P.n_jtiles];
        const int j0 = (tt % P.n_jtiles) * BW_BN + static_cast<int>(crank) * 64;
        for (int st = 0; st < n_sslot; ++st) {
          const int nk = (2 * st + 1 < num_kb) ? 2 : 1;
          const int s = acquire(2 * nk * 8192);
          for (int u = 0; u < nk; ++u)
            tma_load_2d_2sm(ring + s * BP_SLOT + u * 8192, &sg.tm_other, full_bar(s) & kPeerBitMask,
                            (2 * st + u) * BW_BK, j0);
        }
      };
      auto load_a = [&](int t) {  // transposed other operand: this CTA's 128 dim rows of each accumulator chunk
        const int tt = t_begin + t;
        const BwdSegDev& sg = J.seg[tt / P.n_jtiles];
        const int j0 = (tt % P.n_jtiles) * BW_BN;
        for (int kb2 = 0; kb2 < 2; ++kb2)
          for (int c = 0; c < n_chunk; ++c) {
            const int s = acquire(2 * BP_SLOT);
            tma_load_2d_2sm(ring + s * BP_SLOT, &sg.tm_other_t, full_bar(s) & kPeerBitMask, j0 + kb2 * BW_BK,
                            c * 256 + static_cast<int>(crank) * 128);
          }
      };
      // same order as the MMA issuer consumes: S(0), S(1), [S(t+2), A(t)] ...
      load_s(0);
      if (n_tiles > 1) load_s(1);
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 2 < n_tiles) load_s(t + 2);
        load_a(t);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && elect_one() && n_tiles > 0) {
      mbar_wait_cluster(x_full_bar, 0);
      int it = 0;
      auto issue_s = [&](int t) {
        const int b = t % BP_NBUF;
        mbar_wait_cluster(s_empty_bar(b), ((t / BP_NBUF) & 1) ^ 1);
        tc_fence_after();
        for (int st = 0; st < n_sslot; ++st, ++it) {
          const int s = it % BP_STAGES;
          const uint32_t ph = (it / BP_STAGES) & 1;
          mbar_wait_cluster(full_bar(s), ph);
          tc_fence_after();
          const int nk = (2 * st + 1 < num_kb) ? 2 : 1;
          for (int u = 0; u < nk; ++u) {
            const int kb = 2 * st + u;
            const uint64_t ad = umma_desc_k_sw128(x_smem + kb * 8192);
            const uint64_t bd = umma_desc_k_sw128(ring + s * BP_SLOT + u * 8192);
#pragma unroll
            for (int kk = 0; kk < BW_BK / 16; ++kk)
              tc_mma_f16_2sm(tmem + b * 64, ad + 2 * kk, bd + 2 * kk, P.idesc, (kb | kk) != 0);  // M=128 (pair), N=128
          }
          tc_commit_2sm(empty_bar(s), 0x3);
        }
        tc_commit_2sm(s_full_bar(b), 0x3);
      };
      issue_s(0);
      if (n_tiles > 1) issue_s(1);
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 2 < n_tiles) issue_s(t + 2);
        const int gb = t % BP_NBUF;
        mbar_wait_cluster(g_full_bar(gb), (t / BP_NBUF) & 1);
        tc_fence_after();
        for (int kb2 = 0; kb2 < 2; ++kb2) {
          for (int c = 0; c < n_chunk; ++c, ++it) {
            const int s = it % BP_STAGES;
            const uint32_t ph = (it / BP_STAGES) & 1;
            mbar_wait_cluster(full_bar(s), ph);
            tc_fence_after();
            const uint64_t ad = umma_desc_k_sw128(ring + s * BP_SLOT);                          // 128 dim rows per CTA
            const uint64_t bd = umma_desc_k_sw128(g_smem + gb * 16384 + kb2 * 8192);            // 64 G rows per CTA
#pragma unroll
            for (int kk = 0; kk < BW_BK / 16; ++kk)
              tc_mma_f16_2sm(tmem_acc + c * 128, ad + 2 * kk, bd + 2 * kk, P.idesc_m256, (t | kb2 | kk) != 0);  // M=256, N=128
            tc_commit_2sm(empty_bar(s), 0x3);
          }
        }
        tc_commit_2sm(g_empty_bar(gb), 0x3);
      }
      tc_commit_2sm(acc_full_bar, 0x3);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (both CTAs)
    // 2-SM M=128 accumulator of one CTA: TMEM lane L holds self row (L % 64); tile columns (L / 64) * 64 + [0,64)
    // sit in the buffer's 64 TMEM columns.  Two warps per lane quarter split those 64 columns.
    const int q = warp & 3;
    const int ch = (warp - 2) >> 2;                 // 32-column half of the 64 TMEM columns
    const int tl = q * 32 + lane;                   // TMEM lane
    const int r = tl & 63;                          // self row inside this CTA's 64
    const int chalf = tl >> 6;                      // which 64-column half of the tile (== G K-block)
    const int et = threadIdx.x - 64;                // 0..255
    const int grow = i0_own + r;                    // local self row index
    float gs[2] = {0.f, 0.f};
    float gmax = 0.f;
    for (int s = 0; s < J.n_seg; ++s) {
      gs[s] = J.seg[s].grad_scale ? *J.seg[s].grad_scale : 1.f;
      gmax = fmaxf(gmax, fabsf(gs[s]));
    }
    const float inv_gmax = gmax > 0.f ? 1.f / gmax : 0.f;
    if (blockIdx.x == 0 && blockIdx.y == 0 && et == 0) *J.scale_out = gmax * P.out_scale;  // one CTA per job

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
    const uint32_t s_empty_leader0 = map_to_peer(s_empty_bar(0), 0);  // barriers are 8 bytes apart
    const uint32_t g_full_leader0 = map_to_peer(g_full_bar(0), 0);

    for (int t = 0; t < n_tiles; ++t) {
      const int tt = t_begin + t;
      const int si = tt / P.n_jtiles;
      const int j0 = (tt % P.n_jtiles) * BW_BN;
      const int b = t % BP_NBUF;
      const uint32_t bpar = (t / BP_NBUF) & 1;
      if (si != cur_seg) {
        const BwdSegDev& sg = J.seg[si];
        cur_seg = si;
        rr = gs[si] * inv_gmax;
        lse_i = grow < P.n_self ? sg.lse2_self[grow] : 0.f;
        ws = rr * sg.w_self;
        wo_i = rr * sg.w_other * ex2_approx(lse_i - P.c1);
      }
      const float bj_next = load_bj(t + 1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const int ccol0 = chalf * 64 + ch * 32;  // first tile column of this thread's 32
      const float* bjt = bj + (t & 1) * 128 + ccol0;
      const int dl = P.self_offset + grow - j0 - ccol0;  // column of the positive inside our 32 (if any)

      mbar_wait_cluster(s_full_bar(b), bpar);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_addr(tmem + b * 64, q * 32, ch * 32), v);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {  // logit buffer drained: tell the leader's MMA thread
        if (leader) mbar_arrive(s_empty_bar(b)); else mbar_arrive_remote(s_empty_leader0 + 8u * b);
      }
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(v[e]), P.c1, -lse_i));
        const float p1 = ex2_approx(fmaf(__uint_as_float(v[e + 1]), P.c1, -lse_i));
        const float g0 = fmaf(p0, fmaf(wo_i, bjt[e], ws), (e == dl) ? -rr : 0.f);
        const float g1 = fmaf(p1, fmaf(wo_i, bjt[e + 1], ws), (e + 1 == dl) ? -rr : 0.f);
        pk[e >> 1] = pack2<kOp>(g0, g1);
      }
      // G buffer (t % 3) is free once the gradient MMAs of tile t-3 have completed (both CTAs' halves were read)
      const int gb = b;
      mbar_wait_cluster(g_empty_bar(gb), bpar ^ 1);
      // K-major swizzled B operand: K-block = chalf, row r (of 64), 16-byte chunks ch*4 .. ch*4+3
      uint8_t* grow_ptr = base_ptr + PairSmem::g_off(num_kb) + gb * 16384 + chalf * 8192 + r * 128;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4)
        *reinterpret_cast<uint4*>(grow_ptr + (((ch * 4 + c4) ^ (r & 7)) << 4)) =
            make_uint4(pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
      fence_proxy_async_smem();  // generic-proxy stores -> async proxy (the pair's tensor cores read this buffer)
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(g_full_bar(gb)); else mbar_arrive_remote(g_full_leader0 + 8u * gb);
      }
      if (et < 128 && t + 1 < n_tiles) bj[((t + 1) & 1) * 128 + et] = bj_next;
    }

    if (n_tiles > 0) {
      mbar_wait_cluster(acc_full_bar, 0);
      tc_fence_after();
    }
    // accumulator read-out: the accumulator is transposed (TMEM lane = dim row within this CTA's 128 of chunk c,
    // TMEM column = self row), so for a fixed register index the 32 lanes of a warp hold 32 consecutive dim
    // entries of ONE gradient row: every store instruction writes 128 contiguous bytes, no staging needed.
    for (int c = 0; c < n_chunk; ++c) {
      const int d_base = c * 256 + static_cast<int>(crank) * 128 + q * 32;  // dim index of this warp's lane 0
#pragma unroll 1
      for (int cb = ch * 2; cb < ch * 2 + 2; ++cb) {  // this warp's two 32-row blocks of the 128 self rows
        uint32_t v[32];
        if (n_tiles > 0) {
          tmem_ld_32x32b_x32(tmem_addr(tmem_acc + c * 128, q * 32, cb * 32), v);
          tc_wait_ld();
        } else {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = 0u;
        }
        const int d = d_base + lane;
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

int launch_bwd_pair(const BwdParams& P, int n_iblocks, int n_jobs, int op_format, cudaStream_t st) {
  const int smem = static_cast<int>(PairSmem::total(P.num_kb));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * n_iblocks, P.n_split, n_jobs);
  cfg.blockDim = dim3(BW_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
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
    {
      cudaError_t le = cudaLaunchKernelEx(&cfg, ntxent_bwd_pair_kernel<TCL_OP_F16>, P);
      if (le != cudaSuccess) {
        int ncl = -1;
        cudaError_t qe = cudaOccupancyMaxActiveClusters(&ncl, ntxent_bwd_pair_kernel<TCL_OP_F16>, &cfg);
        return set_error(TCL_ERR_CUDA_BASE + (int)le, "pair kernel launch: %s (grid %u,%u,%u smem %d; max active clusters %d, query %s)",
                         cudaGetErrorString(le), cfg.gridDim.x, cfg.gridDim.y, cfg.gridDim.z, smem, ncl, cudaGetErrorString(qe));
      }
    }
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
