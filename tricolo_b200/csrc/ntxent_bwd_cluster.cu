// K3, clustered form (dim > 256): the two dim-half CTAs of one (row block, split, job) form a
// thread-block cluster of 2 and share the logit recompute instead of duplicating it.
//
//   CTA h (= cluster rank = dim half) re-forms only columns [64h, 64h+64) of every 128-column
//   logit tile (M=128, N=64, K=dim), turns them into the 16-bit G half-tile, and writes that
//   half-tile into the G operand buffer of BOTH CTAs (local st.shared + st.shared::cluster into
//   the peer, DSMEM).  Each CTA then runs the gradient MMAs acc[128 x 256] += G[128 x 128] · Zother
//   for its own dim half.  Executed flop per pair: 8·B²·D instead of 12·B²·D.
//
// Synchronisation per CTA (mbarriers, all at identical shared-memory offsets in both CTAs):
//   g_full[kb]   8 arrivals (one per epilogue warp) of CTA kb (local arrive or remote
//                mbarrier.arrive.release.cluster) once their half-tile is stored and fenced
//   g_empty      2 arrivals: tcgen05.commit of the local gradient MMAs and the multicast commit
//                of the peer's -> both CTAs have finished reading G before it is overwritten
// A cluster barrier brackets the kernel so no CTA touches a peer that has not started or has exited.
#include "ntxent_bwd.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

template <int kOp>
__global__ void __launch_bounds__(BW_THREADS, 1) ntxent_bwd_cluster_kernel(const __grid_constant__ BwdParams P) {
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
  auto g_full_bar = [&](int kb) { return bars + 8u * (2 * BW_STAGES + 5 + kb); };
  const uint32_t g_empty_bar = bars + 8u * (2 * BW_STAGES + 7);
  const uint32_t acc_full_bar = bars + 8u * (2 * BW_STAGES + 8);
  const uint32_t tmem_slot = bars + 8u * (2 * BW_STAGES + 9);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(
      base_ptr + BwdSmem::bar_off(num_kb) + 8u * (2 * BW_STAGES + 9));
  float* bj = reinterpret_cast<float*>(base_ptr + BwdSmem::bj_off(num_kb));  // [2][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t h = cluster_ctarank();  // dim half AND logit column half
  const uint32_t peer = h ^ 1u;
  const int ib = blockIdx.x;
  const int split = blockIdx.y >> 1;
  const BwdJobDev& J = P.job[blockIdx.z];
  const int i0 = ib * BW_BM;
  const int d0 = static_cast<int>(h) * BW_DH;
  const int n_dc = (P.dim - d0) >= BW_DH ? 2 : ((P.dim - d0) + 127) / 128;
  const int total_tiles = J.n_seg * P.n_jtiles;
  const int t_begin = static_cast<int>((static_cast<int64_t>(total_tiles) * split) / P.n_split);
  const int t_end = static_cast<int>((static_cast<int64_t>(total_tiles) * (split + 1)) / P.n_split);
  const int n_tiles = t_end - t_begin;
  const int n_sstage = (num_kb + 1) / 2;  // ring stages per logit half-tile (two 64-wide K blocks each)

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
      mbar_init(s_empty_bar(b), BW_EPI_WARPS);  // one elected arrive per epilogue warp
      mbar_init(g_full_bar(b), BW_EPI_WARPS);
    }
    mbar_init(g_empty_bar, 2);
    mbar_init(acc_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const uint32_t tmem_acc = tmem + 128;  // columns 128..383; logit half-tiles at 0..63 / 64..127

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one() && n_tiles > 0) {
      mbar_arrive_expect_tx(x_full_bar, num_kb * BW_KB_BYTES);
      for (int kb = 0; kb < num_kb; ++kb)
        tma_load_2d(x_smem + kb * BW_KB_BYTES, &J.tm_self, x_full_bar, kb * BW_BK, i0);
      int it = 0;
      auto acquire = [&](uint32_t bytes) -> int {
        const int s = it % BW_STAGES;
        const uint32_t ph = (it / BW_STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(s), bytes);
        ++it;
        return s;
      };
      auto load_s = [&](int t) {  // 64 "other" rows of logit half-tile t, two K blocks per stage
        const int tt = t_begin + t;
        const BwdSegDev& sg = J.seg[tt / P.n_jtiles];
        const int j0 = (tt % P.n_jtiles) * BW_BN + static_cast<int>(h) * 64;
        for (int st = 0; st < n_sstage; ++st) {
          const int nk = (2 * st + 1 < num_kb) ? 2 : 1;
          const int s = acquire(nk * (BW_KB_BYTES / 2));
          for (int u = 0; u < nk; ++u)
            tma_load_2d(ring + s * BW_KB_BYTES + u * (BW_KB_BYTES / 2), &sg.tm_other, full_bar(s),
                        (2 * st + u) * BW_BK, j0);
        }
      };
      auto load_a = [&](int t) {  // operands of the gradient MMAs of tile t (this CTA's dim half)
        const int tt = t_begin + t;
        const BwdSegDev& sg = J.seg[tt / P.n_jtiles];
        const int j0 = (tt % P.n_jtiles) * BW_BN;
        for (int kb2 = 0; kb2 < 2; ++kb2)
          for (int dc = 0; dc < n_dc; ++dc) {
            const int s = acquire(BW_KB_BYTES);
            tma_load_2d(ring + s * BW_KB_BYTES, &sg.tm_other_t, full_bar(s), j0 + kb2 * BW_BK, d0 + dc * 128);
          }
      };
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
        for (int st = 0; st < n_sstage; ++st, ++it) {
          const int s = it % BW_STAGES;
          const uint32_t ph = (it / BW_STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const int nk = (2 * st + 1 < num_kb) ? 2 : 1;
          for (int u = 0; u < nk; ++u) {
            const int kb = 2 * st + u;
            const uint64_t ad = umma_desc_k_sw128(x_smem + kb * BW_KB_BYTES);
            const uint64_t bd = umma_desc_k_sw128(ring + s * BW_KB_BYTES + u * (BW_KB_BYTES / 2));
#pragma unroll
            for (int kk = 0; kk < BW_BK / 16; ++kk)
              tc_mma_f16(tmem + b * 64, ad + 2 * kk, bd + 2 * kk, P.idesc_n64, (kb | kk) != 0);
          }
          tc_commit(empty_bar(s));
        }
        tc_commit(s_full_bar(b));
      };
      issue_s(0);
      for (int t = 0; t < n_tiles; ++t) {
        if (t + 1 < n_tiles) issue_s(t + 1);
        mbar_wait_cluster(g_full_bar(0), t & 1);
        mbar_wait_cluster(g_full_bar(1), t & 1);
        fence_proxy_async_smem();
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
              tc_mma_f16(tmem_acc + dc * 128, ad + 2 * kk, bd + 2 * kk, P.idesc, (t | kb2 | kk) != 0);
            tc_commit(empty_bar(s));
          }
        }
        tc_commit_multicast(g_empty_bar, 0x3);  // G of tile t consumed here: tell both CTAs
      }
      tc_commit(acc_full_bar);
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int q = warp & 3;          // TMEM lane quarter
    const int ch = (warp - 2) >> 2;  // 32-column half of this CTA's 64 logit columns
    const int r = q * 32 + lane;     // tile-local row == TMEM lane
    const int et = threadIdx.x - 64;  // 0..255
    const int grow = i0 + r;
    float gs[2] = {0.f, 0.f};
    float gmax = 0.f;
    for (int s = 0; s < J.n_seg; ++s) {
      gs[s] = J.seg[s].grad_scale ? *J.seg[s].grad_scale : 1.f;
      gmax = fmaxf(gmax, fabsf(gs[s]));
    }
    const float inv_gmax = gmax > 0.f ? kGScale / gmax : 0.f;  // G in [-kGScale, kGScale]: see ntxent_bwd.h
    if (blockIdx.x == 0 && blockIdx.y == 0 && et == 0) *J.scale_out = gmax * P.out_scale * (1.f / kGScale);

    auto load_bj = [&](int t) -> float {
      if (et >= 64 || t >= n_tiles) return 0.f;
      const int tt = t_begin + t;
      const BwdSegDev& sg = J.seg[tt / P.n_jtiles];
      const int j = (tt % P.n_jtiles) * BW_BN + static_cast<int>(h) * 64 + et;
      return j < P.n_other ? ex2_approx(P.c1 - sg.lse2_other[j]) : 0.f;
    };
    if (et < 64 && n_tiles > 0) bj[et] = load_bj(0);
    int cur_seg = -1;
    float lse_i = 0.f, ws = 0.f, wo_i = 0.f, rr = 0.f;
    // this thread's 64 bytes of the G half-tile: K block h, row r, 16-byte chunks ch*4 .. ch*4+3
    const uint32_t g_row = g_smem + h * BW_KB_BYTES + r * 128;
    const uint32_t g_row_peer = map_to_peer(g_row, peer);
    const uint32_t g_full_local = g_full_bar(static_cast<int>(h));
    const uint32_t g_full_peer = map_to_peer(g_full_local, peer);

    for (int t = 0; t < n_tiles; ++t) {
      const int tt = t_begin + t;
      const int si = tt / P.n_jtiles;
      const int j0 = (tt % P.n_jtiles) * BW_BN + static_cast<int>(h) * 64;  // first column of our half
      const int b = t & 1;
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
      const float* bjt = bj + (t & 1) * 64 + ch * 32;
      const int dl = P.self_offset + grow - j0 - ch * 32;  // column of the positive inside our 32

      mbar_wait(s_full_bar(b), (t >> 1) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_addr(tmem + b * 64, q * 32, ch * 32), v);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_empty_bar(b));
      uint32_t pk[16];
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        const float p0 = ex2_approx(fmaf(__uint_as_float(v[e]), P.c1, -lse_i));
        const float p1 = ex2_approx(fmaf(__uint_as_float(v[e + 1]), P.c1, -lse_i));
        const float g0 = fmaf(p0, fmaf(wo_i, bjt[e], ws), (e == dl) ? -rr : 0.f);
        const float g1 = fmaf(p1, fmaf(wo_i, bjt[e + 1], ws), (e + 1 == dl) ? -rr : 0.f);
        pk[e >> 1] = pack2<kOp>(g0, g1);
      }
      // both CTAs have finished the gradient MMAs that read the previous G
      mbar_wait_cluster(g_empty_bar, (t & 1) ^ 1);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const uint32_t off = static_cast<uint32_t>(((ch * 4 + c4) ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(base_ptr + (g_row - base) + off) =
            make_uint4(pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
        st_cluster_v4(g_row_peer + off, pk[4 * c4], pk[4 * c4 + 1], pk[4 * c4 + 2], pk[4 * c4 + 3]);
      }
      fence_proxy_async_all();  // generic-proxy stores (local and DSMEM) -> async proxy (tcgen05 operand reads)
      __syncwarp();
      if (lane == 0) {  // one arrive per warp and per CTA: 8 remote arrives per tile instead of 256
        mbar_arrive(g_full_local);
        mbar_arrive_remote(g_full_peer);
      }
      if (et < 64 && t + 1 < n_tiles) bj[((t + 1) & 1) * 64 + et] = bj_next;
    }

    if (n_tiles > 0) {
      mbar_wait(acc_full_bar, 0);
      tc_fence_after();
    }
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
        if (grow < P.n_self) {
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
  cluster_sync_all();  // no CTA exits while its peer may still store / arrive into it
  if (warp == 1) tmem_dealloc(tmem, 512);
}

int launch_bwd_cluster(const BwdParams& P, int n_iblocks, int n_jobs, int op_format, cudaStream_t st) {
  const int smem = static_cast<int>(BwdSmem::total(P.num_kb));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_iblocks, 2 * P.n_split, n_jobs);
  cfg.blockDim = dim3(BW_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 2;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (op_format == TCL_OP_F16) {
    static int set = 0;
    if (set < smem) {
      TCL_CHECK_CUDA(cudaFuncSetAttribute(ntxent_bwd_cluster_kernel<TCL_OP_F16>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      set = smem;
    }
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ntxent_bwd_cluster_kernel<TCL_OP_F16>, P));
  } else {
    static int set = 0;
    if (set < smem) {
      TCL_CHECK_CUDA(cudaFuncSetAttribute(ntxent_bwd_cluster_kernel<TCL_OP_BF16>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
      set = smem;
    }
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ntxent_bwd_cluster_kernel<TCL_OP_BF16>, P));
  }
  return TCL_OK;
}

}  // namespace tcl
