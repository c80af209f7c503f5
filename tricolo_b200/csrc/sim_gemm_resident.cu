// K2' (main form) — retrieval similarity GEMM with the query block resident in shared memory.
//
// S = Q Gᵀ has a short K (dim <= 512): a 128x128 tile needs 256 KB of operands for 64 KB of
// output, so re-streaming both operands per tile is L2-bound.  Here a CTA keeps its 128 query
// rows (dim/64 swizzled 16 KB K-blocks) resident and sweeps a range of 128-row gallery tiles
// that stream through a TMA ring; two TMEM accumulators alternate so that the epilogue of tile t
// (TMEM -> registers -> full-sector streaming stores) overlaps the MMAs of tile t+1.
//
// Measured on B200 the single-CTA form is bound by L2 throughput (12.8 GB of gallery reads + 6.5 GB
// of result writes per 8192 x 200k block, ~11 TB/s).  Two CTAs with adjacent query blocks therefore
// form a cluster: each loads HALF of every gallery K-block (64 rows) and TMA-multicasts it into both
// CTAs' rings, halving the L2 read traffic.  A ring slot is released by tcgen05.commit multicast to
// both CTAs (empty barrier count 2), since each producer writes into both shared memories.
//
// Grid order: query block fastest, so the CTAs running at the same time sweep the same gallery range
// and every gallery tile is fetched from HBM about once.
// Warp roles: 0 TMA producer, 1 MMA issuer, 2..9 epilogue (two per TMEM lane quarter, 64 columns each).
#include "host_common.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

static constexpr int SR_BM = 128, SR_BN = 128, SR_BK = 64;
static constexpr int SR_KB_BYTES = SR_BM * SR_BK * 2;  // 16 KB
static constexpr int SR_STAGES = 6;  // 96 KB in flight: the gallery stream is latency-bound with fewer
static constexpr int SR_EPI_WARPS = 8;
static constexpr int SR_THREADS = 64 + SR_EPI_WARPS * 32;

struct SimResParams {
  CUtensorMap tm_q;  // [n_q, dim] box {64, 128}
  CUtensorMap tm_g;  // [n_g, dim] box {64, 128 / cluster size}
  float* s;
  int64_t ld;
  int n_q, n_g, num_kb, n_gtiles, n_split;
  uint32_t idesc;
};

struct SimResSmem {
  static constexpr uint32_t ring_off(int num_kb) { return num_kb * SR_KB_BYTES; }
  static constexpr uint32_t bar_off(int num_kb) { return ring_off(num_kb) + SR_STAGES * SR_KB_BYTES; }
  static constexpr uint32_t total(int num_kb) { return bar_off(num_kb) + 256 + 1024; }
};

template <int kCluster>
__global__ void __launch_bounds__(SR_THREADS, 1) sim_gemm_resident_kernel(const __grid_constant__ SimResParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const int num_kb = P.num_kb;
  const uint32_t q_smem = base;
  const uint32_t ring = base + SimResSmem::ring_off(num_kb);
  const uint32_t bars = base + SimResSmem::bar_off(num_kb);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (SR_STAGES + s); };
  const uint32_t q_full_bar = bars + 8u * (2 * SR_STAGES);
  auto tmem_full_bar = [&](int b) { return bars + 8u * (2 * SR_STAGES + 1 + b); };
  auto tmem_empty_bar = [&](int b) { return bars + 8u * (2 * SR_STAGES + 3 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * SR_STAGES + 5);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + SimResSmem::bar_off(num_kb) + 8u * (2 * SR_STAGES + 5));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = kCluster > 1 ? cluster_ctarank() : 0u;
  const int m0 = blockIdx.x * SR_BM;
  const int sp = blockIdx.y;
  const int t_begin = static_cast<int>((static_cast<int64_t>(P.n_gtiles) * sp) / P.n_split);
  const int t_end = static_cast<int>((static_cast<int64_t>(P.n_gtiles) * (sp + 1)) / P.n_split);
  const int n_tiles = t_end - t_begin;  // identical for all CTAs of a cluster (same split)
  constexpr uint16_t kMask = static_cast<uint16_t>((1u << kCluster) - 1u);
  constexpr int kSliceRows = SR_BN / kCluster;  // gallery rows this CTA fetches per K-block
  constexpr int kSliceBytes = SR_KB_BYTES / kCluster;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&P.tm_q);
    tma_prefetch_desc(&P.tm_g);
    for (int s = 0; s < SR_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kCluster);  // one tcgen05.commit per CTA of the cluster
    }
    mbar_init(q_full_bar, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), SR_EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();  // peers' barriers exist before any multicast lands
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 0) {
    if (elect_one() && n_tiles > 0) {
      mbar_arrive_expect_tx(q_full_bar, num_kb * SR_KB_BYTES);
      for (int kb = 0; kb < num_kb; ++kb)
        tma_load_2d(q_smem + kb * SR_KB_BYTES, &P.tm_q, q_full_bar, kb * SR_BK, m0);
      int it = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int n0 = (t_begin + t) * SR_BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % SR_STAGES;
          const uint32_t ph = (it / SR_STAGES) & 1;
          if (kCluster > 1) {
            mbar_wait_cluster(empty_bar(s), ph ^ 1);          // every CTA of the cluster has consumed the slot
            mbar_arrive_expect_tx(full_bar(s), SR_KB_BYTES);  // own slice + the peers' multicast slices
            tma_load_2d_multicast(ring + s * SR_KB_BYTES + crank * kSliceBytes, &P.tm_g, full_bar(s), kb * SR_BK,
                                  n0 + static_cast<int>(crank) * kSliceRows, kMask);
          } else {
            mbar_wait(empty_bar(s), ph ^ 1);
            mbar_arrive_expect_tx(full_bar(s), SR_KB_BYTES);
            tma_load_2d(ring + s * SR_KB_BYTES, &P.tm_g, full_bar(s), kb * SR_BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one() && n_tiles > 0) {
      mbar_wait(q_full_bar, 0);
      int it = 0;
      for (int t = 0; t < n_tiles; ++t) {
        const int b = t & 1;
        mbar_wait(tmem_empty_bar(b), ((t >> 1) & 1) ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % SR_STAGES;
          const uint32_t ph = (it / SR_STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint64_t ad = umma_desc_k_sw128(q_smem + kb * SR_KB_BYTES);
          const uint64_t bd = umma_desc_k_sw128(ring + s * SR_KB_BYTES);
#pragma unroll
          for (int kk = 0; kk < SR_BK / 16; ++kk)
            tc_mma_f16(tmem + b * SR_BN, ad + 2 * kk, bd + 2 * kk, P.idesc, (kb | kk) != 0);
          if (kCluster > 1)
            tc_commit_multicast(empty_bar(s), kMask);
          else
            tc_commit(empty_bar(s));
        }
        tc_commit(tmem_full_bar(b));
      }
    }
  } else {
    // Quad TMEM load layout (16x256b): thread (ql = lane/4, p = lane%4) holds two adjacent columns of rows
    // ql and ql+8 per 8-column group, so one store instruction writes eight full 32-byte sectors and the
    // four groups of a chunk complete 128-byte lines back to back.  No shared-memory staging.
    const int q = warp & 3;          // TMEM lane quarter (rows q*32 .. q*32+31 of the tile)
    const int ch = (warp - 2) >> 2;  // column half (64 columns)
    const int ql = lane >> 2, p = lane & 3;
    // this thread's four output rows are fixed for the whole sweep: row pointers are formed once
    float* rowp[2][2];
    bool rowv[2][2];
    bool rows_ok = true;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int grow = m0 + q * 32 + h * 16 + ql + rr * 8;
        rowv[h][rr] = grow < P.n_q;
        rows_ok &= rowv[h][rr];
        rowp[h][rr] = P.s + static_cast<int64_t>(grow) * P.ld + ch * 64 + 2 * p;
      }
    rows_ok = __all_sync(0xffffffffu, rows_ok);
    for (int t = 0; t < n_tiles; ++t) {
      const int b = t & 1;
      const int n0 = (t_begin + t) * SR_BN;
      mbar_wait(tmem_full_bar(b), (t >> 1) & 1);
      tc_fence_after();
      uint32_t v[2][2][16];  // [row half h][column chunk cl]
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int cl = 0; cl < 2; ++cl)
          tmem_ld_16x256b_x4(tmem_addr(tmem + b * SR_BN, q * 32 + h * 16, (2 * ch + cl) * 32), v[h][cl]);
      tc_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(b));  // accumulator is in registers
      if (rows_ok && n0 + SR_BN <= P.n_g) {
        // interior tile: 64 unguarded 8-byte streaming stores with immediate offsets
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            float* dst = rowp[h][rr] + n0;
#pragma unroll
            for (int cl = 0; cl < 2; ++cl)
#pragma unroll
              for (int g = 0; g < 4; ++g)
                __stcs(reinterpret_cast<float2*>(dst + cl * 32 + g * 8),
                       make_float2(__uint_as_float(v[h][cl][4 * g + 2 * rr]),
                                   __uint_as_float(v[h][cl][4 * g + 2 * rr + 1])));
          }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            if (!rowv[h][rr]) continue;
            float* dst = rowp[h][rr] + n0;
            const int c0 = n0 + ch * 64 + 2 * p;  // global column of dst[0]
#pragma unroll
            for (int cl = 0; cl < 2; ++cl)
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const int gcol = c0 + cl * 32 + g * 8;
                const float x0 = __uint_as_float(v[h][cl][4 * g + 2 * rr]);
                const float x1 = __uint_as_float(v[h][cl][4 * g + 2 * rr + 1]);
                if (gcol + 2 <= P.n_g)
                  *reinterpret_cast<float2*>(dst + cl * 32 + g * 8) = make_float2(x0, x1);
                else if (gcol < P.n_g)
                  dst[cl * 32 + g * 8] = x0;
              }
          }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();  // no CTA exits while a peer may still multicast into it
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// pick the gallery split: whole waves of 148 CTAs, >= 8 tiles per CTA when the problem is large
static int sim_split(int n_mblocks, int n_gtiles) {
  const int min_tiles = n_mblocks >= kNumSMsB200 / 2 ? 8 : 1;
  int max_split = n_gtiles / min_tiles > 0 ? n_gtiles / min_tiles : 1;
  if (max_split > 256) max_split = 256;
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= max_split; ++s) {
    const long long total = static_cast<long long>(n_mblocks) * s;
    const long long waves = (total + kNumSMsB200 - 1) / kNumSMsB200;
    const double eff = static_cast<double>(total) / (static_cast<double>(waves) * kNumSMsB200);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = s;
    }
  }
  return best;
}

int launch_sim_gemm_resident(const void* q, const void* g, int64_t n_q, int64_t n_g, int64_t dim, int op_format,
                             float* s, int64_t ld_s, cudaStream_t st) {
  SimResParams P;
  memset(&P, 0, sizeof(P));
  int n_mblocks = static_cast<int>((n_q + SR_BM - 1) / SR_BM);
  // two query blocks per cluster share the gallery stream; a single block runs the plain kernel
  const bool cluster = n_mblocks >= 2;
  if (cluster) n_mblocks = (n_mblocks + 1) / 2 * 2;  // padded block: rows >= n_q, loads zero-filled, stores masked
  if (int e = make_tmap_2d_16bit(&P.tm_q, q, n_q, dim, dim, SR_BM, SR_BK)) return e;
  if (int e = make_tmap_2d_16bit(&P.tm_g, g, n_g, dim, dim, cluster ? SR_BN / 2 : SR_BN, SR_BK)) return e;
  P.s = s;
  P.ld = ld_s;
  P.n_q = static_cast<int>(n_q);
  P.n_g = static_cast<int>(n_g);
  P.num_kb = static_cast<int>(dim / 64);
  P.n_gtiles = static_cast<int>((n_g + SR_BN - 1) / SR_BN);
  P.n_split = sim_split(n_mblocks, P.n_gtiles);
  P.idesc = umma_idesc_f16(SR_BM, SR_BN, op_format);
  TCL_REQUIRE(P.n_split <= 65535, TCL_ERR_BAD_SHAPE, "sim_gemm: split out of range");
  const int smem = static_cast<int>(SimResSmem::total(P.num_kb));
  ProfScope prof(TCL_K_SIM_GEMM, st);
  if (cluster) {
    if (int e = ensure_dyn_smem(sim_gemm_resident_kernel<2>, smem)) return e;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(n_mblocks, P.n_split);
    cfg.blockDim = dim3(SR_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, sim_gemm_resident_kernel<2>, P));
  } else {
    if (int e = ensure_dyn_smem(sim_gemm_resident_kernel<1>, smem)) return e;
    sim_gemm_resident_kernel<1><<<dim3(n_mblocks, P.n_split), SR_THREADS, smem, st>>>(P);
    TCL_CHECK_CUDA(cudaGetLastError());
  }
  return TCL_OK;
}

}  // namespace tcl
