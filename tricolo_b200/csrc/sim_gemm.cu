// K2' — similarity GEMM for retrieval: S[q, g] = sum_d Q[q,d] G[g,d]
// (raw dot product, tricolo/evaluation/eval_retrieval.py:74), 16-bit operands,
// fp32 accumulation in TMEM, fp32 result written to HBM for the top-k kernel.
//
// One 128x128 output tile per CTA, K streamed in 64-element blocks through a
// TMA -> shared-memory ring; one thread issues tcgen05.mma, four warps drain
// TMEM.  Two CTAs fit per SM (96 KB shared memory, 128 TMEM columns each) so
// one CTA's epilogue overlaps the other's MMA.
#include "host_common.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

static constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 64;
static constexpr int SG_STAGES = 3;
static constexpr int SG_TILE_BYTES = SG_BM * SG_BK * 2;  // 16 KB per operand per stage
static constexpr int SG_SMEM_BYTES = SG_STAGES * 2 * SG_TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

__global__ void __launch_bounds__(192, 1)
sim_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                float* __restrict__ c, int64_t ldc, int m, int n, int k, uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t a_smem = base;
  const uint32_t b_smem = base + SG_STAGES * SG_TILE_BYTES;
  const uint32_t bar_base = base + 2 * SG_STAGES * SG_TILE_BYTES;
  // barriers: full[S], empty[S], tmem_full ; then the TMEM base address
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (SG_STAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * SG_STAGES);
  const uint32_t tmem_slot = bar_base + 8u * (2 * SG_STAGES + 1);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * SG_BM;
  const int n0 = blockIdx.x * SG_BN;
  const int num_kb = (k + SG_BK - 1) / SG_BK;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
    for (int s = 0; s < SG_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp == 0) {
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % SG_STAGES;
        const uint32_t ph = (kb / SG_STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_arrive_expect_tx(full_bar(s), 2 * SG_TILE_BYTES);
        tma_load_2d(a_smem + s * SG_TILE_BYTES, &tm_a, full_bar(s), kb * SG_BK, m0);
        tma_load_2d(b_smem + s * SG_TILE_BYTES, &tm_b, full_bar(s), kb * SG_BK, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % SG_STAGES;
        const uint32_t ph = (kb / SG_STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint64_t ad = umma_desc_k_sw128(a_smem + s * SG_TILE_BYTES);
        const uint64_t bd = umma_desc_k_sw128(b_smem + s * SG_TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < SG_BK / 16; ++kk)
          tc_mma_f16(tmem, ad + 2 * kk, bd + 2 * kk, idesc, (kb | kk) != 0);
        tc_commit(empty_bar(s));
      }
      tc_commit(tmem_full_bar);
    }
  } else {
    // epilogue warps 2..5 -> TMEM lane quarter (warp % 4)
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    // All TMA loads have landed and all MMAs have read them (tmem_full): the operand ring is dead and
    // is reused as a per-warp 32x32 fp32 staging tile (row stride 36 floats: conflict-free both ways)
    // so that every global store instruction writes four full 128-byte lines.
    float* stage = reinterpret_cast<float*>(smem_raw + (base - raw)) + q * (32 * 36);
    const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
#pragma unroll 1
    for (int cc = 0; cc < SG_BN / 32; ++cc) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_addr(tmem, q * 32, cc * 32), v);
      tc_wait_ld();
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4)
        *reinterpret_cast<float4*>(stage + lane * 36 + c4 * 4) =
            make_float4(__uint_as_float(v[4 * c4]), __uint_as_float(v[4 * c4 + 1]),
                        __uint_as_float(v[4 * c4 + 2]), __uint_as_float(v[4 * c4 + 3]));
      __syncwarp();
      const int gcol = n0 + cc * 32 + sub_c;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + sub_r;
        const int grow = m0 + q * 32 + r;
        const float4 x = *reinterpret_cast<const float4*>(stage + r * 36 + sub_c);
        if (grow < m) {
          float* dst = c + static_cast<int64_t>(grow) * ldc + gcol;
          if (gcol + 4 <= n) {
            *reinterpret_cast<float4*>(dst) = x;
          } else {
            if (gcol < n) dst[0] = x.x;
            if (gcol + 1 < n) dst[1] = x.y;
            if (gcol + 2 < n) dst[2] = x.z;
          }
        }
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 128);
}

// ---------------------------------------------------------------------------
// bring-up probe for the tcgen05.ld 16x256b register layout
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) tmem_probe_kernel(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    tmem_alloc(smem_u32(&slot), 32);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  for (int col = 0; col < 32; ++col)
    tmem_st_32x32b_x1(tmem_addr(tmem, warp * 32, col), static_cast<uint32_t>((warp * 32 + lane) * 64 + col));
  tc_wait_st();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int half = 0; half < 2; ++half) {
    uint32_t v[16];
    tmem_ld_16x256b_x4(tmem_addr(tmem, warp * 32 + half * 16, 0), v);
    tc_wait_ld();
    for (int r = 0; r < 16; ++r) out[((warp * 2 + half) * 32 + lane) * 16 + r] = v[r];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

}  // namespace tcl

using namespace tcl;

extern "C" int tcl_sim_gemm(const void* q, const void* g, int64_t n_q, int64_t n_g, int64_t dim,
                            int op_format, float* s, int64_t ld_s, void* stream) {
  TCL_REQUIRE(q && g && s, TCL_ERR_BAD_ARG, "sim_gemm: null pointer");
  TCL_REQUIRE(n_q >= 0 && n_g >= 1 && n_q < (1LL << 31) - 256 && n_g < (1LL << 31) - 256,
              TCL_ERR_BAD_SHAPE, "sim_gemm: bad sizes");
  TCL_REQUIRE(dim >= 8 && dim % 8 == 0, TCL_ERR_BAD_SHAPE, "sim_gemm: dim must be a positive multiple of 8 (got %lld)", (long long)dim);
  TCL_REQUIRE(ld_s >= n_g && ld_s % 4 == 0 && aligned_to(s, 16), TCL_ERR_BAD_ALIGN, "sim_gemm: output must be 16-byte aligned with ld_s %% 4 == 0");
  TCL_REQUIRE(op_format == TCL_OP_F16 || op_format == TCL_OP_BF16, TCL_ERR_BAD_ARG, "op_format %d", op_format);
  if (int e = require_sm100()) return e;
  if (n_q == 0) return TCL_OK;
  // main form: query block resident in shared memory (sim_gemm_resident.cu); the tile-per-CTA
  // kernel below covers dims the resident layout cannot hold.
  if (dim % 64 == 0 && dim <= 512)
    return launch_sim_gemm_resident(q, g, n_q, n_g, dim, op_format, s, ld_s, static_cast<cudaStream_t>(stream));
  CUtensorMap tm_a, tm_b;
  if (int e = make_tmap_2d_16bit(&tm_a, q, n_q, dim, dim, SG_BM, SG_BK)) return e;
  if (int e = make_tmap_2d_16bit(&tm_b, g, n_g, dim, dim, SG_BN, SG_BK)) return e;
  if (int e = ensure_dyn_smem(sim_gemm_kernel, SG_SMEM_BYTES)) return e;
  dim3 grid(static_cast<unsigned>((n_g + SG_BN - 1) / SG_BN), static_cast<unsigned>((n_q + SG_BM - 1) / SG_BM));
  TCL_REQUIRE(grid.y <= 65535, TCL_ERR_BAD_SHAPE, "sim_gemm: more than 65535*128 query rows per call; chunk the queries");
  {
    ProfScope prof(TCL_K_SIM_GEMM, static_cast<cudaStream_t>(stream));
    sim_gemm_kernel<<<grid, 192, SG_SMEM_BYTES, static_cast<cudaStream_t>(stream)>>>(
        tm_a, tm_b, s, ld_s, (int)n_q, (int)n_g, (int)dim, umma_idesc_f16(SG_BM, SG_BN, op_format));
  }
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

extern "C" int tcl_debug_tmem_probe(uint32_t* out, void* stream) {
  TCL_REQUIRE(out != nullptr, TCL_ERR_BAD_ARG, "probe: null pointer");
  if (int e = require_sm100()) return e;
  tmem_probe_kernel<<<1, 128, 0, static_cast<cudaStream_t>(stream)>>>(out);
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}
