// Normalise backward folded into the gradient GEMM kernels' accumulator read-out.
//
// The separate l2norm_bwd_kernel re-reads every fp32 partial the GEMM kernel has just stored (plus x) in a launch of its
// own: 43 us of the 693 us step at B = 8192, a whole launch of the five at B = 256.  Here the read-out warps that store
// the LAST piece of a 128-row block finish the block themselves: a counter per (job, row block) is bumped after each
// piece's TMA stores have completed; whoever brings it to the block's piece count (known from the tile table) sums the
// partials in slot order (the same fixed order as the separate kernel: bit-identical), applies
//     dx = clamped ? g * inv : (g - (g.z) z) * inv,   g = scale * sum of partials,  z = x * inv
// (F.normalize backward, tricolo/loss/nt_xent.py:56-57) and writes dx.  The accumulator has been handed back to the MMA
// thread before, so this work overlaps the next piece's MMAs; the partials come from L2.
// MEASURED (B = 8192, one GPU, CTA-pair kernel): kernel B 297 -> 355 us against the 42 us of the kernel it replaces, the
// step 0.688 -> 0.707 ms: a row block is finished by ONE CTA's eight warps, a row at a time, while the next piece's
// accumulator waits for the same warps.  Kept opt-in (TRICOLO_B200_FOLD=1) and tested; not the default.
// Not used when column-side partials come from other ranks (sharded shared-G form: tcl_ntxent_bwd_sharded_finish).
#pragma once
#include "common.cuh"

namespace tcl {

static constexpr int FOLD_MAX_JOBS = 6;
struct FoldJob {
  const void* x;          // [rows, dim] inputs, x_dtype, row stride x_stride
  void* dx;               // [rows, dim] output, x_dtype, contiguous
  const float* inv_norm;  // [rows]
  const float* gpart;     // [slots][slot_rows][dim] fp32 partials of this job
  const float* scale;     // device scalar written by the kernel that formed G
};
struct FoldParams {
  FoldJob job[FOLD_MAX_JOBS];
  uint32_t* counters;     // [FOLD_MAX_JOBS][n_rowblocks], zero before the launch; left at zero by the kernel
  int64_t x_stride;       // elements
  int64_t slot_stride;    // elements between partial slots (slot_rows * dim)
  int enabled, rows, dim, x_dtype, n_rowblocks;
  float eps;
};

__device__ __forceinline__ void fold_load4(const void* p, int dtype, int64_t idx, float (&v)[4]) {
  if (dtype == TCL_DT_F32) {
    const float4 t = *reinterpret_cast<const float4*>(static_cast<const float*>(p) + idx);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else if (dtype == TCL_DT_F16) {
    const uint2 t = *reinterpret_cast<const uint2*>(static_cast<const __half*>(p) + idx);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else if (dtype == TCL_DT_BF16) {
    const uint2 t = *reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(p) + idx);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  } else {
    const double2 a = *reinterpret_cast<const double2*>(static_cast<const double*>(p) + idx);
    const double2 b = *reinterpret_cast<const double2*>(static_cast<const double*>(p) + idx + 2);
    v[0] = static_cast<float>(a.x); v[1] = static_cast<float>(a.y); v[2] = static_cast<float>(b.x); v[3] = static_cast<float>(b.y);
  }
}
__device__ __forceinline__ void fold_store4(void* p, int dtype, int64_t idx, const float (&v)[4]) {
  if (dtype == TCL_DT_F32) {
    *reinterpret_cast<float4*>(static_cast<float*>(p) + idx) = make_float4(v[0], v[1], v[2], v[3]);
  } else if (dtype == TCL_DT_F16) {
    uint2 t;
    *reinterpret_cast<__half2*>(&t.x) = __floats2half2_rn(v[0], v[1]);
    *reinterpret_cast<__half2*>(&t.y) = __floats2half2_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(static_cast<__half*>(p) + idx) = t;
  } else if (dtype == TCL_DT_BF16) {
    uint2 t;
    *reinterpret_cast<__nv_bfloat162*>(&t.x) = __floats2bfloat162_rn(v[0], v[1]);
    *reinterpret_cast<__nv_bfloat162*>(&t.y) = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(p) + idx) = t;
  } else {
    *reinterpret_cast<double2*>(static_cast<double*>(p) + idx) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(static_cast<double*>(p) + idx + 2) = make_double2(v[2], v[3]);
  }
}

// Normalise backward of rows [ib * 128 + dw * 16, +16) of one job (dw = index of the read-out warp, 0..7): one row at a
// time per warp, the row's x and up to three partials in flight before any is consumed (as l2norm_bwd_kernel).
__device__ __forceinline__ void fold_rows(const FoldParams& F, int job, int ib, int n_pieces, int dw, int lane) {
  const FoldJob& J = F.job[job];
  const float scale = __ldcg(J.scale);
  for (int rr = 0; rr < 16; ++rr) {
    const int64_t row = static_cast<int64_t>(ib) * 128 + dw * 16 + rr;
    if (row >= F.rows) break;
    const float inv = J.inv_norm[row];
    const bool clamped = inv >= 1.f / F.eps;
    float g[4][4], z[4][4];
    float dot = 0.f;
    float4 part[3][4];
    float xv[4][4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int c = it * 128 + lane * 4;
#pragma unroll
      for (int s = 0; s < 3; ++s)
        part[s][it] = (c < F.dim && s < n_pieces) ? __ldcg(reinterpret_cast<const float4*>(J.gpart + s * F.slot_stride + row * F.dim + c))
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < F.dim) fold_load4(J.x, F.x_dtype, row * F.x_stride + c, xv[it]);
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < F.dim) {
        // slot order 0, 1, 2, ...: the summation order of the separate kernel
        float acc[4] = {part[0][it].x + part[1][it].x + part[2][it].x, part[0][it].y + part[1][it].y + part[2][it].y,
                        part[0][it].z + part[1][it].z + part[2][it].z, part[0][it].w + part[1][it].w + part[2][it].w};
        for (int s = 3; s < n_pieces; ++s) {
          const float4 p = __ldcg(reinterpret_cast<const float4*>(J.gpart + s * F.slot_stride + row * F.dim + c));
          acc[0] += p.x; acc[1] += p.y; acc[2] += p.z; acc[3] += p.w;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          g[it][e] = acc[e] * scale;
          z[it][e] = xv[it][e] * inv;
          dot += g[it][e] * z[it][e];
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (clamped) dot = 0.f;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < F.dim) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] = (g[it][e] - dot * z[it][e]) * inv;
        fold_store4(J.dx, F.x_dtype, row * F.dim + c, o);
      }
    }
  }
}

// Called by ALL eight read-out warps of a CTA after they have issued the TMA stores of a piece of (job, row block ib).
// `flag` is a word of shared memory, bar_id a named barrier reserved for the 256 read-out threads.
__device__ __forceinline__ void fold_piece_done(const FoldParams& F, int job, int ib, int n_pieces, int dw, int lane,
                                                volatile uint32_t* flag, int bar_id) {
  if (lane == 0) {
    bulk_wait_all();              // this warp's partial stores are complete ...
    fence_proxy_async_generic();  // ... and ordered before the generic-proxy traffic below
    __threadfence();
  }
  asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
  if (dw == 0 && lane == 0) {
    uint32_t* cnt = F.counters + job * F.n_rowblocks + ib;
    __threadfence();  // release: the eight warps' stores (ordered to this thread by the barrier) before the count
    const bool last = atomicAdd(cnt, 1u) + 1u == static_cast<uint32_t>(n_pieces);
    if (last) *cnt = 0u;  // nobody touches it again in this launch
    __threadfence();
    *flag = last ? 1u : 0u;
  }
  asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
  if (*flag != 0u) fold_rows(F, job, ib, n_pieces, dw, lane);
  asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");  // the flag may be rewritten by the next piece
}

}  // namespace tcl
