// K1 (L2-normalise prologue), 16-bit cast, 16-bit transpose and the normalise
// backward.  All HBM-bound streaming kernels: one warp per row, 16-byte
// vector accesses, fp32 math.
//
// Reference semantics: torch.nn.functional.normalize(p=2, dim=1) as called at
// tricolo/loss/nt_xent.py:56-57, i.e. x / max(||x||_2, eps) with eps = 1e-12,
// and its autograd.
#include "host_common.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

struct PtrPack3 {
  const void* in[TCL_MAX_TENSORS];
  void* out[TCL_MAX_TENSORS];
  float* aux[TCL_MAX_TENSORS];
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&v)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<double>(const double* p, float (&v)[4]) {
  double2 a = *reinterpret_cast<const double2*>(p);
  double2 b = *reinterpret_cast<const double2*>(p + 2);
  v[0] = (float)a.x; v[1] = (float)a.y; v[2] = (float)b.x; v[3] = (float)b.y;
}
template <>
__device__ __forceinline__ void load4<__half>(const __half* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  float2 a = __half22float2(*reinterpret_cast<__half2*>(&t.x));
  float2 b = __half22float2(*reinterpret_cast<__half2*>(&t.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <>
__device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&t.y));
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

template <typename T>
__device__ __forceinline__ void store4(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void store4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<double>(double* p, const float (&v)[4]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
  *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}
template <>
__device__ __forceinline__ void store4<__half>(__half* p, const float (&v)[4]) {
  uint2 t;
  *reinterpret_cast<__half2*>(&t.x) = __floats2half2_rn(v[0], v[1]);
  *reinterpret_cast<__half2*>(&t.y) = __floats2half2_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 t;
  *reinterpret_cast<__nv_bfloat162*>(&t.x) = __floats2bfloat162_rn(v[0], v[1]);
  *reinterpret_cast<__nv_bfloat162*>(&t.y) = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------
// K1: one warp per row.  kNormalise=false gives the plain 16-bit cast.
// The row is read once into registers when dim <= 32*4*kMaxIter, else twice.
// ---------------------------------------------------------------------------
template <typename TIn, typename TOut, bool kNormalise>
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(PtrPack3 pk, int64_t rows, int dim,
                                                         int64_t x_stride, int64_t z_stride, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (row >= rows) return;
  const TIn* x = static_cast<const TIn*>(pk.in[blockIdx.y]) + row * x_stride;
  TOut* z = static_cast<TOut*>(pk.out[blockIdx.y]) + row * z_stride;
  constexpr int kMaxIter = 4;  // 512 elements stay in registers
  float v[kMaxIter][4];
  float ss = 0.f;
  const bool in_regs = dim <= 32 * 4 * kMaxIter;
  if (in_regs) {
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < dim) {
        load4<TIn>(x + c, v[it]);
        ss += v[it][0] * v[it][0] + v[it][1] * v[it][1] + v[it][2] * v[it][2] + v[it][3] * v[it][3];
      }
    }
  } else if (kNormalise) {
    for (int c = lane * 4; c < dim; c += 128) {
      float t[4];
      load4<TIn>(x + c, t);
      ss += t[0] * t[0] + t[1] * t[1] + t[2] * t[2] + t[3] * t[3];
    }
  }
  float inv = 1.f;
  if (kNormalise) {
    ss = warp_sum(ss);
    inv = 1.f / fmaxf(sqrtf(ss), eps);
    if (lane == 0) pk.aux[blockIdx.y][row] = inv;
  }
  if (in_regs) {
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < dim) {
        float o[4] = {v[it][0] * inv, v[it][1] * inv, v[it][2] * inv, v[it][3] * inv};
        store4<TOut>(z + c, o);
      }
    }
  } else {
    for (int c = lane * 4; c < dim; c += 128) {
      float t[4];
      load4<TIn>(x + c, t);
      float o[4] = {t[0] * inv, t[1] * inv, t[2] * inv, t[3] * inv};
      store4<TOut>(z + c, o);
    }
  }
}

// K1 + all-gather + flags (the sharded loss's first kernel; tricolo_b200/distributed.py, host_common.h: ShardSync).
//   * block (0,0) tells every peer "my operand buffer may be overwritten" (kReady, epoch e) - stream order guarantees
//     that this rank's readers of the previous step have finished;
//   * every warp normalises one row, stores it locally at once and to peer p as soon as p's kReady flag shows e;
//   * the last block of a 128-row chunk (all modalities) signals kArrived[rank][chunk] = e to every peer: the forward
//     tile kernel of the peer loads a column tile as soon as its chunk has landed (no barrier);
//   * the last block of the grid publishes the new epoch locally.
// Each lane handles 8 consecutive elements: 16-byte stores over NVLink.
struct PushPack {
  const void* in[TCL_MAX_TENSORS];
  void* out[TCL_MAX_PEERS][TCL_MAX_TENSORS];
  float* aux[TCL_MAX_TENSORS];
  uint32_t* sync[TCL_MAX_PEERS];  // sync pad of every rank as mapped here; sync[rank] is the own one
  int rank, world, n_tensors;
  int remote;  // 1: this kernel also stores the rows to the peers and flags the chunks; 0: the forward tile kernel's
               // push warps do that while its MMAs run (tcl_ntxent_fwd_sharded); 2: plain stores to `world` destinations,
               // no flags at all (tcl_l2norm_fwd_bcast: the caller brackets the launch with barriers)
};
template <typename T>
__device__ __forceinline__ uint4 pack8(const float (&o)[8]);
template <>
__device__ __forceinline__ uint4 pack8<__half>(const float (&o)[8]) {
  uint4 u;
  *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(o[0], o[1]);
  *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(o[2], o[3]);
  *reinterpret_cast<__half2*>(&u.z) = __floats2half2_rn(o[4], o[5]);
  *reinterpret_cast<__half2*>(&u.w) = __floats2half2_rn(o[6], o[7]);
  return u;
}
template <>
__device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float (&o)[8]) {
  uint4 u;
  *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(o[0], o[1]);
  *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(o[2], o[3]);
  *reinterpret_cast<__nv_bfloat162*>(&u.z) = __floats2bfloat162_rn(o[4], o[5]);
  *reinterpret_cast<__nv_bfloat162*>(&u.w) = __floats2bfloat162_rn(o[6], o[7]);
  return u;
}
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) l2norm_fwd_push_kernel(const __grid_constant__ PushPack pk, int64_t rows, int dim,
                                                              int64_t x_stride, int64_t z_stride, float eps) {
  griddep_launch();
  griddep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blocks in (row block, tensor) order, tensor fastest: the blocks of one 128-row chunk are scheduled together
  const int tensor = static_cast<int>(blockIdx.x % pk.n_tensors);
  const uint32_t rowblock = blockIdx.x / pk.n_tensors;
  const int64_t row = static_cast<int64_t>(rowblock) * 8 + warp;
  const bool flags = pk.remote != 2;
  uint32_t* sy = pk.sync[pk.rank];
  const uint32_t e = flags ? ld_relaxed_u32(sy + ShardSync::kFwdEpoch) + 1u : 0u;  // written only by the grid's LAST block
  if (flags && blockIdx.x == 0 && threadIdx.x < pk.world && static_cast<int>(threadIdx.x) != pk.rank)
    st_release_sys_u32(pk.sync[threadIdx.x] + ShardSync::kReady + pk.rank, e);
  if (row < rows) {
    const TIn* x = static_cast<const TIn*>(pk.in[tensor]) + row * x_stride;
    float v[2][8];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int c = it * 256 + lane * 8;
      if (c < dim) {
        float a[4], b[4];
        load4<TIn>(x + c, a);
        load4<TIn>(x + c + 4, b);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          v[it][k] = a[k];
          v[it][4 + k] = b[k];
          ss += a[k] * a[k] + b[k] * b[k];
        }
      }
    }
    ss = warp_sum(ss);
    const float inv = 1.f / fmaxf(sqrtf(ss), eps);
    if (lane == 0) pk.aux[tensor][row] = inv;
    uint4 o[2];
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      float t[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) t[k] = v[it][k] * inv;
      o[it] = pack8<TOut>(t);
    }
    const int n_dst = pk.remote ? pk.world : 1;
    for (int d = 0; d < n_dst; ++d) {
      const int p = (pk.rank + d) % pk.world;  // own buffer first, then the peers in a rank-staggered order
      if (d > 0 && flags) {
        if (lane == 0) flag_wait_ge(sy + ShardSync::kReady + p, e);
        __syncwarp();
      }
      if (pk.out[p][tensor] == nullptr) continue;  // this destination does not take this tensor (bcast form only)
      TOut* z = static_cast<TOut*>(pk.out[p][tensor]) + row * z_stride;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int c = it * 256 + lane * 8;
        if (c < dim) *reinterpret_cast<uint4*>(z + c) = o[it];
      }
    }
  }
  if (!flags) return;
  if (pk.remote) __threadfence_system();
  else __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (pk.remote) {
      const uint32_t chunk = rowblock >> 4;  // 16 row blocks of 8 rows, all tensors
      const int64_t rows_left = rows - static_cast<int64_t>(chunk) * 128;
      const uint32_t rows_in_chunk = static_cast<uint32_t>(rows_left < 128 ? rows_left : 128);
      const uint32_t per_chunk = ((rows_in_chunk + 7) / 8) * pk.n_tensors;
      if (atomicAdd(sy + ShardSync::kChunkCnt + chunk, 1u) + 1u == per_chunk) {  // last block of the chunk
        sy[ShardSync::kChunkCnt + chunk] = 0u;  // nobody touches the counter again before the next step
        __threadfence_system();
        for (int p = 0; p < pk.world; ++p)
          if (p != pk.rank) st_release_sys_u32(pk.sync[p] + ShardSync::kArrived + pk.rank * ShardSync::kMaxChunks + chunk, e);
      }
    }
    if (atomicAdd(sy + ShardSync::kK1Done, 1u) + 1u == gridDim.x) {
      sy[ShardSync::kK1Done] = 0u;
      __threadfence();
      *reinterpret_cast<volatile uint32_t*>(sy + ShardSync::kFwdEpoch) = e;
    }
  }
}

__global__ void __launch_bounds__(256) peer_sum_kernel(const float* const* __restrict__ /*unused*/, int n_src,
                                                       const float* s0, const float* s1, const float* s2, const float* s3,
                                                       const float* s4, const float* s5, const float* s6, const float* s7,
                                                       int64_t n4, int n_tail, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const float* src[8] = {s0, s1, s2, s3, s4, s5, s6, s7};
  if (i >= n4) {  // the n % 4 trailing elements (odd local batches at world 2): one thread each, same rank order
    const int64_t t = i - n4;
    if (t < n_tail) {
      float a = 0.f;
      for (int r = 0; r < n_src; ++r) {
        const float v = __ldcv(src[r] + 4 * n4 + t);
        a = r == 0 ? v : a + v;
      }
      out[4 * n4 + t] = a;
    }
    return;
  }
  float4 part[8];
#pragma unroll
  for (int r = 0; r < 8; ++r)
    part[r] = r < n_src ? __ldcv(reinterpret_cast<const float4*>(src[r]) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 a = part[0];
#pragma unroll
  for (int r = 1; r < 8; ++r) {
    if (r < n_src) { a.x += part[r].x; a.y += part[r].y; a.z += part[r].z; a.w += part[r].w; }
  }
  reinterpret_cast<float4*>(out)[i] = a;
}

template <typename TIn, bool kNormalise>
static int launch_fwd_t(const PtrPack3& pk, int n_tensors, int64_t rows, int dim, int64_t stride,
                        int64_t z_stride, int op_format, float eps, cudaStream_t st) {
  dim3 grid(static_cast<unsigned>((rows + 7) / 8), n_tensors);
  ProfScope prof(kNormalise ? TCL_K_L2NORM_FWD : TCL_K_CAST16, st);
  if (op_format == TCL_OP_F16)
    l2norm_fwd_kernel<TIn, __half, kNormalise><<<grid, 256, 0, st>>>(pk, rows, dim, stride, z_stride, eps);
  else
    l2norm_fwd_kernel<TIn, __nv_bfloat16, kNormalise>
        <<<grid, 256, 0, st>>>(pk, rows, dim, stride, z_stride, eps);
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

template <bool kNormalise>
static int launch_fwd(const PtrPack3& pk, int n_tensors, int x_dtype, int64_t rows, int dim,
                      int64_t stride, int64_t z_stride, int op_format, float eps, cudaStream_t st) {
  switch (x_dtype) {
    case TCL_DT_F32: return launch_fwd_t<float, kNormalise>(pk, n_tensors, rows, dim, stride, z_stride, op_format, eps, st);
    case TCL_DT_F64: return launch_fwd_t<double, kNormalise>(pk, n_tensors, rows, dim, stride, z_stride, op_format, eps, st);
    case TCL_DT_F16: return launch_fwd_t<__half, kNormalise>(pk, n_tensors, rows, dim, stride, z_stride, op_format, eps, st);
    case TCL_DT_BF16: return launch_fwd_t<__nv_bfloat16, kNormalise>(pk, n_tensors, rows, dim, stride, z_stride, op_format, eps, st);
  }
  return set_error(TCL_ERR_BAD_ARG, "unknown x_dtype %d", x_dtype);
}

// ---------------------------------------------------------------------------
// Gallery build: out16[g] = round16(src0[index[g]] (+ src1[index[g]])), one warp per output row
// ---------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) gather_sum_cast_kernel(PtrPack3 pk, int n_src, int64_t n_src_rows, int dim,
                                                              int64_t x_stride, const int64_t* __restrict__ index,
                                                              int64_t n_out, TOut* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t g = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (g >= n_out) return;
  const int64_t r = index[g];
  if (r < 0 || r >= n_src_rows) return;  // validated on the host side of the ABI's callers; never read out of range
  const TIn* a = static_cast<const TIn*>(pk.in[0]) + r * x_stride;
  const TIn* b = n_src > 1 ? static_cast<const TIn*>(pk.in[1]) + r * x_stride : nullptr;
  TOut* o = out + g * dim;
  for (int c = lane * 4; c < dim; c += 128) {
    float va[4], vb[4] = {0.f, 0.f, 0.f, 0.f};
    load4<TIn>(a + c, va);
    if (b) load4<TIn>(b + c, vb);
    // reference order: zeros, += image, += voxel  (0 + x is exact)
    float s[4] = {(0.f + va[0]) + vb[0], (0.f + va[1]) + vb[1], (0.f + va[2]) + vb[2], (0.f + va[3]) + vb[3]};
    if (!b) { s[0] = va[0]; s[1] = va[1]; s[2] = va[2]; s[3] = va[3]; }
    store4<TOut>(o + c, s);
  }
}

template <typename TIn>
static int launch_gather_sum(const PtrPack3& pk, int n_src, int64_t n_src_rows, int dim, int64_t stride,
                             const int64_t* index, int64_t n_out, void* out, int op_format, cudaStream_t st) {
  dim3 grid(static_cast<unsigned>((n_out + 7) / 8));
  ProfScope prof(TCL_K_GATHER_SUM, st);
  if (op_format == TCL_OP_F16)
    gather_sum_cast_kernel<TIn, __half><<<grid, 256, 0, st>>>(pk, n_src, n_src_rows, dim, stride, index, n_out,
                                                              static_cast<__half*>(out));
  else
    gather_sum_cast_kernel<TIn, __nv_bfloat16><<<grid, 256, 0, st>>>(pk, n_src, n_src_rows, dim, stride, index, n_out,
                                                                     static_cast<__nv_bfloat16*>(out));
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

static size_t dtype_size(int dt) {
  return dt == TCL_DT_F32 ? 4 : dt == TCL_DT_F64 ? 8 : 2;
}

// ---------------------------------------------------------------------------
// 16-bit transpose: [rows, dim] -> [dim, ld_t], 64x64 tiles through shared memory
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose16_kernel(PtrPack3 pk, int64_t rows, int dim,
                                                          int64_t z_stride, int64_t ld_t) {
  __shared__ uint16_t tile[64][66];
  const uint16_t* in = static_cast<const uint16_t*>(pk.in[blockIdx.z]);
  uint16_t* out = static_cast<uint16_t*>(pk.out[blockIdx.z]);
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int c0 = blockIdx.y * 64;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  // read: each thread reads 2 adjacent columns (4 bytes) of 8 rows
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = ty + i * 8;
    const int c = tx * 2;
    uint32_t w = 0;
    if (r0 + r < rows && c0 + c < dim)
      w = *reinterpret_cast<const uint32_t*>(in + (r0 + r) * z_stride + c0 + c);
    tile[r][c] = static_cast<uint16_t>(w & 0xffffu);
    tile[r][c + 1] = static_cast<uint16_t>(w >> 16);
  }
  __syncthreads();
  // write: out[c0 + c][r0 + r], two adjacent r per thread
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = ty + i * 8;
    const int r = tx * 2;
    if (c0 + c < dim && r0 + r < rows) {
      const uint32_t w = static_cast<uint32_t>(tile[r][c]) |
                         (static_cast<uint32_t>(r0 + r + 1 < rows ? tile[r + 1][c] : 0) << 16);
      uint16_t* dst = out + static_cast<int64_t>(c0 + c) * ld_t + r0 + r;
      if (r0 + r + 1 < ld_t)
        *reinterpret_cast<uint32_t*>(dst) = w;
      else
        *dst = static_cast<uint16_t>(w & 0xffffu);
    }
  }
}

// ---------------------------------------------------------------------------
// Normalise backward (+ reduction of the split-K partials of the gradient GEMM).
//   g  = scale * sum_s gpart[s][row][:]
//   z  = x * inv
//   dx = clamped ? g * inv : (g - (g.z) z) * inv
// "clamped" (||x|| < eps): the reference's clamp_min has zero derivative there, so
// the projection term vanishes.
// ---------------------------------------------------------------------------

template <typename T>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(NormBwdParams pr, int64_t rows, int dim,
                                                         int64_t x_stride, int n_split, float eps) {
  griddep_launch();
  griddep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (row >= rows) return;
  const NormBwdJob& jb = pr.job[blockIdx.y];
  const T* x = static_cast<const T*>(jb.x) + row * x_stride;
  T* dx = static_cast<T*>(jb.dx) + row * dim;
  const float inv = jb.inv_norm[row];
  const float scale = *jb.scale;
  const bool clamped = inv >= 1.f / eps;
  constexpr int kMaxIter = 4;
  float g[kMaxIter][4], z[kMaxIter][4];
  float dot = 0.f;
  const int64_t split_stride = (pr.split_rows ? pr.split_rows : rows) * static_cast<int64_t>(dim);
  if (pr.n_clusters > 0) {
    // persistent gradient kernel: one partial per tile range that touches this row's 128-row unit
    const int64_t u0 = pr.job_tile_base[blockIdx.y] + (row >> (pr.unit_shift ? pr.unit_shift : 7)) * pr.unit_tiles[blockIdx.y];
    n_split = pc_range_of(pr.total_tiles, u0 + pr.unit_tiles[blockIdx.y] - 1, pr.n_clusters) -
              pc_range_of(pr.total_tiles, u0, pr.n_clusters) + 1;
  }
  if (n_split <= 3) {
    // common case (1-3 partials per row): EVERY load of the row - x and all partials, up to 16 x 16 bytes per lane - is
    // issued before anything is consumed; with the loads inside the per-128-column loop the row cost four dependent
    // round trips to memory
    float4 part[3][kMaxIter];
    float xv[kMaxIter][4];
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
      const int c = it * 128 + lane * 4;
#pragma unroll
      for (int s = 0; s < 3; ++s)
        part[s][it] = (c < dim && s < n_split)
                          ? __ldcs(reinterpret_cast<const float4*>(jb.gpart + s * split_stride + row * dim + c))
                          : make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < dim) load4<T>(x + c, xv[it]);
    }
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < dim) {
        // same summation order as the general path: s = 0, 1, 2
        const float acc[4] = {part[0][it].x + part[1][it].x + part[2][it].x, part[0][it].y + part[1][it].y + part[2][it].y,
                              part[0][it].z + part[1][it].z + part[2][it].z, part[0][it].w + part[1][it].w + part[2][it].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          g[it][e] = acc[e] * scale;
          z[it][e] = xv[it][e] * inv;
          dot += g[it][e] * z[it][e];
        }
      }
    }
  } else {
#pragma unroll
  for (int it = 0; it < kMaxIter; ++it) {
    const int c = it * 128 + lane * 4;
    if (c < dim) {
      // all split partials are fetched before any is consumed (8 independent 16-byte loads in flight);
      // the sum order s = 0,1,2,... is fixed, so the result does not depend on the split count's timing
      float4 part[8];
#pragma unroll
      for (int s = 0; s < 8; ++s)
        part[s] = s < n_split ? __ldcs(reinterpret_cast<const float4*>(jb.gpart + s * split_stride + row * dim + c))
                              : make_float4(0.f, 0.f, 0.f, 0.f);
      float xv[4];
      load4<T>(x + c, xv);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        acc[0] += part[s].x; acc[1] += part[s].y; acc[2] += part[s].z; acc[3] += part[s].w;
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        g[it][e] = acc[e] * scale;
        z[it][e] = xv[e] * inv;
        dot += g[it][e] * z[it][e];
      }
    }
  }
  }
  dot = warp_sum(dot);
  if (clamped) dot = 0.f;
#pragma unroll
  for (int it = 0; it < kMaxIter; ++it) {
    const int c = it * 128 + lane * 4;
    if (c < dim) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = (g[it][e] - dot * z[it][e]) * inv;
      store4<T>(dx + c, o);
    }
  }
}

int launch_l2norm_bwd(const NormBwdParams& pr, int n_jobs, int x_dtype, int64_t rows, int dim,
                      int64_t x_stride, int n_split, float eps, cudaStream_t st) {
  TCL_REQUIRE(dim <= 512 && dim % 4 == 0, TCL_ERR_BAD_SHAPE, "normalise backward: dim %d > 512", dim);
  dim3 grid(static_cast<unsigned>((rows + 7) / 8), n_jobs);
  ProfScope prof(TCL_K_L2NORM_BWD, st);
  LaunchCfg L(grid, dim3(256), 0, st);
  switch (x_dtype) {
    case TCL_DT_F32: TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, l2norm_bwd_kernel<float>, pr, rows, dim, x_stride, n_split, eps)); break;
    case TCL_DT_F64: TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, l2norm_bwd_kernel<double>, pr, rows, dim, x_stride, n_split, eps)); break;
    case TCL_DT_F16: TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, l2norm_bwd_kernel<__half>, pr, rows, dim, x_stride, n_split, eps)); break;
    case TCL_DT_BF16: TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, l2norm_bwd_kernel<__nv_bfloat16>, pr, rows, dim, x_stride, n_split, eps)); break;
    default: return set_error(TCL_ERR_BAD_ARG, "unknown x_dtype %d", x_dtype);
  }
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

// ---------------------------------------------------------------------------
// Normalise backward of the sharded shared-G form: g = sum over sources of scale[src] * partial (host_common.h:
// NormShParams), then the same projection as above.  Fixed summation order: row-side slots, then source ranks
// 0..W-1 with their slots.
// ---------------------------------------------------------------------------
template <typename T, int WMAX>
__global__ void __launch_bounds__(256) l2norm_bwd_sharded_kernel(const __grid_constant__ NormShParams pr, int64_t rows,
                                                                 int dim, int64_t x_stride, float eps) {
  griddep_launch();
  griddep_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (pr.sync != nullptr) {  // every source rank's gradient GEMM has completed and its partials have landed here
    if (static_cast<int>(threadIdx.x) < pr.world)
      flag_wait_ge(pr.sync + ShardSync::kGrads + threadIdx.x, ld_relaxed_u32(pr.sync + ShardSync::kBwdEpoch));
    __syncthreads();
  }
  if (row >= rows) return;
  const NormShJob& jb = pr.job[blockIdx.y];
  const T* x = static_cast<const T*>(jb.x) + row * x_stride;
  T* dx = static_cast<T*>(jb.dx) + row * dim;
  const float inv = jb.inv_norm[row];
  const bool clamped = inv >= 1.f / eps;
  // pieces of this row's units (per 256-column half when the accumulators were split)
  int np_row[2] = {0, 0}, np_col[2] = {0, 0};
  for (int dh = 0; dh < pr.n_dsplit; ++dh) {
    if (jb.row_job >= 0) {
      const int T_ = pr.unit_tiles[jb.row_job];
      const int64_t u0 = pr.job_tile_base[jb.row_job] + ((row >> pr.unit_shift) * pr.n_dsplit + dh) * T_;
      np_row[dh] = pc_range_of(pr.total_tiles, u0 + T_ - 1, pr.n_ranges) - pc_range_of(pr.total_tiles, u0, pr.n_ranges) + 1;
    }
    if (jb.col_job >= 0) {
      const int T_ = pr.unit_tiles[jb.col_job];
      const int64_t grow = static_cast<int64_t>(pr.rank) * rows + row;
      const int64_t u0 = pr.job_tile_base[jb.col_job] + ((grow >> pr.unit_shift) * pr.n_dsplit + dh) * T_;
      np_col[dh] = pc_range_of(pr.total_tiles, u0 + T_ - 1, pr.n_ranges) - pc_range_of(pr.total_tiles, u0, pr.n_ranges) + 1;
    }
  }
  float sc[TCL_MAX_PEERS];
#pragma unroll
  for (int r = 0; r < TCL_MAX_PEERS; ++r) sc[r] = r < pr.world ? __ldcv(pr.scales + r) : 0.f;
  const float sc_own = __ldcv(pr.scales + pr.rank);
  constexpr int kMaxIter = 4;
  float g[kMaxIter][4], z[kMaxIter][4];
  float dot = 0.f;
  const int64_t col_slot_stride = rows * static_cast<int64_t>(dim);
  // common case (fp16 column partials, at most two pieces per unit): fixed trip counts, so every load of a 128-column
  // step - x, the row-side slots and W x slots column-side partials - is issued before anything is consumed, and the
  // four steps overlap; with the piece counts as loop bounds the row cost ~12 dependent round trips to memory
  // (45 us per launch at 4096 rows per rank).  Same summation order as the general path below (absent entries add 0).
  constexpr int KR = 6;  // row-side pieces held in registers (a 64-tile row unit spans up to 5 tile ranges at 1024 rows per rank)
  const bool fast = pr.col16 && pr.world <= WMAX && np_row[0] <= KR && np_row[1] <= KR && np_col[0] <= 2 && np_col[1] <= 2;
  if (fast) {
#pragma unroll
    for (int it = 0; it < kMaxIter; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < dim) {
        const int dh = (pr.n_dsplit == 2 && c >= 256) ? 1 : 0;
        const int64_t off = row * dim + c;
        float xv[4];
        load4<T>(x + c, xv);
        float4 rp[KR];
        uint2 h[2][WMAX];
#pragma unroll
        for (int k = 0; k < KR; ++k)
          rp[k] = k < np_row[dh] ? __ldcs(reinterpret_cast<const float4*>(jb.row_part + k * pr.row_slot_stride + off))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
          for (int r = 0; r < WMAX; ++r)
            h[k][r] = (k < np_col[dh] && r < pr.world)
                          ? __ldcv(reinterpret_cast<const uint2*>(static_cast<const __half*>(jb.col_part) +
                                                                  (static_cast<int64_t>(r) * pr.n_slots_col + k) * col_slot_stride + off))
                          : make_uint2(0u, 0u);
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < KR; ++k) {
          acc[0] += sc_own * rp[k].x; acc[1] += sc_own * rp[k].y; acc[2] += sc_own * rp[k].z; acc[3] += sc_own * rp[k].w;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
#pragma unroll
          for (int r = 0; r < WMAX; ++r) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h[k][r].x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&h[k][r].y));
            acc[0] += sc[r] * a.x; acc[1] += sc[r] * a.y; acc[2] += sc[r] * b.x; acc[3] += sc[r] * b.y;
          }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          g[it][e] = acc[e];
          z[it][e] = xv[e] * inv;
          dot += g[it][e] * z[it][e];
        }
      }
    }
  } else {
#pragma unroll
  for (int it = 0; it < kMaxIter; ++it) {
    const int c = it * 128 + lane * 4;
    if (c < dim) {
      const int dh = (pr.n_dsplit == 2 && c >= 256) ? 1 : 0;
      float xv[4];
      load4<T>(x + c, xv);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const int64_t off = row * dim + c;
      for (int k = 0; k < np_row[dh]; ++k) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(jb.row_part + k * pr.row_slot_stride + off));
        acc[0] += sc_own * v.x; acc[1] += sc_own * v.y; acc[2] += sc_own * v.z; acc[3] += sc_own * v.w;
      }
      for (int k = 0; k < np_col[dh]; ++k) {
        float4 v[TCL_MAX_PEERS];
        if (pr.col16) {  // fp16 partials (what crosses NVLink by default)
          uint2 h[TCL_MAX_PEERS];
#pragma unroll
          for (int r = 0; r < TCL_MAX_PEERS; ++r)
            h[r] = r < pr.world ? __ldcv(reinterpret_cast<const uint2*>(static_cast<const __half*>(jb.col_part) +
                                                                       (static_cast<int64_t>(r) * pr.n_slots_col + k) * col_slot_stride + off))
                                : make_uint2(0u, 0u);
#pragma unroll
          for (int r = 0; r < TCL_MAX_PEERS; ++r) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h[r].x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&h[r].y));
            v[r] = make_float4(a.x, a.y, b.x, b.y);
          }
        } else {
#pragma unroll
          for (int r = 0; r < TCL_MAX_PEERS; ++r)
            v[r] = r < pr.world
                       ? __ldcv(reinterpret_cast<const float4*>(static_cast<const float*>(jb.col_part) + (static_cast<int64_t>(r) * pr.n_slots_col + k) * col_slot_stride + off))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int r = 0; r < TCL_MAX_PEERS; ++r) {
          if (r < pr.world) { acc[0] += sc[r] * v[r].x; acc[1] += sc[r] * v[r].y; acc[2] += sc[r] * v[r].z; acc[3] += sc[r] * v[r].w; }
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        g[it][e] = acc[e];
        z[it][e] = xv[e] * inv;
        dot += g[it][e] * z[it][e];
      }
    }
  }
  }
  dot = warp_sum(dot);
  if (clamped) dot = 0.f;
#pragma unroll
  for (int it = 0; it < kMaxIter; ++it) {
    const int c = it * 128 + lane * 4;
    if (c < dim) {
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) o[e] = (g[it][e] - dot * z[it][e]) * inv;
      store4<T>(dx + c, o);
    }
  }
}

int launch_l2norm_bwd_sharded(const NormShParams& pr, int n_jobs, int x_dtype, int64_t rows, int dim, int64_t x_stride,
                              float eps, cudaStream_t st) {
  TCL_REQUIRE(dim <= 512 && dim % 4 == 0, TCL_ERR_BAD_SHAPE, "normalise backward: dim %d > 512", dim);
  dim3 grid(static_cast<unsigned>((rows + 7) / 8), n_jobs);
  ProfScope prof(TCL_K_L2NORM_BWD, st);
  LaunchCfg L(grid, dim3(256), 0, st);
  // the fast path holds 2 x WMAX column-side loads per step in registers: instantiate for the world sizes of one node
#define TCL_NBS(T)                                                                                                    \
  do {                                                                                                                \
    if (pr.world <= 2) TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, l2norm_bwd_sharded_kernel<T, 2>, pr, rows, dim, x_stride, eps)); \
    else if (pr.world <= 4) TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, l2norm_bwd_sharded_kernel<T, 4>, pr, rows, dim, x_stride, eps)); \
    else TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, l2norm_bwd_sharded_kernel<T, TCL_MAX_PEERS>, pr, rows, dim, x_stride, eps)); \
  } while (0)
  switch (x_dtype) {
    case TCL_DT_F32: TCL_NBS(float); break;
    case TCL_DT_F64: TCL_NBS(double); break;
    case TCL_DT_F16: TCL_NBS(__half); break;
    case TCL_DT_BF16: TCL_NBS(__nv_bfloat16); break;
    default: return set_error(TCL_ERR_BAD_ARG, "unknown x_dtype %d", x_dtype);
  }
#undef TCL_NBS
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

}  // namespace tcl

using namespace tcl;

static int launch_push(const struct tcl::PushPack& pk, int n_tensors, int x_dtype, int64_t rows, int dim, int64_t stride,
                       int64_t z_stride, int op_format, float eps, cudaStream_t st);

extern "C" int tcl_l2norm_fwd(int n_tensors, const void* const* x, int x_dtype, int64_t rows,
                              int64_t dim, int64_t x_row_stride, void* const* z, int64_t z_row_stride,
                              int op_format, float* const* inv_norm, float eps, void* stream) {
  if (z_row_stride == 0) z_row_stride = dim;
  TCL_REQUIRE(z_row_stride >= dim && z_row_stride % 8 == 0, TCL_ERR_BAD_ALIGN,
              "l2norm: z_row_stride %lld must be >= dim and a multiple of 8", (long long)z_row_stride);
  TCL_REQUIRE(n_tensors >= 1 && n_tensors <= TCL_MAX_TENSORS, TCL_ERR_BAD_ARG, "n_tensors %d", n_tensors);
  TCL_REQUIRE(rows >= 0 && dim >= 8 && dim % 8 == 0, TCL_ERR_BAD_SHAPE,
              "l2norm: dim must be a positive multiple of 8 (got %lld)", (long long)dim);
  TCL_REQUIRE(x_row_stride >= dim && (x_row_stride * dtype_size(x_dtype)) % 16 == 0, TCL_ERR_BAD_ALIGN,
              "l2norm: row stride %lld not 16-byte aligned", (long long)x_row_stride);
  TCL_REQUIRE(op_format == TCL_OP_F16 || op_format == TCL_OP_BF16, TCL_ERR_BAD_ARG, "op_format %d", op_format);
  if (int e = require_sm100()) return e;
  if (rows == 0) return TCL_OK;
  PtrPack3 pk{};
  for (int i = 0; i < n_tensors; ++i) {
    TCL_REQUIRE(x[i] && z[i] && inv_norm[i], TCL_ERR_BAD_ARG, "l2norm: null pointer (tensor %d)", i);
    TCL_REQUIRE(aligned_to(x[i], 16) && aligned_to(z[i], 16), TCL_ERR_BAD_ALIGN, "l2norm: pointers must be 16-byte aligned");
    pk.in[i] = x[i]; pk.out[i] = z[i]; pk.aux[i] = inv_norm[i];
  }
  if (dim <= 512) {
    // the same kernel as the sharded forms (8 elements per lane, 16-byte stores, one destination): identical bits
    // whether a row is normalised here, by tcl_l2norm_fwd_bcast or by tcl_l2norm_fwd_push
    PushPack pp{};
    pp.rank = 0;
    pp.world = 1;
    pp.n_tensors = n_tensors;
    pp.remote = 2;
    for (int i = 0; i < n_tensors; ++i) {
      pp.in[i] = x[i];
      pp.aux[i] = inv_norm[i];
      pp.out[0][i] = z[i];
    }
    return launch_push(pp, n_tensors, x_dtype, rows, (int)dim, x_row_stride, z_row_stride, op_format, eps,
                       static_cast<cudaStream_t>(stream));
  }
  return launch_fwd<true>(pk, n_tensors, x_dtype, rows, (int)dim, x_row_stride, z_row_stride, op_format, eps,
                          static_cast<cudaStream_t>(stream));
}

extern "C" int tcl_l2norm_fwd_bcast(int n_tensors, const void* const* x, int x_dtype, int64_t rows, int64_t dim,
                                    int64_t x_row_stride, int n_dst, void* const* z_dst, int64_t z_row_stride,
                                    int op_format, float* const* inv_norm, float eps, void* stream) {
  if (z_row_stride == 0) z_row_stride = dim;
  TCL_REQUIRE(z_row_stride >= dim && z_row_stride % 8 == 0, TCL_ERR_BAD_ALIGN, "l2norm_bcast: z_row_stride");
  TCL_REQUIRE(n_tensors >= 1 && n_tensors <= TCL_MAX_TENSORS, TCL_ERR_BAD_ARG, "n_tensors %d", n_tensors);
  TCL_REQUIRE(n_dst >= 1 && n_dst <= TCL_MAX_PEERS, TCL_ERR_BAD_ARG, "l2norm_bcast: n_dst %d", n_dst);
  TCL_REQUIRE(rows >= 0 && dim >= 8 && dim % 8 == 0 && dim <= 512, TCL_ERR_BAD_SHAPE,
              "l2norm_bcast: dim must be a multiple of 8 in [8, 512] (got %lld)", (long long)dim);
  TCL_REQUIRE(x_row_stride >= dim && (x_row_stride * dtype_size(x_dtype)) % 16 == 0, TCL_ERR_BAD_ALIGN, "l2norm_bcast: row stride");
  TCL_REQUIRE(op_format == TCL_OP_F16 || op_format == TCL_OP_BF16, TCL_ERR_BAD_ARG, "op_format %d", op_format);
  PushPack pk{};
  pk.rank = 0;
  pk.world = n_dst;
  pk.n_tensors = n_tensors;
  pk.remote = 2;
  for (int i = 0; i < n_tensors; ++i) {
    TCL_REQUIRE(x[i] && inv_norm[i] && aligned_to(x[i], 16), TCL_ERR_BAD_ALIGN, "l2norm_bcast: input %d", i);
    pk.in[i] = x[i];
    pk.aux[i] = inv_norm[i];
    for (int d = 0; d < n_dst; ++d) {
      void* p = z_dst[d * n_tensors + i];
      // destination 0 (the own buffer) is mandatory; a NULL further destination means "this peer never reads this
      // tensor's rows" (e.g. the text modality under the sharded shared-G backward) and is skipped
      TCL_REQUIRE((p || d > 0) && aligned_to(p, 16), TCL_ERR_BAD_ALIGN, "l2norm_bcast: destination %d of tensor %d", d, i);
      pk.out[d][i] = p;
    }
  }
  if (int e = require_sm100()) return e;
  if (rows == 0) return TCL_OK;
  return launch_push(pk, n_tensors, x_dtype, rows, (int)dim, x_row_stride, z_row_stride, op_format, eps,
                     static_cast<cudaStream_t>(stream));
}

template <typename TIn>
static int launch_push_t(const PushPack& pk, int n_tensors, int64_t rows, int dim, int64_t stride, int64_t z_stride,
                         int op_format, float eps, cudaStream_t st) {
  dim3 grid(static_cast<unsigned>((rows + 7) / 8) * n_tensors);
  ProfScope prof(TCL_K_L2NORM_FWD, st);
  LaunchCfg L(grid, dim3(256), 0, st);
  if (op_format == TCL_OP_F16)
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, l2norm_fwd_push_kernel<TIn, __half>, pk, rows, dim, stride, z_stride, eps));
  else
    TCL_CHECK_CUDA(cudaLaunchKernelEx(&L.cfg, l2norm_fwd_push_kernel<TIn, __nv_bfloat16>, pk, rows, dim, stride, z_stride, eps));
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

static int launch_push(const PushPack& pk, int n_tensors, int x_dtype, int64_t rows, int dim, int64_t stride,
                       int64_t z_stride, int op_format, float eps, cudaStream_t st) {
  switch (x_dtype) {
    case TCL_DT_F32: return launch_push_t<float>(pk, n_tensors, rows, dim, stride, z_stride, op_format, eps, st);
    case TCL_DT_F64: return launch_push_t<double>(pk, n_tensors, rows, dim, stride, z_stride, op_format, eps, st);
    case TCL_DT_F16: return launch_push_t<__half>(pk, n_tensors, rows, dim, stride, z_stride, op_format, eps, st);
    case TCL_DT_BF16: return launch_push_t<__nv_bfloat16>(pk, n_tensors, rows, dim, stride, z_stride, op_format, eps, st);
  }
  return set_error(TCL_ERR_BAD_ARG, "unknown x_dtype %d", x_dtype);
}

extern "C" size_t tcl_shard_sync_bytes(void) { return sizeof(uint32_t) * ShardSync::kWords; }

extern "C" int tcl_l2norm_fwd_push(int n_tensors, const void* const* x, int x_dtype, int64_t rows, int64_t dim,
                                   int64_t x_row_stride, int rank, int world, void* const* z_dst, int64_t z_row_stride,
                                   int op_format, float* const* inv_norm, float eps, void* const* sync_ptrs,
                                   int remote, void* stream) {
  if (z_row_stride == 0) z_row_stride = dim;
  TCL_REQUIRE(z_row_stride >= dim && z_row_stride % 8 == 0, TCL_ERR_BAD_ALIGN, "l2norm_push: z_row_stride");
  TCL_REQUIRE(n_tensors >= 1 && n_tensors <= TCL_MAX_TENSORS, TCL_ERR_BAD_ARG, "n_tensors %d", n_tensors);
  TCL_REQUIRE(world >= 1 && world <= TCL_MAX_PEERS && rank >= 0 && rank < world, TCL_ERR_BAD_ARG, "l2norm_push: rank %d of %d", rank, world);
  TCL_REQUIRE(rows >= 1 && rows <= 128 * ShardSync::kMaxChunks && dim >= 8 && dim % 8 == 0 && dim <= 512, TCL_ERR_BAD_SHAPE,
              "l2norm_push: rows in [1, %d], dim a multiple of 8 in [8, 512] (got %lld x %lld)", 128 * ShardSync::kMaxChunks,
              (long long)rows, (long long)dim);
  TCL_REQUIRE(x_row_stride >= dim && (x_row_stride * dtype_size(x_dtype)) % 16 == 0, TCL_ERR_BAD_ALIGN, "l2norm_push: row stride");
  TCL_REQUIRE(op_format == TCL_OP_F16 || op_format == TCL_OP_BF16, TCL_ERR_BAD_ARG, "op_format %d", op_format);
  TCL_REQUIRE(x && z_dst && inv_norm && sync_ptrs, TCL_ERR_BAD_ARG, "l2norm_push: null pointer");
  if (int e = require_sm100()) return e;
  PushPack pk{};
  pk.rank = rank;
  pk.world = world;
  pk.n_tensors = n_tensors;
  pk.remote = remote != 0;
  for (int r = 0; r < world; ++r) {
    TCL_REQUIRE(sync_ptrs[r] && aligned_to(sync_ptrs[r], 16), TCL_ERR_BAD_ALIGN, "l2norm_push: sync pad %d", r);
    pk.sync[r] = static_cast<uint32_t*>(sync_ptrs[r]);
  }
  for (int i = 0; i < n_tensors; ++i) {
    TCL_REQUIRE(x[i] && inv_norm[i] && aligned_to(x[i], 16), TCL_ERR_BAD_ALIGN, "l2norm_push: input %d", i);
    pk.in[i] = x[i];
    pk.aux[i] = inv_norm[i];
    for (int d = 0; d < world; ++d) {
      void* p = z_dst[d * n_tensors + i];
      TCL_REQUIRE(p && aligned_to(p, 16), TCL_ERR_BAD_ALIGN, "l2norm_push: destination %d of tensor %d", d, i);
      pk.out[d][i] = p;
    }
  }
  return launch_push(pk, n_tensors, x_dtype, rows, (int)dim, x_row_stride, z_row_stride, op_format, eps,
                     static_cast<cudaStream_t>(stream));
}

extern "C" int tcl_peer_sum_f32(int n_src, const float* const* src, int64_t n, float* out, void* stream) {
  TCL_REQUIRE(n_src >= 1 && n_src <= TCL_MAX_PEERS && src && out, TCL_ERR_BAD_ARG, "peer_sum: n_src %d", n_src);
  TCL_REQUIRE(n >= 0 && aligned_to(out, 16), TCL_ERR_BAD_ALIGN, "peer_sum: output alignment");
  if (int e = require_sm100()) return e;
  if (n == 0) return TCL_OK;
  const float* s[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  for (int r = 0; r < n_src; ++r) {
    TCL_REQUIRE(src[r] && aligned_to(src[r], 16), TCL_ERR_BAD_ALIGN, "peer_sum: source %d", r);
    s[r] = src[r];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n4 = n / 4;
  const int n_tail = static_cast<int>(n - 4 * n4);
  ProfScope prof(TCL_K_PEER_SUM, st);
  peer_sum_kernel<<<static_cast<unsigned>((n4 + n_tail + 255) / 256), 256, 0, st>>>(nullptr, n_src, s[0], s[1], s[2], s[3], s[4],
                                                                                    s[5], s[6], s[7], n4, n_tail, out);
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

extern "C" int tcl_cast_16bit(const void* x, int x_dtype, int64_t rows, int64_t dim,
                              int64_t x_row_stride, void* y, int op_format, void* stream) {
  TCL_REQUIRE(rows >= 0 && dim >= 8 && dim % 8 == 0, TCL_ERR_BAD_SHAPE,
              "cast: dim must be a positive multiple of 8 (got %lld)", (long long)dim);
  TCL_REQUIRE(x && y && aligned_to(x, 16) && aligned_to(y, 16), TCL_ERR_BAD_ALIGN, "cast: pointers must be non-null, 16-byte aligned");
  TCL_REQUIRE(x_row_stride >= dim && (x_row_stride * dtype_size(x_dtype)) % 16 == 0, TCL_ERR_BAD_ALIGN, "cast: row stride");
  TCL_REQUIRE(op_format == TCL_OP_F16 || op_format == TCL_OP_BF16, TCL_ERR_BAD_ARG, "op_format %d", op_format);
  if (int e = require_sm100()) return e;
  if (rows == 0) return TCL_OK;
  PtrPack3 pk{};
  pk.in[0] = x; pk.out[0] = y;
  return launch_fwd<false>(pk, 1, x_dtype, rows, (int)dim, x_row_stride, dim, op_format, 0.f,
                           static_cast<cudaStream_t>(stream));
}

extern "C" int tcl_gather_sum_cast16(int n_src, const void* const* src, int src_dtype, int64_t n_src_rows, int64_t dim,
                                     int64_t src_row_stride, const int64_t* index, int64_t n_out, void* out16,
                                     int op_format, void* stream) {
  TCL_REQUIRE(n_src == 1 || n_src == 2, TCL_ERR_BAD_ARG, "gather_sum: n_src %d", n_src);
  TCL_REQUIRE(n_src_rows >= 1 && n_out >= 0 && dim >= 8 && dim % 8 == 0, TCL_ERR_BAD_SHAPE, "gather_sum: sizes");
  TCL_REQUIRE(src && index && out16 && aligned_to(out16, 16), TCL_ERR_BAD_ALIGN, "gather_sum: null or misaligned pointer");
  TCL_REQUIRE(src_row_stride >= dim && (src_row_stride * dtype_size(src_dtype)) % 16 == 0, TCL_ERR_BAD_ALIGN, "gather_sum: row stride");
  TCL_REQUIRE(op_format == TCL_OP_F16 || op_format == TCL_OP_BF16, TCL_ERR_BAD_ARG, "op_format %d", op_format);
  if (int e = require_sm100()) return e;
  if (n_out == 0) return TCL_OK;
  PtrPack3 pk{};
  for (int i = 0; i < n_src; ++i) {
    TCL_REQUIRE(src[i] && aligned_to(src[i], 16), TCL_ERR_BAD_ALIGN, "gather_sum: source %d null or misaligned", i);
    pk.in[i] = src[i];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (src_dtype) {
    case TCL_DT_F32: return launch_gather_sum<float>(pk, n_src, n_src_rows, (int)dim, src_row_stride, index, n_out, out16, op_format, st);
    case TCL_DT_F64: return launch_gather_sum<double>(pk, n_src, n_src_rows, (int)dim, src_row_stride, index, n_out, out16, op_format, st);
    case TCL_DT_F16: return launch_gather_sum<__half>(pk, n_src, n_src_rows, (int)dim, src_row_stride, index, n_out, out16, op_format, st);
    case TCL_DT_BF16: return launch_gather_sum<__nv_bfloat16>(pk, n_src, n_src_rows, (int)dim, src_row_stride, index, n_out, out16, op_format, st);
  }
  return set_error(TCL_ERR_BAD_ARG, "unknown src_dtype %d", src_dtype);
}

extern "C" int tcl_transpose_16bit(int n_tensors, const void* const* z, int64_t rows, int64_t dim,
                                   int64_t z_row_stride, void* const* zt, int64_t ld_t, void* stream) {
  if (z_row_stride == 0) z_row_stride = dim;
  TCL_REQUIRE(z_row_stride >= dim && z_row_stride % 2 == 0, TCL_ERR_BAD_ALIGN, "transpose: z_row_stride");
  TCL_REQUIRE(n_tensors >= 1 && n_tensors <= TCL_MAX_TENSORS, TCL_ERR_BAD_ARG, "n_tensors %d", n_tensors);
  TCL_REQUIRE(rows >= 0 && dim >= 2 && dim % 2 == 0, TCL_ERR_BAD_SHAPE, "transpose: dim %lld", (long long)dim);
  TCL_REQUIRE(ld_t >= rows && ld_t % 8 == 0, TCL_ERR_BAD_ALIGN, "transpose: ld_t %lld must be >= rows and a multiple of 8", (long long)ld_t);
  if (int e = require_sm100()) return e;
  if (rows == 0) return TCL_OK;
  PtrPack3 pk{};
  for (int i = 0; i < n_tensors; ++i) {
    TCL_REQUIRE(z[i] && zt[i] && aligned_to(z[i], 4) && aligned_to(zt[i], 16), TCL_ERR_BAD_ALIGN, "transpose: pointer alignment");
    pk.in[i] = z[i]; pk.out[i] = zt[i];
  }
  dim3 grid(static_cast<unsigned>((rows + 63) / 64), static_cast<unsigned>((dim + 63) / 64), n_tensors);
  {
    ProfScope prof(TCL_K_TRANSPOSE16, static_cast<cudaStream_t>(stream));
    transpose16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(pk, rows, (int)dim, z_row_stride, ld_t);
  }
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}
