// Triplet loss of tricolo/loss/triplet.py (TripletLoss.forward, :202-224, and _pairwise_distances, :11-45) and its
// autograd - SURVEY.md 8f row 4.  The reference builds the B x B distance matrix with one matmul and then walks it
// in two Python loops; it is only ever used at training batch sizes (128-256), so everything here is plain fp32
// CUDA-core work on a materialised B x B matrix: the margin (0.025) is far below what 16-bit tensor-core operands
// resolve in ||a||^2 - 2<a,b> + ||b||^2.
//
// Faithful to the reference, including its index convention (the TODO at :30):
//     d2[a][b] = ||zls_b||^2 - 2 <zls_a, zis_b> + ||zis_a||^2        D = sqrt(max(d2, 0)), D = 0 where d2 <= 0
//     semi-hard terms: j != i with D[i][i] < D[i][j] < D[i][i] + margin  ->  D[i][i] - D[i][j] + margin
//     if there is none ("loss_list is 0"): hard terms: j != i with D[i][j] < D[i][i]
//     loss = mean of the terms                      (no term at all: the reference divides by zero -> info[2] = 2)
#include "host_common.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

template <typename T>
__device__ __forceinline__ float ld1(const T* p);
template <> __device__ __forceinline__ float ld1<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld1<double>(const double* p) { return static_cast<float>(*p); }
template <> __device__ __forceinline__ float ld1<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float ld1<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void st1(T* p, float v);
template <> __device__ __forceinline__ void st1<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<double>(double* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st1<__half>(__half* p, float v) { *p = __float2half_rn(v); }
template <> __device__ __forceinline__ void st1<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// squared norms of the rows of both inputs: one warp per row
template <typename T>
__global__ void __launch_bounds__(256) trip_norms_kernel(const T* __restrict__ zis, const T* __restrict__ zls, int64_t stride,
                                                         int batch, int dim, float* __restrict__ n_is,
                                                         float* __restrict__ n_ls) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= batch) return;
  const T* x = (blockIdx.y == 0 ? zis : zls) + static_cast<int64_t>(row) * stride;
  float s = 0.f;
  for (int c = lane; c < dim; c += 32) {
    const float v = ld1<T>(x + c);
    s = fmaf(v, v, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) (blockIdx.y == 0 ? n_is : n_ls)[row] = s;
}

// D[a][b] from a 32 x 32 tile of <zls_a, zis_b>, K in chunks of 32 through shared memory
template <typename T>
__global__ void __launch_bounds__(256) trip_dist_kernel(const T* __restrict__ zis, const T* __restrict__ zls, int64_t stride,
                                                        int batch, int dim, const float* __restrict__ n_is,
                                                        const float* __restrict__ n_ls, float* __restrict__ D) {
  __shared__ float sa[32][33], sb[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < dim; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = ty + 8 * i, k = k0 + tx;
      sa[r][tx] = (a0 + r < batch && k < dim) ? ld1<T>(zls + static_cast<int64_t>(a0 + r) * stride + k) : 0.f;
      sb[r][tx] = (b0 + r < batch && k < dim) ? ld1<T>(zis + static_cast<int64_t>(b0 + r) * stride + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float bv = sb[tx][k];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(sa[ty + 8 * i][k], bv, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int a = a0 + ty + 8 * i, b = b0 + tx;
    if (a < batch && b < batch) {
      const float d2 = n_ls[b] - 2.f * acc[i] + n_is[a];  // :32, the reference's own pairing of the norms
      D[static_cast<int64_t>(a) * batch + b] = d2 > 0.f ? sqrtf(d2) : 0.f;  // :35-43
    }
  }
}

// per row i: sums and counts of the semi-hard and of the hard terms (fixed-order block reduction, fp64 sums)
__global__ void __launch_bounds__(256) trip_select_kernel(const float* __restrict__ D, int batch, float margin,
                                                          double* __restrict__ row_sum, int* __restrict__ row_cnt) {
  const int i = blockIdx.x;
  const float* d = D + static_cast<int64_t>(i) * batch;
  const float dii = d[i];
  const float hi = dii + margin;
  double s_semi = 0.0, s_hard = 0.0;
  int c_semi = 0, c_hard = 0;
  for (int j = threadIdx.x; j < batch; j += blockDim.x) {
    if (j == i) continue;
    const float dij = d[j];
    const float term = dii - dij + margin;
    if (dii < dij && dij < hi) { s_semi += term; ++c_semi; }
    if (dij < dii) { s_hard += term; ++c_hard; }
  }
  __shared__ double ss[2][256];
  __shared__ int sc[2][256];
  ss[0][threadIdx.x] = s_semi; ss[1][threadIdx.x] = s_hard;
  sc[0][threadIdx.x] = c_semi; sc[1][threadIdx.x] = c_hard;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      ss[0][threadIdx.x] += ss[0][threadIdx.x + o]; ss[1][threadIdx.x] += ss[1][threadIdx.x + o];
      sc[0][threadIdx.x] += sc[0][threadIdx.x + o]; sc[1][threadIdx.x] += sc[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    row_sum[2 * i] = ss[0][0]; row_sum[2 * i + 1] = ss[1][0];
    row_cnt[2 * i] = sc[0][0]; row_cnt[2 * i + 1] = sc[1][0];
  }
}

// totals in row order; info = {n_semi, n_hard, mode (0 semi-hard, 1 hard fallback, 2 no term), n_used}
__global__ void __launch_bounds__(256) trip_finalize_kernel(const double* __restrict__ row_sum, const int* __restrict__ row_cnt,
                                                            int batch, float* __restrict__ loss, int* __restrict__ info) {
  __shared__ double ss[2][256];
  __shared__ long long sc[2][256];
  double a = 0.0, b = 0.0;
  long long ca = 0, cb = 0;
  for (int i = threadIdx.x; i < batch; i += blockDim.x) {
    a += row_sum[2 * i]; b += row_sum[2 * i + 1];
    ca += row_cnt[2 * i]; cb += row_cnt[2 * i + 1];
  }
  ss[0][threadIdx.x] = a; ss[1][threadIdx.x] = b; sc[0][threadIdx.x] = ca; sc[1][threadIdx.x] = cb;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      ss[0][threadIdx.x] += ss[0][threadIdx.x + o]; ss[1][threadIdx.x] += ss[1][threadIdx.x + o];
      sc[0][threadIdx.x] += sc[0][threadIdx.x + o]; sc[1][threadIdx.x] += sc[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const long long n_semi = sc[0][0], n_hard = sc[1][0];
    const int mode = n_semi > 0 ? 0 : (n_hard > 0 ? 1 : 2);
    const long long n = mode == 0 ? n_semi : n_hard;
    info[0] = static_cast<int>(n_semi); info[1] = static_cast<int>(n_hard); info[2] = mode; info[3] = static_cast<int>(n);
    *loss = mode == 2 ? __int_as_float(0x7fc00000) : static_cast<float>((mode == 0 ? ss[0][0] : ss[1][0]) / static_cast<double>(n));
  }
}

// V[a][b] = dLoss/d(d2[a][b]) (the selection is re-derived from D), its row sums; one block per row
__global__ void __launch_bounds__(256) trip_v_kernel(const float* __restrict__ D, int batch, float margin,
                                                     const int* __restrict__ info, const float* __restrict__ grad_loss,
                                                     float* __restrict__ V, float* __restrict__ v_rowsum) {
  const int i = blockIdx.x;
  const float* d = D + static_cast<int64_t>(i) * batch;
  float* v = V + static_cast<int64_t>(i) * batch;
  const int mode = info[2];
  const float g = mode == 2 ? 0.f : *grad_loss / static_cast<float>(info[3]);
  const float dii = d[i];
  const float hi = dii + margin;
  float rs = 0.f;
  int cnt = 0;
  for (int j = threadIdx.x; j < batch; j += blockDim.x) {
    if (j == i) continue;
    const float dij = d[j];
    const bool sel = mode == 0 ? (dii < dij && dij < hi) : (mode == 1 && dij < dii);
    const float w = (sel && dij > 0.f) ? -g / (2.f * dij) : 0.f;  // d(term)/dD[i][j] = -1, dD/dd2 = 1/(2D)
    v[j] = w;
    rs += w;
    cnt += sel ? 1 : 0;
  }
  __shared__ float sr[256];
  __shared__ int sc[256];
  sr[threadIdx.x] = rs; sc[threadIdx.x] = cnt;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) { sr[threadIdx.x] += sr[threadIdx.x + o]; sc[threadIdx.x] += sc[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float wd = dii > 0.f ? g * static_cast<float>(sc[0]) / (2.f * dii) : 0.f;  // every selected term has +D[i][i]
    v[i] = wd;
    v_rowsum[i] = sr[0] + wd;
  }
}

__global__ void __launch_bounds__(256) trip_colsum_kernel(const float* __restrict__ V, int batch, float* __restrict__ v_colsum) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  float s = 0.f;
  for (int a = 0; a < batch; ++a) s += V[static_cast<int64_t>(a) * batch + b];
  v_colsum[b] = s;
}

// out[r][:] = 2 diag[r] self[r][:] - 2 sum_k M[r][k] other[k][:]   (M = V, or V^T when kTrans)
template <typename T, bool kTrans>
__global__ void __launch_bounds__(256) trip_grad_kernel(const float* __restrict__ V, const float* __restrict__ diag,
                                                        const T* __restrict__ self, const T* __restrict__ other,
                                                        int64_t stride, int batch, int dim, T* __restrict__ out) {
  __shared__ float sv[32][33], so[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k0 = 0; k0 < batch; k0 += 32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int rr = ty + 8 * i;
      const int r = r0 + rr, k = k0 + tx;
      sv[rr][tx] = (r < batch && k < batch) ? (kTrans ? V[static_cast<int64_t>(k) * batch + r] : V[static_cast<int64_t>(r) * batch + k]) : 0.f;
      const int kk = k0 + rr, c = c0 + tx;
      so[rr][tx] = (kk < batch && c < dim) ? ld1<T>(other + static_cast<int64_t>(kk) * stride + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float ov = so[k][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(sv[ty + 8 * i][k], ov, acc[i]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    if (r < batch && c < dim)
      st1<T>(out + static_cast<int64_t>(r) * dim + c,
             2.f * diag[r] * ld1<T>(self + static_cast<int64_t>(r) * stride + c) - 2.f * acc[i]);
  }
}

struct TripWs {
  float *D, *V, *n_is, *n_ls, *v_rs, *v_cs;
  double* row_sum;
  int *row_cnt, *info;
};
static size_t trip_ws_bytes(int64_t b) {
  return static_cast<size_t>(b) * b * 8 + static_cast<size_t>(b) * (4 * 4 + 16 + 8) + 64 + 512;
}
static TripWs trip_ws(void* ws, int64_t b) {
  TripWs w;
  char* p = static_cast<char*>(ws);
  w.row_sum = reinterpret_cast<double*>(p); p += static_cast<size_t>(b) * 16;
  w.D = reinterpret_cast<float*>(p); p += static_cast<size_t>(b) * b * 4;
  w.V = reinterpret_cast<float*>(p); p += static_cast<size_t>(b) * b * 4;
  w.n_is = reinterpret_cast<float*>(p); p += static_cast<size_t>(b) * 4;
  w.n_ls = reinterpret_cast<float*>(p); p += static_cast<size_t>(b) * 4;
  w.v_rs = reinterpret_cast<float*>(p); p += static_cast<size_t>(b) * 4;
  w.v_cs = reinterpret_cast<float*>(p); p += static_cast<size_t>(b) * 4;
  w.row_cnt = reinterpret_cast<int*>(p); p += static_cast<size_t>(b) * 8;
  w.info = reinterpret_cast<int*>(p);
  return w;
}

template <typename T>
static int trip_fwd_t(const void* zis, const void* zls, int64_t batch, int64_t dim, int64_t stride, float margin, float* loss,
                      const TripWs& w, cudaStream_t st) {
  const int b = static_cast<int>(batch), d = static_cast<int>(dim);
  const T* pi = static_cast<const T*>(zis);
  const T* pl = static_cast<const T*>(zls);
  trip_norms_kernel<T><<<dim3((b + 7) / 8, 2), 256, 0, st>>>(pi, pl, stride, b, d, w.n_is, w.n_ls);
  trip_dist_kernel<T><<<dim3((b + 31) / 32, (b + 31) / 32), 256, 0, st>>>(pi, pl, stride, b, d, w.n_is, w.n_ls, w.D);
  trip_select_kernel<<<b, 256, 0, st>>>(w.D, b, margin, w.row_sum, w.row_cnt);
  trip_finalize_kernel<<<1, 256, 0, st>>>(w.row_sum, w.row_cnt, b, loss, w.info);
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}
template <typename T>
static int trip_bwd_t(const void* zis, const void* zls, int64_t batch, int64_t dim, int64_t stride, float margin,
                      const float* grad_loss, const TripWs& w, void* d_zis, void* d_zls, cudaStream_t st) {
  const int b = static_cast<int>(batch), d = static_cast<int>(dim);
  const T* pi = static_cast<const T*>(zis);
  const T* pl = static_cast<const T*>(zls);
  trip_v_kernel<<<b, 256, 0, st>>>(w.D, b, margin, w.info, grad_loss, w.V, w.v_rs);
  trip_colsum_kernel<<<(b + 255) / 256, 256, 0, st>>>(w.V, b, w.v_cs);
  const dim3 grid((d + 31) / 32, (b + 31) / 32);
  // d zls_r = 2 colsum[r] zls_r - 2 sum_b V[r][b] zis_b ;  d zis_r = 2 rowsum[r] zis_r - 2 sum_a V[a][r] zls_a
  if (d_zls) trip_grad_kernel<T, false><<<grid, 256, 0, st>>>(w.V, w.v_cs, pl, pi, stride, b, d, static_cast<T*>(d_zls));
  if (d_zis) trip_grad_kernel<T, true><<<grid, 256, 0, st>>>(w.V, w.v_rs, pi, pl, stride, b, d, static_cast<T*>(d_zis));
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}

}  // namespace tcl

using namespace tcl;

extern "C" size_t tcl_triplet_workspace_bytes(int64_t batch) { return batch < 1 ? 0 : trip_ws_bytes(batch); }

static int trip_check(const void* zis, const void* zls, int64_t batch, int64_t dim, int64_t stride, void* ws, size_t ws_bytes) {
  TCL_REQUIRE(zis && zls && ws, TCL_ERR_BAD_ARG, "triplet: null pointer");
  TCL_REQUIRE(batch >= 1 && batch <= 32768 && dim >= 1 && stride >= dim, TCL_ERR_BAD_SHAPE, "triplet: sizes");
  TCL_REQUIRE(ws_bytes >= trip_ws_bytes(batch) && aligned_to(ws, 16), TCL_ERR_WORKSPACE, "triplet: workspace");
  return require_sm100();
}

extern "C" int tcl_triplet_fwd(const void* zis, const void* zls, int x_dtype, int64_t batch, int64_t dim,
                               int64_t row_stride, float margin, float* loss, int32_t* info_out, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (int e = trip_check(zis, zls, batch, dim, row_stride, workspace, workspace_bytes)) return e;
  TCL_REQUIRE(loss && info_out, TCL_ERR_BAD_ARG, "triplet_fwd: null output");
  const TripWs w = trip_ws(workspace, batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  switch (x_dtype) {
    case TCL_DT_F32: rc = trip_fwd_t<float>(zis, zls, batch, dim, row_stride, margin, loss, w, st); break;
    case TCL_DT_F64: rc = trip_fwd_t<double>(zis, zls, batch, dim, row_stride, margin, loss, w, st); break;
    case TCL_DT_F16: rc = trip_fwd_t<__half>(zis, zls, batch, dim, row_stride, margin, loss, w, st); break;
    case TCL_DT_BF16: rc = trip_fwd_t<__nv_bfloat16>(zis, zls, batch, dim, row_stride, margin, loss, w, st); break;
    default: return set_error(TCL_ERR_BAD_ARG, "unknown x_dtype %d", x_dtype);
  }
  if (rc) return rc;
  TCL_CHECK_CUDA(cudaMemcpyAsync(info_out, w.info, 16, cudaMemcpyDeviceToDevice, st));
  return TCL_OK;
}

extern "C" int tcl_triplet_bwd(const void* zis, const void* zls, int x_dtype, int64_t batch, int64_t dim,
                               int64_t row_stride, float margin, const float* grad_loss, void* workspace,
                               size_t workspace_bytes, void* d_zis, void* d_zls, void* stream) {
  if (int e = trip_check(zis, zls, batch, dim, row_stride, workspace, workspace_bytes)) return e;
  TCL_REQUIRE(grad_loss, TCL_ERR_BAD_ARG, "triplet_bwd: null grad_loss");
  const TripWs w = trip_ws(workspace, batch);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (x_dtype) {
    case TCL_DT_F32: return trip_bwd_t<float>(zis, zls, batch, dim, row_stride, margin, grad_loss, w, d_zis, d_zls, st);
    case TCL_DT_F64: return trip_bwd_t<double>(zis, zls, batch, dim, row_stride, margin, grad_loss, w, d_zis, d_zls, st);
    case TCL_DT_F16: return trip_bwd_t<__half>(zis, zls, batch, dim, row_stride, margin, grad_loss, w, d_zis, d_zls, st);
    case TCL_DT_BF16: return trip_bwd_t<__nv_bfloat16>(zis, zls, batch, dim, row_stride, margin, grad_loss, w, d_zis, d_zls, st);
  }
  return set_error(TCL_ERR_BAD_ARG, "unknown x_dtype %d", x_dtype);
}
