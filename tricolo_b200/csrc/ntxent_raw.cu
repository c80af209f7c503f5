// NTXentLoss.forward(zis, zjs, norm=False) — tricolo/loss/nt_xent.py:55-74 with the F.normalize of :56-57 skipped —
// and its autograd, for the same pair list as the normalised path (tricolo/model/tricolo_net.py:56-65).
//
// Without the normalisation the logits z_i . z_j / tau are unbounded, so (a) the fixed-shift sum-exp of the tensor-core
// path does not apply: every row and column carries an online (max, sum) pair, merged in a fixed order; (b) 16-bit
// operands are not good enough: a logit of magnitude 1e2..1e3 needs an absolute error below 1e-3 for rtol 1e-3 on the
// softmax, i.e. fp32 products.  TriCoLoNet never calls this mode (tricolo_net.py:63), so it is the one path of the
// library that runs on the fp32 FMA pipe: 64 x 64 logit tiles, k-major operand slices in shared memory, a 4 x 4
// register tile per thread.  The B x B logits still never reach HBM: the forward keeps only the row / column
// statistics, the backward re-forms each tile, turns it into the gradient tile G in shared memory and multiplies it
// with a 128-dim slice of the other tensor.  No atomics on floating-point data: fixed summation orders throughout.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <float.h>
#include <math.h>

#include "host_common.h"
#include "../../include/tricolo_b200.h"

namespace tcl {
namespace raw {

constexpr int RT = 64;      // tile edge (rows of either tensor)
constexpr int RK = 16;      // dims per operand slice
constexpr int RP = RT + 4;  // padded row of a k-major slice (rows stay 16-byte aligned)

template <typename T>
__device__ __forceinline__ float4 ld4(const T* p);
template <>
__device__ __forceinline__ float4 ld4<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 ld4<__half>(const __half* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <>
__device__ __forceinline__ float4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T>
__device__ __forceinline__ void st4(T* p, const float (&v)[4]);
template <>
__device__ __forceinline__ void st4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <>
__device__ __forceinline__ void st4<__half>(__half* p, const float (&v)[4]) {
  uint2 u;
  *reinterpret_cast<__half2*>(&u.x) = __floats2half2_rn(v[0], v[1]);
  *reinterpret_cast<__half2*>(&u.y) = __floats2half2_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}
template <>
__device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  uint2 u;
  *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(v[0], v[1]);
  *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}

// acc[r][c] = <X[i0 + ty*4 + r], Y[j0 + tx*4 + c]> in fp32, k ascending.  256 threads; thread (ty, tx) = (t / 16, t % 16).
// Rows past n read as zero.  The next operand slice is fetched into registers while the current one is consumed.
template <typename T>
__device__ __forceinline__ void s_tile(const T* __restrict__ X, const T* __restrict__ Y, int64_t stride, int n, int dim,
                                       int i0, int j0, float (*Xs)[RP], float (*Ys)[RP], float (&acc)[4][4]) {
  const int t = threadIdx.x, lr = t >> 2, lk = (t & 3) * 4;  // loader: row lr, dims lk..lk+3 of the slice
  const int ty = t >> 4, tx = t & 15;
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
  const bool xin = i0 + lr < n, yin = j0 + lr < n;
  const T* xp = X + static_cast<int64_t>(i0 + lr) * stride + lk;
  const T* yp = Y + static_cast<int64_t>(j0 + lr) * stride + lk;
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 xr = xin ? ld4<T>(xp) : zero, yr = yin ? ld4<T>(yp) : zero;
  for (int k0 = 0; k0 < dim; k0 += RK) {
    Xs[lk + 0][lr] = xr.x; Xs[lk + 1][lr] = xr.y; Xs[lk + 2][lr] = xr.z; Xs[lk + 3][lr] = xr.w;
    Ys[lk + 0][lr] = yr.x; Ys[lk + 1][lr] = yr.y; Ys[lk + 2][lr] = yr.z; Ys[lk + 3][lr] = yr.w;
    __syncthreads();
    if (k0 + RK < dim) {
      xr = xin ? ld4<T>(xp + k0 + RK) : zero;
      yr = yin ? ld4<T>(yp + k0 + RK) : zero;
    }
#pragma unroll
    for (int k = 0; k < RK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&Xs[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ys[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
    }
    __syncthreads();
  }
}

// (m, s) <- merge of (m, s) and (m2, s2): the sum of exp(. - max) of the union
__device__ __forceinline__ void merge_ms(float& m, float& s, float m2, float s2) {
  const float mn = fmaxf(m, m2);
  s = s * expf(m - mn) + s2 * expf(m2 - mn);
  m = mn;
}

struct FwdParams {
  const void* x[TCL_MAX_TENSORS];
  int pair_row[TCL_MAX_PAIRS], pair_col[TCL_MAX_PAIRS];
  int n, dim, n_rt, n_split;
  int64_t stride;
  float inv_tau;
  float2* row_part;  // [pair][split][n]  (max, sum exp(. - max)) of the split's columns
  float2* col_part;  // [pair][row tile][n] the same of the row tile's rows
  float* diag;       // [pair][n] the positive logit
};

template <typename T>
__global__ void __launch_bounds__(256) raw_fwd_kernel(const __grid_constant__ FwdParams P) {
  __shared__ __align__(16) float Xs[RK][RP];
  __shared__ __align__(16) float Ys[RK][RP];
  __shared__ float Ls[RT][RT + 1];
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int pair = blockIdx.z, i0 = blockIdx.x * RT, n = P.n;
  const T* X = static_cast<const T*>(P.x[P.pair_row[pair]]);
  const T* Y = static_cast<const T*>(P.x[P.pair_col[pair]]);
  const int jt_lo = static_cast<int>(static_cast<int64_t>(P.n_rt) * blockIdx.y / P.n_split);
  const int jt_hi = static_cast<int>(static_cast<int64_t>(P.n_rt) * (blockIdx.y + 1) / P.n_split);
  float M = -INFINITY, S = 0.f;  // running statistics of row i0 + t (threads 0..63)
  for (int jt = jt_lo; jt < jt_hi; ++jt) {
    const int j0 = jt * RT;
    float acc[4][4];
    s_tile<T>(X, Y, P.stride, n, P.dim, i0, j0, Xs, Ys, acc);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) Ls[ty * 4 + r][tx * 4 + c] = acc[r][c] * P.inv_tau;
    __syncthreads();
    if (t < RT) {
      const int i = i0 + t;
      if (i < n) {
        const int nc = n - j0 < RT ? n - j0 : RT;
        float m = Ls[t][0];
        for (int c = 1; c < nc; ++c) m = fmaxf(m, Ls[t][c]);
        float s = 0.f;
        for (int c = 0; c < nc; ++c) s += expf(Ls[t][c] - m);
        merge_ms(M, S, m, s);
        if (i >= j0 && i < j0 + RT) P.diag[static_cast<int64_t>(pair) * n + i] = Ls[t][i - j0];
      }
    } else if (t < 2 * RT) {
      const int c = t - RT, j = j0 + c;
      if (j < n) {
        const int nr = n - i0 < RT ? n - i0 : RT;
        float m = Ls[0][c];
        for (int r = 1; r < nr; ++r) m = fmaxf(m, Ls[r][c]);
        float s = 0.f;
        for (int r = 0; r < nr; ++r) s += expf(Ls[r][c] - m);
        P.col_part[(static_cast<int64_t>(pair) * P.n_rt + blockIdx.x) * n + j] = make_float2(m, s);
      }
    }
    __syncthreads();
  }
  if (t < RT && i0 + t < n)
    P.row_part[(static_cast<int64_t>(pair) * P.n_split + blockIdx.y) * n + i0 + t] = make_float2(M, S);
}

struct FinParams {
  const float2* row_part;
  const float2* col_part;
  const float* diag;
  float* lse_row;  // [pair][n]  natural-log LSE of the logit rows
  float* lse_col;  // [pair][n]  ... of the logit columns
  double2* blk;    // [pair][blocks] per-block sums of (lse_row - diag, lse_col - diag)
  unsigned int* counters;  // [TCL_MAX_PAIRS + 1], zero on entry (memset by the host)
  float* loss;     // [n_pairs + 1]: pair losses, then their sum
  int n, n_rt, n_split, n_pairs;
  float alpha;
};

__global__ void __launch_bounds__(256) raw_finalize_kernel(const __grid_constant__ FinParams P) {
  __shared__ double sa[256], sb[256];
  __shared__ bool last;
  const int pair = blockIdx.y, n = P.n;
  const int i = blockIdx.x * 256 + threadIdx.x;
  double a = 0.0, b = 0.0;
  if (i < n) {
    float m = -INFINITY, s = 0.f;
    for (int sp = 0; sp < P.n_split; ++sp) {
      const float2 v = P.row_part[(static_cast<int64_t>(pair) * P.n_split + sp) * n + i];
      merge_ms(m, s, v.x, v.y);
    }
    const float lr = m + logf(s);
    m = -INFINITY; s = 0.f;
    for (int rt = 0; rt < P.n_rt; ++rt) {
      const float2 v = P.col_part[(static_cast<int64_t>(pair) * P.n_rt + rt) * n + i];
      merge_ms(m, s, v.x, v.y);
    }
    const float lc = m + logf(s);
    const float d = P.diag[static_cast<int64_t>(pair) * n + i];
    P.lse_row[static_cast<int64_t>(pair) * n + i] = lr;
    P.lse_col[static_cast<int64_t>(pair) * n + i] = lc;
    a = static_cast<double>(lr) - static_cast<double>(d);
    b = static_cast<double>(lc) - static_cast<double>(d);
  }
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      sa[threadIdx.x] += sa[threadIdx.x + o];
      sb[threadIdx.x] += sb[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    P.blk[static_cast<int64_t>(pair) * gridDim.x + blockIdx.x] = make_double2(sa[0], sb[0]);
    __threadfence();
    last = atomicAdd(P.counters + pair, 1u) + 1u == gridDim.x;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double ta = 0.0, tb = 0.0;
    for (unsigned blk = 0; blk < gridDim.x; ++blk) {  // block order: reproducible
      const double* v = reinterpret_cast<const double*>(P.blk + static_cast<int64_t>(pair) * gridDim.x + blk);
      ta += __ldcg(v);
      tb += __ldcg(v + 1);
    }
    // nt_xent.py:71-74: alpha * loss_a + (1 - alpha) * loss_b, each a mean over the batch
    __stcg(P.loss + pair, static_cast<float>((P.alpha * ta + (1.0 - static_cast<double>(P.alpha)) * tb) / n));
    __threadfence();
    if (atomicAdd(P.counters + TCL_MAX_PAIRS, 1u) + 1u == static_cast<unsigned>(P.n_pairs)) {
      __threadfence();
      float tot = 0.f;
      for (int p = 0; p < P.n_pairs; ++p) tot += __ldcg(P.loss + p);  // sum(loss_dict.values()), tricolo_net.py:64
      P.loss[P.n_pairs] = tot;
    }
  }
}

struct BwdJob {
  int other;    // index of the partner tensor
  int pair;
  float w_own;  // weight of the softmax taken along the partner (the own row's LSE): alpha when own = row side
  int own_is_row;
};
struct BwdParams {
  const void* x[TCL_MAX_TENSORS];
  void* dx[TCL_MAX_TENSORS];
  BwdJob job[TCL_MAX_TENSORS][2];
  int n_job[TCL_MAX_TENSORS];
  const float* lse_row;
  const float* lse_col;
  const float* grad_losses;  // [n_pairs] or NULL
  const float* grad_total;   // one float or NULL
  int n, dim, n_rt;
  int64_t stride;
  float inv_tau;
};

// DC = dims of the other tensor handled by one backward CTA: the logit tile is re-formed once per DC-wide slice of the
// gradient, so 256 halves the recompute of 128 at twice the accumulator registers (chosen by the host: 128 while the
// grid would not fill the GPU otherwise)
template <int DC>
constexpr int bwd_smem() { return (2 * RK * RP + RT * (RT + 1) + RT * DC) * 4; }

template <typename T, int DC>
__global__ void __launch_bounds__(256) raw_bwd_kernel(const __grid_constant__ BwdParams P) {
  constexpr int NH = DC / 64;  // float4 groups per thread and gradient row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float (*Xs)[RP] = reinterpret_cast<float (*)[RP]>(smem_raw);
  float (*Ys)[RP] = reinterpret_cast<float (*)[RP]>(smem_raw + RK * RP * 4);
  float (*Yc)[DC] = reinterpret_cast<float (*)[DC]>(smem_raw + 2 * RK * RP * 4);
  float (*Gs)[RT + 1] = reinterpret_cast<float (*)[RT + 1]>(smem_raw + 2 * RK * RP * 4 + RT * DC * 4);
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int m = blockIdx.z, i0 = blockIdx.x * RT, dc0 = blockIdx.y * DC, n = P.n;
  if (P.dx[m] == nullptr) return;  // no gradient wanted for this tensor
  const T* X = static_cast<const T*>(P.x[m]);
  float out[4][4 * NH];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4 * NH; ++c) out[r][c] = 0.f;
  for (int jb = 0; jb < P.n_job[m]; ++jb) {
    const BwdJob J = P.job[m][jb];
    const T* Y = static_cast<const T*>(P.x[J.other]);
    const float* lse_own = (J.own_is_row ? P.lse_row : P.lse_col) + static_cast<int64_t>(J.pair) * n;
    const float* lse_oth = (J.own_is_row ? P.lse_col : P.lse_row) + static_cast<int64_t>(J.pair) * n;
    float g = 0.f;
    if (P.grad_losses != nullptr) g += P.grad_losses[J.pair];
    if (P.grad_total != nullptr) g += P.grad_total[0];
    const float coef = g * P.inv_tau / static_cast<float>(n);  // d loss / d S of a unit logit gradient
    const float w_own = J.w_own, w_oth = 1.f - J.w_own;
    float lo[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty * 4 + r;
      lo[r] = i < n ? lse_own[i] : 0.f;
    }
    for (int jt = 0; jt < P.n_rt; ++jt) {
      const int j0 = jt * RT;
      float acc[4][4];
      s_tile<T>(X, Y, P.stride, n, P.dim, i0, j0, Xs, Ys, acc);
      // G = coef * [w_own softmax over the partner + w_oth softmax over the own side - I]  (closed form of the
      // autograd of nt_xent.py:68-74)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + tx * 4 + c;
        const float lj = j < n ? lse_oth[j] : 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = i0 + ty * 4 + r;
          const float l = acc[r][c] * P.inv_tau;
          float gv = w_own * expf(l - lo[r]) + w_oth * expf(l - lj) - (i == j ? 1.f : 0.f);
          Gs[ty * 4 + r][tx * 4 + c] = (i < n && j < n) ? coef * gv : 0.f;
        }
      }
      // the partner's rows j0.., dims dc0..dc0+DC-1
#pragma unroll
      for (int q = 0; q < (RT * DC / 4) / 256; ++q) {
        const int idx = q * 256 + t, row = idx / (DC / 4), c4 = (idx % (DC / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j0 + row < n && dc0 + c4 < P.dim) v = ld4<T>(Y + static_cast<int64_t>(j0 + row) * P.stride + dc0 + c4);
        *reinterpret_cast<float4*>(&Yc[row][c4]) = v;
      }
      __syncthreads();
#pragma unroll 4
      for (int k = 0; k < RT; ++k) {
        float gk[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) gk[r] = Gs[ty * 4 + r][k];
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const float4 y = *reinterpret_cast<const float4*>(&Yc[k][h * 64 + tx * 4]);
          const float yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) out[r][h * 4 + c] = fmaf(gk[r], yv[c], out[r][h * 4 + c]);
        }
      }
      __syncthreads();
    }
  }
  T* dX = static_cast<T*>(P.dx[m]);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= n) continue;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      const int c = dc0 + h * 64 + tx * 4;
      if (c < P.dim) {
        const float v[4] = {out[r][h * 4 + 0], out[r][h * 4 + 1], out[r][h * 4 + 2], out[r][h * 4 + 3]};
        st4<T>(dX + static_cast<int64_t>(i) * P.dim + c, v);
      }
    }
  }
}

static int n_split_for(int n_pairs, int n_rt) {
  // at least ~2 CTAs per SM in the forward grid where the column sweep allows it
  int s = (2 * kNumSMsB200 + n_pairs * n_rt - 1) / (n_pairs * n_rt);
  if (s > n_rt) s = n_rt;
  return s < 1 ? 1 : s;
}
struct Layout {
  int n_rt, n_split, n_blk;
  size_t off_row, off_col, off_diag, off_blk, off_cnt, total;
};
static Layout layout(int n_pairs, int64_t n) {
  Layout L;
  L.n_rt = static_cast<int>((n + RT - 1) / RT);
  L.n_split = n_split_for(n_pairs, L.n_rt);
  L.n_blk = static_cast<int>((n + 255) / 256);
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) / 256 * 256; return at; };
  L.off_cnt = take(sizeof(unsigned int) * (TCL_MAX_PAIRS + 1));
  L.off_row = take(sizeof(float2) * n_pairs * L.n_split * n);
  L.off_col = take(sizeof(float2) * static_cast<size_t>(n_pairs) * L.n_rt * n);
  L.off_diag = take(sizeof(float) * n_pairs * n);
  L.off_blk = take(sizeof(double2) * n_pairs * L.n_blk);
  L.total = o;
  return L;
}
static size_t state_bytes(int n_pairs, int64_t n) { return (sizeof(float) * 2 * n_pairs * n + 255) / 256 * 256; }

static int check_common(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim, int64_t stride,
                        int n_pairs, const int32_t* pair_row, const int32_t* pair_col, float inv_tau) {
  TCL_REQUIRE(n_tensors >= 2 && n_tensors <= TCL_MAX_TENSORS && n_pairs >= 1 && n_pairs <= TCL_MAX_PAIRS, TCL_ERR_BAD_ARG,
              "ntxent_raw: n_tensors %d, n_pairs %d", n_tensors, n_pairs);
  TCL_REQUIRE(x && pair_row && pair_col, TCL_ERR_BAD_ARG, "ntxent_raw: null pointer");
  TCL_REQUIRE(x_dtype == TCL_DT_F32 || x_dtype == TCL_DT_F16 || x_dtype == TCL_DT_BF16, TCL_ERR_BAD_ARG,
              "ntxent_raw: x_dtype %d (f32, f16 or bf16)", x_dtype);
  TCL_REQUIRE(batch >= 1 && batch <= (1 << 20) && dim >= RK && dim % RK == 0 && stride >= dim, TCL_ERR_BAD_SHAPE,
              "ntxent_raw: batch %lld, dim %lld (a multiple of %d), row stride %lld", (long long)batch, (long long)dim, RK,
              (long long)stride);
  const size_t es = x_dtype == TCL_DT_F32 ? 4 : 2;
  TCL_REQUIRE((stride * es) % 16 == 0, TCL_ERR_BAD_ALIGN, "ntxent_raw: rows must be 16-byte aligned");
  for (int m = 0; m < n_tensors; ++m)
    TCL_REQUIRE(x[m] && aligned_to(x[m], 16), TCL_ERR_BAD_ALIGN, "ntxent_raw: x[%d] must be a 16-byte aligned pointer", m);
  for (int p = 0; p < n_pairs; ++p)
    TCL_REQUIRE(pair_row[p] >= 0 && pair_row[p] < n_tensors && pair_col[p] >= 0 && pair_col[p] < n_tensors &&
                    pair_row[p] != pair_col[p],
                TCL_ERR_BAD_ARG, "ntxent_raw: pair %d = (%d, %d)", p, pair_row[p], pair_col[p]);
  TCL_REQUIRE(inv_tau > 0.f && isfinite(inv_tau), TCL_ERR_BAD_ARG, "ntxent_raw: 1/temperature %g", inv_tau);
  return TCL_OK;
}

}  // namespace raw
}  // namespace tcl

using namespace tcl;
using namespace tcl::raw;

extern "C" size_t tcl_ntxent_raw_state_bytes(int n_pairs, int64_t batch) {
  if (n_pairs < 1 || n_pairs > TCL_MAX_PAIRS || batch < 1) return 0;
  return state_bytes(n_pairs, batch);
}
extern "C" size_t tcl_ntxent_raw_workspace_bytes(int n_pairs, int64_t batch) {
  if (n_pairs < 1 || n_pairs > TCL_MAX_PAIRS || batch < 1) return 0;
  return layout(n_pairs, batch).total;
}

extern "C" int tcl_ntxent_raw_fwd(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim,
                                  int64_t x_row_stride, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                                  float inv_tau, float alpha, void* state, size_t state_bytes_in, void* workspace,
                                  size_t workspace_bytes, float* loss, void* stream) {
  if (int e = check_common(n_tensors, x, x_dtype, batch, dim, x_row_stride, n_pairs, pair_row, pair_col, inv_tau)) return e;
  TCL_REQUIRE(state && workspace && loss, TCL_ERR_BAD_ARG, "ntxent_raw_fwd: null pointer");
  const Layout L = layout(n_pairs, batch);
  TCL_REQUIRE(state_bytes_in >= state_bytes(n_pairs, batch) && aligned_to(state, 256), TCL_ERR_WORKSPACE,
              "ntxent_raw_fwd: state buffer (tcl_ntxent_raw_state_bytes, 256-byte aligned)");
  TCL_REQUIRE(workspace_bytes >= L.total && aligned_to(workspace, 256), TCL_ERR_WORKSPACE,
              "ntxent_raw_fwd: workspace (tcl_ntxent_raw_workspace_bytes, 256-byte aligned)");
  if (int e = require_sm100()) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  char* ws = static_cast<char*>(workspace);
  FwdParams F;
  memset(&F, 0, sizeof(F));
  for (int m = 0; m < n_tensors; ++m) F.x[m] = x[m];
  for (int p = 0; p < n_pairs; ++p) { F.pair_row[p] = pair_row[p]; F.pair_col[p] = pair_col[p]; }
  F.n = static_cast<int>(batch); F.dim = static_cast<int>(dim); F.n_rt = L.n_rt; F.n_split = L.n_split;
  F.stride = x_row_stride; F.inv_tau = inv_tau;
  F.row_part = reinterpret_cast<float2*>(ws + L.off_row);
  F.col_part = reinterpret_cast<float2*>(ws + L.off_col);
  F.diag = reinterpret_cast<float*>(ws + L.off_diag);
  TCL_CHECK_CUDA(cudaMemsetAsync(ws + L.off_cnt, 0, sizeof(unsigned int) * (TCL_MAX_PAIRS + 1), st));
  const dim3 grid(L.n_rt, L.n_split, n_pairs);
  {
    ProfScope prof(TCL_K_NTXENT_RAW_FWD, st);
    switch (x_dtype) {
      case TCL_DT_F32: raw_fwd_kernel<float><<<grid, 256, 0, st>>>(F); break;
      case TCL_DT_F16: raw_fwd_kernel<__half><<<grid, 256, 0, st>>>(F); break;
      default: raw_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(F); break;
    }
    TCL_CHECK_CUDA(cudaGetLastError());
  }
  FinParams Q;
  Q.row_part = F.row_part; Q.col_part = F.col_part; Q.diag = F.diag;
  Q.lse_row = static_cast<float*>(state);
  Q.lse_col = Q.lse_row + static_cast<int64_t>(n_pairs) * batch;
  Q.blk = reinterpret_cast<double2*>(ws + L.off_blk);
  Q.counters = reinterpret_cast<unsigned int*>(ws + L.off_cnt);
  Q.loss = loss;
  Q.n = F.n; Q.n_rt = L.n_rt; Q.n_split = L.n_split; Q.n_pairs = n_pairs; Q.alpha = alpha;
  {
    ProfScope prof(TCL_K_FWD_FINALIZE, st);
    raw_finalize_kernel<<<dim3(L.n_blk, n_pairs), 256, 0, st>>>(Q);
    TCL_CHECK_CUDA(cudaGetLastError());
  }
  return TCL_OK;
}

extern "C" int tcl_ntxent_raw_bwd(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim,
                                  int64_t x_row_stride, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                                  float inv_tau, float alpha, const void* state, const float* grad_losses,
                                  const float* grad_total, const uint8_t* need_grad_host, void* const* dx,
                                  void* stream) {
  if (int e = check_common(n_tensors, x, x_dtype, batch, dim, x_row_stride, n_pairs, pair_row, pair_col, inv_tau)) return e;
  TCL_REQUIRE(state && need_grad_host && dx, TCL_ERR_BAD_ARG, "ntxent_raw_bwd: null pointer");
  TCL_REQUIRE(grad_losses || grad_total, TCL_ERR_BAD_ARG, "ntxent_raw_bwd: grad_losses and grad_total are both NULL");
  if (int e = require_sm100()) return e;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  BwdParams B;
  memset(&B, 0, sizeof(B));
  bool any = false;
  for (int m = 0; m < n_tensors; ++m) {
    B.x[m] = x[m];
    if (!need_grad_host[m]) continue;
    TCL_REQUIRE(dx[m] && aligned_to(dx[m], 16), TCL_ERR_BAD_ALIGN, "ntxent_raw_bwd: dx[%d] must be a 16-byte aligned pointer", m);
    B.dx[m] = dx[m];
    any = true;
    for (int p = 0; p < n_pairs; ++p) {
      if (pair_row[p] != m && pair_col[p] != m) continue;
      TCL_REQUIRE(B.n_job[m] < 2, TCL_ERR_BAD_ARG, "ntxent_raw_bwd: tensor %d appears in more than two pairs", m);
      BwdJob& J = B.job[m][B.n_job[m]++];
      J.pair = p;
      J.own_is_row = pair_row[p] == m;
      J.other = J.own_is_row ? pair_col[p] : pair_row[p];
      J.w_own = J.own_is_row ? alpha : 1.f - alpha;
    }
  }
  if (!any) return TCL_OK;
  B.lse_row = static_cast<const float*>(state);
  B.lse_col = B.lse_row + static_cast<int64_t>(n_pairs) * batch;
  B.grad_losses = grad_losses;
  B.grad_total = grad_total;
  B.n = static_cast<int>(batch); B.dim = static_cast<int>(dim); B.n_rt = static_cast<int>((batch + RT - 1) / RT);
  B.stride = x_row_stride; B.inv_tau = inv_tau;
  // grid z = tensor; the CTAs of a tensor that needs no gradient leave at once
  const bool wide = dim > 128 && static_cast<int64_t>(B.n_rt) * ((dim + 255) / 256) * n_tensors >= 2 * kNumSMsB200;
  ProfScope prof(TCL_K_NTXENT_RAW_BWD, st);
#define TCL_RAW_BWD(T, DCV)                                                                                   \
  do {                                                                                                        \
    if (int e = ensure_dyn_smem(raw_bwd_kernel<T, DCV>, bwd_smem<DCV>())) return e;                           \
    raw_bwd_kernel<T, DCV><<<dim3(B.n_rt, static_cast<unsigned>((dim + DCV - 1) / DCV), n_tensors), 256,     \
                             bwd_smem<DCV>(), st>>>(B);                                                       \
  } while (0)
  switch (x_dtype) {
    case TCL_DT_F32: if (wide) TCL_RAW_BWD(float, 256); else TCL_RAW_BWD(float, 128); break;
    case TCL_DT_F16: if (wide) TCL_RAW_BWD(__half, 256); else TCL_RAW_BWD(__half, 128); break;
    default: if (wide) TCL_RAW_BWD(__nv_bfloat16, 256); else TCL_RAW_BWD(__nv_bfloat16, 128); break;
  }
#undef TCL_RAW_BWD
  TCL_CHECK_CUDA(cudaGetLastError());
  return TCL_OK;
}
