// Small-batch form of the whole loss (the reference's own operating points: batch 128 bimodal / 256 trimodal,
// config/config.yaml:62-66): ONE cooperative launch for the forward and ONE for the backward.
//
// At B <= a few hundred the multi-kernel pipeline of ntxent_fused.cu is pure launch and pipeline-fill latency (37-43 us
// in a CUDA graph at B = 256 for ~0.6 GFLOP of work).  Here a CTA owns one 32 x 64 tile of one pair's logit matrix and
// keeps both operand blocks (32 rows of the row tensor, 64 of the column tensor, 16-bit, <= 512 dims) in shared memory
// for the whole kernel:
//
//   forward   normalise the 96 rows straight into shared memory (F.normalize, tricolo/loss/nt_xent.py:56-57; the
//             designated tile of each row block also writes z and 1/||x|| to the state buffer) -> S tile with
//             mma.sync.m16n8k16 (fp32 accumulate) -> row / column sum-exp partials and positives -> grid barrier ->
//             one CTA per pair adds the partials in block order, writes lse2 and the pair's loss
//             (nt_xent.py:59-74; TriCoLoNet._calculate_losses, tricolo_net.py:56-65).
//   backward  bulk-copy the same two z blocks -> recompute the S tile -> G tile = w (alpha p_row + (1 - alpha) p_col - I),
//             16-bit, in shared memory -> dZrow partial = G Zcol and dZcol partial = G^T Zrow from the SAME resident
//             operands (ldmatrix.trans supplies the transposed views) -> grid barrier -> every warp pair of the grid
//             (helper CTAs included) finishes a row: partials added in slot order, F.normalize backward, dx in the
//             input dtype.
//
// Why this shape (profiles/small_trace.py, phase stamps of CTA 0): one SM ingests only 60-80 GB/s from L2, so the
// operand blocks are kept small and the tiles spread over as many SMs as the batch allows; legacy mma.sync (HMMA) is
// slow on sm_100 (a 64 x 64 x 512 tile took 2.4 us with sixteen warps), which is the second reason for the half-height
// tile.  tcgen05 would remove the HMMA time but needs TMEM allocation, tensor maps and swizzled operand layouts whose
// set-up costs about what it saves at this size.  The large-batch kernels (tcgen05) are untouched; the two forms share
// the state layout of ntxent_fused.cu, so a forward of one form can be followed by a backward of the other.
// Results are deterministic: fixed summation orders, no floating-point atomics.
#include <atomic>

#include "ntxent_bwd.h"
#include "../../include/tricolo_b200.h"

namespace tcl {

static constexpr int SM_TM = 32;    // tile rows (row tensor of the pair)
static constexpr int SM_TN = 64;    // tile columns (column tensor)
static constexpr int SM_THREADS = 512;
static constexpr int SM_WARPS = SM_THREADS / 32;
static constexpr int SM_PAD = 8;    // 16-bit elements: row stride = 16 bytes mod 128 -> ldmatrix is conflict-free
static constexpr int SM_GLD = 72;   // row stride of the G tile, same property
static constexpr int SM_MAX_SLOTS = 16;
static constexpr int SM_BAR_LANES = 8;
static constexpr int SM_BAR_MAXN = 192;

// Grid-barrier counters, monotonically increasing, one per (lane, grid size): a counter is only ever used by grids of
// ONE size n, so it is a multiple of n between kernels and needs neither a reset nor a generation word.  A launch takes
// the next lane (host counter), so kernels in flight at the same time on one device do not share a counter.
__device__ unsigned long long g_small_bar[SM_BAR_LANES * SM_BAR_MAXN];
__device__ unsigned int g_small_fin[SM_BAR_LANES];  // pair finalisers that are done; reset by the last one

struct SmallParams {
  const void* x[TCL_MAX_TENSORS];
  void* z[TCL_MAX_TENSORS];      // state: [batch][dim] 16-bit
  float* inv[TCL_MAX_TENSORS];   // state: [batch]
  void* dx[TCL_MAX_TENSORS];
  int pair_row[TCL_MAX_PAIRS], pair_col[TCL_MAX_PAIRS];
  int writer_pair[TCL_MAX_TENSORS];  // the pair whose tiles write z / inv of the tensor (its first pair)
  int slot_base[TCL_MAX_PAIRS][2];   // first gradient-partial slot of the pair in its row tensor / column tensor
  int n_slots[TCL_MAX_TENSORS];
  uint8_t need_grad[TCL_MAX_TENSORS];
  float* row_part;   // [pairs][nbj][batch] sum-exp of a row over one 64-column block
  float* col_part;   // [pairs][nbi][batch] sum-exp of a column over one 32-row block
  float* diag2;      // [pairs][batch] positives, log2 domain
  float* lse_row;    // [pairs][batch] log2 domain
  float* lse_col;
  float* loss_parts;
  float* loss;        // [pairs], then their sum if want_total
  const float* grad_losses;  // [pairs] or null
  const float* grad_total;   // [1] or null: gradient of the sum of the pair losses
  float* dz_part;    // [tensors][max_slots][nbj * 64][dim] fp32
  int64_t x_stride;
  int n_tensors, n_pairs, n_tiles, batch, dim, nbi, nbj, bar_lane, max_slots, want_total;
  float c1, alpha, eps, out_scale;
};

// phase stamps of CTA 0 (globaltimer, ns): [0..15] forward, [16..31] backward; trace build only (make trace)
__device__ unsigned long long g_small_trace[32];
#ifdef TCL_PAIR_TRACE
__device__ __forceinline__ void small_stamp(int i) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_small_trace[i] = t;
  }
}
#define SM_STAMP(i) small_stamp(i)
#else
#define SM_STAMP(i)
#endif

// All CTAs of a cooperative launch (co-resident by construction).
__device__ __forceinline__ void small_grid_barrier(int lane_idx) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long* bar = g_small_bar + lane_idx * SM_BAR_MAXN + gridDim.x;
    const unsigned long long n = gridDim.x;
    __threadfence();
    const unsigned long long old = atomicAdd(bar, 1ull);
    const unsigned long long target = (old / n + 1ull) * n;
    unsigned long long now;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(now) : "l"(bar) : "memory");
    } while (now < target);
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)) : "memory");
}
template <int kOp>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if (kOp == TCL_OP_F16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(smem_u32(p)) : "memory");
}

// four consecutive elements of the input dtype <-> fp32 (the dtype is a template parameter of the kernels: the
// run-time switch of norm_fold.cuh, unrolled 64 times, was most of a 64 KB kernel that runs once, cold)
template <typename T> struct Io4;
template <> struct Io4<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct Io4<double> {
  static __device__ __forceinline__ void ld(const double* p, float (&v)[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = static_cast<float>(a.x); v[1] = static_cast<float>(a.y); v[2] = static_cast<float>(b.x); v[3] = static_cast<float>(b.y);
  }
  static __device__ __forceinline__ void st(double* p, const float (&v)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
  }
};
template <> struct Io4<__half> {
  static __device__ __forceinline__ void ld(const __half* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void st(__half* p, const float (&v)[4]) {
    uint2 t;
    *reinterpret_cast<__half2*>(&t.x) = __floats2half2_rn(v[0], v[1]);
    *reinterpret_cast<__half2*>(&t.y) = __floats2half2_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = t;
  }
};
template <> struct Io4<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x)), b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 t;
    *reinterpret_cast<__nv_bfloat162*>(&t.x) = __floats2bfloat162_rn(v[0], v[1]);
    *reinterpret_cast<__nv_bfloat162*>(&t.y) = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = t;
  }
};

struct SmallTile {
  int pair, bi, bj, i0, j0;
  bool is_tile;  // CTAs past the last tile only help with the row phase of the backward
  __device__ explicit SmallTile(const SmallParams& P) {
    const int nn = P.nbi * P.nbj;
    is_tile = static_cast<int>(blockIdx.x) < P.n_tiles;
    const int t = is_tile ? blockIdx.x : 0;
    pair = t / nn;
    bi = (t % nn) / P.nbj;
    bj = t % P.nbj;
    i0 = bi * SM_TM;
    j0 = bj * SM_TN;
  }
};

// S tile (32 x 64) = Xs Ys^T over `dim`, sixteen warps: warp (wm, wn) owns rows wm*16.. and columns wn*8..; acc[e] is
// element (wm*16 + g + (e >> 1) * 8, wn*8 + 2 tig + (e & 1)), g = lane / 4, tig = lane % 4.
template <int kOp>
__device__ __forceinline__ void small_s_tile(const uint16_t* Xs, const uint16_t* Ys, int ld, int dim, int wm, int wn,
                                             int lane, float (&acc)[4]) {
  // four independent accumulation chains (a dependent HMMA issues only every ~100 cycles on sm_100), added at the end
  float c[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int e = 0; e < 4; ++e) c[u][e] = 0.f;
  const uint16_t* ap = Xs + (wm * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ld + (lane >> 4) * 8;
  const uint16_t* bp = Ys + (wn * 8 + (lane & 7)) * ld + (lane >> 3) * 8;  // four k-octets: two k-steps per load
#pragma unroll 2
  for (int k0 = 0; k0 < dim; k0 += 64) {  // dim % 64 == 0
    uint32_t a[4][4], b[2][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) ldsm_x4(a[u], ap + k0 + u * 16);
    ldsm_x4(b[0], bp + k0);
    ldsm_x4(b[1], bp + k0 + 32);
#pragma unroll
    for (int u = 0; u < 4; ++u) mma16816<kOp>(c[u], a[u], b[u >> 1][(u & 1) * 2], b[u >> 1][(u & 1) * 2 + 1]);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) acc[e] = (c[0][e] + c[1][e]) + (c[2][e] + c[3][e]);
}

// out[16 kM][dim] (fp32, row-major, global) = A (16 kM x 16 kK) * Os (16 kK k-rows x dim), A = Gs or Gs^T; the warp owns
// rows wm*16.. and the 64-column chunks n0, n0 + n_step, ...  A lane pair swaps halves so that every store is 16 bytes.
template <int kOp, bool kTransA, int kK>
__device__ __forceinline__ void small_grad_gemm(const uint16_t* Gs, const uint16_t* Os, int ld, int dim, float* out,
                                                int wm, int n0, int n_step, int lane) {
  uint32_t a[kK][4];
#pragma unroll
  for (int ks = 0; ks < kK; ++ks) {
    const int k0 = ks * 16;
    if (!kTransA) ldsm_x4(a[ks], Gs + (wm * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * SM_GLD + k0 + (lane >> 4) * 8);
    else ldsm_x4_t(a[ks], Gs + (k0 + (lane & 7) + (lane >> 4) * 8) * SM_GLD + wm * 16 + ((lane >> 3) & 1) * 8);
  }
  const int g = lane >> 2, tig = lane & 3;
  const bool odd = tig & 1;
  const uint16_t* bp = Os + ((lane & 7) + ((lane >> 3) & 1) * 8) * ld + (lane >> 4) * 8;
#pragma unroll 1
  for (; n0 < dim; n0 += n_step) {
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < kK; ++ks) {
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        uint32_t b[4];
        ldsm_x4_t(b, bp + ks * 16 * ld + n0 + h * 16);
        mma16816<kOp>(acc[2 * h], a[ks], b[0], b[1]);
        mma16816<kOp>(acc[2 * h + 1], a[ks], b[2], b[3]);
      }
    }
    // even tig keeps row g and takes the neighbour's two columns of it; odd tig keeps row g + 8
    float* o = out + static_cast<int64_t>(wm * 16 + g + (odd ? 8 : 0)) * dim + n0 + 2 * (tig & 2);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float s0 = odd ? acc[nt][0] : acc[nt][2], s1 = odd ? acc[nt][1] : acc[nt][3];
      const float r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
      const float4 v = odd ? make_float4(r0, r1, acc[nt][2], acc[nt][3]) : make_float4(acc[nt][0], acc[nt][1], r0, r1);
      __stcg(reinterpret_cast<float4*>(o + nt * 8), v);
    }
  }
}

// Three of the tile's 96 operand rows (virtual rows v0 .. v0 + 2: the 32 rows of the row tensor, then the 64 of the
// column tensor), normalised, into the padded 16-bit blocks of shared memory; rows past the batch are zero.  All of the
// warp's loads are in flight before any is consumed.
template <int kOp, typename TIn>
__device__ __forceinline__ void small_normalise_rows(const SmallParams& P, const SmallTile& T, int v0, uint16_t* Xs,
                                                     uint16_t* Ys, int ld, int lane) {
  constexpr int kRows = 3;
  const int ma = P.pair_row[T.pair], mb = P.pair_col[T.pair];
  const bool wx = P.writer_pair[ma] == T.pair && T.bj == 0, wy = P.writer_pair[mb] == T.pair && T.bi == 0;
  float v[kRows][4][4];
#pragma unroll
  for (int rr = 0; rr < kRows; ++rr) {
    const bool is_y = v0 + rr >= SM_TM;
    const int row = is_y ? T.j0 + v0 + rr - SM_TM : T.i0 + v0 + rr;
    const TIn* xb = static_cast<const TIn*>(P.x[is_y ? mb : ma]);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int c = it * 128 + lane * 4;
      if (row < P.batch && c < P.dim) Io4<TIn>::ld(xb + static_cast<int64_t>(row) * P.x_stride + c, v[rr][it]);
      else v[rr][it][0] = v[rr][it][1] = v[rr][it][2] = v[rr][it][3] = 0.f;
    }
  }
#pragma unroll
  for (int rr = 0; rr < kRows; ++rr) {
    const bool is_y = v0 + rr >= SM_TM;
    const int lr = is_y ? v0 + rr - SM_TM : v0 + rr, row = (is_y ? T.j0 : T.i0) + lr, m = is_y ? mb : ma;
    uint16_t* dst = is_y ? Ys : Xs;
    const bool write_state = is_y ? wy : wx;
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < 4; ++it)
#pragma unroll
      for (int e = 0; e < 4; ++e) ss = fmaf(v[rr][it][e], v[rr][it][e], ss);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = 1.f / fmaxf(sqrtf(ss), P.eps);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int c = it * 128 + lane * 4;
      if (c < P.dim) {
        uint2 pk;
        pk.x = pack2<kOp>(v[rr][it][0] * inv, v[rr][it][1] * inv);
        pk.y = pack2<kOp>(v[rr][it][2] * inv, v[rr][it][3] * inv);
        *reinterpret_cast<uint2*>(dst + lr * ld + c) = pk;
        if (write_state && row < P.batch)
          *reinterpret_cast<uint2*>(static_cast<uint16_t*>(P.z[m]) + static_cast<int64_t>(row) * P.dim + c) = pk;
      }
    }
    if (write_state && row < P.batch && lane == 0) P.inv[m][row] = inv;
  }
}

template <int kOp, typename TIn>
__global__ void __launch_bounds__(SM_THREADS, 1) ntxent_small_fwd_kernel(const __grid_constant__ SmallParams P) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int ld = P.dim + SM_PAD;
  uint16_t* Xs = reinterpret_cast<uint16_t*>(sm_raw);            // [32][ld]
  uint16_t* Ys = Xs + SM_TM * ld;                                // [64][ld]
  float* red_row = reinterpret_cast<float*>(Ys + SM_TN * ld);    // [8][32]
  float* red_col = red_row + 8 * SM_TM;                          // [2][64]
  double* red_d = reinterpret_cast<double*>(red_col + 2 * SM_TN);  // [2][SM_WARPS]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SmallTile T(P);
  SM_STAMP(0);
  // 96 rows over sixteen warps: two rounds of three rows (the registers hold 48 loaded values per lane and round)
  small_normalise_rows<kOp, TIn>(P, T, warp * 6, Xs, Ys, ld, lane);
  small_normalise_rows<kOp, TIn>(P, T, warp * 6 + 3, Xs, Ys, ld, lane);
  __syncthreads();
  SM_STAMP(1);

  const int wm = warp & 1, wn = warp >> 1, g = lane >> 2, tig = lane & 3;
  float acc[4];
  small_s_tile<kOp>(Xs, Ys, ld, P.dim, wm, wn, lane, acc);
  SM_STAMP(2);

  // sum-exp with the fixed shift 1/tau (|cos| <= 1), as ntxent_fwd.cu: e = 2^(c1 (s - 1))
  float rsum[2] = {0.f, 0.f}, csum[2] = {0.f, 0.f};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int row = wm * 16 + g + (e >> 1) * 8, col = wn * 8 + 2 * tig + (e & 1);
    const bool valid = T.i0 + row < P.batch && T.j0 + col < P.batch;
    const float ev = valid ? ex2_approx(fmaf(acc[e], P.c1, -P.c1)) : 0.f;
    rsum[e >> 1] += ev;
    csum[e & 1] += ev;
    if (T.i0 + row == T.j0 + col && valid) P.diag2[static_cast<int64_t>(T.pair) * P.batch + T.i0 + row] = acc[e] * P.c1;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    rsum[h] += __shfl_xor_sync(0xffffffffu, rsum[h], 1);
    rsum[h] += __shfl_xor_sync(0xffffffffu, rsum[h], 2);
    if (tig == 0) red_row[wn * SM_TM + wm * 16 + g + h * 8] = rsum[h];
  }
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    float c = csum[e];
    c += __shfl_xor_sync(0xffffffffu, c, 4);
    c += __shfl_xor_sync(0xffffffffu, c, 8);
    c += __shfl_xor_sync(0xffffffffu, c, 16);
    if (g == 0) red_col[wm * SM_TN + wn * 8 + 2 * tig + e] = c;
  }
  __syncthreads();
  {
    const int t = threadIdx.x;
    if (t < SM_TM) {
      if (T.i0 + t < P.batch) {
        float r = 0.f;
#pragma unroll
        for (int s = 0; s < 8; ++s) r += red_row[s * SM_TM + t];
        __stcg(P.row_part + (static_cast<int64_t>(T.pair) * P.nbj + T.bj) * P.batch + T.i0 + t, r);
      }
    } else if (t < SM_TM + SM_TN) {
      const int c = t - SM_TM;
      if (T.j0 + c < P.batch)
        __stcg(P.col_part + (static_cast<int64_t>(T.pair) * P.nbi + T.bi) * P.batch + T.j0 + c, red_col[c] + red_col[SM_TN + c]);
    }
  }
  SM_STAMP(3);
  small_grid_barrier(P.bar_lane);
  SM_STAMP(4);
  if (T.bi != 0 || T.bj != 0) return;

  // one CTA per pair: lse2 and the pair's loss; partials added in block order with every load in flight, fp64 sums in
  // a fixed order.  The last of these CTAs to finish adds the pair losses (fp32, pair order).
  const int p = T.pair;
  double sa = 0.0, sb = 0.0;
  for (int i = threadIdx.x; i < P.batch; i += SM_THREADS) {
    const float* rp = P.row_part + static_cast<int64_t>(p) * P.nbj * P.batch + i;
    const float* cp = P.col_part + static_cast<int64_t>(p) * P.nbi * P.batch + i;
    float rv[SM_MAX_SLOTS / 2], cv[SM_MAX_SLOTS];
#pragma unroll
    for (int u = 0; u < SM_MAX_SLOTS / 2; ++u) rv[u] = u < P.nbj ? __ldcg(rp + static_cast<int64_t>(u) * P.batch) : 0.f;
#pragma unroll
    for (int u = 0; u < SM_MAX_SLOTS; ++u) cv[u] = u < P.nbi ? __ldcg(cp + static_cast<int64_t>(u) * P.batch) : 0.f;
    const float dg = __ldcg(P.diag2 + static_cast<int64_t>(p) * P.batch + i);
    float rs = 0.f, cs = 0.f;
#pragma unroll
    for (int u = 0; u < SM_MAX_SLOTS / 2; ++u) rs += rv[u];
#pragma unroll
    for (int u = 0; u < SM_MAX_SLOTS; ++u) cs += cv[u];
    const float lr = log2f(rs) + P.c1, lc = log2f(cs) + P.c1;
    P.lse_row[static_cast<int64_t>(p) * P.batch + i] = lr;
    P.lse_col[static_cast<int64_t>(p) * P.batch + i] = lc;
    sa += static_cast<double>(lr - dg);
    sb += static_cast<double>(lc - dg);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sa += __shfl_xor_sync(0xffffffffu, sa, o);
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
  }
  if (lane == 0) {
    red_d[warp] = sa;
    red_d[SM_WARPS + warp] = sb;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double ta = 0.0, tb = 0.0;
    for (int w = 0; w < SM_WARPS; ++w) {
      ta += red_d[w];
      tb += red_d[SM_WARPS + w];
    }
    const double ln2 = 0.69314718055994530942;
    const double pa = ta * ln2, pbv = tb * ln2;
    P.loss_parts[p * 2 + 0] = static_cast<float>(pa);
    P.loss_parts[p * 2 + 1] = static_cast<float>(pbv);
    __stcg(P.loss + p, static_cast<float>((P.alpha * pa + (1.0 - P.alpha) * pbv) / P.batch));
    if (P.want_total) {
      unsigned int* cnt = g_small_fin + P.bar_lane;
      __threadfence();
      if (atomicAdd(cnt, 1u) + 1u == static_cast<unsigned int>(P.n_pairs)) {
        *cnt = 0u;  // nobody touches it again in this launch
        __threadfence();
        float total = 0.f;
        for (int q = 0; q < P.n_pairs; ++q) total += __ldcg(P.loss + q);  // sum(loss_dict.values()), tricolo_net.py:64
        P.loss[P.n_pairs] = total;
      }
    }
  }
  SM_STAMP(5);
}

template <int kOp, typename TIn>
__global__ void __launch_bounds__(SM_THREADS, 1) ntxent_small_bwd_kernel(const __grid_constant__ SmallParams P) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  const int ld = P.dim + SM_PAD;
  uint16_t* Xs = reinterpret_cast<uint16_t*>(sm_raw);                   // [32][ld]
  uint16_t* Ys = Xs + SM_TM * ld;                                       // [64][ld]
  uint16_t* Gs = Ys + SM_TN * ld;                                       // [32][SM_GLD]
  float* lr_s = reinterpret_cast<float*>(Gs + SM_TM * SM_GLD);          // [32]
  float* lc_s = lr_s + SM_TM;                                           // [64]
  float* dot_s = lc_s + SM_TN;                                          // [SM_WARPS]
  uint64_t* mbar = reinterpret_cast<uint64_t*>(dot_s + SM_WARPS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SmallTile T(P);
  const int a = P.pair_row[T.pair], b = P.pair_col[T.pair];
  const bool active = T.is_tile && (P.need_grad[a] || P.need_grad[b]);
  SM_STAMP(16);

  if (active) {
    // the two resident operand blocks: one bulk copy per row (rows are padded in shared memory), rows past the batch zero
    const uint32_t bar = smem_u32(mbar);
    const int rows_x = min(SM_TM, P.batch - T.i0), rows_y = min(SM_TN, P.batch - T.j0);
    const uint32_t row_bytes = static_cast<uint32_t>(P.dim) * 2u;
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      fence_mbar_init();
      mbar_arrive_expect_tx(bar, row_bytes * static_cast<uint32_t>(rows_x + rows_y));
    }
    __syncthreads();
    if (threadIdx.x < SM_TM + SM_TN) {
      const bool which = threadIdx.x >= SM_TM;
      const int r = which ? threadIdx.x - SM_TM : threadIdx.x;
      uint16_t* dst = (which ? Ys : Xs) + r * ld;
      if (r < (which ? rows_y : rows_x)) {
        const uint16_t* src = static_cast<const uint16_t*>(P.z[which ? b : a]) + static_cast<int64_t>((which ? T.j0 : T.i0) + r) * P.dim;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(dst)), "l"(src), "r"(row_bytes), "r"(bar) : "memory");
      } else {
        for (int c = 0; c < P.dim; c += 8) *reinterpret_cast<uint4*>(dst + c) = make_uint4(0u, 0u, 0u, 0u);
      }
    } else if (threadIdx.x < 128 + SM_TM) {
      if (threadIdx.x >= 128) {
        const int r = threadIdx.x - 128, i = T.i0 + r;
        lr_s[r] = i < P.batch ? P.lse_row[static_cast<int64_t>(T.pair) * P.batch + i] : 1e30f;
      }
    } else if (threadIdx.x >= 192 && threadIdx.x < 192 + SM_TN) {
      const int r = threadIdx.x - 192, j = T.j0 + r;
      lc_s[r] = j < P.batch ? P.lse_col[static_cast<int64_t>(T.pair) * P.batch + j] : 1e30f;
    }
  }
  // upstream gradient of pair p: d/d loss[p] plus d/d (sum of the losses); either may be absent
  const float g_tot = P.grad_total ? *P.grad_total : 0.f;
  float gmax = 0.f, g_pair = 0.f;
  for (int p = 0; p < P.n_pairs; ++p) {
    const float gp = (P.grad_losses ? P.grad_losses[p] : 0.f) + g_tot;
    gmax = fmaxf(gmax, fabsf(gp));
    if (p == T.pair) g_pair = gp;
  }
  if (active) {
    mbar_wait(smem_u32(mbar), 0);
    __syncthreads();
    SM_STAMP(17);

    const int wm = warp & 1, wn = warp >> 1, g = lane >> 2, tig = lane & 3;
    float acc[4];
    small_s_tile<kOp>(Xs, Ys, ld, P.dim, wm, wn, lane, acc);
    SM_STAMP(18);
    // G = w (alpha p_row + (1 - alpha) p_col - I), w = grad_scale kGScale / max|grad_scale| (ntxent_bwd.h)
    const float rr = gmax > 0.f ? g_pair * (kGScale / gmax) : 0.f;
    const float wr = rr * P.alpha, wc = rr * (1.f - P.alpha);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = wm * 16 + g + h * 8, col = wn * 8 + 2 * tig;
      float gv[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float s = acc[2 * h + e];
        const float pr = ex2_approx(fmaf(s, P.c1, -lr_s[row])), pc = ex2_approx(fmaf(s, P.c1, -lc_s[col + e]));
        const bool valid = T.i0 + row < P.batch && T.j0 + col + e < P.batch;
        float v = fmaf(wr, pr, wc * pc);
        if (T.i0 + row == T.j0 + col + e) v -= rr;
        gv[e] = valid ? v : 0.f;
      }
      *reinterpret_cast<uint32_t*>(Gs + row * SM_GLD + col) = pack2<kOp>(gv[0], gv[1]);
    }
    __syncthreads();
    SM_STAMP(19);
    const int64_t b_pad = static_cast<int64_t>(P.nbj) * SM_TN;
    if (P.need_grad[a]) {  // dZrow[32 x dim] = G (32 x 64) Zcol: two row tiles x eight 64-column chunks
      float* out = P.dz_part + ((static_cast<int64_t>(a) * P.max_slots + P.slot_base[T.pair][0] + T.bj) * b_pad + T.i0) * P.dim;
      small_grad_gemm<kOp, false, 4>(Gs, Ys, ld, P.dim, out, warp & 1, (warp >> 1) * 64, 512, lane);
    }
    if (P.need_grad[b]) {  // dZcol[64 x dim] = G^T (64 x 32) Zrow: four row tiles x four chunks, twice
      float* out = P.dz_part + ((static_cast<int64_t>(b) * P.max_slots + P.slot_base[T.pair][1] + T.bi) * b_pad + T.j0) * P.dim;
      small_grad_gemm<kOp, true, 2>(Gs, Xs, ld, P.dim, out, warp & 3, (warp >> 2) * 64, 256, lane);
    }
  }
  SM_STAMP(20);
  small_grid_barrier(P.bar_lane);
  SM_STAMP(21);

  // F.normalize backward (nt_xent.py:56-57): g = scale * sum of the row's partials in slot order;
  // dx = clamped ? g / ||x|| : (g - (g.z) z) / ||x||, z = x / ||x|| in fp32 (as l2norm_bwd_kernel).
  // A PAIR of warps per row (256 columns each), eight slots x two 128-column segments in flight per lane.
  const float scale = gmax * P.out_scale * (1.f / kGScale);
  const int64_t b_pad = static_cast<int64_t>(P.nbj) * SM_TN;
  const int n_items = P.n_tensors * P.batch;
  constexpr int kRowsPerCta = SM_WARPS / 2;
  const int wp = warp >> 1, wh = warp & 1;
  for (int base = blockIdx.x * kRowsPerCta; base < n_items; base += gridDim.x * kRowsPerCta) {
    const int item = base + wp;
    const bool live = item < n_items && P.need_grad[item < n_items ? item / P.batch : 0];
    const int m = live ? item / P.batch : 0, row = live ? item % P.batch : 0;
    const TIn* xr = static_cast<const TIn*>(P.x[m]) + static_cast<int64_t>(row) * P.x_stride;
    const float inv = live ? P.inv[m][row] : 0.f;
    const int ns = live ? P.n_slots[m] : 0;
    const float* part = P.dz_part + (static_cast<int64_t>(m) * P.max_slots * b_pad + row) * P.dim;
    float xv[2][4], gq[2][4], dot = 0.f;
    float4 s4[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int c = (wh * 2 + h) * 128 + lane * 4;
      if (live && c < P.dim) Io4<TIn>::ld(xr + c, xv[h]);
      else xv[h][0] = xv[h][1] = xv[h][2] = xv[h][3] = 0.f;
    }
    for (int s0 = 0; s0 < ns; s0 += 8) {
      float4 pv[2][8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = (wh * 2 + h) * 128 + lane * 4;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          pv[h][u] = (s0 + u < ns && c < P.dim) ? __ldcg(reinterpret_cast<const float4*>(part + (s0 + u) * b_pad * P.dim + c))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          s4[h].x += pv[h][u].x; s4[h].y += pv[h][u].y; s4[h].z += pv[h][u].z; s4[h].w += pv[h][u].w;
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      gq[h][0] = s4[h].x * scale; gq[h][1] = s4[h].y * scale; gq[h][2] = s4[h].z * scale; gq[h][3] = s4[h].w * scale;
#pragma unroll
      for (int e = 0; e < 4; ++e) dot = fmaf(gq[h][e], xv[h][e] * inv, dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) dot_s[warp] = dot;
    __syncthreads();
    dot = dot_s[wp * 2] + dot_s[wp * 2 + 1];  // the same order in both warps of the pair
    __syncthreads();
    if (inv >= 1.f / P.eps) dot = 0.f;  // clamped row: F.normalize divides by eps, a constant
    if (live) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = (wh * 2 + h) * 128 + lane * 4;
        if (c < P.dim) {
          float o4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) o4[e] = (gq[h][e] - dot * (xv[h][e] * inv)) * inv;
          Io4<TIn>::st(static_cast<TIn*>(P.dx[m]) + static_cast<int64_t>(row) * P.dim + c, o4);
        }
      }
    }
  }
  SM_STAMP(22);
}

}  // namespace tcl
extern "C" int tcl_debug_small_trace(unsigned long long* out32) {
  using namespace tcl;
  TCL_CHECK_CUDA(cudaDeviceSynchronize());
  TCL_CHECK_CUDA(cudaMemcpyFromSymbol(out32, g_small_trace, sizeof(unsigned long long) * 32));
  return TCL_OK;
}
namespace tcl {
// ---------------------------------------------------------------------------------------------------------------
static std::atomic<unsigned> g_next_bar_lane{0};

static inline int small_nbi(int64_t batch) { return static_cast<int>((batch + SM_TM - 1) / SM_TM); }
static inline int small_nbj(int64_t batch) { return static_cast<int>((batch + SM_TN - 1) / SM_TN); }
static inline size_t small_smem_fwd(int64_t dim) {
  return (SM_TM + SM_TN) * (dim + SM_PAD) * 2 + (8 * SM_TM + 2 * SM_TN) * 4 + 2 * TCL_MAX_PAIRS * SM_WARPS * 8;
}
static inline size_t small_smem_bwd(int64_t dim) {
  return (SM_TM + SM_TN) * (dim + SM_PAD) * 2 + SM_TM * SM_GLD * 2 + (SM_TM + SM_TN + SM_WARPS) * 4 + 16;
}
static inline int small_max_slots(int n_pairs, int64_t batch) {
  // a tensor is in at most two pairs; as a pair's row tensor it gets one partial per column block, as its column
  // tensor one per (smaller) row block
  return (n_pairs > 1 ? 2 : 1) * small_nbi(batch);
}

// TRICOLO_B200_SMALL=0 keeps the multi-kernel pipeline at every size (read per call)
bool small_enabled(int n_tensors, int n_pairs, int64_t batch, int64_t dim, int op_format) {
  const char* e = getenv("TRICOLO_B200_SMALL");
  if (e && e[0] == '0') return false;
  if (op_format != TCL_OP_F16 && op_format != TCL_OP_BF16) return false;
  if (dim % 64 != 0 || dim < 64 || dim > 512 || batch < 1 || batch > 4096) return false;
  if (n_tensors > TCL_MAX_TENSORS || n_pairs > TCL_MAX_PAIRS) return false;
  if (small_max_slots(n_pairs, batch) > SM_MAX_SLOTS) return false;
  int dev = 0, n_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return false;
  if (n_sm >= SM_BAR_MAXN) return false;
  // one tile per CTA, one CTA per SM, all co-resident
  return static_cast<int64_t>(n_pairs) * small_nbi(batch) * small_nbj(batch) <= n_sm;
}
size_t small_fwd_workspace_bytes(int n_pairs, int64_t batch) {
  return static_cast<size_t>(n_pairs) * (small_nbi(batch) + small_nbj(batch)) * batch * sizeof(float);
}
size_t small_bwd_workspace_bytes(int n_tensors, int64_t batch, int64_t dim) {
  return static_cast<size_t>(n_tensors) * small_max_slots(n_tensors > 2 ? 3 : 1, batch) * small_nbj(batch) * SM_TN * dim * sizeof(float);
}

template <typename K>
static int small_launch(K kernel, const SmallParams& P, int grid, size_t smem, cudaStream_t st) {
  if (int e = ensure_dyn_smem(kernel, static_cast<int>(smem))) return e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(SM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  // (programmatic dependent launch on top of the cooperative attribute is accepted but gains nothing: measured)
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // the grid barrier needs every CTA resident
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  TCL_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, P));
  return TCL_OK;
}

template <bool kFwd, int kOp, typename TIn>
static int small_dispatch2(const SmallParams& P, int grid, size_t smem, cudaStream_t st) {
  if (kFwd) return small_launch(ntxent_small_fwd_kernel<kOp, TIn>, P, grid, smem, st);
  return small_launch(ntxent_small_bwd_kernel<kOp, TIn>, P, grid, smem, st);
}
template <bool kFwd, int kOp>
static int small_dispatch1(const SmallParams& P, int x_dtype, int grid, size_t smem, cudaStream_t st) {
  switch (x_dtype) {
    case TCL_DT_F32: return small_dispatch2<kFwd, kOp, float>(P, grid, smem, st);
    case TCL_DT_F64: return small_dispatch2<kFwd, kOp, double>(P, grid, smem, st);
    case TCL_DT_F16: return small_dispatch2<kFwd, kOp, __half>(P, grid, smem, st);
    case TCL_DT_BF16: return small_dispatch2<kFwd, kOp, __nv_bfloat16>(P, grid, smem, st);
  }
  TCL_REQUIRE(false, TCL_ERR_BAD_ARG, "loss: x_dtype %d", x_dtype);
  return TCL_OK;
}
template <bool kFwd>
static int small_dispatch(const SmallParams& P, int op_format, int x_dtype, int grid, size_t smem, cudaStream_t st) {
  return op_format == TCL_OP_F16 ? small_dispatch1<kFwd, TCL_OP_F16>(P, x_dtype, grid, smem, st)
                                 : small_dispatch1<kFwd, TCL_OP_BF16>(P, x_dtype, grid, smem, st);
}

static int small_fill_common(SmallParams& P, int n_tensors, const void* const* x, int64_t batch, int64_t dim,
                             int64_t x_row_stride, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                             float inv_tau, float alpha, float eps) {
  memset(&P, 0, sizeof(P));
  P.n_tensors = n_tensors;
  P.n_pairs = n_pairs;
  P.batch = static_cast<int>(batch);
  P.dim = static_cast<int>(dim);
  P.nbi = small_nbi(batch);
  P.nbj = small_nbj(batch);
  P.n_tiles = n_pairs * P.nbi * P.nbj;
  P.x_stride = x_row_stride;
  P.c1 = inv_tau * 1.4426950408889634f;
  P.alpha = alpha;
  P.eps = eps;
  P.out_scale = inv_tau / static_cast<float>(batch);
  P.max_slots = small_max_slots(n_tensors > 2 ? 3 : 1, batch);
  P.bar_lane = static_cast<int>(g_next_bar_lane.fetch_add(1u) % SM_BAR_LANES);
  int n_seg[TCL_MAX_TENSORS] = {0};
  for (int m = 0; m < n_tensors; ++m) {
    P.x[m] = x[m];
    P.writer_pair[m] = -1;
  }
  for (int p = 0; p < n_pairs; ++p) {
    const int a = pair_row[p], b = pair_col[p];
    TCL_REQUIRE(a >= 0 && a < n_tensors && b >= 0 && b < n_tensors && a != b, TCL_ERR_BAD_ARG, "loss: pair %d out of range", p);
    TCL_REQUIRE(n_seg[a] < 2 && n_seg[b] < 2, TCL_ERR_BAD_ARG, "loss: a tensor takes part in more than two pairs (pair %d)", p);
    P.pair_row[p] = a;
    P.pair_col[p] = b;
    if (P.writer_pair[a] < 0) P.writer_pair[a] = p;
    if (P.writer_pair[b] < 0) P.writer_pair[b] = p;
    P.slot_base[p][0] = P.n_slots[a];
    P.slot_base[p][1] = P.n_slots[b];
    P.n_slots[a] += P.nbj;  // one partial per column block
    P.n_slots[b] += P.nbi;  // one partial per row block
    ++n_seg[a];
    ++n_seg[b];
  }
  for (int m = 0; m < n_tensors; ++m)
    TCL_REQUIRE(P.n_slots[m] <= P.max_slots, TCL_ERR_BAD_SHAPE, "loss: %d partial slots for tensor %d", P.n_slots[m], m);
  return TCL_OK;
}

int launch_small_fwd(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim, int64_t x_row_stride,
                     int n_pairs, const int32_t* pair_row, const int32_t* pair_col, int op_format, float inv_tau,
                     float alpha, float eps, void* const* z, float* const* inv, float* diag2, float* lse_row,
                     float* lse_col, float* loss_parts, float* loss, bool want_total, void* workspace,
                     size_t workspace_bytes, cudaStream_t st) {
  if (int e = require_sm100()) return e;
  TCL_REQUIRE(inv_tau > 0.f && 2.f * inv_tau * 1.4426950408889634f < 120.f, TCL_ERR_BAD_ARG,
              "ntxent_fwd: temperature %g too small for the fixed-shift sum-exp (need tau >= 0.025)", 1.0 / inv_tau);
  TCL_REQUIRE(workspace_bytes >= small_fwd_workspace_bytes(n_pairs, batch), TCL_ERR_WORKSPACE, "loss_fwd: workspace too small");
  SmallParams P;
  if (int e = small_fill_common(P, n_tensors, x, batch, dim, x_row_stride, n_pairs, pair_row, pair_col, inv_tau, alpha, eps)) return e;
  for (int m = 0; m < n_tensors; ++m) {
    P.z[m] = z[m];
    P.inv[m] = inv[m];
  }
  P.row_part = static_cast<float*>(workspace);
  P.col_part = P.row_part + static_cast<size_t>(n_pairs) * P.nbj * batch;
  P.diag2 = diag2;
  P.lse_row = lse_row;
  P.lse_col = lse_col;
  P.loss_parts = loss_parts;
  P.loss = loss;
  P.want_total = want_total ? 1 : 0;
  prof_begin(TCL_K_NTXENT_SMALL_FWD, st);
  const int e = small_dispatch<true>(P, op_format, x_dtype, P.n_tiles, small_smem_fwd(dim), st);
  prof_end(TCL_K_NTXENT_SMALL_FWD, st);
  return e;
}

int launch_small_bwd(int n_tensors, const void* const* x, int x_dtype, int64_t batch, int64_t dim, int64_t x_row_stride,
                     int n_pairs, const int32_t* pair_row, const int32_t* pair_col, int op_format, float inv_tau,
                     float alpha, float eps, const void* const* z, const float* inv_base, const float* lse_row,
                     const float* lse_col, const float* grad_losses, const float* grad_total, const uint8_t* need_grad,
                     void* const* dx, void* workspace, size_t workspace_bytes, cudaStream_t st) {
  if (int e = require_sm100()) return e;
  TCL_REQUIRE(workspace_bytes >= small_bwd_workspace_bytes(n_tensors, batch, dim), TCL_ERR_WORKSPACE, "loss_bwd: workspace too small");
  SmallParams P;
  if (int e = small_fill_common(P, n_tensors, x, batch, dim, x_row_stride, n_pairs, pair_row, pair_col, inv_tau, alpha, eps)) return e;
  bool any = false;
  for (int m = 0; m < n_tensors; ++m) {
    P.z[m] = const_cast<void*>(z[m]);
    P.inv[m] = const_cast<float*>(inv_base) + static_cast<size_t>(m) * batch;
    P.need_grad[m] = need_grad[m] ? 1 : 0;
    TCL_REQUIRE(!need_grad[m] || dx[m] != nullptr, TCL_ERR_BAD_ARG, "loss_bwd: dx[%d] is null", m);
    P.dx[m] = dx[m];
    any = any || (need_grad[m] && P.n_slots[m] > 0);
  }
  if (!any) return TCL_OK;
  P.lse_row = const_cast<float*>(lse_row);
  P.lse_col = const_cast<float*>(lse_col);
  P.grad_losses = grad_losses;
  P.grad_total = grad_total;
  P.dz_part = static_cast<float*>(workspace);
  // tiles first; further CTAs (up to one per SM) only take rows of the final phase, which is one L2 round trip per row
  int grid = P.n_tiles, dev = 0, n_sm = 0;
  TCL_CHECK_CUDA(cudaGetDevice(&dev));
  TCL_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const int rows_per_cta = SM_WARPS / 2;
  const int want = static_cast<int>((static_cast<int64_t>(n_tensors) * batch + rows_per_cta - 1) / rows_per_cta);
  if (grid < want) grid = want < n_sm ? want : n_sm;
  prof_begin(TCL_K_NTXENT_SMALL_BWD, st);
  const int e = small_dispatch<false>(P, op_format, x_dtype, grid, small_smem_bwd(dim), st);
  prof_end(TCL_K_NTXENT_SMALL_BWD, st);
  return e;
}

}  // namespace tcl
