/*
 * tricolo_b200 — C ABI of the sm_100a embedding-similarity hot path.
 *
 * This is the drop-in boundary for the two reference entry points
 *   tricolo/loss/nt_xent.py:24        NTXentLoss.forward(zis, zjs, norm=True)
 *   tricolo/evaluation/eval_retrieval.py:249  compute_metrics(dataset, embeddings_dict)
 * (and the callers tricolo/model/tricolo_net.py:56-65 / :90-97).  The reference has no
 * FFI of its own (pure PyTorch / NumPy); the Python host in
 * tricolo_b200/ binds these symbols with ctypes, INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the name says host; the caller allocates every buffer, including
 *     workspaces (query the *_workspace_bytes functions); the library never
 *     allocates, frees or retains device memory.
 *   - `stream` is a cudaStream_t passed as void*; every call only enqueues work
 *     (no hidden synchronisation).
 *   - return value: 0 = OK, otherwise one of TCL_ERR_* (>= 1000: 1000 +
 *     cudaError_t); tcl_last_error_string() gives the text (thread-local).
 *   - sm_100 only: any other device returns TCL_ERR_BAD_ARCH.  There is no
 *     fallback path of any kind.
 */
#ifndef TRICOLO_B200_H_
#define TRICOLO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TCL_ABI_VERSION 1

enum {
  TCL_OK = 0,
  TCL_ERR_BAD_SHAPE = 1,
  TCL_ERR_BAD_ALIGN = 2,
  TCL_ERR_BAD_ARCH = 3,
  TCL_ERR_BAD_ARG = 4,
  TCL_ERR_DRIVER = 5,
  TCL_ERR_WORKSPACE = 6,
  TCL_ERR_CUDA_BASE = 1000
};

/* 16-bit tensor-core operand formats (tcgen05 kind::f16 takes either). */
enum { TCL_OP_F16 = 0, TCL_OP_BF16 = 1 };
/* element types of user tensors */
enum { TCL_DT_F32 = 0, TCL_DT_F16 = 1, TCL_DT_BF16 = 2, TCL_DT_F64 = 3 };

#define TCL_MAX_TENSORS 3 /* text, image, voxel */
#define TCL_MAX_PAIRS 3   /* (text,image) (text,voxel) (image,voxel): tricolo_net.py:59-63 */

int tcl_version(void);
const char* tcl_last_error_string(void);

/* ---------------------------------------------------------------------------
 * K1 — L2-normalise prologue.           replaces nt_xent.py:56-57 (F.normalize)
 * For each of n_tensors row-major [rows, dim] inputs: inv_norm[r] = 1/max(||x_r||, eps),
 * z[r,:] = x[r,:] * inv_norm[r] rounded to the 16-bit operand format.
 * x_row_stride in elements.  dim % 8 == 0.
 * z_row_stride: row stride of every 16-bit operand matrix in elements (0 = dim, i.e. contiguous;
 * otherwise >= dim and a multiple of 8).  A stride lets several modalities share one
 * [rows, M*dim] buffer so that a single all-gather moves them all (tricolo_b200/distributed.py).
 * ------------------------------------------------------------------------- */
int tcl_l2norm_fwd(int n_tensors, const void* const* x_host_ptrs, int x_dtype, int64_t rows,
                   int64_t dim, int64_t x_row_stride, void* const* z_host_ptrs, int64_t z_row_stride,
                   int op_format, float* const* inv_norm_host_ptrs, float eps, void* stream);

/* K1 fused with the all-gather of the sharded loss (tricolo_b200/distributed.py): the same computation as
 * tcl_l2norm_fwd, but every normalised row is stored to n_dst destinations, dst r being rank r's copy of the gathered
 * operand buffer (peer-mapped device memory, e.g. a torch symmetric-memory allocation: plain st.global over NVLink).
 *   z_dst_host_ptrs[r * n_tensors + m] = address of THIS rank's first row of modality m inside rank r's buffer.
 * No synchronisation: the caller brackets the call with cross-device barriers (nobody still reads the buffers before,
 * every rank's rows have landed after).  n_dst <= TCL_MAX_PEERS.  Destination 0 is mandatory for every tensor; a NULL
 * address at r > 0 skips that (destination, tensor): a modality whose remote rows nobody reads (the text modality
 * under the sharded shared-G backward, or in a forward without gradients) then costs no NVLink traffic. */
#define TCL_MAX_PEERS 8
int tcl_l2norm_fwd_bcast(int n_tensors, const void* const* x_host_ptrs, int x_dtype, int64_t rows, int64_t dim,
                         int64_t x_row_stride, int n_dst, void* const* z_dst_host_ptrs, int64_t z_row_stride,
                         int op_format, float* const* inv_norm_host_ptrs, float eps, void* stream);

/* ---------------------------------------------------------------------------
 * Sharded forward without barriers (tricolo_b200/distributed.py; semantics: nt_xent.py:55-74 on the concatenated
 * global batch).  Every rank owns a zero-initialised, peer-mapped "sync pad" of
 * tcl_shard_sync_bytes() (flags + step counters, tricolo_b200/csrc/host_common.h: ShardSync) and a peer-mapped
 * statistics buffer of tcl_shard_stats_bytes(); sync_host_ptrs[r] / stats_host_ptrs[r] are rank r's buffers as
 * mapped into this process.  Per step, in stream order on every rank:
 *   tcl_l2norm_fwd_push      K1: normalised rows into the own gathered buffer + the flag "my operand buffer may be
 *                            overwritten" to every peer at kernel start.  remote = 1: also the all-gather (as
 *                            tcl_l2norm_fwd_bcast) with the flag "chunk c of my rows has landed" per 128 rows;
 *                            remote = 0: the tile kernel's push warps do that while its MMAs run (n_push > 0 below)
 *   tcl_ntxent_fwd_sharded   K2 on the local row block against all b_glob columns; column tiles are consumed in
 *                            arrival order, each gated on its chunk's flag, so the NVLink gather overlaps the sweep.
 *                            n_push > 0: the all-gather is FUSED into this kernel - two extra warps per CTA copy the
 *                            rank's rows of the modalities at element offsets push_offsets[] of a gathered row into
 *                            every rank's gathered buffer (z_base_host_ptrs[r]) while the MMAs run, and flag each
 *                            128-row chunk.  Then the sum-exp statistics of this rank are stored into every rank's
 *                            statistics buffer + flag
 *   tcl_ntxent_finalize_sharded  waits for the W statistics flags, adds the column partials in rank order and
 *                            finalises all rows: lse2_row [P][b_glob], lse2_col [P][b_glob], loss [P] (identical
 *                            on every rank)
 * No host synchronisation, no barrier kernel, capturable in a CUDA graph.  rows per rank % 128 == 0, <= 8192.
 * Barrier form of the same two calls (sync pointers NULL; measured faster on 8 B200: system-scope fences are dear):
 * tcl_ntxent_fwd_sharded sweeps contiguous tile ranges without gating (the caller ran a barrier after the gather) and
 * writes this rank's statistics into slot `rank` of its OWN buffer; after a barrier, tcl_ntxent_finalize_sharded
 * PULLS slot s from rank s's buffer over NVLink (W x ~100 KB) - one barrier and two kernels for the statistics
 * exchange instead of zero/copy kernels + barrier + tcl_peer_sum_f32 + tcl_ntxent_finalize.
 * ------------------------------------------------------------------------- */
size_t tcl_shard_sync_bytes(void);
size_t tcl_shard_stats_bytes(int n_pairs, int64_t b_loc, int world);
int tcl_l2norm_fwd_push(int n_tensors, const void* const* x_host_ptrs, int x_dtype, int64_t rows, int64_t dim,
                        int64_t x_row_stride, int rank, int world, void* const* z_dst_host_ptrs, int64_t z_row_stride,
                        int op_format, float* const* inv_norm_host_ptrs, float eps, void* const* sync_host_ptrs,
                        int remote, void* stream);
int tcl_ntxent_fwd_sharded(int n_pairs, const void* const* zrow_host_ptrs, const void* const* zcol_host_ptrs,
                           int64_t b_loc, int64_t b_glob, int64_t dim, int64_t z_row_stride, int rank, int world,
                           int op_format, float inv_tau, float* diag2, void* workspace, size_t workspace_bytes,
                           void* const* stats_host_ptrs, void* const* sync_host_ptrs, void* const* z_base_host_ptrs,
                           int n_push, const int64_t* push_offsets_host, void* stream);
int tcl_ntxent_finalize_sharded(int n_pairs, int64_t b_loc, int64_t b_glob, int rank, int world, float inv_tau,
                                float alpha, void* const* stats_host_ptrs, const void* sync_own, float* lse2_row,
                                float* lse2_col, float* loss, void* stream);

/* out[i] = sum over r < n_src (in that order) of src[r][i], fp32: the one-shot all-reduce of the sum-exp statistics
 * over peer-mapped buffers (every rank reads all ranks' partials and adds them in rank order, so all ranks get
 * bit-identical sums).  16-byte aligned pointers; any n. */
int tcl_peer_sum_f32(int n_src, const float* const* src_host_ptrs, int64_t n, float* out, void* stream);

/* Copy-engine transfer of `rows` rows of `width_bytes` between two pitched device buffers (cudaMemcpy2DAsync, device to
 * device; either side may be peer-mapped memory of another GPU).  The sharded loss uses it to send the rows of a modality
 * that only the BACKWARD reads remotely (text under the directional backward) to the peers on a side stream while the
 * forward tile kernel runs: no SM time, NVLink busy during the MMAs (tricolo_b200/distributed.py). */
int tcl_copy_rows(void* dst, int64_t dst_pitch_bytes, const void* src, int64_t src_pitch_bytes, int64_t width_bytes,
                  int64_t rows, void* stream);

/* 16-bit cast without normalisation (retrieval uses the raw dot product,
 * eval_retrieval.py:74).  Accepts f32/f64/f16/bf16 input. */
int tcl_cast_16bit(const void* x, int x_dtype, int64_t rows, int64_t dim, int64_t x_row_stride,
                   void* y, int op_format, void* stream);

/* Gallery build of the device-resident evaluation hand-off: replaces the host loops of
 * TriCoLoNet._collate_output (tricolo_net.py:125-158: shape = zeros + image + voxel, fp32) and the
 * first-occurrence de-duplication of construct_embeddings_matrix (eval_retrieval.py:49-56):
 *   out16[g, :] = round16( src[0][index[g], :] (+ src[1][index[g], :]) )      g < n_out
 * The sum is formed in fp32 in the reference's order.  n_src in {1, 2}; all sources share dtype,
 * shape [n_src_rows, dim] and row stride; index values must lie in [0, n_src_rows). dim % 8 == 0. */
int tcl_gather_sum_cast16(int n_src, const void* const* src_host_ptrs, int src_dtype, int64_t n_src_rows,
                          int64_t dim, int64_t src_row_stride, const int64_t* index, int64_t n_out,
                          void* out16, int op_format, void* stream);

/* [rows, dim] 16-bit -> [dim, ld_t] 16-bit (ld_t >= rows, ld_t % 8 == 0); operand
 * layout of the gradient GEMM in tcl_ntxent_bwd. */
int tcl_transpose_16bit(int n_tensors, const void* const* z_host_ptrs, int64_t rows, int64_t dim,
                        int64_t z_row_stride, void* const* zt_host_ptrs, int64_t ld_t, void* stream);

/* ---------------------------------------------------------------------------
 * K2 — similarity GEMM + fused sum-exp epilogue.   replaces nt_xent.py:62-72 forward
 * For each pair p: S = Zrow_p [n_rows, dim] · Zcol_p [n_cols, dim]^T on tcgen05 (fp32
 * accumulate in TMEM); the logits never leave the SM.  With c1 = log2(e)/tau:
 *   row_sumexp[p][i] = sum_j 2^(c1*S_ij - c1)      (complete for the given rows)
 *   col_sumexp[p][j] = sum_i 2^(c1*S_ij - c1)      (partial: only the given rows)
 *   diag2[p][i]      = c1 * S_{i, row_offset+i}    (log2-domain logit of the positive)
 * The fixed shift c1 is exact for normalised inputs (|S| <= 1); requires
 * 2/tau*log2(e) < 120, i.e. tau >= 0.025 (checked).
 * n_rows = rows held locally (global index row_offset + i), n_cols = global batch.
 * dim % 64 == 0, dim <= 512.
 * ------------------------------------------------------------------------- */
size_t tcl_ntxent_fwd_workspace_bytes(int n_pairs, int64_t n_rows, int64_t n_cols);
int tcl_ntxent_fwd(int n_pairs, const void* const* zrow_host_ptrs,
                   const void* const* zcol_host_ptrs, int64_t n_rows, int64_t n_cols, int64_t dim,
                   int64_t z_row_stride, int64_t row_offset, int op_format, float inv_tau, float* row_sumexp,
                   float* col_sumexp, float* diag2, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Finalise: log2-domain LSEs and the per-pair loss.     nt_xent.py:20-21,71-74
 *   lse2_row[p][i] = log2(row_sumexp) + c1,  lse2_col[p][j] likewise (needs the
 *   all-reduced col_sumexp when rows are sharded over ranks)
 *   loss_parts[p][0] = ln2 * sum_i (lse2_row_i - diag2_i)
 *   loss_parts[p][1] = ln2 * sum_i (lse2_col_{row_offset+i} - diag2_i)
 *   loss[p] = (alpha*parts0 + (1-alpha)*parts1) / n_cols     (only meaningful when the
 *             call holds all rows; sharded callers all-reduce loss_parts instead) */
int tcl_ntxent_finalize(int n_pairs, int64_t n_rows, int64_t n_cols, int64_t row_offset,
                        float inv_tau, float alpha, const float* row_sumexp,
                        const float* col_sumexp, const float* diag2, float* lse2_row,
                        float* lse2_col, float* loss_parts, float* loss, void* stream);

/* ---------------------------------------------------------------------------
 * K3 — recompute-based backward.       replaces autograd of nt_xent.py:55-74
 * One "job" per tensor that needs a gradient; a job has 1-2 "segments", one per
 * pair the tensor takes part in.  For a segment the kernel re-forms the logit
 * tile S = Zself · Zother^T, builds
 *   G_ij = w_self * 2^(c1 S_ij - lse2_self_i) + w_other * 2^(c1 S_ij - lse2_other_j) - [i==j]
 * in registers, writes it as a 16-bit operand tile to shared memory and
 * accumulates  dZself += G · Zother  in TMEM.  Then (second kernel) the
 * normalise backward  dx = (g - (g·z) z) * inv_norm,  g = grad_scale/(tau*n_other) * acc.
 * w_self / lse2_self belong to the softmax taken along the OTHER index for a fixed
 * self row (row softmax when self is the pair's first argument: w = alpha).
 * ------------------------------------------------------------------------- */
typedef struct {
  const void* z_other;     /* [n_other, dim] 16-bit */
  const void* z_other_t;   /* [dim, ld_t]   16-bit (tcl_transpose_16bit); may be NULL when
                              tcl_ntxent_bwd_needs_transpose(dim) == 0 */
  const float* lse2_self;  /* [n_self]  */
  const float* lse2_other; /* [n_other] */
  const float* grad_scale; /* device scalar dL/d(loss of this pair); NULL = 1 */
  float w_self;
  float w_other;
} tcl_bwd_segment;

typedef struct {
  const void* z_self;    /* [n_self, dim] 16-bit */
  const void* x_self;    /* original input rows [n_self, dim], x_dtype */
  const float* inv_norm; /* [n_self] from tcl_l2norm_fwd */
  void* dx;              /* [n_self, dim] output, x_dtype */
  int32_t n_segments;
  int32_t reserved;
  tcl_bwd_segment seg[2];
} tcl_bwd_job;

size_t tcl_ntxent_bwd_workspace_bytes(int n_jobs, int64_t n_self, int64_t dim);
/* 1 if tcl_ntxent_bwd needs the transposed copies z_other_t / ld_t for this dim (the kernels for dim <= 256 and the
 * non-default kernels selected by TRICOLO_B200_BWD), 0 if it reads the row-major operands only. */
int tcl_ntxent_bwd_needs_transpose(int64_t dim);
int tcl_ntxent_bwd(int n_jobs, const tcl_bwd_job* jobs_host, int64_t n_self, int64_t n_other,
                   int64_t dim, int64_t z_row_stride, int64_t self_offset, int64_t ld_t, int x_dtype,
                   int64_t x_row_stride, int op_format, float inv_tau, float eps,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * K3, sharded over ranks (north_star: "each rank owns a row block of the logits, and gradients
 * are reduce-scattered").  Semantics: the backward of the reference loss on the CONCATENATED global batch
 * (nt_xent.py:68-74 per pair, pairs as tricolo_net.py:56-65), rank r holding rows [r*b_loc, (r+1)*b_loc).
 *   gemm    forms the rank's row block G[b_loc x b_glob] of every pair once (16-bit, in `workspace`), then
 *           dRow = G * Zcol for the local rows (local partials) and dCol = G^T * Zrow_local for ALL b_glob rows of
 *           the column tensor; every 128-row piece of dCol is TMA-stored from the kernel's accumulator drain into the
 *           OWNER rank's receive buffer (recv_ptrs[owner], peer-mapped memory: the reduce-scatter is the kernel's own
 *           store traffic over NVLink, one slot per source rank, no atomics).  6*b_loc*b_glob*dim flop per pair.
 *   finish  sums the local and the received partials in a fixed order and applies the normalise backward ->
 *           dx[m] [b_loc, dim].  It must not read before every rank's gemm has completed: either the caller runs a
 *           cross-rank barrier between the two calls (sync pointers NULL), or both calls get the ranks' sync pads
 *           (tcl_shard_sync_bytes): the last CTA of each gemm kernel then signals every rank and finish waits on
 *           the device for the W flags - no barrier kernel, capturable in a CUDA graph.
 *   z_all[m]      gathered 16-bit operands [b_glob, dim] (row stride z_row_stride elements), identical on all ranks
 *   lse_row/col   [n_pairs][b_glob] log2-domain LSEs of all rows / columns (tcl_ntxent_finalize)
 *   recv_ptrs[r]  rank r's receive buffer as mapped into this process (recv_ptrs[rank] = the own one), each
 *                 tcl_ntxent_bwd_sharded_recv_bytes() large, 256-byte aligned
 *   workspace     local scratch, tcl_ntxent_bwd_sharded_workspace_bytes(), 256-byte aligned, kept until finish
 * Requires 256 < dim <= 512, b_loc % 128 == 0, world <= TCL_MAX_PEERS, the same sizes, pair list and need_grad flags on
 * every rank (the slot a partial lands in is a pure function of them).
 * ------------------------------------------------------------------------- */
size_t tcl_ntxent_bwd_sharded_workspace_bytes(int n_tensors, int n_pairs, const int32_t* pair_row,
                                              const int32_t* pair_col, const uint8_t* need_grad_host, int64_t b_loc,
                                              int64_t b_glob, int64_t dim, int world);
size_t tcl_ntxent_bwd_sharded_recv_bytes(int n_tensors, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                                         const uint8_t* need_grad_host, int64_t b_loc, int64_t b_glob, int64_t dim,
                                         int world);
int tcl_ntxent_bwd_sharded_gemm(int n_tensors, const void* const* z_all_host_ptrs, int64_t b_loc, int64_t b_glob,
                                int64_t dim, int64_t z_row_stride, int rank, int world, int n_pairs,
                                const int32_t* pair_row, const int32_t* pair_col, int op_format, float inv_tau,
                                float alpha, const float* lse_row, const float* lse_col, const float* grad_losses,
                                const uint8_t* need_grad_host, void* workspace, size_t workspace_bytes,
                                void* const* recv_host_ptrs, size_t recv_bytes, void* const* sync_host_ptrs,
                                void* stream);
int tcl_ntxent_bwd_sharded_finish(int n_tensors, const void* const* x_host_ptrs, int x_dtype, int64_t b_loc,
                                  int64_t b_glob, int64_t dim, int64_t x_row_stride, int rank, int world, int n_pairs,
                                  const int32_t* pair_row, const int32_t* pair_col, const float* inv_norm,
                                  const uint8_t* need_grad_host, float eps, const void* workspace,
                                  const void* recv_own, const void* sync_own, void* const* dx_host_ptrs, void* stream);

/* ---------------------------------------------------------------------------
 * Triplet loss: TripletLoss.forward(zis, zls) of tricolo/loss/triplet.py:202-224 with
 * _pairwise_distances (:11-45) and their autograd; selected by loss.name=TripletLoss (config/config.yaml:102-104).
 * Plain fp32 on a materialised B x B distance matrix (training batch sizes; the margin is below 16-bit operand
 * resolution).  Semantics, including the reference's pairing of the norms (:32):
 *   d2[a][b] = |zls_b|^2 - 2 <zls_a, zis_b> + |zis_a|^2, D = sqrt(max(d2,0));
 *   terms D[i][i] - D[i][j] + margin over j != i with D[i][i] < D[i][j] < D[i][i] + margin (semi-hard), or, when there
 *   is none, over D[i][j] < D[i][i] (hard); loss = mean of the terms.
 * info_out (device int32[4]) = {n_semi_hard, n_hard, mode (0 semi-hard, 1 hard, 2 no term: loss = NaN, the reference
 * divides by zero), n_terms_used}.  The workspace carries D and info from the forward to the backward.
 * d_zis / d_zls: [batch, dim] contiguous, input dtype; either may be NULL.
 * ------------------------------------------------------------------------- */
size_t tcl_triplet_workspace_bytes(int64_t batch);
int tcl_triplet_fwd(const void* zis, const void* zls, int x_dtype, int64_t batch, int64_t dim, int64_t row_stride,
                    float margin, float* loss, int32_t* info_out, void* workspace, size_t workspace_bytes, void* stream);
int tcl_triplet_bwd(const void* zis, const void* zls, int x_dtype, int64_t batch, int64_t dim, int64_t row_stride,
                    float margin, const float* grad_loss, void* workspace, size_t workspace_bytes, void* d_zis,
                    void* d_zls, void* stream);

/* ---------------------------------------------------------------------------
 * Whole-loss entry points for one GPU: the kernels above sequenced by the library, so that a
 * training step is two calls.  Replaces TriCoLoNet._calculate_losses (tricolo_net.py:56-65) +
 * NTXentLoss.forward (nt_xent.py:24-74) and their autograd.
 *   x[m]           n_tensors inputs [batch, dim] (x_dtype, common row stride)
 *   pair_row/col   host arrays: pair p is loss(x[pair_row[p]], x[pair_col[p]])  (argument order matters)
 *   state          caller-allocated, 256-byte aligned, tcl_ntxent_loss_state_bytes(); written by the
 *                  forward, read by the backward (normalised operands, 1/norms, LSEs)
 *   workspace      scratch, tcl_ntxent_loss_workspace_bytes(), 256-byte aligned
 *   loss           [n_pairs] fp32 out
 *   backward: grad_losses [n_pairs] device fp32 (upstream gradients), need_grad host flags [n_tensors],
 *             dx[m] device outputs [batch, dim] in x_dtype (ignored where need_grad[m] == 0)
 * ------------------------------------------------------------------------- */
size_t tcl_ntxent_loss_state_bytes(int n_tensors, int n_pairs, int64_t batch, int64_t dim);
size_t tcl_ntxent_loss_workspace_bytes(int n_tensors, int n_pairs, int64_t batch, int64_t dim);
int tcl_ntxent_loss_fwd(int n_tensors, const void* const* x_host_ptrs, int x_dtype, int64_t batch, int64_t dim,
                        int64_t x_row_stride, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                        int op_format, float inv_tau, float alpha, float eps, void* state, size_t state_bytes,
                        void* workspace, size_t workspace_bytes, float* loss, void* stream);
int tcl_ntxent_loss_bwd(int n_tensors, const void* const* x_host_ptrs, int x_dtype, int64_t batch, int64_t dim,
                        int64_t x_row_stride, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                        int op_format, float inv_tau, float alpha, float eps, const void* state,
                        const float* grad_losses, const uint8_t* need_grad_host, void* const* dx_host_ptrs,
                        void* workspace, size_t workspace_bytes, void* stream);
/* The same two calls with the SUM of the pair losses (loss_dict["<prefix>/total_loss"], tricolo_net.py:64) as an
 * output of the forward and its upstream gradient as an input of the backward, so that the training step
 * `total_loss.backward()` needs no framework kernels between the two library calls:
 *   loss          [n_pairs + 1] floats: the pair losses, then their fp32 sum in pair order
 *   grad_losses   [n_pairs] device floats or NULL, grad_total one device float or NULL (not both NULL):
 *                 the upstream gradient of pair p is grad_losses[p] + *grad_total
 * At batch sizes where every 32 x 64 logit tile gets its own SM (n_pairs * ceil(B/32) * ceil(B/64) <= SM count,
 * dim % 64 == 0, dim <= 512) all four entry points run ONE cooperative kernel each (csrc/ntxent_small.cu);
 * TRICOLO_B200_SMALL=0 keeps the multi-kernel form. */
int tcl_ntxent_loss_fwd_total(int n_tensors, const void* const* x_host_ptrs, int x_dtype, int64_t batch, int64_t dim,
                              int64_t x_row_stride, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                              int op_format, float inv_tau, float alpha, float eps, void* state, size_t state_bytes,
                              void* workspace, size_t workspace_bytes, float* loss, void* stream);
int tcl_ntxent_loss_bwd_total(int n_tensors, const void* const* x_host_ptrs, int x_dtype, int64_t batch, int64_t dim,
                              int64_t x_row_stride, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                              int op_format, float inv_tau, float alpha, float eps, const void* state,
                              const float* grad_losses, const float* grad_total, const uint8_t* need_grad_host,
                              void* const* dx_host_ptrs, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * NTXentLoss.forward(zis, zjs, norm=False): nt_xent.py:55-74 with the F.normalize of :56-57 skipped, and its autograd.
 * The logits x_i . x_j / tau are unbounded, so this path carries online (max, sum) statistics per row and column and
 * computes in fp32 on the FMA pipe (16-bit operands cannot give rtol 1e-3 on logits of magnitude 1e2..1e3); the
 * B x B logits never reach HBM here either.  TriCoLoNet itself only ever calls norm=True (tricolo_net.py:63).
 *   x, pair_row/col, loss [n_pairs + 1], grad_losses / grad_total, need_grad_host, dx: as tcl_ntxent_loss_*_total
 *   dim % 16 == 0; x_dtype f32 / f16 / bf16 (converted to fp32 on load; dx in x_dtype, [batch, dim] contiguous)
 *   state      tcl_ntxent_raw_state_bytes(): natural-log LSEs of the logit rows and columns, forward -> backward
 *   workspace  tcl_ntxent_raw_workspace_bytes(): forward scratch only
 * ------------------------------------------------------------------------- */
size_t tcl_ntxent_raw_state_bytes(int n_pairs, int64_t batch);
size_t tcl_ntxent_raw_workspace_bytes(int n_pairs, int64_t batch);
int tcl_ntxent_raw_fwd(int n_tensors, const void* const* x_host_ptrs, int x_dtype, int64_t batch, int64_t dim,
                       int64_t x_row_stride, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                       float inv_tau, float alpha, void* state, size_t state_bytes, void* workspace,
                       size_t workspace_bytes, float* loss, void* stream);
int tcl_ntxent_raw_bwd(int n_tensors, const void* const* x_host_ptrs, int x_dtype, int64_t batch, int64_t dim,
                       int64_t x_row_stride, int n_pairs, const int32_t* pair_row, const int32_t* pair_col,
                       float inv_tau, float alpha, const void* state, const float* grad_losses,
                       const float* grad_total, const uint8_t* need_grad_host, void* const* dx_host_ptrs,
                       void* stream);

/* ---------------------------------------------------------------------------
 * K2' — similarity GEMM for retrieval.           replaces eval_retrieval.py:74 (np.dot)
 * S[q, g] = Q[q,:] · G[g,:] (raw dot product, no normalisation), 16-bit operands,
 * fp32 accumulate, fp32 output with leading dimension ld_s (ld_s % 4 == 0).
 * dim % 8 == 0.
 * ------------------------------------------------------------------------- */
int tcl_sim_gemm(const void* q, const void* g, int64_t n_q, int64_t n_g, int64_t dim,
                 int op_format, float* s, int64_t ld_s, void* stream);

/* ---------------------------------------------------------------------------
 * K4 — per-query top-k + rank of the ground truth.  replaces eval_retrieval.py:75-82,184-186
 * Order: (similarity descending, gallery index ascending) — the stated tie-break.
 * For each query row of S [n_q, n_g] (ld_s):
 *   topk_val/topk_idx[q][0..k)  best k entries, idx = idx_base + column
 *   gt_sim[q]   = S[q, label[q]-idx_base]         (only written when gt_sim_in == NULL)
 *   n_before[q] = #{ j : S_qj > s_gt  or (S_qj == s_gt and idx_base+j < label[q]) }
 * so that rank = 1 + n_before (summed over gallery shards).  When the gallery is
 * sharded, pass the all-reduced ground-truth similarity in gt_sim_in.
 * k <= 16.  labels are global gallery indices.
 * ------------------------------------------------------------------------- */
int tcl_topk_rank(const float* s, int64_t ld_s, int64_t n_q, int64_t n_g, int k,
                  const int64_t* labels, int64_t idx_base, const float* gt_sim_in,
                  float* topk_val, int32_t* topk_idx, float* gt_sim_out, int32_t* n_before,
                  void* stream);

/* gt_sim[q] = S[q, label[q]-idx_base] if the label falls in [idx_base, idx_base+n_g) else 0
 * (shard owner contributes, the others add zero in the all-reduce). */
int tcl_gather_gt_sim(const float* s, int64_t ld_s, int64_t n_q, int64_t n_g,
                      const int64_t* labels, int64_t idx_base, float* gt_sim, void* stream);

/* Merge n_shards sorted candidate lists per query (layout [n_shards][n_q][k]) into one
 * top-k under the same order. */
int tcl_topk_merge(const float* cand_val, const int32_t* cand_idx, int n_shards, int64_t n_q,
                   int k, float* topk_val, int32_t* topk_idx, void* stream);

/* ---------------------------------------------------------------------------
 * K5 - metric reduction.                       replaces eval_retrieval.py:169-201 for the tensor-in entry
 * With exactly one relevant gallery item per query (the gallery is de-duplicated by model id, :49-56) every metric is
 * a function of the 1-based rank r_q of the ground truth:
 *   recall_rate[j] = recall[j] = mean(r <= j+1), precision[j] = mean(r <= j+1)/(j+1),
 *   ndcg[j] = mean([r <= j+1] / log2(r+1)), mrr = mean(1/r).
 * out (device, k+1 doubles): out[j] = #{q : r_q == j+1} for j < k, out[k] = sum_q 1/r_q (fp64, fixed order).
 * workspace: tcl_rank_metrics_workspace_bytes(), 256-byte aligned, zero-initialised once (the kernel leaves its
 * counter at zero).  k <= 16.
 * ------------------------------------------------------------------------- */
size_t tcl_rank_metrics_workspace_bytes(void);
int tcl_rank_metrics(const int32_t* rank, int64_t n_q, int k, double* out, void* workspace, size_t workspace_bytes,
                     void* stream);

/* ---------------------------------------------------------------------------
 * K2'+K4 fused — retrieval without materialising S (dim % 64 == 0, dim <= 512, k <= 16).
 * Same results as tcl_sim_gemm + tcl_topk_rank (same order, same rank definition), bit for bit:
 * the per-query ground-truth similarity gt_sim[q] must be the number the tensor core produces for
 * (q, label[q]); tcl_gt_sim_mma computes it with the identical MMA sequence on gathered gallery rows
 * (0 where the label is outside [idx_base, idx_base+n_g): all-reduce(sum) over gallery shards).
 *   tcl_gt_sim_mma      workspace >= n_q*dim*2 bytes (gathered rows), 16-byte aligned
 *   tcl_sim_topk_fused  workspace tcl_sim_topk_fused_workspace_bytes() (only used when the gallery sweep
 *                       is split over several CTAs per query block, i.e. few queries)
 * ------------------------------------------------------------------------- */
int tcl_gt_sim_mma(const void* q, const void* g, int64_t n_q, int64_t n_g, int64_t dim, int op_format,
                   const int64_t* labels, int64_t idx_base, float* gt_sim, void* workspace,
                   size_t workspace_bytes, void* stream);
size_t tcl_sim_topk_fused_workspace_bytes(int64_t n_q, int64_t n_g, int k);
int tcl_sim_topk_fused(const void* q, const void* g, int64_t n_q, int64_t n_g, int64_t dim, int op_format,
                       int k, const int64_t* labels, int64_t idx_base, const float* gt_sim,
                       float* topk_val, int32_t* topk_idx, int32_t* n_before, void* workspace,
                       size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Measurement hooks (bench.py).  tcl_launch_count: kernels launched by this library
 * since load.  With profiling enabled every kernel launch is bracketed by CUDA events
 * on its launch stream; tcl_profile_read synchronises on those events and returns the
 * summed device time and launch count of one kernel id since the last enable.
 * ------------------------------------------------------------------------- */
enum {
  TCL_K_L2NORM_FWD = 0,
  TCL_K_CAST16 = 1,
  TCL_K_TRANSPOSE16 = 2,
  TCL_K_NTXENT_FWD = 3,
  TCL_K_FWD_REDUCE = 4,
  TCL_K_FWD_FINALIZE = 5,
  TCL_K_NTXENT_BWD = 6,
  TCL_K_L2NORM_BWD = 7,
  TCL_K_SIM_GEMM = 8,
  TCL_K_TOPK_RANK = 9,
  TCL_K_GATHER_GT = 10,
  TCL_K_TOPK_MERGE = 11,
  TCL_K_SIM_TOPK_FUSED = 12,
  TCL_K_GATHER_SUM = 13,
  TCL_K_PEER_SUM = 14,
  TCL_K_NTXENT_G = 15, /* shared-G backward, kernel A (logit recompute -> G); its GEMM kernel is TCL_K_NTXENT_BWD */
  TCL_K_RANK_METRICS = 16,
  TCL_K_NTXENT_SMALL_FWD = 17, /* small-batch whole-loss forward, one cooperative launch (csrc/ntxent_small.cu) */
  TCL_K_NTXENT_SMALL_BWD = 18, /* small-batch whole-loss backward, one cooperative launch */
  TCL_K_NTXENT_RAW_FWD = 19,   /* norm=False forward tile kernel (csrc/ntxent_raw.cu); its finalise is TCL_K_FWD_FINALIZE */
  TCL_K_NTXENT_RAW_BWD = 20,   /* norm=False backward */
  TCL_K_COUNT = 21
};
int64_t tcl_launch_count(void);
int tcl_profile_enable(int on);
int tcl_profile_read(int kernel_id, double* total_ms, int64_t* launches);

/* ---------------------------------------------------------------------------
 * Bring-up / test hooks (not part of the product path).
 * tcl_debug_tmem_probe: writes lane*64+col into a 128x32 TMEM block and reads it back
 * through the 16x256b load shape; out[(warp*2+half)*32*16 + thread*16 + reg].
 * ------------------------------------------------------------------------- */
int tcl_debug_tmem_probe(uint32_t* out, void* stream);
/* tcl_debug_pc_trace: cycles the roles of the first cluster of the producer/consumer backward kernel (ntxent_bwd_pc.cu)
 * spent in each wait (32 counters; all zero unless the library was built with `make trace`). */
int tcl_debug_pc_trace(unsigned long long* out32, int reset);
/* tcl_debug_gb_trace: the same for the first ([0..31]) and the last ([32..63]) CTA of the shared-G gradient GEMM kernel
 * (ntxent_bwd_g.cu): 0/1 TMA warp waits (G slot, operand stage), 2 TMA warp total, 3/4/5 MMA warp waits (accumulator,
 * G tile, operand stage), 6 MMA warp total, 7 read-out wait, 8 read-out work, 9 read-out total, 10 tiles, 11 pieces. */
int tcl_debug_gb_trace(unsigned long long* out64, int reset);
/* tcl_debug_fwd_trace: the same for the first cluster of the CTA-pair forward kernel (ntxent_fwd.cu). */
int tcl_debug_fwd_trace(unsigned long long* out32, int reset);
/* tcl_debug_small_trace: %globaltimer stamps (ns) of CTA 0 at the phase boundaries of the small-batch kernels
 * (csrc/ntxent_small.cu): [0..5] forward, [16..22] backward; zeros unless built with `make trace`. */
int tcl_debug_small_trace(unsigned long long* out32);
/* tcl_debug_max_clusters: cudaOccupancyMaxActiveClusters for the producer/consumer kernel's footprint. */
int tcl_debug_max_clusters(int cluster_size, int* out);

#ifdef __cplusplus
}
#endif
#endif /* TRICOLO_B200_H_ */
