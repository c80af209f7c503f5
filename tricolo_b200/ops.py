"""Tensor-level wrappers of the C-ABI entry points (one function per export).

PyTorch is used here for device memory and streams only; every computation is
an sm_100a kernel behind include/tricolo_b200.h.  All functions enqueue on the
current CUDA stream and return without synchronising.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib as L

LIB = L.LIB
F16, BF16 = L.TCL_OP_F16, L.TCL_OP_BF16
EPS = 1e-12  # F.normalize default used at tricolo/loss/nt_xent.py:56-57


def _rows_2d(t: torch.Tensor) -> torch.Tensor:
    if t.dim() != 2:
        raise ValueError(f"expected a [rows, dim] matrix, got shape {tuple(t.shape)}")
    if t.stride(1) != 1 or (t.stride(0) * t.element_size()) % 16 != 0 or t.data_ptr() % 16 != 0:
        t = t.contiguous()
    return t


def _dense16(t: torch.Tensor) -> torch.Tensor:
    """The retrieval kernels build their TMA maps with row stride == dim: a strided 16-bit view (e.g. a column slice of
    an interleaved buffer) is copied to a dense matrix instead of being silently misread."""
    if t.dim() != 2:
        raise ValueError(f"expected a [rows, dim] matrix, got shape {tuple(t.shape)}")
    if t.stride(1) != 1 or t.stride(0) != t.shape[1] or t.data_ptr() % 16 != 0:
        t = t.contiguous()
    return t


def _z_stride(zs) -> int:
    """Common row stride (elements) of a set of 16-bit operand matrices (contiguous or strided views)."""
    st = zs[0].stride(0)
    for z in zs:
        if z.stride(1) != 1 or z.stride(0) != st:
            raise ValueError("16-bit operand matrices must share one row stride and have unit column stride")
    return st


def l2norm_fwd(xs: Sequence[torch.Tensor], op_format: int = F16, eps: float = EPS, out=None):
    """K1. Returns ([z 16-bit [rows, dim]], [inv_norm fp32 [rows]]). nt_xent.py:56-57.
    `out`: optional pre-allocated z views (e.g. column slices of one [rows, M*dim] buffer)."""
    dev = L.require_cuda(*xs)
    xs = [_rows_2d(x) for x in xs]
    rows, dim = xs[0].shape
    for x in xs:
        if x.shape != xs[0].shape or x.dtype != xs[0].dtype:
            raise ValueError("l2norm_fwd: all tensors must share shape and dtype")
    same_stride = all(x.stride(0) == xs[0].stride(0) for x in xs)
    if not same_stride:
        xs = [x.contiguous() for x in xs]
    zs = out if out is not None else [torch.empty((rows, dim), dtype=L.op_torch_dtype(op_format), device=dev) for _ in xs]
    inv_all = torch.empty((len(xs), rows), dtype=torch.float32, device=dev)
    invs = [inv_all[m] for m in range(len(xs))]
    with torch.cuda.device(dev):
        L.check(LIB.tcl_l2norm_fwd(len(xs), L.ptr_array(xs), L.dtype_code(xs[0]), rows, dim, xs[0].stride(0),
                                   L.ptr_array(zs), _z_stride(zs), op_format, L.ptr_array(invs), eps,
                                   L.stream_ptr(dev)))
    return zs, invs, xs


def l2norm_fwd_bcast(xs: Sequence[torch.Tensor], dsts: Sequence[Sequence], z_row_stride: int, op_format: int = F16,
                     eps: float = EPS):
    """K1 fused with the all-gather: dsts[r][m] = device address (int) of THIS rank's first row of modality m inside
    destination r - a peer-mapped buffer of rank r, or ONE NVLink-multicast mapping that the switch replicates to
    every rank.  z_row_stride in elements.  Returns ([inv_norm], xs).  The caller provides the cross-device barriers."""
    dev = L.require_cuda(*xs)
    xs = [_rows_2d(x) for x in xs]
    rows, dim = xs[0].shape
    for x in xs:
        if x.shape != xs[0].shape or x.dtype != xs[0].dtype:
            raise ValueError("l2norm_fwd_bcast: all tensors must share shape and dtype")
    if not all(x.stride(0) == xs[0].stride(0) for x in xs):
        xs = [x.contiguous() for x in xs]
    flat = [int(d) for per_dst in dsts for d in per_dst]
    if len(flat) != len(dsts) * len(xs):
        raise ValueError("l2norm_fwd_bcast: one destination address per (destination, modality)")
    inv_all = torch.empty((len(xs), rows), dtype=torch.float32, device=dev)
    invs = [inv_all[m] for m in range(len(xs))]
    dst_arr = (C.c_void_p * len(flat))(*flat)
    with torch.cuda.device(dev):
        L.check(LIB.tcl_l2norm_fwd_bcast(len(xs), L.ptr_array(xs), L.dtype_code(xs[0]), rows, dim, xs[0].stride(0),
                                         len(dsts), dst_arr, z_row_stride, op_format, L.ptr_array(invs),
                                         eps, L.stream_ptr(dev)))
    return invs, xs


def copy_rows(dst_addr: int, dst_pitch: int, src_addr: int, src_pitch: int, width: int, rows: int, stream: int) -> None:
    """Copy-engine (cudaMemcpy2DAsync) transfer of `rows` x `width` bytes between pitched device buffers; `stream` is a
    raw cudaStream_t."""
    L.check(LIB.tcl_copy_rows(int(dst_addr), dst_pitch, int(src_addr), src_pitch, width, rows, int(stream)))


def shard_sync_bytes() -> int:
    return int(LIB.tcl_shard_sync_bytes())


def shard_stats_bytes(n_pairs: int, b_loc: int, world: int) -> int:
    return int(LIB.tcl_shard_stats_bytes(n_pairs, b_loc, world))


def _addr_array(addrs):
    return (C.c_void_p * len(addrs))(*[int(a) for a in addrs])


def l2norm_fwd_push(xs: Sequence[torch.Tensor], dsts: Sequence[Sequence], z_row_stride: int, rank: int, world: int,
                    sync_addrs: Sequence[int], op_format: int = F16, eps: float = EPS, remote: bool = True):
    """K1 (+ all-gather + arrival flags when remote) (tcl_l2norm_fwd_push): dsts[r][m] = address of THIS rank's first
    row of modality m inside rank r's gathered buffer (r = 0..world-1, the own buffer included), sync_addrs[r] = rank
    r's sync pad.  remote=False: rows go to the own buffer only; the forward tile kernel's push warps do the gather.
    Returns (inv_norm [n, rows] fp32, xs).  No barrier: the kernel waits for each peer's "ready" flag itself."""
    dev = L.require_cuda(*xs)
    xs = [_rows_2d(x) for x in xs]
    rows, dim = xs[0].shape
    for x in xs:
        if x.shape != xs[0].shape or x.dtype != xs[0].dtype:
            raise ValueError("l2norm_fwd_push: all tensors must share shape and dtype")
    if not all(x.stride(0) == xs[0].stride(0) for x in xs):
        xs = [x.contiguous() for x in xs]
    flat = [int(d) for per_dst in dsts for d in per_dst]
    if len(dsts) != world or len(flat) != world * len(xs):
        raise ValueError("l2norm_fwd_push: one destination address per (rank, modality)")
    inv_all = torch.empty((len(xs), rows), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(LIB.tcl_l2norm_fwd_push(len(xs), L.ptr_array(xs), L.dtype_code(xs[0]), rows, dim, xs[0].stride(0), rank, world,
                                        _addr_array(flat), z_row_stride, op_format,
                                        L.ptr_array([inv_all[m] for m in range(len(xs))]), eps, _addr_array(sync_addrs),
                                        1 if remote else 0, L.stream_ptr(dev)))
    return inv_all, xs


def ntxent_fwd_sharded(zrows: Sequence[torch.Tensor], zcols: Sequence[torch.Tensor], rank: int, world: int, inv_tau: float,
                       stats_addrs: Sequence[int], sync_addrs: Optional[Sequence[int]], op_format: int = F16,
                       z_base_addrs: Optional[Sequence[int]] = None, push_offsets: Sequence[int] = ()) -> None:
    """K2 on the local row block, column tiles gated on the arrival flags, then this rank's sum-exp statistics stored
    into every rank's statistics buffer (+ flag).  push_offsets (element offsets of modalities inside a gathered row)
    + z_base_addrs (every rank's gathered buffer): the all-gather of those modalities is fused into the kernel (push
    warps).  Results: ntxent_finalize_sharded."""
    dev = L.require_cuda(*zrows, *zcols)
    p = len(zrows)
    b_loc, dim = zrows[0].shape
    b_glob = zcols[0].shape[0]
    diag2 = torch.empty((p, b_loc), dtype=torch.float32, device=dev)
    ws_bytes = LIB.tcl_ntxent_fwd_workspace_bytes(p, b_loc, b_glob)
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(LIB.tcl_ntxent_fwd_sharded(p, L.ptr_array(list(zrows)), L.ptr_array(list(zcols)), b_loc, b_glob, dim,
                                           _z_stride(list(zrows) + list(zcols)), rank, world, op_format, inv_tau,
                                           L.ptr(diag2), L.ptr(ws), ws_bytes, _addr_array(stats_addrs),
                                           None if sync_addrs is None else _addr_array(sync_addrs),
                                           None if not push_offsets else _addr_array(z_base_addrs), len(push_offsets),
                                           (C.c_int64 * max(len(push_offsets), 1))(*[int(o) for o in push_offsets]),
                                           L.stream_ptr(dev)))


def ntxent_finalize_sharded(n_pairs: int, b_loc: int, rank: int, world: int, inv_tau: float, alpha: float,
                            stats_addrs: Sequence[int], sync_own_addr: int, device):
    """lse2_row [P, B], lse2_col [P, B], loss [P] of the global batch.  sync_own_addr != 0: waits on the device for every
    rank's statistics flag (they were pushed into this rank's buffer); 0: pulls slot s from rank s's buffer
    (stats_addrs[s]) - the caller ran a barrier after ntxent_fwd_sharded."""
    b_glob = b_loc * world
    lse2_row = torch.empty((n_pairs, b_glob), dtype=torch.float32, device=device)
    lse2_col = torch.empty((n_pairs, b_glob), dtype=torch.float32, device=device)
    loss = torch.empty((n_pairs,), dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        L.check(LIB.tcl_ntxent_finalize_sharded(n_pairs, b_loc, b_glob, rank, world, inv_tau, alpha, _addr_array(stats_addrs),
                                                C.c_void_p(int(sync_own_addr)), L.ptr(lse2_row), L.ptr(lse2_col), L.ptr(loss),
                                                L.stream_ptr(device)))
    return lse2_row, lse2_col, loss


def peer_sum(srcs: Sequence[torch.Tensor], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = srcs[0] + srcs[1] + ... in that order (fp32, same shape): one-shot all-reduce over peer-mapped buffers."""
    dev = L.require_cuda(srcs[0])
    if out is None:
        out = torch.empty(srcs[0].shape, dtype=torch.float32, device=dev)
    n = out.numel()
    if any(t.numel() != n or t.dtype != torch.float32 or not t.is_contiguous() for t in list(srcs) + [out]):
        raise ValueError("peer_sum: contiguous fp32 tensors of one size")
    with torch.cuda.device(dev):
        L.check(LIB.tcl_peer_sum_f32(len(srcs), L.ptr_array(list(srcs)), n, L.ptr(out), L.stream_ptr(dev)))
    return out


def cast_16bit(x: torch.Tensor, op_format: int = BF16) -> torch.Tensor:
    dev = L.require_cuda(x)
    x = _rows_2d(x)
    y = torch.empty(x.shape, dtype=L.op_torch_dtype(op_format), device=dev)
    with torch.cuda.device(dev):
        L.check(LIB.tcl_cast_16bit(L.ptr(x), L.dtype_code(x), x.shape[0], x.shape[1], x.stride(0), L.ptr(y),
                                   op_format, L.stream_ptr(dev)))
    return y


def gather_sum_cast16(srcs: Sequence[torch.Tensor], index: torch.Tensor, op_format: int = BF16) -> torch.Tensor:
    """out16[g] = round16(srcs[0][index[g]] (+ srcs[1][index[g]])): gallery build of the device-resident evaluation
    hand-off (tricolo_net.py:125-158 + eval_retrieval.py:49-56).  index: int64 device tensor."""
    dev = L.require_cuda(*srcs, index)
    srcs = [_rows_2d(x) for x in srcs]
    if not 1 <= len(srcs) <= 2:
        raise ValueError("gather_sum_cast16: one or two sources")
    rows, dim = srcs[0].shape
    for x in srcs:
        if x.shape != srcs[0].shape or x.dtype != srcs[0].dtype:
            raise ValueError("gather_sum_cast16: sources must share shape and dtype")
    if any(x.stride(0) != srcs[0].stride(0) for x in srcs):
        srcs = [x.contiguous() for x in srcs]
    index = index.to(torch.int64).contiguous()
    if index.numel() and (int(index.min()) < 0 or int(index.max()) >= rows):
        raise IndexError("gather_sum_cast16: index out of range")
    out = torch.empty((index.numel(), dim), dtype=L.op_torch_dtype(op_format), device=dev)
    with torch.cuda.device(dev):
        L.check(LIB.tcl_gather_sum_cast16(len(srcs), L.ptr_array(srcs), L.dtype_code(srcs[0]), rows, dim,
                                          srcs[0].stride(0), L.ptr(index), index.numel(), L.ptr(out), op_format,
                                          L.stream_ptr(dev)))
    return out


def bwd_needs_transpose(dim: int) -> bool:
    """Whether tcl_ntxent_bwd wants the transposed operand copies for this dim (the default kernel for
    dim > 256 reads the row-major operands directly)."""
    return bool(LIB.tcl_ntxent_bwd_needs_transpose(int(dim)))


def transpose_for_bwd(zs: Sequence[torch.Tensor]) -> Tuple[List[Optional[torch.Tensor]], int]:
    """The transposed copies tcl_ntxent_bwd needs for these operands: ([None, ...], 0) when it needs none."""
    if not bwd_needs_transpose(zs[0].shape[1]):
        return [None] * len(zs), 0
    return transpose_16bit(zs)


def transpose_16bit(zs: Sequence[torch.Tensor]) -> Tuple[List[torch.Tensor], int]:
    """[rows, dim] -> [dim, ld_t] (ld_t = rows rounded up to 8)."""
    dev = L.require_cuda(*zs)
    rows, dim = zs[0].shape
    ld_t = (rows + 7) // 8 * 8
    zts = [torch.empty((dim, ld_t), dtype=z.dtype, device=dev) for z in zs]
    with torch.cuda.device(dev):
        L.check(LIB.tcl_transpose_16bit(len(zs), L.ptr_array(zs), rows, dim, _z_stride(zs), L.ptr_array(zts), ld_t,
                                        L.stream_ptr(dev)))
    return zts, ld_t


def ntxent_fwd(zrows: Sequence[torch.Tensor], zcols: Sequence[torch.Tensor], row_offset: int, inv_tau: float,
               op_format: int = F16):
    """K2. Returns row_sumexp [P, n_rows], col_sumexp [P, n_cols] (partial over the given rows), diag2 [P, n_rows]."""
    dev = L.require_cuda(*zrows, *zcols)
    p = len(zrows)
    n_rows, dim = zrows[0].shape
    n_cols = zcols[0].shape[0]
    row_sum = torch.empty((p, n_rows), dtype=torch.float32, device=dev)
    col_sum = torch.empty((p, n_cols), dtype=torch.float32, device=dev)
    diag2 = torch.empty((p, n_rows), dtype=torch.float32, device=dev)
    ws_bytes = LIB.tcl_ntxent_fwd_workspace_bytes(p, n_rows, n_cols)
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(LIB.tcl_ntxent_fwd(p, L.ptr_array(zrows), L.ptr_array(zcols), n_rows, n_cols, dim,
                                   _z_stride(list(zrows) + list(zcols)), row_offset,
                                   op_format, inv_tau, L.ptr(row_sum), L.ptr(col_sum), L.ptr(diag2), L.ptr(ws),
                                   ws_bytes, L.stream_ptr(dev)))
    return row_sum, col_sum, diag2


def ntxent_finalize(row_sum, col_sum, diag2, row_offset: int, inv_tau: float, alpha: float, want_loss: bool = True):
    dev = L.require_cuda(row_sum, col_sum, diag2)
    p, n_rows = row_sum.shape
    n_cols = col_sum.shape[1]
    lse2_row = torch.empty_like(row_sum)
    lse2_col = torch.empty_like(col_sum)
    parts = torch.empty((p, 2), dtype=torch.float32, device=dev)
    loss = torch.empty((p,), dtype=torch.float32, device=dev) if want_loss else None
    with torch.cuda.device(dev):
        L.check(LIB.tcl_ntxent_finalize(p, n_rows, n_cols, row_offset, inv_tau, alpha, L.ptr(row_sum), L.ptr(col_sum),
                                        L.ptr(diag2), L.ptr(lse2_row), L.ptr(lse2_col), L.ptr(parts), L.ptr(loss),
                                        L.stream_ptr(dev)))
    return lse2_row, lse2_col, parts, loss


class BwdSegmentSpec:
    __slots__ = ("z_other", "z_other_t", "lse2_self", "lse2_other", "grad_scale", "w_self", "w_other")

    def __init__(self, z_other, z_other_t, lse2_self, lse2_other, grad_scale, w_self, w_other):
        self.z_other, self.z_other_t = z_other, z_other_t
        self.lse2_self, self.lse2_other = lse2_self, lse2_other
        self.grad_scale, self.w_self, self.w_other = grad_scale, w_self, w_other


class BwdJobSpec:
    __slots__ = ("z_self", "x_self", "inv_norm", "segments")

    def __init__(self, z_self, x_self, inv_norm, segments):
        self.z_self, self.x_self, self.inv_norm, self.segments = z_self, x_self, inv_norm, segments


def ntxent_bwd(jobs: Sequence[BwdJobSpec], n_other: int, self_offset: int, ld_t: int, inv_tau: float,
               op_format: int = F16, eps: float = EPS) -> List[torch.Tensor]:
    """K3 + normalise backward. Returns dx (dtype/shape of x_self) per job."""
    x0 = jobs[0].x_self
    dev = L.require_cuda(x0)
    n_self, dim = x0.shape
    arr = (L.BwdJob * len(jobs))()
    dxs, keep = [], []
    for j, job in enumerate(jobs):
        if job.x_self.shape != x0.shape or job.x_self.dtype != x0.dtype or job.x_self.stride(0) != x0.stride(0):
            raise ValueError("ntxent_bwd: all jobs must share shape, dtype and row stride")
        dx = torch.empty((n_self, dim), dtype=x0.dtype, device=dev)
        dxs.append(dx)
        arr[j].z_self = job.z_self.data_ptr()
        arr[j].x_self = job.x_self.data_ptr()
        arr[j].inv_norm = job.inv_norm.data_ptr()
        arr[j].dx = dx.data_ptr()
        arr[j].n_segments = len(job.segments)
        for s, sg in enumerate(job.segments):
            a = arr[j].seg[s]
            a.z_other = sg.z_other.data_ptr()
            a.z_other_t = 0 if sg.z_other_t is None else sg.z_other_t.data_ptr()
            a.lse2_self = sg.lse2_self.data_ptr()
            a.lse2_other = sg.lse2_other.data_ptr()
            a.grad_scale = 0 if sg.grad_scale is None else sg.grad_scale.data_ptr()
            a.w_self, a.w_other = sg.w_self, sg.w_other
            keep.append(sg)
    ws_bytes = LIB.tcl_ntxent_bwd_workspace_bytes(len(jobs), n_self, dim)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        z_stride = _z_stride([j.z_self for j in jobs] + [s.z_other for j in jobs for s in j.segments])
        L.check(LIB.tcl_ntxent_bwd(len(jobs), arr, n_self, n_other, dim, z_stride, self_offset, ld_t, L.dtype_code(x0),
                                   x0.stride(0), op_format, inv_tau, eps, L.ptr(ws), ws_bytes, L.stream_ptr(dev)))
    return dxs


class ShardedBwdPlan:
    """Sizes and host-side argument arrays of the sharded shared-G backward (tcl_ntxent_bwd_sharded_*): a pure function
    of (pairs, need_grad, b_loc, world, dim), identical on every rank."""

    def __init__(self, n_tensors: int, pairs, need_grad, b_loc: int, world: int, dim: int):
        self.n_tensors, self.pairs, self.b_loc, self.world, self.dim = n_tensors, list(pairs), b_loc, world, dim
        self.b_glob = b_loc * world
        self.pair_row = (C.c_int32 * len(self.pairs))(*[a for a, _ in self.pairs])
        self.pair_col = (C.c_int32 * len(self.pairs))(*[b for _, b in self.pairs])
        self.need = (C.c_uint8 * n_tensors)(*[1 if g else 0 for g in need_grad])
        self.need_grad = [bool(g) for g in need_grad]
        args = (n_tensors, len(self.pairs), self.pair_row, self.pair_col, self.need, b_loc, self.b_glob, dim, world)
        self.workspace_bytes = int(LIB.tcl_ntxent_bwd_sharded_workspace_bytes(*args))
        self.recv_bytes = int(LIB.tcl_ntxent_bwd_sharded_recv_bytes(*args))
        if self.workspace_bytes == 0 or self.recv_bytes == 0:
            raise ValueError("sharded backward: unsupported configuration: " + LIB.tcl_last_error_string().decode())

    @staticmethod
    def supported(b_loc: int, dim: int, world: int) -> bool:
        return 256 < dim <= 512 and dim % 64 == 0 and b_loc >= 128 and b_loc % 128 == 0 and 1 <= world <= 8


def ntxent_bwd_sharded_gemm(plan: ShardedBwdPlan, z_all: Sequence[torch.Tensor], rank: int, inv_tau: float, alpha: float,
                            lse2_row: torch.Tensor, lse2_col: torch.Tensor, grad_losses: torch.Tensor,
                            workspace: torch.Tensor, recv_addrs: Sequence[int], op_format: int = F16,
                            sync_addrs: Optional[Sequence[int]] = None) -> None:
    """Kernel A (row block of G per pair) + kernel B (row-side gradients into local partials, column-side partials
    stored into every owner's receive buffer, recv_addrs[r] = device address of rank r's buffer as mapped here)."""
    dev = L.require_cuda(*z_all, lse2_row, lse2_col, grad_losses, workspace)
    if lse2_row.shape != (len(plan.pairs), plan.b_glob) or lse2_col.shape != lse2_row.shape:
        raise ValueError("sharded backward: LSEs must be [n_pairs, b_glob]")
    if not (lse2_row.is_contiguous() and lse2_col.is_contiguous() and grad_losses.is_contiguous()):
        raise ValueError("sharded backward: contiguous statistics expected")
    recv = (C.c_void_p * plan.world)(*[int(a) for a in recv_addrs])
    with torch.cuda.device(dev):
        L.check(LIB.tcl_ntxent_bwd_sharded_gemm(plan.n_tensors, L.ptr_array(list(z_all)), plan.b_loc, plan.b_glob, plan.dim,
                                                _z_stride(list(z_all)), rank, plan.world, len(plan.pairs), plan.pair_row,
                                                plan.pair_col, op_format, inv_tau, alpha, L.ptr(lse2_row), L.ptr(lse2_col),
                                                L.ptr(grad_losses), plan.need, L.ptr(workspace), workspace.numel(), recv,
                                                plan.recv_bytes, None if sync_addrs is None else _addr_array(sync_addrs),
                                                L.stream_ptr(dev)))


def ntxent_bwd_sharded_finish(plan: ShardedBwdPlan, xs: Sequence[torch.Tensor], invs: torch.Tensor, rank: int,
                              workspace: torch.Tensor, recv_own_addr: int, eps: float = EPS,
                              sync_own_addr: int = 0) -> List[Optional[torch.Tensor]]:
    """Sum of the local and received gradient partials + normalise backward.  Returns dx per tensor (None where no
    gradient was requested).  Every rank's gemm call must have completed: either the caller ran a cross-rank barrier,
    or both calls were given the sync pads (sync_own_addr: the kernel waits for the ranks' flags itself)."""
    dev = L.require_cuda(*xs, invs, workspace)
    x0 = xs[0]
    if any(x.shape != x0.shape or x.dtype != x0.dtype or x.stride(0) != x0.stride(0) or x.stride(1) != 1 for x in xs):
        raise ValueError("sharded backward: inputs must share shape, dtype and row stride")
    if invs.shape != (plan.n_tensors, plan.b_loc) or not invs.is_contiguous():
        raise ValueError("sharded backward: inv_norm must be a contiguous [n_tensors, b_loc] tensor")
    dxs = [torch.empty((plan.b_loc, plan.dim), dtype=x0.dtype, device=dev) if g else None for g in plan.need_grad]
    dx_arr = (C.c_void_p * plan.n_tensors)(*[0 if d is None else d.data_ptr() for d in dxs])
    with torch.cuda.device(dev):
        L.check(LIB.tcl_ntxent_bwd_sharded_finish(plan.n_tensors, L.ptr_array(list(xs)), L.dtype_code(x0), plan.b_loc,
                                                  plan.b_glob, plan.dim, x0.stride(0), rank, plan.world, len(plan.pairs),
                                                  plan.pair_row, plan.pair_col, L.ptr(invs), plan.need, eps, L.ptr(workspace),
                                                  C.c_void_p(int(recv_own_addr)), C.c_void_p(int(sync_own_addr)), dx_arr,
                                                  L.stream_ptr(dev)))
    return dxs


def sim_gemm(q16: torch.Tensor, g16: torch.Tensor, out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, int]:
    """K2'. S = Q G^T in fp32 with leading dimension ld (multiple of 32 floats = one 128-byte line).
    Returns (S [n_q, ld], n_g)."""
    dev = L.require_cuda(q16, g16)
    if q16.dtype != g16.dtype or q16.dtype not in (torch.float16, torch.bfloat16):
        raise TypeError("sim_gemm: operands must both be float16 or both bfloat16")
    q16, g16 = _dense16(q16), _dense16(g16)
    n_q, dim = q16.shape
    n_g = g16.shape[0]
    ld = (n_g + 31) // 32 * 32
    if out is None:
        out = torch.empty((n_q, ld), dtype=torch.float32, device=dev)
    op = F16 if q16.dtype == torch.float16 else BF16
    with torch.cuda.device(dev):
        L.check(LIB.tcl_sim_gemm(L.ptr(q16), L.ptr(g16), n_q, n_g, dim, op, L.ptr(out), out.stride(0), L.stream_ptr(dev)))
    return out, n_g


def topk_rank(s: torch.Tensor, n_g: int, k: int, labels: torch.Tensor, idx_base: int = 0,
              gt_sim_in: Optional[torch.Tensor] = None):
    """K4. Returns (topk_val [Q,k] f32, topk_idx [Q,k] i32, gt_sim [Q] f32, n_before [Q] i32)."""
    dev = L.require_cuda(s, labels)
    n_q = s.shape[0]
    labels = labels.to(torch.int64).contiguous()
    if s.stride(1) != 1:
        raise ValueError("topk_rank: similarity rows must have unit column stride")
    val = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    gt = torch.empty((n_q,), dtype=torch.float32, device=dev) if gt_sim_in is None else gt_sim_in
    nb = torch.empty((n_q,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(LIB.tcl_topk_rank(L.ptr(s), s.stride(0), n_q, n_g, k, L.ptr(labels), idx_base, L.ptr(gt_sim_in),
                                  L.ptr(val), L.ptr(idx), L.ptr(gt), L.ptr(nb), L.stream_ptr(dev)))
    return val, idx, gt, nb


def gather_gt_sim(s: torch.Tensor, n_g: int, labels: torch.Tensor, idx_base: int) -> torch.Tensor:
    dev = L.require_cuda(s, labels)
    out = torch.empty((s.shape[0],), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        L.check(LIB.tcl_gather_gt_sim(L.ptr(s), s.stride(0), s.shape[0], n_g, L.ptr(labels), idx_base, L.ptr(out),
                                      L.stream_ptr(dev)))
    return out


def topk_merge(cand_val: torch.Tensor, cand_idx: torch.Tensor):
    """cand_* [n_shards, Q, k] -> merged (val [Q,k], idx [Q,k])."""
    dev = L.require_cuda(cand_val, cand_idx)
    n_shards, n_q, k = cand_val.shape
    cand_val, cand_idx = cand_val.contiguous(), cand_idx.contiguous()
    val = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        L.check(LIB.tcl_topk_merge(L.ptr(cand_val), L.ptr(cand_idx), n_shards, n_q, k, L.ptr(val), L.ptr(idx),
                                   L.stream_ptr(dev)))
    return val, idx


_RM_WS = {}


def rank_metrics(rank: torch.Tensor, k: int) -> torch.Tensor:
    """K5: device tensor of k+1 doubles - #{rank == j+1} for j < k, then sum 1/rank (fp64, fixed order)."""
    dev = L.require_cuda(rank)
    if rank.dtype != torch.int32 or not rank.is_contiguous():
        raise TypeError("rank_metrics: contiguous int32 ranks expected")
    ws = _RM_WS.get(dev)
    if ws is None:
        ws = _RM_WS[dev] = torch.zeros((int(LIB.tcl_rank_metrics_workspace_bytes()),), dtype=torch.uint8, device=dev)
    out = torch.empty((k + 1,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        L.check(LIB.tcl_rank_metrics(L.ptr(rank), rank.numel(), k, L.ptr(out), L.ptr(ws), ws.numel(), L.stream_ptr(dev)))
    return out


def gt_sim_mma(q16: torch.Tensor, g16: torch.Tensor, labels: torch.Tensor, idx_base: int = 0) -> torch.Tensor:
    """Ground-truth similarity q . g[label - idx_base] produced by the same MMA sequence as the GEMM
    (0 where the label is outside this gallery shard)."""
    dev = L.require_cuda(q16, g16, labels)
    q16, g16, labels = _dense16(q16), _dense16(g16), labels.to(torch.int64).contiguous()
    n_q, dim = q16.shape
    op = F16 if q16.dtype == torch.float16 else BF16
    out = torch.empty((n_q,), dtype=torch.float32, device=dev)
    ws = torch.empty((max(n_q * dim * 2, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(LIB.tcl_gt_sim_mma(L.ptr(q16), L.ptr(g16), n_q, g16.shape[0], dim, op, L.ptr(labels), idx_base,
                                   L.ptr(out), L.ptr(ws), ws.numel(), L.stream_ptr(dev)))
    return out


def sim_topk_fused(q16: torch.Tensor, g16: torch.Tensor, k: int, labels: torch.Tensor, gt_sim: torch.Tensor,
                   idx_base: int = 0):
    """K2'+K4 fused. Returns (topk_val [Q,k] f32, topk_idx [Q,k] i32, n_before [Q] i32)."""
    dev = L.require_cuda(q16, g16, labels, gt_sim)
    if q16.dtype != g16.dtype or q16.dtype not in (torch.float16, torch.bfloat16):
        raise TypeError("sim_topk_fused: operands must both be float16 or both bfloat16")
    q16, g16, labels, gt_sim = _dense16(q16), _dense16(g16), labels.to(torch.int64).contiguous(), gt_sim.contiguous()
    n_q, dim = q16.shape
    n_g = g16.shape[0]
    op = F16 if q16.dtype == torch.float16 else BF16
    val = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    nb = torch.empty((n_q,), dtype=torch.int32, device=dev)
    ws_bytes = LIB.tcl_sim_topk_fused_workspace_bytes(n_q, n_g, k)
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(LIB.tcl_sim_topk_fused(L.ptr(q16), L.ptr(g16), n_q, n_g, dim, op, k, L.ptr(labels), idx_base,
                                       L.ptr(gt_sim), L.ptr(val), L.ptr(idx), L.ptr(nb), L.ptr(ws), ws_bytes,
                                       L.stream_ptr(dev)))
    return val, idx, nb


def debug_tmem_probe(device="cuda") -> torch.Tensor:
    out = torch.zeros((4, 2, 32, 16), dtype=torch.int32, device=device)
    with torch.cuda.device(out.device):
        L.check(LIB.tcl_debug_tmem_probe(L.ptr(out), L.stream_ptr(out.device)))
    return out
