"""NT-Xent / InfoNCE loss on sm_100a behind the reference's call signatures.

    NTXentLoss(temperature, alpha_weight)(zis, zjs, norm=True) -> 0-dim tensor
        mirrors tricolo/loss/nt_xent.py:6-74 (constructor kwargs from config/config.yaml:97-100)
    calculate_losses(output_dict, loss_prefix, loss_fn) -> dict
        mirrors TriCoLoNet._calculate_losses (tricolo/model/tricolo_net.py:56-65): same keys,
        same pair order, but ONE fused forward and ONE fused backward for all pairs.

Host code is a torch.autograd.Function over the C ABI; there is no PyTorch
implementation of the math in this package.
"""
from __future__ import annotations

import functools
import os
from itertools import combinations
from typing import Dict, List, Sequence, Tuple

import ctypes as C

import torch

from .. import _lib as L
from .. import ops

LIB = L.LIB

# fp16 operands by default: same tcgen05 kind::f16 throughput as bf16, 3 more
# mantissa bits; normalised embeddings lie in [-1, 1] so range is a non-issue.
# bf16 operands miss the 1e-3 gradient tolerance ("Operand format").
DEFAULT_OP_FORMAT = ops.F16


def _pairs(n: int) -> List[Tuple[int, int]]:
    return list(combinations(range(n), 2))


@functools.lru_cache(maxsize=64)
def _plan(n: int, pairs: Tuple[Tuple[int, int], ...], b: int, d: int, env):
    """Per-shape constants of the two library calls (the pair index arrays and the buffer sizes): a training loop
    calls with one shape, and at small batch the host side of a step is longer than its device side.  `env`: the
    switches the library reads per call and that change the workspace size."""
    p = len(pairs)
    pr = (C.c_int32 * p)(*[a for a, _ in pairs])
    pc = (C.c_int32 * p)(*[c for _, c in pairs])
    return pr, pc, LIB.tcl_ntxent_loss_state_bytes(n, p, b, d), LIB.tcl_ntxent_loss_workspace_bytes(n, p, b, d)


class _OnDevice:
    """`with torch.cuda.device(dev)` only when `dev` is not already current (the context manager costs ~10 us)."""

    def __init__(self, dev):
        self.ctx = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def _prep(feats):
    """Row-major, 16-byte aligned rows with one common row stride (what the C ABI takes)."""
    out = []
    for f in feats:
        if f.stride(1) != 1 or (f.stride(0) * f.element_size()) % 16 != 0 or f.data_ptr() % 16 != 0:
            f = f.contiguous()
        out.append(f)
    if any(f.stride(0) != out[0].stride(0) for f in out):
        out = [f.contiguous() for f in out]
    return out


class _FusedNTXent(torch.autograd.Function):
    """(losses[p] for every pair (a, b), a < b, of the given feature matrices; their sum).

    Two library calls per step: tcl_ntxent_loss_fwd_total (K1 -> K2 -> reduce -> finalise, or ONE cooperative kernel
    at small batch) and tcl_ntxent_loss_bwd_total (G -> gradient GEMMs -> normalise backward, or ONE cooperative
    kernel); see include/tricolo_b200.h.  The sum is an output of the forward and its upstream gradient an input of
    the backward, so `total_loss.backward()` (tricolo_net.py:64, the training step) runs no framework kernels between
    the two calls except autograd's own ones_like for the root gradient."""

    @staticmethod
    def forward(ctx, temperature: float, alpha: float, op_format: int, pairs, *feats: torch.Tensor):
        dev = L.require_cuda(*feats)
        xs = _prep([f.detach() for f in feats])
        n, p = len(xs), len(pairs)
        b, d = xs[0].shape
        inv_tau = 1.0 / float(temperature)
        env = os.environ.get
        pr, pc, state_bytes, ws_bytes = _plan(n, tuple(pairs), b, d, (env("TRICOLO_B200_BWD"), env("TRICOLO_B200_SMALL")))
        state = torch.empty((state_bytes,), dtype=torch.uint8, device=dev)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        loss = torch.empty((p + 1,), dtype=torch.float32, device=dev)
        with _OnDevice(dev):
            L.check(LIB.tcl_ntxent_loss_fwd_total(n, L.ptr_array(xs), L.dtype_code(xs[0]), b, d, xs[0].stride(0), p, pr,
                                                  pc, op_format, inv_tau, alpha, ops.EPS, state.data_ptr(), state_bytes,
                                                  ws.data_ptr(), ws_bytes, loss.data_ptr(), L.stream_ptr(dev)))
        ctx.cfg = (inv_tau, float(alpha), op_format, pr, pc, ws_bytes)
        ctx.save_for_backward(state, *xs)
        ctx.set_materialize_grads(False)  # an unused output arrives as None, not as a zeros kernel
        return loss[:p], loss[p]

    @staticmethod
    def backward(ctx, grad_losses, grad_total):
        inv_tau, alpha, op_format, pr, pc, ws_bytes = ctx.cfg
        state, *xs = ctx.saved_tensors
        n, p = len(xs), len(pr)
        b, d = xs[0].shape
        dev = xs[0].device
        if grad_losses is None and grad_total is None:
            return (None,) * (4 + n)
        if grad_losses is not None and (grad_losses.dtype != torch.float32 or not grad_losses.is_contiguous()):
            grad_losses = grad_losses.to(torch.float32).contiguous()
        if grad_total is not None and grad_total.dtype != torch.float32:
            grad_total = grad_total.to(torch.float32)
        need = (C.c_uint8 * n)(*[1 if ctx.needs_input_grad[4 + m] else 0 for m in range(n)])
        dx_all = torch.empty((n, b, d), dtype=xs[0].dtype, device=dev)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        dxs = (C.c_void_p * n)(*[dx_all.data_ptr() + m * b * d * dx_all.element_size() for m in range(n)])
        with _OnDevice(dev):
            L.check(LIB.tcl_ntxent_loss_bwd_total(n, L.ptr_array(xs), L.dtype_code(xs[0]), b, d, xs[0].stride(0), p, pr,
                                                  pc, op_format, inv_tau, alpha, ops.EPS, state.data_ptr(),
                                                  None if grad_losses is None else grad_losses.data_ptr(),
                                                  None if grad_total is None else grad_total.data_ptr(),
                                                  need, dxs, ws.data_ptr(), ws_bytes, L.stream_ptr(dev)))
        grads = [dx_all[m] if need[m] else None for m in range(n)]
        return (None, None, None, None, *grads)


class _RawNTXent(torch.autograd.Function):
    """The same outputs as _FusedNTXent for norm=False (nt_xent.py:55: the F.normalize of :56-57 is skipped):
    tcl_ntxent_raw_fwd / _bwd (csrc/ntxent_raw.cu) - fp32 logit tiles with online (max, sum) statistics."""

    @staticmethod
    def forward(ctx, temperature: float, alpha: float, pairs, *feats: torch.Tensor):
        dev = L.require_cuda(*feats)
        xs = _prep([f.detach() for f in feats])
        n, p = len(xs), len(pairs)
        b, d = xs[0].shape
        inv_tau = 1.0 / float(temperature)
        pr = (C.c_int32 * p)(*[a for a, _ in pairs])
        pc = (C.c_int32 * p)(*[c for _, c in pairs])
        state_bytes, ws_bytes = LIB.tcl_ntxent_raw_state_bytes(p, b), LIB.tcl_ntxent_raw_workspace_bytes(p, b)
        state = torch.empty((state_bytes,), dtype=torch.uint8, device=dev)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        loss = torch.empty((p + 1,), dtype=torch.float32, device=dev)
        with _OnDevice(dev):
            L.check(LIB.tcl_ntxent_raw_fwd(n, L.ptr_array(xs), L.dtype_code(xs[0]), b, d, xs[0].stride(0), p, pr, pc,
                                           inv_tau, alpha, state.data_ptr(), state_bytes, ws.data_ptr(), ws_bytes,
                                           loss.data_ptr(), L.stream_ptr(dev)))
        ctx.cfg = (inv_tau, float(alpha), pr, pc)
        ctx.save_for_backward(state, *xs)
        ctx.set_materialize_grads(False)
        return loss[:p], loss[p]

    @staticmethod
    def backward(ctx, grad_losses, grad_total):
        inv_tau, alpha, pr, pc = ctx.cfg
        state, *xs = ctx.saved_tensors
        n, p = len(xs), len(pr)
        b, d = xs[0].shape
        dev = xs[0].device
        if grad_losses is None and grad_total is None:
            return (None,) * (3 + n)
        if grad_losses is not None and (grad_losses.dtype != torch.float32 or not grad_losses.is_contiguous()):
            grad_losses = grad_losses.to(torch.float32).contiguous()
        if grad_total is not None and grad_total.dtype != torch.float32:
            grad_total = grad_total.to(torch.float32)
        need = (C.c_uint8 * n)(*[1 if ctx.needs_input_grad[3 + m] else 0 for m in range(n)])
        dx_all = torch.empty((n, b, d), dtype=xs[0].dtype, device=dev)
        dxs = (C.c_void_p * n)(*[dx_all.data_ptr() + m * b * d * dx_all.element_size() for m in range(n)])
        with _OnDevice(dev):
            L.check(LIB.tcl_ntxent_raw_bwd(n, L.ptr_array(xs), L.dtype_code(xs[0]), b, d, xs[0].stride(0), p, pr, pc,
                                           inv_tau, alpha, state.data_ptr(),
                                           None if grad_losses is None else grad_losses.data_ptr(),
                                           None if grad_total is None else grad_total.data_ptr(),
                                           need, dxs, L.stream_ptr(dev)))
        return (None, None, None, *[dx_all[m] if need[m] else None for m in range(n)])


def trimodal_ntxent(feats: Sequence[torch.Tensor], temperature: float, alpha: float,
                    op_format: int = DEFAULT_OP_FORMAT, norm: bool = True) -> torch.Tensor:
    """Per-pair losses [n_pairs] (fp32) for all unordered pairs of `feats`, in combinations() order."""
    return trimodal_ntxent_total(feats, temperature, alpha, op_format, norm)[0]


def trimodal_ntxent_total(feats: Sequence[torch.Tensor], temperature: float, alpha: float,
                          op_format: int = DEFAULT_OP_FORMAT, norm: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """(per-pair losses [n_pairs], their sum as a 0-dim tensor), both fp32 and both differentiable; backpropagating
    through the sum alone costs no framework kernels (see _FusedNTXent).  norm=False (nt_xent.py:55) takes the fp32
    path of csrc/ntxent_raw.cu (`op_format` does not apply there)."""
    feats = list(feats)
    if len(feats) < 2 or len(feats) > 3:
        raise ValueError("expected 2 or 3 feature matrices")
    b, d = feats[0].shape
    for f in feats:
        if f.shape != (b, d):
            raise ValueError(f"feature matrices must share one [batch, dim] shape, got {tuple(f.shape)} vs {(b, d)}")
    dt = feats[0].dtype
    feats = [f if f.dtype == dt else f.to(dt) for f in feats]
    if not norm:
        return _RawNTXent.apply(float(temperature), float(alpha), _pairs(len(feats)), *feats)
    return _FusedNTXent.apply(float(temperature), float(alpha), op_format, _pairs(len(feats)), *feats)


class NTXentLoss(torch.nn.Module):
    """Drop-in for tricolo.loss.nt_xent.NTXentLoss (Hydra `_target_`, config/config.yaml:97).

    No parameters and no buffers, so state_dict() is unchanged (test.py:29 strict load).
    """

    def __init__(self, temperature, alpha_weight, op_format: int = DEFAULT_OP_FORMAT):
        super().__init__()
        self.temperature = temperature
        self.alpha_weight = alpha_weight
        self.op_format = op_format

    def forward(self, zis, zjs, norm=True):
        # one pair: the sum output IS the pair's loss (0 + l in fp32), and needs no select / select-backward kernels.
        # norm=False (nt_xent.py:55; never used by TriCoLoNet, tricolo_net.py:63) runs the fp32 kernels of
        # csrc/ntxent_raw.cu: unbounded logits need online (max, sum) statistics and fp32 products.
        return trimodal_ntxent_total([zis, zjs], self.temperature, self.alpha_weight, self.op_format, norm=bool(norm))[1]

    def fused(self, feats: Sequence[torch.Tensor], norm: bool = True) -> torch.Tensor:
        return trimodal_ntxent(feats, self.temperature, self.alpha_weight, self.op_format, norm)

    def fused_total(self, feats: Sequence[torch.Tensor], norm: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        return trimodal_ntxent_total(feats, self.temperature, self.alpha_weight, self.op_format, norm)


def calculate_losses(output_dict: Dict[str, torch.Tensor], loss_prefix: str, loss_fn: NTXentLoss) -> Dict[str, torch.Tensor]:
    """Fused replacement of TriCoLoNet._calculate_losses (tricolo_net.py:56-65).

    Usable as a monkey-patch:  TriCoLoNet._calculate_losses = lambda self, out, prefix:
    calculate_losses(out, prefix, self.loss_fn).
    """
    keys = list(output_dict.keys())
    losses, total = loss_fn.fused_total([output_dict[k] for k in keys])
    loss_dict = {}
    for p, (a, b) in enumerate(combinations(keys, 2)):
        loss_dict[f"{loss_prefix}/{a[:-9]}_{b[:-9]}_loss"] = losses[p]
    # sum(loss_dict.values()) of tricolo_net.py:64, formed by the forward itself (fp32, pair order)
    loss_dict[f"{loss_prefix}/total_loss"] = total
    return loss_dict
