"""Drop-in for tricolo/loss/triplet.py (TripletLoss, selected by loss.name=TripletLoss, config/config.yaml:102-104).

Same constructor and call signature: TripletLoss(margin)(zis, zls) -> 0-dim tensor, differentiable w.r.t. both inputs.
The B x B distance matrix, the semi-hard / hard selection (triplet.py:202-224) and the backward run in the CUDA
library (tcl_triplet_fwd / tcl_triplet_bwd); like the reference, the forward is synchronous (the reference's Python
loops compare device scalars), which is where the "no term at all" case raises ZeroDivisionError as it does there."""
from __future__ import annotations

import torch

from .. import _lib as L

LIB = L.LIB


class _Triplet(torch.autograd.Function):
    @staticmethod
    def forward(ctx, margin: float, zis: torch.Tensor, zls: torch.Tensor):
        dev = L.require_cuda(zis, zls)
        if zis.shape != zls.shape or zis.dim() != 2 or zis.dtype != zls.dtype:
            raise ValueError("TripletLoss: zis and zls must be [B, D] tensors of one dtype")
        zis_c, zls_c = zis.detach().contiguous(), zls.detach().contiguous()
        b, d = zis_c.shape
        ws_bytes = LIB.tcl_triplet_workspace_bytes(b)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        info = torch.empty((4,), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            L.check(LIB.tcl_triplet_fwd(L.ptr(zis_c), L.ptr(zls_c), L.dtype_code(zis_c), b, d, zis_c.stride(0), float(margin),
                                        L.ptr(loss), L.ptr(info), L.ptr(ws), ws_bytes, L.stream_ptr(dev)))
        n_semi, n_hard, mode, _ = info.tolist()
        if mode != 0:
            print("loss_list is 0")  # triplet.py:214
        if mode == 2:
            raise ZeroDivisionError("division by zero")  # triplet.py:222 with an empty loss_list
        ctx.margin = float(margin)
        ctx.save_for_backward(zis_c, zls_c, ws)
        return loss.to(zis.dtype) if zis.dtype != torch.float64 else loss.double()

    @staticmethod
    def backward(ctx, grad_loss):
        zis_c, zls_c, ws = ctx.saved_tensors
        dev = zis_c.device
        b, d = zis_c.shape
        g = grad_loss.detach().to(torch.float32).reshape(1).contiguous()
        d_zis = torch.empty_like(zis_c) if ctx.needs_input_grad[1] else None
        d_zls = torch.empty_like(zls_c) if ctx.needs_input_grad[2] else None
        with torch.cuda.device(dev):
            L.check(LIB.tcl_triplet_bwd(L.ptr(zis_c), L.ptr(zls_c), L.dtype_code(zis_c), b, d, zis_c.stride(0), ctx.margin,
                                        L.ptr(g), L.ptr(ws), ws.numel(), L.ptr(d_zis), L.ptr(d_zls), L.stream_ptr(dev)))
        return None, d_zis, d_zls


class TripletLoss(torch.nn.Module):
    """TripletLoss(margin).forward(zis, zls) — triplet.py:5-9, 202-224. No parameters, no buffers."""

    def __init__(self, margin):
        super().__init__()
        self.margin = margin

    def forward(self, zis, zls):
        return _Triplet.apply(float(self.margin), zis, zls)
