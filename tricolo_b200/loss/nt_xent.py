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

from itertools import combinations
from typing import Dict, List, Sequence, Tuple

import torch

from .. import ops

# fp16 operands by default: same tcgen05 kind::f16 throughput as bf16, 3 more
# mantissa bits; normalised embeddings lie in [-1, 1] so range is a non-issue.
# bf16 operands miss the 1e-3 gradient tolerance ("Operand format").
DEFAULT_OP_FORMAT = ops.F16


def _pairs(n: int) -> List[Tuple[int, int]]:
    return list(combinations(range(n), 2))


class _FusedNTXent(torch.autograd.Function):
    """losses[p] for every pair (a, b), a < b, of the given feature matrices."""

    @staticmethod
    def forward(ctx, temperature: float, alpha: float, op_format: int, pairs, *feats: torch.Tensor):
        inv_tau = 1.0 / float(temperature)
        feats = [f.detach() for f in feats]
        zs, invs, xs = ops.l2norm_fwd(feats, op_format)
        zrows = [zs[a] for a, _ in pairs]
        zcols = [zs[b] for _, b in pairs]
        row_sum, col_sum, diag2 = ops.ntxent_fwd(zrows, zcols, 0, inv_tau, op_format)
        lse2_row, lse2_col, _parts, loss = ops.ntxent_finalize(row_sum, col_sum, diag2, 0, inv_tau, alpha)
        ctx.cfg = (inv_tau, float(alpha), op_format, tuple(pairs))
        ctx.save_for_backward(lse2_row, lse2_col, *xs, *zs, *invs)
        return loss

    @staticmethod
    def backward(ctx, grad_losses: torch.Tensor):
        inv_tau, alpha, op_format, pairs = ctx.cfg
        saved = ctx.saved_tensors
        lse2_row, lse2_col = saved[0], saved[1]
        n = (len(saved) - 2) // 3
        xs, zs, invs = saved[2:2 + n], saved[2 + n:2 + 2 * n], saved[2 + 2 * n:]
        grad_losses = grad_losses.to(torch.float32).contiguous()
        zts, ld_t = ops.transpose_16bit(zs)
        jobs, owners = [], []
        for m in range(n):
            if not ctx.needs_input_grad[4 + m]:
                continue
            segs = []
            for p, (a, b) in enumerate(pairs):
                if m == a:  # self is the row side: row softmax weight alpha (nt_xent.py:71,74)
                    segs.append(ops.BwdSegmentSpec(zs[b], zts[b], lse2_row[p], lse2_col[p], grad_losses[p:p + 1],
                                                   alpha, 1.0 - alpha))
                elif m == b:  # self is the column side (nt_xent.py:72,74)
                    segs.append(ops.BwdSegmentSpec(zs[a], zts[a], lse2_col[p], lse2_row[p], grad_losses[p:p + 1],
                                                   1.0 - alpha, alpha))
            if segs:
                jobs.append(ops.BwdJobSpec(zs[m], xs[m], invs[m], segs))
                owners.append(m)
        grads: List = [None] * n
        if jobs:
            dxs = ops.ntxent_bwd(jobs, zs[0].shape[0], 0, ld_t, inv_tau, op_format)
            for m, dx in zip(owners, dxs):
                grads[m] = dx
        return (None, None, None, None, *grads)


def trimodal_ntxent(feats: Sequence[torch.Tensor], temperature: float, alpha: float,
                    op_format: int = DEFAULT_OP_FORMAT) -> torch.Tensor:
    """Per-pair losses [n_pairs] (fp32) for all unordered pairs of `feats`, in combinations() order."""
    feats = list(feats)
    if len(feats) < 2 or len(feats) > 3:
        raise ValueError("expected 2 or 3 feature matrices")
    b, d = feats[0].shape
    for f in feats:
        if f.shape != (b, d):
            raise ValueError(f"feature matrices must share one [batch, dim] shape, got {tuple(f.shape)} vs {(b, d)}")
    dt = feats[0].dtype
    feats = [f if f.dtype == dt else f.to(dt) for f in feats]
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
        if not norm:
            # nt_xent.py:55 allows norm=False; the fused sum-exp relies on |cos| <= 1.
            raise NotImplementedError(
                "tricolo_b200.NTXentLoss supports norm=True only (the only mode TriCoLoNet uses, "
                "tricolo_net.py:63); there is no fallback path")
        return trimodal_ntxent([zis, zjs], self.temperature, self.alpha_weight, self.op_format)[0]

    def fused(self, feats: Sequence[torch.Tensor]) -> torch.Tensor:
        return trimodal_ntxent(feats, self.temperature, self.alpha_weight, self.op_format)


def calculate_losses(output_dict: Dict[str, torch.Tensor], loss_prefix: str, loss_fn: NTXentLoss) -> Dict[str, torch.Tensor]:
    """Fused replacement of TriCoLoNet._calculate_losses (tricolo_net.py:56-65).

    Usable as a monkey-patch:  TriCoLoNet._calculate_losses = lambda self, out, prefix:
    calculate_losses(out, prefix, self.loss_fn).
    """
    keys = list(output_dict.keys())
    losses = loss_fn.fused([output_dict[k] for k in keys])
    loss_dict = {}
    for p, (a, b) in enumerate(combinations(keys, 2)):
        loss_dict[f"{loss_prefix}/{a[:-9]}_{b[:-9]}_loss"] = losses[p]
    loss_dict[f"{loss_prefix}/total_loss"] = sum(loss_dict.values())
    return loss_dict
