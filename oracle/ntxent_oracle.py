"""Oracle for the NT-Xent / InfoNCE loss — test infrastructure, not product code.

Restates tricolo/loss/nt_xent.py:24-74 (forward) and its autograd (closed form,
SURVEY.md §8a-L4) and the trimodal pairing of tricolo/model/tricolo_net.py:56-65.
"""
from __future__ import annotations

from itertools import combinations

import numpy as np

EPS = 1e-12  # torch.nn.functional.normalize default, nt_xent.py:56-57


def _normalise(x: np.ndarray):
    """x / max(||x||_2, eps) row-wise (nt_xent.py:56-57). Returns (z, clamped_norm)."""
    n = np.maximum(np.sqrt((x * x).sum(axis=1, keepdims=True)), EPS)
    return x / n, n


def _lse(z: np.ndarray, axis: int) -> np.ndarray:
    m = z.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(z - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def ntxent_forward(zis, zjs, temperature: float, alpha: float) -> float:
    """alpha * CE_rows(Z) + (1 - alpha) * CE_cols(Z), Z = zi_hat zj_hat^T / tau, identity targets.

    nt_xent.py:68-74 with _softXEnt (nt_xent.py:15-22): -(I * log_softmax).sum() / B.
    """
    zi, _ = _normalise(np.asarray(zis, dtype=np.float64))
    zj, _ = _normalise(np.asarray(zjs, dtype=np.float64))
    z = zi @ zj.T / temperature
    d = np.diagonal(z)
    loss_a = float(np.mean(_lse(z, 1) - d))  # rows of logits_ab        (nt_xent.py:71)
    loss_b = float(np.mean(_lse(z, 0) - d))  # rows of logits_ba = Z^T  (nt_xent.py:72)
    return alpha * loss_a + (1.0 - alpha) * loss_b


def ntxent_forward_backward(zis, zjs, temperature: float, alpha: float, grad_out: float = 1.0, norm: bool = True):
    """Loss and d(loss)/d(zis), d(loss)/d(zjs) in fp64 (closed form of the autograd graph).
    norm=False: nt_xent.py:55 skips the F.normalize of :56-57; the logits are the raw inner products / tau.

    dL/dZ = [alpha softmax_rows(Z) + (1-alpha) softmax_cols(Z) - I] / B
    dzi_hat = dL/dZ zj_hat / tau ; dzj_hat = dL/dZ^T zi_hat / tau
    dx = (g - (g . x_hat) x_hat) / ||x||   (no projection term where the norm was clamped)
    """
    xi = np.asarray(zis, dtype=np.float64)
    xj = np.asarray(zjs, dtype=np.float64)
    b = xi.shape[0]
    if norm:
        zi, ni = _normalise(xi)
        zj, nj = _normalise(xj)
    else:
        zi, zj = xi, xj
    z = zi @ zj.T / temperature
    lr = _lse(z, 1)
    lc = _lse(z, 0)
    d = np.diagonal(z)
    loss = alpha * float(np.mean(lr - d)) + (1.0 - alpha) * float(np.mean(lc - d))
    g = (alpha * np.exp(z - lr[:, None]) + (1.0 - alpha) * np.exp(z - lc[None, :]) - np.eye(b)) / b
    g *= grad_out
    gzi = g @ zj / temperature
    gzj = g.T @ zi / temperature
    if not norm:
        return loss, gzi, gzj

    def norm_bwd(gz, zh, n, x):
        clamped = (np.sqrt((x * x).sum(axis=1, keepdims=True)) < EPS)
        proj = (gz * zh).sum(axis=1, keepdims=True)
        proj = np.where(clamped, 0.0, proj)
        return (gz - proj * zh) / n

    return loss, norm_bwd(gzi, zi, ni, xi), norm_bwd(gzj, zj, nj, xj)


def pair_order(keys):
    """combinations(keys, 2) in insertion order, tricolo_net.py:59-61."""
    return list(combinations(list(keys), 2))


def trimodal_forward_backward(features: dict, temperature: float, alpha: float, prefix: str = "train_loss",
                              norm: bool = True):
    """tricolo_net.py:56-65: one bimodal loss per unordered key pair, summed.

    Returns (loss_dict with the reference's key names, grads dict keyed like `features`).
    """
    losses = {}
    grads = {k: np.zeros(np.asarray(v).shape, dtype=np.float64) for k, v in features.items()}
    for a, b in pair_order(features.keys()):
        loss, ga, gb = ntxent_forward_backward(features[a], features[b], temperature, alpha, norm=norm)
        losses[f"{prefix}/{a[:-9]}_{b[:-9]}_loss"] = loss  # key[:-9] strips "_features" (tricolo_net.py:62)
        grads[a] += ga
        grads[b] += gb
    losses[f"{prefix}/total_loss"] = sum(losses.values())
    return losses, grads


# ---------------------------------------------------------------------------
# CPU-PyTorch restatement used for the timed CPU baseline (fp32, autograd), so
# the baseline executes the same library ops as the reference: normalise, two
# matmuls, two log_softmax, autograd backward.
# ---------------------------------------------------------------------------
def torch_ntxent(zis, zjs, temperature: float, alpha: float):
    import torch

    hi = zis / zis.norm(p=2, dim=1, keepdim=True).clamp_min(EPS)
    hj = zjs / zjs.norm(p=2, dim=1, keepdim=True).clamp_min(EPS)
    eye = torch.eye(hi.shape[0], device=hi.device, dtype=torch.float32)
    lab = hi @ hj.t() / temperature
    lba = hj @ hi.t() / temperature
    la = -(eye * torch.log_softmax(lab, dim=1)).sum() / lab.shape[0]
    lb = -(eye * torch.log_softmax(lba, dim=1)).sum() / lba.shape[0]
    return alpha * la + (1 - alpha) * lb


def torch_trimodal(features: dict, temperature: float, alpha: float, prefix: str = "train_loss"):
    out = {}
    for a, b in pair_order(features.keys()):
        out[f"{prefix}/{a[:-9]}_{b[:-9]}_loss"] = torch_ntxent(features[a], features[b], temperature, alpha)
    out[f"{prefix}/total_loss"] = sum(out.values())
    return out
