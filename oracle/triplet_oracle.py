"""Oracle for the triplet loss — test infrastructure, not product code.

Restates tricolo/loss/triplet.py: _pairwise_distances (:11-45, including the reference's pairing of the squared
norms at :32) and forward (:202-224: semi-hard terms, hard terms when there is no semi-hard one, mean), plus the
analytic gradient that the reference obtains from autograd.  NumPy, float64 arithmetic on the given inputs; the
selection is made on float32-rounded distances like the reference's comparisons."""
from __future__ import annotations

import numpy as np


def pairwise_distances(zis, zls):
    zis, zls = np.asarray(zis, np.float64), np.asarray(zls, np.float64)
    dot = zls @ zis.T                                     # :20
    a2 = np.sum(zls * zls, axis=1)                        # :22
    b2 = np.sum(zis * zis, axis=1)                        # :23
    d2 = a2[None, :] - 2.0 * dot + b2[:, None]            # :32
    d2 = np.where(d2 < 0, 0.0, d2)                        # :35
    return np.sqrt(d2), d2                                # :37-43 (D = 0 exactly where d2 == 0)


def selection(D, margin):
    """Boolean [B,B] mask of the (i, j) terms and the mode: 0 semi-hard, 1 hard, 2 none."""
    Df = D.astype(np.float32)
    dii = np.diag(Df)[:, None]
    off = ~np.eye(D.shape[0], dtype=bool)
    semi = off & (dii < Df) & (Df < (dii + np.float32(margin)).astype(np.float32))   # :208
    if semi.any():
        return semi, 0
    hard = off & (Df < dii)                                                            # :216
    return hard, (1 if hard.any() else 2)


def triplet_forward_backward(zis, zls, margin):
    """(loss, d_zis, d_zls, info) with info = (n_semi, n_hard, mode)."""
    zis, zls = np.asarray(zis, np.float64), np.asarray(zls, np.float64)
    D, d2 = pairwise_distances(zis, zls)
    sel, mode = selection(D, margin)
    if mode == 2:
        raise ZeroDivisionError("division by zero")
    n = int(sel.sum())
    dii = np.diag(D)[:, None]
    loss = float(np.sum(np.where(sel, dii - D + margin, 0.0)) / n)
    # dLoss/dD: -1/n on selected (i, j), +count_i/n on (i, i); dD/dd2 = 1/(2D) where D > 0
    W = np.where(sel, -1.0 / n, 0.0)
    W[np.arange(len(D)), np.arange(len(D))] = sel.sum(axis=1) / n
    with np.errstate(divide="ignore", invalid="ignore"):
        V = np.where(D > 0, W / (2.0 * D), 0.0)
    d_zls = 2.0 * V.sum(axis=0)[:, None] * zls - 2.0 * V @ zis
    d_zis = 2.0 * V.sum(axis=1)[:, None] * zis - 2.0 * V.T @ zls
    Df = D.astype(np.float32)
    off = ~np.eye(len(D), dtype=bool)
    d32 = np.diag(Df)[:, None]
    n_semi = int((off & (d32 < Df) & (Df < (d32 + np.float32(margin)).astype(np.float32))).sum())
    n_hard = int((off & (Df < d32)).sum())
    return loss, d_zis, d_zls, (n_semi, n_hard, mode)


def make_triplet_case(seed=0, batch=96, dim=512, noise=0.35, normalise=True):
    """Paired features (zis_i, zls_i correlated), float32, bf16-rounded so that every dtype path sees the same values."""
    from .retrieval_oracle import bf16_round

    rng = np.random.default_rng(seed)
    base = rng.standard_normal((batch, dim)).astype(np.float32)
    zis = base + noise * rng.standard_normal((batch, dim)).astype(np.float32)
    zls = base + noise * rng.standard_normal((batch, dim)).astype(np.float32)
    if normalise:
        zis /= np.linalg.norm(zis, axis=1, keepdims=True)
        zls /= np.linalg.norm(zls, axis=1, keepdims=True)
    return bf16_round(zis), bf16_round(zls)
