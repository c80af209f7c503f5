"""Golden outputs of the UNMODIFIED reference at the two large configurations:

  C4  TriCoLoNet._calculate_losses (tricolo_net.py:56-65 -> nt_xent.py:24-74) on the concatenated global batch
      B = 8192, dim 512, seed 7 (bench.py's own inputs): 3 pair losses + total, gradient norms, every 64th gradient row.
  C5  one 3000-query block (the reference's own block size, eval_retrieval.py:110) against the full 200 000-shape
      gallery: _compute_nearest_neighbors_cosine (:68-99) + compute_pr_at_k (:149-207), fp64 as in the reference.
      Stored: top-5 indices, the rank of the ground truth, the fp64 margins that decide which rows an fp32 path must
      reproduce exactly, and the metric dict.

Run in the build container only (needs /root/reference; ~20 GB of host memory, a few minutes):
    python tests/golden/make_golden_large.py
"""
import json
import os
import sys
import time
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from make_golden import install_shim  # noqa: E402
from oracle import retrieval_oracle as RO  # noqa: E402  (input generator only)

RANK_DELTA = 2e-6  # bound on |fp32-accumulated - fp64| similarity of bf16-exact unit vectors (512 terms)


def c4_features(batch=8192, dim=512, seed=7):
    """bench.py make_features: correlated modalities, base + 0.5 noise."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(batch, dim, generator=g)
    return {k: (base + 0.5 * torch.randn(batch, dim, generator=g)).contiguous()
            for k in ("text_features", "image_features", "voxel_features")}


def main():
    install_shim()
    from tricolo.evaluation import eval_retrieval as ER
    from tricolo.loss.nt_xent import NTXentLoss
    from tricolo.model.tricolo_net import TriCoLoNet

    out = {"torch": torch.__version__, "numpy": np.__version__}
    # ------------------------------------------------------------------ C4
    torch.set_num_threads(os.cpu_count() or 1)
    fake_self = types.SimpleNamespace(loss_fn=NTXentLoss(temperature=0.1, alpha_weight=0.25))
    feats = {k: v.requires_grad_(True) for k, v in c4_features().items()}
    t0 = time.perf_counter()
    losses = TriCoLoNet._calculate_losses(fake_self, feats, "train_loss")
    losses["train_loss/total_loss"].backward()
    out["c4"] = {"batch": 8192, "dim": 512, "seed": 7, "temperature": 0.1, "alpha_weight": 0.25,
                 "losses": {k: float(v) for k, v in losses.items()},
                 "grad_norm": {k: float(v.grad.double().norm()) for k, v in feats.items()},
                 "row_stride": 64, "reference_cpu_seconds": time.perf_counter() - t0}
    np.savez_compressed(os.path.join(HERE, "c4_grads.npz"),
                        **{k: v.grad.numpy()[::64].astype(np.float32) for k, v in feats.items()})
    print("C4", out["c4"])
    # ------------------------------------------------------------------ C5 block
    text, gal, labels = RO.make_large_retrieval(seed=0, n_shapes=200_000, n_queries=3000, dim=512)
    text64 = text.astype(np.float64)  # the reference's text matrix is float64 (eval_retrieval.py:25)
    t0 = time.perf_counter()
    dist, idx, sort_idx = ER._compute_nearest_neighbors_cosine(gal, text64, 5, False)
    t_nn = time.perf_counter() - t0
    fit_labels = np.arange(gal.shape[0])
    t0 = time.perf_counter()
    metrics = ER.compute_pr_at_k(idx, sort_idx, labels, 5, text.shape[0], fit_labels)
    t_pr = time.perf_counter() - t0
    q = text.shape[0]
    rank = np.empty(q, dtype=np.int64)
    for i in range(q):
        rank[i] = int(np.nonzero(sort_idx[i] == labels[i])[0][0]) + 1
    del sort_idx
    sim = np.dot(text64, gal.T.astype(np.float64))
    s_gt = sim[np.arange(q), labels][:, None]
    rank_lo = (sim > s_gt + RANK_DELTA).sum(axis=1) + 1           # every fp32 path must land in [lo, hi]
    rank_hi = (sim >= s_gt - RANK_DELTA).sum(axis=1)
    top6 = -np.sort(-sim, axis=1)[:, :6]
    margin = np.min(-np.diff(top6, axis=1), axis=1)
    out["c5_block"] = {"queries": q, "gallery": int(gal.shape[0]), "dim": 512, "seed": 0, "rank_delta": RANK_DELTA,
                       "metrics": {k: (v.tolist() if hasattr(v, "tolist") else float(v)) for k, v in metrics.items()},
                       "reference_seconds": {"nearest_neighbors": t_nn, "compute_pr_at_k": t_pr},
                       "rows_rank_exact": int((rank_lo == rank_hi).sum()), "rows_top5_safe": int((margin > 1e-5).sum())}
    np.savez_compressed(os.path.join(HERE, "c5_block.npz"), indices=idx.astype(np.int32), rank=rank.astype(np.int32),
                        rank_lo=rank_lo.astype(np.int32), rank_hi=rank_hi.astype(np.int32), margin=margin,
                        top5_val=top6[:, :5], gt_sim=s_gt[:, 0], distances_head=dist[:4], distances_tail=dist[-4:])
    print("C5 block", out["c5_block"])
    with open(os.path.join(HERE, "large_outputs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
