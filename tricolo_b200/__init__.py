"""tricolo_b200 — sm_100a implementation of TriCoLo's embedding-similarity hot path.

Drop-in surface (same names and call signatures as the reference):
  tricolo_b200.loss.nt_xent.NTXentLoss           <- tricolo/loss/nt_xent.py:6
  tricolo_b200.loss.nt_xent.calculate_losses     <- TriCoLoNet._calculate_losses, tricolo_net.py:56-65
  tricolo_b200.evaluation.eval_retrieval.compute_metrics (+ the inner trio)
                                                 <- tricolo/evaluation/eval_retrieval.py:249
Importing the package loads lib/libtricolo_b200.so and fails if it is missing.
"""
from . import _lib  # noqa: F401  (fails loudly when the CUDA library is absent)

__all__ = ["_lib"]
