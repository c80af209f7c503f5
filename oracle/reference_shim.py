"""Import shim for the UNMODIFIED reference (test / baseline infrastructure, not product code).

The reference's two hot-path files (tricolo/loss/nt_xent.py, tricolo/evaluation/eval_retrieval.py) and
tricolo/model/tricolo_net.py import once `lightning.pytorch`, `jsonlines`, `hydra` and `clip` are stubbed
(SURVEY.md §8c / Appendix A).  Nothing of the reference is copied: it is imported from where it lies
(TRICOLO_REFERENCE, /root/reference) when that exists - in the build container, never on the GPU box.
"""
from __future__ import annotations

import os
import sys
import types


def reference_root():
    for p in (os.environ.get("TRICOLO_REFERENCE"), "/root/reference"):
        if p and os.path.isfile(os.path.join(p, "tricolo", "loss", "nt_xent.py")):
            return p
    return None


def install_shim(root: str) -> None:
    import torch

    def stub(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class LightningModule(torch.nn.Module):  # only what nt_xent.py:6,62 and tricolo_net.py:11-14 touch
        def __init__(self):
            super().__init__()
            self._device = torch.device("cpu")

        @property
        def device(self):
            return self._device

        def _apply(self, fn, *a, **k):  # keep .device in sync with .to()/.cuda()
            r = super()._apply(fn, *a, **k)
            self._device = fn(torch.empty(0, device=self._device)).device
            return r

        def save_hyperparameters(self, *a, **k):
            pass

    ltp = stub("lightning.pytorch", LightningModule=LightningModule)
    stub("lightning", pytorch=ltp)

    class _Writer:
        def write(self, obj):
            pass

    stub("jsonlines", open=lambda *a, **k: _Writer())
    stub("hydra", utils=stub("hydra.utils"))
    stub("clip")
    if root not in sys.path:
        sys.path.insert(0, root)


def load():
    """(NTXentLoss, TriCoLoNet, eval_retrieval module) of the unmodified reference, or None when it is not present."""
    root = reference_root()
    if root is None:
        return None
    try:
        install_shim(root)
        from tricolo.evaluation import eval_retrieval as ER
        from tricolo.loss.nt_xent import NTXentLoss
        from tricolo.model.tricolo_net import TriCoLoNet
    except Exception:
        return None
    return types.SimpleNamespace(NTXentLoss=NTXentLoss, TriCoLoNet=TriCoLoNet, ER=ER, root=root)


def trimodal_loss_fn(temperature: float, alpha: float):
    """fn(feature_dict) -> loss dict through the reference's own TriCoLoNet._calculate_losses (tricolo_net.py:56-65),
    or None when the reference is not importable."""
    ref = load()
    if ref is None:
        return None
    fake_self = types.SimpleNamespace(loss_fn=ref.NTXentLoss(temperature=temperature, alpha_weight=alpha))
    return lambda feats: ref.TriCoLoNet._calculate_losses(fake_self, feats, "train_loss")
