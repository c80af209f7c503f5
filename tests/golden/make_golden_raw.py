"""Goldens of the UNMODIFIED reference for NTXentLoss.forward(zis, zjs, norm=False) (nt_xent.py:55).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_raw.py
For every case of tests/cases.py:RAW_CASES: the per-pair losses (reference forward on every unordered pair of the
features, pair order of tricolo_net.py:59-61, summed like :64) and, after backward of the sum, every 4th gradient row.
Only numeric outputs on seeded synthetic inputs are stored.
"""
import json
import os
import sys
from itertools import combinations

import numpy as np
import torch

REF = os.environ.get("TRICOLO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from tests.cases import RAW_CASES, raw_case  # noqa: E402


def main():
    from oracle.reference_shim import install_shim

    install_shim(REF)
    from tricolo.loss.nt_xent import NTXentLoss

    tau, alpha = 0.1, 0.25
    loss_fn = NTXentLoss(temperature=tau, alpha_weight=alpha)
    out = {"temperature": tau, "alpha_weight": alpha, "torch": torch.__version__, "cases": {}}
    grads = {}
    for name in RAW_CASES:
        feats = {k: v.clone().requires_grad_(True) for k, v in raw_case(name).items()}
        losses = {}
        for a, b in combinations(feats.keys(), 2):
            losses[f"{a[:-9]}_{b[:-9]}_loss"] = loss_fn(feats[a], feats[b], norm=False)
        total = sum(losses.values())
        total.backward()
        out["cases"][name] = {"losses": {k: float(v) for k, v in losses.items()}, "total": float(total),
                              "grad_norm": {k: float(v.grad.double().norm()) for k, v in feats.items()}}
        for k, v in feats.items():
            grads[f"{name}.{k}"] = v.grad.numpy()[::4].astype(np.float32)
        print(name, out["cases"][name])
    with open(os.path.join(HERE, "raw_outputs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "raw_grads.npz"), **grads)


if __name__ == "__main__":
    main()
