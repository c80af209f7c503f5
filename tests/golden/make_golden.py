"""Generate tests/golden/*.json|npz by running the UNMODIFIED reference code.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference's two hot-path files import once `lightning.pytorch`, `jsonlines`,
`hydra` and `clip` are stubbed.  Nothing from the
reference is copied: only its numeric outputs on seeded synthetic inputs are
stored, together with this script.
"""
import json
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = os.environ.get("TRICOLO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import retrieval_oracle as RO  # noqa: E402  (data generators only)


def install_shim():
    from oracle.reference_shim import install_shim as _install  # the one shim used by tests, goldens and the baseline arm

    _install(REF)


def loss_cases():
    """(name, feature dict builder) — SURVEY.md §8a G1-G3, KAT1-3."""
    def g1():
        g = torch.Generator().manual_seed(1234)
        t = torch.randn(128, 512, generator=g); v = torch.randn(128, 512, generator=g)
        return {"text_features": t, "voxel_features": v}

    def g2():
        g = torch.Generator().manual_seed(1234)
        t = torch.randn(256, 512, generator=g); i = torch.randn(256, 512, generator=g); v = torch.randn(256, 512, generator=g)
        return {"text_features": t, "image_features": i, "voxel_features": v}

    def g3():
        g = torch.Generator().manual_seed(7)
        base = torch.randn(256, 512, generator=g)
        t = base + 0.5 * torch.randn(256, 512, generator=g)
        i = base + 0.5 * torch.randn(256, 512, generator=g)
        v = base + 0.5 * torch.randn(256, 512, generator=g)
        return {"text_features": t, "image_features": i, "voxel_features": v}

    def kat1():
        e = torch.eye(128, 512)
        return {"text_features": e.clone(), "voxel_features": e.clone()}

    def kat2():
        o = torch.ones(128, 512)
        return {"text_features": o.clone(), "voxel_features": o.clone()}

    def kat3():
        g = torch.Generator().manual_seed(1)
        t = torch.randn(64, 512, generator=g); v = torch.randn(64, 512, generator=g)
        t[0] = 0
        return {"text_features": t, "voxel_features": v}

    def ragged():  # batch not a multiple of the 128-row tile
        g = torch.Generator().manual_seed(99)
        t = torch.randn(200, 512, generator=g); i = torch.randn(200, 512, generator=g); v = torch.randn(200, 512, generator=g)
        return {"text_features": t, "image_features": i, "voxel_features": v}

    def dim256():
        g = torch.Generator().manual_seed(5)
        t = torch.randn(384, 256, generator=g); v = torch.randn(384, 256, generator=g)
        return {"text_features": t, "image_features": v}

    return {"G1": g1, "G2": g2, "G3": g3, "KAT1": kat1, "KAT2": kat2, "KAT3": kat3, "RAGGED200": ragged, "DIM256": dim256}


def main():
    install_shim()
    from tricolo.loss.nt_xent import NTXentLoss
    from tricolo.evaluation import eval_retrieval as ER
    from tricolo.model.tricolo_net import TriCoLoNet

    tau, alpha = 0.1, 0.25
    fake_self = types.SimpleNamespace(loss_fn=NTXentLoss(temperature=tau, alpha_weight=alpha))
    out = {"temperature": tau, "alpha_weight": alpha, "torch": torch.__version__, "numpy": np.__version__, "cases": {}}
    grads_npz = {}
    for name, build in loss_cases().items():
        for variant in ("fp32", "bf16"):  # bf16: inputs rounded to bf16, reference math still fp32
            feats = build()
            if variant == "bf16":
                feats = {k: v.bfloat16().float() for k, v in feats.items()}
            feats = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
            losses = TriCoLoNet._calculate_losses(fake_self, feats, "train_loss")
            losses["train_loss/total_loss"].backward()
            rec = {"losses": {k: float(v) for k, v in losses.items()}, "grad_norm": {}, "grad_sample": {}}
            for k, v in feats.items():
                g = v.grad
                rec["grad_norm"][k] = float(g.double().norm())
                flat = g.flatten()
                rec["grad_sample"][k] = [float(x) for x in flat[::997][:64]]
                if variant == "bf16" and name in ("G1", "G3", "KAT3", "RAGGED200"):
                    grads_npz[f"{name}.{variant}.{k}"] = g.numpy()[::4].astype(np.float32)  # every 4th row
            out["cases"][f"{name}.{variant}"] = rec
    # argument order matters (alpha != 0.5)
    g = torch.Generator().manual_seed(3)
    a = torch.randn(64, 512, generator=g); b = torch.randn(64, 512, generator=g)
    out["order"] = {"ab": float(fake_self.loss_fn(a, b)), "ba": float(fake_self.loss_fn(b, a))}

    # ---------------- retrieval ----------------
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())
    ev = {}
    def run_eval(tuples):
        np.random.seed(0)
        m = ER.compute_metrics("Text2ShapeChairTable", {"caption_embedding_tuples": tuples})
        return {k: (v.tolist() if hasattr(v, "tolist") else float(v)) for k, v in m.items()}
    ev["KAT_E1"] = run_eval(RO.make_integer_kat())
    ev["C3.fp32"] = run_eval(RO.make_val_shaped(round_bf16=False))
    ev["C3.bf16"] = run_eval(RO.make_val_shaped(round_bf16=True))
    ev["C3TRI.bf16"] = run_eval(RO.make_val_shaped(round_bf16=True, trimodal_gallery=True))
    ev["SMALL.bf16"] = run_eval(RO.make_val_shaped(seed=3, n_shapes=300, n_queries=1000, dim=128, round_bf16=True))
    # indices / ranks of the reference itself on the tie-free C3.bf16 case
    tuples = RO.make_val_shaped(round_bf16=True)
    text, gal, labels, fit_labels, _, nq, _ = ER.construct_embeddings_matrix("x", {"caption_embedding_tuples": tuples})
    dist, idx, sort_idx = ER.compute_nearest_neighbors(gal, text, 5)
    rank = np.array([int(np.nonzero(sort_idx[i] == labels[i])[0][0]) + 1 for i in range(nq)])
    sim = np.dot(text, gal.T)
    srt = np.sort(sim, axis=1)
    ev["C3.bf16.min_gap"] = float(np.min(np.diff(srt, axis=1)))
    os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "eval_c3_bf16.npz"), indices=idx.astype(np.int32), rank=rank.astype(np.int32),
                        labels=labels.astype(np.int32), distances_head=dist[:8], distances_tail=dist[-8:])
    out["eval"] = ev
    with open(os.path.join(HERE, "reference_outputs.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "loss_grads.npz"), **grads_npz)
    print("wrote", HERE)
    for k, v in out["cases"].items():
        print(k, v["losses"])
    print(json.dumps(ev, indent=1)[:1500])


if __name__ == "__main__":
    main()
