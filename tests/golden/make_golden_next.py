"""Golden outputs of the UNMODIFIED reference for the rows SURVEY.md §8f marks "next":
device-resident evaluation hand-off (validation_step/_collate_output + compute_metrics), the self-retrieval
branch of compute_nearest_neighbors and the nearest.jsonl writer.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden_next.py
Writes tests/golden/next_rows.json and next_rows.npz.  Nothing from the reference is copied: only its outputs on
seeded synthetic inputs (generators: oracle/retrieval_oracle.py) are stored, together with this script."""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from make_golden import install_shim  # noqa: E402
from oracle import retrieval_oracle as RO  # noqa: E402


def main():
    install_shim()
    records = []

    class _Writer:
        def write(self, obj):
            records.append(obj)

    sys.modules["jsonlines"].open = lambda *a, **k: _Writer()
    from tricolo.evaluation import eval_retrieval as ER
    from tricolo.model.tricolo_net import TriCoLoNet

    out, npz = {}, {}
    cwd = os.getcwd()
    os.chdir(tempfile.mkdtemp())

    # ---- validation epoch: (data_dict, output_dict) batches -> _collate_output -> compute_metrics
    for name, kw in {"VAL_TRI": {}, "VAL_BI_VOXEL": {"drop": "image_features"}}.items():
        batches = RO.make_val_batches()
        if "drop" in kw:
            for _, o in batches:
                del o[kw["drop"]]
        fake = types.SimpleNamespace(val_test_step_outputs=[(d, dict(o)) for d, o in batches])
        emb = TriCoLoNet._collate_output(fake)
        del records[:]
        np.random.seed(0)
        m = ER.compute_metrics("Text2Shape", emb)
        out[name] = {k: (v.tolist() if hasattr(v, "tolist") else float(v)) for k, v in m.items()}
        text, gal, labels, fit_labels, m2l, nq, l2m = ER.construct_embeddings_matrix("Text2Shape", emb)
        out[name]["n_gallery"] = int(gal.shape[0])
        out[name]["label_to_model_id_head"] = [l2m[i] for i in range(8)]
        npz[f"{name}.labels"] = labels.astype(np.int32)
        npz[f"{name}.gallery_head"] = gal[:4].astype(np.float32)
        if name == "VAL_TRI":
            # nearest.jsonl as the reference writes it (np.random.seed(0) before compute_metrics)
            out["NEAREST"] = {"n": len(records), "head": records[:12],
                              "groundtruth_all": [r["groundtruth"] for r in records],
                              "retrieved_all": [r["retrieved_models"] for r in records]}
            npz["NEAREST.distance"] = np.asarray([r["distance"] for r in records], dtype=np.float64)

    # ---- self-retrieval: fit == query
    x = RO.make_self_retrieval()
    dist, idx, _ = ER.compute_nearest_neighbors(x, x, 5)
    npz["SELF.indices"] = idx.astype(np.int32)
    npz["SELF.distances"] = dist
    xb = RO.make_self_retrieval(seed=6, n=9000, dim=32)  # > 8000 queries: blocks of 3000 with range_start
    dist, idx, _ = ER.compute_nearest_neighbors(xb, xb, 5)
    npz["SELFBLK.indices"] = idx.astype(np.int32)
    npz["SELFBLK.distances_head"] = dist[:4]
    npz["SELFBLK.distances_tail"] = dist[-4:]
    # ---- triplet loss: the reference module on seeded inputs, autograd gradients
    import torch
    from tricolo.loss.triplet import TripletLoss
    from oracle import triplet_oracle as TO

    trip = {}
    for name, (margin, noise, norm) in {"T1_SEMI": (0.025, 3.0, True), "T2_HARD": (1e-7, 3.0, True),
                                        "T3_NONE": (0.025, 0.35, True), "T4_RAW": (0.025, 3.0, False)}.items():
        zis, zls = TO.make_triplet_case(noise=noise, normalise=norm)
        a = torch.from_numpy(zis).requires_grad_(True)
        b = torch.from_numpy(zls).requires_grad_(True)
        try:
            loss = TripletLoss(margin)(a, b)
            loss.backward()
            trip[name] = {"loss": float(loss), "margin": margin, "noise": noise, "normalise": norm}
            npz[f"{name}.d_zis"] = a.grad.numpy()
            npz[f"{name}.d_zls"] = b.grad.numpy()
        except ZeroDivisionError:
            trip[name] = {"loss": None, "error": "ZeroDivisionError", "margin": margin, "noise": noise, "normalise": norm}
    out["TRIPLET"] = trip
    os.chdir(cwd)
    out["numpy"] = np.__version__
    with open(os.path.join(HERE, "next_rows.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "next_rows.npz"), **npz)
    print({k: (v if not isinstance(v, dict) else {kk: vv for kk, vv in v.items() if kk in ("recall_rate", "ndcg", "mrr", "n_gallery", "n")}) for k, v in out.items()})


if __name__ == "__main__":
    main()
