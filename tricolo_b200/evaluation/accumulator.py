"""Device-resident evaluation hand-off.

Replaces, for the retrieval metrics, the host round trip of the reference's validation/test epoch:
  validation_step / test_step   tricolo_net.py:78-88, 98-109   per-batch .cpu().numpy() of every feature tensor
  _collate_output               tricolo_net.py:125-158         vstack + shape = zeros + image + voxel + O(Q) tuple list
  construct_embeddings_matrix   eval_retrieval.py:6-65         first-occurrence gallery de-duplication, labels
  compute_metrics               eval_retrieval.py:249-278
The feature tensors stay on the GPU; the only host work is one pass over the model-id strings.  The gallery is built
by one kernel (tcl_gather_sum_cast16: gather of the first occurrences + modality sum + 16-bit cast), the metrics by the
same retrieval kernels as compute_metrics.  embeddings_dict() materialises the reference's list-of-tuples format for
the on-disk writers (output.p, nearest.jsonl)."""
from __future__ import annotations

import pickle
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from .. import ops
from . import eval_retrieval as ER

SHAPE_KEYS = ("image_features", "voxel_features")  # summation order of tricolo_net.py:135-139


class RetrievalAccumulator:
    def __init__(self, operand_format: int = ER.OPERAND_FORMAT):
        self.operand_format = operand_format
        self.clear()

    def clear(self):
        self._feats: Dict[str, List[torch.Tensor]] = {}
        self._model_ids: List[str] = []
        self._categories: List[str] = []

    def __len__(self):
        return len(self._model_ids)

    def update(self, output_dict: Dict[str, torch.Tensor], model_ids: Sequence[str], categories: Sequence[str]):
        """One validation/test step: `output_dict` as returned by TriCoLoNet.forward (device tensors
        text_features and image_features and/or voxel_features, [b, D]); data_dict["model_id"], data_dict["category"]."""
        if "text_features" not in output_dict:
            raise KeyError("RetrievalAccumulator.update: output_dict needs text_features")
        b = output_dict["text_features"].shape[0]
        if len(model_ids) != b or len(categories) != b:
            raise ValueError("RetrievalAccumulator.update: one model id and category per row")
        keys = ["text_features"] + [k for k in SHAPE_KEYS if k in output_dict]
        if self._feats and sorted(self._feats) != sorted(keys):
            raise ValueError("RetrievalAccumulator.update: the set of modalities changed between steps")
        for k in keys:
            t = output_dict[k].detach()
            if not t.is_cuda:
                raise RuntimeError("RetrievalAccumulator keeps features on the GPU; got a CPU tensor (no CPU path)")
            self._feats.setdefault(k, []).append(t)
        self._model_ids.extend(model_ids)
        self._categories.extend(categories)

    # ------------------------------------------------------------------ device side
    def _labels_and_first(self, dataset: str):
        """eval_retrieval.py:38-60 on the id strings only: label per query, row of the first occurrence per label."""
        keys = self._categories if dataset == "Primitives" else self._model_ids  # :45-46
        model_id_to_label, first_rows = {}, []
        labels = np.empty(len(keys), dtype=np.int64)
        for q, mid in enumerate(keys):
            lab = model_id_to_label.get(mid)
            if lab is None:
                lab = len(first_rows)
                model_id_to_label[mid] = lab
                first_rows.append(q)
            labels[q] = lab
        return labels, np.asarray(first_rows, dtype=np.int64), model_id_to_label

    def matrices(self, dataset: str = "Text2Shape"):
        """Device-resident (text16 [Q,D], gallery16 [G,D], labels [Q] int64) + label -> model id."""
        if not self._model_ids:
            raise ValueError("RetrievalAccumulator: no steps accumulated")
        shape_keys = [k for k in SHAPE_KEYS if k in self._feats]
        if not shape_keys:
            raise ValueError("RetrievalAccumulator: no shape modality (image_features / voxel_features)")
        labels, first_rows, model_id_to_label = self._labels_and_first(dataset)
        dev = self._feats["text_features"][0].device
        text = torch.cat(self._feats["text_features"], dim=0)
        srcs = [torch.cat(self._feats[k], dim=0) for k in shape_keys]
        gallery16 = ops.gather_sum_cast16(srcs, torch.from_numpy(first_rows).to(dev), self.operand_format)
        text16 = ops.cast_16bit(text, self.operand_format)
        label_to_model_id = {v: k for k, v in model_id_to_label.items()}
        return text16, gallery16, torch.from_numpy(labels).to(dev), labels, label_to_model_id

    def compute(self, dataset: str = "Text2Shape", print_results: bool = False, write_nearest: bool = False,
                n_neighbors: int = 5) -> dict:
        """Same returned dict as compute_metrics (eval_retrieval.py:249-278)."""
        text16, gallery16, labels_dev, labels, label_to_model_id = self.matrices(dataset)
        val, idx, rank = ER.retrieve(text16, gallery16, labels_dev, n_neighbors, operand_format=self.operand_format)
        indices = idx.cpu().numpy().astype(np.int64)
        pr_at_k = ER.metrics_from_ranks(indices, rank.cpu().numpy().astype(np.int64), labels, n_neighbors,
                                        np.arange(gallery16.shape[0]))
        if write_nearest:
            n_q = len(labels)
            distances = ER._flip_distances_like_reference(val.cpu().numpy().astype(np.float64), 3000 if n_q > 8000 else None)
            # 'groundtruth' is always the raw model id (get_nearest_info reads caption_tuples[idx][2], eval_retrieval.py:
            # 232-240); the Primitives swap (:45-46) applies to the labels only
            nearest = [[label_to_model_id[int(c)] for c in row] for row in indices]
            ER.print_nearest_info(self._categories, list(self._model_ids), nearest, distances)
        if print_results:
            ER._print_results(pr_at_k)
        return pr_at_k

    # ------------------------------------------------------------------ on-disk formats (host by nature)
    def embeddings_dict(self) -> dict:
        """The reference's hand-off format, tricolo_net.py:149-157: (None, category, model_id, text_vec, shape_vec)."""
        text = torch.cat(self._feats["text_features"], dim=0).cpu().numpy()
        shape = np.zeros_like(text)
        for k in SHAPE_KEYS:
            if k in self._feats:
                shape += torch.cat(self._feats[k], dim=0).cpu().numpy()
        tuples = [(None, self._categories[i], self._model_ids[i], text[i], shape[i]) for i in range(text.shape[0])]
        return {"caption_embedding_tuples": tuples}

    def save_predictions(self, path: str):
        """output.p of on_test_epoch_end (tricolo_net.py:118-122): pickle of the embeddings dict."""
        save_predictions(self.embeddings_dict(), path)


def save_predictions(embeddings_dict: dict, path: str):
    with open(path, "wb") as f:
        pickle.dump(embeddings_dict, f)
