"""Text->shape retrieval and RR@k / NDCG@k / MRR on sm_100a behind the reference's names.

Mirrors tricolo/evaluation/eval_retrieval.py:
    construct_embeddings_matrix (:6-65)   host: tuple list -> matrices + labels
    compute_nearest_neighbors   (:133-146) GPU: similarity GEMM (K2') + top-k (K4)
    compute_pr_at_k             (:149-207) host fp64 finalise from top-k indices and GT ranks
    get_nearest_info            (:210-246) host id mapping
    compute_metrics             (:249-278) orchestration, returns the same dict
plus the tensor-in fast entry `retrieve()` for device-resident inputs (C5-sized runs,
where a list of a million Python tuples would itself be the bottleneck).

Deviations from the reference, all stated:
  * similarities are fp32 (16-bit operands, fp32 accumulate) instead of fp64;
  * ties: (similarity desc, gallery index asc) — the reference's order is undefined;
  * `sort_indices` is not materialised as a [Q, G] int64 array (88 MB at the val size,
    1.6 TB at 1M x 200k): compute_nearest_neighbors returns a lazy `SimilarityOrder`
    that yields the ground-truth ranks on the GPU and materialises rows only on demand.
"""
from __future__ import annotations

import json
from typing import Optional

import numpy as np
import torch

from .. import ops

OPERAND_FORMAT = ops.BF16  # raw, un-normalised embeddings: bf16 keeps fp32's range (BASELINE north_star)
BLOCK_QUERIES = 65536      # upper bound of the queries per GEMM/top-k chunk of the two-kernel form


def two_kernel_block_queries(n_gallery: int) -> int:
    """Queries per chunk of the two-kernel form: the fp32 similarity block is kept around 4 GB, so that a top-k launch
    is several waves of rows even for a small gallery shard (at 25 000 shapes per rank a block of 8192 queries is ONE
    wave of one-row warps: every warp streams, then every warp sorts, and HBM idles in between - 58 % of peak; with
    32768+ rows per launch the phases of different waves overlap)."""
    ld = (n_gallery + 31) // 32 * 32
    return max(8192, min(BLOCK_QUERIES, (4 << 30) // (4 * ld) // 1024 * 1024))


def construct_embeddings_matrix(dataset, embeddings_dict, model_id_to_label=None, label_to_model_id=None,
                                _text_dtype=np.float64):
    """Same outputs as eval_retrieval.py:6-65 (text float64 [Q,D], gallery [G,D], labels int64, ...).
    (_text_dtype: compute_metrics keeps the text matrix in the vectors' own dtype on its way to the GPU - the float64
    copy of :25 changes no value and doubles the bytes.)"""
    assert (model_id_to_label is None) == (label_to_model_id is None)
    tuples = embeddings_dict["caption_embedding_tuples"]
    sample = tuples[0][-1]
    assert sample.ndim == 1
    num_embeddings = len(tuples)
    new_dicts = model_id_to_label is None
    if new_dicts:
        model_id_to_label, label_to_model_id = {}, {}
    # one pass over the ids (first occurrence of a model id defines its gallery row, :49-56), then bulk copies: the
    # per-caption row assignments of the reference's loop are 2/3 of this function's time at the val size
    id_pos = 1 if dataset == "Primitives" else 2  # :45-46
    shapes, labels_shape = [], []
    labels = np.empty(num_embeddings, dtype=np.int64)
    for q, tup in enumerate(tuples):
        model_id = tup[id_pos]
        lab = model_id_to_label.get(model_id)
        if lab is None and new_dicts:
            lab = len(model_id_to_label)
            model_id_to_label[model_id] = lab
            label_to_model_id[lab] = model_id
            shapes.append(tup[4])
            labels_shape.append(lab)
        labels[q] = model_id_to_label[model_id] if lab is None else lab
    text = np.concatenate([tup[3] for tup in tuples]).reshape(num_embeddings, sample.shape[0])
    if _text_dtype is not None and text.dtype != _text_dtype:
        text = text.astype(_text_dtype)  # float64, as :25
    gallery = np.vstack(shapes)
    return (text, gallery, labels, np.array(labels_shape).astype(int), model_id_to_label, num_embeddings,
            label_to_model_id)


class SimilarityOrder:
    """Stand-in for the reference's `sort_indices` [Q, G] (eval_retrieval.py:82).

    Holds the device-resident fp32 similarity matrix. `ranks(labels)` gives the 1-based
    position of gallery item labels[q] in query q's ordering (what :184-186 and :217-219
    use sort_indices for) with the K4 kernel; `materialize()` / `[i]` produce explicit
    orderings (stable device sort) for callers that really want them.
    """

    def __init__(self, sim: torch.Tensor, n_gallery: int):
        self.sim, self.n_gallery = sim, n_gallery

    def __len__(self):
        return self.sim.shape[0]

    @property
    def shape(self):
        return (self.sim.shape[0], self.n_gallery)

    def ranks(self, labels) -> np.ndarray:
        lab = torch.as_tensor(np.asarray(labels), dtype=torch.int64, device=self.sim.device)
        _, _, _, nb = ops.topk_rank(self.sim, self.n_gallery, 1, lab)
        return nb.cpu().numpy().astype(np.int64) + 1

    def materialize(self, rows=None) -> np.ndarray:
        s = self.sim[:, : self.n_gallery] if rows is None else self.sim[rows, : self.n_gallery]
        return torch.sort(s, dim=-1, descending=True, stable=True).indices.cpu().numpy()

    def __getitem__(self, i):
        return self.materialize(i)

    def __array__(self, dtype=None, copy=None):
        a = self.materialize()
        return a if dtype is None else a.astype(dtype)


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("tricolo_b200.evaluation needs a CUDA device (sm_100a); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _flip_distances_like_reference(val: np.ndarray, block: Optional[int]) -> np.ndarray:
    """eval_retrieval.py:78 calls np.flip without an axis: rows come back in reverse query order
    (per 3000-query block when Q > 8000, :105-125). Reproduced so `distance` in nearest.jsonl matches."""
    if block is None:
        return val[::-1].copy()
    out = np.empty_like(val)
    for s in range(0, val.shape[0], block):
        out[s:s + block] = val[s:s + block][::-1]
    return out


def _drop_self(indices: np.ndarray, n_neighbors: int, block: Optional[int]) -> np.ndarray:
    """Self-retrieval post-pass of eval_retrieval.py:84-98 on the k+1 nearest neighbours: a row that contains its own
    index (block-relative row number + range_start, :87/:118) loses that entry, any other row keeps its first k."""
    n_q = indices.shape[0]
    final = np.zeros((n_q, n_neighbors), dtype=int)
    for s in range(0, n_q, block or n_q):
        blk = indices[s:s + (block or n_q)]
        own = np.arange(s, s + blk.shape[0]).reshape(-1, 1)
        has_self = np.equal(own, blk)
        for r in range(blk.shape[0]):
            if has_self[r].any():
                final[s + r] = np.delete(blk[r], np.nonzero(has_self[r])[0])[:n_neighbors]
            else:
                final[s + r] = blk[r, :n_neighbors]
    return final


def compute_nearest_neighbors(fit_embeddings_matrix, query_embeddings_matrix, n_neighbors):
    """(distances, indices [Q,k] i64, sort_indices) like eval_retrieval.py:133-146.  When fit == query (:139-140) the
    reference's self-retrieval branch applies: k+1 neighbours are taken, `distances` keeps k+1 columns (:76-78) and
    each row drops its own index (:84-98)."""
    fit = np.asarray(fit_embeddings_matrix)
    query = np.asarray(query_embeddings_matrix)
    fit_eq_query = fit.shape == query.shape and np.allclose(fit, query)
    k_eff = n_neighbors + 1 if fit_eq_query else n_neighbors
    dev = _device()
    q16 = ops.cast_16bit(torch.from_numpy(np.ascontiguousarray(query)).to(dev), OPERAND_FORMAT)
    g16 = ops.cast_16bit(torch.from_numpy(np.ascontiguousarray(fit)).to(dev), OPERAND_FORMAT)
    sim, n_g = ops.sim_gemm(q16, g16)
    dummy = torch.zeros((sim.shape[0],), dtype=torch.int64, device=dev)
    val, idx, _, _ = ops.topk_rank(sim, n_g, k_eff, dummy)
    n_q = query.shape[0]
    block = 3000 if n_q > 8000 else None
    distances = _flip_distances_like_reference(val.cpu().numpy().astype(np.float64), block)
    indices = idx.cpu().numpy().astype(np.int64)
    if fit_eq_query:
        indices = _drop_self(indices, n_neighbors, block)
    return distances, indices, SimilarityOrder(sim, n_g)


def metrics_from_ranks(indices: np.ndarray, rank: np.ndarray, labels: np.ndarray, n_neighbors: int,
                       fit_labels: Optional[np.ndarray] = None) -> dict:
    """Host fp64 finalise with the reference's NumPy op order (eval_retrieval.py:161-206)."""
    q = indices.shape[0]
    fit_labels = labels if fit_labels is None else np.asarray(fit_labels)
    if (np.asarray(indices) < 0).any():  # retrieve() marks the slots a gallery smaller than k cannot fill with -1
        raise ValueError("metrics: fewer gallery items than n_neighbors (the reference fails on this input too, "
                         "eval_retrieval.py:175)")
    rel_score = np.equal(fit_labels[indices], labels[:, None]).astype(np.float32)
    num_correct = np.cumsum(rel_score, axis=1, dtype=np.float32)
    num_relevant = np.bincount(fit_labels)[labels]
    rel_score_ideal = (np.arange(n_neighbors)[None, :] < np.minimum(num_relevant, n_neighbors)[:, None]).astype(np.float32)
    # sequential fp64 accumulation, as the Python loop at :184-187
    r_rank = float(np.cumsum(1.0 / rank.astype(np.float64))[-1]) / q
    dcg_d = np.log2(np.arange(1, n_neighbors + 1) + 1)
    dcg = np.cumsum((np.exp2(rel_score) - 1) / dcg_d, axis=1)
    dcg_ideal = np.cumsum((np.exp2(rel_score_ideal) - 1) / dcg_d, axis=1)
    ndcg = dcg / dcg_ideal
    return {
        "precision": np.sum(num_correct / np.arange(1, n_neighbors + 1), axis=0) / q,
        "recall": np.sum(num_correct / num_relevant[:, None], axis=0) / q,
        "recall_rate": np.sum(num_correct > 0, axis=0) / q,
        "ndcg": np.sum(ndcg, axis=0) / q,
        "mrr": r_rank,
    }


def compute_pr_at_k(indices, sort_indices, labels, n_neighbors, num_embeddings, fit_labels=None):
    """eval_retrieval.py:149-207. `sort_indices` may be a SimilarityOrder (ranks on the GPU) or an
    explicit [Q, G] array (ranks by search, as the reference)."""
    labels = np.asarray(labels)
    fl = labels if fit_labels is None else np.asarray(fit_labels)
    unique_gallery = fl.shape[0] == np.unique(fl).shape[0] and np.array_equal(fl, np.arange(fl.shape[0]))
    if isinstance(sort_indices, SimilarityOrder) and unique_gallery:
        rank = sort_indices.ranks(labels)
    else:
        order = np.asarray(sort_indices)
        rank = np.argmax(fl[order] == labels[:, None], axis=1).astype(np.int64) + 1
    return metrics_from_ranks(np.asarray(indices)[:num_embeddings], rank[:num_embeddings], labels[:num_embeddings],
                              n_neighbors, fl)


def get_nearest_info(indices, sort_indices, fit_labels, labels, label_to_model_id, caption_tuples):
    """eval_retrieval.py:210-246: model ids of queries and neighbours (+ reciprocal ranks)."""
    labels = np.asarray(labels)
    if isinstance(sort_indices, SimilarityOrder):
        rank = sort_indices.ranks(labels)
    else:
        rank = np.argmax(np.asarray(fit_labels)[np.asarray(sort_indices)] == labels[:, None], axis=1) + 1
    r_rank_list = [1 / int(r) for r in rank]
    query_model_ids = [t[2] for t in caption_tuples[: len(labels)]]
    cat_ids = [t[1] for t in caption_tuples[: len(labels)]]
    nearest_model_ids = [[label_to_model_id[int(c)] for c in row] for row in np.asarray(indices)]
    assert len(query_model_ids) == len(nearest_model_ids)
    return query_model_ids, cat_ids, nearest_model_ids, r_rank_list


def print_nearest_info(categories, query_model_ids, nearest_model_ids, distances, path="nearest.jsonl"):
    """eval_retrieval.py:281-304: one JSON line per query, in np.random.permutation order (the
    global NumPy RNG is consumed exactly as the reference does, :289)."""
    perm = np.random.permutation(len(nearest_model_ids))
    with open(path, "w") as f:
        for i in perm:
            f.write(json.dumps({"cat_id": categories[i], "groundtruth": query_model_ids[i] + ("-%04d" % i),
                                "retrieved_models": nearest_model_ids[i], "distance": distances[i].tolist()}) + "\n")


def _print_results(pr_at_k):
    print("\nRR@1 RR@5 NDCG@5 MRR")
    print(f'{round(pr_at_k["recall_rate"][0] * 100, 2)} {round(pr_at_k["recall_rate"][4] * 100, 2)} '
          f'{round(pr_at_k["ndcg"][4] * 100, 2)} {round(pr_at_k["mrr"] * 100, 2)}')


FUSED_BLOCK_QUERIES = 148 * 128 * 8  # fused path: only bounds the gathered-row scratch (1 KB per query)


def _fusable(dim: int, k: int) -> bool:
    return dim % 64 == 0 and 64 <= dim <= 512 and 1 <= k <= 16


def retrieve(text: torch.Tensor, gallery: torch.Tensor, labels: torch.Tensor, k: int = 5,
             block_queries: int = None, operand_format: int = OPERAND_FORMAT, fused: bool = None):
    """Tensor-in fast entry: device-resident text [Q,D], gallery [G,D], labels [Q] (int64 gallery index).

    Returns (topk_val [Q,k] f32, topk_idx [Q,k] i32, rank [Q] i32) on the device.

    fused=True  (default when dim % 64 == 0 and dim <= 512): K2'+K4 fused kernel, the similarities stay in
                TMEM/registers (tensor-bound).
    fused=False : similarity GEMM (K2') into an fp32 block of block_queries x G, then the top-k/rank kernel
                (K4, HBM-bound).  Both give identical results bit for bit."""
    dt = ops.L.op_torch_dtype(operand_format)
    g16 = gallery if gallery.dtype == dt else ops.cast_16bit(gallery, operand_format)
    n_q, n_g, dim = text.shape[0], gallery.shape[0], text.shape[1]
    dev = text.device
    if fused is None:
        fused = _fusable(dim, k)
    elif fused and not _fusable(dim, k):
        raise ValueError(f"fused retrieval needs dim % 64 == 0, dim <= 512, k <= 16 (got dim={dim}, k={k})")
    if block_queries is None:
        block_queries = FUSED_BLOCK_QUERIES if fused else two_kernel_block_queries(n_g)
    val = torch.empty((n_q, k), dtype=torch.float32, device=dev)
    idx = torch.empty((n_q, k), dtype=torch.int32, device=dev)
    rank = torch.empty((n_q,), dtype=torch.int32, device=dev)
    labels = labels.to(torch.int64)
    if not fused:
        ld = (n_g + 31) // 32 * 32
        buf = torch.empty((min(block_queries, n_q), ld), dtype=torch.float32, device=dev)
    for s in range(0, n_q, block_queries):
        e = min(s + block_queries, n_q)
        tq = text[s:e]
        q16 = tq if tq.dtype == dt else ops.cast_16bit(tq, operand_format)
        if fused:
            gt = ops.gt_sim_mma(q16, g16, labels[s:e])
            v, i, nb = ops.sim_topk_fused(q16, g16, k, labels[s:e], gt)
        else:
            sim, _ = ops.sim_gemm(q16, g16, out=buf[: e - s])
            v, i, _, nb = ops.topk_rank(sim, n_g, k, labels[s:e])
        val[s:e], idx[s:e] = v, i
        rank[s:e] = nb + 1
    return val, idx, rank


def metrics_from_rank_counts(counts, sum_rr: float, n_queries: int) -> dict:
    """The metric dict from #{rank == j+1} (j < k) and sum 1/rank: the closed form of eval_retrieval.py:161-206 when
    every query has exactly one relevant gallery item (equal to metrics_from_ranks to ~1e-15)."""
    counts = np.asarray(counts, dtype=np.float64)
    k = counts.shape[0]
    hit = np.cumsum(counts)  # #{rank <= j+1}
    rr = hit / n_queries
    return {"precision": rr / np.arange(1, k + 1), "recall": rr.copy(), "recall_rate": rr,
            "ndcg": np.cumsum(counts / np.log2(np.arange(1, k + 1) + 1)) / n_queries, "mrr": float(sum_rr) / n_queries}


def retrieve_metrics(text: torch.Tensor, gallery: torch.Tensor, labels: torch.Tensor, k: int = 5, rank=None, **kw) -> dict:
    """Tensor-in entry all the way to the metric dict: retrieve() (or ranks already computed, e.g. by
    distributed.sharded_retrieve) + the K5 reduction on the device; only k+1 numbers cross PCIe.  Valid when labels
    index a de-duplicated gallery (one relevant item per query), which is what construct_embeddings_matrix builds."""
    if rank is None:
        _, _, rank = retrieve(text, gallery, labels, k, **kw)
    red = ops.rank_metrics(rank.contiguous(), k).cpu().numpy()
    return metrics_from_rank_counts(red[:k], red[k], rank.numel())


def compute_metrics(dataset, embeddings_dict, print_results=False, write_nearest=True):
    """Drop-in for eval_retrieval.py:249-278: same input format, same returned dict."""
    (text, gallery, labels, fit_labels, _model_id_to_label, num_embeddings,
     label_to_model_id) = construct_embeddings_matrix(dataset, embeddings_dict, _text_dtype=None)
    n_neighbors = 5  # :257
    dev = _device()
    t_dev = torch.from_numpy(text).to(dev)
    g_dev = torch.from_numpy(np.ascontiguousarray(gallery)).to(dev)
    l_dev = torch.from_numpy(labels).to(dev)
    val, idx, rank = retrieve(t_dev, g_dev, l_dev, n_neighbors)
    indices = idx.cpu().numpy().astype(np.int64)
    rank_h = rank.cpu().numpy().astype(np.int64)
    pr_at_k = metrics_from_ranks(indices, rank_h, labels, n_neighbors, fit_labels)
    if write_nearest:
        distances = _flip_distances_like_reference(val.cpu().numpy().astype(np.float64),
                                                   3000 if num_embeddings > 8000 else None)
        tuples = embeddings_dict["caption_embedding_tuples"]
        nearest_model_ids = [[label_to_model_id[int(c)] for c in row] for row in indices]
        print_nearest_info([t[1] for t in tuples], [t[2] for t in tuples], nearest_model_ids, distances)
    if print_results:
        _print_results(pr_at_k)
    return pr_at_k
