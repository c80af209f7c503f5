"""Oracle for text->shape retrieval and its metrics — test infrastructure, not product code.

Restates tricolo/evaluation/eval_retrieval.py:
  construct_embeddings_matrix (:6-65), _compute_nearest_neighbors_cosine (:68-99),
  compute_pr_at_k (:149-207), compute_metrics (:249-278).
Tie order: the reference leaves it to np.argsort's introsort (undefined); the
oracle uses the build's stated order — similarity descending, gallery index
ascending (a stable sort of the negated similarities).
"""
from __future__ import annotations

import numpy as np


def build_matrices(tuples):
    """(:6-65) text matrix [Q,D] float64, gallery = shape vector of the FIRST occurrence of
    each model_id in first-seen order, labels[q] = gallery index of the query's shape."""
    dim = tuples[0][-1].shape[0]
    text = np.zeros((len(tuples), dim))  # float64, as :25
    labels = np.zeros(len(tuples), dtype=np.int64)
    first = {}
    gallery = []
    for q, (_cap, _cat, model_id, t, s) in enumerate(tuples):
        if model_id not in first:
            first[model_id] = len(gallery)
            gallery.append(s)
        text[q] = t
        labels[q] = first[model_id]
    id_of = {v: k for k, v in first.items()}
    return text, np.vstack(gallery), labels, np.arange(len(gallery)), first, id_of


def similarities(text, gallery):
    """(:74) raw dot product — NOT cosine — in the promoted dtype (float64 for the reference's inputs)."""
    return np.dot(text, gallery.T)


def rank_order(sim):
    """Full ordering per query under (similarity desc, index asc)."""
    return np.argsort(-sim, axis=1, kind="stable")


def topk_and_rank(sim, labels, k):
    """top-k indices/values and the 1-based rank of the ground-truth column, without a full sort.

    rank = 1 + #{j : s_j > s_gt} + #{j < gt : s_j == s_gt}   (tied inputs)
    """
    q = sim.shape[0]
    order = rank_order(sim)
    idx = order[:, :k]
    val = np.take_along_axis(sim, idx, axis=1)
    s_gt = sim[np.arange(q), labels][:, None]
    cols = np.arange(sim.shape[1])[None, :]
    n_before = ((sim > s_gt) | ((sim == s_gt) & (cols < labels[:, None]))).sum(axis=1)
    return val, idx, n_before + 1


def metrics_from_topk(indices, rank, labels, k, fit_labels=None):
    """(:149-207) precision / recall / recall_rate / ndcg @1..k and MRR, same NumPy op order as
    the reference for the final reductions (float32 work arrays, float64 sums)."""
    q = indices.shape[0]
    if fit_labels is None:
        fit_labels = labels
    fit_labels = np.asarray(fit_labels)
    nearest_classes = fit_labels[indices]                       # :172
    rel = np.equal(nearest_classes, labels[:, None]).astype(np.float32)  # :175
    num_correct = np.cumsum(rel, axis=1, dtype=np.float32)      # :178-181 (exact: small integers)
    num_relevant = np.bincount(fit_labels)[labels]              # :163-164
    rel_ideal = np.zeros((q, k), dtype=np.float32)
    clamp = np.minimum(num_relevant, k)
    rel_ideal[np.arange(k)[None, :] < clamp[:, None]] = 1       # :176
    mrr = 0.0
    for r in rank:                                              # :184-187 sequential fp64 sum
        mrr += 1 / int(r)
    mrr = mrr / q
    dcg_n = np.exp2(rel) - 1                                    # :190
    dcg_d = np.log2(np.arange(1, k + 1) + 1)                    # :191
    dcg = np.cumsum(dcg_n / dcg_d, axis=1)                      # :192
    dcg_ideal = np.cumsum((np.exp2(rel_ideal) - 1) / dcg_d, axis=1)  # :194-195
    ndcg = dcg / dcg_ideal                                      # :197
    return {
        "precision": np.sum(num_correct / np.arange(1, k + 1), axis=0) / q,     # :201
        "recall": np.sum(num_correct / num_relevant[:, None], axis=0) / q,      # :200
        "recall_rate": np.sum(num_correct > 0, axis=0) / q,                     # :199
        "ndcg": np.sum(ndcg, axis=0) / q,                                       # :198
        "mrr": mrr,
    }


def compute_metrics(tuples, k: int = 5, sim=None):
    """(:249-278) end to end. `sim` lets a test substitute the GPU's own fp32 similarity
    matrix so that the selection stage can be checked bit-exactly on identical numbers."""
    text, gallery, labels, fit_labels, _, _ = build_matrices(tuples)
    if sim is None:
        sim = similarities(text, gallery)
    val, idx, rank = topk_and_rank(sim, labels, k)
    out = metrics_from_topk(idx, rank, labels, k, fit_labels)
    out["_indices"] = idx
    out["_values"] = val
    out["_rank"] = rank
    return out


# ---------------------------------------------------------------------------
# synthetic data generators shared by tests and bench
# ---------------------------------------------------------------------------
def bf16_round(a: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even to bfloat16 precision, returned as float32."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def make_val_shaped(seed=0, n_shapes=1486, n_queries=7424, dim=512, noise=10.5, round_bf16=False,
                    trimodal_gallery=False):
    """C3 generator: unit-norm Gaussian shapes, every shape owns >= 1 caption."""
    rng = np.random.default_rng(seed)
    shape = rng.standard_normal((n_shapes, dim)).astype(np.float32)
    shape /= np.linalg.norm(shape, axis=1, keepdims=True)
    if trimodal_gallery:  # gallery = image + voxel, un-normalised sum (tricolo_net.py:135-139)
        other = rng.standard_normal((n_shapes, dim)).astype(np.float32)
        other /= np.linalg.norm(other, axis=1, keepdims=True)
        gal = (shape + other).astype(np.float32)
    else:
        gal = shape
    owner = rng.permutation(np.concatenate([np.arange(n_shapes), rng.integers(0, n_shapes, n_queries - n_shapes)]))
    tuples = []
    for s in owner:
        te = shape[s] + noise * rng.standard_normal(dim).astype(np.float32) / np.sqrt(dim)
        te /= np.linalg.norm(te)
        te = te.astype(np.float32)
        sv = gal[s]
        if round_bf16:
            te, sv = bf16_round(te), bf16_round(sv)
        tuples.append((None, "c", f"m{s}", te, sv))
    return tuples


def make_large_retrieval(seed=0, n_shapes=200_000, n_queries=3000, dim=512, noise=10.5):
    """C5-shaped block (the C3 recipe with owners drawn uniformly): unit-norm Gaussian gallery of
    n_shapes, n_queries captions = normalise(shape[owner] + noise * N(0, I) / sqrt(dim)); everything bf16-exact.
    Returns (text [Q,D] f32, gallery [G,D] f32, labels [Q] i64) — the tensor-in form (a 200k-entry tuple list would
    itself be the bottleneck)."""
    rng = np.random.default_rng(seed)
    shape = rng.standard_normal((n_shapes, dim), dtype=np.float32)
    shape /= np.linalg.norm(shape, axis=1, keepdims=True)
    owner = rng.integers(0, n_shapes, n_queries)
    text = shape[owner] + np.float32(noise) * rng.standard_normal((n_queries, dim), dtype=np.float32) / np.float32(np.sqrt(dim))
    text /= np.linalg.norm(text, axis=1, keepdims=True)
    return bf16_round(text), bf16_round(shape), owner.astype(np.int64)


def make_integer_kat(seed=42, n_shapes=50, dim=16, captions=3):
    """KAT-E1: integer-valued vectors -> every dot product is exact in any precision."""
    rng = np.random.default_rng(seed)
    shape = rng.integers(-1, 2, (n_shapes, dim)).astype(np.float32)
    tuples = []
    for s in range(n_shapes):
        for _ in range(captions):
            te = shape[s].copy()
            flip = rng.integers(0, dim, 4)
            te[flip] = rng.integers(-1, 2, 4)
            tuples.append((None, "c", f"m{s}", te, shape[s]))
    return tuples


def make_val_batches(seed=11, n_shapes=300, n_items=2000, dim=128, batch=192, noise=14.0):
    """Validation-epoch shaped input of tricolo_net.py:78-88: a list of (data_dict, output_dict) batches with
    text / image / voxel features (float32 numpy), model ids with repeats and categories.  image and voxel lie on a
    1/64 grid so that image + voxel (tricolo_net.py:135-139) is exact in bf16 as well as in fp32; text is bf16-exact."""
    rng = np.random.default_rng(seed)
    img = np.clip(np.round(rng.standard_normal((n_shapes, dim)) * 16) / 64, -1.5, 1.5).astype(np.float32)
    vox = np.clip(np.round(rng.standard_normal((n_shapes, dim)) * 16) / 64, -1.5, 1.5).astype(np.float32)
    shape = img + vox
    owner = rng.permutation(np.concatenate([np.arange(n_shapes), rng.integers(0, n_shapes, n_items - n_shapes)]))
    batches = []
    for s in range(0, n_items, batch):
        own = owner[s:s + batch]
        te = shape[own] + noise * rng.standard_normal((len(own), dim)).astype(np.float32) / np.sqrt(dim)
        te = bf16_round(te.astype(np.float32))
        data = {"model_id": [f"m{int(o):04d}" for o in own], "category": [f"c{int(o) % 3}" for o in own]}
        out = {"text_features": te, "image_features": img[own].copy(), "voxel_features": vox[own].copy()}
        batches.append((data, out))
    return batches


def make_self_retrieval(seed=5, n=500, dim=64):
    """fit == query input of the self-retrieval branch (eval_retrieval.py:84-98), bf16-exact."""
    rng = np.random.default_rng(seed)
    return bf16_round(rng.standard_normal((n, dim)).astype(np.float32))


def topk_margin(sim, k):
    """Per row: the smallest gap between consecutive values among the k+1 largest (fp64); rows whose margin
    exceeds the fp32 rounding of the GPU's similarities must reproduce the reference's order exactly."""
    part = -np.sort(-sim, axis=1)[:, :k + 1]
    return np.min(-np.diff(part, axis=1), axis=1)
