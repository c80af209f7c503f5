"""CPU: the oracle (oracle/) against the golden vectors produced by the reference itself."""
import os

import numpy as np
import pytest

from oracle import ntxent_oracle as NO
from oracle import retrieval_oracle as RO
from tests.cases import GRAD_CASES, LOSS_CASES, bf16_rounded, loss_case

HERE = os.path.dirname(os.path.abspath(__file__))
TAU, ALPHA = 0.1, 0.25


@pytest.mark.parametrize("name", LOSS_CASES)
@pytest.mark.parametrize("variant", ["fp32", "bf16"])
def test_loss_matches_reference(golden, name, variant):
    feats = loss_case(name)
    if variant == "bf16":
        feats = bf16_rounded(feats)
    losses, grads = NO.trimodal_forward_backward({k: v.numpy() for k, v in feats.items()}, TAU, ALPHA)
    ref = golden["cases"][f"{name}.{variant}"]
    assert set(losses) == set(ref["losses"])
    for k, v in ref["losses"].items():
        # the reference is fp32; the oracle fp64
        assert losses[k] == pytest.approx(v, rel=2e-6, abs=2e-7), k
    for k, gn in ref["grad_norm"].items():
        assert np.linalg.norm(grads[k]) == pytest.approx(gn, rel=1e-4)
        sample = grads[k].flatten()[::997][:64]
        np.testing.assert_allclose(sample, np.array(ref["grad_sample"][k]), rtol=2e-3, atol=1e-7 * max(1.0, gn))


@pytest.mark.parametrize("name", GRAD_CASES)
def test_full_gradients_match_reference(name):
    z = np.load(os.path.join(HERE, "golden", "loss_grads.npz"))
    feats = bf16_rounded(loss_case(name))
    _, grads = NO.trimodal_forward_backward({k: v.numpy() for k, v in feats.items()}, TAU, ALPHA)
    for k, g in grads.items():
        ref = z[f"{name}.bf16.{k}"]
        got = g[::4]
        scale = np.abs(ref).max()
        # fp32 autograd of the reference vs fp64 closed form
        assert np.abs(got - ref).max() <= 2e-5 * scale + 1e-12
        assert np.linalg.norm(got - ref) <= 1e-5 * np.linalg.norm(ref) + 1e-12


def test_argument_order_matters(golden):
    import torch

    g = torch.Generator().manual_seed(3)
    a = torch.randn(64, 512, generator=g).numpy()
    b = torch.randn(64, 512, generator=g).numpy()
    assert NO.ntxent_forward(a, b, TAU, ALPHA) == pytest.approx(golden["order"]["ab"], rel=2e-6)
    assert NO.ntxent_forward(b, a, TAU, ALPHA) == pytest.approx(golden["order"]["ba"], rel=2e-6)
    assert golden["order"]["ab"] != golden["order"]["ba"]


def test_known_answers():
    e = np.eye(128, 512)
    assert NO.ntxent_forward(e, e, TAU, ALPHA) == pytest.approx(np.log(1 + 127 * np.exp(-10)), rel=1e-12)
    o = np.ones((128, 512))
    assert NO.ntxent_forward(o, o, TAU, ALPHA) == pytest.approx(np.log(128), rel=1e-12)


def test_closed_form_gradient_against_autograd():
    import torch

    g = torch.Generator().manual_seed(11)
    a = torch.randn(48, 64, generator=g, dtype=torch.float64)
    b = torch.randn(48, 64, generator=g, dtype=torch.float64)
    a[3] = 0  # clamped-norm row
    ta, tb = a.clone().requires_grad_(), b.clone().requires_grad_()
    NO.torch_ntxent(ta, tb, TAU, ALPHA).backward()
    loss, ga, gb = NO.ntxent_forward_backward(a.numpy(), b.numpy(), TAU, ALPHA)
    np.testing.assert_allclose(ga, ta.grad.numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gb, tb.grad.numpy(), rtol=1e-9, atol=1e-12)


# ---------------------------------------------------------------- retrieval
def _cmp_metrics(got, ref):
    for k in ("precision", "recall", "recall_rate", "ndcg"):
        assert np.array_equal(np.asarray(got[k]), np.asarray(ref[k])), k  # bit-identical fp64
    assert got["mrr"] == ref["mrr"]


def test_eval_integer_kat_brackets_reference(golden):
    # Integer data: every dot product is exact, but the rows contain real ties and the
    # reference's tie order is undefined (introsort).  Its metrics must lie between the
    # oracle's with every tie resolved against / in favour of the ground truth, and the
    # stated order (lowest index first) must lie in the same bracket.
    tuples = RO.make_integer_kat()
    text, gal, labels, fit_labels, _, _ = RO.build_matrices(tuples)
    sim = RO.similarities(text, gal)
    s_gt = sim[np.arange(len(labels)), labels][:, None]
    best = 1 + (sim > s_gt).sum(axis=1)
    worst = (sim >= s_gt).sum(axis=1)
    ref = golden["eval"]["KAT_E1"]
    stated = RO.compute_metrics(tuples)
    assert np.all(best <= stated["_rank"]) and np.all(stated["_rank"] <= worst)
    for k in range(5):
        lo, hi = np.mean(worst <= k + 1), np.mean(best <= k + 1)
        assert lo <= ref["recall_rate"][k] <= hi
        assert lo <= stated["recall_rate"][k] <= hi
    assert np.mean(1 / worst) <= ref["mrr"] <= np.mean(1 / best)
    assert np.mean(1 / worst) <= stated["mrr"] <= np.mean(1 / best)


@pytest.mark.parametrize("key,kw", [
    ("C3.fp32", dict(round_bf16=False)),
    ("C3.bf16", dict(round_bf16=True)),
    ("C3TRI.bf16", dict(round_bf16=True, trimodal_gallery=True)),
    ("SMALL.bf16", dict(seed=3, n_shapes=300, n_queries=1000, dim=128, round_bf16=True)),
])
def test_eval_val_shaped_bit_exact(golden, key, kw):
    got = RO.compute_metrics(RO.make_val_shaped(**kw))
    _cmp_metrics(got, golden["eval"][key])


def test_eval_indices_and_ranks_match_reference():
    z = np.load(os.path.join(HERE, "golden", "eval_c3_bf16.npz"))
    got = RO.compute_metrics(RO.make_val_shaped(round_bf16=True))
    assert np.array_equal(got["_indices"], z["indices"])
    assert np.array_equal(got["_rank"], z["rank"])


def test_tie_break_is_lowest_index():
    sim = np.array([[1.0, 3.0, 3.0, 2.0, 3.0, 0.0]])
    val, idx, rank = RO.topk_and_rank(sim, np.array([4]), 5)
    assert idx.tolist() == [[1, 2, 4, 3, 0]]
    assert rank.tolist() == [3]
    sim = np.zeros((1, 10))
    _, idx, rank = RO.topk_and_rank(sim, np.array([7]), 5)
    assert idx.tolist() == [[0, 1, 2, 3, 4]] and rank.tolist() == [8]


def test_bf16_round_matches_torch():
    import torch

    x = np.random.default_rng(0).standard_normal(10000).astype(np.float32)
    assert np.array_equal(RO.bf16_round(x), torch.from_numpy(x).bfloat16().float().numpy())


# ---------------------------------------------------------------- norm=False (nt_xent.py:55)
@pytest.mark.parametrize("name", ["R1", "R2", "R3", "R4"])
def test_unnormalised_oracle_matches_reference(name):
    """oracle(norm=False) against the reference's own forward/backward (tests/golden/make_golden_raw.py)."""
    import json

    from tests.cases import raw_case

    with open(os.path.join(HERE, "golden", "raw_outputs.json")) as f:
        ref = json.load(f)["cases"][name]
    z = np.load(os.path.join(HERE, "golden", "raw_grads.npz"))
    feats = {k: v.numpy() for k, v in raw_case(name).items()}
    losses, grads = NO.trimodal_forward_backward(feats, TAU, ALPHA, prefix="x", norm=False)
    for k, v in ref["losses"].items():
        assert losses[f"x/{k}"] == pytest.approx(v, rel=5e-6, abs=1e-6), k  # fp32 reference vs fp64 oracle
    assert losses["x/total_loss"] == pytest.approx(ref["total"], rel=5e-6, abs=1e-6)
    for k, g in grads.items():
        r = z[f"{name}.{k}"]
        assert np.linalg.norm(g[::4] - r) <= 2e-5 * np.linalg.norm(r) + 1e-9, k
