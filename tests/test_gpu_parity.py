"""GPU parity tests (run with -m gpu on a B200): the sm_100a path, called through the
C ABI, against the CPU oracle and the golden vectors produced by the reference itself.

Tolerances (BASELINE.json north_star): loss and gradients rtol 1e-3 on bf16-rounded
inputs; top-k indices, ranks and RR/NDCG bit-exact on the fp32 similarities under the
(similarity desc, index asc) order.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ntxent_oracle as NO
from oracle import retrieval_oracle as RO
from tests.cases import GRAD_CASES, LOSS_CASES, bf16_rounded, loss_case

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
TAU, ALPHA = 0.1, 0.25
RTOL = 1e-3  # north_star: "Loss and gradients must agree within rtol 1e-3"


@pytest.fixture(scope="module")
def tb():
    import tricolo_b200  # noqa: F401  (raises if the CUDA library is missing)
    from tricolo_b200 import ops
    from tricolo_b200 import loss as L
    from tricolo_b200 import evaluation as E

    class NS:
        pass

    ns = NS()
    ns.ops, ns.loss, ns.eval = ops, L, E
    return ns


def _grad_close(got: torch.Tensor, ref: np.ndarray, scale_hint: float):
    """Gradient check: norm-wise and element-wise (against the largest reference entry) rtol 1e-3.
    `scale_hint` is the natural magnitude of a gradient entry; used as the absolute floor when the
    reference gradient vanishes analytically (e.g. identical rows)."""
    got = got.double().cpu().numpy()
    ref_norm = np.linalg.norm(ref)
    floor = scale_hint * np.sqrt(ref.size)
    assert np.linalg.norm(got - ref) <= RTOL * max(ref_norm, floor), (np.linalg.norm(got - ref), ref_norm)
    assert np.abs(got - ref).max() <= 2 * RTOL * max(np.abs(ref).max(), scale_hint)


@pytest.mark.parametrize("name", LOSS_CASES)
def test_trimodal_loss_and_grads_vs_oracle(tb, name):
    feats = bf16_rounded(loss_case(name))
    dev = {k: v.cuda().requires_grad_(True) for k, v in feats.items()}
    fn = tb.loss.NTXentLoss(TAU, ALPHA)
    losses = tb.loss.calculate_losses(dev, "train_loss", fn)
    losses["train_loss/total_loss"].backward()
    ref_l, ref_g = NO.trimodal_forward_backward({k: v.numpy() for k, v in feats.items()}, TAU, ALPHA)
    assert list(losses.keys()) == list(ref_l.keys())
    for k, v in ref_l.items():
        assert losses[k].dtype == torch.float32 and losses[k].dim() == 0
        assert float(losses[k].detach()) == pytest.approx(v, rel=RTOL)
    b = next(iter(feats.values())).shape[0]
    for k, v in dev.items():
        norms = feats[k].norm(dim=1)
        hint = 1.0 / (b * TAU * float(norms[norms > 0].median()) * np.sqrt(feats[k].shape[1])) * 1e-3
        _grad_close(v.grad, ref_g[k], hint)


@pytest.mark.parametrize("name", LOSS_CASES)
def test_loss_vs_reference_golden(tb, golden, name):
    """Directly against the numbers the reference code produced (tests/golden/make_golden.py)."""
    feats = bf16_rounded(loss_case(name))
    dev = {k: v.cuda() for k, v in feats.items()}
    fn = tb.loss.NTXentLoss(TAU, ALPHA)
    with torch.no_grad():
        losses = tb.loss.calculate_losses(dev, "train_loss", fn)
    for k, v in golden["cases"][f"{name}.bf16"]["losses"].items():
        assert float(losses[k]) == pytest.approx(v, rel=RTOL)


@pytest.mark.parametrize("name", GRAD_CASES)
def test_grads_vs_reference_golden(tb, name):
    z = np.load(os.path.join(HERE, "golden", "loss_grads.npz"))
    feats = bf16_rounded(loss_case(name))
    dev = {k: v.cuda().requires_grad_(True) for k, v in feats.items()}
    fn = tb.loss.NTXentLoss(TAU, ALPHA)
    tb.loss.calculate_losses(dev, "train_loss", fn)["train_loss/total_loss"].backward()
    for k, v in dev.items():
        ref = z[f"{name}.bf16.{k}"].astype(np.float64)
        got = v.grad[::4]
        if name == "KAT3":
            # row 0 of the text matrix is zero: its gradient is g / eps ~ 1e8 and dominates every norm;
            # check it separately from the regular rows
            if k == "text_features":
                _grad_close(got[:1], ref[:1], 0.0)
                got, ref = got[1:], ref[1:]
        _grad_close(got, ref, 0.0)


def test_bimodal_module_signature_and_order(tb, golden):
    g = torch.Generator().manual_seed(3)
    a = torch.randn(64, 512, generator=g).cuda()
    b = torch.randn(64, 512, generator=g).cuda()
    fn = tb.loss.NTXentLoss(temperature=TAU, alpha_weight=ALPHA)
    assert len(list(fn.parameters())) == 0 and len(fn.state_dict()) == 0
    ab, ba = float(fn(a, b)), float(fn(b, a))
    assert ab == pytest.approx(golden["order"]["ab"], rel=RTOL)
    assert ba == pytest.approx(golden["order"]["ba"], rel=RTOL)
    assert ab != ba  # alpha != 0.5: argument order matters
    # norm=False is a different function of the inputs (nt_xent.py:55), same signature
    assert float(fn(a, b, norm=False)) == pytest.approx(NO.ntxent_forward_backward(a.cpu().numpy(), b.cpu().numpy(), TAU, ALPHA,
                                                                               norm=False)[0], rel=1e-5)
    with pytest.raises(RuntimeError):
        fn(a.cpu(), b.cpu())  # no CPU path


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_input_dtypes(tb, dtype):
    g = torch.Generator().manual_seed(21)
    a32 = torch.randn(256, 512, generator=g).bfloat16().float()
    b32 = torch.randn(256, 512, generator=g).bfloat16().float()
    a = a32.to(dtype).cuda().requires_grad_(True)
    b = b32.to(dtype).cuda().requires_grad_(True)
    fn = tb.loss.NTXentLoss(TAU, ALPHA)
    loss = fn(a, b)
    loss.backward()
    ref, ga, gb = NO.ntxent_forward_backward(a32.numpy(), b32.numpy(), TAU, ALPHA)
    assert float(loss.detach()) == pytest.approx(ref, rel=RTOL)
    assert a.grad.dtype == dtype
    tol = RTOL if dtype == torch.float32 else 6e-3  # the gradient itself is rounded to the 16-bit input dtype
    assert np.linalg.norm(a.grad.double().cpu().numpy() - ga) <= tol * np.linalg.norm(ga)
    assert np.linalg.norm(b.grad.double().cpu().numpy() - gb) <= tol * np.linalg.norm(gb)


def test_upstream_gradient_scaling_and_partial_requires_grad(tb):
    g = torch.Generator().manual_seed(8)
    f = [torch.randn(256, 512, generator=g).bfloat16().float() for _ in range(3)]
    t = f[0].cuda().requires_grad_(True)
    i = f[1].cuda()  # no gradient wanted
    v = f[2].cuda().requires_grad_(True)
    losses = tb.loss.trimodal_ntxent([t, i, v], TAU, ALPHA)
    w = torch.tensor([2.0, -0.5, 4096.0], device="cuda")
    (losses * w).sum().backward()
    assert i.grad is None
    gt = np.zeros_like(f[0].numpy(), dtype=np.float64)
    gv = np.zeros_like(gt)
    for p, (a, b) in enumerate([(0, 1), (0, 2), (1, 2)]):
        _, ga, gb = NO.ntxent_forward_backward(f[a].numpy(), f[b].numpy(), TAU, ALPHA, grad_out=float(w[p]))
        if a == 0:
            gt += ga
        if b == 2:
            gv += gb
    assert np.linalg.norm(t.grad.double().cpu().numpy() - gt) <= RTOL * np.linalg.norm(gt)
    assert np.linalg.norm(v.grad.double().cpu().numpy() - gv) <= RTOL * np.linalg.norm(gv)


def test_large_batch_properties(tb):
    """B = 4096 (oracle too slow in fp64 loops? no — 4096^2 is fine) plus size-independent properties:
    gradient rows are orthogonal to the inputs (normalise backward) and permutation equivariance."""
    g = torch.Generator().manual_seed(13)
    base = torch.randn(4096, 512, generator=g)
    f = [(base + 0.5 * torch.randn(4096, 512, generator=g)).bfloat16().float() for _ in range(3)]
    dev = [x.cuda().requires_grad_(True) for x in f]
    losses = tb.loss.trimodal_ntxent(dev, TAU, ALPHA)
    losses.sum().backward()
    ref_l, ref_g = NO.trimodal_forward_backward(
        {"text_features": f[0].numpy(), "image_features": f[1].numpy(), "voxel_features": f[2].numpy()}, TAU, ALPHA)
    for p, k in enumerate(["train_loss/text_image_loss", "train_loss/text_voxel_loss", "train_loss/image_voxel_loss"]):
        assert float(losses[p].detach()) == pytest.approx(ref_l[k], rel=RTOL)
    for x, k in zip(dev, ["text_features", "image_features", "voxel_features"]):
        assert np.linalg.norm(x.grad.double().cpu().numpy() - ref_g[k]) <= RTOL * np.linalg.norm(ref_g[k])
        # dx_i . x_i = 0 for every row
        dots = (x.grad.double() * x.detach().double()).sum(1).abs().max().item()
        assert dots <= 1e-5 * x.grad.double().norm(dim=1).max().item() * x.detach().double().norm(dim=1).max().item()
    # permuting the batch permutes gradients and leaves the losses unchanged (up to summation order)
    perm = torch.randperm(4096, generator=g)
    dev_p = [x.detach()[perm.cuda()].clone().requires_grad_(True) for x in dev]
    losses_p = tb.loss.trimodal_ntxent(dev_p, TAU, ALPHA)
    losses_p.sum().backward()
    assert torch.allclose(losses_p.detach(), losses.detach(), rtol=1e-5)
    for x, xp in zip(dev, dev_p):
        assert torch.allclose(xp.grad, x.grad[perm.cuda()], rtol=1e-3, atol=1e-3 * x.grad.abs().max().item())


def test_bf16_operand_mode_runs_and_is_less_accurate(tb):
    """bf16 tensor-core operands are supported (north_star wording) but cannot meet rtol 1e-3 on gradients;
    the default is fp16 operands. This test documents the measured gap."""
    feats = bf16_rounded(loss_case("G2"))
    ref_l, ref_g = NO.trimodal_forward_backward({k: v.numpy() for k, v in feats.items()}, TAU, ALPHA)
    errs = {}
    for op in (tb.ops.F16, tb.ops.BF16):
        dev = [v.cuda().requires_grad_(True) for v in feats.values()]
        losses = tb.loss.trimodal_ntxent(dev, TAU, ALPHA, op_format=op)
        losses.sum().backward()
        assert float(losses.sum().detach()) == pytest.approx(ref_l["train_loss/total_loss"], rel=RTOL)
        errs[op] = max(np.linalg.norm(x.grad.double().cpu().numpy() - ref_g[k]) / np.linalg.norm(ref_g[k])
                       for x, k in zip(dev, feats.keys()))
    assert errs[tb.ops.F16] <= RTOL
    assert errs[tb.ops.BF16] <= 1e-2
    assert errs[tb.ops.BF16] > errs[tb.ops.F16]


@pytest.mark.parametrize("name", ["R1", "R2", "R3", "R4"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_unnormalised_loss_and_grads(tb, name, dtype):
    """norm=False (nt_xent.py:55; csrc/ntxent_raw.cu) against the fp64 oracle AND the reference's own outputs
    (tests/golden/raw_*.{json,npz}): logits of magnitude 1e1..1e3, online (max, sum) statistics, ragged tiles."""
    import json

    from tests.cases import raw_case

    feats = raw_case(name)  # bf16-representable values: both input dtypes carry them exactly
    ref_losses, ref_grads = NO.trimodal_forward_backward({k: v.numpy() for k, v in feats.items()}, TAU, ALPHA, prefix="x",
                                                        norm=False)
    with open(os.path.join(HERE, "golden", "raw_outputs.json")) as f:
        gold = json.load(f)["cases"][name]
    gz = np.load(os.path.join(HERE, "golden", "raw_grads.npz"))
    dev = [v.to(dtype).cuda().requires_grad_(True) for v in feats.values()]
    fn = tb.loss.NTXentLoss(TAU, ALPHA)
    if len(dev) == 2:
        total = fn(dev[0], dev[1], norm=False)  # the reference signature
        losses = total.reshape(1)
    else:
        losses, total = fn.fused_total(dev, norm=False)
    from itertools import combinations
    names = [f"{a[:-9]}_{b[:-9]}_loss" for a, b in combinations(feats.keys(), 2)]  # pair order of tricolo_net.py:59-61
    for got, k in zip(losses.tolist(), names):
        assert got == pytest.approx(gold["losses"][k], rel=RTOL), k
        assert got == pytest.approx(ref_losses[f"x/{k}"], rel=1e-5, abs=1e-6), k  # fp32 arithmetic: far inside rtol 1e-3
    assert float(total.detach()) == pytest.approx(gold["total"], rel=RTOL)
    total.backward()
    # bf16 gradients are rounded on output: half an ulp = 2^-9 relative per entry
    # fp32 logits of magnitude 1e3 (R3, R4) carry absolute errors of ~1e-4, in this library and in the fp32 reference
    # alike: the softmax inherits them as relative errors
    rt = (1e-5 if name in ("R1", "R2") else 5e-4) if dtype == torch.float32 else 4e-3
    for x, k in zip(dev, feats.keys()):
        got = x.grad.double().cpu().numpy()
        assert x.grad.dtype == dtype
        ref = ref_grads[k]
        assert np.linalg.norm(got - ref) <= rt * np.linalg.norm(ref), k
        assert np.linalg.norm(got[::4] - gz[f"{name}.{k}"]) <= max(rt, 2e-5) * np.linalg.norm(gz[f"{name}.{k}"]), k
    # upstream scaling and a tensor that needs no gradient
    dev2 = [v.to(dtype).cuda().requires_grad_(i != 0) for i, v in enumerate(feats.values())]
    l2 = fn.fused(dev2, norm=False)
    w = torch.arange(1, l2.numel() + 1, device="cuda", dtype=torch.float32)
    (l2 * w).sum().backward()
    assert dev2[0].grad is None
    acc = np.zeros_like(ref_grads[list(feats)[1]])
    keys = list(feats)
    for p, (a, b) in enumerate(combinations(keys, 2)):
        _, ga, gb = NO.ntxent_forward_backward(feats[a].numpy(), feats[b].numpy(), TAU, ALPHA, grad_out=float(p + 1), norm=False)
        if a == keys[1]:
            acc += ga
        if b == keys[1]:
            acc += gb
    got = dev2[1].grad.double().cpu().numpy()
    assert np.linalg.norm(got - acc) <= rt * np.linalg.norm(acc)


def test_unsupported_inputs_raise(tb):
    a = torch.randn(64, 100).cuda()  # dim not a multiple of 64
    fn = tb.loss.NTXentLoss(TAU, ALPHA)
    with pytest.raises(Exception):
        fn(a, a.clone())
    with pytest.raises(Exception):
        tb.loss.NTXentLoss(0.001, ALPHA)(torch.randn(64, 128).cuda(), torch.randn(64, 128).cuda())  # tau too small


# ------------------------------------------------------------------------------ retrieval
def _eval_gpu(tb, tuples, k=5):
    text, gal, labels, fit_labels, _, _ = RO.build_matrices(tuples)
    t = torch.from_numpy(text).cuda()
    g = torch.from_numpy(gal).cuda()
    lab = torch.from_numpy(labels).cuda()
    val, idx, rank = tb.eval.retrieve(t, g, lab, k)
    sim, n_g = tb.ops.sim_gemm(tb.ops.cast_16bit(t, tb.ops.BF16), tb.ops.cast_16bit(g, tb.ops.BF16))
    return text, gal, labels, fit_labels, val.cpu().numpy(), idx.cpu().numpy(), rank.cpu().numpy(), sim[:, :n_g].cpu().numpy()


def test_retrieval_integer_kat_bit_exact(tb):
    """Integer-valued embeddings: every dot product is exact in any precision, rows are full of ties ->
    indices, ranks and metrics must equal the stable-sort restatement bit for bit."""
    tuples = RO.make_integer_kat()
    ref = RO.compute_metrics(tuples)
    text, gal, labels, fit_labels, val, idx, rank, sim = _eval_gpu(tb, tuples)
    assert np.array_equal(sim.astype(np.float64), RO.similarities(text, gal))
    assert np.array_equal(idx, ref["_indices"])
    assert np.array_equal(rank, ref["_rank"])
    assert np.array_equal(val.astype(np.float64), ref["_values"])
    os.chdir("/tmp")
    got = tb.eval.compute_metrics("Text2ShapeChairTable", {"caption_embedding_tuples": tuples})
    for k in ("precision", "recall", "recall_rate", "ndcg"):
        assert np.array_equal(got[k], ref[k]) and got[k].dtype == np.float64 and got[k].shape == (5,)
    assert got["mrr"] == ref["mrr"] and isinstance(got["mrr"], float)


@pytest.mark.parametrize("kw", [
    dict(seed=3, n_shapes=300, n_queries=1000, dim=128, round_bf16=True),
    dict(round_bf16=True),
    dict(round_bf16=True, trimodal_gallery=True),
])
def test_retrieval_val_shaped(tb, golden, kw):
    """C3-shaped data. (1) selection stage bit-exact against the stable-sort restatement on the GPU's own
    fp32 similarity matrix; (2) against the fp64 reference: identical wherever the fp64 margin exceeds 1e-5,
    metrics within 1e-3."""
    tuples = RO.make_val_shaped(**kw)
    text, gal, labels, fit_labels, val, idx, rank, sim = _eval_gpu(tb, tuples)
    # (1) same numbers in, same selection out
    on_gpu_sims = RO.compute_metrics(tuples, sim=sim)
    assert np.array_equal(idx, on_gpu_sims["_indices"])
    assert np.array_equal(rank, on_gpu_sims["_rank"])
    assert np.array_equal(val, on_gpu_sims["_values"])
    os.chdir("/tmp")
    got = tb.eval.compute_metrics("Text2ShapeChairTable", {"caption_embedding_tuples": tuples})
    for k in ("precision", "recall", "recall_rate", "ndcg"):
        assert np.array_equal(got[k], on_gpu_sims[k])
    assert got["mrr"] == on_gpu_sims["mrr"]
    # (2) fp64 reference
    ref = RO.compute_metrics(tuples)
    sim64 = RO.similarities(text, gal)
    assert np.abs(sim - sim64).max() < 5e-6  # bf16-exact inputs, fp32 accumulation over 512 terms
    srt = -np.sort(-sim64, axis=1)
    safe_rows = (srt[:, :5] - srt[:, 1:6]).min(axis=1) > 1e-5
    assert safe_rows.mean() > 0.9
    assert np.array_equal(idx[safe_rows], ref["_indices"][safe_rows])
    s_gt = sim64[np.arange(len(labels)), labels]
    gap = np.abs(sim64 - s_gt[:, None])
    gap[np.arange(len(labels)), labels] = np.inf
    safe_rank = gap.min(axis=1) > 1e-5
    assert safe_rank.mean() > 0.9
    assert np.array_equal(rank[safe_rank], ref["_rank"][safe_rank])
    for k in ("recall_rate", "ndcg"):
        assert np.abs(got[k] - ref[k]).max() < 1e-3
    assert abs(got["mrr"] - ref["mrr"]) < 1e-3


def test_retrieval_vs_reference_golden(tb, golden):
    """The reference's own outputs (fp64) for the C3 generator: metrics within 1e-3."""
    os.chdir("/tmp")
    got = tb.eval.compute_metrics("Text2ShapeChairTable", {"caption_embedding_tuples": RO.make_val_shaped(round_bf16=True)})
    ref = golden["eval"]["C3.bf16"]
    for k in ("precision", "recall", "recall_rate", "ndcg"):
        assert np.abs(got[k] - np.array(ref[k])).max() < 1e-3
    assert abs(got["mrr"] - ref["mrr"]) < 1e-3


def test_inner_trio_signatures(tb):
    tuples = RO.make_val_shaped(seed=3, n_shapes=300, n_queries=1000, dim=128, round_bf16=True)
    E = tb.eval
    (text, gal, labels, fit_labels, m2l, n, l2m) = E.construct_embeddings_matrix("x", {"caption_embedding_tuples": tuples})
    r_text, r_gal, r_labels, r_fit, r_first, r_ids = RO.build_matrices(tuples)
    assert text.dtype == np.float64 and np.array_equal(text, r_text) and np.array_equal(gal, r_gal)
    assert labels.dtype == np.int64 and np.array_equal(labels, r_labels) and np.array_equal(fit_labels, r_fit)
    dist, idx, order = E.compute_nearest_neighbors(gal, text, 5)
    assert dist.shape == (1000, 5) and idx.shape == (1000, 5) and idx.dtype == np.int64 and dist.dtype == np.float64
    ref = RO.compute_metrics(tuples, sim=order.sim[:, :300].cpu().numpy())
    assert np.array_equal(idx, ref["_indices"])
    # reference quirk: distances come back in reverse query order (np.flip without axis, :78)
    assert np.array_equal(dist, ref["_values"][::-1].astype(np.float64))
    pr = E.compute_pr_at_k(idx, order, labels, 5, n, fit_labels)
    for k in ("precision", "recall", "recall_rate", "ndcg"):
        assert np.array_equal(pr[k], ref[k])
    assert pr["mrr"] == ref["mrr"]
    # explicit ordering path (what the reference passes): same result
    pr2 = E.compute_pr_at_k(idx, np.asarray(order), labels, 5, n, fit_labels)
    assert pr2["mrr"] == pr["mrr"]
    q_ids, cats, near, rr = E.get_nearest_info(idx, order, fit_labels, labels, l2m, tuples)
    assert q_ids[0] == tuples[0][2] and len(near) == 1000 and near[0][0] == l2m[int(idx[0, 0])]
    assert rr == [1 / int(r) for r in ref["_rank"]]


def test_retrieval_edge_cases(tb):
    ops = tb.ops
    # gallery smaller than k, ragged sizes, single query
    for (q, g, d) in ((1, 3, 64), (5, 130, 72), (129, 257, 520)):
        gen = torch.Generator().manual_seed(q * g)
        t = torch.randint(-2, 3, (q, d), generator=gen).float().cuda()
        gal = torch.randint(-2, 3, (g, d), generator=gen).float().cuda()
        lab = torch.randint(0, g, (q,), generator=gen).cuda()
        val, idx, rank = tb.eval.retrieve(t, gal, lab, 5)
        sim = (t.double() @ gal.double().t()).cpu().numpy()
        rv, ri, rr = RO.topk_and_rank(sim, lab.cpu().numpy(), min(5, g))
        assert np.array_equal(idx.cpu().numpy()[:, : min(5, g)], ri)
        assert np.array_equal(rank.cpu().numpy(), rr)
        if g < 5:
            assert (idx.cpu().numpy()[:, g:] == -1).all()
    # blocked queries give the same answer as one block
    gen = torch.Generator().manual_seed(77)
    t = torch.randn(1000, 128, generator=gen).bfloat16().float().cuda()
    gal = torch.randn(333, 128, generator=gen).bfloat16().float().cuda()
    lab = torch.randint(0, 333, (1000,), generator=gen).cuda()
    a = tb.eval.retrieve(t, gal, lab, 5)
    b = tb.eval.retrieve(t, gal, lab, 5, block_queries=96)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_large_retrieval_properties(tb):
    """C5-shaped slice (65k queries x 25k gallery = one rank's shard): consistency between the sharded
    (3 gallery shards + merge) and unsharded paths, and planted nearest neighbours are found."""
    ops = tb.ops
    gen = torch.Generator(device="cuda").manual_seed(0)
    G, Q, D = 25000, 16384, 512
    gal = torch.randn(G, D, generator=gen, device="cuda").bfloat16()
    owner = torch.randint(0, G, (Q,), generator=gen, device="cuda")
    text = (gal[owner].float() + 0.5 * torch.randn(Q, D, generator=gen, device="cuda")).bfloat16()
    val, idx, rank = tb.eval.retrieve(text, gal, owner, 5)
    assert (rank == 1).float().mean().item() > 0.99  # planted neighbour wins
    assert (idx[:, 0].long() == owner).float().mean().item() > 0.99
    assert (val[:, :-1] >= val[:, 1:]).all()  # sorted descending
    # sharded path
    bounds = [0, 8000, 17000, G]
    sims = [ops.sim_gemm(text, gal[lo:hi].contiguous()) for lo, hi in zip(bounds[:-1], bounds[1:])]
    gts = sum(ops.gather_gt_sim(s, n, owner, lo) for (s, n), lo in zip(sims, bounds[:-1]))
    cv, ci, nb = [], [], 0
    for (s, n), lo in zip(sims, bounds[:-1]):
        v, i, _, b = ops.topk_rank(s, n, 5, owner, lo, gts)
        cv.append(v); ci.append(i); nb = nb + b
    mv, mi = ops.topk_merge(torch.stack(cv), torch.stack(ci))
    assert torch.equal(mi, idx) and torch.equal(mv, val) and torch.equal(nb + 1, rank)


@pytest.mark.parametrize("q,g,d,k,ties", [
    (1000, 300, 128, 5, False), (7424, 1486, 512, 5, False), (777, 1486, 512, 5, True), (300, 4099, 64, 16, False),
    (50, 3, 64, 5, True), (129, 257, 192, 3, True), (20000, 25000, 512, 5, False),
])
def test_fused_retrieval_matches_unfused_bit_exact(tb, q, g, d, k, ties):
    """K2'+K4 fused (similarities never leave the SM) against GEMM -> HBM -> top-k: identical values, indices
    and ranks, including the ground-truth similarity produced by the diagonal-MMA pre-pass."""
    ops = tb.ops
    gen = torch.Generator().manual_seed(q * 7 + g)
    if ties:
        text = torch.randint(-2, 3, (q, d), generator=gen).float()
        gal = torch.randint(-2, 3, (g, d), generator=gen).float()
    else:
        text = torch.randn(q, d, generator=gen)
        gal = torch.randn(g, d, generator=gen)
    lab = torch.randint(0, g, (q,), generator=gen).cuda()
    t16, g16 = text.cuda().bfloat16(), gal.cuda().bfloat16()
    sim, n_g = ops.sim_gemm(t16, g16)
    gt_ref = torch.gather(sim[:, :n_g], 1, lab[:, None])[:, 0]
    gt = ops.gt_sim_mma(t16, g16, lab)
    assert torch.equal(gt, gt_ref)  # same MMA sequence -> same bits
    a = tb.eval.retrieve(t16, g16, lab, k, fused=False)
    b = tb.eval.retrieve(t16, g16, lab, k, fused=True)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    if ties:  # exact arithmetic: also equal to the oracle
        rv, ri, rr = RO.topk_and_rank((text.double() @ gal.double().t()).numpy(), lab.cpu().numpy(), min(k, g))
        assert np.array_equal(b[1].cpu().numpy()[:, : min(k, g)], ri) and np.array_equal(b[2].cpu().numpy(), rr)


def test_rank_metrics_reduction(tb):
    """K5 (tcl_rank_metrics) + the closed-form finalise equal the reference-order host finalise on the same ranks."""
    ops = tb.ops
    tuples = RO.make_val_shaped(seed=3, n_shapes=1486, n_queries=7424, dim=512, round_bf16=True)
    text, gal, labels, fit_labels, _, _ = RO.build_matrices(tuples)
    t, g, lab = torch.from_numpy(text).float().cuda(), torch.from_numpy(gal).cuda(), torch.from_numpy(labels).cuda()
    val, idx, rank = tb.eval.retrieve(t, g, lab, 5)
    ref = tb.eval.metrics_from_ranks(idx.cpu().numpy().astype(np.int64), rank.cpu().numpy().astype(np.int64), labels, 5, fit_labels)
    got = tb.eval.retrieve_metrics(t, g, lab, 5)
    for k in ("precision", "recall", "recall_rate", "ndcg"):
        assert np.abs(got[k] - ref[k]).max() <= 1e-12, k
    assert abs(got["mrr"] - ref["mrr"]) <= 1e-12
    red = ops.rank_metrics(rank, 5).cpu().numpy()  # counts are exact integers; two launches give identical sums
    r = rank.cpu().numpy()
    assert [int(x) for x in red[:5]] == [int((r == j + 1).sum()) for j in range(5)]
    assert np.array_equal(red, ops.rank_metrics(rank, 5).cpu().numpy())
    big = torch.randint(1, 200000, (1_000_003,), device="cuda", dtype=torch.int32)
    rb = ops.rank_metrics(big, 16).cpu().numpy()
    assert abs(rb[16] - float((1.0 / big.double()).sum())) <= 1e-9 * rb[16]
    assert [int(x) for x in rb[:16]] == [int((big == j + 1).sum()) for j in range(16)]


def test_fused_sharded_matches_unsharded(tb):
    ops = tb.ops
    gen = torch.Generator(device="cuda").manual_seed(5)
    G, Q, D = 25000, 4096, 512
    gal = torch.randn(G, D, generator=gen, device="cuda").bfloat16()
    text = torch.randn(Q, D, generator=gen, device="cuda").bfloat16()
    lab = torch.randint(0, G, (Q,), generator=gen, device="cuda")
    val, idx, rank = tb.eval.retrieve(text, gal, lab, 5, fused=True)
    bounds = [0, 8000, 17000, G]
    shards = [gal[lo:hi].contiguous() for lo, hi in zip(bounds[:-1], bounds[1:])]
    gt = sum(ops.gt_sim_mma(text, s, lab, lo) for s, lo in zip(shards, bounds[:-1]))
    cv, ci, nb = [], [], 0
    for s, lo in zip(shards, bounds[:-1]):
        v, i, b = ops.sim_topk_fused(text, s, 5, lab, gt, lo)
        cv.append(v); ci.append(i); nb = nb + b
    mv, mi = ops.topk_merge(torch.stack(cv), torch.stack(ci))
    assert torch.equal(mi, idx) and torch.equal(mv, val) and torch.equal(nb + 1, rank)


@pytest.mark.parametrize("b_glob,world,interleaved", [(4096, 2, True), (2500, 1, False), (4224, 2, False)])
def test_rank_emulation_offsets_match_unsharded(tb, b_glob, world, interleaved):
    """The sharded form's kernel calls (row_offset / self_offset, rows per rank >= 2048 -> CTA-pair forward and
    persistent backward with cut units, ragged 128-row blocks, modality-interleaved operand rows as in
    tricolo_b200/distributed.py) replayed rank by rank on ONE GPU must reproduce the unsharded loss and gradients."""
    ops = tb.ops
    F16 = 0
    g = torch.Generator().manual_seed(21)
    base = torch.randn(b_glob, 512, generator=g)
    f = [(base + 0.5 * torch.randn(b_glob, 512, generator=g)).bfloat16().float().cuda() for _ in range(3)]
    dev = [x.clone().requires_grad_(True) for x in f]
    losses = tb.loss.trimodal_ntxent(dev, TAU, ALPHA)
    losses.sum().backward()
    pairs = [(0, 1), (0, 2), (1, 2)]
    inv_tau = 1.0 / TAU
    b_loc = b_glob // world
    if interleaved:
        zbuf = torch.empty((b_glob, 3 * 512), dtype=torch.float16, device="cuda")
        outs = [zbuf.view(b_glob, 3, 512)[:, m] for m in range(3)]
        z_all, invs, xs = ops.l2norm_fwd(f, F16, out=outs)
    else:
        z_all, invs, xs = ops.l2norm_fwd(f, F16)
    fw = []
    for r in range(world):
        sl = slice(r * b_loc, (r + 1) * b_loc)
        fw.append(ops.ntxent_fwd([z_all[a][sl] for a, _ in pairs], [z_all[b] for _, b in pairs], r * b_loc, inv_tau, F16))
    col_sum = sum(x[1] for x in fw)
    fin = [ops.ntxent_finalize(fw[r][0], col_sum, fw[r][2], r * b_loc, inv_tau, ALPHA, want_loss=False) for r in range(world)]
    lse2_row_all = torch.cat([x[0] for x in fin], dim=1).contiguous()
    lse2_col = fin[0][1]
    parts = sum(x[2] for x in fin)
    loss = (ALPHA * parts[:, 0] + (1.0 - ALPHA) * parts[:, 1]) / b_glob
    assert torch.allclose(loss, losses.detach(), rtol=1e-5)
    ones = torch.ones((3,), dtype=torch.float32, device="cuda")
    zts, ld_t = ops.transpose_for_bwd(z_all)
    for r in range(world):
        sl = slice(r * b_loc, (r + 1) * b_loc)
        jobs = []
        for m in range(3):
            segs = []
            for p, (a, b) in enumerate(pairs):
                if m == a:
                    segs.append(ops.BwdSegmentSpec(z_all[b], zts[b], lse2_row_all[p, sl], lse2_col[p], ones[p:p + 1], ALPHA, 1.0 - ALPHA))
                elif m == b:
                    segs.append(ops.BwdSegmentSpec(z_all[a], zts[a], lse2_col[p, sl], lse2_row_all[p], ones[p:p + 1], 1.0 - ALPHA, ALPHA))
            jobs.append(ops.BwdJobSpec(z_all[m][sl], xs[m][sl], invs[m][sl], segs))
        dxs = ops.ntxent_bwd(jobs, b_glob, r * b_loc, ld_t, inv_tau, F16)
        for m in range(3):
            # the unsharded call may run the shared-G backward (different fp32 summation order): norm-wise 1e-5,
            # element-wise against the largest entry
            ref = dev[m].grad[sl]
            rel = float((dxs[m] - ref).norm()) / float(ref.norm())
            assert rel <= 1e-4, (r, m, rel)
            assert torch.allclose(dxs[m], ref, rtol=1e-3, atol=1e-4 * ref.abs().max().item()), (r, m)


def _emulated_forward(tb, f, pairs, world, op=0):
    """Forward of the sharded form replayed rank by rank on one GPU: gathered operands [B, n*D] (modalities
    interleaved per row as in tricolo_b200/distributed.py), LSEs of all rows / columns, losses."""
    ops = tb.ops
    n, (b_glob, dim) = len(f), f[0].shape
    b_loc = b_glob // world
    zbuf = torch.empty((b_glob, n * dim), dtype=torch.float16 if op == 0 else torch.bfloat16, device="cuda")
    outs = [zbuf.view(b_glob, n, dim)[:, m] for m in range(n)]
    z_all, invs, xs = ops.l2norm_fwd(f, op, out=outs)
    fw = [ops.ntxent_fwd([z_all[a][r * b_loc:(r + 1) * b_loc] for a, _ in pairs], [z_all[b] for _, b in pairs],
                         r * b_loc, 1.0 / TAU, op) for r in range(world)]
    col_sum = sum(x[1] for x in fw)
    fin = [ops.ntxent_finalize(fw[r][0], col_sum, fw[r][2], r * b_loc, 1.0 / TAU, ALPHA, want_loss=False) for r in range(world)]
    lse2_row_all = torch.cat([x[0] for x in fin], dim=1).contiguous()
    parts = sum(x[2] for x in fin)
    loss = (ALPHA * parts[:, 0] + (1.0 - ALPHA) * parts[:, 1]) / b_glob
    return z_all, torch.stack(invs), xs, lse2_row_all, fin[0][1].contiguous(), loss


@pytest.mark.parametrize("b_glob,world,gsplit,need,rs16", [
    (1024, 8, None, (1, 1, 1), "0"),      # one 128-row block per rank
    (4096, 2, None, (1, 1, 1), "0"),      # accumulator halves (G of all pairs fits L2)
    (4096, 2, "1", (1, 1, 1), "0"),       # full-width accumulators, cut units, CTA-pair gradient GEMM
    (4096, 2, "1", (1, 1, 1), "1"),       # the same with fp16 column-side partials (the default)
    (3072, 4, None, (1, 0, 1), "1"),      # image needs no gradient: its jobs disappear, G of (text,image) is still needed
    (8192, 8, None, (1, 1, 1), "1"),      # BASELINE configs[3] at N=8
])
def test_sharded_sharedg_backward_rank_emulation(tb, monkeypatch, b_glob, world, gsplit, need, rs16):
    """tcl_ntxent_bwd_sharded_gemm / _finish replayed rank by rank on ONE GPU (every rank's receive buffer is a local
    allocation; on a multi-GPU box they are peer-mapped: tests/gpu_multirank.py): the row block of G formed once per
    pair, row-side gradients local, column-side partials stored into the owner's slots, must reproduce the unsharded
    gradients of the concatenated batch."""
    ops = tb.ops
    if gsplit is not None:
        monkeypatch.setenv("TRICOLO_B200_GSPLIT", gsplit)
    monkeypatch.setenv("TRICOLO_B200_RS16", rs16)
    # fp32 partials: only the summation order differs from the unsharded kernels; fp16 partials (what crosses NVLink by
    # default) add one 11-bit rounding per source rank - still inside the 1e-3 budget against the oracle (test below)
    tol = 1e-4 if rs16 == "0" else 6e-4
    g = torch.Generator().manual_seed(33)
    base = torch.randn(b_glob, 512, generator=g)
    f = [(base + 0.5 * torch.randn(b_glob, 512, generator=g)).bfloat16().float().cuda() for _ in range(3)]
    dev = [x.clone().requires_grad_(bool(nd)) for x, nd in zip(f, need)]
    scales = torch.tensor([1.0, 0.5, 2.0], device="cuda")  # unequal upstream gradients per pair
    losses = tb.loss.trimodal_ntxent(dev, TAU, ALPHA)
    (losses * scales).sum().backward()
    pairs = [(0, 1), (0, 2), (1, 2)]
    z_all, invs, xs, lse2_row_all, lse2_col, loss = _emulated_forward(tb, f, pairs, world)
    assert torch.allclose(loss, losses.detach(), rtol=1e-5)
    b_loc = b_glob // world
    plan = ops.ShardedBwdPlan(3, pairs, need, b_loc, world, 512)
    recv = [torch.full((plan.recv_bytes,), 0xFF, dtype=torch.uint8, device="cuda") for _ in range(world)]  # NaN-poisoned
    wss = [torch.empty((plan.workspace_bytes,), dtype=torch.uint8, device="cuda") for _ in range(world)]
    addrs = [r.data_ptr() for r in recv]
    for r in range(world):
        ops.ntxent_bwd_sharded_gemm(plan, z_all, r, 1.0 / TAU, ALPHA, lse2_row_all, lse2_col, scales, wss[r], addrs)
    for r in range(world):
        sl = slice(r * b_loc, (r + 1) * b_loc)
        dxs = ops.ntxent_bwd_sharded_finish(plan, [x[sl] for x in xs], invs[:, sl].contiguous(), r, wss[r], addrs[r])
        for m in range(3):
            if not need[m]:
                assert dxs[m] is None
                continue
            ref = dev[m].grad[sl]
            rel = float((dxs[m] - ref).norm()) / float(ref.norm())
            assert rel <= tol, (r, m, rel)
            # element-wise: the LSEs of the replayed forward differ from the fused forward's in the last fp32 bit, which
            # flips the 16-bit rounding of single entries of G (one ulp of a diagonal entry = 4e-4 of the largest
            # gradient element); both results are equally close to the fp64 oracle (profiles/shard_g_debug.py)
            assert torch.allclose(dxs[m], ref, rtol=1e-3, atol=1e-3 * ref.abs().max().item()), (r, m)


@pytest.mark.parametrize("b_glob,world,fused", [(1024, 8, True), (4096, 2, True), (1024, 8, False), (2048, 4, False),
                                                (2048, 4, None), (1400, 2, None)])
def test_flag_protocol_rank_emulation(tb, monkeypatch, b_glob, world, fused):
    """The barrier-free sharded step (tcl_l2norm_fwd_push -> tcl_ntxent_fwd_sharded -> tcl_ntxent_finalize_sharded ->
    tcl_ntxent_bwd_sharded_gemm/_finish with sync pads) replayed rank by rank on ONE GPU, two steps in a row (the flags
    carry the step number and are never reset).  fused: the all-gather of the column modalities (image, voxel) is done
    by the tile kernel's push warps, text rows stay local; otherwise by K1.  Every wait finds its flag already set by
    an earlier kernel of the replay, except the "ready" flags (a rank signals them at the start of ITS K1), which the
    test sets by hand (and see the two-pass replay of the fused form below)."""
    ops = tb.ops
    pairs = [(0, 1), (0, 2), (1, 2)]
    b_loc = b_glob // world
    n, dim = 3, 512
    if fused is None:
        return _barrier_form_rank_emulation(tb, b_glob, world)
    monkeypatch.setenv("TRICOLO_B200_RS16", "0")  # fp32 partials: compared tightly with the unsharded gradients below
    zbufs = [torch.zeros((b_glob, n * dim), dtype=torch.float16, device="cuda") for _ in range(world)]
    syncs = [torch.zeros((ops.shard_sync_bytes() // 4,), dtype=torch.int32, device="cuda") for _ in range(world)]
    stats = [torch.full((ops.shard_stats_bytes(3, b_loc, world) // 4,), float("nan"), device="cuda") for _ in range(world)]
    plan = ops.ShardedBwdPlan(3, pairs, (1, 1, 1), b_loc, world, dim)
    recv = [torch.full((plan.recv_bytes,), 0xFF, dtype=torch.uint8, device="cuda") for _ in range(world)]
    wss = [torch.empty((plan.workspace_bytes,), dtype=torch.uint8, device="cuda") for _ in range(world)]
    sync_addrs, stats_addrs, recv_addrs = [x.data_ptr() for x in syncs], [x.data_ptr() for x in stats], [x.data_ptr() for x in recv]
    for step in (1, 2):
        g = torch.Generator().manual_seed(40 + step)
        base = torch.randn(b_glob, dim, generator=g)
        f = [(base + 0.5 * torch.randn(b_glob, dim, generator=g)).bfloat16().float().cuda() for _ in range(n)]
        dev = [x.clone().requires_grad_(True) for x in f]
        losses = tb.loss.trimodal_ntxent(dev, TAU, ALPHA)
        losses.sum().backward()
        for sy in syncs:
            sy[16:16 + world] = step  # ShardSync::kReady: every peer is ready for this step
        fwd = []
        for r in range(world):
            sl = slice(r * b_loc, (r + 1) * b_loc)
            dsts = [[zb.data_ptr() + (r * b_loc * n * dim + m * dim) * 2 for m in range(n)] for zb in zbufs]
            fwd.append(ops.l2norm_fwd_push([x[sl] for x in f], dsts, n * dim, r, world, sync_addrs, remote=not fused))
        if not fused:
            for zb in zbufs[1:]:
                assert torch.equal(zb, zbufs[0])
        z_addrs = [zb.data_ptr() for zb in zbufs]
        # fused: rank r's tile kernel needs the rows that the LATER ranks' tile kernels push.  Replay in two passes: in
        # the first every arrival flag is pre-set, so no kernel waits (its remote tiles read rows that may not have
        # landed yet: results discarded) but every rank's push warps deliver and flag their rows; the second pass then
        # finds all rows in place.  (The waiting itself is exercised across real GPUs by tests/gpu_multirank.py.)
        for replay in ((0, 1) if fused else (1,)):
            if fused and replay == 0:
                for sy in syncs:
                    sy[64:64 + 8 * 64] = step  # ShardSync::kArrived
            for r in range(world):
                z3 = zbufs[r].view(b_glob, n, dim)
                z_all = [z3[:, m] for m in range(n)]
                sl = slice(r * b_loc, (r + 1) * b_loc)
                ops.ntxent_fwd_sharded([z_all[a][sl] for a, _ in pairs], [z_all[b] for _, b in pairs], r, world, 1.0 / TAU,
                                       stats_addrs, sync_addrs, z_base_addrs=z_addrs,
                                       push_offsets=[dim, 2 * dim] if fused else ())
        if fused:  # image and voxel rows are everywhere, text rows only at home
            for zb in zbufs[1:]:
                assert torch.equal(zb.view(b_glob, n, dim)[:, 1:], zbufs[0].view(b_glob, n, dim)[:, 1:])
        fin = [ops.ntxent_finalize_sharded(3, b_loc, r, world, 1.0 / TAU, ALPHA, stats_addrs, sync_addrs[r], f[0].device)
               for r in range(world)]
        for r in range(world):
            assert torch.equal(fin[r][0], fin[0][0]) and torch.equal(fin[r][1], fin[0][1]) and torch.equal(fin[r][2], fin[0][2])
        assert torch.allclose(fin[0][2], losses.detach(), rtol=1e-5)
        ones = torch.ones((3,), device="cuda")
        for r in range(world):
            z3 = zbufs[r].view(b_glob, n, dim)
            ops.ntxent_bwd_sharded_gemm(plan, [z3[:, m] for m in range(n)], r, 1.0 / TAU, ALPHA, fin[r][0], fin[r][1], ones,
                                        wss[r], recv_addrs, sync_addrs=sync_addrs)
        for r in range(world):
            sl = slice(r * b_loc, (r + 1) * b_loc)
            inv_all, xs = fwd[r]
            dxs = ops.ntxent_bwd_sharded_finish(plan, xs, inv_all, r, wss[r], recv_addrs[r], sync_own_addr=sync_addrs[r])
            for m in range(n):
                ref = dev[m].grad[sl]
                assert float((dxs[m] - ref).norm()) <= 1e-4 * float(ref.norm()), (step, r, m)
        for sy in syncs:  # step counters: forward epoch, backward epoch, and every flag at this step's value
            assert int(sy[0]) == step and int(sy[1]) == step
            assert all(int(v) == step for v in sy[32:32 + world]) and all(int(v) == step for v in sy[48:48 + world])


def _barrier_form_rank_emulation(tb, b_glob, world):
    """fused=None: the barrier form of the same entry points (sync pointers NULL): contiguous tile ranges, statistics
    written into the own slot and PULLED by the finalise kernel from every rank's buffer; any rows per rank."""
    ops = tb.ops
    pairs = [(0, 1), (0, 2), (1, 2)]
    b_loc, n, dim = b_glob // world, 3, 512
    g = torch.Generator().manual_seed(50)
    base = torch.randn(b_glob, dim, generator=g)
    f = [(base + 0.5 * torch.randn(b_glob, dim, generator=g)).bfloat16().float().cuda() for _ in range(n)]
    losses = tb.loss.trimodal_ntxent(f, TAU, ALPHA)
    zbuf = torch.zeros((b_glob, n * dim), dtype=torch.float16, device="cuda")
    stats = [torch.full((ops.shard_stats_bytes(3, b_loc, world) // 4,), float("nan"), device="cuda") for _ in range(world)]
    stats_addrs = [x.data_ptr() for x in stats]
    for r in range(world):  # tcl_l2norm_fwd_bcast (16-byte stores): here all destinations are one buffer
        sl = slice(r * b_loc, (r + 1) * b_loc)
        ops.l2norm_fwd_bcast([x[sl] for x in f], [[zbuf.data_ptr() + (r * b_loc * n * dim + m * dim) * 2 for m in range(n)]], n * dim)
    z_ref, _, _ = ops.l2norm_fwd(f, 0)
    for m in range(n):
        assert torch.equal(zbuf.view(b_glob, n, dim)[:, m], z_ref[m])
    z_all = [zbuf.view(b_glob, n, dim)[:, m] for m in range(n)]
    for r in range(world):
        sl = slice(r * b_loc, (r + 1) * b_loc)
        ops.ntxent_fwd_sharded([z_all[a][sl] for a, _ in pairs], [z_all[b] for _, b in pairs], r, world, 1.0 / TAU, stats_addrs, None)
    fin = [ops.ntxent_finalize_sharded(3, b_loc, r, world, 1.0 / TAU, ALPHA, stats_addrs, 0, f[0].device) for r in range(world)]
    for r in range(world):
        assert all(torch.equal(fin[r][k], fin[0][k]) for k in range(3))
    assert torch.allclose(fin[0][2], losses, rtol=1e-5)
    z_all2, invs, xs, lse2_row_all, lse2_col, loss = _emulated_forward(tb, f, pairs, world)
    assert torch.allclose(fin[0][0], lse2_row_all, rtol=1e-6, atol=1e-6) and torch.allclose(fin[0][1], lse2_col, rtol=1e-6, atol=1e-6)


def test_bcast_normalise_and_peer_sum_single_gpu(tb):
    """tcl_l2norm_fwd_bcast with several local destinations must equal tcl_l2norm_fwd bit for bit (on a multi-GPU
    box the destinations are peer-mapped buffers: tests/gpu_multirank.py); tcl_peer_sum_f32 adds in the given order."""
    ops = tb.ops
    g = torch.Generator().manual_seed(5)
    xs = [torch.randn(300, 512, generator=g).cuda() for _ in range(3)]
    xs[1][7] = 0  # clamped row
    zs, invs, _ = ops.l2norm_fwd(xs, 0)
    bufs = [torch.zeros((700, 3 * 512), dtype=torch.float16, device="cuda") for _ in range(3)]
    lo = 200
    dsts = [[b.data_ptr() + (lo * 3 * 512 + m * 512) * 2 for m in range(3)] for b in bufs]
    invs2, _ = ops.l2norm_fwd_bcast(xs, dsts, 3 * 512, 0)
    for b in bufs:
        v = b.view(700, 3, 512)
        for m in range(3):
            assert torch.equal(v[lo:lo + 300, m], zs[m])
            assert torch.equal(invs2[m], invs[m])
        assert float(v[:lo].abs().max()) == 0 and float(v[lo + 300:].abs().max()) == 0
    # a NULL address at a further destination skips that (destination, tensor): the modality nobody reads remotely
    # (text under the sharded shared-G backward) costs no NVLink traffic
    bufs2 = [torch.zeros((700, 3 * 512), dtype=torch.float16, device="cuda") for _ in range(3)]
    dsts2 = [[b.data_ptr() + (lo * 3 * 512 + m * 512) * 2 for m in range(3)] for b in bufs2]
    dsts2 = [dsts2[0]] + [[0] + list(d[1:]) for d in dsts2[1:]]
    ops.l2norm_fwd_bcast(xs, dsts2, 3 * 512, 0)
    for r, b in enumerate(bufs2):
        v = b.view(700, 3, 512)
        for m in range(3):
            if r > 0 and m == 0:
                assert float(v[:, m].abs().max()) == 0
            else:
                assert torch.equal(v[lo:lo + 300, m], zs[m])
    # tcl_copy_rows: copy-engine transfer between pitched buffers (columns 8..15 of 24 int16 per row)
    src = torch.arange(8 * 24, dtype=torch.int16, device="cuda").view(8, 24) + 1
    dst = torch.zeros_like(src)
    ops.copy_rows(dst.data_ptr() + 16, 48, src.data_ptr() + 16, 48, 16, 8, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(dst[:, 8:16], src[:, 8:16])
    assert int(dst[:, :8].abs().sum()) == 0 and int(dst[:, 16:].abs().sum()) == 0
    parts = [torch.randn(3, 3, 1000, generator=g).cuda() for _ in range(5)]
    want = parts[0].clone()
    for p in parts[1:]:
        want = want + p
    assert torch.equal(ops.peer_sum(parts), want)
    # odd element counts (3 * P * b_glob with an odd local batch at world 2: 3 * 3 * 114 = 1026, not a multiple of 4)
    parts = [torch.randn(3, 3, 114, generator=g).cuda() for _ in range(2)]
    assert torch.equal(ops.peer_sum(parts), parts[0] + parts[1])


def test_gather_sum_cast16(tb):
    ops = tb.ops
    g = torch.Generator().manual_seed(6)
    a = torch.randn(500, 256, generator=g).cuda()
    b = torch.randn(500, 256, generator=g).cuda()
    idx = torch.randint(0, 500, (137,), generator=g).cuda()
    got = ops.gather_sum_cast16([a, b], idx, 1)
    assert torch.equal(got, (a[idx] + b[idx]).bfloat16())
    assert torch.equal(ops.gather_sum_cast16([a], idx, 0), a[idx].half())
    with pytest.raises(IndexError):
        ops.gather_sum_cast16([a], torch.tensor([500], device="cuda"), 1)


def test_host_pipelined_loss_matches_direct(tb):
    """HostPipelinedLoss (H2D / graph replay / D2H of consecutive steps on three streams) returns, for every step,
    exactly what the direct call returns for that step's inputs."""
    from tricolo_b200.graphs import HostPipelinedLoss

    g = torch.Generator().manual_seed(31)
    steps = []
    for _ in range(7):
        base = torch.randn(384, 512, generator=g)
        steps.append([(base + 0.5 * torch.randn(384, 512, generator=g)).pin_memory() for _ in range(3)])
    want = []
    for fs in steps:
        dv = [f.cuda().requires_grad_(True) for f in fs]
        ls = tb.loss.trimodal_ntxent(dv, TAU, ALPHA)
        ls.sum().backward()
        want.append((ls.detach().cpu(), [d.grad.cpu() for d in dv]))
    pipe = HostPipelinedLoss(steps[0], TAU, ALPHA, depth=3)
    tickets, got = [], []
    for k, fs in enumerate(steps):
        tickets.append(pipe.submit(fs))
        if k >= 2:
            ls, gr = pipe.result(tickets[k - 2])
            got.append((ls.clone(), [x.clone() for x in gr]))
    for t in tickets[-2:]:
        ls, gr = pipe.result(t)
        got.append((ls.clone(), [x.clone() for x in gr]))
    assert len(got) == len(want)
    for (l0, g0), (l1, g1) in zip(got, want):
        assert torch.equal(l0, l1)
        for a, b in zip(g0, g1):
            assert torch.equal(a, b)


@pytest.mark.parametrize("b,n_mod", [(333, 3), (2500, 3), (1100, 2)])
def test_backward_forms_agree(tb, b, n_mod, monkeypatch):
    """The shared-G backward (G of a pair formed once, two GEMM kernels) and the producer/consumer backward (logits
    recomputed per direction) are two implementations of the same gradient: both within rtol 1e-3 of the oracle and
    within fp32 summation noise of each other, including ragged batches, two modalities and unequal upstream scales."""
    monkeypatch.setenv("TRICOLO_B200_SMALL", "0")  # b = 333 would take the single-launch form (test_small_batch_*)
    g = torch.Generator().manual_seed(77 + b)
    base = torch.randn(b, 512, generator=g)
    feats = [(base + 0.5 * torch.randn(b, 512, generator=g)).bfloat16().float() for _ in range(n_mod)]
    keys = ["text_features", "image_features", "voxel_features"][:n_mod]
    n_pairs = n_mod * (n_mod - 1) // 2
    w = torch.tensor([1.0, -0.5, 3.0][:n_pairs])
    ref_l, _ = NO.trimodal_forward_backward({k: f.numpy() for k, f in zip(keys, feats)}, TAU, ALPHA)
    grads = {}
    for mode in ("sharedg", "pc"):
        monkeypatch.setenv("TRICOLO_B200_BWD", mode)
        dev = [f.cuda().requires_grad_(True) for f in feats]
        losses = tb.loss.trimodal_ntxent(dev, TAU, ALPHA)
        (losses * w.cuda()).sum().backward()
        grads[mode] = [d.grad.clone() for d in dev]
    # oracle gradient of the weighted sum: linear in the per-pair gradients
    tot = [np.zeros((b, 512)) for _ in range(n_mod)]
    pairs = [(i, j) for i in range(n_mod) for j in range(i + 1, n_mod)]
    for p, (i, j) in enumerate(pairs):
        _, gp = NO.trimodal_forward_backward({keys[i]: feats[i].numpy(), keys[j]: feats[j].numpy()}, TAU, ALPHA)
        tot[i] += float(w[p]) * gp[keys[i]]
        tot[j] += float(w[p]) * gp[keys[j]]
    for m in range(n_mod):
        for mode in ("sharedg", "pc"):
            err = np.linalg.norm(grads[mode][m].double().cpu().numpy() - tot[m]) / np.linalg.norm(tot[m])
            assert err <= RTOL, (mode, m, err)
        # with unequal upstream scales the two forms round DIFFERENT multiples of G to 16 bit (per-tensor vs global
        # max|grad_scale|): they differ by independent rounding noise of the size of each one's own error
        rel = float((grads["sharedg"][m] - grads["pc"][m]).norm()) / float(grads["pc"][m].norm())
        assert rel <= RTOL, (m, rel)


@pytest.mark.parametrize("dim", [64, 192, 320, 448])
@pytest.mark.parametrize("b", [300, 2200])
def test_other_dims_all_kernels(tb, dim, b, monkeypatch):
    """dim != 512: odd numbers of 64-wide K-blocks (ring slots with one block, partly empty accumulator chunks), the
    dim <= 256 kernels, both forward kernels (b >= 2048 -> CTA pair) and both backward forms, against the oracle."""
    monkeypatch.setenv("TRICOLO_B200_SMALL", "0")  # b = 300 would take the single-launch form (test_small_batch_*)
    g = torch.Generator().manual_seed(1000 + dim + b)
    base = torch.randn(b, dim, generator=g)
    feats = [(base + 0.5 * torch.randn(b, dim, generator=g)).bfloat16().float() for _ in range(3)]
    keys = ["text_features", "image_features", "voxel_features"]
    ref_l, ref_g = NO.trimodal_forward_backward({k: f.numpy() for k, f in zip(keys, feats)}, TAU, ALPHA)
    names = ["train_loss/text_image_loss", "train_loss/text_voxel_loss", "train_loss/image_voxel_loss"]
    for mode in (("sharedg", "pc") if dim > 256 else ("pc",)):
        monkeypatch.setenv("TRICOLO_B200_BWD", mode)
        dev = [f.cuda().requires_grad_(True) for f in feats]
        losses = tb.loss.trimodal_ntxent(dev, TAU, ALPHA)
        losses.sum().backward()
        for p, k in enumerate(names):
            assert float(losses[p].detach()) == pytest.approx(ref_l[k], rel=RTOL), (mode, k)
        for x, k in zip(dev, keys):
            err = np.linalg.norm(x.grad.double().cpu().numpy() - ref_g[k]) / np.linalg.norm(ref_g[k])
            assert err <= RTOL, (mode, k, err)


# ---------------------------------------------------------------------------------------------------------------
# small-batch single-launch form (csrc/ntxent_small.cu): taken automatically when every 64 x 64 tile gets its own SM
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("b,dim,n_mod", [(128, 512, 2), (256, 512, 3), (57, 512, 3), (200, 320, 3), (200, 64, 3),
                                         (230, 192, 3), (448, 448, 2), (2, 512, 2), (64, 128, 3)])
def test_small_batch_form_vs_oracle_and_pipeline(tb, b, dim, n_mod, monkeypatch):
    """C1 / C2 (config/config.yaml:62-66: batch 128 bimodal, 256 trimodal) and ragged / other-dim shapes: the
    single-launch form against the fp64 oracle (rtol 1e-3), and against the multi-kernel pipeline on the same inputs
    (same 16-bit operands, different summation orders: fp32 noise), with unequal upstream scales."""
    from tricolo_b200 import _lib as LB

    g = torch.Generator().manual_seed(4000 + b + dim)
    base = torch.randn(b, dim, generator=g)
    feats = [(base + 0.5 * torch.randn(b, dim, generator=g)).bfloat16().float() for _ in range(n_mod)]
    keys = ["text_features", "image_features", "voxel_features"][:n_mod]
    pairs = [(i, j) for i in range(n_mod) for j in range(i + 1, n_mod)]
    w = torch.tensor([1.0, -0.5, 3.0][:len(pairs)])
    out = {}
    for form in ("1", "0"):
        monkeypatch.setenv("TRICOLO_B200_SMALL", form)
        LB.profile_enable(True)
        dev = [f.cuda().requires_grad_(True) for f in feats]
        losses = tb.loss.trimodal_ntxent(dev, TAU, ALPHA)
        (losses * w.cuda()).sum().backward()
        torch.cuda.synchronize()
        out[form] = (losses.detach().cpu().numpy(), [d.grad.double().cpu().numpy() for d in dev])
        ran = LB.profile_read()  # the form that ran is the form that was asked for
        LB.profile_enable(False)
        for name, want in (("ntxent_small_fwd", form == "1"), ("ntxent_small_bwd", form == "1"), ("ntxent_fwd", form == "0")):
            assert (name in ran) == want, (form, name, sorted(ran))
    tot = [np.zeros((b, dim)) for _ in range(n_mod)]
    for p, (i, j) in enumerate(pairs):
        ref_l, gp = NO.trimodal_forward_backward({keys[i]: feats[i].numpy(), keys[j]: feats[j].numpy()}, TAU, ALPHA)
        (ref,) = [v for k, v in ref_l.items() if not k.endswith("total_loss")]
        for form in ("1", "0"):
            assert float(out[form][0][p]) == pytest.approx(ref, rel=RTOL, abs=1e-6), (form, p)
        tot[i] += float(w[p]) * gp[keys[i]]
        tot[j] += float(w[p]) * gp[keys[j]]
    for m in range(n_mod):
        den = max(np.linalg.norm(tot[m]), 1e-12)
        for form in ("1", "0"):
            assert np.linalg.norm(out[form][1][m] - tot[m]) / den <= RTOL, (form, m)
        assert np.linalg.norm(out["1"][1][m] - out["0"][1][m]) / den <= RTOL


def test_small_batch_form_mixed_with_pipeline_and_replay(tb, monkeypatch):
    """The two forms share the state layout: a single-launch forward followed by a pipeline backward (and the reverse)
    gives the same gradients; the single-launch form is deterministic (bit-identical on replay, also from a CUDA
    graph) and honours requires_grad subsets."""
    g = torch.Generator().manual_seed(99)
    feats = [torch.randn(256, 512, generator=g).bfloat16().float() for _ in range(3)]

    def run(fwd_form, bwd_form, need=(True, True, True)):
        dev = [f.cuda().requires_grad_(n) for f, n in zip(feats, need)]
        monkeypatch.setenv("TRICOLO_B200_SMALL", fwd_form)
        losses = tb.loss.trimodal_ntxent(dev, TAU, ALPHA)
        monkeypatch.setenv("TRICOLO_B200_SMALL", bwd_form)
        losses.sum().backward()
        return losses.detach().clone(), [None if d.grad is None else d.grad.clone() for d in dev]

    l11, g11 = run("1", "1")
    l11b, g11b = run("1", "1")
    assert torch.equal(l11, l11b) and all(torch.equal(a, b) for a, b in zip(g11, g11b))
    for ff, bf in (("1", "0"), ("0", "1"), ("0", "0")):
        l, gr = run(ff, bf)
        assert torch.allclose(l, l11, rtol=1e-5)
        for a, b in zip(gr, g11):
            assert float((a - b).norm()) <= 5e-4 * float(b.norm()), (ff, bf)
    # the sum output and its upstream gradient (calculate_losses -> total_loss.backward(), the training step): the
    # same numbers as summing the pair losses with framework ops, in both forms, also when both outputs are used
    for form in ("1", "0"):
        monkeypatch.setenv("TRICOLO_B200_SMALL", form)
        ref_l, ref_g = run(form, form)
        dev = [f.cuda().requires_grad_(True) for f in feats]
        d = tb.loss.calculate_losses(dict(zip(["text_features", "image_features", "voxel_features"], dev)), "train_loss",
                                     tb.loss.NTXentLoss(TAU, ALPHA))
        assert list(d) == ["train_loss/text_image_loss", "train_loss/text_voxel_loss", "train_loss/image_voxel_loss",
                           "train_loss/total_loss"]
        tot = d["train_loss/total_loss"]
        assert tot.dim() == 0 and float(tot) == pytest.approx(float(ref_l.sum()), rel=1e-6)
        tot.backward()
        assert all(torch.equal(x.grad, r) for x, r in zip(dev, ref_g)), form
        dev = [f.cuda().requires_grad_(True) for f in feats]
        ls, tot = tb.loss.trimodal_ntxent_total(dev, TAU, ALPHA)
        (2.0 * tot + (ls * torch.tensor([1.0, -3.0, 0.5], device="cuda")).sum()).backward()
        dev2 = [f.cuda().requires_grad_(True) for f in feats]
        (tb.loss.trimodal_ntxent(dev2, TAU, ALPHA) * torch.tensor([3.0, -1.0, 2.5], device="cuda")).sum().backward()
        for x, y in zip(dev, dev2):
            assert float((x.grad - y.grad).norm()) <= 1e-6 * float(y.grad.norm()), form
    _, gsub = run("1", "1", need=(True, False, True))
    assert gsub[1] is None
    for m in (0, 2):  # the other tensors' gradients do not depend on who else wants one
        assert torch.equal(gsub[m], g11[m])
    # CUDA graph capture of the two cooperative launches
    monkeypatch.setenv("TRICOLO_B200_SMALL", "1")
    static = [f.cuda().requires_grad_(True) for f in feats]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            tb.loss.trimodal_ntxent(static, TAU, ALPHA).sum().backward()
    torch.cuda.current_stream().wait_stream(s)
    for t in static:
        t.grad = None
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        lg = tb.loss.trimodal_ntxent(static, TAU, ALPHA)
        lg.sum().backward()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(lg.detach(), l11)
    assert all(torch.equal(t.grad, b) for t, b in zip(static, g11))
