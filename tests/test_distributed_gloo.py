"""CPU, world_size 2, gloo: host logic of the sharded (multi-GPU) paths of tricolo_b200/distributed.py
with the CUDA ops replaced by contract-equivalent CPU stand-ins (tests/cpu_ops_mock.py), checked against
the single-process oracle on the concatenated batch / unsharded gallery."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAU, ALPHA = 0.1, 0.25


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, what, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import tests.cpu_ops_mock as mock
    import tricolo_b200.distributed as D

    D.ops = mock  # swap the kernel layer for its CPU contract
    try:
        if what == "loss":
            g = torch.Generator().manual_seed(7)
            b, d = 64 * world, 64
            base = torch.randn(b, d, generator=g)
            full = [(base + 0.5 * torch.randn(b, d, generator=g)) for _ in range(3)]
            bl = b // world
            loc = [f[rank * bl:(rank + 1) * bl].clone().requires_grad_(True) for f in full]
            feats = {"text_features": loc[0], "image_features": loc[1], "voxel_features": loc[2]}
            out = D.global_calculate_losses(feats, "train_loss", TAU, ALPHA)
            out["train_loss/total_loss"].backward()
            q.put((rank, {k: float(v) for k, v in out.items()}, [x.grad.numpy() for x in loc], [f.numpy() for f in full]))
        else:
            from oracle import retrieval_oracle as RO
            tuples = RO.make_val_shaped(seed=3, n_shapes=301, n_queries=500, dim=64, round_bf16=True)
            text, gal, labels, *_ = RO.build_matrices(tuples)
            per = (gal.shape[0] + world - 1) // world
            lo, hi = rank * per, min((rank + 1) * per, gal.shape[0])
            v, i, r = D.sharded_retrieve(torch.from_numpy(text).float(), torch.from_numpy(gal[lo:hi]),
                                         torch.from_numpy(labels), lo, 5, block_queries=128)
            q.put((rank, v.numpy(), i.numpy(), r.numpy()))
    finally:
        dist.destroy_process_group()


def _run(what, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, what, q)) for r in range(world)]
    for p in procs:
        p.start()
    import queue as _queue
    res = []
    while len(res) < world:
        try:
            res.append(q.get(timeout=2))
        except _queue.Empty:
            assert all(p.is_alive() or p.exitcode == 0 for p in procs), "a worker died"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(res, key=lambda t: t[0])


@pytest.mark.timeout(300)
def test_global_negative_loss_world2():
    from oracle import ntxent_oracle as NO

    res = _run("loss")
    full = res[0][3]
    ref_l, ref_g = NO.trimodal_forward_backward(
        {"text_features": full[0], "image_features": full[1], "voxel_features": full[2]}, TAU, ALPHA)
    for rank, losses, grads, _ in res:
        assert list(losses) == list(ref_l)  # same keys, same order on every rank
        for k, v in ref_l.items():
            assert losses[k] == pytest.approx(v, rel=2e-4)
        bl = full[0].shape[0] // 2
        for m, key in enumerate(["text_features", "image_features", "voxel_features"]):
            ref = ref_g[key][rank * bl:(rank + 1) * bl]
            assert np.linalg.norm(grads[m] - ref) <= 1e-3 * np.linalg.norm(ref)
    assert res[0][1] == res[1][1]  # identical loss on both ranks


@pytest.mark.timeout(300)
def test_sharded_retrieval_world2():
    from oracle import retrieval_oracle as RO

    res = _run("retrieval")
    tuples = RO.make_val_shaped(seed=3, n_shapes=301, n_queries=500, dim=64, round_bf16=True)
    ref = RO.compute_metrics(tuples)
    for rank, v, i, r in res:
        assert np.array_equal(i, ref["_indices"])
        assert np.array_equal(r, ref["_rank"])
    assert np.array_equal(res[0][1], res[1][1])


def test_gather_plan_follows_the_backward_form(monkeypatch):
    """Which modalities cross NVLink in K1's gather (tricolo_b200/distributed.py:_gather_plan): the column side of the
    pairs always; the row-only modality (text) only for the directional backward."""
    from tricolo_b200.distributed import _gather_plan

    monkeypatch.delenv("TRICOLO_B200_GATHER_ALL", raising=False)
    monkeypatch.delenv("TRICOLO_B200_DEFER_GATHER", raising=False)
    pairs = [(0, 1), (0, 2), (1, 2)]  # (text, image), (text, voxel), (image, voxel): tricolo_net.py:59-61
    dsts = [[100 + 10 * r + m for m in range(3)] for r in range(4)]  # own buffer first
    sg, d = _gather_plan(dsts, pairs, 3, True, True, False)
    assert sg[0] == dsts[0] and d == []
    assert all(row[0] == 0 and row[1:] == ref[1:] for row, ref in zip(sg[1:], dsts[1:]))  # text stays local
    assert _gather_plan(dsts, pairs, 3, False, False, False)[0] == sg  # forward without gradients: the same
    assert _gather_plan(dsts, pairs, 3, False, True, False) == (dsts, [])  # directional backward: everything
    monkeypatch.setenv("TRICOLO_B200_DEFER_GATHER", "1")
    assert _gather_plan(dsts, pairs, 3, False, True, False) == (sg, [0])  # text by the copy engines
    assert _gather_plan(dsts, pairs, 3, True, True, False) == (sg, [])
    monkeypatch.setenv("TRICOLO_B200_GATHER_ALL", "1")
    assert _gather_plan(dsts, pairs, 3, True, True, False) == (dsts, [])
    monkeypatch.delenv("TRICOLO_B200_GATHER_ALL")
    assert _gather_plan(dsts, pairs, 3, True, True, True) == (dsts, [])  # one multicast mapping: nothing to skip
    # bimodal (text, voxel): voxel is the column side
    assert _gather_plan([[1, 2], [3, 4]], [(0, 1)], 2, True, True, False)[0] == [[1, 2], [0, 4]]
