"""Multi-GPU (NCCL) check of tricolo_b200/distributed.py against the single-process oracle.
   torchrun --nproc-per-node N tests/gpu_multirank.py        (also used by tests/test_gpu_multigpu.py)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TAU, ALPHA = 0.1, 0.25


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import ntxent_oracle as NO
    from oracle import retrieval_oracle as RO
    from tricolo_b200.distributed import global_calculate_losses, sharded_retrieve
    from tricolo_b200.evaluation import retrieve

    ok = True
    keys = ["text_features", "image_features", "voxel_features"]

    def run_loss(loc, transport, bwd, sync="barrier"):
        os.environ["TRICOLO_B200_SYMM"] = transport
        os.environ["TRICOLO_B200_SHARDED_BWD"] = bwd
        # sync "defer": barrier form with the text rows sent by the copy engines during the forward (opt-in experiment)
        os.environ["TRICOLO_B200_SHARD_SYNC"] = "barrier" if sync == "defer" else sync
        os.environ["TRICOLO_B200_DEFER_GATHER"] = "1" if sync == "defer" else "0"
        for x in loc:
            x.grad = None
        out = global_calculate_losses(dict(zip(keys, loc)), "train_loss", TAU, ALPHA)
        out["train_loss/total_loss"].backward()
        return {k: float(v.detach()) for k, v in out.items()}, [x.grad.clone() for x in loc]

    def close(a, b):  # different kernels / summation orders / fp16 partials over NVLink: norm-wise 6e-4 (every form is
        # checked against the fp64 oracle at 1e-3 above), element-wise against the largest entry
        return (float((a - b).norm()) <= 6e-4 * float(b.norm()) and
                torch.allclose(a, b, rtol=1e-3, atol=1e-3 * float(b.abs().max())))

    # ---- global-negative trimodal loss vs the single-process fp64 oracle.  Rows per rank: 256 (one or two 128-row
    # blocks per unit), 57 (odd: ragged blocks, the statistics buffer is not a multiple of four floats), 1024 (cut units)
    for bl in (256, 57, 1024):
        b = bl * world
        g = torch.Generator().manual_seed(7)
        base = torch.randn(b, 512, generator=g)
        full = [(base + 0.5 * torch.randn(b, 512, generator=g)).bfloat16().float() for _ in range(3)]
        loc = [f[rank * bl:(rank + 1) * bl].cuda().requires_grad_(True) for f in full]
        ref_l, ref_g = NO.trimodal_forward_backward(dict(zip(keys, [f.numpy() for f in full])), TAU, ALPHA)
        results = {}
        # symm/sharedg: NVLink peer memory, sharded shared-G backward with the in-kernel reduce-scatter; symm/pc:
        # directional backward; .../flags: the barrier-free flag protocol instead of barrier kernels (fused all-gather
        # push warps, flag-signalled reduce-scatter); nccl: NCCL collectives
        for name, transport, bwd, sync in (("symm/sharedg", "1", "sharedg", "barrier"), ("symm/sharedg/flags", "1", "sharedg", "flags"),
                                           ("symm/pc", "1", "pc", "barrier"), ("symm/pc/flags", "1", "pc", "flags"),
                                           ("symm/pc/defer", "1", "pc", "defer"), ("nccl/pc", "0", "pc", "barrier")):
            losses, grads = run_loss(loc, transport, bwd, sync)
            lerr = max(abs(losses[k] - v) / abs(v) for k, v in ref_l.items())
            errs = [np.linalg.norm(grads[m].double().cpu().numpy() - ref_g[k][rank * bl:(rank + 1) * bl]) /
                    np.linalg.norm(ref_g[k][rank * bl:(rank + 1) * bl]) for m, k in enumerate(keys)]
            ok &= lerr < 1e-3 and max(errs) < 1e-3
            results[name] = (losses["train_loss/total_loss"], grads)
            print(f"[rank {rank}] rows/rank {bl} {name}: loss rel err {lerr:.2e} grad errs {['%.2e' % e for e in errs]}", flush=True)
        same = all(abs(results[n][0] - results["nccl/pc"][0]) <= 1e-6 * abs(results["nccl/pc"][0]) for n in results)
        # (the flag protocol sweeps the column tiles in arrival order: the sum-exp statistics differ in the last fp32
        # bit from the barrier / NCCL forms, which flips the 16-bit rounding of single entries of G)
        same &= all(close(a, b) for n in results for a, b in zip(results[n][1], results["nccl/pc"][1]))
        ok &= same
        print(f"[rank {rank}] rows/rank {bl}: transports and backward forms agree: {same}", flush=True)
    # two forwards before the two backwards: the second forward must not overwrite the operands the first backward needs
    os.environ["TRICOLO_B200_SYMM"] = "1"
    os.environ["TRICOLO_B200_SHARDED_BWD"] = "sharedg"
    os.environ["TRICOLO_B200_SHARD_SYNC"] = "barrier"
    first = results["symm/sharedg"][1]
    loc2 = [(x.detach() * 0.5 + 0.1).requires_grad_(True) for x in loc]
    for x in loc:
        x.grad = None
    o1 = global_calculate_losses(dict(zip(keys, loc)), "a", TAU, ALPHA)
    o2 = global_calculate_losses(dict(zip(keys, loc2)), "a", TAU, ALPHA)
    o2["a/total_loss"].backward()
    o1["a/total_loss"].backward()
    inter = all(close(x.grad, g) for x, g in zip(loc, first))
    ok &= inter
    print(f"[rank {rank}] interleaved forwards keep their operands: {inter}", flush=True)

    # ---- BASELINE configs[3] itself (global batch 8192, bench.py's inputs) against the golden output of the UNMODIFIED
    # reference (tests/golden/make_golden_large.py: 3 losses + every 64th gradient row), when 8192 splits over the ranks
    if 8192 % (128 * world) == 0:
        import json

        from bench import make_features

        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "large_outputs.json")))["c4"]
        gg = np.load(os.path.join(ROOT, "tests", "golden", "c4_grads.npz"))
        bl = 8192 // world
        feats = make_features(8192, bl, rank * bl)
        loc = [feats[k].cuda().requires_grad_(True) for k in keys]
        for name, transport, bwd in (("symm/sharedg", "1", "sharedg"), ("nccl/pc", "0", "pc")):
            losses, grads = run_loss(loc, transport, bwd)
            lerr = max(abs(losses[k] - v) / abs(v) for k, v in gold["losses"].items())
            rows = np.arange(0, 8192, 64)
            mine = rows[(rows >= rank * bl) & (rows < (rank + 1) * bl)]
            num = torch.zeros(3, dtype=torch.float64, device="cuda")
            den = torch.zeros(3, dtype=torch.float64, device="cuda")
            for m, k in enumerate(keys):
                ref = torch.from_numpy(gg[k][mine // 64]).cuda().double()
                got = grads[m][torch.from_numpy(mine - rank * bl).cuda()].double()
                num[m], den[m] = ((got - ref) ** 2).sum(), (ref ** 2).sum()
            dist.all_reduce(num)
            dist.all_reduce(den)
            gerr = float((num / den).sqrt().max())
            ok &= lerr < 1e-3 and gerr < 1e-3
            print(f"[rank {rank}] C4 (B=8192, {bl} rows/rank) {name} vs reference golden: loss rel err {lerr:.2e} "
                  f"grad rel err {gerr:.2e}", flush=True)

    # ---- gallery-sharded retrieval vs the unsharded single-GPU path and the oracle
    tuples = RO.make_val_shaped(seed=3, n_shapes=1486, n_queries=3000, dim=512, round_bf16=True)
    text, gal, labels, *_ = RO.build_matrices(tuples)
    per = (gal.shape[0] + world - 1) // world
    lo, hi = rank * per, min((rank + 1) * per, gal.shape[0])
    t = torch.from_numpy(text).float().cuda()
    lab = torch.from_numpy(labels).cuda()
    v, i, r = sharded_retrieve(t, torch.from_numpy(gal[lo:hi]).cuda(), lab, lo, 5, block_queries=1024)
    v1, i1, r1 = retrieve(t, torch.from_numpy(gal).cuda(), lab, 5)
    same = torch.equal(i, i1) and torch.equal(r, r1) and torch.equal(v, v1)
    ok &= same
    print(f"[rank {rank}] sharded retrieval == unsharded: {same}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("MULTIRANK", "OK" if int(flag.item()) == 1 else "FAIL", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
