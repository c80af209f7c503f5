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
    # ---- global-negative trimodal loss: B_global = 256 * world, rank r owns rows [256 r, 256 (r+1))
    b = 256 * world
    g = torch.Generator().manual_seed(7)
    base = torch.randn(b, 512, generator=g)
    full = [(base + 0.5 * torch.randn(b, 512, generator=g)).bfloat16().float() for _ in range(3)]
    bl = b // world
    loc = [f[rank * bl:(rank + 1) * bl].cuda().requires_grad_(True) for f in full]
    ref_l, ref_g = NO.trimodal_forward_backward(
        {"text_features": full[0].numpy(), "image_features": full[1].numpy(), "voxel_features": full[2].numpy()}, TAU, ALPHA)
    results = {}
    for transport in ("1", "0"):  # NVLink peer memory (symmetric memory) and NCCL collectives
        os.environ["TRICOLO_B200_SYMM"] = transport
        for x in loc:
            x.grad = None
        out = global_calculate_losses({"text_features": loc[0], "image_features": loc[1], "voxel_features": loc[2]},
                                      "train_loss", TAU, ALPHA)
        out["train_loss/total_loss"].backward()
        for k, v in ref_l.items():
            rel = abs(float(out[k].detach()) - v) / abs(v)
            ok &= rel < 1e-3
        errs = []
        for m, key in enumerate(["text_features", "image_features", "voxel_features"]):
            ref = ref_g[key][rank * bl:(rank + 1) * bl]
            errs.append(np.linalg.norm(loc[m].grad.double().cpu().numpy() - ref) / np.linalg.norm(ref))
        ok &= max(errs) < 1e-3
        results[transport] = (float(out["train_loss/total_loss"].detach()), [x.grad.clone() for x in loc])
        print(f"[rank {rank}] transport {'symm' if transport == '1' else 'nccl'}: loss total "
              f"{results[transport][0]:.6f} ref {ref_l['train_loss/total_loss']:.6f} grad errs {['%.2e' % e for e in errs]}", flush=True)
    same = abs(results["1"][0] - results["0"][0]) <= 1e-6 * abs(results["0"][0])
    same &= all(torch.allclose(a, b, rtol=1e-5, atol=1e-9) for a, b in zip(results["1"][1], results["0"][1]))
    ok &= same
    # two forwards before the two backwards: the second forward must not overwrite the operands the first backward needs
    os.environ["TRICOLO_B200_SYMM"] = "1"
    loc2 = [(x.detach() * 0.5 + 0.1).requires_grad_(True) for x in loc]
    for x in loc:
        x.grad = None
    o1 = global_calculate_losses({"text_features": loc[0], "image_features": loc[1], "voxel_features": loc[2]}, "a", TAU, ALPHA)
    o2 = global_calculate_losses({"text_features": loc2[0], "image_features": loc2[1], "voxel_features": loc2[2]}, "a", TAU, ALPHA)
    o2["a/total_loss"].backward()
    o1["a/total_loss"].backward()
    inter = all(torch.allclose(x.grad, g, rtol=1e-5, atol=1e-9) for x, g in zip(loc, results["1"][1]))
    ok &= inter
    print(f"[rank {rank}] symm == nccl: {same}; interleaved forwards keep their operands: {inter}", flush=True)

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
