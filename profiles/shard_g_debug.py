"""Numerics of the sharded shared-G backward replayed on one GPU against the fp64 oracle and the unsharded kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import ntxent_oracle as NO
from tricolo_b200 import ops
from tricolo_b200.loss import trimodal_ntxent
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
TAU, ALPHA = 0.1, 0.25
b_glob, world = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator().manual_seed(33)
base = torch.randn(b_glob, 512, generator=g)
fh = [(base + 0.5 * torch.randn(b_glob, 512, generator=g)).bfloat16().float() for _ in range(3)]
f = [x.cuda() for x in fh]
dev = [x.clone().requires_grad_(True) for x in f]
scales = torch.tensor([1.0, 0.5, 2.0], device="cuda")
losses = trimodal_ntxent(dev, TAU, ALPHA)
(losses * scales).sum().backward()
pairs = [(0, 1), (0, 2), (1, 2)]
b_loc = b_glob // world
zbuf = torch.empty((b_glob, 3 * 512), dtype=torch.float16, device="cuda")
z_all, invs, xs = ops.l2norm_fwd(f, 0, out=[zbuf.view(b_glob, 3, 512)[:, m] for m in range(3)])
fw = [ops.ntxent_fwd([z_all[a][r * b_loc:(r + 1) * b_loc] for a, _ in pairs], [z_all[b] for _, b in pairs], r * b_loc, 1 / TAU, 0) for r in range(world)]
col_sum = sum(x[1] for x in fw)
fin = [ops.ntxent_finalize(fw[r][0], col_sum, fw[r][2], r * b_loc, 1 / TAU, ALPHA, want_loss=False) for r in range(world)]
lse2_row_all = torch.cat([x[0] for x in fin], dim=1).contiguous()
lse2_col = fin[0][1].contiguous()
plan = ops.ShardedBwdPlan(3, pairs, (1, 1, 1), b_loc, world, 512)
recv = [torch.full((plan.recv_bytes,), 0xFF, dtype=torch.uint8, device="cuda") for _ in range(world)]
wss = [torch.empty((plan.workspace_bytes,), dtype=torch.uint8, device="cuda") for _ in range(world)]
addrs = [r.data_ptr() for r in recv]
for r in range(world):
    ops.ntxent_bwd_sharded_gemm(plan, z_all, r, 1 / TAU, ALPHA, lse2_row_all, lse2_col, scales, wss[r], addrs)
inv_all = torch.stack(invs)
# fp64 oracle with the same upstream scales: gradient of sum_p scale_p loss_p
keys = ["text_features", "image_features", "voxel_features"]
t = [x.clone().double().requires_grad_(True) for x in fh]
out = NO.torch_trimodal(dict(zip(keys, t)), TAU, ALPHA)
names = [f"train_loss/{a[:-9]}_{b[:-9]}_loss" for a, b in [(keys[0], keys[1]), (keys[0], keys[2]), (keys[1], keys[2])]]
sum(s * out[n] for s, n in zip([1.0, 0.5, 2.0], names)).backward()
for r in range(world):
    sl = slice(r * b_loc, (r + 1) * b_loc)
    dxs = ops.ntxent_bwd_sharded_finish(plan, [x[sl] for x in xs], inv_all[:, sl].contiguous(), r, wss[r], addrs[r])
    for m in range(3):
        ref = dev[m].grad[sl]
        orc = t[m].grad[sl].cuda()
        d = (dxs[m] - ref)
        e_sh = float((dxs[m].double() - orc).norm() / orc.norm())
        e_un = float((ref.double() - orc).norm() / orc.norm())
        viol = (d.abs() > 1e-3 * ref.abs() + 1e-4 * ref.abs().max()).sum().item()
        print(f"rank {r} tensor {m}: sharded-vs-unsharded rel {float(d.norm() / ref.norm()):.2e} max|d|/max|ref| {float(d.abs().max() / ref.abs().max()):.2e} "
              f"violations {viol}; vs fp64 oracle: sharded {e_sh:.2e} unsharded {e_un:.2e} nan {int(torch.isnan(dxs[m]).sum())}")
