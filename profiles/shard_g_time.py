"""Sharded shared-G backward of ONE rank of a W-way sharded global batch replayed on one GPU (receive buffers are local
allocations, so the NVLink part of the stores is not in these numbers): per-kernel device time of kernel A (G row
block), kernel B (gradient GEMMs + drain into the receive slots) and the summing normalise backward, next to the
producer/consumer kernel on the same shapes.
   python profiles/shard_g_time.py [B] [W] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_features, TAU, ALPHA
from tricolo_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
W = int(sys.argv[2]) if len(sys.argv) > 2 else 8
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
F16 = 0
f = [v.cuda() for v in make_features(B, B, 0).values()]
pairs = [(0, 1), (0, 2), (1, 2)]
inv_tau = 1.0 / TAU
b_loc = B // W
zbuf = torch.empty((B, 3 * 512), dtype=torch.float16, device="cuda")
z_all, invs, xs = ops.l2norm_fwd(f, F16, out=[zbuf.view(B, 3, 512)[:, m] for m in range(3)])
fw = [ops.ntxent_fwd([z_all[a][r * b_loc:(r + 1) * b_loc] for a, _ in pairs], [z_all[b] for _, b in pairs], r * b_loc, inv_tau, F16)
      for r in range(W)]
col_sum = sum(x[1] for x in fw)
row_sum = torch.cat([x[0] for x in fw], dim=1).contiguous()
diag = torch.cat([x[2] for x in fw], dim=1).contiguous()
lse2_row_all, lse2_col, _, _ = ops.ntxent_finalize(row_sum, col_sum, diag, 0, inv_tau, ALPHA)
ones = torch.ones((3,), dtype=torch.float32, device="cuda")
plan = ops.ShardedBwdPlan(3, pairs, (1, 1, 1), b_loc, W, 512)
recv = torch.zeros((plan.recv_bytes,), dtype=torch.uint8, device="cuda")
work = torch.empty((plan.workspace_bytes,), dtype=torch.uint8, device="cuda")
addrs = [recv.data_ptr()] * W  # every "owner" is the same local buffer: same store traffic, no NVLink
inv_all = torch.stack(invs)[:, :b_loc].contiguous()
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")


def step():
    ops.ntxent_bwd_sharded_gemm(plan, z_all, 0, inv_tau, ALPHA, lse2_row_all, lse2_col, ones, work, addrs, F16)
    ops.ntxent_bwd_sharded_finish(plan, [x[:b_loc] for x in xs], inv_all, 0, work, addrs[0])


for _ in range(3):
    step()
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(iters):
    flush.zero_()
    step()
torch.cuda.synchronize()
prof = _lib.profile_read()
_lib.profile_enable(False)
res = {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]}
fl_g, fl_b = 2.0 * b_loc * B * 512 * 3, 4.0 * b_loc * B * 512 * 3
print(f"sharded-G  B={B} W={W} gsplit={os.environ.get('TRICOLO_B200_GSPLIT', 'auto')}", res,
      "G %.0f TF/s, GEMM %.0f TF/s" % (fl_g / res["ntxent_g"] / 1e9, fl_b / res["ntxent_bwd"] / 1e9))
# the producer/consumer kernel on the same rank's work
sl = slice(0, b_loc)
jobs = []
for m in range(3):
    segs = []
    for p, (a, b) in enumerate(pairs):
        if m == a:
            segs.append(ops.BwdSegmentSpec(z_all[b], None, lse2_row_all[p, sl], lse2_col[p], ones[p:p + 1], ALPHA, 1.0 - ALPHA))
        elif m == b:
            segs.append(ops.BwdSegmentSpec(z_all[a], None, lse2_col[p, sl], lse2_row_all[p], ones[p:p + 1], 1.0 - ALPHA, ALPHA))
    jobs.append(ops.BwdJobSpec(z_all[m][sl], xs[m][sl], invs[m][sl], segs))
for _ in range(3):
    ops.ntxent_bwd(jobs, B, 0, 0, inv_tau, F16)
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(iters):
    flush.zero_()
    ops.ntxent_bwd(jobs, B, 0, 0, inv_tau, F16)
torch.cuda.synchronize()
prof = _lib.profile_read()
_lib.profile_enable(False)
print("producer/consumer", {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]})
