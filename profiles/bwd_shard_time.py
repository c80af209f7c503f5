"""Backward of ONE rank of a W-way sharded global batch, replayed on one GPU (n_self = B/W local rows against all B
rows): kernel time and, with the trace build, the wait-time accounting of the first cluster.
   python profiles/bwd_shard_time.py [B] [W] [iters]"""
import ctypes as C, os, sys
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
buf = (C.c_uint64 * 32)()
if hasattr(_lib.LIB, "tcl_debug_pc_trace"):
    _lib.check(_lib.LIB.tcl_debug_pc_trace(buf, 1))
_lib.profile_enable(True)
for _ in range(iters):
    ops.ntxent_bwd(jobs, B, 0, 0, inv_tau, F16)
torch.cuda.synchronize()
prof = _lib.profile_read()
_lib.profile_enable(False)
print({k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]})
_lib.check(_lib.LIB.tcl_debug_pc_trace(buf, 1))
if buf[31]:
    names = {0: "P-tma p_empty", 1: "P-mma s_empty", 2: "P-mma p_full", 3: "P-mma total", 12: "P-mma x_full", 4: "P-epi s_full",
             6: "P-epi math", 7: "P-epi g_empty", 10: "P-epi total", 13: "P-epi piece prologue", 17: "C-mma g_full",
             18: "C-mma c_full", 19: "C-mma total", 21: "C-mma acc_empty", 20: "C-epi read-out", 22: "C-epi acc_full"}
    tiles, pieces = int(buf[31]) / iters, int(buf[30]) / iters
    print("per launch: tiles", tiles, "pieces", pieces)
    for i, n in sorted(names.items()):
        print(f"{n:22s} per launch {int(buf[i]) / iters:10.0f}  per tile {int(buf[i]) / iters / max(tiles, 1):8.1f}")
