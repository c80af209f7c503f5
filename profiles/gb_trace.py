"""Wait-time accounting of the shared-G gradient GEMM kernel (first and last CTA) on one rank's share of a W-way
sharded batch replayed on one GPU.
   make trace && TRICOLO_B200_LIB=tricolo_b200/lib/libtricolo_b200_trace.so python profiles/gb_trace.py [B] [W] [iters]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_features, TAU, ALPHA
from tricolo_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
W = int(sys.argv[2]) if len(sys.argv) > 2 else 8
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
f = [v.cuda() for v in make_features(B, B, 0).values()]
pairs = [(0, 1), (0, 2), (1, 2)]
b_loc = B // W
zbuf = torch.empty((B, 3 * 512), dtype=torch.float16, device="cuda")
z_all, invs, xs = ops.l2norm_fwd(f, 0, out=[zbuf.view(B, 3, 512)[:, m] for m in range(3)])
lse = torch.full((3, B), 14.0, device="cuda")  # any plausible LSE: the trace does not depend on the values
ones = torch.ones((3,), device="cuda")
plan = ops.ShardedBwdPlan(3, pairs, (1, 1, 1), b_loc, W, 512)
recv = torch.zeros((plan.recv_bytes,), dtype=torch.uint8, device="cuda")
work = torch.empty((plan.workspace_bytes,), dtype=torch.uint8, device="cuda")
addrs = [recv.data_ptr()] * W
for _ in range(3):
    ops.ntxent_bwd_sharded_gemm(plan, z_all, 0, 1.0 / TAU, ALPHA, lse, lse, ones, work, addrs, 0)
buf = (C.c_uint64 * 64)()
_lib.check(_lib.LIB.tcl_debug_gb_trace(buf, 1))
for _ in range(iters):
    ops.ntxent_bwd_sharded_gemm(plan, z_all, 0, 1.0 / TAU, ALPHA, lse, lse, ones, work, addrs, 0)
_lib.check(_lib.LIB.tcl_debug_gb_trace(buf, 1))
names = ["tma wait g_empty", "tma wait c_empty", "tma total", "mma wait acc_empty", "mma wait g_full", "mma wait c_full",
         "mma total", "drain wait acc_full", "drain work", "drain total", "tiles", "pieces"]
for base, who in ((0, "CTA 0 (column-side units)"), (32, "last CTA (row-side units)")):
    print(who, f"B={B} W={W} gsplit={os.environ.get('TRICOLO_B200_GSPLIT', 'auto')}")
    for i, n in enumerate(names):
        print(f"  {n:22s} {int(buf[base + i]) / iters:10.0f}")
