"""Wait-time accounting of the CTA-pair gradient GEMM kernel (first and last CTA) on the single-GPU step, B = 8192.
   make trace && TRICOLO_B200_LIB=tricolo_b200/lib/libtricolo_b200_trace.so python profiles/gb2_trace.py [B] [iters]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_features, TAU, ALPHA
from tricolo_b200 import _lib
from tricolo_b200.loss import trimodal_ntxent

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
f = [v.cuda().requires_grad_(True) for v in make_features(B, B, 0).values()]
def step():
    for x in f: x.grad = None
    trimodal_ntxent(f, TAU, ALPHA).sum().backward()
for _ in range(3): step()
buf = (C.c_uint64 * 64)()
_lib.check(_lib.LIB.tcl_debug_gb_trace(buf, 1))
for _ in range(iters): step()
_lib.check(_lib.LIB.tcl_debug_gb_trace(buf, 1))
names = ["tma wait g_empty", "tma wait c_empty", "tma total", "mma wait acc_empty", "mma wait g_full", "mma wait c_full",
         "mma total", "drain wait acc_full", "drain work", "drain total", "tiles", "pieces"]
for base, who in ((0, "CTA 0 (leader of the first pair)"), (32, "last CTA (peer of the last pair)")):
    print(who, f"B={B} ggemm={os.environ.get('TRICOLO_B200_GGEMM', '2sm')}")
    for i, n in enumerate(names):
        print(f"  {n:22s} {int(buf[base + i]) / iters:10.0f}")
