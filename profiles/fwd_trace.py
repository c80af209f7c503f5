"""Wait-time accounting of the CTA-pair forward kernel (first cluster only).
   make trace && TRICOLO_B200_LIB=tricolo_b200/lib/libtricolo_b200_trace.so python profiles/fwd_trace.py [B]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_features, TAU, ALPHA
from tricolo_b200 import _lib
from tricolo_b200.loss import trimodal_ntxent

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = torch.device("cuda", 0)
feats = [v.to(dev).requires_grad_(True) for v in make_features(B, B, 0).values()]
buf = (C.c_uint64 * 32)()
for it in range(3):
    with torch.no_grad():
        trimodal_ntxent(feats, TAU, ALPHA)
    if it == 1:
        _lib.check(_lib.LIB.tcl_debug_fwd_trace(buf, 1))
_lib.check(_lib.LIB.tcl_debug_fwd_trace(buf, 1))
names = {0: "mma x_full", 1: "mma tmem_empty", 2: "mma full", 3: "mma total", 5: "tma empty (leader)", 6: "tma empty (peer)",
         7: "epi tmem_full (leader)", 8: "epi loads+math", 9: "epi column sums", 10: "epi total (leader)",
         11: "epi tmem_full (peer)", 12: "epi prologue", 13: "epi total (peer)"}
tiles = max(int(buf[4]), 1)
print("tiles", tiles)
for i, n in sorted(names.items()):
    print(f"{n:24s} {int(buf[i]):12d}  per tile {int(buf[i]) / tiles:9.1f}")
