"""Wait-time accounting of the producer/consumer backward kernel (first cluster only).
   make trace && TRICOLO_B200_LIB=tricolo_b200/lib/libtricolo_b200_trace.so python profiles/pc_trace.py [B]"""
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
    for f in feats:
        f.grad = None
    trimodal_ntxent(feats, TAU, ALPHA).sum().backward()
    if it == 1:
        _lib.check(_lib.LIB.tcl_debug_pc_trace(buf, 1))  # drop the warm-up launches
_lib.check(_lib.LIB.tcl_debug_pc_trace(buf, 1))
names = {0: "P-tma p_empty", 1: "P-mma s_empty", 2: "P-mma p_full", 3: "P-mma total", 4: "P-epi s_full", 5: "P-epi tmem-ld",
         6: "P-epi math", 7: "P-epi g_empty", 8: "P-epi stores", 9: "P-epi fence+arrive", 10: "P-epi total",
         11: "P-epi bar.sync", 16: "C-tma c_empty", 17: "C-mma g_full", 18: "C-mma c_full", 19: "C-mma total",
         20: "C-epi read-out", 12: "P-mma x_full", 13: "P-epi piece prologue", 21: "C-mma acc_empty", 22: "C-epi acc_full",
         30: "pieces", 31: "tiles"}
tiles = max(int(buf[31]), 1)
print("tiles", tiles)
print("pieces", int(buf[30]))
for i, n in sorted(names.items()):
    if i < 30:
        print(f"{n:22s} {int(buf[i]):12d}  per tile {int(buf[i]) / tiles:9.1f}")
