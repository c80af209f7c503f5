"""Wait-time accounting of the CTA-pair backward kernel (first cluster only).
   make trace && TRICOLO_B200_LIB=tricolo_b200/lib/libtricolo_b200_trace.so python profiles/pair_trace.py [B]"""
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
        _lib.check(_lib.LIB.tcl_debug_pair_trace(buf, 1))  # drop the warm-up launches
_lib.check(_lib.LIB.tcl_debug_pair_trace(buf, 1))
names = ["prod empty-wait", "mma s_empty", "mma full(S)", "mma g_full", "mma full(A)", "mma total", "epi s_full",
         "epi g_empty", "epi tmem-ld", "epi compute+st", "epi fence+arrive", "epi total", "mma issue(S)", "mma commits",
         "mma issue(A)", "steps"]
for cta in range(2):
    print("CTA", cta, {n: int(buf[16 * cta + i]) for i, n in enumerate(names) if buf[16 * cta + i]})
