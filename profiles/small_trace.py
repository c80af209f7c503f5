"""Phase timeline of the small-batch kernels (CTA 0, %globaltimer): run with the trace build,
   make trace && TRICOLO_B200_LIB=tricolo_b200/lib/libtricolo_b200_trace.so python profiles/small_trace.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tricolo_b200 import _lib as L  # noqa: E402
from tricolo_b200.loss import trimodal_ntxent  # noqa: E402

FWD = ["normalise -> smem", "S tile", "stats + partial stores", "grid barrier", "finalise (pair CTA)"]
BWD = ["z blocks -> smem", "S tile", "G tile", "two gradient GEMMs + stores", "grid barrier", "row phase"]
for name, b, n, d in (("C1 B=128 bi", 128, 2, 512), ("C2 B=256 tri", 256, 3, 512), ("B=256 tri dim 128", 256, 3, 128)):
    gen = torch.Generator().manual_seed(1)
    feats = [torch.randn(b, d, generator=gen).cuda().requires_grad_(True) for _ in range(n)]
    for _ in range(5):
        for f in feats:
            f.grad = None
        trimodal_ntxent(feats, 0.1, 0.25).sum().backward()
    buf = (C.c_uint64 * 32)()
    L.check(L.LIB.tcl_debug_small_trace(buf))
    t = list(buf)
    print(name)
    for i, lab in enumerate(FWD):
        print(f"  fwd {lab:32s} {(t[i + 1] - t[i]) / 1e3:7.2f} us")
    print(f"  fwd kernel body (CTA 0)          {(t[5] - t[0]) / 1e3:7.2f} us")
    for i, lab in enumerate(BWD):
        print(f"  bwd {lab:32s} {(t[17 + i] - t[16 + i]) / 1e3:7.2f} us")
    print(f"  bwd kernel body (CTA 0)          {(t[22] - t[16]) / 1e3:7.2f} us")
