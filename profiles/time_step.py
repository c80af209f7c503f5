"""Per-kernel device times of the trimodal loss step (library event profiler).
   python profiles/time_step.py [B] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_features, TAU, ALPHA
from tricolo_b200 import _lib
from tricolo_b200.loss import trimodal_ntxent

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
feats = [v.to(dev).requires_grad_(True) for v in make_features(B, B, 0).values()]


def step():
    for f in feats:
        f.grad = None
    losses = trimodal_ntxent(feats, TAU, ALPHA)
    losses.sum().backward()
    return losses


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    step()
e1.record()
torch.cuda.synchronize()
print("eager ms/step", e0.elapsed_time(e1) / iters)
_lib.profile_enable(True)
for _ in range(iters):
    step()
torch.cuda.synchronize()
prof = _lib.profile_read()
_lib.profile_enable(False)
print({k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]})
