"""Where the host time of one EAGER small-batch training step goes (C2: Tri B=256): cProfile over 300 steps."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tricolo_b200.loss import trimodal_ntxent_total  # noqa: E402

gen = torch.Generator().manual_seed(1)
feats = [torch.randn(256, 512, generator=gen).cuda().requires_grad_(True) for _ in range(3)]


def step():
    for f in feats:
        f.grad = None
    trimodal_ntxent_total(feats, 0.1, 0.25)[1].backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(300):
    step()
torch.cuda.synchronize()
print(f"eager step: {(time.perf_counter() - t0) / 300 * 1e6:.1f} us wall")
pr = cProfile.Profile()
pr.enable()
for _ in range(300):
    step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
