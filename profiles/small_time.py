"""Small-batch loss: CUDA-graph replay time of the forward alone and of forward + backward, single-launch form
(TRICOLO_B200_SMALL=1, csrc/ntxent_small.cu) against the multi-kernel pipeline (=0).  C1 = Bi(V) B=128, C2 = Tri B=256."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tricolo_b200.loss import trimodal_ntxent, trimodal_ntxent_total  # noqa: E402


def graph_time(fn, iters=200):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(10):
        g.replay()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(iters):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


for name, b, n in (("C1 B=128 bi", 128, 2), ("C2 B=256 tri", 256, 3), ("B=512 bi", 512, 2), ("B=64 tri", 64, 3)):
    gen = torch.Generator().manual_seed(1)
    feats = [torch.randn(b, 512, generator=gen).cuda().requires_grad_(True) for _ in range(n)]
    for form in ("1", "0"):
        os.environ["TRICOLO_B200_SMALL"] = form

        def fwd():
            with torch.no_grad():
                return trimodal_ntxent(feats, 0.1, 0.25)

        def both():
            for f in feats:
                f.grad = None
            trimodal_ntxent(feats, 0.1, 0.25).sum().backward()

        def step():  # the training step: total_loss.backward()
            for f in feats:
                f.grad = None
            trimodal_ntxent_total(feats, 0.1, 0.25)[1].backward()

        t = [min(graph_time(fn) for _ in range(3)) for fn in (fwd, both, step)]
        print(f"{name} small={form}: fwd {t[0]:.1f} us, fwd + .sum().backward() {t[1]:.1f} us, total.backward() {t[2]:.1f} us", flush=True)

