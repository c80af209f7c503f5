"""Short driver for ncu captures: a few fwd+bwd steps of the C4 loss and one retrieval block.
   ncu --set full ... python profiles/prof_step.py [loss|retrieval]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_features, TAU, ALPHA
from tricolo_b200.loss import trimodal_ntxent
from tricolo_b200.evaluation import retrieve

what = sys.argv[1] if len(sys.argv) > 1 else "loss"
dev = torch.device("cuda", 0)
if what == "loss":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    feats = [v.to(dev).requires_grad_(True) for v in make_features(B, B, 0).values()]
    for _ in range(3):
        for f in feats: f.grad = None
        trimodal_ntxent(feats, TAU, ALPHA).sum().backward()
else:
    g = torch.Generator(device=dev).manual_seed(0)
    gal = torch.randn(200000, 512, generator=g, device=dev).bfloat16()
    text = torch.randn(8192 * 3, 512, generator=g, device=dev).bfloat16()
    lab = torch.randint(0, 200000, (8192 * 3,), generator=g, device=dev)
    retrieve(text, gal, lab, 5, block_queries=8192)
torch.cuda.synchronize()
