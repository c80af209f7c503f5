"""Short drivers for ncu captures.
   ncu ... python profiles/prof_step.py loss [B]        three fwd+bwd steps of the trimodal loss
   ncu ... python profiles/prof_step.py retrieval       two-kernel form: GEMM -> HBM -> top-k (3 blocks of 8192 x 200k)
   ncu ... python profiles/prof_step.py fused           fused kernel, 18944 queries x 200k (one CTA per SM)
   ncu ... python profiles/prof_step.py shard           two-kernel form at one rank's shard of 8: 41984 queries x 25k"""
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
elif what == "shard":
    g = torch.Generator(device=dev).manual_seed(0)
    gal = torch.randn(25000, 512, generator=g, device=dev).bfloat16()
    text = torch.randn(41984, 512, generator=g, device=dev).bfloat16()
    lab = torch.randint(0, 25000, (41984,), generator=g, device=dev)
    retrieve(text, gal, lab, 5, fused=False)
else:
    g = torch.Generator(device=dev).manual_seed(0)
    gal = torch.randn(200000, 512, generator=g, device=dev).bfloat16()
    nq = 8192 * 3 if what == "retrieval" else 148 * 128
    text = torch.randn(nq, 512, generator=g, device=dev).bfloat16()
    lab = torch.randint(0, 200000, (nq,), generator=g, device=dev)
    for _ in range(1 if what == "retrieval" else 2):
        retrieve(text, gal, lab, 5, block_queries=8192 if what == "retrieval" else None, fused=(what == "fused"))
torch.cuda.synchronize()
