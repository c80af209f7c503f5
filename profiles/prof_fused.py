"""ncu driver: one fused retrieval launch (18944 queries x 200k gallery = one CTA per SM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tricolo_b200.evaluation import retrieve
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(0)
gal = torch.randn(200000, 512, generator=g, device=dev).bfloat16()
text = torch.randn(148 * 128, 512, generator=g, device=dev).bfloat16()
lab = torch.randint(0, 200000, (148 * 128,), generator=g, device=dev)
for _ in range(2):
    retrieve(text, gal, lab, 5, fused=True)
torch.cuda.synchronize()
