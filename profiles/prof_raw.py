"""Driver for ncu captures of the kernels added in the second half of round 2:
   ncu --set full -k regex:"fwd_reduce_finalize|raw_" python profiles/prof_raw.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import make_features, TAU, ALPHA
from tricolo_b200.loss import NTXentLoss, trimodal_ntxent_total

fn = NTXentLoss(TAU, ALPHA)
feats = [v.cuda().requires_grad_(True) for v in make_features(8192, 8192, 0).values()]
for _ in range(2):
    trimodal_ntxent_total(feats, TAU, ALPHA)[1].backward()      # fwd_reduce_finalize_kernel (B = 8192)
raw = [(0.05 * torch.randn(2048, 512, device="cuda")).requires_grad_(True) for _ in range(3)]
for _ in range(2):
    fn.fused_total(raw, norm=False)[1].backward()               # raw_fwd / raw_finalize / raw_bwd (B = 2048, trimodal)
torch.cuda.synchronize()
