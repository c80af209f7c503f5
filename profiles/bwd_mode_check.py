"""Parity + timing of one implementation of the NT-Xent backward (TRICOLO_B200_BWD=pc|pair|cluster|indep).
   TRICOLO_B200_BWD=pc python profiles/bwd_mode_check.py [B_time] [iters]
The mode is read once by the library, so every mode needs its own process."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from bench import make_features, TAU, ALPHA
from oracle import ntxent_oracle as NO
from tricolo_b200 import _lib
from tricolo_b200.loss import NTXentLoss, calculate_losses, trimodal_ntxent

mode = os.environ.get("TRICOLO_B200_BWD", "default")
B_time = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dev = torch.device("cuda", 0)
worst = 0.0
for (B, D, seed) in [(300, 512, 1), (1024, 512, 2), (200, 320, 3), (128, 448, 4), (640, 384, 5)]:
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(B, D, generator=g)
    feats = {k: (base + 0.7 * torch.randn(B, D, generator=g)).bfloat16().float()
             for k in ("text_features", "image_features", "voxel_features")}
    d = {k: v.to(dev).requires_grad_(True) for k, v in feats.items()}
    losses = calculate_losses(d, "l", NTXentLoss(TAU, ALPHA))
    losses["l/total_loss"].backward()
    ref_l, ref_g = NO.trimodal_forward_backward({k: v.numpy() for k, v in feats.items()}, TAU, ALPHA)
    errs = {k[:1]: float(np.linalg.norm(v.grad.double().cpu().numpy() - ref_g[k]) / np.linalg.norm(ref_g[k])) for k, v in d.items()}
    worst = max(worst, *errs.values())
    print(f"[{mode}] B={B} D={D} grad rel err {errs}", flush=True)
print(f"[{mode}] worst {worst:.3e} {'OK' if worst < 1e-3 else 'FAIL'}", flush=True)

feats = [v.to(dev).requires_grad_(True) for v in make_features(B_time, B_time, 0).values()]


def step():
    for f in feats:
        f.grad = None
    losses = trimodal_ntxent(feats, TAU, ALPHA)
    losses.sum().backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
_lib.profile_enable(True)
for _ in range(iters):
    step()
torch.cuda.synchronize()
prof = _lib.profile_read()
_lib.profile_enable(False)
print(f"[{mode}] B={B_time} ms/launch", {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if v[1]}, flush=True)
