"""norm=False path (csrc/ntxent_raw.cu): time per fwd+bwd step, the reference's ops eager on the same GPU beside it.
   python profiles/raw_time.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tricolo_b200.loss import NTXentLoss

def ref(zis, zjs, tau, alpha):
    eye = torch.eye(zis.shape[0], device=zis.device)
    lab = zis @ zjs.t() / tau
    lba = zjs @ zis.t() / tau
    la = -(eye * torch.log_softmax(lab, 1)).sum() / lab.shape[0]
    lb = -(eye * torch.log_softmax(lba, 1)).sum() / lab.shape[0]
    return alpha * la + (1 - alpha) * lb

torch.backends.cuda.matmul.allow_tf32 = False
fn = NTXentLoss(0.1, 0.25)
for b, n_mod in ((128, 2), (256, 3), (2048, 3), (8192, 3)):
    xs = [(0.05 * torch.randn(b, 512, device="cuda")).requires_grad_(True) for _ in range(n_mod)]
    def ours():
        l, t = fn.fused_total(xs, norm=False)
        t.backward()
        return t
    def theirs():
        t = sum(ref(xs[i], xs[j], 0.1, 0.25) for i in range(n_mod) for j in range(i + 1, n_mod))
        t.backward()
        return t
    res = {}
    for name, f in (("ours", ours), ("torch_eager_fp32", theirs)):
        for _ in range(2):
            f()
        for x in xs:
            x.grad = None
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        n = 5 if b >= 2048 else 20
        e0.record()
        for _ in range(n):
            t = f()
        e1.record()
        torch.cuda.synchronize()
        res[name] = (e0.elapsed_time(e1) / n, float(t), [x.grad.clone() / (n + 0) for x in xs])
        for x in xs:
            x.grad = None
    err = max(float((a - c).norm() / c.norm()) for a, c in zip(res["ours"][2], res["torch_eager_fp32"][2]))
    print(f"B={b} mods={n_mod}: ours {res['ours'][0]:.3f} ms, torch eager fp32 {res['torch_eager_fp32'][0]:.3f} ms; "
          f"loss {res['ours'][1]:.6f} vs {res['torch_eager_fp32'][1]:.6f}; grad rel diff {err:.2e}")
