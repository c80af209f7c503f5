"""Seeded synthetic inputs shared by the CPU and GPU tests (same builders as
tests/golden/make_golden.py)."""
import torch


def _randn(seed, n, b, d):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(b, d, generator=g) for _ in range(n)]


def loss_case(name):
    if name == "G1":
        t, v = _randn(1234, 2, 128, 512)
        return {"text_features": t, "voxel_features": v}
    if name == "G2":
        t, i, v = _randn(1234, 3, 256, 512)
        return {"text_features": t, "image_features": i, "voxel_features": v}
    if name == "G3":
        g = torch.Generator().manual_seed(7)
        base = torch.randn(256, 512, generator=g)
        t = base + 0.5 * torch.randn(256, 512, generator=g)
        i = base + 0.5 * torch.randn(256, 512, generator=g)
        v = base + 0.5 * torch.randn(256, 512, generator=g)
        return {"text_features": t, "image_features": i, "voxel_features": v}
    if name == "KAT1":
        e = torch.eye(128, 512)
        return {"text_features": e.clone(), "voxel_features": e.clone()}
    if name == "KAT2":
        o = torch.ones(128, 512)
        return {"text_features": o.clone(), "voxel_features": o.clone()}
    if name == "KAT3":
        t, v = _randn(1, 2, 64, 512)
        t[0] = 0
        return {"text_features": t, "voxel_features": v}
    if name == "RAGGED200":
        t, i, v = _randn(99, 3, 200, 512)
        return {"text_features": t, "image_features": i, "voxel_features": v}
    if name == "DIM256":
        t, v = _randn(5, 2, 384, 256)
        return {"text_features": t, "image_features": v}
    raise KeyError(name)


LOSS_CASES = ["G1", "G2", "G3", "KAT1", "KAT2", "KAT3", "RAGGED200", "DIM256"]
GRAD_CASES = ["G1", "G3", "KAT3", "RAGGED200"]  # full gradients (every 4th row) stored in loss_grads.npz


def bf16_rounded(feats):
    return {k: v.bfloat16().float() for k, v in feats.items()}


# norm=False cases (nt_xent.py:55): unnormalised embeddings, i.e. logits of magnitude 1e1..1e3.  Inputs are rounded to
# bf16 so that every supported input dtype carries them exactly.  Goldens: tests/golden/make_golden_raw.py.
RAW_CASES = ["R1", "R2", "R3", "R4"]


def raw_case(name):
    if name == "R1":  # Bi(V), logits ~ N(0, 20)
        t, v = _randn(11, 2, 128, 512)
        f = {"text_features": 0.3 * t, "voxel_features": 0.3 * v}
    elif name == "R2":  # trimodal, ragged batch, correlated modalities
        g = torch.Generator().manual_seed(12)
        base = torch.randn(200, 512, generator=g)
        f = {k: 0.03 * (base + torch.randn(200, 512, generator=g))
             for k in ("text_features", "image_features", "voxel_features")}
    elif name == "R3":  # large logits (std ~ 80): near one-hot softmax, needs the online maximum
        t, i = _randn(13, 2, 256, 64)
        f = {"text_features": t, "image_features": i}
    elif name == "R4":  # rows of very different norms; batch of one tile plus one row
        g = torch.Generator().manual_seed(14)
        t = torch.randn(65, 128, generator=g) * torch.logspace(-2, 0.5, 65).unsqueeze(1)
        v = torch.randn(65, 128, generator=g)
        f = {"text_features": t, "voxel_features": v}
    else:
        raise KeyError(name)
    return bf16_rounded(f)
