"""CPU stand-ins for tricolo_b200.ops with the same contracts (include/tricolo_b200.h), built on the
oracle's formulas.  TEST INFRASTRUCTURE: lets the multi-process host logic of
tricolo_b200/distributed.py (sharding offsets, pair -> segment mapping, collectives, merge) run
under gloo on a CPU-only box.  Never imported by the product package."""
import math

import numpy as np
import torch

F16, BF16 = 0, 1
_DT = {F16: torch.float16, BF16: torch.bfloat16}
LOG2E = 1.4426950408889634


class _L:
    @staticmethod
    def op_torch_dtype(op):
        return _DT[op]


L = _L()


class BwdSegmentSpec:
    def __init__(self, z_other, z_other_t, lse2_self, lse2_other, grad_scale, w_self, w_other):
        self.__dict__.update(locals())


class BwdJobSpec:
    def __init__(self, z_self, x_self, inv_norm, segments):
        self.__dict__.update(locals())


def l2norm_fwd(xs, op_format=F16, eps=1e-12, out=None):
    zs, invs = [], []
    for m, x in enumerate(xs):
        inv = 1.0 / x.float().norm(dim=1).clamp_min(eps)
        z = (x.float() * inv[:, None]).to(_DT[op_format])
        if out is not None:
            out[m].copy_(z)
            z = out[m]
        zs.append(z)
        invs.append(inv)
    return zs, invs, list(xs)


def cast_16bit(x, op_format=BF16):
    return x.to(_DT[op_format])


def transpose_16bit(zs):
    rows = zs[0].shape[0]
    ld = (rows + 7) // 8 * 8
    out = []
    for z in zs:
        t = torch.zeros((z.shape[1], ld), dtype=z.dtype)
        t[:, :rows] = z.t()
        out.append(t)
    return out, ld


def transpose_for_bwd(zs):
    """The mock always asks for the transposed copies (the contract of the kernels for dim <= 256)."""
    return transpose_16bit(zs)


def ntxent_fwd(zrows, zcols, row_offset, inv_tau, op_format=F16):
    c1 = inv_tau * LOG2E
    rs, cs, dg = [], [], []
    for zr, zc in zip(zrows, zcols):
        s = zr.double() @ zc.double().t()
        e = torch.exp2(c1 * s - c1)
        rs.append(e.sum(1))
        cs.append(e.sum(0))
        dg.append(c1 * torch.diagonal(s, offset=row_offset))
    return torch.stack(rs).float(), torch.stack(cs).float(), torch.stack(dg).float()


def ntxent_finalize(row_sum, col_sum, diag2, row_offset, inv_tau, alpha, want_loss=True):
    c1 = inv_tau * LOG2E
    n_rows, n_cols = row_sum.shape[1], col_sum.shape[1]
    lr = torch.log2(row_sum.double()) + c1
    lc = torch.log2(col_sum.double()) + c1
    ln2 = math.log(2.0)
    p0 = ln2 * (lr - diag2.double()).sum(1)
    p1 = ln2 * (lc[:, row_offset:row_offset + n_rows] - diag2.double()).sum(1)
    parts = torch.stack([p0, p1], dim=1).float()
    loss = ((alpha * p0 + (1 - alpha) * p1) / n_cols).float() if want_loss else None
    return lr.float(), lc.float(), parts, loss


def ntxent_bwd(jobs, n_other, self_offset, ld_t, inv_tau, op_format=F16, eps=1e-12):
    c1 = inv_tau * LOG2E
    out = []
    for job in jobs:
        n_self = job.z_self.shape[0]
        acc = torch.zeros(job.z_self.shape, dtype=torch.float64)
        for sg in job.segments:
            s = job.z_self.double() @ sg.z_other.double().t()
            g = (sg.w_self * torch.exp2(c1 * s - sg.lse2_self.double()[:, None])
                 + sg.w_other * torch.exp2(c1 * s - sg.lse2_other.double()[None, :]))
            g[torch.arange(n_self), self_offset + torch.arange(n_self)] -= 1.0
            go = 1.0 if sg.grad_scale is None else float(sg.grad_scale)
            # the transposed operand must describe the same matrix
            assert torch.equal(sg.z_other_t[:, :n_other].t().contiguous(), sg.z_other.contiguous())
            acc += go * (g @ sg.z_other_t[:, :n_other].t().double())
        gz = acc * inv_tau / n_other
        x = job.x_self.double()
        inv = job.inv_norm.double()[:, None]
        zh = x * inv
        proj = (gz * zh).sum(1, keepdim=True)
        proj = torch.where(inv >= 1.0 / eps, torch.zeros_like(proj), proj)
        out.append(((gz - proj * zh) * inv).to(job.x_self.dtype))
    return out


def sim_gemm(q16, g16, out=None):
    n_g = g16.shape[0]
    ld = (n_g + 3) // 4 * 4
    s = torch.zeros((q16.shape[0], ld), dtype=torch.float32)
    s[:, :n_g] = (q16.double() @ g16.double().t()).float()
    return s, n_g


def gather_gt_sim(s, n_g, labels, idx_base):
    loc = labels - idx_base
    ok = (loc >= 0) & (loc < n_g)
    v = torch.gather(s, 1, loc.clamp(0, n_g - 1)[:, None])[:, 0]
    return torch.where(ok, v, torch.zeros_like(v))


def topk_rank(s, n_g, k, labels, idx_base=0, gt_sim_in=None):
    sv = s[:, :n_g]
    order = torch.sort(sv, dim=1, descending=True, stable=True).indices[:, :k]
    val = torch.gather(sv, 1, order)
    gt = gt_sim_in if gt_sim_in is not None else gather_gt_sim(s, n_g, labels, idx_base)
    cols = idx_base + torch.arange(n_g)[None, :]
    nb = ((sv > gt[:, None]) | ((sv == gt[:, None]) & (cols < labels[:, None]))).sum(1).to(torch.int32)
    return val, (order + idx_base).to(torch.int32), gt, nb


def topk_merge(cand_val, cand_idx):
    w, q, k = cand_val.shape
    v = cand_val.permute(1, 0, 2).reshape(q, w * k).double().numpy()
    i = cand_idx.permute(1, 0, 2).reshape(q, w * k).numpy()
    order = np.lexsort((i, -v), axis=1)[:, :k]
    return (torch.from_numpy(np.take_along_axis(v, order, 1)).float(),
            torch.from_numpy(np.take_along_axis(i, order, 1)).to(torch.int32))


def gt_sim_mma(q16, g16, labels, idx_base=0):
    loc = labels - idx_base
    ok = (loc >= 0) & (loc < g16.shape[0])
    rows = g16[loc.clamp(0, g16.shape[0] - 1)]
    v = (q16.double() * rows.double()).sum(1).float()
    return torch.where(ok, v, torch.zeros_like(v))


def sim_topk_fused(q16, g16, k, labels, gt_sim, idx_base=0):
    s, n_g = sim_gemm(q16, g16)
    val, idx, _, nb = topk_rank(s, n_g, k, labels, idx_base, gt_sim)
    return val, idx, nb
