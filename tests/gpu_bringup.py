"""Stage-by-stage hardware bring-up of the sm_100a kernels (diagnostics, not a pytest module).

    python tests/gpu_bringup.py <stage>        stage in: probe elementwise gemm topk fwd bwd eval all
Each stage runs in its own process under `timeout` when driven by `all`, so a trap in one
kernel does not poison the CUDA context of the next stage.
"""
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["probe", "elementwise", "gemm", "topk", "fwd", "bwd", "eval"]


def stage_probe():
    from tricolo_b200 import ops
    out = ops.debug_tmem_probe().cpu().numpy()  # [warp][half][thread][reg]
    ok = True
    bad = []
    for w in range(4):
        for h in range(2):
            for t in range(32):
                q, p = t // 4, t % 4
                for g in range(4):
                    for e in range(2):
                        col = 8 * g + 2 * p + e
                        for half_row, reg in ((0, 4 * g + e), (8, 4 * g + 2 + e)):
                            lane = w * 32 + h * 16 + q + half_row
                            exp = lane * 64 + col
                            got = int(out[w, h, t, reg])
                            if got != exp:
                                ok = False
                                if len(bad) < 12:
                                    bad.append((w, h, t, reg, got // 64, got % 64, lane, col))
    print("probe 16x256b layout as assumed:", ok)
    if not ok:
        print("  (warp, half, thread, reg, got_lane, got_col, exp_lane, exp_col):")
        for b in bad:
            print("  ", b)
        print("  raw warp0 half0 thread0..3:", out[0, 0, :4].tolist())
    return ok


def stage_elementwise():
    from tricolo_b200 import ops
    ok = True
    for dt in (torch.float32, torch.bfloat16, torch.float16, torch.float64):
        for rows, dim in ((128, 512), (200, 256), (1000, 64)):
            g = torch.Generator().manual_seed(rows + dim)
            x = torch.randn(rows, dim, generator=g).to(dt).cuda()
            x[3] = 0
            for op, tdt in ((ops.F16, torch.float16), (ops.BF16, torch.bfloat16)):
                (z,), (inv,), _ = ops.l2norm_fwd([x], op)
                n = x.double().norm(dim=1).clamp_min(1e-12)
                zr = (x.double() / n[:, None])
                e1 = (z.double() - zr.to(tdt).double()).abs().max().item()
                e2 = ((inv.double() - 1 / n).abs() / (1 / n)).max().item()
                y = ops.cast_16bit(x, op)
                e3 = (y.double() - x.double().to(tdt).double()).abs().max().item()
                (zt,), ld = ops.transpose_16bit([z])
                e4 = (zt[:, :rows] != z.t()).sum().item()
                good = e1 <= 2 ** -8 and e2 < 1e-5 and e3 <= 2 ** -7 * x.abs().max().item() and e4 == 0
                ok &= good
                if not good:
                    print("  FAIL", dt, rows, dim, op, e1, e2, e3, e4)
    print("elementwise ok:", ok)
    return ok


def stage_gemm():
    from tricolo_b200 import ops
    ok = True
    for (m, n, k) in ((128, 128, 64), (128, 128, 512), (256, 384, 512), (200, 1486, 512), (1000, 300, 136), (7424, 1486, 512)):
        for tdt in (torch.bfloat16, torch.float16):
            g = torch.Generator().manual_seed(m + n + k)
            a = torch.randn(m, k, generator=g).to(tdt).cuda()
            b = torch.randn(n, k, generator=g).to(tdt).cuda()
            s, n_g = ops.sim_gemm(a, b)
            torch.cuda.synchronize()
            ref = a.double() @ b.double().t()
            err = (s[:, :n_g].double() - ref).abs().max().item()
            # integer operands: exact
            ai = torch.randint(-2, 3, (m, k), generator=g).to(tdt).cuda()
            bi = torch.randint(-2, 3, (n, k), generator=g).to(tdt).cuda()
            si, _ = ops.sim_gemm(ai, bi)
            exact = torch.equal(si[:, :n].double(), ai.double() @ bi.double().t())
            good = err < 2e-3 * (k ** 0.5) / 22 and exact
            ok &= good
            print(f"  gemm {m}x{n}x{k} {tdt}: max err {err:.3e} integer-exact {exact} {'ok' if good else 'FAIL'}")
            if not good:
                d = (s[:, :n_g].double() - ref).abs()
                bad = (d > 1e-2).nonzero()
                print("   bad count", bad.shape[0], "first", bad[:8].tolist())
                print("   s[0,:8]", s[0, :8].tolist(), "ref", ref[0, :8].tolist())
    print("gemm ok:", ok)
    return ok


def ref_topk(s, labels, k):
    order = torch.sort(s, dim=1, descending=True, stable=True).indices
    idx = order[:, :k]
    val = torch.gather(s, 1, idx)
    sgt = torch.gather(s, 1, labels[:, None])
    cols = torch.arange(s.shape[1], device=s.device)[None, :]
    nb = ((s > sgt) | ((s == sgt) & (cols < labels[:, None]))).sum(dim=1)
    return val, idx, nb


def stage_topk():
    from tricolo_b200 import ops
    ok = True
    for (q, g, k, ties) in ((64, 40, 5, True), (1000, 1486, 5, False), (777, 1486, 5, True), (300, 4099, 16, False), (50, 3, 5, True),
                            (2000, 25000, 5, False)):
        gen = torch.Generator().manual_seed(q + g)
        if ties:
            s = torch.randint(-3, 4, (q, g), generator=gen).float()
        else:
            s = torch.randn(q, g, generator=gen)
        ld = (g + 3) // 4 * 4
        sp = torch.full((q, ld), float("nan"))
        sp[:, :g] = s
        labels = torch.randint(0, g, (q,), generator=gen)
        sp, s, labels = sp.cuda(), s.cuda(), labels.cuda()
        val, idx, gt, nb = ops.topk_rank(sp, g, k, labels)
        kk = min(k, g)
        rv, ri, rnb = ref_topk(s, labels, kk)
        good = torch.equal(idx[:, :kk].long(), ri) and torch.equal(val[:, :kk], rv) and torch.equal(nb.long(), rnb)
        good &= torch.equal(gt, torch.gather(s, 1, labels[:, None])[:, 0])
        # sharded: 3 shards + merge
        bounds = [0, g // 3, 2 * g // 3, g]
        if g >= 12:
            gts = torch.zeros(q, device="cuda")
            for i in range(3):
                lo, hi = bounds[i], bounds[i + 1]
                shard = torch.zeros((q, (hi - lo + 3) // 4 * 4), device="cuda")
                shard[:, : hi - lo] = s[:, lo:hi]
                gts += ops.gather_gt_sim(shard, hi - lo, labels, lo)
            cv, ci, nbs = [], [], torch.zeros(q, dtype=torch.int32, device="cuda")
            for i in range(3):
                lo, hi = bounds[i], bounds[i + 1]
                shard = torch.zeros((q, (hi - lo + 3) // 4 * 4), device="cuda")
                shard[:, : hi - lo] = s[:, lo:hi]
                v, ix, _, nb_i = ops.topk_rank(shard, hi - lo, k, labels, lo, gts)
                cv.append(v); ci.append(ix); nbs += nb_i
            mv, mi = ops.topk_merge(torch.stack(cv), torch.stack(ci))
            good &= torch.equal(mi[:, :kk].long(), ri) and torch.equal(mv[:, :kk], rv) and torch.equal(nbs.long(), rnb)
        ok &= bool(good)
        print(f"  topk q={q} g={g} k={k} ties={ties}: {'ok' if good else 'FAIL'}")
        if not good:
            bad = (idx[:, :kk].long() != ri).any(dim=1).nonzero()[:3, 0].tolist()
            for b in bad:
                print("   row", b, "got", idx[b].tolist(), val[b].tolist(), "ref", ri[b].tolist(), rv[b].tolist())
            badn = (nb.long() != rnb).nonzero()[:3, 0].tolist()
            for b in badn:
                print("   nb row", b, int(nb[b]), int(rnb[b]))
    print("topk ok:", ok)
    return ok


def stage_fwd():
    from tricolo_b200 import ops
    ok = True
    tau = 0.1
    c1 = 1.4426950408889634 / tau
    for (b, d, npairs) in ((128, 512, 1), (256, 512, 3), (200, 512, 3), (384, 256, 1), (1024, 512, 3), (2048, 512, 1)):
        g = torch.Generator().manual_seed(b + d)
        feats = [torch.randn(b, d, generator=g).cuda() for _ in range(3)]
        zs, invs, _ = ops.l2norm_fwd(feats, ops.F16)
        pairs = [(0, 1), (0, 2), (1, 2)][:npairs]
        rs, cs, dg = ops.ntxent_fwd([zs[a] for a, _ in pairs], [zs[bb] for _, bb in pairs], 0, 1 / tau, ops.F16)
        torch.cuda.synchronize()
        for p, (a, bb) in enumerate(pairs):
            s = zs[a].double() @ zs[bb].double().t()
            e = torch.exp2(c1 * s - c1)
            er = ((rs[p].double() - e.sum(1)).abs() / e.sum(1)).max().item()
            ec = ((cs[p].double() - e.sum(0)).abs() / e.sum(0)).max().item()
            ed = (dg[p].double() - c1 * s.diagonal()).abs().max().item()
            good = er < 2e-5 and ec < 2e-5 and ed < 1e-4
            ok &= good
            print(f"  fwd B={b} D={d} pair{p}: row {er:.2e} col {ec:.2e} diag {ed:.2e} {'ok' if good else 'FAIL'}")
            if not good:
                print("   rs", rs[p][:4].tolist(), e.sum(1)[:4].tolist(), "cs", cs[p][:4].tolist(), e.sum(0)[:4].tolist())
                badr = ((rs[p].double() - e.sum(1)).abs() / e.sum(1) > 1e-3).nonzero()[:8, 0].tolist()
                badc = ((cs[p].double() - e.sum(0)).abs() / e.sum(0) > 1e-3).nonzero()[:8, 0].tolist()
                print("   bad rows", badr, "bad cols", badc)
    # sharded rows: rows [256,512) of a 512 batch
    g = torch.Generator().manual_seed(5)
    feats = [torch.randn(512, 512, generator=g).cuda() for _ in range(2)]
    zs, _, _ = ops.l2norm_fwd(feats, ops.F16)
    rs, cs, dg = ops.ntxent_fwd([zs[0][256:]], [zs[1]], 256, 1 / tau, ops.F16)
    s = zs[0][256:].double() @ zs[1].double().t()
    e = torch.exp2(c1 * s - c1)
    er = ((rs[0].double() - e.sum(1)).abs() / e.sum(1)).max().item()
    ec = ((cs[0].double() - e.sum(0)).abs() / e.sum(0)).max().item()
    ed = (dg[0].double() - c1 * torch.diagonal(s, offset=256)).abs().max().item()
    good = er < 2e-5 and ec < 2e-5 and ed < 1e-4
    ok &= good
    print(f"  fwd sharded rows: row {er:.2e} col {ec:.2e} diag {ed:.2e} {'ok' if good else 'FAIL'}")
    print("fwd ok:", ok)
    return ok


def stage_bwd():
    from oracle import ntxent_oracle as NO
    from tests.cases import LOSS_CASES, bf16_rounded, loss_case
    from tricolo_b200.loss import NTXentLoss, calculate_losses
    ok = True
    fn = NTXentLoss(0.1, 0.25)
    for name in LOSS_CASES + ["BIG1024"]:
        if name == "BIG1024":
            g = torch.Generator().manual_seed(7)
            base = torch.randn(1024, 512, generator=g)
            feats = {k: base + 0.5 * torch.randn(1024, 512, generator=g) for k in ("text_features", "image_features", "voxel_features")}
        else:
            feats = loss_case(name)
        feats = bf16_rounded(feats)
        dev = {k: v.cuda().requires_grad_(True) for k, v in feats.items()}
        losses = calculate_losses(dev, "train_loss", fn)
        losses["train_loss/total_loss"].backward()
        torch.cuda.synchronize()
        ref_l, ref_g = NO.trimodal_forward_backward({k: v.numpy() for k, v in feats.items()}, 0.1, 0.25)
        worst = 0.0
        for k, v in ref_l.items():
            rel = abs(float(losses[k]) - v) / abs(v)
            worst = max(worst, rel)
        gerr = {}
        for k, v in dev.items():
            gr = torch.from_numpy(ref_g[k])
            got = v.grad.double().cpu()
            gerr[k] = ((got - gr).norm() / gr.norm()).item(), ((got - gr).abs().max() / gr.abs().max()).item()
        good = worst < 1e-3 and all(a < 1e-3 and b < 2e-3 for a, b in gerr.values())
        ok &= good
        print(f"  {name}: loss rel {worst:.2e} grads " + " ".join(f"{k[:3]}:fro {a:.2e} max {b:.2e}" for k, (a, b) in gerr.items()),
              "ok" if good else "FAIL")
    print("bwd ok:", ok)
    return ok


def stage_eval():
    from oracle import retrieval_oracle as RO
    from tricolo_b200.evaluation import compute_metrics
    ok = True
    os.chdir("/tmp")
    for name, tuples in (("KAT_E1", RO.make_integer_kat()),
                         ("SMALL", RO.make_val_shaped(seed=3, n_shapes=300, n_queries=1000, dim=128, round_bf16=True)),
                         ("C3", RO.make_val_shaped(round_bf16=True))):
        t0 = time.time()
        got = compute_metrics("Text2Shape", {"caption_embedding_tuples": tuples})
        torch.cuda.synchronize()
        t1 = time.time()
        ref = RO.compute_metrics(tuples)
        good = all(np.array_equal(got[k], ref[k]) for k in ("precision", "recall", "recall_rate", "ndcg")) and got["mrr"] == ref["mrr"]
        ok &= good
        print(f"  eval {name}: RR@1 {got['recall_rate'][0]:.4f} RR@5 {got['recall_rate'][4]:.4f} ndcg5 {got['ndcg'][4]:.4f} mrr {got['mrr']:.4f}"
              f" vs ref mrr {ref['mrr']:.4f}  {t1 - t0:.3f}s  {'ok' if good else 'DIFF'}")
    print("eval ok:", ok)
    return ok


def main():
    stage = sys.argv[1] if len(sys.argv) > 1 else "all"
    if stage == "all":
        results = {}
        for s in STAGES:
            t0 = time.time()
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), s], timeout=300, capture_output=True, text=True)
                out = (r.stdout + r.stderr).strip().splitlines()
                print("\n".join(out[-60:]))
                results[s] = r.returncode
            except subprocess.TimeoutExpired:
                print(f"stage {s}: TIMEOUT")
                results[s] = "timeout"
            print(f"== stage {s}: rc={results[s]} ({time.time() - t0:.1f}s)", flush=True)
        print("SUMMARY", results)
        return 0
    fn = globals()["stage_" + stage]
    return 0 if fn() else 1


if __name__ == "__main__":
    sys.exit(main())
