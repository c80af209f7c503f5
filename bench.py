#!/usr/bin/env python
"""Benchmark of the TriCoLo embedding-similarity hot path on B200 (contract: task spec, "bench.py").

    python bench.py --gpus N --steps K --warmup W            # our arm (sm_100a kernels)
    python bench.py --impl reference --gpus N --steps K ...  # reference arm: CPU port of the reference path
    torchrun-style launch for N > 1 (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* from the env).

Workload (config.workload = "c4"): BASELINE.json configs[3] — global-negative trimodal InfoNCE, global
batch 8192, dim 512, tau 0.1, alpha 0.25, forward + backward, STRONG scaling: the global batch is fixed and
each of the N ranks owns B/N rows (at N=1 the whole problem runs on one GPU).  metric = pairs/s =
global batch / time of one fwd+bwd of the whole three-term loss.  A secondary object "retrieval" reports
configs[4]-shaped sharded top-5 retrieval (queries/s); "small_batch" reports configs[1] (B=256) latency.

One JSON line on stdout (rank 0).  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TAU, ALPHA, DIM = 0.1, 0.25, 512
FEATURE_KEYS = ("text_features", "image_features", "voxel_features")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=["c4", "c2"])
    ap.add_argument("--batch", type=int, default=0, help="override the global batch")
    ap.add_argument("--no-retrieval", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--retrieval-queries", type=int, default=1_000_000)
    ap.add_argument("--retrieval-gallery", type=int, default=200_000)
    ap.add_argument("--op-format", default="f16", choices=["f16", "bf16"])
    return ap.parse_args()


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tf_burst": p["bf16_tflops"], "tf_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.004):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.max_mhz = index, period_s, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(self.period)

    def finish(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def make_features(batch, rows, row0, seed=7, device="cpu", dtype=None):
    """SURVEY.md §8d C4 inputs: correlated modalities (base + 0.5 noise), seed 7; rows [row0, row0+rows)."""
    import torch

    g = torch.Generator().manual_seed(seed)
    base = torch.randn(batch, DIM, generator=g)
    out = {}
    for k in FEATURE_KEYS:
        out[k] = (base + 0.5 * torch.randn(batch, DIM, generator=g))[row0:row0 + rows].contiguous()
    return out


# ----------------------------------------------------------------------------------------------- reference arm
def run_reference(args, rank, world):
    """CPU port of the reference path (oracle.torch_trimodal = the same PyTorch ops as nt_xent.py, autograd
    backward) on the host cores; rank 0 only."""
    if rank != 0:
        return
    import torch

    from oracle import ntxent_oracle as NO

    batch = args.batch or (8192 if args.workload == "c4" else 256)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    feats = make_features(batch, batch, 0)

    def step(fd):
        fd = {k: v.clone().requires_grad_(True) for k, v in fd.items()}
        out = NO.torch_trimodal(fd, TAU, ALPHA)
        out["train_loss/total_loss"].backward()
        return float(out["train_loss/total_loss"])

    # probe once; if a full-size step is too slow, time a bounded sample: a smaller global batch
    t0 = time.perf_counter()
    step(feats)
    probe = time.perf_counter() - t0
    sample_batch = batch
    budget_s = 150.0
    while probe * (sample_batch / batch) ** 2 * (args.steps + args.warmup) > budget_s and sample_batch > 1024:
        sample_batch //= 2
    if sample_batch != batch:
        feats = make_features(sample_batch, sample_batch, 0)
    for _ in range(max(args.warmup - 1, 1)):
        step(feats)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(feats)
    dt = (time.perf_counter() - t0) / args.steps
    # cost is quadratic in the batch: a full-size step costs (batch/sample)^2 sample steps
    full_step = dt * (batch / sample_batch) ** 2
    value = batch / full_step
    sample = (f"full workload per step (B={batch}, 3 pair terms, fwd+bwd, fp32)" if sample_batch == batch else
              f"B={sample_batch} sub-batch per step, extrapolated x{(batch // sample_batch) ** 2} (cost is quadratic in B)")
    line = {
        "impl": "reference", "metric": "infonce_fwd_bwd_pairs_per_s", "value": value, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": full_step * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: global-negative trimodal InfoNCE fwd+bwd", "global_batch": batch,
                   "dim": DIM, "temperature": TAU, "alpha_weight": ALPHA, "pairs": 3},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from tricolo_b200 import _lib, ops
    from tricolo_b200.distributed import global_trimodal_ntxent, sharded_retrieve
    from tricolo_b200.evaluation import retrieve
    from tricolo_b200.loss import trimodal_ntxent

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    peaks = load_peaks()
    op = ops.F16 if args.op_format == "f16" else ops.BF16
    batch = args.batch or (8192 if args.workload == "c4" else 256)
    assert batch % (128 * world) == 0 or world == 1, "global batch must split into 128-row blocks per rank"
    b_loc = batch // world
    row0 = rank * b_loc
    host = make_features(batch, b_loc, row0)
    host = {k: v.pin_memory() for k, v in host.items()}
    feats = [host[k].to(dev).requires_grad_(True) for k in FEATURE_KEYS]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()

    def loss_fn(fs):
        if world > 1:
            return global_trimodal_ntxent(fs, TAU, ALPHA, op_format=op)
        return trimodal_ntxent(fs, TAU, ALPHA, op_format=op)

    def step(fs):
        for f in fs:
            f.grad = None
        losses = loss_fn(fs)
        losses.sum().backward()
        return losses

    def timed(fn, steps, warmup):
        """Device time of `steps` calls, L2 flushed (untimed) before each; returns per-step ms list."""
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize(dev)
        barrier()
        evs = []
        for _ in range(steps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize(dev)
        barrier()
        return [a.elapsed_time(b) for a, b in evs]

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- main timed region: K steps, kernel events + clocks sampled meanwhile
    # The step is captured once in a CUDA graph and replayed (same kernels, same NCCL collectives, no host work
    # between launches); the eager number is reported next to it.  The per-kernel event profile is taken on
    # eager steps (events recorded inside a capture would be baked into the graph).
    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        step(feats)
    torch.cuda.synchronize(dev)
    _lib.profile_enable(True)
    launches0 = _lib.launch_count()
    eager_times = timed(lambda: step(feats), args.steps, 0)
    launches = _lib.launch_count() - launches0
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    eager_ms = max_over_ranks(sum(eager_times)) / args.steps
    launch_mode, graphed = "eager", None
    try:
        from tricolo_b200.graphs import GraphedTrimodalLoss

        graphed = GraphedTrimodalLoss(feats, TAU, ALPHA, distributed=world > 1, op_format=op, warmup=3)
        launch_mode = "cuda_graph"
    except Exception as e:  # report, never hide
        launch_mode = "eager (graph capture failed: %s)" % repr(e)[:120]
    run_step = graphed.replay if graphed is not None else (lambda: step(feats))
    sampler.start()
    wall0 = time.perf_counter()
    times = timed(run_step, args.steps, args.warmup)
    wall = time.perf_counter() - wall0
    clocks = sampler.finish()
    total_ms = max_over_ranks(sum(times))
    ms_per_step = total_ms / args.steps
    if eager_ms < ms_per_step:  # never report the slower of the two launch modes as the headline
        ms_per_step, launch_mode = eager_ms, "eager"
    value = batch / (ms_per_step * 1e-3)

    # ---------------- per-kernel numbers and the roofline of the dominant kernel
    pairs = 3
    flops_fwd = 2.0 * b_loc * batch * DIM * pairs           # one similarity GEMM per pair
    flops_bwd = 4.0 * b_loc * batch * DIM * pairs           # two gradient GEMMs per pair (algorithmic)
    kern = {}
    for name, (ms, n) in prof.items():
        kern[name] = {"ms_per_launch": ms / n, "launches_per_step": n / args.steps}
    def tf(flops, name):
        return flops / (kern[name]["ms_per_launch"] * 1e-3) / 1e12 if name in kern else None
    bytes_l2n = 3 * (b_loc * DIM * 4 + b_loc * DIM * 2 + b_loc * 4)
    if "ntxent_fwd" in kern:
        kern["ntxent_fwd"].update({"bound": "tensor", "algorithmic_tflops": tf(flops_fwd, "ntxent_fwd")})
    if "ntxent_bwd" in kern:
        # executed flop of the gradient kernel per algorithmic flop: the producer/consumer and pair kernels recompute
        # the logits once per direction (2x), the independent-CTA kernel once per dim half (3x)
        # the shared-G form (one GPU, >= 2048 rows) forms the logits once per pair (1.5x), the producer/consumer
        # kernel once per direction (2x), the independent-CTA kernel once per dim half (3x)
        bwd_mode = os.environ.get("TRICOLO_B200_BWD") or ("sharedg" if (world == 1 and batch >= 2048) else "pc")
        exec_factor = {"indep": 3.0, "sharedg": 1.5}.get(bwd_mode, 2.0)
        if os.environ.get("TRICOLO_B200_BWD_NOPAIR"):
            exec_factor = 3.0
        if bwd_mode == "sharedg" and "ntxent_g" in kern:
            # two kernels: ntxent_g (logit recompute -> 16-bit G, 2 B^2 D per pair, not algorithmic) and the gradient
            # GEMMs ntxent_ggemm (event id ntxent_bwd: exactly the algorithmic 4 B^2 D per pair)
            kern["ntxent_g"].update({"bound": "tensor", "executed_tflops": tf(flops_fwd, "ntxent_g"),
                                     "note": "logit recompute for the backward: executed, not algorithmic"})
            kern["ntxent_bwd"].update({"bound": "tensor", "algorithmic_tflops": tf(flops_bwd, "ntxent_bwd"),
                                       "executed_tflops": tf(flops_bwd, "ntxent_bwd"), "mode": bwd_mode,
                                       "kernel": "ntxent_ggemm_kernel"})
        else:
            kern["ntxent_bwd"].update({"bound": "tensor", "algorithmic_tflops": tf(flops_bwd, "ntxent_bwd"),
                                       "executed_tflops": tf(flops_bwd * exec_factor, "ntxent_bwd"), "mode": bwd_mode})
    if "l2norm_fwd" in kern:
        kern["l2norm_fwd"].update({"bound": "hbm", "gbs": bytes_l2n / (kern["l2norm_fwd"]["ms_per_launch"] * 1e-3) / 1e9})
    peak_tf = peaks["tf_sustained"]
    bwd_is_sharedg = kern.get("ntxent_bwd", {}).get("mode") == "sharedg"
    ach = kern.get("ntxent_bwd", {}).get("algorithmic_tflops")
    traffic = None
    try:  # DRAM bytes per launch of the same kernel at the same shapes, from the committed ncu capture
        with open(os.path.join(ROOT, "profiles", "r1h_traffic.json")) as f:
            tj = json.load(f)
            traffic = tj.get("ntxent_ggemm_kernel") if bwd_is_sharedg else tj.get("ntxent_bwd_pc_kernel")
            if not (world == 1 and batch == 8192):
                traffic = None
    except Exception:
        traffic = None
    bwd_kernel = "ntxent_ggemm_kernel" if bwd_is_sharedg else "ntxent_bwd_pc_kernel"
    roofline = {"kernel": bwd_kernel, "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": (ach / peak_tf) if ach else None, "traffic": traffic,
                "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                "algorithmic_flops_per_launch": flops_bwd,
                "whole_step": {"algorithmic_flops": flops_fwd + flops_bwd,
                               "achieved": (flops_fwd + flops_bwd) / (ms_per_step * 1e-3) / 1e12,
                               "frac": (flops_fwd + flops_bwd) / (ms_per_step * 1e-3) / 1e12 / peak_tf}}

    # ---------------- e2e: host buffers in, loss + gradients back to the host, every step
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    out_host = [torch.empty((b_loc, DIM), dtype=torch.float32).pin_memory() for _ in FEATURE_KEYS]
    loss_host = torch.empty((3,), dtype=torch.float32).pin_memory()
    d2h = sum(t.numel() * 4 for t in out_host) + 12

    def e2e_step():
        fs = [host[k].to(dev, non_blocking=True).requires_grad_(True) for k in FEATURE_KEYS]
        losses = loss_fn(fs)
        losses.sum().backward()
        loss_host.copy_(losses.detach(), non_blocking=True)
        for o, f in zip(out_host, fs):
            o.copy_(f.grad, non_blocking=True)

    e2e_times = timed(e2e_step, args.steps, max(3, args.warmup))
    seq_ms = max_over_ranks(sum(e2e_times)) / args.steps
    e2e = {"value": batch / (seq_ms * 1e-3), "unit": "pairs/s", "ms_per_step": seq_ms,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "api": "tricolo_b200.loss.trimodal_ntxent(...).sum().backward() on pinned host fp32 embeddings"}
    # The same work through the pipelined public entry: every step still copies its own inputs from pinned host memory
    # and its own losses + gradients back; the H2D of step k+1, the loss graph of step k and the D2H of step k-1 overlap.
    try:
        from tricolo_b200.graphs import HostPipelinedLoss

        depth = 3
        pipe = HostPipelinedLoss([host[k] for k in FEATURE_KEYS], TAU, ALPHA, depth=depth, distributed=world > 1,
                                 op_format=op)
        feats_h = [host[k] for k in FEATURE_KEYS]
        for _ in range(max(3, args.warmup)):
            pipe.result(pipe.submit(feats_h))
        torch.cuda.synchronize(dev)
        barrier()
        e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_a.record(pipe.s_h2d)
        tickets = []
        for k in range(args.steps):
            tickets.append(pipe.submit(feats_h))
            if k >= depth - 1:
                pipe.result(tickets[k - (depth - 1)])
        for t in tickets[-(depth - 1):]:
            pipe.result(t)
        e_b.record(pipe.s_d2h)
        torch.cuda.synchronize(dev)
        barrier()
        pipe_ms = max_over_ranks(e_a.elapsed_time(e_b)) / args.steps
        e2e.update({"value": batch / (pipe_ms * 1e-3), "ms_per_step": pipe_ms, "sequential_ms_per_step": seq_ms,
                    "sequential_value": batch / (seq_ms * 1e-3),
                    "api": "tricolo_b200.graphs.HostPipelinedLoss.submit/result on pinned host fp32 embeddings "
                           f"(depth {depth}: H2D, loss graph and D2H of consecutive steps overlap; K steps timed from the "
                           "first H2D to the last D2H); sequential_* = trimodal_ntxent(...).sum().backward() one step at a time"})
    except Exception as ex:  # report, never hide
        e2e["pipelined_error"] = repr(ex)[:200]

    # ---------------- secondary: small batch (configs[1]) latency
    small = None
    if world == 1 and args.workload == "c4":
        sf = [x.to(dev).requires_grad_(True) for x in make_features(256, 256, 0, seed=1234).values()]
        st = timed(lambda: step(sf), 50, 10)
        small = {"workload": "c2: trimodal B=256 fwd+bwd", "ms_per_step": statistics.median(st),
                 "pairs_per_s": 256 / (statistics.median(st) * 1e-3)}
        # the same step captured once in a CUDA graph (what a launch-bound training loop would replay)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    step(sf)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            for f in sf:
                f.grad = None
            with torch.cuda.graph(graph):
                trimodal_ntxent(sf, TAU, ALPHA, op_format=op).sum().backward()
            gt = timed(graph.replay, 50, 10)
            small["graph_ms_per_step"] = statistics.median(gt)
            small["graph_pairs_per_s"] = 256 / (statistics.median(gt) * 1e-3)
        except Exception as e:  # report, never hide
            small["graph_error"] = repr(e)[:200]

    # ---------------- secondary: sharded retrieval (configs[4]-shaped)
    retrieval = None
    if not args.no_retrieval:
        retrieval = bench_retrieval(args, rank, world, dev, peaks, timed, max_over_ranks, retrieve, sharded_retrieve, _lib)

    # ---------------- CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(batch)
        # the same PyTorch ops as the reference loss, eager, ON this GPU (fp32 inputs, torch defaults):
        # "the kernel to beat on the same box".  Part of the baseline leg; never on the product path.
        try:
            cpu["torch_eager_on_gpu"] = torch_eager_gpu_baseline(batch, dev)
        except Exception as ex:
            cpu["torch_eager_on_gpu"] = {"error": repr(ex)[:200]}

    if rank == 0:
        line = {
            "metric": "infonce_fwd_bwd_pairs_per_s", "value": value, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16" if op == ops.F16 else "bf16", "data": "synthetic",
            "config": {"workload": (f"{args.workload}: global-negative trimodal InfoNCE fwd+bwd (BASELINE configs[3])" if args.workload == "c4"
                                    else f"{args.workload}: Tri(I+V) trimodal loss fwd+bwd, batch 256 (BASELINE configs[1])"),
                       "global_batch": batch, "rows_per_rank": b_loc, "dim": DIM, "temperature": TAU, "alpha_weight": ALPHA,
                       "pairs": 3, "accumulate": "f32", "l2": "flushed (256 MB write) before every timed step",
                       "launch": launch_mode, "eager_ms_per_step": eager_ms,
                       "parallelism": f"row-block x{world}"},
            "roofline": roofline, "kernels": kern, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "cpu_baseline": cpu, "small_batch": small, "retrieval": retrieval,
            "wall_s_timed_region": wall,
        }
        print(json.dumps(line), flush=True)


def bench_retrieval(args, rank, world, dev, peaks, timed, max_over_ranks, retrieve, sharded_retrieve, _lib):
    """configs[4]: Q queries against a gallery of G shapes sharded over the ranks, top-5 + rank.
    Headline = the fused kernel (similarities stay on chip, tensor-bound).  The two-kernel form the north star
    describes (GEMM -> fp32 block in HBM -> top-k kernel) is timed on a slice of the same queries so that the
    top-k kernel's HBM fraction and the GEMM's numbers stay on record."""
    import torch

    n_q, n_g = args.retrieval_queries, args.retrieval_gallery
    g_loc = n_g // world
    base = rank * g_loc
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    gal = torch.randn(g_loc, DIM, generator=gen, device=dev).bfloat16()
    genq = torch.Generator(device=dev).manual_seed(99)  # same queries / labels on every rank
    text = torch.randn(n_q, DIM, generator=genq, device=dev).bfloat16()
    labels = torch.randint(0, n_g, (n_q,), generator=genq, device=dev)

    def run(fused, nq, block=None):
        if world > 1:
            return sharded_retrieve(text[:nq], gal, labels[:nq], base, 5, block_queries=block, fused=fused)
        return retrieve(text[:nq], gal, labels[:nq], 5, block_queries=block, fused=fused)

    steps = 2
    run(True, n_q)  # warm-up outside the kernel profile (its launches must not enter the per-launch flop count)
    _lib.profile_enable(True)
    t = timed(lambda: run(True, n_q), steps, 0)
    prof = _lib.profile_read()
    _lib.profile_enable(False)
    ms = max_over_ranks(sum(t)) / steps
    out = {"metric": "retrieval_queries_per_s", "value": n_q / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms,
           "config": {"workload": "c5: sharded top-5 retrieval (BASELINE configs[4]), fused GEMM+top-k kernel",
                      "queries": n_q, "gallery": n_g, "gallery_per_rank": g_loc, "dim": DIM, "k": 5, "operands": "bf16"}}
    if "sim_topk_fused" in prof:
        ms_f, n_f = prof["sim_topk_fused"]
        fl = 2.0 * n_q * g_loc * DIM / (n_f / steps)  # per launch
        tfs = fl / (ms_f / n_f * 1e-3) / 1e12
        out["fused_roofline"] = {"kernel": "sim_topk_fused_kernel", "bound": "tensor", "achieved": tfs,
                                 "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": tfs / peaks["tf_sustained"],
                                 "ms_per_launch": ms_f / n_f, "launches_per_step": n_f / steps}
    # two-kernel form on a slice (bounded: block x G_loc x 4 bytes of fp32 similarities per block)
    block = 8192
    nq2 = min(n_q, 16 * block)
    run(False, nq2, block)
    _lib.profile_enable(True)
    t2 = timed(lambda: run(False, nq2, block), steps, 0)
    prof2 = _lib.profile_read()
    _lib.profile_enable(False)
    ms2 = max_over_ranks(sum(t2)) / steps
    two = {"queries": nq2, "block_queries": block, "ms_per_step": ms2, "queries_per_s": nq2 / (ms2 * 1e-3)}
    if "topk_rank" in prof2:
        ms_k, n_k = prof2["topk_rank"]
        bytes_k = block * g_loc * 4 + block * 5 * 8 + block * 8  # per launch (one query block)
        gbs = bytes_k / (ms_k / n_k * 1e-3) / 1e9
        two["topk_roofline"] = {"kernel": "topk_rank_kernel", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                                "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "ms_per_launch": ms_k / n_k}
    if "sim_gemm" in prof2:
        ms_g, n_g_l = prof2["sim_gemm"]
        fl = 2.0 * block * g_loc * DIM
        tfs = fl / (ms_g / n_g_l * 1e-3) / 1e12
        two["gemm_roofline"] = {"kernel": "sim_gemm_resident_kernel", "bound": "tensor", "achieved": tfs,
                                "peak": peaks["tf_sustained"], "unit": "TFLOP/s", "frac": tfs / peaks["tf_sustained"],
                                "ms_per_launch": ms_g / n_g_l,
                                "hbm_write_gbs": block * g_loc * 4 / (ms_g / n_g_l * 1e-3) / 1e9}
    out["two_kernel_form"] = two
    if world > 1:
        # SURVEY 8(e) comparison: QUERY sharding (gallery replicated, every rank scores Q/W queries, no merge, no
        # collective) against the gallery sharding above
        try:
            gen_full = torch.Generator(device=dev).manual_seed(1234)
            gal_full = torch.randn(n_g, DIM, generator=gen_full, device=dev).bfloat16()
            q_loc = (n_q + world - 1) // world
            lo, hi = rank * q_loc, min((rank + 1) * q_loc, n_q)
            fq = lambda: retrieve(text[lo:hi], gal_full, labels[lo:hi], 5)
            fq()
            tq = timed(fq, steps, 0)
            msq = max_over_ranks(sum(tq)) / steps
            out["query_sharded_comparison"] = {"ms_per_step": msq, "queries_per_s": n_q / (msq * 1e-3),
                                               "note": "gallery replicated on every rank, Q/W queries per rank, no merge"}
        except Exception as ex:
            out["query_sharded_comparison"] = {"error": repr(ex)[:200]}
    return out


def cpu_baseline(batch):
    """The oracle's CPU-PyTorch port of the reference loss on the host cores: bounded sample (~10-30 s)."""
    import torch

    from oracle import ntxent_oracle as NO

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sample_batch = min(batch, 2048)
    feats = make_features(sample_batch, sample_batch, 0)

    def step():
        fd = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
        out = NO.torch_trimodal(fd, TAU, ALPHA)
        out["train_loss/total_loss"].backward()

    step()
    t0 = time.perf_counter()
    n = 0
    while n < 3 or (time.perf_counter() - t0 < 10.0 and n < 50):
        step()
        n += 1
    dt = (time.perf_counter() - t0) / n
    full = dt * (batch / sample_batch) ** 2
    return {"value": batch / full, "unit": "pairs/s", "cores": cores, "kind": "port",
            "sample": f"{n} fwd+bwd steps of the trimodal loss at B={sample_batch} (fp32, torch CPU, {cores} threads), "
                      f"scaled x{(batch // sample_batch) ** 2} to B={batch} (cost is quadratic in B)",
            "ms_per_sample_step": dt * 1e3}


def torch_eager_gpu_baseline(batch, dev):
    """The oracle's PyTorch port of nt_xent.py:24-74 + tricolo_net.py:56-65 run eagerly on the GPU (autograd backward)."""
    import torch

    from oracle import ntxent_oracle as NO

    feats = {k: v.to(dev) for k, v in make_features(batch, batch, 0).items()}

    def step():
        fd = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
        out = NO.torch_trimodal(fd, TAU, ALPHA)
        out["train_loss/total_loss"].backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / n
    return {"value": batch / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms, "kind": "port",
            "sample": f"{n} eager fwd+bwd steps at B={batch} on the GPU, fp32 tensors, "
                      f"allow_tf32={torch.backends.cuda.matmul.allow_tf32}"}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a B200; there is no CPU path")
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the single JSON line: NCCL's version banner goes to stderr
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    elif args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist

            dist.destroy_process_group()


if __name__ == "__main__":
    main()
