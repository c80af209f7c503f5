"""CUDA-graph replay of the loss step (forward + backward) for fixed shapes.

The loss at training batch sizes is launch-bound on the host side: ~10 kernel launches, a few allocations
and (multi-GPU) three NCCL collectives per step against 50-300 us of device time.  Capturing the step once
and replaying it removes the host from the loop; the device work is unchanged (same kernels, same collectives).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .loss.nt_xent import DEFAULT_OP_FORMAT, trimodal_ntxent_total


class GraphedTrimodalLoss:
    """step(feats) -> (losses [n_pairs], grads [one per modality]); static buffers, one CUDA graph.

    distributed=True uses tricolo_b200.distributed.global_trimodal_ntxent (global negatives across the ranks of
    `group`; NCCL collectives are captured into the graph)."""

    def __init__(self, example_feats: Sequence[torch.Tensor], temperature: float, alpha: float,
                 distributed: bool = False, group=None, op_format: int = DEFAULT_OP_FORMAT, warmup: int = 3):
        self.static_in: List[torch.Tensor] = [torch.empty_like(f).copy_(f.detach()).requires_grad_(True)
                                              for f in example_feats]
        if distributed:
            from .distributed import global_trimodal_ntxent

            def fn(fs):
                losses = global_trimodal_ntxent(fs, temperature, alpha, group, op_format=op_format)
                return losses, losses.sum()
        else:
            def fn(fs):  # the sum comes out of the forward call itself: no framework kernels inside the graph
                return trimodal_ntxent_total(fs, temperature, alpha, op_format=op_format)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                for f in self.static_in:
                    f.grad = None
                fn(self.static_in)[1].backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for f in self.static_in:
            f.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses, total = fn(self.static_in)
            total.backward()
        self.grads: List[Optional[torch.Tensor]] = [f.grad for f in self.static_in]

    def replay(self):
        """Re-run the captured step on whatever currently sits in the static input buffers."""
        self.graph.replay()
        return self.losses, self.grads

    def step(self, feats: Sequence[torch.Tensor]):
        for dst, src in zip(self.static_in, feats):
            dst.detach().copy_(src, non_blocking=True)
        return self.replay()


class HostPipelinedLoss:
    """Host-resident embeddings in, losses and gradients back on the host, step after step, with the three stages of
    a step on three streams: H2D copy of step k+1, the captured loss graph of step k and the D2H copy of step k-1
    overlap (PCIe is full duplex).  Every step still moves its own inputs from pinned host memory and its own
    results back; only the stages of DIFFERENT steps overlap.

        pipe = HostPipelinedLoss(example_pinned_feats, temperature, alpha, depth=3)
        t = pipe.submit(pinned_feats)              # enqueue one step, returns a ticket
        losses, grads = pipe.result(t)             # pinned host tensors of that step (valid until `depth` later submits)
    """

    class _Slot:
        pass

    def __init__(self, example_host_feats: Sequence[torch.Tensor], temperature: float, alpha: float, depth: int = 3,
                 distributed: bool = False, group=None, op_format: int = DEFAULT_OP_FORMAT):
        if depth < 2:
            raise ValueError("HostPipelinedLoss: depth >= 2")
        if any(f.is_cuda for f in example_host_feats):
            raise ValueError("HostPipelinedLoss takes host tensors; use trimodal_ntxent for device tensors")
        dev = torch.device("cuda", torch.cuda.current_device())
        self.depth, self.k = depth, 0
        self.s_h2d, self.s_comp, self.s_d2h = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
        self.slots = []
        for _ in range(depth):
            sl = HostPipelinedLoss._Slot()
            ex = [f.to(dev) for f in example_host_feats]
            sl.graph = GraphedTrimodalLoss(ex, temperature, alpha, distributed=distributed, group=group,
                                           op_format=op_format)
            sl.host_losses = torch.empty(sl.graph.losses.shape, dtype=torch.float32).pin_memory()
            sl.host_grads = [torch.empty(f.shape, dtype=f.dtype).pin_memory() for f in example_host_feats]
            sl.ev_in, sl.ev_done, sl.ev_out = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
            sl.used = False
            self.slots.append(sl)
        torch.cuda.synchronize()

    def submit(self, host_feats: Sequence[torch.Tensor]) -> int:
        ticket = self.k
        sl = self.slots[ticket % self.depth]
        self.k += 1
        with torch.cuda.stream(self.s_h2d):
            if sl.used:
                self.s_h2d.wait_event(sl.ev_done)   # the previous step of this slot has read its inputs
            for dst, src in zip(sl.graph.static_in, host_feats):
                dst.detach().copy_(src, non_blocking=True)
            sl.ev_in.record(self.s_h2d)
        with torch.cuda.stream(self.s_comp):
            self.s_comp.wait_event(sl.ev_in)
            if sl.used:
                self.s_comp.wait_event(sl.ev_out)   # ... and its results have left the device
            sl.graph.replay()
            sl.ev_done.record(self.s_comp)
        with torch.cuda.stream(self.s_d2h):
            self.s_d2h.wait_event(sl.ev_done)
            sl.host_losses.copy_(sl.graph.losses.detach(), non_blocking=True)
            for dst, g in zip(sl.host_grads, sl.graph.grads):
                dst.copy_(g, non_blocking=True)
            sl.ev_out.record(self.s_d2h)
        sl.used = True
        return ticket

    def result(self, ticket: int):
        if ticket < self.k - self.depth or ticket >= self.k:
            raise ValueError("HostPipelinedLoss.result: ticket no longer (or not yet) held")
        sl = self.slots[ticket % self.depth]
        sl.ev_out.synchronize()
        return sl.host_losses, sl.host_grads
