"""CUDA-graph replay of the loss step (forward + backward) for fixed shapes.

The loss at training batch sizes is launch-bound on the host side: ~10 kernel launches, a few allocations
and (multi-GPU) three NCCL collectives per step against 50-300 us of device time.  Capturing the step once
and replaying it removes the host from the loop; the device work is unchanged (same kernels, same collectives).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .loss.nt_xent import DEFAULT_OP_FORMAT, trimodal_ntxent


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
                return global_trimodal_ntxent(fs, temperature, alpha, group, op_format=op_format)
        else:
            def fn(fs):
                return trimodal_ntxent(fs, temperature, alpha, op_format=op_format)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                for f in self.static_in:
                    f.grad = None
                fn(self.static_in).sum().backward()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        for f in self.static_in:
            f.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.losses = fn(self.static_in)
            self.losses.sum().backward()
        self.grads: List[Optional[torch.Tensor]] = [f.grad for f in self.static_in]

    def replay(self):
        """Re-run the captured step on whatever currently sits in the static input buffers."""
        self.graph.replay()
        return self.losses, self.grads

    def step(self, feats: Sequence[torch.Tensor]):
        for dst, src in zip(self.static_in, feats):
            dst.detach().copy_(src, non_blocking=True)
        return self.replay()
