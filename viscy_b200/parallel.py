"""Data-parallel gradient exchange for the hot path: one process per GPU, NCCL all-reduce of the fp32 gradients.

The reference shards over the batch with Lightning's `strategy: ddp`: the only collective is the
gradient all-reduce (mean).  Stock `DistributedDataParallel` works with the modules of this package in eager mode;
this helper performs the same exchange in a form that can be recorded into the whole-step CUDA graph: gradients are
gathered into one flat fp32 buffer, all-reduced in place over NCCL/NVLink (NVLS when available), averaged, and handed
back to the parameters as views.
"""

from __future__ import annotations

from typing import Iterable

import torch
import torch.distributed as dist


class FlatGradAllReduce:
    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.numel = sum(p.numel() for p in self.params)

    @torch.no_grad()
    def broadcast_parameters(self, src: int = 0) -> None:
        """Same initial weights on every rank (what DDP's constructor does)."""
        if self.world == 1:
            return
        flat = torch.cat([p.detach().reshape(-1) for p in self.params])
        dist.broadcast(flat, src, group=self.group)
        off = 0
        for p in self.params:
            p.copy_(flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    @torch.no_grad()
    def __call__(self) -> None:
        """After backward: average `.grad` of all parameters across ranks (parameters without a grad count as 0)."""
        if self.world == 1:
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.mul_(1.0 / self.world)
        off = 0
        for p in self.params:
            p.grad = flat[off:off + p.numel()].view_as(p)
            off += p.numel()
