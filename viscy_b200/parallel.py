"""Data-parallel gradient exchange for the hot path: one process per GPU, NCCL all-reduce of the fp32 gradients.

The reference shards over the batch with Lightning's `strategy: ddp`: the only collective is the
gradient all-reduce (mean).  Stock `DistributedDataParallel` works with the modules of this package in eager mode;
`BucketedGradAllReduce` performs the same exchange in a form that can be recorded into the whole-step CUDA graph and
overlaps it with the backward pass:

  * gradients live in ONE persistent flat fp32 buffer, cut into a few buckets in the order the backward pass produces
    them (recorded from the first backward: head / decoder first, stem last);
  * a post-accumulate hook per parameter counts a bucket down; when its last gradient lands, the bucket is packed
    (one multi-tensor copy) and all-reduced (average) on a communication stream while backward keeps running on the
    main stream -- inside a CUDA graph this is a parallel branch;
  * `finish()` joins the communication stream and re-points every `.grad` at its slice of the flat buffer, so the
    optimizer reads the averaged values without another pass.  Only the last (small) bucket is exposed.

Parameters that receive no gradient (e.g. ConvBlock3D.resid_conv when residual=False) keep `.grad = None`, as under
stock DDP, and contribute zeros to their bucket.
"""

from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.distributed as dist


def _slot(p: torch.Tensor) -> int:
    """Elements a gradient occupies in the flat buffer: padded to 16 bytes so that every view is float4-aligned (the
    one-launch AdamW reads gradients 16 bytes at a time; the pads stay zero and ride along in the all-reduce)."""
    return (p.numel() + 3) // 4 * 4


class _Bucket:
    def __init__(self, params: list[torch.nn.Parameter], flat: torch.Tensor):
        self.params = params
        self.flat = flat
        self.views = []
        off = 0
        for p in params:
            self.views.append(flat[off:off + p.numel()].view_as(p))
            off += _slot(p)
        self.pending = len(params)
        self.launched = False


class BucketedGradAllReduce:
    """fractions: cumulative share of the gradient bytes at which buckets end, in backward order; the remainder (the
    parameters whose gradients arrive last) forms the final, exposed bucket and should be small."""

    def __init__(self, params: Iterable[torch.nn.Parameter], group=None, fractions: Sequence[float] = (0.3, 0.6, 0.9, 0.985)):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.numel = sum(_slot(p) for p in self.params)
        self.fractions = tuple(fractions)
        self.order: list[torch.nn.Parameter] = []
        self._seen: set[int] = set()
        self.buckets: list[_Bucket] | None = None
        self._bucket_of: dict[int, _Bucket] = {}
        self.comm: torch.cuda.Stream | None = None
        self._avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        for p in self.params:
            p.register_post_accumulate_grad_hook(self._hook)

    # ------------------------------------------------------------------ setup
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

    def _build(self) -> None:
        """Cut the recorded backward order into buckets of one persistent flat buffer."""
        seen = {id(p) for p in self.order}
        ordered = self.order + [p for p in self.params if id(p) not in seen]  # never-ready parameters go last
        dev = ordered[0].device
        self.flat = torch.zeros((self.numel,), device=dev, dtype=torch.float32)
        cuts = [f * self.numel for f in self.fractions] + [float("inf")]
        groups, cur, acc, ci = [], [], 0, 0
        for p in ordered:
            cur.append(p)
            acc += _slot(p)
            if acc >= cuts[ci]:
                groups.append(cur)
                cur = []
                while acc >= cuts[ci]:
                    ci += 1
        if cur:
            groups.append(cur)
        self.buckets, off = [], 0
        for g in groups:
            n = sum(_slot(p) for p in g)
            b = _Bucket(g, self.flat[off:off + n])
            off += n
            self.buckets.append(b)
            for p in g:
                self._bucket_of[id(p)] = b
        if dev.type == "cuda":
            self.comm = torch.cuda.Stream(device=dev)

    # ------------------------------------------------------------------ per step
    def _hook(self, p: torch.nn.Parameter) -> None:
        if self.world == 1:
            return
        if self.buckets is None:  # first backward: record the order in which gradients become final
            if id(p) not in self._seen:
                self._seen.add(id(p))
                self.order.append(p)
            return
        b = self._bucket_of[id(p)]
        b.pending -= 1
        if b.pending == 0:
            self._launch(b)

    @torch.no_grad()
    def _launch(self, b: _Bucket) -> None:
        """Pack and all-reduce one bucket on the communication stream (behind everything queued on the main stream)."""
        b.launched = True
        have = [(v, p.grad) for v, p in zip(b.views, b.params) if p.grad is not None]
        missing = [v for v, p in zip(b.views, b.params) if p.grad is None]
        ctx = None
        if self.comm is not None:
            self.comm.wait_stream(torch.cuda.current_stream())
            ctx = torch.cuda.stream(self.comm)
            ctx.__enter__()
        try:
            if have:
                torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
            for v in missing:
                v.zero_()
            if self._avg:
                dist.all_reduce(b.flat, op=dist.ReduceOp.AVG, group=self.group)
            else:
                dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.group)
                b.flat.mul_(1.0 / self.world)
        finally:
            if ctx is not None:
                ctx.__exit__(None, None, None)

    @torch.no_grad()
    def finish(self) -> None:
        """After backward: exchange whatever has not been sent yet, join, and hand the averaged gradients back."""
        if self.world == 1:
            return
        if self.buckets is None:  # first step: build the buckets from the recorded order, then exchange them all now
            self._build()
        for b in self.buckets:
            if not b.launched:
                self._launch(b)
        if self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)
        for b in self.buckets:
            for v, p in zip(b.views, b.params):
                if p.grad is not None:
                    p.grad = v
            b.pending = len(b.params)
            b.launched = False

    __call__ = finish


# Round-1 name: one flat all-reduce after backward.  Kept as an alias of the bucketed exchange with a single bucket.
class FlatGradAllReduce(BucketedGradAllReduce):
    def __init__(self, params: Iterable[torch.nn.Parameter], group=None):
        super().__init__(params, group=group, fractions=())
