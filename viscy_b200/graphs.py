"""CUDA-graph capture of a whole training step (forward + loss + backward + optimizer).

The hot path is ~800 short kernels per step; launched eagerly from Python the host becomes the bottleneck, so the
step is recorded once on a side stream and replayed.  Everything the step touches must be static: inputs are
copied into fixed buffers, the optimizer must be capturable (torch optimizers: `capturable=True`;
`viscy_b200.optim.AdamW` always is, its state being created by the eager warm-up steps below), no host synchronisation
inside the step.
"""

from __future__ import annotations

from typing import Callable, Sequence

import torch

from . import _lib


class GraphedStep:
    """`step_fn(*static_inputs) -> Tensor` captured into one CUDA graph.

    Call the instance with fresh inputs of the same shapes / dtypes; they are copied into the static buffers
    (device-to-device, or host-to-device when the sources are pinned host tensors) and the graph is replayed.
    """

    def __init__(self, step_fn: Callable[..., torch.Tensor], example_inputs: Sequence[torch.Tensor], warmup: int = 3):
        self.static_inputs = [torch.empty_like(t, device="cuda") if not t.is_cuda else t.clone() for t in example_inputs]
        for dst, src in zip(self.static_inputs, example_inputs):
            dst.copy_(src)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step_fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_output = step_fn(*self.static_inputs)
        self.launches_per_replay = _lib.launch_count() - n0

    def copy_inputs(self, *inputs: torch.Tensor) -> None:
        """Copy fresh inputs into the graph's static buffers (current stream)."""
        for dst, src in zip(self.static_inputs, inputs):
            if src is not dst:
                dst.copy_(src, non_blocking=True)

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.static_output

    def __call__(self, *inputs: torch.Tensor) -> torch.Tensor:
        self.copy_inputs(*inputs)
        return self.replay()
