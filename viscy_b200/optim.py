"""AdamW of the training step on the sm_100a path.

The reference builds ``torch.optim.AdamW(self.model.parameters(), lr)`` (VU/optimizers.py:10-61
``configure_adamw_scheduler``; CY/engine.py:547-554; the contrastive engine's ``configure_optimizers``).  This class has the
same constructor, semantics, ``state_dict`` layout (``step`` / ``exp_avg`` / ``exp_avg_sq`` per parameter) and GradScaler
protocol (``_step_supports_amp_scaling``), but the whole parameter group is ONE kernel launch
(``vb200_adamw_step``, csrc/optim_sm100.cu) instead of one multi-tensor launch per ~36 tensors, and it is always
CUDA-graph capturable (the step counts live on the device, one fp32 scalar per parameter as in torch's capturable mode).
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

CHUNK = 2048  # elements per work item (csrc/optim_sm100.cu ADAM_CHUNK)
WINDOW = 448  # tensors per launch (ADAM_MAX_TENSORS)


class _Plan:
    """Static launch tables of one parameter group: {p, m, v, step} pointers and the chunk list, both on the device."""

    def __init__(self, params, moments):
        dev = params[0].device
        n = len(params)
        rows = []
        chunks, start = [], [0]
        for i, (p, (m, v, st)) in enumerate(zip(params, moments)):
            rows.append((p.data_ptr(), m.data_ptr(), v.data_ptr(), st.data_ptr()))
            numel = p.numel()
            for off in range(0, numel, CHUNK):
                chunks.append((i % WINDOW, off, min(CHUNK, numel - off), 0))
            start.append(len(chunks))
        self.key = tuple((p.data_ptr(), m.data_ptr()) for p, (m, _, _) in zip(params, moments))
        self.n = n
        self.table = torch.tensor(rows, dtype=torch.int64).to(dev)
        self.chunks = torch.tensor(chunks, dtype=torch.int32).reshape(-1, 4).to(dev)
        self.chunk_start = (C.c_int32 * (n + 1))(*start)
        self.grads = (C.c_void_p * n)()
        self.done = torch.zeros((1,), dtype=torch.int32, device=dev)


class AdamW(torch.optim.Optimizer):
    """``torch.optim.AdamW`` (decoupled weight decay, no amsgrad) for CUDA fp32 parameters, one launch per group."""

    _step_supports_amp_scaling = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, *, maximize=False):
        if isinstance(lr, torch.Tensor) and lr.numel() != 1:
            raise ValueError("Tensor lr must be 1-element")
        if not isinstance(lr, torch.Tensor) and lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameters: {betas}")
        if weight_decay < 0.0:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, maximize=maximize))
        self._plans: dict[int, _Plan] = {}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._plans = {}
        for group in self.param_groups:  # step counts live on the device, one fp32 scalar per parameter (as torch capturable)
            for p in group["params"]:
                st = self.state.get(p)
                if st and "step" in st:
                    st["step"] = torch.as_tensor(st["step"], dtype=torch.float32).reshape(()).to(p.device).clone()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        capturing = torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            dev = params[0].device
            moments = []
            for p in params:
                if not p.is_cuda:
                    raise RuntimeError("viscy_b200.optim.AdamW runs on CUDA parameters only (no CPU fallback)")
                if p.dtype != torch.float32 or not p.is_contiguous() or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise NotImplementedError("viscy_b200.optim.AdamW takes dense contiguous fp32 parameters and gradients")
                st = self.state[p]
                if "exp_avg" not in st:
                    if capturing:  # the zero-initialisation would be replayed with the graph
                        raise RuntimeError("viscy_b200.optim.AdamW: run one eager step before CUDA-graph capture "
                                           "(state is created on the first step)")
                    st["step"] = torch.zeros((), dtype=torch.float32, device=dev)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                moments.append((st["exp_avg"], st["exp_avg_sq"], st["step"]))
            plan = self._plans.get(gi)
            key = tuple((p.data_ptr(), m.data_ptr()) for p, (m, _, _) in zip(params, moments))
            if plan is None or plan.key != key:
                if capturing:  # building the launch tables copies from pageable host memory
                    raise RuntimeError("viscy_b200.optim.AdamW: the set of parameters with gradients changed under CUDA-graph "
                                       "capture; run one eager step with the same set first")
                plan = self._plans[gi] = _Plan(params, moments)
            for i, p in enumerate(params):
                g = p.grad
                if not g.is_contiguous():
                    g = p.grad = g.contiguous()
                plan.grads[i] = g.data_ptr()
            lr = group["lr"]
            lr_t = lr if isinstance(lr, torch.Tensor) else None
            if lr_t is not None and (lr_t.device != dev or lr_t.dtype != torch.float32):
                lr_t = lr_t.to(device=dev, dtype=torch.float32)
            for t, name in ((grad_scale, "grad_scale"), (found_inf, "found_inf")):
                if t is not None and (t.device != dev or t.dtype != torch.float32):
                    raise ValueError(f"{name} must be an fp32 scalar on {dev}")
            b1, b2 = group["betas"]
            L.check(
                L.lib().vb200_adamw_step(
                    L.ptr(plan.table), plan.grads, plan.chunk_start, plan.n, L.ptr(plan.chunks), L.ptr(plan.done),
                    C.c_float(0.0 if lr_t is not None else float(lr)), L.ptr(lr_t), C.c_float(b1), C.c_float(b2), C.c_float(group["eps"]), C.c_float(group["weight_decay"]), int(bool(group["maximize"])),
                    L.ptr(grad_scale), L.ptr(found_inf), L.stream_ptr(),
                ),
                "vb200_adamw_step",
            )
        return loss
