"""NT-Xent losses of the DynaCLR training step (VM/contrastive/loss.py:20-186) without the pytorch-metric-learning
dependency: `NTXentLoss` (PML semantics: cosine similarity, all same-label ordered pairs as positives, all
different-label embeddings as negatives, mean over positive pairs) and `NTXentHCL` (hard-negative reweighting).

These are a handful of small torch ops on (2B, D) projections ("negligible FLOPs"); they stay in torch.
"""

from __future__ import annotations

import math
from typing import Literal

import torch
from torch import Tensor, nn


def cosine_anneal(start: float, end: float, epoch: int, warmup_epochs: int) -> float:
    """Cosine anneal from `start` to `end` over `warmup_epochs` (VM/schedule.py)."""
    if epoch >= warmup_epochs:
        return end
    return end + (start - end) * 0.5 * (1.0 + math.cos(math.pi * epoch / warmup_epochs))


class NTXentLoss(nn.Module):
    def __init__(self, temperature: float = 0.07, temperature_schedule: Literal["cosine", "constant"] = "constant",
                 temperature_start: float = 0.1, temperature_warmup_epochs: int = 50, **kwargs) -> None:
        super().__init__()
        self.temperature = temperature
        self.temperature_schedule = temperature_schedule
        self.temperature_start = temperature_start
        self.temperature_end = temperature
        self.temperature_warmup_epochs = temperature_warmup_epochs
        self.beta = 0.0

    def step(self, epoch: int) -> None:
        """Update the temperature for the given (0-indexed) epoch."""
        if self.temperature_schedule == "cosine":
            self.temperature = cosine_anneal(self.temperature_start, self.temperature_end, epoch,
                                             self.temperature_warmup_epochs)

    def forward(self, embeddings: Tensor, labels: Tensor) -> Tensor:
        emb = torch.nn.functional.normalize(embeddings.float(), dim=1)
        sim = emb @ emb.t()
        same = labels[:, None] == labels[None, :]
        eye = torch.eye(len(labels), dtype=torch.bool, device=sim.device)
        a1, p = torch.where(same & ~eye)
        a2, n = torch.where(~same)
        if a1.numel() == 0 or a2.numel() == 0:
            return sim.sum() * 0.0
        dtype = sim.dtype
        pos = sim[a1, p].unsqueeze(1) / self.temperature
        neg_raw = sim[a2, n]
        neg = neg_raw / self.temperature
        n_per_p = (a2.unsqueeze(0) == a1.unsqueeze(1)).to(dtype)
        neg_m = neg * n_per_p
        neg_m[n_per_p == 0] = torch.finfo(dtype).min
        mx = torch.max(pos, neg_m.max(dim=1, keepdim=True)[0]).detach()
        num = torch.exp(pos - mx).squeeze(1)
        w = torch.exp(neg_m - mx)
        if self.beta != 0.0:  # hard-negative concentration: reweight negatives by exp(beta * sim), renormalised
            hw = torch.exp(self.beta * neg_raw) * n_per_p
            hw = hw * n_per_p.sum(dim=1, keepdim=True) / hw.sum(dim=1, keepdim=True).clamp(min=1e-8)
            w = hw * w
        den = w.sum(dim=1) + num
        return (-torch.log(num / den + torch.finfo(dtype).tiny)).mean()


class NTXentHCL(NTXentLoss):
    def __init__(self, temperature: float = 0.07, beta: float = 0.5, **kwargs) -> None:
        super().__init__(temperature=temperature, **kwargs)
        self.beta = beta
