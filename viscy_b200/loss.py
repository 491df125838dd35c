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
        """Dense, static-shape form (no index gathering, no host synchronisation: the loss can be recorded into the
        training step's CUDA graph).  For every ordered positive pair (a, p):
            loss_ap = -log( exp(s_ap / T) / (exp(s_ap / T) + sum_{n in neg(a)} w_an exp(s_an / T)) ),   mean over pairs
        with w_an = 1 (PML NTXentLoss) or the hard-negative weights exp(beta s_an) renormalised to the number of
        negatives (NTXentHCL, VM/contrastive/loss.py:120-185)."""
        emb = torch.nn.functional.normalize(embeddings.float(), dim=1)
        sim = emb @ emb.t()
        same = labels[:, None] == labels[None, :]
        eye = torch.eye(len(labels), dtype=torch.bool, device=sim.device)
        pos_m = same & ~eye
        neg_m = ~same
        logits = sim / self.temperature
        ninf = torch.finfo(sim.dtype).min
        neg_logits = logits
        if self.beta != 0.0:
            hw = self.beta * sim  # log of the un-normalised hard-negative weights
            lse_hw = torch.logsumexp(hw.masked_fill(~neg_m, ninf), dim=1, keepdim=True)
            cnt = neg_m.sum(dim=1, keepdim=True).clamp_min(1).to(sim.dtype)
            neg_logits = logits + hw + torch.log(cnt) - lse_hw
        lneg = torch.logsumexp(neg_logits.masked_fill(~neg_m, ninf), dim=1, keepdim=True)  # [2B, 1]
        lneg = torch.where(neg_m.any(dim=1, keepdim=True), lneg, torch.full_like(lneg, ninf))
        per_pair = torch.logaddexp(logits, lneg) - logits  # -log(num / den) for every (a, p)
        n_pos = pos_m.sum()
        return (per_pair * pos_m).sum() / n_pos.clamp_min(1).to(sim.dtype)


class NTXentHCL(NTXentLoss):
    def __init__(self, temperature: float = 0.07, beta: float = 0.5, **kwargs) -> None:
        super().__init__(temperature=temperature, **kwargs)
        self.beta = beta
