"""NT-Xent product loss vs the oracle restatement of PML's NTXentLoss (reference test: test_loss.py:22-32, beta=0
equals the PML parent), plus the qualitative HCL properties the reference tests."""
import torch

from oracle.models import ntxent
from viscy_b200.loss import NTXentHCL, NTXentLoss


def _batch(b=8, d=16, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(2 * b, d, generator=g), torch.cat([torch.arange(b), torch.arange(b)])


def test_matches_oracle_and_beta0_equals_parent():
    e, l = _batch()
    for t in (0.07, 0.2, 0.5):
        ref = ntxent(e, l, t)
        torch.testing.assert_close(NTXentLoss(temperature=t)(e, l), ref)
        torch.testing.assert_close(NTXentHCL(temperature=t, beta=0.0)(e, l), ref)


def test_gradient_flows_and_hcl_differs():
    e, l = _batch(seed=1)
    e.requires_grad_(True)
    NTXentHCL(beta=0.5)(e, l).backward()
    assert e.grad is not None and torch.isfinite(e.grad).all() and e.grad.abs().sum() > 0
    assert not torch.allclose(NTXentHCL(beta=1.0)(e.detach(), l), NTXentLoss()(e.detach(), l))


def test_temperature_schedule():
    loss = NTXentLoss(temperature=0.07, temperature_schedule="cosine", temperature_start=0.1, temperature_warmup_epochs=10)
    loss.step(0)
    assert abs(loss.temperature - 0.1) < 1e-9
    loss.step(10)
    assert abs(loss.temperature - 0.07) < 1e-9
