"""MixedLoss / ms_ssim_25d host-side mirror (CPU / torch backend) against the reference's own code and its golden vectors;
the reference's own tests (packages/viscy-utils/tests/test_mixed_loss.py) restated for the CPU path."""
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

from oracle import reference_loader as RL
from viscy_b200.losses import MixedLoss, ms_ssim_25d, ssim_25d

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name", ["mixed_loss_default", "mixed_loss_all"])
def test_cpu_backend_matches_reference_golden(name):
    g = torch.load(GOLD / f"{name}.pt", weights_only=False)
    x = g["x"].float().requires_grad_(True)
    y = g["y"].float()
    loss = MixedLoss(**g["kw"])(x, y)
    loss.backward()
    assert abs(loss.item() - g["loss"]) < 1e-6
    torch.testing.assert_close(x.grad[..., ::3, ::3], g["grad_sub"], rtol=1e-4, atol=1e-9)
    assert abs(ms_ssim_25d(x.detach(), y, clamp=True).item() - g["ms_ssim"]) < 1e-6
    s, c = ssim_25d(x.detach(), y, return_contrast_sensitivity=True)
    torch.testing.assert_close(s, g["ssim"])
    torch.testing.assert_close(c, g["cs"])


@pytest.mark.skipif(not RL.available(), reason="/root/reference is not present")
def test_mirror_equals_reference_code():
    ns = RL.load_losses()
    torch.manual_seed(0)
    p, t = torch.rand(2, 1, 5, 192, 192), torch.rand(2, 1, 5, 192, 192)
    a, b = p.clone().requires_grad_(True), p.clone().requires_grad_(True)
    la, lb = ns.MixedLoss(0.3, 0.2, 0.5)(a, t), MixedLoss(0.3, 0.2, 0.5)(b, t)
    la.backward()
    lb.backward()
    assert torch.equal(la, lb) and torch.equal(a.grad, b.grad)
    assert torch.equal(ns.ms_ssim_25d(p, t, clamp=True), ms_ssim_25d(p, t, clamp=True))
    assert torch.equal(ns.ssim_25d(p, t, (7, 9)), ssim_25d(p, t, (7, 9)))


def test_l1_only_matches_torch_l1():
    """test_mixed_loss.py:92-107: ms_dssim_alpha=0 collapses to alpha * F.l1_loss bit-exact."""
    torch.manual_seed(3)
    pred, target = torch.rand(2, 1, 8, 64, 64), torch.rand(2, 1, 8, 64, 64)
    torch.testing.assert_close(MixedLoss(0.5, 0.0, 0.0)(pred, target), F.l1_loss(pred, target) * 0.5, rtol=0, atol=0)


def test_errors():
    with pytest.raises(ValueError, match="cannot be all zero"):
        MixedLoss(0.0, 0.0, 0.0)
    with pytest.raises(ValueError, match=r"Input shape must be \(B, C, D, W, H\)"):
        ssim_25d(torch.rand(1, 1, 32, 32), torch.rand(1, 1, 32, 32))
    with pytest.raises(ValueError, match="must have same shape"):
        ssim_25d(torch.rand(1, 1, 2, 32, 32), torch.rand(1, 1, 2, 32, 30))
