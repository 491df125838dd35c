"""Unet2d / ConvBlock2D host-side mirror (CPU / torch backend) against the reference's own code (when present) and the
golden vectors generated from it."""
from pathlib import Path

import pytest
import torch

from oracle import reference_loader as RL
from viscy_b200 import ConvBlock2D, Unet2d

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name", ["unet2d", "unet2d_res"])
def test_cpu_backend_matches_reference_golden(name):
    g = torch.load(GOLD / f"{name}.pt", weights_only=False)
    torch.manual_seed(g["seed"])
    m = Unet2d(**g["cfg"])
    assert len(m.state_dict()) == g["n_keys"]
    out = m(g["x"])
    torch.testing.assert_close(out, g["out"], rtol=1e-5, atol=1e-6)
    loss = torch.nn.functional.mse_loss(out, g["target"])
    loss.backward()
    assert abs(loss.item() - g["loss"]) < 1e-5
    for n, p in m.named_parameters():
        if n in g["grad_norms"]:
            ref = g["grad_norms"][n]
            assert abs(p.grad.norm().item() - ref) <= 1e-4 * max(ref, 1e-3), n
        else:
            assert p.grad is None and "resid_conv" in n, n  # registered but unused (reference behaviour)


@pytest.mark.skipif(not RL.available(), reason="/root/reference is not present")
def test_mirror_equals_reference_code():
    ns = RL.load()
    for cfg in (dict(), dict(in_channels=2, out_channels=3, task="reg", residual=True, num_blocks=3,
                             num_filters=(8, 16, 32, 64)), dict(kernel_size=(5, 3), num_block_layers=3)):
        torch.manual_seed(1)
        r = ns.Unet2d(**cfg)
        torch.manual_seed(1)
        m = Unet2d(**cfg)
        sr, sm = r.state_dict(), m.state_dict()
        assert list(sr) == list(sm) and all(torch.equal(sr[k], sm[k]) for k in sr)
        x = torch.randn(2, cfg.get("in_channels", 1), 1, 64, 64)
        assert torch.equal(r(x), m(x))
    for kw in (dict(norm="instance", activation="elu", layer_order="cna", filter_steps="linear"),
               dict(filter_steps="last", residual=True, dropout=True), dict(activation="linear", norm="none", num_repeats=1)):
        torch.manual_seed(1)
        r = ns.ConvBlock2D(8, 16, **kw)
        torch.manual_seed(1)
        m = ConvBlock2D(8, 16, **kw)
        assert list(r.state_dict()) == list(m.state_dict())
        x = torch.randn(2, 8, 16, 16)
        torch.manual_seed(2)
        a = r(x)
        torch.manual_seed(2)
        assert torch.equal(a, m(x))


def test_constructor_errors_and_surface():
    with pytest.raises(ValueError, match="Kernel dims must be odd"):
        ConvBlock2D(4, 8, kernel_size=4)
    with pytest.raises(ValueError, match="kernel_size length must be 2"):
        ConvBlock2D(4, 8, kernel_size=(3, 3, 3))
    with pytest.raises(AttributeError, match="must be either int or tuple"):
        ConvBlock2D(4, 8, kernel_size=[3, 3])
    with pytest.raises(NotImplementedError, match="'same' padding in ConvTranspose2d"):
        ConvBlock2D(4, 8, transpose=True)
    with pytest.raises(NotImplementedError, match="Activation type tanh not supported"):
        ConvBlock2D(4, 8, activation="tanh")
    m = Unet2d()
    assert m.num_blocks == 4 and len(m.state_dict()) == 148
    assert "down_conv_block_0.resid_conv.weight" in m.state_dict()
    assert not any(k.startswith("up_samp") or "dropout" in k for k in m.state_dict())
