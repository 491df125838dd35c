"""Unet3d / UNet3DBase / Unet25d host mirrors (CPU, fp32) against golden vectors from the reference's own code
(BASELINE configs 1 and 5 at reduced size)."""
from pathlib import Path

import pytest
import torch

from viscy_b200 import Unet3d, Unet25d, UNet3DBase
from viscy_b200.unet3d import ConvBottleneck3D

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name,cls", [("unet3d", Unet3d), ("unet25d", Unet25d)])
def test_cpu_matches_reference_golden(name, cls):
    g = torch.load(GOLD / f"{name}.pt", weights_only=False)
    torch.manual_seed(g["seed"])  # same RNG consumption as the reference constructor -> identical weights
    m = cls(**g["cfg"])
    assert len(m.state_dict()) == g["n_keys"]
    out = m(g["x"])
    torch.testing.assert_close(out, g["outs"][0], rtol=1e-5, atol=1e-6)
    loss = torch.nn.functional.mse_loss(out, g["targets"][0])
    loss.backward()
    assert abs(loss.item() - g["loss"]) < 1e-5
    for n, p in m.named_parameters():
        if p.grad is None:  # ConvBlock3D.resid_conv is registered but unused when residual=False (reference behaviour)
            assert n not in g["grad_norms"]
            continue
        ref = g["grad_norms"][n]
        assert abs(p.grad.norm().item() - ref) <= 1e-4 * max(ref, 1e-3), n


def test_key_counts_pinned_by_reference():
    """test_state_dict_compat.py:286 (Unet25d = 147) and the 146 keys of Unet3d(3,3,4,32)."""
    sd = Unet25d().state_dict()
    assert len(sd) == 147
    for k in ["down_conv_block_2.Conv3d_1.weight", "up_conv_block_3.resid_conv.bias"]:
        assert k in sd
    assert len(Unet3d(3, 3, 4, 32).state_dict()) == 146
    assert Unet3d(1, 1, depth=2, mult_chan=4).num_blocks == 2 and Unet3d().downsamples_z


def test_unet3d_validation_errors():
    m = Unet3d(1, 1, depth=3, mult_chan=4)
    with pytest.raises(ValueError, match="must be divisible by 8"):
        m(torch.randn(1, 1, 12, 16, 16))
    with pytest.raises(ValueError, match="must equal"):
        UNet3DBase(1, 1, dims=[4, 8], num_res_block=[1, 1], bottleneck=ConvBottleneck3D(8))
    # F-Net initialisation: conv weights ~ N(0, 0.02)
    assert abs(m.inconv.weight.std().item() - 0.02) < 0.01


def test_unet25d_unsupported_configs_raise_on_cuda():
    """Configurations without sm_100a kernels raise (never a cuDNN fallback); fp32 input without autocast raises too."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA tensor")
    from viscy_b200.unet25d import ConvBlock3D
    with pytest.raises(NotImplementedError):
        Unet25d().cuda()(torch.randn(1, 1, 5, 32, 32, device="cuda"))  # fp32, no autocast
    blk = ConvBlock3D(16, 16, norm="instance").cuda()
    with pytest.raises(NotImplementedError):
        blk.forward_cl(torch.randn(1, 4, 8, 8, 16, device="cuda").half())


@pytest.mark.parametrize("name", ["unet3d_base_gn", "unet3d_base_gn_odd"])
def test_unet3d_base_groupnorm_cpu_matches_reference_golden(name):
    """UNet3DBase with the class defaults (GroupNorm + SiLU), residual blocks, timestep + conditioning inputs: the host
    mirror against vectors from the reference's own code (tests/golden/make_golden_unet3d_base.py)."""
    g = torch.load(GOLD / f"{name}.pt", weights_only=False)
    cfg = g["cfg"]
    bott = ConvBottleneck3D(cfg["dims"][-1], time_emb_dim=g["time_embed_dim"], residual=True, groups=cfg["groups"])
    m = UNet3DBase(bottleneck=bott, time_embed_dim=g["time_embed_dim"], cond_channels=g["cond_channels"], **cfg)
    assert list(m.state_dict()) == list(g["state_dict"])
    m.load_state_dict(g["state_dict"])
    out = m(g["x"], g["cond"], g["t"])
    torch.testing.assert_close(out, g["out"], rtol=1e-5, atol=1e-6)
    torch.nn.functional.mse_loss(out, g["target"]).backward()
    for n, p in m.named_parameters():
        torch.testing.assert_close(p.grad, g["grads"][n], rtol=1e-4, atol=1e-6, msg=n)
