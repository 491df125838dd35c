"""FullyConvolutionalMAE host-side mirror (CPU / torch backend): state_dict surface, RNG consumption and numerics against
the reference's own code (when /root/reference is present) and against the golden vectors generated from it."""
from pathlib import Path

import pytest
import torch

from oracle import reference_loader as RL
from viscy_b200 import FullyConvolutionalMAE

GOLD = Path(__file__).resolve().parent / "golden"


def perturb(model):
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "grn" in n or n.endswith("bias"):
                p.normal_(0, 0.2)


@pytest.mark.parametrize("name", ["fcmae_dense", "fcmae_masked", "fcmae_2d"])
def test_cpu_backend_matches_reference_golden(name):
    g = torch.load(GOLD / f"{name}.pt", weights_only=False)
    torch.manual_seed(g["seed"])
    m = FullyConvolutionalMAE(**g["cfg"])
    assert len(m.state_dict()) == g["n_keys"]
    perturb(m)
    torch.manual_seed(g["seed"] + 2000)  # the mask draw
    out = m(g["x"].clone(), g["mask_ratio"]) if g["mask_ratio"] > 0 else m(g["x"].clone())
    if g["mask"] is not None:
        out, mask = out
        assert torch.equal(mask, g["mask"])
    torch.testing.assert_close(out, g["out"], rtol=1e-5, atol=1e-6)
    loss = torch.nn.functional.mse_loss(out, g["target"])
    loss.backward()
    assert abs(loss.item() - g["loss"]) < 1e-5
    for n, p in m.named_parameters():
        if n in g["grad_norms"]:
            ref = g["grad_norms"][n]
            assert abs(p.grad.norm().item() - ref) <= 1e-4 * max(ref, 1e-3), n
        else:
            assert p.grad is None, n  # stem.conv2d / conv3d: the branch the input depth does not take


@pytest.mark.skipif(not RL.available(), reason="/root/reference is not present")
@pytest.mark.parametrize("cfg,mask_ratio", [
    (dict(in_channels=1, out_channels=1, in_stack_depth=5), 0.5),
    (dict(in_channels=2, out_channels=2, in_stack_depth=15, head_conv=True, pretraining=False,
          encoder_blocks=(1, 1, 2, 1), dims=(48, 96, 192, 384)), 0.0),
    (dict(in_channels=1, out_channels=3, in_stack_depth=1, stem_kernel_size=(1, 4, 4), encoder_drop_path_rate=0.1), 0.25),
])
def test_mirror_equals_reference_code(cfg, mask_ratio):
    """Same seeds -> identical state_dict (keys, order, values), identical mask draw, identical forward (train mode, so
    DropPath draws are compared too)."""
    ns = RL.load()
    torch.manual_seed(3)
    r = ns.FullyConvolutionalMAE(**cfg)
    torch.manual_seed(3)
    m = FullyConvolutionalMAE(**cfg)
    sr, sm = r.state_dict(), m.state_dict()
    assert list(sr) == list(sm)
    assert all(torch.equal(sr[k], sm[k]) for k in sr)
    x = torch.randn(2, cfg["in_channels"], cfg["in_stack_depth"], 64, 64)
    torch.manual_seed(5)
    a = r(x.clone(), mask_ratio)
    torch.manual_seed(5)
    b = m(x.clone(), mask_ratio)
    if isinstance(a, tuple):
        assert torch.equal(a[0], b[0])
        assert (a[1] is None and b[1] is None) or torch.equal(a[1], b[1])
    else:
        assert torch.equal(a, b)


def test_attribute_surface_and_errors():
    m = FullyConvolutionalMAE(1, 2, in_stack_depth=5, head_conv=True)
    assert m.num_blocks == 8 and m.out_stack_depth == 5 and m.pretraining is True
    assert {n for n, _ in m.named_children()} == {"encoder", "decoder", "head"}
    assert m.encoder.total_stride == 32
    with pytest.raises(ValueError, match="length of drop_path_rates"):
        from viscy_b200.fcmae import MaskedConvNeXtV2Stage
        MaskedConvNeXtV2Stage(8, 8, num_blocks=2, drop_path_rates=[0.1])
    from viscy_b200.fcmae import generate_mask, upsample_mask
    mk = generate_mask(torch.Size((2, 1, 5, 64, 64)), 32, 0.5, "cpu")
    assert mk.shape == (2, 1, 2, 2) and mk.sum().item() == 4
    with pytest.raises(ValueError, match="must be divisible by mask shape"):
        upsample_mask(mk, torch.Size((2, 1, 5, 5)))
