"""The oracle (oracle/models.py + restated timm/monai) pinned against the golden vectors generated from the
reference's own code, and - when /root/reference is present - against that code directly."""
from pathlib import Path

import pytest
import torch

from oracle import models as OM
from oracle import reference_loader as RL

GOLD = Path(__file__).resolve().parent / "golden"
ORACLE_CLS = {"UNeXt2": OM.UNeXt2, "ContrastiveEncoder": OM.ContrastiveEncoder}


def _run(model, g):
    out = model(g["x"])
    outs = out if isinstance(out, tuple) else (out,)
    loss = sum(torch.nn.functional.mse_loss(o, t) for o, t in zip(outs, g["targets"]))
    loss.backward()
    return outs, loss


@pytest.mark.parametrize("name", ["unext2_atto", "unext2_tiny", "contrastive_tiny"])
def test_oracle_reproduces_golden(name):
    g = torch.load(GOLD / f"{name}.pt", weights_only=False)
    torch.manual_seed(g["seed"])
    m = ORACLE_CLS[g["cls"]](**g["cfg"])
    assert len(m.state_dict()) == g["n_keys"]
    outs, loss = _run(m, g)
    for o, ref in zip(outs, g["outs"]):
        torch.testing.assert_close(o, ref, rtol=1e-5, atol=1e-6)
    assert abs(loss.item() - g["loss"]) < 1e-5 * max(1.0, abs(g["loss"]))
    for n, p in m.named_parameters():
        ref = g["grad_norms"][n]
        assert abs(p.grad.norm().item() - ref) <= 1e-4 * max(ref, 1e-3), n


def test_reference_key_counts_and_sentinels():
    """packages/viscy-models/tests/test_state_dict_compat.py:30-124: 213 (UNeXt2 atto), 194 (ContrastiveEncoder)."""
    m = OM.UNeXt2(backbone="convnextv2_atto")
    sd = m.state_dict()
    assert len(sd) == 213
    for k in ["encoder_stages.stages_1.blocks.1.mlp.fc2.bias", "decoder.decoder_stages.0.conv.blocks.0.conv_dw.weight",
              "head.conv.1.weight", "stem.conv.weight", "head.conv.0.adn.A.weight"]:
        assert k in sd
    assert {k.split(".")[0] for k in sd} == {"encoder_stages", "stem", "decoder", "head"}
    ce = OM.ContrastiveEncoder("convnext_tiny", 2, 15)
    sd = ce.state_dict()
    assert len(sd) == 194
    for k in ["encoder.head.norm.bias", "encoder.stages.2.blocks.4.gamma", "projection.4.weight"]:
        assert k in sd
    assert len(OM.UNeXt2(in_stack_depth=21, stem_kernel_size=(7, 4, 4)).state_dict()) == 273


def test_unext2_shapes_like_reference_tests():
    """test_unext2.py:9-43 (shape contract)."""
    m = OM.UNeXt2(in_channels=1, out_channels=3, in_stack_depth=5, out_stack_depth=5, backbone="convnextv2_atto").eval()
    with torch.no_grad():
        y = m(torch.randn(1, 1, 5, 64, 64))
    assert y.shape == (1, 3, 5, 64, 64)


@pytest.mark.skipif(not RL.available(), reason="/root/reference only exists in the authoring container")
@pytest.mark.parametrize("cls,cfg,xshape", [
    ("UNeXt2", dict(in_channels=2, out_channels=1, in_stack_depth=10, backbone="convnextv2_atto", head_pool=False), (1, 2, 10, 32, 32)),
    ("ContrastiveEncoder", dict(backbone="convnext_tiny", in_channels=1, in_stack_depth=10), (3, 1, 10, 32, 32)),
])
def test_oracle_equals_reference_code(cls, cfg, xshape):
    ns = RL.load()
    torch.manual_seed(5)
    ref = getattr(ns, cls)(**cfg)
    torch.manual_seed(5)
    ora = ORACLE_CLS[cls](**cfg)
    for (ka, va), (kb, vb) in zip(ref.state_dict().items(), ora.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    x = torch.randn(xshape)
    a, b = ref(x), ora(x)
    for u, v in zip(a if isinstance(a, tuple) else (a,), b if isinstance(b, tuple) else (b,)):
        assert torch.equal(u, v)


def test_ntxent_oracle_properties():
    """NT-Xent restatement: permutation invariant, lower when positives align."""
    torch.manual_seed(0)
    a = torch.randn(8, 16)
    labels = torch.cat([torch.arange(8), torch.arange(8)])
    good = OM.ntxent(torch.cat([a, a + 0.01 * torch.randn(8, 16)]), labels)
    bad = OM.ntxent(torch.cat([a, torch.randn(8, 16)]), labels)
    assert good < bad
    assert torch.isfinite(good) and good > 0
