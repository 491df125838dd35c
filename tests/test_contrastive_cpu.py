"""ContrastiveEncoder host mirror (CPU / torch backend) against the golden vectors from the reference's own code."""
from pathlib import Path

import pytest
import torch

from oracle import models as OM
from viscy_b200 import ContrastiveEncoder

GOLD = Path(__file__).resolve().parent / "golden"


def test_cpu_backend_matches_reference_golden():
    g = torch.load(GOLD / "contrastive_tiny.pt", weights_only=False)
    torch.manual_seed(g["seed"])
    o = OM.ContrastiveEncoder(**g["cfg"])
    m = ContrastiveEncoder(**g["cfg"])
    assert list(m.state_dict()) == list(o.state_dict()) and len(m.state_dict()) == 194
    m.load_state_dict(o.state_dict())
    emb, proj = m(g["x"])
    torch.testing.assert_close(emb, g["outs"][0], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(proj, g["outs"][1], rtol=1e-4, atol=1e-4)
    loss = sum(torch.nn.functional.mse_loss(o_, t) for o_, t in zip((emb, proj), g["targets"]))
    loss.backward()
    assert abs(loss.item() - g["loss"]) < 1e-4


def test_surface_and_errors():
    m = ContrastiveEncoder("convnext_tiny", in_channels=2, in_stack_depth=15)
    assert {n for n, _ in m.named_children()} == {"stem", "encoder", "projection"}
    assert m.encoder.num_features == 768
    for k in ["encoder.head.norm.bias", "encoder.stages.2.blocks.4.gamma", "projection.4.weight"]:
        assert k in m.state_dict()
    with pytest.raises(ValueError, match="Stem needs to output"):
        ContrastiveEncoder("convnext_tiny", in_channels=1, in_stack_depth=25, stem_kernel_size=(5, 4, 4))  # 96 % 5
    with pytest.raises(NotImplementedError):
        ContrastiveEncoder("resnet50", in_channels=1, in_stack_depth=15)
