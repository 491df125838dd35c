"""Generate golden vectors from the reference's OWN code (authoring container only: needs /root/reference).

    python tests/golden/make_golden.py

Each fixture holds: constructor config, RNG seeds, the seeded input, the reference forward output, the MSE loss
against a seeded target and the L2 norm of every parameter gradient.  Weights are NOT stored: the reference modules
and oracle/models.py consume the torch CPU RNG identically, so `torch.manual_seed(seed)` before construction
reproduces them (checked by tests/test_oracle.py, which also compares against the reference directly when present).
UNeXt2 / ContrastiveEncoder run the reference's composition code over the restated timm/monai blocks
(oracle/ref_timm.py, ref_monai.py); Unet25d / Unet3d are 100 % reference code.
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import reference_loader as RL  # noqa: E402

OUT = Path(__file__).resolve().parent

CASES = {
    "unext2_atto": dict(cls="UNeXt2", seed=11, xshape=(2, 1, 5, 64, 64),
                        cfg=dict(in_channels=1, out_channels=2, in_stack_depth=5, backbone="convnextv2_atto",
                                 stem_kernel_size=(5, 4, 4), head_pool=True)),
    "unext2_tiny": dict(cls="UNeXt2", seed=12, xshape=(1, 1, 14, 64, 64),
                        cfg=dict(in_channels=1, out_channels=2, in_stack_depth=14, backbone="convnextv2_tiny",
                                 stem_kernel_size=(7, 4, 4), head_pool=True)),
    "contrastive_tiny": dict(cls="ContrastiveEncoder", seed=13, xshape=(4, 2, 15, 64, 64),
                             cfg=dict(backbone="convnext_tiny", in_channels=2, in_stack_depth=15,
                                      stem_kernel_size=(5, 4, 4), stem_stride=(5, 4, 4), embedding_dim=768,
                                      projection_dim=128)),
    "unet25d": dict(cls="Unet25d", seed=14, xshape=(2, 1, 5, 64, 64),
                    cfg=dict(in_channels=1, out_channels=1, in_stack_depth=5, out_stack_depth=1, task="reg", dropout=0.0)),
    "unet3d": dict(cls="Unet3d", seed=15, xshape=(1, 3, 16, 32, 32),
                   cfg=dict(in_channels=3, out_channels=3, depth=3, mult_chan=8)),
}


def run_case(model, xshape, seed):
    g = torch.Generator().manual_seed(seed + 1000)
    x = torch.randn(xshape, generator=g)
    out = model(x)
    outs = out if isinstance(out, tuple) else (out,)
    tg = [torch.randn(o.shape, generator=g) for o in outs]
    loss = sum(torch.nn.functional.mse_loss(o, t) for o, t in zip(outs, tg))
    loss.backward()
    gn = {n: p.grad.norm().item() for n, p in model.named_parameters() if p.grad is not None}
    return x, [o.detach() for o in outs], tg, loss.item(), gn


def main():
    ns = RL.load()
    for name, c in CASES.items():
        torch.manual_seed(c["seed"])
        model = getattr(ns, c["cls"])(**c["cfg"])
        x, outs, tg, loss, gn = run_case(model, c["xshape"], c["seed"])
        torch.save({"cls": c["cls"], "cfg": c["cfg"], "seed": c["seed"], "x": x, "outs": outs, "targets": tg, "loss": loss,
                    "grad_norms": gn, "n_keys": len(model.state_dict()), "third_party": ns.third_party,
                    "torch": torch.__version__}, OUT / f"{name}.pt")
        print(name, [tuple(o.shape) for o in outs], f"loss={loss:.6f}", "keys", len(model.state_dict()))


if __name__ == "__main__":
    main()
