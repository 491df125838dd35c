"""Golden vectors for Unet2d from the reference's OWN code (100 % reference code; authoring container only).

    python tests/golden/make_golden_unet2d.py
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import reference_loader as RL  # noqa: E402

OUT = Path(__file__).resolve().parent
CASES = {
    "unet2d": dict(seed=41, xshape=(2, 1, 1, 128, 128), cfg=dict(in_channels=1, out_channels=2, task="reg")),
    "unet2d_res": dict(seed=42, xshape=(2, 2, 1, 64, 96),
                       cfg=dict(in_channels=2, out_channels=1, task="seg", residual=True, num_blocks=3,
                                num_filters=(16, 32, 64, 128), kernel_size=(3, 5))),
}


def main():
    ns = RL.load()
    for name, c in CASES.items():
        torch.manual_seed(c["seed"])
        model = ns.Unet2d(**c["cfg"])
        g = torch.Generator().manual_seed(c["seed"] + 1000)
        x = torch.randn(c["xshape"], generator=g)
        out = model(x)
        tgt = torch.randn(out.shape, generator=g)
        loss = torch.nn.functional.mse_loss(out, tgt)
        loss.backward()
        gn = {n: p.grad.norm().item() for n, p in model.named_parameters() if p.grad is not None}
        torch.save({"cfg": c["cfg"], "seed": c["seed"], "x": x, "out": out.detach(), "target": tgt, "loss": loss.item(),
                    "grad_norms": gn, "n_keys": len(model.state_dict()), "torch": torch.__version__}, OUT / f"{name}.pt")
        print(name, tuple(out.shape), f"loss={loss.item():.6f}", "keys", len(model.state_dict()))


if __name__ == "__main__":
    main()
