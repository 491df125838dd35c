"""Full-shape golden vectors from the reference's OWN code (100 % reference code; authoring container only):

    python tests/golden/make_golden_fullshape.py

  unet25d_c1 : BASELINE config 1 exactly - Unet25d(1 -> 1, 5 -> 1 slices, task "reg", dropout 0) on (2, 1, 5, 128, 128)
  unet3d_d4  : the config-5 architecture Unet3d(3, 3, depth 4, mult_chan 32) on a (1, 3, 64, 64, 64) volume: every level
               (64^3 x 32 ... 4^3 x 512) runs the tilings of the patch-form / resident-filter / implicit-GEMM conv kernels
Inputs and targets are regenerated from the stored seeds (torch CPU generator); the unet3d_d4 output is stored in fp16.
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import reference_loader as RL  # noqa: E402

OUT = Path(__file__).resolve().parent
CASES = {
    "unet25d_c1": dict(cls="Unet25d", seed=51, xshape=(2, 1, 5, 128, 128), half=False,
                       cfg=dict(in_channels=1, out_channels=1, in_stack_depth=5, out_stack_depth=1, task="reg", dropout=0.0)),
    "unet3d_d4": dict(cls="Unet3d", seed=52, xshape=(1, 3, 64, 64, 64), half=True,
                      cfg=dict(in_channels=3, out_channels=3, depth=4, mult_chan=32)),
}


def inputs(seed, xshape, out_shape=None):
    g = torch.Generator().manual_seed(seed + 1000)
    x = torch.randn(xshape, generator=g)
    return x, (torch.randn(out_shape, generator=g) if out_shape is not None else g)


def main():
    ns = RL.load()
    for name, c in CASES.items():
        torch.manual_seed(c["seed"])
        model = getattr(ns, c["cls"])(**c["cfg"])
        g = torch.Generator().manual_seed(c["seed"] + 1000)
        x = torch.randn(c["xshape"], generator=g)
        out = model(x)
        tgt = torch.randn(out.shape, generator=g)
        loss = torch.nn.functional.mse_loss(out, tgt)
        loss.backward()
        gn = {n: p.grad.norm().item() for n, p in model.named_parameters() if p.grad is not None}
        torch.save({"cls": c["cls"], "cfg": c["cfg"], "seed": c["seed"], "xshape": c["xshape"],
                    "out": out.detach().half() if c["half"] else out.detach(), "loss": loss.item(), "grad_norms": gn,
                    "n_keys": len(model.state_dict()), "torch": torch.__version__}, OUT / f"{name}.pt")
        print(name, tuple(out.shape), f"loss={loss.item():.6f}", "keys", len(model.state_dict()))


if __name__ == "__main__":
    main()
