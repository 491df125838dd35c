"""Golden vectors for MixedLoss / ms_ssim_25d from the reference's OWN code (authoring container only).

    python tests/golden/make_golden_mixed_loss.py

viscy_utils/evaluation/metrics.py and viscy_utils/losses/mixed_loss.py are executed unmodified
(oracle/reference_loader.load_losses); the SSIM path in them is pure torch, so these fixtures are 100 % reference code.
Inputs are bf16-exact uniform random volumes (stored as bf16); the prediction gradient is stored on an in-plane 1/3 grid.
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import reference_loader as RL  # noqa: E402

OUT = Path(__file__).resolve().parent
CASES = {
    "mixed_loss_default": dict(seed=31, shape=(2, 1, 5, 192, 192), kw=dict(l1_alpha=0.5, l2_alpha=0.0, ms_dssim_alpha=0.5)),
    "mixed_loss_all": dict(seed=32, shape=(1, 1, 9, 200, 232), kw=dict(l1_alpha=0.3, l2_alpha=0.2, ms_dssim_alpha=0.5)),
}


def main():
    ns = RL.load_losses()
    for name, c in CASES.items():
        g = torch.Generator().manual_seed(c["seed"])
        x = torch.rand(c["shape"], generator=g).to(torch.bfloat16)
        # target: a blurred / shifted relative of the prediction plus noise, so the SSIM terms are far from both 0 and 1
        y = (0.6 * x.float() + 0.4 * torch.rand(c["shape"], generator=g)).to(torch.bfloat16)
        xp = x.float().requires_grad_(True)
        loss = ns.MixedLoss(**c["kw"])(xp, y.float())
        loss.backward()
        ms = ns.ms_ssim_25d(x.float(), y.float(), clamp=True)
        ssim, cs = ns.ssim_25d(x.float(), y.float(), return_contrast_sensitivity=True)
        torch.save({"kw": c["kw"], "x": x, "y": y, "loss": loss.item(), "ms_ssim": ms.item(), "ssim": ssim, "cs": cs,
                    "grad_norm": xp.grad.norm().item(), "grad_sub": xp.grad[..., ::3, ::3].clone(),
                    "torch": torch.__version__}, OUT / f"{name}.pt")
        print(name, f"loss={loss.item():.6f} ms_ssim={ms.item():.6f}", ssim.tolist(), cs.tolist())


if __name__ == "__main__":
    main()
