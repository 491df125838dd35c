"""Golden vectors for UNet3DBase with the class defaults (GroupNorm + SiLU), residual blocks, timestep conditioning and
a conditioning input, from the reference's OWN code (authoring container only: needs /root/reference; 100 % reference
code: VM/unet/unet3d_base.py, VM/unet/blocks.py).

    python tests/golden/make_golden_unet3d_base.py

Each fixture: constructor config, state_dict, inputs (x, cond, t), output, MSE loss against a seeded target, and every
parameter gradient (the models are tiny, so the full gradients are stored, not only their norms)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import reference_loader as RL  # noqa: E402

OUT = Path(__file__).resolve().parent

CASES = {
    # dims % 8 == 0; two ResnetBlocks on the deeper level (each decoder block pops its own skip)
    "unet3d_base_gn": dict(seed=21, xshape=(2, 2, 8, 16, 16), cond_channels=2, time_embed_dim=32,
                           cfg=dict(in_channels=2, out_channels=3, dims=[8, 16, 32], num_res_block=[1, 2], residual=True,
                                    downsample_z=True, groups=4)),
    # channel counts that are not multiples of 8: rows are padded, concatenations must compact
    "unet3d_base_gn_odd": dict(seed=22, xshape=(1, 1, 4, 16, 16), cond_channels=None, time_embed_dim=None,
                               cfg=dict(in_channels=1, out_channels=1, dims=[12, 20], num_res_block=[1], residual=True,
                                        downsample_z=False, groups=4)),
}


def main():
    ns = RL.load()
    import viscy_models.unet.blocks as RB
    for name, c in CASES.items():
        torch.manual_seed(c["seed"])
        cfg = dict(c["cfg"])
        bott = RB.ConvBottleneck3D(cfg["dims"][-1], time_emb_dim=c["time_embed_dim"], residual=True, groups=cfg["groups"])
        model = ns.UNet3DBase(bottleneck=bott, time_embed_dim=c["time_embed_dim"], cond_channels=c["cond_channels"], **cfg)
        with torch.no_grad():  # spread the affine parameters so that every path carries signal
            for n, p in model.named_parameters():
                if n.endswith("norm.weight"):
                    p.normal_(1.0, 0.2)
                elif n.endswith("bias"):
                    p.normal_(0.0, 0.2)
        g = torch.Generator().manual_seed(c["seed"] + 1000)
        x = torch.randn(c["xshape"], generator=g)
        cond = torch.randn((c["xshape"][0], c["cond_channels"], *c["xshape"][2:]), generator=g) if c["cond_channels"] else None
        t = torch.rand(c["xshape"][0], generator=g) * 100 if c["time_embed_dim"] else None
        out = model(x, cond, t)
        tgt = torch.randn(out.shape, generator=g)
        loss = torch.nn.functional.mse_loss(out, tgt)
        loss.backward()
        torch.save({"cfg": cfg, "time_embed_dim": c["time_embed_dim"], "cond_channels": c["cond_channels"],
                    "state_dict": {k: v.clone() for k, v in model.state_dict().items()}, "x": x, "cond": cond, "t": t,
                    "out": out.detach(), "target": tgt, "loss": loss.item(),
                    "grads": {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None},
                    "torch": torch.__version__}, OUT / f"{name}.pt")
        print(name, tuple(out.shape), f"loss={loss.item():.6f}", "keys", len(model.state_dict()))


if __name__ == "__main__":
    main()
