"""Golden vectors for FullyConvolutionalMAE from the reference's OWN code (authoring container only: needs /root/reference).

    python tests/golden/make_golden_fcmae.py

VM/unet/fcmae.py is executed unmodified (oracle/reference_loader.py) over the restated timm symbols it imports
(Downsample, DropPath, GlobalResponseNormMlp, LayerNorm2d, create_conv2d, trunc_normal_: oracle/ref_timm.py) and the
restated monai UpSample / Convolution (oracle/ref_monai.py).  Weights are not stored: `torch.manual_seed(seed)` before
construction plus `perturb()` reproduces them (viscy_b200.fcmae consumes the RNG identically; tests/test_fcmae_cpu.py
checks that bit for bit against the reference when it is present).  The masked case stores the mask the reference drew.
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import reference_loader as RL  # noqa: E402

OUT = Path(__file__).resolve().parent

CASES = {
    # fine-tuning / inference configuration of the published VSCyto3D-style models: dense encoder, conv head
    "fcmae_dense": dict(seed=21, xshape=(2, 1, 5, 64, 64), mask_ratio=0.0,
                        cfg=dict(in_channels=1, out_channels=2, in_stack_depth=5, stem_kernel_size=(5, 4, 4),
                                 pretraining=False, head_conv=True, head_conv_expansion_ratio=4, head_conv_pool=True)),
    # pretraining configuration: sparse masked encoder, pixel-shuffle head, returns (reconstruction, mask)
    "fcmae_masked": dict(seed=22, xshape=(2, 1, 5, 128, 128), mask_ratio=0.5,
                         cfg=dict(in_channels=1, out_channels=1, in_stack_depth=5, stem_kernel_size=(5, 4, 4),
                                  encoder_blocks=(2, 2, 2, 2), dims=(64, 128, 256, 512), pretraining=True)),
    # 2-D input (VSCyto2D): Conv2d stem branch, depth 1
    "fcmae_2d": dict(seed=23, xshape=(2, 2, 1, 64, 64), mask_ratio=0.0,
                     cfg=dict(in_channels=2, out_channels=3, in_stack_depth=1, stem_kernel_size=(1, 4, 4),
                              encoder_blocks=(1, 1, 2, 1), dims=(32, 64, 128, 256), pretraining=False)),
}


def perturb(model):
    """GRN weights / biases are zero-initialised: give them (and every bias) values so all paths carry signal."""
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "grn" in n or n.endswith("bias"):
                p.normal_(0, 0.2)


def main():
    ns = RL.load()
    for name, c in CASES.items():
        torch.manual_seed(c["seed"])
        model = ns.FullyConvolutionalMAE(**c["cfg"])
        perturb(model)
        g = torch.Generator().manual_seed(c["seed"] + 1000)
        x = torch.randn(c["xshape"], generator=g)
        torch.manual_seed(c["seed"] + 2000)  # the mask draw
        out = model(x.clone(), c["mask_ratio"]) if c["mask_ratio"] > 0 else model(x.clone())
        mask = None
        if isinstance(out, tuple):
            out, mask = out
        tgt = torch.randn(out.shape, generator=g)
        loss = torch.nn.functional.mse_loss(out, tgt)
        loss.backward()
        gn = {n: p.grad.norm().item() for n, p in model.named_parameters() if p.grad is not None}
        torch.save({"cfg": c["cfg"], "seed": c["seed"], "mask_ratio": c["mask_ratio"], "x": x, "out": out.detach(),
                    "mask": mask, "target": tgt, "loss": loss.item(), "grad_norms": gn,
                    "n_keys": len(model.state_dict()), "third_party": ns.third_party, "torch": torch.__version__},
                   OUT / f"{name}.pt")
        print(name, tuple(out.shape), f"loss={loss.item():.6f}", "keys", len(model.state_dict()),
              "mask" if mask is not None else "")


if __name__ == "__main__":
    main()
