"""The kernels added in round 2, launched at BASELINE-config shapes for `ncu --set full`:
  head  : streaming head-tail backward (B=8, 21 x 128 x 128 voxels, Cmid 32)
  ssim  : MixedLoss levels on the config-2 output (B=8, 2 x 21 x 256 x 256; bf16 prediction, fp32 target), forward + backward
  gn    : GroupNorm(8) + SiLU apply / backward on a Unet3d level (1 x 128^3 x 32, fp16)
  fcmae : row gather / scatter at FCMAE stage 0 (B=8, 64 x 64 x 96, half the rows kept), shuffle-pool head (r = 4)
  blend : crop + blend of one Z window (B=1, 2 x 21 x 2048 x 2048 fp32 volume slab from a bf16 window)
Usage: python tools/profile_r2_kernels.py [which ...]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from viscy_b200 import functional as VF  # noqa: E402
from viscy_b200 import losses, ops, predict  # noqa: E402

which = set(sys.argv[1:])
if not which - {"once"}:
    which |= {"head", "ssim", "gn", "fcmae", "blend"}
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(s, device=dev, generator=g)  # noqa: E731
bf = torch.bfloat16
for _ in range(1 if "once" in which else 2):
    if "head" in which:
        B, Dz, H, W = 8, 21, 128, 128
        z = rn(B, Dz * H * W, 32).to(bf)
        dout = rn(B, 2, Dz, 2 * H, 2 * W).to(bf)
        mean, rstd = ops.instnorm_stats(z)
        ops.head_tail_bwd(z, mean, rstd, torch.tensor([0.25], device=dev), rn(8, 32) * 0.2, dout, Dz, H, W)
        del z, dout
    if "ssim" in which:
        p = torch.rand((8, 2, 21, 256, 256), device=dev, generator=g).to(bf).requires_grad_(True)
        t = torch.rand((8, 2, 21, 256, 256), device=dev, generator=g)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            losses.MixedLoss()(p, t).backward()
        del p, t
    if "gn" in which:
        x = rn(1, 128, 128, 128, 32).half().requires_grad_(True)
        gn = torch.nn.GroupNorm(8, 32).to(dev)
        VF.groupnorm_act_cl(x, gn, "silu").backward(rn(1, 128, 128, 128, 32).half())
        del x
    if "fcmae" in which:
        B, H, C = 8, 64, 96
        keep = torch.zeros(B, H * H, dtype=torch.bool, device=dev)
        keep[:, ::2] = True
        mi = VF.MaskIndex(keep.view(B, H, H), H * H // 2)
        x = rn(B * H * H, C).to(bf)
        rows = ops.rows_select(x, mi.idx)
        ops.rows_select(rows, mi.inv, base=x)
        dec = rn(8, 64, 64, 2 * 21 * 16).to(bf)
        y = ops.shuffle_pool_fwd(dec, 4, True)
        ops.shuffle_pool_bwd(y, 4, True)
        del x, rows, dec, y
    if "blend" in which:
        out = torch.zeros((1, 2, 32, 2048, 2048), device=dev)
        pred = rn(1, 2, 21, 2048, 2048).to(bf)
        predict.blend_window_(out, pred, 3)
        del out, pred
torch.cuda.synchronize()
print("done")
