"""Level-0 conv launches of Unet3d config 5 (128^3, 32 -> 32 channels, fp16) for `ncu --set full` captures:
patch-form forward (gemm_kernel<64, MODE_CONVKP, ...>) and patch-form weight gradient (conv3d_wgrad_kh3_kernel)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from viscy_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
S, ci, co = 128, 32, 32
g = torch.Generator(device=dev).manual_seed(3)
x = torch.randn((1, S, S, S, ci), device=dev, generator=g).half()
dy = torch.randn((1, S, S, S, co), device=dev, generator=g).half()
w = (torch.randn((co, 27 * ci), device=dev, generator=g) * 0.02).half()
y = torch.empty((1, S, S, S, co), device=dev, dtype=torch.float16)
# interior level: 128 -> 128 channels at 32^3 (patch form, 128-wide tile)
x2 = torch.randn((1, 32, 32, 32, 128), device=dev, generator=g).half()
w2 = (torch.randn((128, 27 * 128), device=dev, generator=g) * 0.02).half()
y2 = torch.empty((1, 32, 32, 32, 128), device=dev, dtype=torch.float16)
calls = [lambda: ops.conv3d_igemm(x, w, None, (3, 3, 3), (1, 1, 1), out=y),
         lambda: ops.conv3d_wgrad_kh3(x, dy, (3, 3, 3), (1, 1, 1)),
         lambda: ops.conv3d_igemm(x2, w2, None, (3, 3, 3), (1, 1, 1), out=y2)]
for _ in range(2):
    for f in calls:
        f()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for f in calls:
    f()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
