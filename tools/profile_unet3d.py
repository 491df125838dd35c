"""One Unet3d(3,3,4,32) 128^3 fp16 training step (BASELINE config 5) between cudaProfilerStart/Stop."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from viscy_b200 import Unet3d  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m = Unet3d(3, 3, 4, 32).to(dev)
from viscy_b200.optim import AdamW  # noqa: E402

opt = AdamW(m.parameters(), lr=1e-3)
x = torch.randn(1, 3, S, S, S, device=dev)
y = torch.randn(1, 3, S, S, S, device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.float16):
        loss = torch.nn.functional.mse_loss(m(x).float(), y)
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
