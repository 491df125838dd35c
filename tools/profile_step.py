"""One UNeXt2 training step (BASELINE config 2) between cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from viscy_b200 import UNeXt2  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = UNeXt2(**bench.CFG).to(dev)
from viscy_b200.optim import AdamW  # noqa: E402

opt = AdamW(model.parameters(), lr=1e-3)
B = int(sys.argv[1]) if len(sys.argv) > 1 else bench.BATCH
x = torch.randn((B, *bench.SHAPE_IN), device=dev)
y = torch.randn((B, *bench.SHAPE_OUT), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.mse_loss(model(x).float(), y)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
