"""One ContrastiveEncoder training step (BASELINE config 4: 2 views x 64 x (2,15,224,224) bf16 + NT-Xent) between
cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from viscy_b200 import ContrastiveEncoder  # noqa: E402
from viscy_b200.loss import NTXentLoss  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = ContrastiveEncoder("convnext_tiny", in_channels=2, in_stack_depth=15).to(dev)
opt = torch.optim.AdamW(m.parameters(), lr=2e-4, fused=True)
a = torch.randn(64, 2, 15, 224, 224, device=dev)
p = torch.randn(64, 2, 15, 224, 224, device=dev)
labels = torch.cat([torch.arange(64), torch.arange(64)]).to(dev)
crit = NTXentLoss(temperature=0.07)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        _, pa = m(a)
        _, pp = m(p)
    loss = crit(torch.cat([pa, pp]).float(), labels)
    loss.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
