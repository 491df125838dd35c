"""In-situ kernel times of the CUDA-graph replayed UNeXt2 step via torch.profiler (kineto/CUPTI)."""
import collections
import re
import sys
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from viscy_b200 import UNeXt2  # noqa: E402
from viscy_b200.graphs import GraphedStep  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
model = UNeXt2(**bench.CFG).to(dev)
opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True, capturable=True)
x = torch.randn((bench.BATCH, *bench.SHAPE_IN), device=dev)
y = torch.randn((bench.BATCH, *bench.SHAPE_OUT), device=dev)


def step(a, b):
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = torch.nn.functional.mse_loss(model(a).float(), b)
    loss.backward()
    opt.step()
    return loss


g = GraphedStep(step, (x, y))
for _ in range(3):
    g(x, y)
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        g(x, y)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", e.name)[:80]
        agg[name][0] += 1
        agg[name][1] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
        tot += agg[name][1] * 0
if len(sys.argv) > 1:
    sel = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and sys.argv[1] in e.name]
    sel = sel[: len(sel) // N]
    print(sys.argv[1], [round(e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, 1) for e in sel])
tot = sum(v[1] for v in agg.values())
print(f"sum of kernel time per step: {tot / N:.1f} us")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t / N:10.1f} us {100 * t / tot:5.1f}% n={c // N:4d}  {n}")
