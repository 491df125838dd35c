"""Secondary BASELINE configs at full size (eager): config 4 ContrastiveEncoder 2x64x(2,15,224,224) bf16 + NT-Xent,
config 5 Unet3d(3,3,4,32) 1x(3,128,128,128) fp16.  Prints ms/step (CUDA events)."""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle.models import ntxent  # loss only (checker-side torch restatement of PML NT-Xent)
from viscy_b200 import ContrastiveEncoder, Unet3d

dev = torch.device("cuda:0")


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


which = sys.argv[1:] or ["c4", "c5"]
if "c5" in which:
    torch.manual_seed(0)
    m = Unet3d(3, 3, 4, 32).to(dev)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-3, fused=True)
    x = torch.randn(1, 3, 128, 128, 128, device=dev)
    y = torch.randn(1, 3, 128, 128, 128, device=dev)

    def step5():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.float16):
            loss = torch.nn.functional.mse_loss(m(x).float(), y)
        loss.backward()
        opt.step()

    ms = timeit(step5)
    print(f"config5 Unet3d 128^3 fp16 B=1: {ms:.1f} ms/step = {1000 / ms:.2f} samples/s; "
          f"{3.696 / (ms * 1e-3):.0f} TFLOP/s algorithmic; peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    del m, opt, x, y
    torch.cuda.empty_cache()
if "c4" in which:
    torch.manual_seed(0)
    m = ContrastiveEncoder("convnext_tiny", in_channels=2, in_stack_depth=15).to(dev)
    opt = torch.optim.AdamW(m.parameters(), lr=2e-4, fused=True)
    a = torch.randn(64, 2, 15, 224, 224, device=dev)
    p = torch.randn(64, 2, 15, 224, 224, device=dev)
    labels = torch.cat([torch.arange(64), torch.arange(64)]).to(dev)

    def step4():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            _, pa = m(a)
            _, pp = m(p)
        loss = ntxent(torch.cat([pa, pp]).float(), labels, 0.07)
        loss.backward()
        opt.step()

    ms = timeit(step4)
    print(f"config4 ContrastiveEncoder 2x64 bf16: {ms:.1f} ms/step = {128000 / ms:.0f} samples/s; "
          f"{3.448 / (ms * 1e-3):.0f} TFLOP/s algorithmic; peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
