"""CUDA-event timing of the depthwise 7x7 kernels (forward, data gradient + residual add, weight gradient) at the UNeXt2
config-2 shapes.  A/B two builds with VB200_LIB=/path/to/other/libviscy_b200.so."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from viscy_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


tot = 0.0
for (B, H, C, reps) in ((8, 64, 736, 2), (8, 64, 96, 3), (8, 32, 192, 5), (8, 16, 384, 11)):
    x = torch.randn((B, H, H, C), device=dev, generator=g).bfloat16()
    dy = torch.randn((B, H, H, C), device=dev, generator=g).bfloat16()
    wt, wtf = ops.dw_pack(torch.randn((C, 1, 7, 7), device=dev, generator=g) * 0.1)
    bias = torch.randn((C,), device=dev, generator=g)
    f = timeit(lambda: ops.dwconv7(x, wt, bias))
    d = timeit(lambda: ops.dwconv7(dy, wtf, None, add=x))
    w = timeit(lambda: ops.dwconv7_wgrad(x, dy))
    tot += reps * (f + d + w)
    print(f"B={B} {H}x{H} C={C}: fwd {f:.1f} us, dgrad+add {d:.1f} us, wgrad {w:.1f} us   (x{reps} blocks per step)")
print(f"per-step total {tot:.0f} us")
