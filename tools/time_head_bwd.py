"""A/B timing of the UNeXt2 head tail backward at the BASELINE geometry (B=8, 21x128x128 voxels, Cmid 32, Co4 8):
streaming kernels vs the generic two-phase kernels + dW1 GEMM.  CUDA events, 10 reps."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from viscy_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, Dz, H, W, Cmid, Co = 8, 21, 128, 128, 32, 2
g = torch.Generator(device=dev).manual_seed(0)
z = torch.randn((B, Dz * H * W, Cmid), device=dev, generator=g).bfloat16()
alpha = torch.tensor([0.25], device=dev)
w1 = torch.randn((Co * 4, Cmid), device=dev, generator=g) * 0.2
dout = torch.randn((B, Co, Dz, 2 * H, 2 * W), device=dev, generator=g).bfloat16()
mean, rstd = ops.instnorm_stats(z)


def timeit(fn, n=10, warm=3):
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


res = {}
for stream in (True, False):
    ops.HEAD_BWD_STREAM = stream
    res[stream] = ops.head_tail_bwd(z, mean, rstd, alpha, w1, dout, Dz, H, W)
    us = timeit(lambda: ops.head_tail_bwd(z, mean, rstd, alpha, w1, dout, Dz, H, W))
    mb = (z.numel() * 2 * 3 + dout.numel() * 2 * 2) / 1e6
    print(f"head tail backward, {'streaming' if stream else 'generic'}: {us:.1f} us  ({mb / us * 1e-3 * 1e3:.0f} GB/s of the "
          f"{mb:.0f} MB algorithmic traffic: z twice + dout twice + dz)")
names = ("dz", "dW1", "db1", "dalpha", "dbz")
for n, a, b in zip(names, res[True], res[False]):
    print(n, "rel diff streaming vs generic:", ((a.float() - b.float()).norm() / b.float().norm()).item())
