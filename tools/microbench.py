"""Per-kernel micro-benchmarks at the decoder-stage-2 / head sizes of UNeXt2 config 2 (CUDA events, back-to-back launches).
Reports time and effective HBM GB/s (algorithmic bytes / time) per kernel."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from viscy_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(s, device=dev, generator=g)  # noqa: E731
bf = torch.bfloat16


def bench(name, fn, nbytes, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:44s} {ms * 1e3:8.1f} us  {nbytes / ms / 1e6:8.0f} GB/s")


B, H, W, C, C4 = 8, 64, 64, 736, 2944
M = B * H * W
x = rn(B, H, W, C).to(bf)
dy = rn(B, H, W, C).to(bf)
hid = rn(B, H * W, C4).to(bf)
wt = rn(49, C)
bias = rn(C)
gam, bet = rn(C), rn(C)
nb = x.numel() * 2
bench("colreduce sum  [32768 x 736]", lambda: ops.colreduce(x.view(1, M, C), 0), nb)
bench("colreduce sum  [32768 x 2944]", lambda: ops.colreduce(hid.view(1, M, C4), 0), hid.numel() * 2)
bench("colreduce sumsq [8 x 4096 x 2944]", lambda: ops.colreduce(hid, 1), hid.numel() * 2)
x96 = rn(1, M, 96).to(bf)
bench("colreduce sum  [32768 x 96]", lambda: ops.colreduce(x96, 0), x96.numel() * 2)
bench("dwconv7 fwd    [8,64,64,736]", lambda: ops.dwconv7(x, wt, bias), 2 * nb)
bench("dwconv7 dgrad+add", lambda: ops.dwconv7(dy, wt, None, add=x), 3 * nb)
bench("dwconv7 wgrad", lambda: ops.dwconv7_wgrad(x, dy), 2 * nb)
y, mean, rstd = ops.layernorm_fwd(x, gam, bet, 1e-6)
bench("layernorm fwd  [32768 x 736]", lambda: ops.layernorm_fwd(x, gam, bet, 1e-6), 2 * nb)
bench("layernorm bwd  [32768 x 736]", lambda: ops.layernorm_bwd(dy, x, mean, rstd, gam), 5 * nb)
P = rn(8, C, C4)
w2 = rn(C, C4)
s = rn(8, C4)
bench("grn_wgrad_finish [8,736,2944]", lambda: ops.grn_wgrad_finish(P, w2, s, rn(C4), rn(C)), P.numel() * 4 + 2 * w2.numel() * 4)
bench("grn_prepare", lambda: ops.grn_prepare(s.abs() + 0.1, rn(C4), rn(C4), w2, rn(C), bf), w2.numel() * 4 + 8 * w2.numel() * 2)
xs = rn(8, 32, 32, 192).to(bf)
wts, bs = rn(49, 192), rn(192)
bench("dwconv7 fwd    [8,32,32,192]", lambda: ops.dwconv7(xs, wts, bs), 2 * xs.numel() * 2)
xs2 = rn(8, 16, 16, 384).to(bf)
wts2, bs2 = rn(49, 384), rn(384)
bench("dwconv7 fwd    [8,16,16,384]", lambda: ops.dwconv7(xs2, wts2, bs2), 2 * xs2.numel() * 2)
bench("dwconv7 wgrad  [8,16,16,384]", lambda: ops.dwconv7_wgrad(xs2, xs2), 2 * xs2.numel() * 2)
dec = rn(8, 64, 64, 736).to(bf)
u = ops.head_shuffle_pool_fwd(dec, 23, True, 8)
bench("head shuffle+pool fwd", lambda: ops.head_shuffle_pool_fwd(dec, 23, True, 8), dec.numel() * 2 + u.numel() * 2)
bench("head shuffle+pool bwd", lambda: ops.head_shuffle_pool_bwd(u, 184, True), dec.numel() * 2 + u.numel() * 2)
z = rn(8, 21 * 128 * 128, 32).to(bf)
mz, rz = ops.instnorm_stats(z)
bench("instnorm stats [8 x 344064 x 32]", lambda: ops.instnorm_stats(z), z.numel() * 2)
al, W1, b1 = torch.tensor([0.25], device=dev), rn(8, 32) * 0.2, rn(8)
out = ops.head_tail_fwd(z, mz, rz, al, W1, b1, 21, 128, 128)
bench("head tail fwd", lambda: ops.head_tail_fwd(z, mz, rz, al, W1, b1, 21, 128, 128), z.numel() * 2 + out.numel() * 2)
bench("head tail bwd (2 phases + dW1 GEMM)", lambda: ops.head_tail_bwd(z, mz, rz, al, W1, out, 21, 128, 128),
      (3 * z.numel() + 2 * out.numel() + 2 * z.numel() + out.numel()) * 2)
xin = rn(8, 1, 21, 256, 256)
bench("stem patchify", lambda: ops.stem_patchify(xin, 4, 4, bf), xin.numel() * 4 + xin.numel() * 2)
