"""CUDA-event timing of the decoder-stage-2 GEMM launches of UNeXt2 config 2 (M=32768, 736<->2944)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from viscy_b200 import _lib as LL, ops  # noqa: E402

dev = torch.device("cuda:0")
B, C, C4, R = 8, 736, 2944, 4096
M = B * R
g = torch.Generator(device=dev).manual_seed(7)
rn = lambda *sh: torch.randn(sh, device=dev, generator=g)  # noqa: E731
a_c, a_c4 = rn(M, C).bfloat16(), rn(M, C4).bfloat16()
w1, w2t = (rn(C4, C) * 0.03).bfloat16(), (rn(C4, C) * 0.03).bfloat16()
w2s, w1t = (rn(B * C, C4) * 0.02).bfloat16(), (rn(C, C4) * 0.02).bfloat16()
b_c4, b_c = rn(C4), rn(C)
sv, tv = rn(B, C4) * 0.1 + 1.0, rn(B, C4) * 0.1
o_c4a, o_c4b, o_c = torch.empty_like(a_c4), torch.empty_like(a_c4), torch.empty_like(a_c)
calls = {
    "fc1+gelu_gp": lambda: ops.gemm(a_c, w1, bias=b_c4, epilogue=LL.EPI_GELU_GP, out=o_c4a, out2=o_c4b),
    "fc2+residual": lambda: ops.gemm(a_c4, w2s, bias=b_c, residual=a_c, b_batch_rows=R, out=o_c),
    "dgrad_fc2+grn_gelu_bwd": lambda: ops.gemm(a_c, w2t, epilogue=LL.EPI_DGELU_GRN, aux=a_c4, aux2=o_c4b, tvec=tv, svec=sv,
                                                rows_per_sample=R, out=o_c4a),
    "dgrad_fc1": lambda: ops.gemm(a_c4, w1t, out=o_c),
    "wgrad_fc2_slabs": lambda: ops.gemm(a_c, a_c4, mn_major=True, epilogue=LL.EPI_F32, k_splits=8, split_slabs=True),
    "wgrad_fc1": lambda: ops.gemm(a_c4, a_c, mn_major=True, epilogue=LL.EPI_F32, k_splits=4),
}
tot = 0.0
for name, fn in calls.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tot += ms
    print(f"{name:26s} {ms * 1e3:7.1f} us  {2.0 * M * C * C4 / ms / 1e9:7.0f} TFLOP/s")
print(f"mean {tot / len(calls) * 1e3:.1f} us")
