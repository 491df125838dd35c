"""The four decoder-stage-2 GEMM launches of one block (M=32768, 736<->2944), for `ncu --set full` captures."""
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
calls = [
    lambda: ops.gemm(a_c, w1, bias=b_c4, epilogue=LL.EPI_GELU_GP, out=o_c4a, out2=o_c4b),
    lambda: ops.gemm(a_c4, w2s, bias=b_c, residual=a_c, b_batch_rows=R, out=o_c),
    lambda: ops.gemm(a_c, w2t, epilogue=LL.EPI_DGELU_GRN, aux=a_c4, aux2=o_c4b, tvec=tv, svec=sv, rows_per_sample=R, out=o_c4a),
    lambda: ops.gemm(a_c4, w1t, out=o_c),
    lambda: ops.gemm(a_c, a_c4, mn_major=True, epilogue=LL.EPI_F32, k_splits=8, split_slabs=True),
    lambda: ops.gemm(a_c4, a_c, mn_major=True, epilogue=LL.EPI_F32, k_splits=4),
]
for _ in range(2):
    for f in calls:
        f()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for f in calls:
    f()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
