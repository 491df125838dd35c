"""The HBM / FMA-bound tail kernels of one ConvNeXt block at the decoder-stage-2 shape of BASELINE config 2 (B=8, 64x64,
C=736, C4=2944) and at one small-map shape, launched back to back for `ncu --set full -k regex:...`.
Usage: python tools/profile_tail.py [which ...]   which in {dw, ln, grn, red}"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from viscy_b200 import ops  # noqa: E402

which = set(sys.argv[1:]) or {"dw", "ln", "grn", "red"}
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(s, device=dev, generator=g)  # noqa: E731
bf = torch.bfloat16
for (B, H, C) in ((8, 64, 736), (8, 16, 384)):
    C4, M, R = 4 * C, B * H * H, H * H
    x, dy = rn(B, H, H, C).to(bf), rn(B, H, H, C).to(bf)
    w = rn(C, 1, 7, 7) * 0.1
    wt, wtf = ops.dw_pack(w)
    bias = rn(C)
    for _ in range(2):
        if "dw" in which:
            ops.dwconv7(x, wt, bias)
            ops.dwconv7(dy, wtf, None, add=x)
            ops.dwconv7_wgrad(x, dy)
        if "ln" in which:
            gm, bt = rn(C), rn(C)
            gbuf = torch.empty((M, C4 + 16), device=dev, dtype=bf)
            lb, mean, rstd = ops.layernorm_fwd(x, gm, bt, 1e-6, ones=True, ones2=gbuf, ones2_col=C4)
            ops.layernorm_bwd(dy, x, mean, rstd, gm)
        if "grn" in which or "red" in which:
            hid = rn(M, C4 + 16).to(bf)
            sumsq = ops.colreduce(hid.view(B, R, C4 + 16), 1, width=C4)
            if "grn" in which:
                w2 = rn(C, C4) * 0.02
                s, w2s, b2e = ops.grn_prepare(sumsq, rn(C4), rn(C4), w2, rn(C), bf)
                P = rn(B, C, C4 + 16)
                dW2, S1, dbg, db2 = ops.grn_wgrad_finish(P, w2, s, rn(C4), None)
                t = torch.empty_like(S1)
                dgw = torch.zeros(C4, device=dev)
                ops._call("vb200_grn_coef_bwd", ops._p(sumsq), ops._p(S1.contiguous()), ops._p(rn(C4)), ops._p(t), ops._p(dgw), B, C4,
                          ops.C.c_float(1e-6))
torch.cuda.synchronize()
print("done")
