#!/usr/bin/env python
"""Stage-by-stage comparison of the sm_100a UNeXt2 path with the fp32 oracle (run on the GPU box).

Usage: python tools/bisect_parity.py [hw] [depth] [batch] [dtype]
Prints rel-L2 of every intermediate (stem, encoder stages, decoder stages, head) and, inside the first stage that
deviates, of every ConvNeXt block.
"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))

import incumbent as I  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nhwc(t):  # oracle NCHW -> NHWC for comparison
    return t.permute(0, 2, 3, 1).contiguous()


def main():
    hw = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 21
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[sys.argv[4] if len(sys.argv) > 4 else "fp16"]
    cfg = dict(I.CFG, in_stack_depth=depth)
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    o, m = I.seeded_pair(cfg)
    o, m = o.to(dev).eval(), m.to(dev).eval()
    torch.manual_seed(1)
    x = torch.randn(batch, 1, depth, hw, hw, device=dev)
    from viscy_b200 import ops
    with torch.no_grad():
        # oracle
        so = o.stem(x)
        fo = o.encoder_stages(so)
        # ours, piece by piece, each piece fed with the ORACLE's input so errors do not accumulate
        m._weight_packs(dt)
        sm = m.stem.forward_cl(x, dt)
        print(f"stem            {rel(sm, nhwc(so)):.3e}")
        cur = nhwc(so).to(dt)
        cur = m.encoder_stages.stem_1.forward_cl(cur)
        cur_o = o.encoder_stages.stem_1(so)
        print(f"stem_1 LN       {rel(cur, nhwc(cur_o)):.3e}")
        for i in range(4):
            st_m = getattr(m.encoder_stages, f"stages_{i}")
            st_o = getattr(o.encoder_stages, f"stages_{i}")
            xin_o = cur_o
            xin_m = nhwc(xin_o).to(dt)
            if not isinstance(st_m.downsample, torch.nn.Identity):
                from viscy_b200 import functional as F
                xin_m = F.ln_conv(xin_m, st_m.downsample[0], st_m.downsample[1])
                xin_o = st_o.downsample(xin_o)
                print(f"enc{i} downsample {rel(xin_m, nhwc(xin_o)):.3e}")
                xin_m = nhwc(xin_o).to(dt)
            for j, (bm, bo) in enumerate(zip(st_m.blocks, st_o.blocks)):
                ym = bm.forward_cl(xin_m)
                yo = bo(xin_o)
                print(f"enc{i} block{j} C={yo.shape[1]} HW={yo.shape[2]}  {rel(ym, nhwc(yo)):.3e}")
                xin_o = yo
                xin_m = nhwc(yo).to(dt)
            cur_o = xin_o
        feats_o = list(fo)
        feats_o.reverse()
        feat_o = feats_o[0]
        from viscy_b200 import functional as F
        for k, (sm_, so_) in enumerate(zip(m.decoder.decoder_stages, o.decoder.decoder_stages)):
            skip_o = feats_o[k + 1] if k + 1 < len(feats_o) else None
            up_o = so_.upsample(feat_o)
            cat_o = torch.cat([up_o, skip_o], 1) if skip_o is not None else up_o
            cat_m = F.pixshuf_cat(nhwc(feat_o).to(dt), None if skip_o is None else nhwc(skip_o).to(dt))
            print(f"dec{k} pixshuf+cat  {rel(cat_m, nhwc(cat_o)):.3e}")
            st_m, st_o = sm_.conv, so_.conv
            xin_o = cat_o
            xin_m = nhwc(cat_o).to(dt)
            if not isinstance(st_m.downsample, torch.nn.Identity):
                xin_m = F.ln_conv(xin_m, st_m.downsample[0], st_m.downsample[1])
                xin_o = st_o.downsample(xin_o)
                print(f"dec{k} ln+conv1x1   {rel(xin_m, nhwc(xin_o)):.3e}")
                xin_m = nhwc(xin_o).to(dt)
            for j, (bm, bo) in enumerate(zip(st_m.blocks, st_o.blocks)):
                ym = bm.forward_cl(xin_m)
                yo = bo(xin_o)
                print(f"dec{k} block{j} C={yo.shape[1]} HW={yo.shape[2]}  {rel(ym, nhwc(yo)):.3e}")
                if rel(ym, nhwc(yo)) > 0.05:
                    # inside the block
                    d_o = bo.conv_dw(xin_o)
                    wt, _ = ops.dw_taps(bm.conv_dw.weight)
                    d_m = ops.dwconv7(xin_m, wt, bm.conv_dw.bias)
                    print(f"    dwconv        {rel(d_m, nhwc(d_o)):.3e}")
                    l_o = bo.norm(d_o)
                    l_m, _, _ = ops.layernorm_fwd(nhwc(d_o).to(dt), bm.norm.weight, bm.norm.bias, 1e-6)
                    print(f"    layernorm     {rel(l_m, nhwc(l_o)):.3e}")
                    h_o = bo.mlp.act(bo.mlp.fc1(l_o))
                    from viscy_b200 import _lib as L
                    M = l_m.numel() // l_m.shape[-1]
                    hp, g_m = F.linear_fwd(nhwc(l_o).to(dt).view(M, -1), bm.mlp.fc1.weight, bm.mlp.fc1.bias, epilogue=L.EPI_GELU_GP)
                    print(f"    fc1+gelu      {rel(g_m.view(nhwc(h_o).shape), nhwc(h_o)):.3e}")
                    y_o = bo.mlp.grn(h_o)
                    z_o = bo.mlp.fc2(y_o)
                    print(f"    (oracle fc2 out norm {z_o.norm().item():.3e}, grn out norm {y_o.norm().item():.3e})")
                xin_o = yo
                xin_m = nhwc(yo).to(dt)
            feat_o = xin_o
        ho = o.head(feat_o)
        hm = m.head.forward_cl(nhwc(feat_o).to(dt))
        print(f"head            {rel(hm, ho):.3e}")
        with torch.autocast('cuda', dtype=dt):
            full = m(x)
        print(f"whole model     {rel(full, o(x)):.3e}")


if __name__ == "__main__":
    main()
