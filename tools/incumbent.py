#!/usr/bin/env python
"""The GPU incumbent and the parity yardstick (SURVEY.md 8(d) last bullet).

The "reference GPU path" is the reference math (oracle/models.py: the reference's composition code over restated
timm/monai blocks, plain torch modules) moved to CUDA under stock `torch.autocast` -> cuDNN / cuBLAS kernels.
This tool
  * times it at BASELINE config 2 (B=8, 21x256x256, fwd + MSE + bwd + fused AdamW; contiguous and channels_last) and
  * measures, per tensor, how far stock autocast and the sm_100a path each land from the fp32 oracle (output and
    every parameter gradient, rel-L2 of the element-wise difference) on the same seeded weights and inputs.

The oracle is test infrastructure: it is the baseline that is measured here, never part of the product path.
Writes one JSON document (stdout, and --out FILE).
"""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

CFG = dict(in_channels=1, out_channels=2, in_stack_depth=21, backbone="convnextv2_tiny",
           stem_kernel_size=(7, 4, 4), head_pool=True, head_expansion_ratio=4)


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def seeded_pair(cfg: dict, seed: int = 0):
    """fp32 oracle + sm_100a module with the same weights; zero-initialised GRN / biases get values so that every
    path carries signal."""
    from oracle import models as OM
    from viscy_b200 import UNeXt2
    torch.manual_seed(seed)
    o = OM.UNeXt2(**cfg)
    with torch.no_grad():
        for n, p in o.named_parameters():
            if "grn" in n or n.endswith("bias"):
                p.normal_(0, 0.2)
    m = UNeXt2(**cfg)
    m.load_state_dict(o.state_dict())
    return o, m


def run_fp32(o, x, tgt):
    o.zero_grad(set_to_none=True)
    out = o(x)
    torch.nn.functional.mse_loss(out, tgt).backward()
    return out.detach(), {n: p.grad.detach().clone() for n, p in o.named_parameters()}


def run_autocast(model, x, tgt, dtype, loss_scale: float = 1.0):
    """`loss_scale`: static stand-in for torch.amp.GradScaler (fp16 gradients of a mean-reduced loss underflow
    without it, for the stock path and for ours alike); gradients are returned unscaled."""
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=dtype):
        out = model(x)
        loss = torch.nn.functional.mse_loss(out.float(), tgt)
    (loss * loss_scale).backward()
    return out.detach().float(), {n: p.grad.detach().float() / loss_scale for n, p in model.named_parameters()}


def yardstick(cfg: dict, batch: int, hw: int, dtype: torch.dtype, seed: int = 0, fp32_on: str = "cpu",
              loss_scale: float = 1.0):
    """-> dict(out=(ours, stock), grads={name: (ours, stock)}): rel-L2 errors vs the fp32 oracle."""
    import copy
    dev = torch.device("cuda:0")
    o, m = seeded_pair(cfg, seed)
    torch.manual_seed(seed + 1)
    D = cfg["in_stack_depth"]
    x = torch.randn(batch, cfg["in_channels"], D, hw, hw)
    tgt = torch.randn(batch, cfg["out_channels"], D, hw, hw)
    stock = copy.deepcopy(o).to(dev)
    if fp32_on == "cuda":
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        o = o.to(dev)
        ref_out, ref_g = run_fp32(o, x.to(dev), tgt.to(dev))
    else:
        ref_out, ref_g = run_fp32(o, x, tgt)
    s_out, s_g = run_autocast(stock, x.to(dev), tgt.to(dev), dtype, loss_scale)
    m = m.to(dev)
    m_out, m_g = run_autocast(m, x.to(dev), tgt.to(dev), dtype, loss_scale)
    res = {"out": (rel(m_out, ref_out), rel(s_out, ref_out)), "grads": {}}
    for n, g in ref_g.items():
        if g.norm().item() < 1e-12:
            continue
        res["grads"][n] = (rel(m_g[n], g), rel(s_g[n], g), g.norm().item())
    return res


def time_incumbent(batch: int, hw: int, steps: int, warmup: int, channels_last: bool, dtype=torch.bfloat16):
    from oracle import models as OM
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = OM.UNeXt2(**CFG).to(dev)
    if channels_last:  # 4-D (2-D conv) weights only: the stem / head Conv3d weights are rank 5
        for mod in model.modules():
            if isinstance(mod, torch.nn.Conv2d):
                mod.weight.data = mod.weight.data.contiguous(memory_format=torch.channels_last)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=True)
    x = torch.randn(batch, 1, 21, hw, hw, device=dev)
    y = torch.randn(batch, 2, 21, hw, hw, device=dev)
    torch.backends.cudnn.benchmark = True

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=dtype):
            loss = torch.nn.functional.mse_loss(model(x).float(), y)
        loss.backward()
        opt.step()

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del model, opt, x, y
    torch.cuda.empty_cache()
    return {"ms_per_step": ms, "samples_per_s": batch * 1e3 / ms, "channels_last": channels_last,
            "mode": "eager torch.autocast + cuDNN/cuBLAS, fused AdamW, cudnn.benchmark"}


def summarize(res):
    rows = sorted(((o / max(s, 1e-12), n, o, s) for n, (o, s, _) in res["grads"].items()), reverse=True)
    import statistics
    return {
        "out_ours": res["out"][0], "out_stock": res["out"][1],
        "grad_median_ours": statistics.median(o for _, _, o, _ in rows),
        "grad_median_stock": statistics.median(s for _, _, _, s in rows),
        "grad_worst_ours": max((o, n) for _, n, o, _ in rows),
        "grad_worst_stock": max((s, n) for _, n, _, s in rows),
        "worst_ratio": [(f"{r:.2f}", n, f"{o:.2e}", f"{s:.2e}") for r, n, o, s in rows[:8]],
        "n_ratio_gt_1.5": sum(1 for r, *_ in rows if r > 1.5),
        "n_tensors": len(rows),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-time", action="store_true")
    ap.add_argument("--hw", type=int, default=256)
    ap.add_argument("--batch", type=int, default=2)
    args = ap.parse_args()
    doc = {}
    if not args.no_time:
        for cl in (False, True):
            try:
                doc[f"incumbent_bf16_b8{'_channels_last' if cl else ''}"] = time_incumbent(8, 256, 10, 3, cl)
            except Exception as exc:
                doc[f"incumbent_bf16_b8{'_channels_last' if cl else ''}"] = {"error": f"{type(exc).__name__}: {exc}"}
            print(json.dumps(doc), flush=True)
    for name, dt in (("bf16", torch.bfloat16), ("fp16", torch.float16)):
        t0 = time.time()
        res = yardstick(CFG, args.batch, args.hw, dt, loss_scale=65536.0 if dt == torch.float16 else 1.0)
        doc[f"yardstick_{name}_b{args.batch}_{args.hw}"] = summarize(res)
        doc[f"yardstick_{name}_b{args.batch}_{args.hw}"]["seconds"] = time.time() - t0
        print(json.dumps(doc[f"yardstick_{name}_b{args.batch}_{args.hw}"]), flush=True)
    if args.out:
        Path(args.out).write_text(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
