"""FullyConvolutionalMAE through the sm_100a kernels: the new kernels (row gather / scatter, pixel-shuffle + pad + pool
head) vs torch's fp32 ops, and the whole model (dense, sparse-masked, 2-D) vs the reference-generated golden vectors and,
element-wise for every parameter gradient, vs the fp32 CPU mirror (== the reference, tests/test_fcmae_cpu.py)."""
from pathlib import Path

import pytest
import torch
import torch.nn.functional as TF

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def perturb(model):
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "grn" in n or n.endswith("bias"):
                p.normal_(0, 0.2)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,h,w,Cq,r,pool", [(2, 8, 8, 5, 4, True), (1, 5, 19, 6, 2, True), (2, 16, 16, 10, 4, False),
                                             (1, 3, 4, 8, 1, True), (1, 64, 64, 42, 4, True)])
def test_shuffle_pool_vs_torch(cuda, dtype, B, h, w, Cq, r, pool):
    from viscy_b200 import functional as VF
    g = torch.Generator(device=cuda).manual_seed(1)
    x = torch.randn(B, Cq * r * r, h, w, device=cuda, generator=g).to(dtype)
    xc = x.permute(0, 2, 3, 1).contiguous().requires_grad_(True)
    y = VF.shuffle_pool(xc, r, pool)
    xf = x.float().requires_grad_(True)
    ref = TF.pixel_shuffle(xf, r)
    if pool:
        ref = TF.avg_pool2d(TF.pad(ref, (r - 1, 0, r - 1, 0)), r, stride=1)
    tol = 1e-3 if dtype == torch.float16 else 6e-3
    assert y.shape == ref.shape and rel(y, ref) < tol
    dy = torch.randn(ref.shape, device=cuda, generator=g).to(dtype)
    ref.backward(dy.float())
    y.backward(dy)
    assert rel(xc.grad.permute(0, 3, 1, 2), xf.grad) < tol


def test_rows_select_gather_scatter(cuda):
    from viscy_b200 import functional as VF
    g = torch.Generator(device=cuda).manual_seed(2)
    B, H, W, C, keep = 3, 8, 8, 40, 24
    unmasked = torch.zeros(B, H * W, dtype=torch.bool, device=cuda)
    for b in range(B):
        unmasked[b, torch.randperm(H * W, device=cuda, generator=g)[:keep]] = True
    mi = VF.MaskIndex(unmasked.view(B, H, W), keep)
    x = torch.randn(B * H * W, C, device=cuda, generator=g).half().requires_grad_(True)
    rows = VF.RowsSelectFn.apply(x, None, mi.idx, mi.inv)
    assert torch.equal(rows, x.detach().view(B, H * W, C)[unmasked])  # masked_patchify's row order
    base = torch.randn(B * H * W, C, device=cuda, generator=g).half().requires_grad_(True)
    out = VF.RowsSelectFn.apply(rows, base, mi.inv, mi.idx)
    ref = base.detach().float().clone().view(B, H * W, C)
    ref[unmasked] += rows.detach().float()
    assert rel(out, ref.view(-1, C)) < 1e-3
    dy = torch.randn_like(out)
    out.backward(dy)
    assert torch.equal(base.grad, dy)
    gx = torch.zeros_like(dy).view(B, H * W, C)
    gx[unmasked] = dy.view(B, H * W, C)[unmasked]
    assert torch.equal(x.grad, gx.view(-1, C))


def _run_golden(cuda, name, dtype):
    from viscy_b200 import FullyConvolutionalMAE
    g = torch.load(GOLD / f"{name}.pt", weights_only=False)
    torch.manual_seed(g["seed"])
    ref = FullyConvolutionalMAE(**g["cfg"])
    perturb(ref)
    m = FullyConvolutionalMAE(**g["cfg"])
    m.load_state_dict(ref.state_dict())
    m = m.to(cuda)
    # fp32 CPU mirror: element-wise gradient reference
    torch.manual_seed(g["seed"] + 2000)
    o = ref(g["x"].clone(), g["mask_ratio"]) if g["mask_ratio"] > 0 else ref(g["x"].clone())
    o = o[0] if isinstance(o, tuple) else o
    TF.mse_loss(o, g["target"]).backward()
    unmasked = None
    if g["mask"] is not None:
        s = m.encoder.total_stride
        unmasked = ~g["mask"][:, :, ::s, ::s].to(cuda)
    with torch.autocast("cuda", dtype=dtype):
        out, mask = m._forward_sm100(g["x"].to(cuda), g["mask_ratio"], unmasked)
        loss = TF.mse_loss(out.float(), g["target"].to(cuda))
    scale = 1024.0 if dtype == torch.float16 else 1.0
    (loss * scale).backward()
    return g, ref, m, out, mask, loss, scale


@pytest.mark.parametrize("dtype,ftol,gtol", [(torch.float16, 2e-3, 3e-2), (torch.bfloat16, 1.2e-2, 1e-1)])
@pytest.mark.parametrize("name", ["fcmae_dense", "fcmae_masked", "fcmae_2d"])
def test_fcmae_against_reference_golden(cuda, name, dtype, ftol, gtol):
    g, ref, m, out, mask, loss, scale = _run_golden(cuda, name, dtype)
    assert out.shape == g["out"].shape and out.dtype == dtype
    if g["mask"] is not None:
        assert torch.equal(mask.cpu(), g["mask"])
    e = rel(out.float().cpu(), g["out"])
    print(f"\n[{name} {dtype}] forward rel-L2 vs reference golden {e:.3e}")
    assert e < ftol
    assert abs(loss.item() - g["loss"]) < 3 * ftol * abs(g["loss"])
    refg = dict(ref.named_parameters())
    worst = []
    for n, p in m.named_parameters():
        gr = refg[n].grad
        if gr is None:
            assert p.grad is None, n
            continue
        assert p.grad is not None, n
        if n == "head.conv.0.conv.bias" or gr.norm().item() < 1e-7:  # bias before InstanceNorm: analytically zero
            continue
        worst.append((rel(p.grad.cpu() / scale, gr), n))
        gn = g["grad_norms"][n]
        assert abs(p.grad.float().norm().item() / scale - gn) < max(gtol, 5e-2) * gn, n
    worst.sort(reverse=True)
    print("worst grads:", [(f"{w:.2e}", n) for w, n in worst[:5]])
    assert worst[0][0] < gtol, worst[:3]


def test_fcmae_mask_draw_and_eval(cuda):
    """forward() draws its own mask on the device (same generator recipe as the reference): masked pixels of the encoder
    features are zero, the returned mask has the input's resolution; eval / no_grad works."""
    from viscy_b200 import FullyConvolutionalMAE
    torch.manual_seed(0)
    m = FullyConvolutionalMAE(1, 1, in_stack_depth=5, encoder_blocks=(1, 1, 1, 1), dims=(32, 64, 128, 256)).to(cuda).eval()
    x = torch.randn(2, 1, 5, 128, 128, device=cuda)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y, mask = m(x, 0.5)
        feats, mask2 = m.encoder.forward_cl(x, torch.bfloat16, 0.75)
    assert y.shape == (2, 1, 5, 128, 128) and mask.shape == (2, 1, 128, 128) and mask.dtype == torch.bool
    assert mask.float().mean().item() == 0.5 and torch.isfinite(y).all()
    f0 = feats[0]  # [B, 32, 32, C] NHWC
    up = mask2[:, 0, ::4, ::4]
    assert f0[up].abs().max().item() == 0.0 and f0[~up].abs().max().item() > 0.0
