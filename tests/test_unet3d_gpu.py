"""Unet3d (F-Net) through the sm_100a kernels vs the golden vectors from the reference's own code."""
from pathlib import Path

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 3e-3), (torch.bfloat16, 2.5e-2)])
def test_unet3d_against_reference_golden(cuda, dtype, tol):
    from viscy_b200 import Unet3d
    g = torch.load(GOLD / "unet3d.pt", weights_only=False)
    torch.manual_seed(g["seed"])
    m = Unet3d(**g["cfg"]).to(cuda)
    with torch.autocast("cuda", dtype=dtype):
        out = m(g["x"].to(cuda))
        loss = F.mse_loss(out.float(), g["targets"][0].to(cuda))
    loss.backward()
    e = rel(out.float().cpu(), g["outs"][0])
    print(f"\n[{dtype}] Unet3d forward rel-L2 vs reference golden {e:.3e}")
    assert out.shape == g["outs"][0].shape and e < tol
    assert abs(loss.item() - g["loss"]) < 3 * tol * abs(g["loss"])
    bad = []
    for n, p in m.named_parameters():
        ref = g["grad_norms"][n]
        if ref < 1e-5:  # conv biases in front of BatchNorm: analytically zero
            continue
        got = p.grad.float().norm().item()
        if abs(got - ref) > (0.05 if dtype == torch.float16 else 0.25) * ref:
            bad.append((n, got, ref))
    assert not bad, bad[:6]
    assert int(m.bottleneck.block.block1.norm.num_batches_tracked) == 1


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 4e-3), (torch.bfloat16, 3e-2)])
def test_unet3d_depth4_against_reference_golden(cuda, dtype, tol):
    """The config-5 architecture Unet3d(3, 3, depth 4, mult_chan 32) on a 64^3 volume (fixture from the reference's own
    code): 32-channel 64^3 levels run the patch-form / resident-filter kernels, 64 ... 512 channels the implicit-GEMM forms,
    transposed convs the parity-class form.  Output vs the golden, every weight gradient element-wise vs the fp32 mirror,
    bounded by 1.5 x the 16-bit-storage yardstick (tests/yardstick.py)."""
    from viscy_b200 import Unet3d
    from yardstick import add_storage_rounding, gradient_ratios
    g = torch.load(GOLD / "unet3d_d4.pt", weights_only=False)
    gen = torch.Generator().manual_seed(g["seed"] + 1000)
    x = torch.randn(g["xshape"], generator=gen)
    tgt = torch.randn(g["out"].shape, generator=gen)
    torch.manual_seed(g["seed"])
    ref = Unet3d(**g["cfg"])
    m = Unet3d(**g["cfg"])
    m.load_state_dict(ref.state_dict())
    m = m.to(cuda)
    F.mse_loss(ref(x), tgt).backward()
    torch.manual_seed(g["seed"])
    emu = add_storage_rounding(Unet3d(**g["cfg"]), dtype, (torch.nn.Conv3d, torch.nn.ConvTranspose3d, torch.nn.BatchNorm3d, torch.nn.ReLU))
    F.mse_loss(emu(x.to(dtype).float()), tgt).backward()
    with torch.autocast("cuda", dtype=dtype):
        out = m(x.to(cuda))
        loss = F.mse_loss(out.float(), tgt.to(cuda))
    scale = 1024.0 if dtype == torch.float16 else 1.0
    (loss * scale).backward()
    e = rel(out.float().cpu(), g["out"].float())
    print(f"\n[{dtype}] Unet3d depth-4 64^3 forward rel-L2 vs reference golden {e:.3e}; loss {loss.item():.5f} vs {g['loss']:.5f}")
    assert out.shape == g["out"].shape and e < tol
    assert abs(loss.item() - g["loss"]) < 3 * tol * abs(g["loss"])
    worst = gradient_ratios(m, ref, emu, scale)
    print("worst weight grads (ratio, ours, yardstick):", [(f"{r:.2f}", f"{a:.2e}", f"{b:.2e}", n) for r, a, b, n in worst[:4]])
    assert worst[0][0] < 1.5, worst[:3]
    for n, p in m.named_parameters():
        gn = g["grad_norms"][n]
        if gn < 1e-5:
            continue
        got = p.grad.float().norm().item() / scale
        assert abs(got - gn) <= (0.05 if dtype == torch.float16 else 0.25) * gn, (n, got, gn)


@pytest.mark.parametrize("stride,pad,k", [((1, 1, 1), (1, 1, 1), 3), ((2, 2, 2), (1, 1, 1), 3), ((1, 1, 1), (0, 0, 0), 1)])
def test_conv3d_cl_vs_torch(cuda, stride, pad, k):
    from viscy_b200 import functional as VF
    torch.manual_seed(0)
    conv = torch.nn.Conv3d(16, 24, k, stride=stride, padding=pad).to(cuda)
    x = torch.randn(2, 16, 8, 8, 16, device=cuda).half()
    xc = x.permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True)
    y = VF.conv3d_cl(xc, conv)
    xf = x.float().requires_grad_(True)
    ref = F.conv3d(xf, conv.weight.half().float(), conv.bias, stride=stride, padding=pad)
    assert rel(y.permute(0, 4, 1, 2, 3), ref) < 1e-3
    dy = torch.randn_like(ref).half()
    wf = conv.weight.detach().clone().requires_grad_(True)
    ref2 = F.conv3d(xf, wf, conv.bias, stride=stride, padding=pad)
    gx, gw = torch.autograd.grad(ref2, [xf, wf], dy.float())
    y.backward(dy.permute(0, 2, 3, 4, 1).contiguous())
    assert rel(xc.grad.permute(0, 4, 1, 2, 3), gx) < 2e-3
    assert rel(conv.weight.grad, gw) < 2e-3
    assert rel(conv.bias.grad, dy.float().sum((0, 2, 3, 4))) < 2e-3


def test_conv_transpose3d_cl_vs_torch(cuda):
    from viscy_b200 import functional as VF
    torch.manual_seed(0)
    ct = torch.nn.ConvTranspose3d(32, 16, kernel_size=3, stride=(2, 2, 2), padding=1, output_padding=1).to(cuda)
    x = torch.randn(1, 32, 4, 6, 8, device=cuda).half()
    xc = x.permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True)
    y = VF.conv_transpose3d_cl(xc, ct)
    xf = x.float().requires_grad_(True)
    wf = ct.weight.detach().clone().requires_grad_(True)
    bf = ct.bias.detach().clone().requires_grad_(True)
    ref = F.conv_transpose3d(xf, wf, bf, stride=2, padding=1, output_padding=1)
    assert y.shape == (1, 8, 12, 16, 16)
    assert rel(y.permute(0, 4, 1, 2, 3), ref) < 2e-3
    dy = torch.randn_like(ref).half()
    ref.backward(dy.float())
    y.backward(dy.permute(0, 2, 3, 4, 1).contiguous())
    assert rel(xc.grad.permute(0, 4, 1, 2, 3), xf.grad) < 2e-3
    assert rel(ct.weight.grad, wf.grad) < 2e-3 and rel(ct.bias.grad, bf.grad) < 2e-3


def test_batchnorm_relu_cat_cl(cuda):
    from viscy_b200 import functional as VF
    torch.manual_seed(0)
    bn = torch.nn.BatchNorm3d(24).to(cuda)
    with torch.no_grad():
        bn.weight.normal_(1, 0.2)
        bn.bias.normal_(0, 0.2)
    x = torch.randn(2, 24, 4, 8, 8, device=cuda).half()
    xc = x.permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True)
    y = VF.batchnorm_act_cl(xc, bn, relu=True)
    xf = x.float().requires_grad_(True)
    ref = F.relu(F.batch_norm(xf, None, None, bn.weight, bn.bias, True, 0.1, bn.eps))
    assert rel(y.permute(0, 4, 1, 2, 3), ref) < 2e-3
    dy = torch.randn_like(ref).half()
    gx, gw, gb = torch.autograd.grad(ref, [xf, bn.weight, bn.bias], dy.float())
    y.backward(dy.permute(0, 2, 3, 4, 1).contiguous())
    assert rel(xc.grad.permute(0, 4, 1, 2, 3), gx) < 3e-3
    assert rel(bn.weight.grad, gw) < 3e-3 and rel(bn.bias.grad, gb) < 3e-3
    a = torch.randn(3, 5, 16, device=cuda).half().requires_grad_(True)
    b = torch.randn(3, 5, 8, device=cuda).half().requires_grad_(True)
    c = VF.cat_cl(a, b)
    assert torch.equal(c, torch.cat([a, b], -1))
    c.backward(torch.ones_like(c))
    assert a.grad.shape == a.shape and b.grad.shape == b.shape


def test_boundary_convs_3_to_32_and_32_to_3(cuda):
    """UNet3DBase inconv / outconv (3 <-> 32 channels, padded to 8): line-kernel forward / data gradient, no patch matrix."""
    from viscy_b200 import functional as VF
    torch.manual_seed(2)
    for cin, cout in ((3, 32), (32, 3)):
        conv = torch.nn.Conv3d(cin, cout, 3, padding=1).to(cuda)
        x = torch.randn(1, cin, 8, 16, 32, device=cuda).half()
        xc = VF.to_channels_last_3d(x, torch.float16).requires_grad_(True)
        y = VF.conv3d_cl(xc, conv)
        xf = x.float().requires_grad_(True)
        wf = conv.weight.detach().half().float().requires_grad_(True)
        ref = F.conv3d(xf, wf, conv.bias, padding=1)
        assert rel(y.permute(0, 4, 1, 2, 3)[:, :cout], ref) < 2e-3
        dy = torch.randn_like(ref).half()
        gx, gw = torch.autograd.grad(ref, [xf, wf], dy.float())
        dyc = VF.to_channels_last_3d(dy, torch.float16)
        y.backward(dyc)
        assert rel(xc.grad.permute(0, 4, 1, 2, 3)[:, :cin], gx) < 2e-3
        assert rel(conv.weight.grad, gw) < 2e-3


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,R,C,G,act,cond", [(2, 300, 32, 8, "silu", True), (1, 4096, 64, 8, "silu", False),
                                               (3, 77, 16, 4, "relu", True), (2, 128, 12, 4, "silu", True)])
def test_groupnorm_scale_shift_act_vs_torch(cuda, N, R, C, G, act, cond, dtype):
    """GroupNorm -> * (scale + 1) + shift -> SiLU | ReLU on channels-last rows vs plain torch ops (fp32) on the same
    16-bit input: forward, input gradient and the gradients of gamma, beta, scale, shift."""
    from viscy_b200 import functional as VF
    torch.manual_seed(N + R + C)
    gn = torch.nn.GroupNorm(G, C).to(cuda)
    with torch.no_grad():
        gn.weight.normal_(1.0, 0.3)
        gn.bias.normal_(0.0, 0.3)
    Cp = -(-C // 8) * 8
    x = torch.zeros(N, 1, R, 1, Cp, device=cuda, dtype=dtype)  # [N, D, H, W, Cp], padded channels stay zero
    x[..., :C] = (torch.randn(N, 1, R, 1, C, device=cuda) * 1.5 + 0.3).to(dtype)
    scale = (torch.randn(N, C, device=cuda) * 0.3).requires_grad_(True) if cond else None
    shift = (torch.randn(N, C, device=cuda) * 0.3).requires_grad_(True) if cond else None
    dy = torch.zeros_like(x)
    dy[..., :C] = torch.randn(N, 1, R, 1, C, device=cuda).to(dtype)
    xm = x.clone().requires_grad_(True)
    y = VF.groupnorm_act_cl(xm, gn, act, scale, shift)
    y.backward(dy)
    got = dict(dx=xm.grad[..., :C].float(), dg=gn.weight.grad.clone(), db=gn.bias.grad.clone(),
               ds=None if scale is None else scale.grad.clone(), dt=None if shift is None else shift.grad.clone())
    gn.zero_grad()
    xr = x[..., :C].float().permute(0, 4, 1, 2, 3).clone().requires_grad_(True)  # NCDHW
    sr = None if scale is None else scale.detach().clone().requires_grad_(True)
    tr = None if shift is None else shift.detach().clone().requires_grad_(True)
    h = gn(xr)
    if cond:
        h = h * (sr[:, :, None, None, None] + 1) + tr[:, :, None, None, None]
    ref = F.silu(h) if act == "silu" else F.relu(h)
    ref.backward(dy[..., :C].float().permute(0, 4, 1, 2, 3))
    tol = 8e-3 if dtype == torch.bfloat16 else 1.5e-3
    assert rel(y[..., :C].float().permute(0, 4, 1, 2, 3), ref) < tol
    assert (y[..., C:] == 0).all()
    assert rel(got["dx"].permute(0, 4, 1, 2, 3), xr.grad) < tol
    assert rel(got["dg"], gn.weight.grad) < 2e-3 and rel(got["db"], gn.bias.grad) < 2e-3
    if cond:
        assert rel(got["ds"], sr.grad) < 2e-3 and rel(got["dt"], tr.grad) < 2e-3


@pytest.mark.parametrize("name", ["unet3d_base_gn", "unet3d_base_gn_odd"])
@pytest.mark.parametrize("dtype,tol", [(torch.float16, 4e-3), (torch.bfloat16, 3e-2)])
def test_unet3d_base_groupnorm_silu_timestep_golden(cuda, name, dtype, tol):
    """UNet3DBase with the reference's class defaults (GroupNorm + SiLU), residual blocks, timestep scale / shift and a
    conditioning input through the sm_100a kernels vs vectors from the reference's own code."""
    from viscy_b200 import UNet3DBase
    from viscy_b200.unet3d import ConvBottleneck3D
    g = torch.load(GOLD / f"{name}.pt", weights_only=False)
    cfg = g["cfg"]
    bott = ConvBottleneck3D(cfg["dims"][-1], time_emb_dim=g["time_embed_dim"], residual=True, groups=cfg["groups"])
    m = UNet3DBase(bottleneck=bott, time_embed_dim=g["time_embed_dim"], cond_channels=g["cond_channels"], **cfg)
    m.load_state_dict(g["state_dict"])
    m = m.to(cuda)
    dev = lambda v: None if v is None else v.to(cuda)  # noqa: E731
    with torch.autocast("cuda", dtype=dtype):
        out = m(dev(g["x"]), dev(g["cond"]), dev(g["t"]))
        loss = F.mse_loss(out.float(), g["target"].to(cuda))
    scale = 1024.0 if dtype == torch.float16 else 1.0
    (loss * scale).backward()
    e = rel(out.float().cpu(), g["out"])
    print(f"\n[{name} {dtype}] forward rel-L2 vs reference golden {e:.3e}")
    assert out.shape == g["out"].shape and e < tol
    worst = max((rel(p.grad.cpu() / scale, g["grads"][n]), n) for n, p in m.named_parameters() if g["grads"][n].norm() > 1e-6)
    print("worst gradient:", worst)
    assert worst[0] < 12 * tol


@pytest.mark.parametrize("make,xshape", [
    (lambda: __import__("viscy_b200").Unet3d(1, 1, depth=2, mult_chan=4), (1, 1, 16, 32, 32)),
    (lambda: __import__("viscy_b200").Unet3d(2, 3, depth=2, mult_chan=12), (1, 2, 16, 32, 32)),
    (lambda: __import__("viscy_b200").Unet25d(num_filters=(4, 8, 12), num_blocks=2, dropout=0.0, task="reg", residual=True),
     (2, 1, 5, 32, 32)),
])
def test_channel_counts_not_multiple_of_8(cuda, make, xshape):
    """Skip concatenation / residual growth with channel counts that are not multiples of 8: rows are padded per tensor, the
    concatenated rows must be [a | b | pad] (what the next conv's weight rows assume), not [a | pad | b | pad]."""
    torch.manual_seed(3)
    ref = make()
    m = make()
    m.load_state_dict(ref.state_dict())
    m = m.to(cuda)
    x = torch.randn(xshape)
    out_ref = ref(x)
    tgt = torch.randn_like(out_ref)
    F.mse_loss(out_ref, tgt).backward()
    with torch.autocast("cuda", dtype=torch.float16):
        out = m(x.to(cuda))
        loss = F.mse_loss(out.float(), tgt.to(cuda))
    (loss * 256.0).backward()
    e = rel(out.float().cpu(), out_ref)
    print(f"\nforward rel-L2 vs the fp32 mirror {e:.3e}")
    assert out.shape == out_ref.shape and e < 1e-2
    refg = dict(ref.named_parameters())
    for n, p in m.named_parameters():
        gr = refg[n].grad
        if gr is None:
            assert p.grad is None, n
        elif gr.dim() > 1 and gr.norm() > 1e-6:
            assert rel(p.grad.cpu() / 256.0, gr) < 0.25, n  # a mis-laid concat shows up as O(1) errors
