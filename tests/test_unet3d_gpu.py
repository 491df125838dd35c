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
