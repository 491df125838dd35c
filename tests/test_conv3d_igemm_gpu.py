"""Implicit-GEMM Conv3d (TMA boxes at tap-shifted coordinates -> tcgen05) vs torch's fp32 conv3d on the same 16-bit
inputs: forward (+bias, +ReLU, +residual), data gradient, weight gradient.  Tolerances: 16-bit outputs 2e-3 rel-L2
(fp16) / 8e-3 (bf16); fp32 weight gradients 2e-3."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


# (N, D, H, W, Cin, Cout, kernel, padding)
CASES = [
    (1, 8, 16, 16, 64, 64, (3, 3, 3), (1, 1, 1)),     # box (16, 8, 1, 1)
    (2, 4, 8, 8, 64, 32, (3, 3, 3), (1, 1, 1)),       # box (8, 8, 2, 1), Cout < tile
    (1, 2, 4, 128, 128, 96, (3, 3, 3), (1, 1, 1)),    # box (128, 1, 1, 1), two channel chunks per tap
    (1, 4, 4, 256, 64, 64, (3, 3, 3), (1, 1, 1)),     # OW > 128: two tiles per row
    (1, 8, 8, 8, 32, 32, (3, 3, 3), (1, 1, 1)),       # 32-channel operands: SWIZZLE_64B rows
    (1, 8, 8, 16, 96, 64, (3, 3, 3), (1, 1, 1)),      # 96 = 3 x 32: SWIZZLE_64B, three chunks per tap
    (4, 4, 4, 4, 64, 320, (3, 3, 3), (1, 1, 1)),      # box (4, 4, 4, 2): tile spans samples; Cout > 256
    (1, 2, 2, 8, 64, 64, (3, 3, 3), (1, 1, 1)),       # 32 voxels: tile larger than the tensor
    (1, 5, 16, 16, 64, 64, (1, 3, 3), (0, 1, 1)),     # Unet25d decoder filter
    (1, 5, 8, 16, 64, 64, (5, 1, 1), (0, 0, 0)),      # Unet25d skip / bottom filter: valid in Z -> OD = 1
    (1, 10, 16, 16, 64, 32, (3, 3, 3), (0, 1, 1)),    # valid in Z (head style): OD = 8
    # extents in whole 16 x 8 patches + kh == 3: the patch form (one y-haloed box serves the three kh taps)
    (1, 4, 32, 32, 32, 32, (3, 3, 3), (1, 1, 1)),     # 32-channel operand, 4 patches per plane
    (1, 2, 16, 32, 64, 128, (3, 3, 3), (1, 1, 1)),    # 128-wide output tile
    (2, 3, 8, 16, 128, 96, (3, 3, 3), (1, 1, 1)),     # two channel chunks, two samples
    (1, 3, 24, 32, 64, 64, (3, 3, 1), (1, 1, 0)),     # kw == 1
    # >= 2 patches per CTA and a filter <= 112 KB: the resident-filter variant (weights fetched once per CTA)
    (1, 8, 64, 128, 32, 32, (3, 3, 3), (1, 1, 1)),    # 32-row weight sub-tiles, MMA N = 32; dgrad the same form
    (1, 5, 64, 128, 64, 32, (3, 3, 3), (1, 1, 1)),    # 64 -> 32; its dgrad is the 32 -> 64 form (64-row sub-tiles)
    (2, 3, 64, 64, 32, 16, (1, 3, 3), (0, 1, 1)),     # 9 taps, two samples
    # >= 148 patches, 64-channel K blocks, filter too large to stay resident: the patch form on CTA pairs (cta_group::2)
    (1, 16, 32, 64, 64, 64, (3, 3, 3), (1, 1, 1)),    # 256 patches, 64-wide tile (each CTA stages 32 filter rows per tap)
    (2, 8, 64, 32, 128, 128, (3, 3, 3), (1, 1, 1)),   # 256 patches, 128-wide tile, two channel chunks, two samples
    (1, 20, 24, 32, 64, 128, (3, 3, 3), (1, 1, 1)),   # 180 patches: ragged last wave of pairs
]


@pytest.mark.parametrize("case", CASES, ids=[str(c) for c in CASES])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_igemm_forward_dgrad_wgrad(cuda, case, dtype):
    from viscy_b200 import _lib as L, ops
    N, D, H, W, Ci, Co, ks, pad = case
    tol = 2e-3 if dtype == torch.float16 else 8e-3
    g = torch.Generator(device=cuda).manual_seed(1)
    x = torch.randn(N, D, H, W, Ci, device=cuda, generator=g).to(dtype)
    w = (torch.randn(Co, Ci, *ks, device=cuda, generator=g) / (Ci * ks[0] * ks[1] * ks[2]) ** 0.5)
    b = torch.randn(Co, device=cuda, generator=g)
    assert ops.conv3d_igemm_supported((N, D, H, W, Ci), Co, ks, pad)
    w16 = w.permute(0, 2, 3, 4, 1).reshape(Co, -1).contiguous().to(dtype)
    xf = x.float().permute(0, 4, 1, 2, 3)
    wf = w16.float().view(Co, *ks, Ci).permute(0, 4, 1, 2, 3)
    ref = F.conv3d(xf, wf, b, padding=pad)
    y = ops.conv3d_igemm(x, w16, b, ks, pad)
    assert y.shape == (N, *ref.shape[2:], Co)
    assert rel(y.float().permute(0, 4, 1, 2, 3), ref) < tol
    # fused ReLU + residual
    res = torch.randn_like(y)
    y2 = ops.conv3d_igemm(x, w16, b, ks, pad, act=L.ACT_RELU, residual=res)
    assert rel(y2.float().permute(0, 4, 1, 2, 3), F.relu(ref) + res.float().permute(0, 4, 1, 2, 3)) < tol
    # data gradient = conv of dout with the flipped / transposed filter, padding k-1-p
    dy = torch.randn(y.shape, device=cuda, generator=g).to(dtype)
    dyf = dy.float().permute(0, 4, 1, 2, 3)
    gx = torch.nn.grad.conv3d_input(xf.shape, wf, dyf, padding=pad)
    gw = torch.nn.grad.conv3d_weight(xf, wf.shape, dyf, padding=pad)
    bpad = tuple(k - 1 - p for k, p in zip(ks, pad))
    if Co % 32 == 0 and ops.conv3d_igemm_supported(tuple(dy.shape), Ci, ks, bpad):
        wflip = w16.view(Co, *ks, Ci).flip(1, 2, 3).permute(4, 1, 2, 3, 0).reshape(Ci, -1).contiguous()
        dx = ops.conv3d_igemm(dy, wflip, None, ks, bpad)
        assert dx.shape == x.shape
        assert rel(dx.float().permute(0, 4, 1, 2, 3), gx) < tol
    assert ops.conv3d_igemm_supported((N, D, H, W, Ci), Co, ks, pad, wgrad=True)
    for splits in (0, 1, 3):
        dw = ops.conv3d_igemm_wgrad(x, dy, ks, pad, k_splits=splits)
        assert rel(dw.view(Co, *ks, Ci).permute(0, 4, 1, 2, 3), gw) < 2e-3, splits


def test_igemm_wgrad_narrow_channels(cuda):
    """weight gradient with Cin = 8 / 24 (below the forward form's 32-channel granule): zero-filled channel boxes"""
    from viscy_b200 import ops
    g = torch.Generator(device=cuda).manual_seed(2)
    for Ci, Co in ((8, 32), (24, 8), (160, 16)):
        x = torch.randn(1, 4, 8, 16, Ci, device=cuda, generator=g).half()
        dy = torch.randn(1, 4, 8, 16, Co, device=cuda, generator=g).half()
        gw = torch.nn.grad.conv3d_weight(x.float().permute(0, 4, 1, 2, 3), (Co, Ci, 3, 3, 3),
                                         dy.float().permute(0, 4, 1, 2, 3), padding=1)
        dw = ops.conv3d_igemm_wgrad(x, dy, (3, 3, 3), (1, 1, 1))
        assert rel(dw.view(Co, 3, 3, 3, Ci).permute(0, 4, 1, 2, 3), gw) < 2e-3, (Ci, Co)


def test_conv3d_cl_uses_igemm_and_matches_torch(cuda):
    """Conv3dFn picks the implicit-GEMM form for box-shaped geometries; same results as torch autograd."""
    from viscy_b200 import _lib, functional as VF
    torch.manual_seed(0)
    conv = torch.nn.Conv3d(64, 48, 3, padding=1).to(cuda)
    x = torch.randn(2, 64, 8, 8, 16, device=cuda).half()
    xc = x.permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True)
    n0 = _lib.launch_count()
    y = VF.conv3d_cl(xc, conv)
    assert _lib.launch_count() - n0 == 2  # weight cast + one implicit-GEMM launch: no im2col
    xf = x.float().requires_grad_(True)
    wf = conv.weight.detach().half().float().requires_grad_(True)
    ref = F.conv3d(xf, wf, conv.bias, padding=1)
    assert rel(y.permute(0, 4, 1, 2, 3), ref) < 2e-3
    dy = torch.randn_like(ref).half()
    gx, gw = torch.autograd.grad(ref, [xf, wf], dy.float())
    y.backward(dy.permute(0, 2, 3, 4, 1).contiguous())
    assert rel(xc.grad.permute(0, 4, 1, 2, 3), gx) < 2e-3
    assert rel(conv.weight.grad, gw) < 2e-3
    assert rel(conv.bias.grad, dy.float().sum((0, 2, 3, 4))) < 2e-3


def test_igemm_rejects_unboxable_geometry(cuda):
    from viscy_b200 import ops
    assert not ops.conv3d_igemm_supported((1, 7, 12, 20, 64), 64, (3, 3, 3), (1, 1, 1))
    x = torch.randn(1, 7, 12, 20, 64, device=cuda).half()
    w16 = torch.randn(64, 27 * 64, device=cuda).half()
    with pytest.raises(NotImplementedError):
        ops.conv3d_igemm(x, w16, None, (3, 3, 3), (1, 1, 1))


# strided convs: TMA element strides deliver every s-th voxel of a box spanning s*b voxels
STRIDED = [
    (1, 16, 16, 16, 64, 64, (3, 3, 3), (1, 1, 1), (2, 2, 2)),
    (1, 8, 32, 32, 32, 64, (3, 3, 3), (1, 1, 1), (2, 2, 2)),     # 32-channel operand (SWIZZLE_64B)
    (2, 4, 16, 16, 64, 32, (1, 3, 3), (0, 1, 1), (1, 2, 2)),     # UNet3DBase down_stride (1,2,2)
    (1, 4, 4, 256, 64, 64, (3, 3, 3), (1, 1, 1), (2, 2, 2)),     # 128 output voxels per row: the 256-voxel box limit
    (1, 8, 8, 8, 128, 256, (3, 3, 3), (1, 1, 1), (2, 2, 2)),     # 64 output voxels: tile larger than the tensor
]


@pytest.mark.parametrize("case", STRIDED, ids=[str(c) for c in STRIDED])
def test_igemm_strided_forward_and_wgrad(cuda, case):
    from viscy_b200 import ops
    N, D, H, W, Ci, Co, ks, pad, st = case
    g = torch.Generator(device=cuda).manual_seed(5)
    x = torch.randn(N, D, H, W, Ci, device=cuda, generator=g).half()
    w = torch.randn(Co, Ci, *ks, device=cuda, generator=g) / (Ci * ks[0] * ks[1] * ks[2]) ** 0.5
    b = torch.randn(Co, device=cuda, generator=g)
    assert ops.conv3d_igemm_supported((N, D, H, W, Ci), Co, ks, pad, stride=st)
    w16 = w.permute(0, 2, 3, 4, 1).reshape(Co, -1).contiguous().half()
    xf = x.float().permute(0, 4, 1, 2, 3)
    wf = w16.float().view(Co, *ks, Ci).permute(0, 4, 1, 2, 3)
    ref = F.conv3d(xf, wf, b, stride=st, padding=pad)
    y = ops.conv3d_igemm(x, w16, b, ks, pad, stride=st)
    assert y.shape == (N, *ref.shape[2:], Co)
    assert rel(y.float().permute(0, 4, 1, 2, 3), ref) < 2e-3
    dy = torch.randn(y.shape, device=cuda, generator=g).half()
    gw = torch.nn.grad.conv3d_weight(xf, wf.shape, dy.float().permute(0, 4, 1, 2, 3), stride=st, padding=pad)
    assert ops.conv3d_igemm_supported((N, D, H, W, Ci), Co, ks, pad, wgrad=True, stride=st)
    dw = ops.conv3d_igemm_wgrad(x, dy, ks, pad, stride=st)
    assert rel(dw.view(Co, *ks, Ci).permute(0, 4, 1, 2, 3), gw) < 2e-3


def test_strided_conv_and_transposed_conv_use_igemm(cuda):
    """Conv3d(stride 2) forward / wgrad and ConvTranspose3d backward run without a patch matrix on box geometries."""
    from viscy_b200 import functional as VF
    torch.manual_seed(1)
    conv = torch.nn.Conv3d(64, 96, 3, stride=2, padding=1).to(cuda)
    x = torch.randn(1, 64, 16, 16, 16, device=cuda).half()
    xc = x.permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True)
    y = VF.conv3d_cl(xc, conv)
    xf = x.float().requires_grad_(True)
    wf = conv.weight.detach().half().float().requires_grad_(True)
    ref = F.conv3d(xf, wf, conv.bias, stride=2, padding=1)
    assert rel(y.permute(0, 4, 1, 2, 3), ref) < 2e-3
    dy = torch.randn_like(ref).half()
    gx, gw = torch.autograd.grad(ref, [xf, wf], dy.float())
    y.backward(dy.permute(0, 2, 3, 4, 1).contiguous())
    assert rel(xc.grad.permute(0, 4, 1, 2, 3), gx) < 2e-3 and rel(conv.weight.grad, gw) < 2e-3

    ct = torch.nn.ConvTranspose3d(64, 32, kernel_size=3, stride=(2, 2, 2), padding=1, output_padding=1).to(cuda)
    x = torch.randn(1, 64, 4, 8, 8, device=cuda).half()
    xc = x.permute(0, 2, 3, 4, 1).contiguous().requires_grad_(True)
    y = VF.conv_transpose3d_cl(xc, ct)
    xf = x.float().requires_grad_(True)
    wf = ct.weight.detach().clone().requires_grad_(True)
    bf = ct.bias.detach().clone().requires_grad_(True)
    ref = F.conv_transpose3d(xf, wf, bf, stride=2, padding=1, output_padding=1)
    assert y.shape == (1, 8, 16, 16, 32) and rel(y.permute(0, 4, 1, 2, 3), ref) < 2e-3
    dy = torch.randn_like(ref).half()
    ref.backward(dy.float())
    y.backward(dy.permute(0, 2, 3, 4, 1).contiguous())
    assert rel(xc.grad.permute(0, 4, 1, 2, 3), xf.grad) < 2e-3
    assert rel(ct.weight.grad, wf.grad) < 2e-3 and rel(ct.bias.grad, bf.grad) < 2e-3


# (N, D, H, W, Cin, Cout, kernel, padding): few-channel patch-form weight gradient (kh taps = views of one haloed box)
KH3 = [
    (1, 4, 16, 16, 32, 32, (3, 3, 3), (1, 1, 1)),
    (2, 3, 8, 24, 64, 32, (3, 3, 3), (1, 1, 1)),
    (1, 2, 16, 8, 8, 16, (3, 3, 3), (1, 1, 1)),       # 8-channel input (padded stem): zero-filled channel box
    (1, 4, 8, 8, 96, 160, (3, 3, 3), (1, 1, 1)),      # two channel tiles, two M tiles
    (1, 5, 16, 16, 48, 16, (1, 3, 3), (0, 1, 1)),     # Unet25d decoder filter
    (1, 6, 10, 16, 32, 64, (3, 3, 3), (0, 0, 1)),     # valid in Z and Y: output rows 8
]


@pytest.mark.parametrize("case", KH3, ids=[str(c) for c in KH3])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_wgrad_kh3_vs_torch(cuda, case, dtype):
    from viscy_b200 import ops
    N, D, H, W, Ci, Co, ks, pad = case
    g = torch.Generator(device=cuda).manual_seed(9)
    x = torch.randn(N, D, H, W, Ci, device=cuda, generator=g).to(dtype)
    assert ops.conv3d_wgrad_kh3_supported((N, D, H, W, Ci), Co, ks, pad)
    od, oh, ow = (D + 2 * pad[0] - ks[0] + 1, H + 2 * pad[1] - ks[1] + 1, W + 2 * pad[2] - ks[2] + 1)
    dy = torch.randn(N, od, oh, ow, Co, device=cuda, generator=g).to(dtype)
    gw = torch.nn.grad.conv3d_weight(x.float().permute(0, 4, 1, 2, 3), (Co, Ci, *ks), dy.float().permute(0, 4, 1, 2, 3),
                                     padding=pad)
    for splits in (0, 1, 5):
        dw = ops.conv3d_wgrad_kh3(x, dy, ks, pad, k_splits=splits)
        assert rel(dw.view(Co, *ks, Ci).permute(0, 4, 1, 2, 3), gw) < 2e-3, splits
    assert not ops.conv3d_wgrad_kh3_supported((N, D, H, W, Ci), Co, ks, pad, stride=(1, 2, 2))
    assert not ops.conv3d_wgrad_kh3_supported((N, D, H + 3, W, Ci), Co, ks, pad)
