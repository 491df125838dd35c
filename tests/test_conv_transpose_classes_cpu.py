"""Host logic of the transposed-conv lowering (functional._parity_classes): ConvTranspose3d forward = one stride-1
sub-convolution per output parity class over the taps {k : (k - r - p) mod s == 0}.  Here every class is evaluated with
plain torch ops exactly as the kernel launches are parameterised (kernel = class taps, pad_lo, extra output extent,
output written to the class's sub-lattice) and the assembly is compared with torch's conv_transpose3d."""
import itertools

import pytest
import torch
import torch.nn.functional as F

from viscy_b200.functional import _parity_classes


def _emulate(x, w, bias, ks, stride, padding, out_sp):
    """x [N,Ci,d,h,w], w [Ci,Co,kd,kh,kw] (ConvTranspose3d layout) -> [N,Co,*out_sp] via per-class stride-1 convs."""
    N, Ci = x.shape[:2]
    Co = w.shape[1]
    per_dim = [_parity_classes(k, s, p, n, o) for k, s, p, n, o in zip(ks, stride, padding, x.shape[2:], out_sp)]
    assert all(c is not None for c in per_dim)
    out = torch.zeros((N, Co, *out_sp))
    covered = torch.zeros(out_sp, dtype=torch.int32)
    for (rz, tz, pz, ez), (ry, ty, py, ey), (rx, tx, px, ex) in itertools.product(*per_dim):
        # class filter [Co, Ci, Kz, Ky, Kx]: tap t of the class is tap tz[t] of the full filter (the kernel's tapmap)
        wsub = w[:, :, tz][:, :, :, ty][:, :, :, :, tx].permute(1, 0, 2, 3, 4)
        # correlation with low-side padding pad_lo and enough high-side zeros for `extra` more outputs
        xp = F.pad(x, (px, px + ex, py, py + ey, pz, pz + ez))
        y = F.conv3d(xp, wsub, bias)
        J = [len(range(r, o, s)) for r, o, s in zip((rz, ry, rx), out_sp, stride)]
        y = y[:, :, :J[0], :J[1], :J[2]]
        assert list(y.shape[2:]) == J  # the class extent the launch computes is exactly the sub-lattice size
        out[:, :, rz::stride[0], ry::stride[1], rx::stride[2]] = y
        covered[rz::stride[0], ry::stride[1], rx::stride[2]] += 1
    assert int(covered.min()) == 1 and int(covered.max()) == 1  # every output voxel belongs to exactly one class
    return out


@pytest.mark.parametrize("ks,stride,padding,opad", [
    ((3, 3, 3), (2, 2, 2), (1, 1, 1), (1, 1, 1)),   # UNet3DBase upsample (unet3d_base.py:118-126)
    ((1, 3, 3), (1, 2, 2), (0, 1, 1), (0, 1, 1)),   # downsample_z=False variant
    ((3, 3, 3), (2, 2, 2), (1, 1, 1), (0, 0, 0)),   # odd output extent
    ((4, 4, 4), (2, 2, 2), (0, 0, 0), (0, 0, 0)),   # even filter: two taps per class everywhere
    ((5, 3, 3), (1, 1, 1), (2, 1, 1), (0, 0, 0)),   # stride 1: a single class = the flipped conv
    ((3, 3, 3), (3, 3, 3), (0, 0, 0), (0, 0, 0)),   # kernel == stride: one tap per class
])
def test_parity_class_decomposition_equals_conv_transpose3d(ks, stride, padding, opad):
    torch.manual_seed(0)
    x = torch.randn(2, 3, 4, 5, 6)
    ct = torch.nn.ConvTranspose3d(3, 5, ks, stride=stride, padding=padding, output_padding=opad)
    ref = ct(x)
    got = _emulate(x, ct.weight.detach(), ct.bias.detach(), ks, stride, padding, tuple(ref.shape[2:]))
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)


def test_strided_conv_data_gradient_is_a_transposed_conv():
    """Conv3d(k3, s2, p1) dgrad = transposed conv of dout with the conv weight read as [in=Co, out=Ci] (Conv3dFn.backward)."""
    torch.manual_seed(1)
    conv = torch.nn.Conv3d(4, 6, 3, stride=2, padding=1)
    x = torch.randn(1, 4, 8, 8, 8, requires_grad=True)
    y = conv(x)
    dy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, dy)
    got = _emulate(dy, conv.weight.detach(), None, (3, 3, 3), (2, 2, 2), (1, 1, 1), (8, 8, 8))
    torch.testing.assert_close(got, gx, rtol=1e-5, atol=1e-5)


def test_unsupported_decompositions_are_declined():
    # stride 3 with a 2-tap filter: the class r = 2 (p = 0) has no tap -> the caller keeps the scatter lowering
    assert _parity_classes(2, 3, 0, 4, 11) is None
    # k4 s2 p1: a class would need FEWER outputs than its symmetric-padding extent -> declined as well
    assert _parity_classes(4, 2, 1, 6, 12) is None
