"""Host-side mirror of the reference's building blocks (same module / parameter names, so `state_dict()` keys
and tensor layouts equal the reference's and published checkpoints load unchanged).

Every module has two backends:
  * `forward(x)`      -- the reference math in plain torch ops on NCHW tensors (CPU tensors, `example_input_array`,
                         BASELINE config #1);
  * `forward_cl(x)`   -- channels-last 16-bit tensors through the hand-written sm_100a kernels (mandatory for CUDA
                         tensors; unsupported configurations raise, there is no cuDNN / CPU fallback).

Reference: timm ConvNeXt pieces as composed by VM/unet/unext2.py:40-49 and VM/components/blocks.py:54-74 (SURVEY.md
Appendix B.1), VM/components/stems.py, VM/components/blocks.py:77-243, VM/components/heads.py:594-641.
"""

from __future__ import annotations

from typing import Callable, Literal, Sequence

import torch
import torch.nn.functional as TF
from torch import Tensor, nn

from . import functional as F

CONVNEXT_CFGS = {
    # name: depths, dims, use_grn, ls_init_value, conv_mlp   (timm 1.0.x convnext.py model defs)
    "convnext_tiny": ((3, 3, 9, 3), (96, 192, 384, 768), False, 1e-6, False),
    "convnextv2_atto": ((2, 2, 6, 2), (40, 80, 160, 320), True, None, True),
    "convnextv2_femto": ((2, 2, 6, 2), (48, 96, 192, 384), True, None, True),
    "convnextv2_pico": ((2, 2, 6, 2), (64, 128, 256, 512), True, None, True),
    "convnextv2_nano": ((2, 2, 8, 2), (80, 160, 320, 640), True, None, True),
    "convnextv2_tiny": ((3, 3, 9, 3), (96, 192, 384, 768), True, None, False),
    "convnextv2_base": ((3, 3, 27, 3), (128, 256, 512, 1024), True, None, False),
}


def init_convnext_weights(module: nn.Module) -> None:
    """timm.models.convnext._init_weights: trunc_normal(std=.02) weights, zero biases."""
    if isinstance(module, (nn.Conv2d, nn.Linear)):
        nn.init.trunc_normal_(module.weight, std=0.02)
        if module.bias is not None:
            nn.init.zeros_(module.bias)


def icnr_init(conv: nn.Module, upsample_factor: int, upsample_dims: int, init: Callable = nn.init.kaiming_normal_):
    """ICNR initialisation of a sub-pixel convolution (VM/components/blocks.py:14-51)."""
    out_channels, in_channels, *dims = conv.weight.shape
    scale = upsample_factor**upsample_dims
    oc2 = int(out_channels / scale)
    kernel = init(torch.zeros([oc2, in_channels] + dims)).transpose(0, 1)
    kernel = kernel.reshape(oc2, in_channels, -1).repeat(1, 1, scale)
    kernel = kernel.reshape([in_channels, out_channels] + dims).transpose(0, 1)
    conv.weight.data.copy_(kernel)


class LayerNorm2d(nn.LayerNorm):
    """LayerNorm over the channel dim of NCHW (timm.layers.LayerNorm2d, eps 1e-6)."""

    def __init__(self, num_channels: int, eps: float = 1e-6):
        super().__init__(num_channels, eps=eps)

    def forward(self, x: Tensor) -> Tensor:
        x = TF.layer_norm(x.permute(0, 2, 3, 1), self.normalized_shape, self.weight, self.bias, self.eps)
        return x.permute(0, 3, 1, 2)

    def forward_cl(self, x: Tensor) -> Tensor:
        return F.layernorm(x, self.weight, self.bias, self.eps)


class GlobalResponseNorm(nn.Module):
    """timm GlobalResponseNorm (zero-init weight / bias of shape [dim])."""

    def __init__(self, dim: int, eps: float = 1e-6, channels_last: bool = True):
        super().__init__()
        self.eps = eps
        self.channels_last = channels_last
        self.weight = nn.Parameter(torch.zeros(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def forward(self, x: Tensor) -> Tensor:
        sp, ch, shape = ((1, 2), -1, (1, 1, 1, -1)) if self.channels_last else ((2, 3), 1, (1, -1, 1, 1))
        x_g = x.norm(p=2, dim=sp, keepdim=True)
        x_n = x_g / (x_g.mean(dim=ch, keepdim=True) + self.eps)
        return x + torch.addcmul(self.bias.view(shape), self.weight.view(shape), x * x_n)


class ConvNeXtMlp(nn.Module):
    """timm Mlp / GlobalResponseNormMlp (fc1 -> GELU -> [GRN] -> fc2); 1x1 Conv2d when `use_conv`."""

    def __init__(self, dim: int, hidden: int, use_grn: bool, use_conv: bool):
        super().__init__()
        lin = (lambda i, o: nn.Conv2d(i, o, kernel_size=1)) if use_conv else nn.Linear
        self.fc1 = lin(dim, hidden)
        self.act = nn.GELU()
        if use_grn:
            self.grn = GlobalResponseNorm(hidden, channels_last=not use_conv)
        self.fc2 = lin(hidden, dim)

    def forward(self, x: Tensor) -> Tensor:
        x = self.act(self.fc1(x))
        if hasattr(self, "grn"):
            x = self.grn(x)
        return self.fc2(x)


class DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x: Tensor) -> Tensor:
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0:
            mask.div_(keep)
        return x * mask


class ConvNeXtBlock(nn.Module):
    """timm ConvNeXtBlock: dw7x7 -> LN -> fc1 -> GELU -> [GRN] -> fc2 -> [*gamma] -> drop_path -> + x."""

    def __init__(self, dim: int, use_grn: bool, conv_mlp: bool, ls_init_value: float | None, drop_path: float = 0.0):
        super().__init__()
        self.use_conv_mlp = conv_mlp
        self.conv_dw = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim, bias=True)
        self.norm = LayerNorm2d(dim) if conv_mlp else nn.LayerNorm(dim, eps=1e-6)
        self.mlp = ConvNeXtMlp(dim, 4 * dim, use_grn, conv_mlp)
        self.gamma = nn.Parameter(ls_init_value * torch.ones(dim)) if ls_init_value is not None else None
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x: Tensor) -> Tensor:
        shortcut = x
        x = self.conv_dw(x)
        if self.use_conv_mlp:
            x = self.mlp(self.norm(x))
        else:
            x = self.mlp(self.norm(x.permute(0, 2, 3, 1))).permute(0, 3, 1, 2)
        if self.gamma is not None:
            x = x * self.gamma.reshape(1, -1, 1, 1)
        return self.drop_path(x) + shortcut

    def forward_cl(self, x: Tensor, keep: Tensor | None = None) -> Tensor:
        """keep: per-sample stochastic-depth scale (fp32 [B]); drawn here like timm's DropPath when not given."""
        if keep is None and isinstance(self.drop_path, DropPath):
            keep = F.drop_path_scale(x, self.drop_path.drop_prob, self.training)
        return F.convnext_block(x, self, keep)


class ConvNeXtStage(nn.Module):
    """timm ConvNeXtStage: [LayerNorm2d + Conv2d(k = 2 if stride > 1 else 1)] + `depth` blocks."""

    def __init__(self, in_chs: int, out_chs: int, stride: int, depth: int, use_grn: bool, conv_mlp: bool,
                 ls_init_value: float | None, drop_path_rates: Sequence[float] | None = None):
        super().__init__()
        if in_chs != out_chs or stride > 1:
            k = 2 if stride > 1 else 1
            self.downsample = nn.Sequential(LayerNorm2d(in_chs), nn.Conv2d(in_chs, out_chs, kernel_size=k, stride=stride))
        else:
            self.downsample = nn.Identity()
        dpr = drop_path_rates or [0.0] * depth
        self.blocks = nn.Sequential(
            *[ConvNeXtBlock(out_chs, use_grn, conv_mlp, ls_init_value, dpr[i]) for i in range(depth)]
        )

    def forward(self, x: Tensor) -> Tensor:
        return self.blocks(self.downsample(x))

    def forward_cl(self, x: Tensor) -> Tensor:
        if not isinstance(self.downsample, nn.Identity):
            x = F.ln_conv(x, self.downsample[0], self.downsample[1])
        for blk in self.blocks:
            x = blk.forward_cl(x)
        return x


def _make_stages(backbone: str, drop_path_rate: float):
    if backbone not in CONVNEXT_CFGS:
        raise ValueError(f"backbone {backbone!r} is not available in viscy_b200 (have {sorted(CONVNEXT_CFGS)})")
    depths, dims, use_grn, ls, conv_mlp = CONVNEXT_CFGS[backbone]
    dpr = [x.tolist() for x in torch.linspace(0, drop_path_rate, sum(depths)).split(depths)]
    stages, prev = [], dims[0]
    for i in range(4):
        stages.append(ConvNeXtStage(prev, dims[i], 2 if i > 0 else 1, depths[i], use_grn, conv_mlp, ls, dpr[i]))
        prev = dims[i]
    return stages, list(dims)


class _FeatureInfo:
    def __init__(self, chans):
        self._chans = list(chans)

    def channels(self):
        return list(self._chans)


class ConvNeXtFeatures(nn.Module):
    """`timm.create_model(backbone, features_only=True)` (FeatureListNet: stem_0, stem_1, stages_0..3)."""

    def __init__(self, backbone: str, drop_path_rate: float = 0.0, in_chans: int = 3):
        super().__init__()
        stages, dims = _make_stages(backbone, drop_path_rate)
        self.stem_0 = nn.Conv2d(in_chans, dims[0], kernel_size=4, stride=4)
        self.stem_1 = LayerNorm2d(dims[0])
        for i, s in enumerate(stages):
            setattr(self, f"stages_{i}", s)
        self.feature_info = _FeatureInfo(dims)
        self.apply(init_convnext_weights)

    def stages(self):
        return [getattr(self, f"stages_{i}") for i in range(4)]

    def forward(self, x: Tensor) -> list[Tensor]:
        x = self.stem_1(self.stem_0(x))
        out = []
        for s in self.stages():
            x = s(x)
            out.append(x)
        return out

    def forward_cl(self, x: Tensor) -> list[Tensor]:
        if not isinstance(self.stem_0, nn.Identity):
            raise NotImplementedError("sm_100a path expects the 3-D stem in place of timm's stem_0")
        x = self.stem_1.forward_cl(x)
        out = []
        for s in self.stages():
            x = s.forward_cl(x)
            out.append(x)
        return out


# ------------------------------------------------------------------------------------------------ stems
class UNeXt2Stem(nn.Module):
    """Stem for UNeXt2 and ContrastiveEncoder networks (VM/components/stems.py:8-50)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: tuple[int, int, int], in_stack_depth: int):
        super().__init__()
        if in_stack_depth < kernel_size[0]:
            raise ValueError(f"in_stack_depth ({in_stack_depth}) must be >= kernel_size[0] ({kernel_size[0]})")
        ratio = in_stack_depth // kernel_size[0]
        if out_channels % ratio != 0:
            raise ValueError(
                f"out_channels ({out_channels}) must be divisible by in_stack_depth // kernel_size[0] ({ratio})"
            )
        self.conv = nn.Conv3d(in_channels, out_channels // ratio, kernel_size=kernel_size, stride=kernel_size)

    def forward(self, x: Tensor) -> Tensor:
        x = self.conv(x)
        b, c, d, h, w = x.shape
        return x.reshape(b, c * d, h, w)

    def forward_cl(self, x: Tensor, dtype: torch.dtype) -> Tensor:
        return F.stem(x, self.conv, dtype)


class StemDepthtoChannels(nn.Module):
    """Stem with 3D convolution that maps depth to channels (VM/components/stems.py:53-134)."""

    def __init__(self, in_channels: int, in_stack_depth: int, in_channels_encoder: int,
                 stem_kernel_size: tuple[int, int, int] = (5, 4, 4), stem_stride: tuple[int, int, int] = (5, 4, 4)):
        super().__init__()
        c = self.compute_stem_channels(in_stack_depth, stem_kernel_size, stem_stride[0], in_channels_encoder)
        self.conv = nn.Conv3d(in_channels, c, kernel_size=stem_kernel_size, stride=stem_stride)

    def compute_stem_channels(self, in_stack_depth, stem_kernel_size, stem_stride_depth, in_channels_encoder) -> int:
        stem3d_out_depth = (in_stack_depth - stem_kernel_size[0]) // stem_stride_depth + 1
        stem3d_out_channels = in_channels_encoder // stem3d_out_depth
        channel_mismatch = in_channels_encoder - stem3d_out_depth * stem3d_out_channels
        if channel_mismatch != 0:
            raise ValueError(
                f"Stem needs to output {channel_mismatch} more channels "
                "to match the encoder. Adjust the in_stack_depth."
            )
        return stem3d_out_channels

    def forward(self, x: Tensor) -> Tensor:
        x = self.conv(x)
        b, c, d, h, w = x.shape
        return x.reshape(b, c * d, h, w)

    def forward_cl(self, x: Tensor, dtype: torch.dtype) -> Tensor:
        return F.stem(x, self.conv, dtype)


# ------------------------------------------------------------------------------------------------ decoder
class PixelShuffleUpSample(nn.Module):
    """monai UpSample(mode='pixelshuffle') = SubpixelUpsample: [conv_block] -> pixel shuffle -> [pad + avg-pool]."""

    def __init__(self, in_channels: int, out_channels: int, scale_factor: int, pre_conv, apply_pad_pool: bool):
        super().__init__()
        r = scale_factor
        self.scale_factor = r
        if pre_conv == "default":
            self.conv_block = nn.Conv2d(in_channels, out_channels * r * r, kernel_size=3, stride=1, padding=1)
            icnr_init(self.conv_block, r, 2)
        elif pre_conv is None:
            self.conv_block = nn.Identity()
        else:
            self.conv_block = pre_conv
        self.pad_pool = nn.Identity()
        if apply_pad_pool:
            self.pad_pool = nn.Sequential(nn.ConstantPad2d((r - 1, 0) * 2, 0.0), nn.AvgPool2d(kernel_size=r, stride=1))

    def forward(self, x: Tensor) -> Tensor:
        return self.pad_pool(TF.pixel_shuffle(self.conv_block(x), self.scale_factor))


def _get_convnext_stage(in_channels: int, out_channels: int, depth: int, upsample_factor: int | None = None):
    """VM/components/blocks.py:54-74: ConvNeXt-V2 stage with 1x1-conv MLP, stride 1, no layer scale."""
    stage = ConvNeXtStage(in_channels, out_channels, stride=1, depth=depth, use_grn=True, conv_mlp=True,
                          ls_init_value=None)
    stage.apply(init_convnext_weights)
    if upsample_factor:
        icnr_init(stage.blocks[-1].mlp.fc2, upsample_factor, upsample_dims=2)
    return stage


class UNeXt2UpStage(nn.Module):
    """Single upsampling stage for the UNeXt2 decoder (VM/components/blocks.py:77-172)."""

    def __init__(self, in_channels: int, skip_channels: int, out_channels: int, scale_factor: int,
                 mode: Literal["deconv", "pixelshuffle"], conv_blocks: int, norm_name: str,
                 upsample_pre_conv: Literal["default"] | Callable | None):
        super().__init__()
        if mode == "deconv":
            raise NotImplementedError(
                "decoder_mode='deconv' is known-broken in the reference (strict xfail, "
                "packages/viscy-models/tests/test_unet/test_unext2.py:45-57) and is not provided"
            )
        if mode != "pixelshuffle":
            raise ValueError(f"unknown decoder mode {mode!r}")
        mid_channels = in_channels // scale_factor**2
        self.upsample = PixelShuffleUpSample(in_channels, mid_channels, scale_factor, upsample_pre_conv, False)
        self.conv = _get_convnext_stage(mid_channels + skip_channels, out_channels, conv_blocks,
                                        upsample_factor=None if upsample_pre_conv else scale_factor)

    def forward(self, inp: Tensor, skip: Tensor) -> Tensor:
        inp = self.upsample(inp)
        inp = torch.cat([inp, skip], dim=1)
        return self.conv(inp)

    def forward_cl(self, inp: Tensor, skip: Tensor) -> Tensor:
        if not isinstance(self.upsample.conv_block, nn.Identity) or self.upsample.scale_factor != 2:
            raise NotImplementedError("sm_100a decoder stage: pixel-shuffle x2 without pre-conv only")
        return self.conv.forward_cl(F.pixshuf_cat(inp, skip))


class UNeXt2Decoder(nn.Module):
    """Multi-stage UNeXt2 decoder (VM/components/blocks.py:175-243)."""

    def __init__(self, num_channels: list[int], norm_name: str, mode: Literal["deconv", "pixelshuffle"],
                 conv_blocks: int, strides: list[int], upsample_pre_conv: Literal["default"] | Callable | None):
        super().__init__()
        self.decoder_stages = nn.ModuleList([])
        for i in range(len(num_channels) - 1):
            self.decoder_stages.append(
                UNeXt2UpStage(num_channels[i], num_channels[i] // 2, num_channels[i + 1], strides[i], mode,
                              conv_blocks, norm_name, upsample_pre_conv)
            )

    def forward(self, features: Sequence[Tensor]) -> Tensor:
        feat = features[0]
        skips = list(features[1:]) + [None]
        for skip, stage in zip(skips, self.decoder_stages):
            feat = stage(feat, skip)
        return feat

    def forward_cl(self, features: Sequence[Tensor]) -> Tensor:
        feat = features[0]
        skips = list(features[1:]) + [None]
        for skip, stage in zip(skips, self.decoder_stages):
            feat = stage.forward_cl(feat, skip)
        return feat


# ------------------------------------------------------------------------------------------------ head
class _ADN(nn.Sequential):
    def __init__(self, channels: int):
        super().__init__()
        self.add_module("N", nn.InstanceNorm3d(channels))
        self.add_module("A", nn.PReLU())


class _Convolution(nn.Sequential):
    """monai Convolution(spatial_dims=3): Conv3d + ADN(InstanceNorm3d, PReLU)  (keys conv.*, adn.A.weight)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size, padding):
        super().__init__()
        self.add_module("conv", nn.Conv3d(in_channels, out_channels, kernel_size, 1, padding, bias=True))
        self.add_module("adn", _ADN(out_channels))


class PixelToVoxelHead(nn.Module):
    """Pixel-shuffle head that upsamples 2D features to 3D voxel output (VM/components/heads.py:594-641)."""

    def __init__(self, in_channels: int, out_channels: int, out_stack_depth: int, expansion_ratio: int, pool: bool):
        super().__init__()
        first_scale = 2
        self.upsample = PixelShuffleUpSample(in_channels, in_channels // first_scale**2, first_scale, None, pool)
        mid_channels = out_channels * expansion_ratio * 2**2
        self.conv = nn.Sequential(
            _Convolution(in_channels // first_scale**2 // (out_stack_depth + 2), mid_channels, 3, (0, 1, 1)),
            nn.Conv3d(mid_channels, out_channels * 2**2, 1),
        )
        icnr_init(self.conv[-1], 2, upsample_dims=2)
        self.out = nn.PixelShuffle(2)
        self.out_stack_depth = out_stack_depth
        self.pool = pool

    def forward(self, x: Tensor) -> Tensor:
        x = self.upsample(x)
        d = self.out_stack_depth + 2
        b, c, h, w = x.shape
        x = x.reshape((b, c // d, d, h, w))
        x = self.conv(x)
        x = x.transpose(1, 2)
        x = self.out(x)
        return x.transpose(1, 2)

    def forward_cl(self, x: Tensor) -> Tensor:
        c0 = self.conv[0]
        return F.pixel_to_voxel_head(x, c0.conv, c0.adn.A, self.conv[1], self.out_stack_depth, self.pool)


class PixelToVoxelShuffleHead(nn.Module):
    """Pixel-shuffle head that reshapes 2D features into a 3D volume (VM/components/heads.py:657-695); no parameters."""

    def __init__(self, in_channels: int, out_channels: int, out_stack_depth: int = 5, xy_scaling: int = 4,
                 pool: bool = False):
        super().__init__()
        self.out_channels = out_channels
        self.out_stack_depth = out_stack_depth
        self.upsample = PixelShuffleUpSample(in_channels, out_stack_depth * out_channels, xy_scaling, None, pool)
        self.pool = pool

    def forward(self, x: Tensor) -> Tensor:
        x = self.upsample(x)
        b, _, h, w = x.shape
        return x.reshape(b, self.out_channels, self.out_stack_depth, h, w)

    def forward_cl(self, x: Tensor) -> Tensor:
        y = F.shuffle_pool(x, self.upsample.scale_factor, self.pool)  # [B, out * depth, H, W]
        b, _, h, w = y.shape
        return y.view(b, self.out_channels, self.out_stack_depth, h, w)
