"""Fully convolutional masked autoencoder (VM/unet/fcmae.py:26-539) with the reference constructor / forward /
state_dict surface: `FullyConvolutionalMAE` (the model behind the published VSCyto2D / VSCyto3D / VSNeuromast
checkpoints, registry key "fcmae", CY/engine.py:36-43) and its masked ConvNeXt-V2 encoder.

CPU tensors run the reference math in plain torch ops.  CUDA tensors run channels-last 16-bit through the sm_100a kernels:
dense blocks (mask_ratio == 0: fine-tuning / inference) are the fused ConvNeXt-V2 block kernels with nn.LayerNorm's eps;
with mask_ratio > 0 the blocks gather the unmasked rows behind the depthwise conv and run LayerNorm / fc1 / GELU / GRN / fc2
on those rows alone (GRN statistics over the kept pixels, exactly like masked_patchify), then scatter onto the masked
shortcut - the reference's sparse path, not a dense emulation.
"""

from __future__ import annotations

import math
from typing import Sequence

import torch
import torch.nn.functional as TF
from torch import BoolTensor, Size, Tensor, nn

from . import functional as F
from .components import (ConvNeXtMlp, DropPath, LayerNorm2d, PixelToVoxelHead, PixelToVoxelShuffleHead, UNeXt2Decoder)
from .unext2 import resolve_compute_dtype


def _init_weights(module: nn.Module) -> None:
    """fcmae.py:26-37"""
    if isinstance(module, nn.Conv2d):
        nn.init.trunc_normal_(module.weight, std=0.02)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, nn.Linear):
        nn.init.trunc_normal_(module.weight, std=0.02)
        nn.init.zeros_(module.bias)
    elif isinstance(module, nn.LayerNorm):
        nn.init.ones_(module.weight)
        nn.init.zeros_(module.bias)


def generate_mask(target: Size, stride: int, mask_ratio: float, device: str) -> BoolTensor:
    """Random binary mask at low resolution, B1HW, True = masked (fcmae.py:40-65)."""
    m_height = target[-2] // stride
    m_width = target[-1] // stride
    mask_numel = m_height * m_width
    masked_elements = int(mask_numel * mask_ratio)
    mask = torch.rand(target[0], mask_numel, device=device).argsort(1) < masked_elements
    return mask.reshape(target[0], 1, m_height, m_width)


def upsample_mask(mask: BoolTensor, target: Size) -> BoolTensor:
    """Nearest-neighbour upsampling of a low-resolution mask to `target` (BCHW) (fcmae.py:68-88)."""
    if target[-2:] != mask.shape[-2:]:
        if not all(i % j == 0 for i, j in zip(target, mask.shape)):
            raise ValueError(f"feature map shape {target} must be divisible by mask shape {mask.shape}.")
        mask = mask.repeat_interleave(target[-2] // mask.shape[-2], dim=-2).repeat_interleave(
            target[-1] // mask.shape[-1], dim=-1
        )
    return mask


def masked_patchify(features: Tensor, unmasked: BoolTensor | None = None) -> Tensor:
    """BCHW -> channels-last rows of the unmasked pixels, BLC (fcmae.py:91-113)."""
    if unmasked is None:
        return features.flatten(2).permute(0, 2, 1)
    b, c = features.shape[:2]
    features = features.permute(0, 2, 3, 1)
    return features[unmasked[:, 0]].reshape(b, -1, c)


def masked_unpatchify(features: Tensor, out_shape: Size, unmasked: BoolTensor | None = None) -> Tensor:
    """BLC rows back to BCHW, zeros at the masked pixels (fcmae.py:116-141)."""
    if unmasked is None:
        return features.permute(0, 2, 1).reshape(out_shape)
    b, c, w, h = out_shape
    out = torch.zeros((b, w, h, c), device=features.device, dtype=features.dtype)
    out[unmasked[:, 0]] = features.reshape(-1, c)
    return out.permute(0, 3, 1, 2)


class _Downsample(nn.Module):
    """timm.models.convnext.Downsample: average pool (stride > 1) + 1x1 conv (channel change)."""

    def __init__(self, in_chs: int, out_chs: int, stride: int = 1):
        super().__init__()
        self.pool = nn.AvgPool2d(2, stride, ceil_mode=True, count_include_pad=False) if stride > 1 else nn.Identity()
        self.conv = nn.Conv2d(in_chs, out_chs, 1, stride=1) if in_chs != out_chs else nn.Identity()

    def forward(self, x: Tensor) -> Tensor:
        return self.conv(self.pool(x))


class MaskedConvNeXtV2Block(nn.Module):
    """Masked ConvNeXt V2 block (fcmae.py:144-227): dw7x7 -> nn.LayerNorm (eps 1e-5) -> Linear -> GELU -> GRN -> Linear."""

    def __init__(self, in_channels: int, out_channels: int | None = None, kernel_size: int = 7, stride: int = 1,
                 mlp_ratio: int = 4, drop_path: float = 0.0) -> None:
        super().__init__()
        out_channels = out_channels or in_channels
        self.dwconv = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                                padding=(kernel_size - 1) // 2, groups=out_channels)
        self.layernorm = nn.LayerNorm(out_channels)
        self.mlp = ConvNeXtMlp(out_channels, mlp_ratio * out_channels, use_grn=True, use_conv=False)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        if in_channels != out_channels or stride > 1:
            self.shortcut = _Downsample(in_channels, out_channels, stride=stride)
        else:
            self.shortcut = nn.Identity()

    def forward(self, x: Tensor, unmasked: BoolTensor | None = None) -> Tensor:
        shortcut = self.shortcut(x)
        if unmasked is not None:
            x *= unmasked
        x = self.dwconv(x)
        if unmasked is not None:
            x *= unmasked
        out_shape = x.shape
        x = masked_patchify(x, unmasked=unmasked)
        x = self.layernorm(x)
        x = self.mlp(x.unsqueeze(1)).squeeze(1)
        x = masked_unpatchify(x, out_shape=out_shape, unmasked=unmasked)
        return self.drop_path(x) + shortcut

    def forward_cl(self, x: Tensor, mi: F.MaskIndex | None = None) -> Tensor:
        if not isinstance(self.shortcut, nn.Identity) or self.dwconv.kernel_size != (7, 7):
            raise NotImplementedError("sm_100a MaskedConvNeXtV2Block: 7x7 depthwise conv with an identity shortcut only")
        keep = None
        if isinstance(self.drop_path, DropPath):
            keep = F.drop_path_scale(x, self.drop_path.drop_prob, self.training)
        return F.fcmae_block(x, self, mi, keep)


class MaskedConvNeXtV2Stage(nn.Module):
    """Masked ConvNeXt V2 stage (fcmae.py:230-308): [LayerNorm2d + Conv2d(k = stride)] + blocks."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 7, stride: int = 2, num_blocks: int = 2,
                 drop_path_rates: Sequence[float] | None = None) -> None:
        super().__init__()
        if drop_path_rates is None:
            drop_path_rates = [0.0] * num_blocks
        elif len(drop_path_rates) != num_blocks:
            raise ValueError(
                "length of drop_path_rates must be equal to "
                f"the number of blocks {num_blocks}, got {len(drop_path_rates)}."
            )
        if in_channels != out_channels or stride > 1:
            k = stride if stride > 1 else 1
            self.downsample = nn.Sequential(
                LayerNorm2d(in_channels), nn.Conv2d(in_channels, out_channels, kernel_size=k, stride=stride, padding=0)
            )
            in_channels = out_channels
        else:
            self.downsample = nn.Identity()
        self.blocks = nn.ModuleList()
        for i in range(num_blocks):
            self.blocks.append(MaskedConvNeXtV2Block(in_channels, out_channels, kernel_size=kernel_size, stride=1,
                                                     drop_path=drop_path_rates[i]))
            in_channels = out_channels

    def forward(self, x: Tensor, unmasked: BoolTensor | None = None) -> Tensor:
        x = self.downsample(x)
        if unmasked is not None:
            unmasked = upsample_mask(unmasked, x.shape)
        for block in self.blocks:
            x = block(x, unmasked)
        return x

    def forward_cl(self, x: Tensor, unmasked: BoolTensor | None = None, kept_cells: int = 0) -> Tensor:
        if not isinstance(self.downsample, nn.Identity):
            x = F.ln_conv(x, self.downsample[0], self.downsample[1])
        mi = None
        if unmasked is not None:
            B, H, W, _ = x.shape
            up = upsample_mask(unmasked, (B, 1, H, W))
            cell = (H // unmasked.shape[-2]) * (W // unmasked.shape[-1])
            mi = F.MaskIndex(up[:, 0], kept_cells * cell)
        for block in self.blocks:
            x = block.forward_cl(x, mi)
        return x


class MaskedAdaptiveProjection(nn.Module):
    """Masked patchifying layer projecting 2-D or 3-D input into 2-D feature maps (fcmae.py:311-385)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size_2d: tuple[int, int] | int = 4,
                 kernel_depth: int = 5, in_stack_depth: int = 5) -> None:
        super().__init__()
        ratio = in_stack_depth // kernel_depth
        if isinstance(kernel_size_2d, int):
            kernel_size_2d = [kernel_size_2d] * 2
        kernel_size_3d = [kernel_depth, *kernel_size_2d]
        self.conv3d = nn.Conv3d(in_channels, out_channels // ratio, kernel_size=kernel_size_3d, stride=kernel_size_3d)
        self.conv2d = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size_2d, stride=kernel_size_2d)
        self.norm = nn.LayerNorm(out_channels)

    def forward(self, x: Tensor, unmasked: BoolTensor = None) -> Tensor:
        # no need to mask before convolutions since patches do not spill over
        if x.shape[2] > 1:
            x = self.conv3d(x)
            b, c, d, h, w = x.shape
            x = x.reshape(b, c * d, h, w)
        else:
            x = self.conv2d(x.squeeze(2))
        out_shape = x.shape
        if unmasked is not None:
            unmasked = upsample_mask(unmasked, x.shape)
        x = masked_patchify(x, unmasked=unmasked)
        x = self.norm(x)
        return masked_unpatchify(x, out_shape=out_shape, unmasked=unmasked)

    def forward_cl(self, x: Tensor, dtype: torch.dtype) -> Tensor:
        """NCDHW fp32/16-bit in -> NHWC rows (the encoder calls the stem without a mask, fcmae.py:441)."""
        if x.shape[2] > 1:
            f = F.stem(x, self.conv3d, dtype)
        else:
            c = self.conv2d
            f = F.StemFn.apply(x, c.weight.unsqueeze(2), c.bias, (1, *c.stride), dtype)
        return F.layernorm(f, self.norm.weight, self.norm.bias, self.norm.eps)


class MaskedMultiscaleEncoder(nn.Module):
    """Multiscale encoder with optional sparse masking for MAE pretraining (fcmae.py:388-446)."""

    def __init__(self, in_channels: int, stage_blocks: Sequence[int] = (3, 3, 9, 3),
                 dims: Sequence[int] = (96, 192, 384, 768), drop_path_rate: float = 0.0,
                 stem_kernel_size: Sequence[int] = (5, 4, 4), in_stack_depth: int = 5) -> None:
        super().__init__()
        self.stem = MaskedAdaptiveProjection(in_channels, dims[0], kernel_size_2d=stem_kernel_size[1:],
                                             kernel_depth=stem_kernel_size[0], in_stack_depth=in_stack_depth)
        self.stages = nn.ModuleList()
        chs = [dims[0], *dims]
        for i, num_blocks in enumerate(stage_blocks):
            stride = 1 if i == 0 else 2
            self.stages.append(MaskedConvNeXtV2Stage(chs[i], chs[i + 1], kernel_size=7, stride=stride,
                                                     num_blocks=num_blocks, drop_path_rates=[drop_path_rate] * num_blocks))
        self.total_stride = stem_kernel_size[1] * 2 ** (len(self.stages) - 1)
        self.apply(_init_weights)

    def _masks(self, x: Tensor, mask_ratio: float):
        if mask_ratio > 0.0:
            mask = generate_mask(x.shape, self.total_stride, mask_ratio, device=x.device)
            b, c, d, h, w = x.shape
            return upsample_mask(mask, (b, 1, h, w)), ~mask
        return None, None

    def forward(self, x: Tensor, mask_ratio: float = 0.0) -> tuple[list[Tensor], BoolTensor | None]:
        mask, unmasked = self._masks(x, mask_ratio)
        x = self.stem(x)
        features = []
        for stage in self.stages:
            x = stage(x, unmasked=unmasked)
            features.append(x)
        return features, mask

    def forward_cl(self, x: Tensor, dtype: torch.dtype, mask_ratio: float = 0.0, unmasked: BoolTensor | None = None):
        """-> (NHWC feature maps, full-resolution mask or None).  `unmasked` (B1hw, low resolution) overrides the draw."""
        mask = None
        if unmasked is None:
            mask, unmasked = self._masks(x, mask_ratio)
        else:
            mask = upsample_mask(~unmasked, (x.shape[0], 1, *x.shape[-2:]))
        kept = 0
        if unmasked is not None:
            cells = unmasked.shape[-2] * unmasked.shape[-1]
            kept = cells - int(cells * mask_ratio) if mask_ratio > 0.0 else int(unmasked[0].sum().item())
        f = self.stem.forward_cl(x, dtype)
        features = []
        for stage in self.stages:
            f = stage.forward_cl(f, unmasked, kept)
            features.append(f)
        return features, mask


class FullyConvolutionalMAE(nn.Module):
    """Fully Convolutional Masked Autoencoder (fcmae.py:449-539)."""

    def __init__(self, in_channels: int, out_channels: int, encoder_blocks: Sequence[int] = (3, 3, 9, 3),
                 dims: Sequence[int] = (96, 192, 384, 768), encoder_drop_path_rate: float = 0.0,
                 stem_kernel_size: Sequence[int] = (5, 4, 4), in_stack_depth: int = 5, decoder_conv_blocks: int = 1,
                 pretraining: bool = True, head_conv: bool = False, head_conv_expansion_ratio: int = 4,
                 head_conv_pool: bool = True) -> None:
        super().__init__()
        self.encoder = MaskedMultiscaleEncoder(in_channels=in_channels, stage_blocks=encoder_blocks, dims=dims,
                                               drop_path_rate=encoder_drop_path_rate, stem_kernel_size=stem_kernel_size,
                                               in_stack_depth=in_stack_depth)
        decoder_channels = list(dims)
        decoder_channels.reverse()
        if head_conv:
            decoder_channels[-1] = (in_stack_depth + 2) * in_channels * 2**2 * head_conv_expansion_ratio
        else:
            decoder_channels[-1] = out_channels * in_stack_depth * stem_kernel_size[-1] ** 2
        self.decoder = UNeXt2Decoder(decoder_channels, norm_name="instance", mode="pixelshuffle",
                                     conv_blocks=decoder_conv_blocks,
                                     strides=[2] * (len(dims) - 1) + [stem_kernel_size[-1]], upsample_pre_conv=None)
        if head_conv:
            self.head = PixelToVoxelHead(in_channels=decoder_channels[-1], out_channels=out_channels,
                                         out_stack_depth=in_stack_depth, expansion_ratio=head_conv_expansion_ratio,
                                         pool=head_conv_pool)
        else:
            self.head = PixelToVoxelShuffleHead(in_channels=decoder_channels[-1], out_channels=out_channels,
                                                out_stack_depth=in_stack_depth, xy_scaling=stem_kernel_size[-1], pool=True)
        self.out_stack_depth = in_stack_depth
        self.num_blocks = len(dims) * int(math.log2(stem_kernel_size[-1]))
        self.pretraining = pretraining
        self.compute_dtype: torch.dtype | None = None
        self._packs = None

    def _weight_packs(self, dt: torch.dtype) -> None:
        from . import ops
        from .components import ConvNeXtBlock
        if self._packs is None or self._packs.dtype != dt or self._packs.stale():
            lin, dws = [], []
            for m in self.modules():
                if isinstance(m, (ConvNeXtBlock, MaskedConvNeXtV2Block)):
                    lin += [m.mlp.fc1.weight, m.mlp.fc2.weight]
                    dws.append(m.conv_dw.weight if isinstance(m, ConvNeXtBlock) else m.dwconv.weight)
            self._packs = ops.WeightPacks(lin, dws, dt)
        self._packs.refresh()
        ops.ACTIVE_PACKS = self._packs

    def _forward_sm100(self, x: Tensor, mask_ratio: float, unmasked: BoolTensor | None = None):
        from . import ops
        dt = resolve_compute_dtype(x, self.compute_dtype)
        self._weight_packs(dt)
        ops.STEP.begin(x.device, torch.is_grad_enabled())
        with torch.autocast("cuda", enabled=False):
            feats, mask = self.encoder.forward_cl(x, dt, mask_ratio, unmasked)
            feats.reverse()
            return self.head.forward_cl(self.decoder.forward_cl(feats)), mask

    def forward(self, x: Tensor, mask_ratio: float = 0.0) -> Tensor:
        if x.is_cuda:
            x, mask = self._forward_sm100(x, mask_ratio)
        else:
            x, mask = self.encoder(x, mask_ratio=mask_ratio)
            x.reverse()
            x = self.decoder(x)
            x = self.head(x)
        if self.pretraining:
            return x, mask
        return x
