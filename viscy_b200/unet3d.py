"""UNet3DBase / Unet3d (F-Net) and their conv blocks (VM/unet/unet3d_base.py:19-198, VM/unet/unet3d.py:37-86,
VM/unet/blocks.py:62-292) with the reference constructor / forward / state_dict surface.

CPU tensors: plain torch ops.  CUDA tensors: channels-last 16-bit activations through the sm_100a kernels
(`functional.conv3d` lowering on the tcgen05 GEMM; fused BatchNorm+ReLU kernels for the `Unet3d` preset, BASELINE
config 5; GroupNorm + timestep scale/shift + SiLU kernels for the class defaults, csrc/groupnorm_sm100.cu).
"""

from __future__ import annotations

import math
from typing import Literal

import torch
from torch import Tensor, nn

from . import functional as F
from .unext2 import resolve_compute_dtype


def _make_norm(norm: str, channels: int, groups: int) -> nn.Module:
    if norm == "group":
        return nn.GroupNorm(groups, channels)
    if norm == "batch":
        return nn.BatchNorm3d(channels)
    raise ValueError(f"Unknown norm type: {norm!r}")


def _make_activation(activation: str) -> nn.Module:
    if activation == "silu":
        return nn.SiLU()
    if activation == "relu":
        return nn.ReLU(inplace=True)
    raise ValueError(f"Unknown activation type: {activation!r}")


class Block(nn.Module):
    """Conv3d(k3, p1) -> norm -> optional (scale, shift) -> activation."""

    def __init__(self, dim: int, dim_out: int, norm: Literal["group", "batch"] = "group",
                 activation: Literal["silu", "relu"] = "silu", groups: int = 8) -> None:
        super().__init__()
        self.proj = nn.Conv3d(dim, dim_out, 3, padding=1)
        self.norm = _make_norm(norm, dim_out, groups)
        self.act = _make_activation(activation)

    def forward(self, x: Tensor, scale_shift: tuple[Tensor, Tensor] | None = None) -> Tensor:
        x = self.norm(self.proj(x))
        if scale_shift is not None:
            scale, shift = scale_shift
            x = x * (scale + 1) + shift
        return self.act(x)

    def forward_cl(self, x: Tensor, scale_shift: tuple[Tensor, Tensor] | None = None) -> Tensor:
        h = F.conv3d_cl(x, self.proj)
        act = "silu" if isinstance(self.act, nn.SiLU) else "relu"
        if isinstance(self.norm, nn.GroupNorm):
            scale, shift = scale_shift if scale_shift is not None else (None, None)
            return F.groupnorm_act_cl(h, self.norm, act, scale, shift)
        if scale_shift is not None or act != "relu":
            raise NotImplementedError("sm_100a Block: BatchNorm3d comes with ReLU and without timestep conditioning")
        return F.batchnorm_act_cl(h, self.norm, relu=True)


class ResnetBlock(nn.Module):
    """Two `Block`s; optional residual (1x1x1 conv when channels change); optional timestep scale/shift."""

    def __init__(self, dim: int, dim_out: int, *, time_emb_dim: int | None = None, residual: bool = True,
                 norm: Literal["group", "batch"] = "group", activation: Literal["silu", "relu"] = "silu",
                 groups: int = 8) -> None:
        super().__init__()
        self.mlp = nn.Sequential(nn.SiLU(), nn.Linear(time_emb_dim, dim_out * 2)) if time_emb_dim is not None else None
        self.block1 = Block(dim, dim_out, norm=norm, activation=activation, groups=groups)
        self.block2 = Block(dim_out, dim_out, norm=norm, activation=activation, groups=groups)
        self.res_conv = (nn.Conv3d(dim, dim_out, 1) if dim != dim_out else nn.Identity()) if residual else None

    def forward(self, x: Tensor, time_emb: Tensor | None = None) -> Tensor:
        scale_shift = None
        if self.mlp is not None and time_emb is not None:
            scale_shift = self.mlp(time_emb)[:, :, None, None, None].chunk(2, dim=1)
        h = self.block2(self.block1(x, scale_shift=scale_shift))
        return h if self.res_conv is None else h + self.res_conv(x)

    def forward_cl(self, x: Tensor, time_emb: Tensor | None = None) -> Tensor:
        scale_shift = None
        if self.mlp is not None and time_emb is not None:
            # [N, 2 C] timestep projection (a few KFLOP): plain fp32 torch ops; only its (scale, shift) enter the kernels
            scale_shift = self.mlp(time_emb.float()).chunk(2, dim=1)
        h = self.block2.forward_cl(self.block1.forward_cl(x, scale_shift))
        if self.res_conv is None:
            return h
        r = x if isinstance(self.res_conv, nn.Identity) else F.conv3d_cl(x, self.res_conv)
        return F.add_cl(h, r)


class TimestepEmbedder(nn.Module):
    """Sinusoidal timestep embedding followed by a 2-layer MLP."""

    def __init__(self, hidden_size: int, frequency_embedding_size: int = 256) -> None:
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden_size, bias=True), nn.SiLU(),
                                 nn.Linear(hidden_size, hidden_size, bias=True))
        half = frequency_embedding_size // 2
        freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
        self.register_buffer("freqs", freqs)

    def forward(self, t: Tensor) -> Tensor:
        args = t[:, None].float() * self.freqs[None]
        return self.mlp(torch.cat([torch.cos(args), torch.sin(args)], dim=-1))


class ConvBottleneck3D(nn.Module):
    """Bottleneck = one ResnetBlock, `forward(x, time_embeds=None)`."""

    def __init__(self, channels: int, *, time_emb_dim: int | None = None, residual: bool = True,
                 norm: Literal["group", "batch"] = "group", activation: Literal["silu", "relu"] = "silu",
                 groups: int = 8) -> None:
        super().__init__()
        self.block = ResnetBlock(channels, channels, time_emb_dim=time_emb_dim, residual=residual, norm=norm,
                                 activation=activation, groups=groups)

    def forward(self, x: Tensor, time_embeds: Tensor | None = None) -> Tensor:
        return self.block(x, time_embeds)

    def forward_cl(self, x: Tensor, time_embeds: Tensor | None = None) -> Tensor:
        return self.block.forward_cl(x, time_embeds)


class UNet3DBase(nn.Module):
    """Parametrized 3D U-Net with injected bottleneck: encoder -> bottleneck -> decoder with concat skips."""

    def __init__(self, in_channels: int, out_channels: int, dims: list[int], num_res_block: list[int],
                 bottleneck: nn.Module, downsample_z: bool = False, residual: bool = True,
                 norm: Literal["group", "batch"] = "group", activation: Literal["silu", "relu"] = "silu",
                 groups: int = 8, time_embed_dim: int | None = None, cond_channels: int | None = None) -> None:
        super().__init__()
        if len(dims) != len(num_res_block) + 1:
            raise ValueError(f"len(dims)={len(dims)} must equal len(num_res_block)+1={len(num_res_block) + 1}")
        self._num_res_block = list(num_res_block)
        self._divisor = 2 ** len(num_res_block)
        self.downsamples_z: bool = downsample_z
        kw = dict(norm=norm, activation=activation, groups=groups)

        self._time_embedder: TimestepEmbedder | None = None
        if time_embed_dim is not None:
            self._time_embedder = TimestepEmbedder(hidden_size=time_embed_dim)
        self.inconv = nn.Conv3d(in_channels, dims[0], kernel_size=3, stride=1, padding=1)
        self._cond_inconv: nn.Conv3d | None = None
        if cond_channels is not None:
            self._cond_inconv = nn.Conv3d(cond_channels, dims[0], kernel_size=3, stride=1, padding=1)
        if downsample_z:
            down_stride = (2, 2, 2)
            up_kw = dict(kernel_size=3, stride=(2, 2, 2), padding=1, output_padding=1)
        else:
            down_stride = (1, 2, 2)
            up_kw = dict(kernel_size=(1, 3, 3), stride=(1, 2, 2), padding=(0, 1, 1), output_padding=(0, 1, 1))

        levels = len(num_res_block)
        self._encoder_blocks = nn.ModuleList()
        self._downsamples = nn.ModuleList()
        for lv in range(levels):
            self._encoder_blocks.append(nn.ModuleList(
                [ResnetBlock(dims[lv], dims[lv], time_emb_dim=time_embed_dim, residual=residual, **kw)
                 for _ in range(num_res_block[lv])]))
            self._downsamples.append(nn.Conv3d(dims[lv], dims[lv + 1], kernel_size=3, stride=down_stride, padding=1))
        self.bottleneck = bottleneck
        self._upsamples = nn.ModuleList()
        self._decoder_blocks = nn.ModuleList()
        for lv in reversed(range(levels)):
            self._upsamples.append(nn.ConvTranspose3d(dims[lv + 1], dims[lv], **up_kw))
            self._decoder_blocks.append(nn.ModuleList(
                [ResnetBlock(dims[lv] * 2, dims[lv], time_emb_dim=time_embed_dim, residual=residual, **kw)
                 for _ in range(num_res_block[lv])]))
        self.outconv = nn.Conv3d(dims[0], out_channels, kernel_size=3, stride=1, padding=1)
        self.compute_dtype: torch.dtype | None = None

    @property
    def num_blocks(self) -> int:
        """Number of spatial downsampling stages."""
        return len(self._num_res_block)

    def _check_divisible(self, x: Tensor) -> None:
        for dim_name, size in zip(("D", "H", "W"), x.shape[2:]):
            if self.downsamples_z or dim_name != "D":
                if size % self._divisor != 0:
                    raise ValueError(
                        f"Spatial dim {dim_name}={size} must be divisible by "
                        f"{self._divisor} (2^{self.num_blocks} levels)."
                    )

    def forward(self, x: Tensor, cond: Tensor | None = None, t: Tensor | None = None) -> Tensor:
        self._check_divisible(x)
        if x.is_cuda:
            return self._forward_sm100(x, cond, t)
        time_embeds = self._time_embedder(t) if (self._time_embedder is not None and t is not None) else None
        h = self.inconv(x)
        if self._cond_inconv is not None and cond is not None:
            h = h + self._cond_inconv(cond)
        skips: list[Tensor] = []
        for blocks, down in zip(self._encoder_blocks, self._downsamples):
            for blk in blocks:
                h = blk(h, time_embeds)
                skips.append(h)
            h = down(h)
        h = self.bottleneck(h, time_embeds=time_embeds)
        for up, blocks in zip(self._upsamples, self._decoder_blocks):
            h = up(h)
            for blk in blocks:
                h = blk(torch.cat([h, skips.pop()], dim=1), time_embeds)
        return self.outconv(h)

    def _forward_sm100(self, x: Tensor, cond: Tensor | None, t: Tensor | None) -> Tensor:
        if not hasattr(self.bottleneck, "forward_cl"):
            raise NotImplementedError(f"sm_100a UNet3DBase: bottleneck {type(self.bottleneck).__name__}")
        dt = resolve_compute_dtype(x, self.compute_dtype)
        F.ops.ACTIVE_PACKS = None  # weight packs are scoped to the model that registered them
        F.ops.STEP.begin(x.device, torch.is_grad_enabled())  # one zero-filled allocation for the step's accumulators
        with torch.autocast("cuda", enabled=False):
            # timestep embedding: sinusoid + 2-layer MLP on [N] scalars (fp32 torch ops; negligible work)
            te = self._time_embedder(t) if (self._time_embedder is not None and t is not None) else None
            h = F.conv3d_cl(F.to_channels_last_3d(x, dt), self.inconv)
            if self._cond_inconv is not None and cond is not None:
                h = F.add_cl(h, F.conv3d_cl(F.to_channels_last_3d(cond, dt), self._cond_inconv))
            skips: list[Tensor] = []
            for blocks, down in zip(self._encoder_blocks, self._downsamples):
                for blk in blocks:
                    h = blk.forward_cl(h, te)
                    skips.append((h, down.in_channels))
                h = F.conv3d_cl(h, down)
            h = self.bottleneck.forward_cl(h, te)
            for up, blocks in zip(self._upsamples, self._decoder_blocks):
                h = F.conv_transpose3d_cl(h, up)
                ch = up.out_channels
                for blk in blocks:
                    skip, cs = skips.pop()
                    h = blk.forward_cl(F.cat_cl(h, skip, ch, cs), te)  # real channel counts: rows are padded to 8
            return F.from_channels_last_3d(F.conv3d_cl(h, self.outconv), self.outconv.out_channels)


def _fnet_weights_init(m: nn.Module) -> None:
    """F-Net initialisation: N(0, 0.02) conv weights, N(1, 0.02) BatchNorm weights."""
    if isinstance(m, (nn.Conv3d, nn.ConvTranspose3d)):
        nn.init.normal_(m.weight, 0.0, 0.02)
    elif isinstance(m, nn.BatchNorm3d):
        nn.init.normal_(m.weight, 1.0, 0.02)
        nn.init.constant_(m.bias, 0)


class Unet3d(UNet3DBase):
    """3D U-Net following Ounkomol et al. 2018 (F-Net): BatchNorm + ReLU double-conv blocks, Z downsampled too."""

    def __init__(self, in_channels: int = 1, out_channels: int = 1, depth: int = 4, mult_chan: int = 32,
                 in_stack_depth: int | None = None) -> None:
        dims = [mult_chan * (2**i) for i in range(depth + 1)]
        bottleneck = ConvBottleneck3D(dims[-1], residual=False, norm="batch", activation="relu")
        super().__init__(in_channels=in_channels, out_channels=out_channels, dims=dims, num_res_block=[1] * depth,
                         bottleneck=bottleneck, downsample_z=True, residual=False, norm="batch", activation="relu")
        self.in_stack_depth = in_stack_depth
        self.out_stack_depth = in_stack_depth
        self.apply(_fnet_weights_init)
