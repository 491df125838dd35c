"""Unet2d and its ConvBlock2D (VM/unet/unet2d.py:11-244, VM/components/conv_block_2d.py:11-388): same constructor, forward,
sub-module names and state_dict as the reference (including its quirks: `nn.Dropout2d(int(dropout))`, unregistered dropout /
upsampling modules, the always-registered `resid_conv`).

CPU tensors run in plain torch ops.  CUDA tensors run channels-last 16-bit through the sm_100a kernels as depth-1 volumes:
every Conv2d is the tcgen05 implicit-GEMM conv with a (1, kh, kw) filter, BatchNorm2d / InstanceNorm2d, dropout + activation,
2x2 average pooling and bilinear x2 upsampling are the HBM-bound kernels of the 2.5-D / 3-D families.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as TF
from torch import Tensor, nn

from . import functional as F
from .unet25d import _ACTS, conv_block_forward_cl
from .unext2 import resolve_compute_dtype


class ConvBlock2D(nn.Module):
    """`num_repeats` x [conv('same') -> (dropout) -> act -> norm] in `layer_order`, plus an optional residual path."""

    def __init__(self, in_filters, out_filters, dropout=False, norm="batch", residual=True, activation="relu",
                 transpose=False, kernel_size=3, num_repeats=3, filter_steps="first", layer_order="can"):
        super().__init__()
        self.in_filters, self.out_filters = in_filters, out_filters
        self.dropout, self.norm, self.residual = dropout, norm, residual
        self.activation, self.transpose = activation, transpose
        self.num_repeats, self.filter_steps, self.layer_order = num_repeats, filter_steps, layer_order
        ks = kernel_size
        if isinstance(ks, int):
            if ks % 2 != 1:
                raise ValueError("Kernel dims must be odd")
        elif isinstance(ks, tuple):
            if any(k % 2 != 1 for k in ks):
                raise ValueError("Kernel dims must be odd")
            if len(ks) != 2:
                raise ValueError("kernel_size length must be 2")
        else:
            raise AttributeError("'kernel_size' must be either int or tuple")
        self.kernel_size = kernel_size
        if self.dropout:  # conv_block_2d.py:101-104: the probability goes through int(); the modules stay unregistered
            self.drop_list = [nn.Dropout2d(int(self.dropout)) for _ in range(num_repeats)]
        steps = np.linspace(in_filters, out_filters, num_repeats + 1).astype(int)

        def norm_width(i):
            if filter_steps == "linear":
                return int(steps[i + 1])
            if filter_steps == "first":
                return int(steps[-1])
            return int(steps[0]) if i < num_repeats - 1 else int(steps[-1])

        self.norm_list = [None] * num_repeats
        if filter_steps in ("linear", "first", "last"):
            if norm == "batch":
                self.norm_list = [nn.BatchNorm2d(norm_width(i)) for i in range(num_repeats)]
            elif norm == "instance":
                self.norm_list = [nn.InstanceNorm2d(norm_width(i)) for i in range(num_repeats)]
        self._register(self.norm_list, f"{norm}_norm")
        self.conv_list = []
        if filter_steps == "linear":
            cls = nn.ConvTranspose2d if transpose else nn.Conv2d
            for i in range(num_repeats):
                pair = (int(steps[i]), int(steps[i + 1])) if i + 1 < num_repeats else (int(steps[i]), int(steps[-1]))
                self.conv_list.append(cls(pair[0], pair[1], kernel_size=kernel_size, padding="same"))
        elif filter_steps == "first":
            if transpose:
                raise NotImplementedError("PyTorch-side problem with 'same' padding in ConvTranspose2d.")
            for i in range(num_repeats):
                self.conv_list.append(nn.Conv2d(in_filters if i == 0 else out_filters, out_filters,
                                                kernel_size=kernel_size, padding="same"))
        elif filter_steps == "last":
            if transpose:
                raise NotImplementedError("Problem with 'same' padding in ConvTranspose2d.")
            for i in range(num_repeats):
                self.conv_list.append(nn.Conv2d(in_filters, out_filters if i == num_repeats - 1 else in_filters,
                                                kernel_size=kernel_size, padding="same"))
        self._register(self.conv_list, "Conv2d")
        self.resid_conv = nn.Conv2d(in_filters, out_filters, kernel_size=1, padding=0)
        self.act_list = []
        if activation in _ACTS:
            self.act_list = [_ACTS[activation]() for _ in range(num_repeats)]
        elif activation != "linear":
            raise NotImplementedError(f"Activation type {self.activation} not supported.")
        self._register(self.act_list, f"{activation}_act")

    def _register(self, modules, name):
        for i, m in enumerate(modules):
            self.add_module(f"{name}_{i}", m)

    register_modules = _register

    def forward(self, x: Tensor, validate_input: bool = False) -> Tensor:
        if x.is_cuda:
            raise NotImplementedError("ConvBlock2D: NCHW CUDA tensors enter through Unet2d.forward (channels-last sm_100a "
                                      "path) or forward_cl; there is no cuDNN fallback")
        if validate_input:
            kh, kw = (self.kernel_size,) * 2 if isinstance(self.kernel_size, int) else self.kernel_size
            if not (x.shape[-1] > kw and x.shape[-2] > kh):
                raise ValueError(f"Input size {x.shape} too small for kernel of size {self.kernel_size}")
        x_0 = x
        for i in range(self.num_repeats):
            for layer in self.layer_order:
                if layer == "c":
                    x = self.conv_list[i](x)
                    if self.dropout:
                        x = self.drop_list[i](x)
                elif layer == "a":
                    if i < self.num_repeats - 1 or self.activation != "linear":
                        x = self.act_list[i](x)
                elif layer == "n" and self.norm_list[i]:
                    x = self.norm_list[i](x)
        if self.residual:
            if self.in_filters > self.out_filters:
                x_0 = self.resid_conv(x_0)
            elif self.in_filters < self.out_filters:
                x_0 = TF.pad(x_0, (*[0] * 4, self.out_filters - self.in_filters, *[0] * 3), mode="constant", value=0)
            x = torch.add(x_0, x)
        return x

    # ---- sm_100a path: rows [N, 1, H, W, C] ----
    def _conv_cl(self, x: Tensor, i: int, pad) -> Tensor:
        conv = self.conv_list[i]
        if self.transpose:
            raise NotImplementedError("sm_100a ConvBlock2D: transpose=True (the reference cannot build it either: "
                                      "ConvTranspose2d rejects padding='same')")
        return F.Conv3dFn.apply(x, conv.weight.unsqueeze(2), conv.bias, (1, 1, 1), pad)

    def _act_cl(self, x: Tensor, scale) -> Tensor:
        if self.activation == "relu":
            return F.scale_relu_cl(x, scale, True)
        return F.scale_act_cl(x, scale, self.activation)

    def _has_act(self, i: int) -> bool:
        return self.activation != "linear"

    def forward_cl(self, x: Tensor) -> Tensor:
        kh, kw = (self.kernel_size,) * 2 if isinstance(self.kernel_size, int) else self.kernel_size
        return conv_block_forward_cl(self, x, (0, kh // 2, kw // 2), float(int(self.dropout)) if self.dropout else 0.0)


class Unet2d(nn.Module):
    """2D U-Net (VM/unet/unet2d.py:11-244): (B, C, 1, H, W) in and out."""

    def __name__(self):
        return "Unet2d"

    def __init__(self, in_channels=1, out_channels=1, kernel_size=(3, 3), residual=False, dropout=0.2, num_blocks=4,
                 num_block_layers=2, num_filters=(), task="seg"):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.residual, self.dropout = kernel_size, residual, dropout
        self.num_blocks, self.num_block_layers, self.task = num_blocks, num_block_layers, task
        self.block_padding = "same"
        self.bottom_block_spatial = False
        if len(num_filters) != 0:
            assert len(num_filters) == num_blocks + 1, (
                "Length of num_filters must be equal to num_blocks + 1 (number of convolutional blocks per path)."
            )
            self.num_filters = list(num_filters)
        else:
            self.num_filters = [pow(2, i) * 16 for i in range(num_blocks + 1)]
        nf = self.num_filters
        down_f = [in_channels] + nf
        up_f = [nf[-(i + 1)] + nf[-(i + 2)] for i in range(len(nf) - 1)] + [out_channels]
        blk = dict(dropout=dropout, residual=residual, activation="relu", kernel_size=kernel_size,
                   num_repeats=num_block_layers)
        self.down_list = [nn.AvgPool2d(kernel_size=2) for _ in range(num_blocks)]
        self._register(self.down_list, "down_samp")
        self.up_list = [nn.Upsample(mode="bilinear", scale_factor=2, align_corners=False) for _ in range(num_blocks)]
        self.down_conv_blocks = [ConvBlock2D(down_f[i], down_f[i + 1], **blk) for i in range(num_blocks)]
        self._register(self.down_conv_blocks, "down_conv_block")
        self.bottom_transition_block = ConvBlock2D(nf[-2], nf[-1], **blk)
        self.up_conv_blocks = [ConvBlock2D(up_f[i], down_f[-(i + 2)], **blk) for i in range(num_blocks)]
        self._register(self.up_conv_blocks, "up_conv_block")
        self.terminal_block = ConvBlock2D(down_f[1], out_channels, dropout=dropout, residual=False,
                                          activation="linear" if task == "reg" else "relu", num_repeats=1, norm="none",
                                          kernel_size=kernel_size)

    def _register(self, modules, name):
        for i, m in enumerate(modules):
            self.add_module(f"{name}_{i}", m)

    register_modules = _register

    compute_dtype: torch.dtype | None = None

    def _forward_sm100(self, x: Tensor) -> Tensor:
        dt = resolve_compute_dtype(x, self.compute_dtype)
        F.ops.ACTIVE_PACKS = None
        F.ops.STEP.begin(x.device, torch.is_grad_enabled())
        with torch.autocast("cuda", enabled=False):
            h = F.to_channels_last_3d(x, dt)  # [N, 1, H, W, C]
            skips = []
            for blk in self.down_conv_blocks:
                h = blk.forward_cl(h)
                skips.append(h)
                h = F.avgpool_hw2_cl(h)
            h = self.bottom_transition_block.forward_cl(h)
            prev = self.bottom_transition_block.out_filters
            for i, blk in enumerate(self.up_conv_blocks):
                skip = skips[-(i + 1)]
                h = blk.forward_cl(F.cat_cl(F.upsample2x_hw_cl(h), skip, prev, self.down_conv_blocks[-(i + 1)].out_filters))
                prev = blk.out_filters
            h = self.terminal_block.forward_cl(h)
            return F.from_channels_last_3d(h, self.out_channels)

    def forward(self, x: Tensor, validate_input: bool = False) -> Tensor:
        if validate_input:
            assert x.shape[-1] == x.shape[-2], "Input must be square in xy"
            assert x.shape[-4] == self.in_channels, f"Input channels must equal network input channels: {self.in_channels}"
        if x.is_cuda:
            if x.shape[2] != 1:
                raise ValueError(f"Unet2d takes (B, C, 1, H, W) input, got {tuple(x.shape)}")
            return self._forward_sm100(x)
        x = x.squeeze(2)
        skips = []
        for i in range(self.num_blocks):
            x = self.down_conv_blocks[i](x, validate_input=validate_input)
            skips.append(x)
            x = self.down_list[i](x)
        x = self.bottom_transition_block(x)
        for i in range(self.num_blocks):
            x = self.up_list[i](x)
            x = torch.cat([x, skips[-1 * (i + 1)]], 1)
            x = self.up_conv_blocks[i](x, validate_input=validate_input)
        return self.terminal_block(x).unsqueeze(2)
