"""Unet25d and its ConvBlock3D (VM/unet/unet25d.py:11-251, VM/components/conv_block_3d.py:11-346): same constructor,
forward, sub-module names and state_dict as the reference.

This family is BASELINE config #1 (CPU, fp32 parity): it runs in plain torch ops.  CUDA tensors are rejected with
NotImplementedError until its sm_100a kernels land - there is no silent cuDNN fallback.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as TF
from torch import Tensor, nn

_ACTS = {"relu": nn.ReLU, "leakyrelu": nn.LeakyReLU, "elu": nn.ELU, "selu": nn.SELU}


def _reject_cuda(x: Tensor, what: str) -> None:
    if x.is_cuda:
        raise NotImplementedError(
            f"{what}: no sm_100a kernels yet (BASELINE config 1 is the CPU parity configuration); "
            "move the module and its input to the CPU"
        )


class ConvBlock3D(nn.Module):
    """`num_repeats` x [pad -> conv -> (dropout) -> act -> norm] in the order given by `layer_order`, plus an optional
    residual path (1x1x1 conv when channels shrink, zero channels prepended when they grow)."""

    def __init__(self, in_filters, out_filters, dropout=False, norm="batch", residual=True, activation="relu",
                 transpose=False, kernel_size=(3, 3, 3), num_repeats=3, filter_steps="first", layer_order="can",
                 padding=None):
        super().__init__()
        self.in_filters, self.out_filters = in_filters, out_filters
        self.dropout, self.norm, self.residual = dropout, norm, residual
        self.activation, self.transpose = activation, transpose
        self.num_repeats, self.filter_steps, self.layer_order = num_repeats, filter_steps, layer_order
        if isinstance(kernel_size, int):
            if kernel_size % 2 != 1:
                raise ValueError("Kernel dims must be odd")
            ks = (kernel_size,) * 3
        elif isinstance(kernel_size, tuple):
            if any(k % 2 != 1 for k in kernel_size):
                raise ValueError("Kernel dims must be odd")
            if len(kernel_size) != 3:
                raise ValueError("kernel_size length must be 3")
            ks = kernel_size
        else:
            raise AttributeError("'kernel_size' must be either int or tuple")
        self.kernel_size = kernel_size
        self.pad_type = "same"
        pad3 = (ks[2] // 2, ks[1] // 2, ks[0] // 2)  # (W, H, D) halves, F.pad order
        if padding == "valid":
            pad3 = (0, 0, 0)
        elif isinstance(padding, tuple):
            pad3 = padding
        self.padding = tuple(pad3[i // 2] for i in range(6)) + (0,) * 4

        widths = np.linspace(in_filters, out_filters, num_repeats + 1).astype(int)

        def norm_width(i):
            if filter_steps == "linear":
                return int(widths[i + 1])
            if filter_steps == "first":
                return int(widths[-1])
            return int(widths[0]) if i < num_repeats - 1 else int(widths[-1])

        def conv_io(i):
            if filter_steps == "linear":
                return int(widths[i]), int(widths[i + 1]) if i + 1 < num_repeats else int(widths[-1])
            if filter_steps == "first":
                return (in_filters if i == 0 else out_filters), out_filters
            return in_filters, (out_filters if i == num_repeats - 1 else in_filters)

        self.drop_list = [nn.Dropout3d(dropout) for _ in range(num_repeats)] if dropout else []
        self._register(self.drop_list, "dropout")
        if norm == "batch":
            self.norm_list = [nn.BatchNorm3d(norm_width(i)) for i in range(num_repeats)]
        elif norm == "instance":
            self.norm_list = [nn.InstanceNorm3d(norm_width(i)) for i in range(num_repeats)]
        else:
            self.norm_list = [None] * num_repeats
        self._register(self.norm_list, f"{norm}_norm")
        conv_cls = nn.ConvTranspose3d if transpose else nn.Conv3d
        self.conv_list = [conv_cls(*conv_io(i), kernel_size=kernel_size) for i in range(num_repeats)]
        self._register(self.conv_list, "Conv3d")
        # always registered, even when unused (reference behaviour: an unused parameter under DDP)
        self.resid_conv = nn.Conv3d(in_filters, out_filters, kernel_size=1, padding=0)
        if activation in _ACTS:
            self.act_list = [_ACTS[activation]() for _ in range(num_repeats)]
        elif activation == "linear":
            self.act_list = []
        else:
            raise NotImplementedError(f"Activation type {activation} not supported.")
        self._register(self.act_list, f"{activation}_act")

    def _register(self, modules, name):
        for i, m in enumerate(modules):
            self.add_module(f"{name}_{i}", m)

    register_modules = _register

    def forward(self, x: Tensor) -> Tensor:
        _reject_cuda(x, "ConvBlock3D")
        x0 = x
        for i in range(self.num_repeats):
            for layer in self.layer_order:
                if layer == "c":
                    x = self.conv_list[i](TF.pad(x, self.padding, "constant", 0))
                    if self.dropout:
                        x = self.drop_list[i](x)
                elif layer == "a":
                    if i < self.num_repeats - 1 or self.activation != "linear":
                        x = self.act_list[i](x)
                elif layer == "n" and self.norm_list[i] is not None:
                    x = self.norm_list[i](x)
        if self.residual:
            if self.in_filters > self.out_filters:
                x0 = self.resid_conv(x0)
            elif self.in_filters < self.out_filters:
                x0 = TF.pad(x0, (0,) * 6 + (self.out_filters - self.in_filters, 0, 0, 0), mode="constant", value=0)
            x = x + x0
        return x


class Unet25d(nn.Module):
    """2.5D U-Net: 3-D conv encoder, Z collapsed by (1+Din-Dout,1,1) convs on the skips, 2-D (1,k,k) decoder."""

    def __init__(self, in_channels=1, out_channels=1, in_stack_depth=5, out_stack_depth=1, xy_kernel_size=(3, 3),
                 residual=False, dropout=0.2, num_blocks=4, num_block_layers=2, num_filters=(), task="seg"):
        super().__init__()
        self.in_channels, self.num_blocks = in_channels, num_blocks
        self.kernel_size, self.residual = xy_kernel_size, residual
        assert dropout >= 0 and dropout <= 0.5, f"Dropout {dropout} not in allowed range: [0, 0.5]"
        self.dropout, self.task = dropout, task
        self.debug_mode = False
        self.block_padding = "same"
        self.bottom_block_spatial = False
        if len(num_filters) != 0:
            assert len(num_filters) == num_blocks + 1, (
                "Length of num_filters must be equal to num_blocks + 1 (number of convolutional blocks per path)."
            )
            self.num_filters = list(num_filters)
        else:
            self.num_filters = [16 * 2**i for i in range(num_blocks + 1)]
        nf = self.num_filters
        down_f = [in_channels] + nf
        up_f = [nf[-(i + 1)] + nf[-(i + 2)] for i in range(len(nf) - 1)] + [out_channels]
        kz = 1 + in_stack_depth - out_stack_depth
        ky, kx = xy_kernel_size

        self.down_list = [nn.AvgPool3d(kernel_size=(1, 2, 2), stride=(1, 2, 2)) for _ in range(num_blocks)]
        self._register(self.down_list, "down_samp")
        self.up_list = [nn.Upsample(scale_factor=(1, 2, 2), mode="trilinear", align_corners=False) for _ in range(num_blocks)]
        self.down_conv_blocks = [
            ConvBlock3D(down_f[i], down_f[i + 1], dropout=dropout, residual=residual, activation="relu",
                        kernel_size=(3, ky, kx), num_repeats=num_block_layers)
            for i in range(num_blocks)
        ]
        self._register(self.down_conv_blocks, "down_conv_block")
        self.bottom_transition_block = nn.Conv3d(nf[-2], nf[-1], kernel_size=(kz, 1, 1), padding=0)
        self.up_conv_blocks = [
            ConvBlock3D(up_f[i], down_f[-(i + 2)], dropout=dropout, residual=residual, activation="relu",
                        kernel_size=(1, ky, kx), num_repeats=num_block_layers)
            for i in range(num_blocks)
        ]
        self._register(self.up_conv_blocks, "up_conv_block")
        self.skip_conv_layers = [nn.Conv3d(down_f[i + 1], down_f[i + 1], kernel_size=(kz, 1, 1)) for i in range(num_blocks)]
        self._register(self.skip_conv_layers, "skip_conv_layer")
        if task == "reg":
            self.terminal_block = ConvBlock3D(down_f[1], out_channels, dropout=False, residual=False,
                                              activation="linear", kernel_size=(1, 3, 3), norm="none", num_repeats=1)
        else:
            self.terminal_block = ConvBlock3D(down_f[1], out_channels, dropout=dropout, residual=False,
                                              activation="relu", kernel_size=(1, 3, 3), num_repeats=1)
        self.log_save_folder = None

    def _register(self, modules, name):
        for i, m in enumerate(modules):
            self.add_module(f"{name}_{i}", m)

    register_modules = _register

    def forward(self, x: Tensor) -> Tensor:
        _reject_cuda(x, "Unet25d")
        skips = []
        for blk, down in zip(self.down_conv_blocks, self.down_list):
            x = blk(x)
            skips.append(x)
            x = down(x)
        x = self.bottom_transition_block(x)
        skips = [conv(s) for conv, s in zip(self.skip_conv_layers, skips)]
        for i, (up, blk) in enumerate(zip(self.up_list, self.up_conv_blocks)):
            x = blk(torch.cat([up(x), skips[-(i + 1)]], 1))
        return self.terminal_block(x)
