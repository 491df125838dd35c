"""Unet25d and its ConvBlock3D (VM/unet/unet25d.py:11-251, VM/components/conv_block_3d.py:11-346): same constructor,
forward, sub-module names and state_dict as the reference.

CPU tensors (BASELINE config #1, fp32 parity) run in plain torch ops.  CUDA tensors run channels-last 16-bit through the
sm_100a kernels: every convolution on the tcgen05 GEMM (implicit-GEMM TMA form where the geometry tiles, im2col lowering
otherwise), Dropout3d + ReLU, BatchNorm3d, (1,2,2) average pooling and trilinear upsampling as HBM-bound kernels.
BatchNorm3d / InstanceNorm3d / no norm, relu | leakyrelu | elu | selu | linear, any layer order and transpose=True all run
on the kernels; there is no cuDNN fallback.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as TF
from torch import Tensor, nn

from . import functional as F
from .unext2 import resolve_compute_dtype

_ACTS = {"relu": nn.ReLU, "leakyrelu": nn.LeakyReLU, "elu": nn.ELU, "selu": nn.SELU}


def _reject_cuda(x: Tensor, what: str) -> None:
    if x.is_cuda:
        raise NotImplementedError(
            f"{what}: NCDHW CUDA tensors enter through Unet25d.forward (channels-last sm_100a path) or "
            "ConvBlock3D.forward_cl; there is no cuDNN fallback"
        )


def conv_block_forward_cl(blk, x: Tensor, pad, drop_p: float) -> Tensor:
    """The layer loop shared by ConvBlock3D and ConvBlock2D (conv_block_3d.py:261-298, conv_block_2d.py forward) on
    channels-last rows: `blk` provides _conv_cl(x, i, pad), _act_cl, _has_act, norm_list, norm, activation, layer_order,
    residual, resid_conv, in_filters / out_filters; drop_p is the channel-dropout probability."""
    x0 = x
    for i in range(blk.num_repeats):
        order = blk.layer_order
        k = 0
        while k < len(order):
            layer = order[k]
            act_next = k + 1 < len(order) and order[k + 1] == "a" and blk._has_act(i)
            if layer == "c":
                x = blk._conv_cl(x, i, pad)
                scale = F.dropout3d_scale(x, drop_p, blk.training) if drop_p else None
                if act_next:  # conv -> dropout -> activation in one pass over the tensor
                    x = blk._act_cl(x, scale)
                    k += 1
                elif scale is not None:
                    x = F.scale_relu_cl(x, scale, False)
            elif layer == "a":
                if blk._has_act(i):
                    x = blk._act_cl(x, None)
            elif layer == "n" and blk.norm_list[i] is not None:
                if blk.norm == "instance":  # norm (-> activation) in one pass
                    x = F.instancenorm_act_cl(x, blk.norm_list[i].eps, blk.activation if act_next else "none")
                    k += int(act_next)
                else:
                    fuse = act_next and blk.activation == "relu"
                    x = F.batchnorm_act_cl(x, blk.norm_list[i], relu=fuse)
                    k += int(fuse)
            k += 1
    if blk.residual:
        if blk.in_filters > blk.out_filters:
            x0 = F.Conv3dFn.apply(x0, blk.resid_conv.weight.reshape(blk.out_filters, blk.in_filters, 1, 1, 1),
                                  blk.resid_conv.bias, (1, 1, 1), (0, 0, 0))
        elif blk.in_filters < blk.out_filters:
            grow = blk.out_filters - blk.in_filters
            zeros = torch.zeros((*x0.shape[:-1], -(-grow // 8) * 8), device=x0.device, dtype=x0.dtype)
            # identity lands on the LAST in_filters channels (conv_block_3d.py:281-287); real channel counts so that
            # padded rows are compacted to [zeros | x0 | pad]
            x0 = F.cat_cl(zeros, x0, grow, blk.in_filters)
        x = F.add_cl(x, x0)
    return x


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

    def _conv_cl(self, x: Tensor, i: int, pad) -> Tensor:
        conv = self.conv_list[i]
        if not self.transpose:
            return F.Conv3dFn.apply(x, conv.weight, conv.bias, (1, 1, 1), pad)
        # F.pad(p) then ConvTranspose3d(k, stride 1, padding 0) == the correlation with the flipped, transposed filter
        # under zero padding k - 1 + p (zero padding composes); autograd carries dW back through flip / transpose
        w = conv.weight.flip(2, 3, 4).transpose(0, 1)
        full = tuple(k - 1 + p for k, p in zip(conv.kernel_size, pad))
        return F.Conv3dFn.apply(x, w, conv.bias, (1, 1, 1), full)

    def _act_cl(self, x: Tensor, scale) -> Tensor:
        if self.activation == "relu":
            return F.scale_relu_cl(x, scale, True)
        return F.scale_act_cl(x, scale, self.activation)

    def forward_cl(self, x: Tensor) -> Tensor:
        """sm_100a path on channels-last rows [N,D,H,W,C] (C padded to a multiple of 8)."""
        pw, ph, pd = self.padding[0], self.padding[2], self.padding[4]
        return conv_block_forward_cl(self, x, (pd, ph, pw), self.dropout if self.dropout else 0.0)

    def _has_act(self, i: int) -> bool:
        return self.activation != "linear"

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

    compute_dtype: torch.dtype | None = None  # sm_100a arithmetic type override (default: autocast / input dtype)

    def _forward_sm100(self, x: Tensor) -> Tensor:
        dt = resolve_compute_dtype(x, self.compute_dtype)
        F.ops.ACTIVE_PACKS = None  # weight packs are scoped to the model that registered them
        F.ops.STEP.begin(x.device, torch.is_grad_enabled())  # one zero-filled allocation for the step's accumulators
        with torch.autocast("cuda", enabled=False):
            h = F.to_channels_last_3d(x, dt)
            skips = []
            for blk in self.down_conv_blocks:
                h = blk.forward_cl(h)
                skips.append(h)
                h = F.avgpool_hw2_cl(h)
            h = F.conv3d_cl(h, self.bottom_transition_block)
            skips = [F.conv3d_cl(s, conv) for conv, s in zip(self.skip_conv_layers, skips)]
            for i, blk in enumerate(self.up_conv_blocks):
                ch = self.up_conv_blocks[i - 1].out_filters if i > 0 else self.bottom_transition_block.out_channels
                h = blk.forward_cl(F.cat_cl(F.upsample2x_hw_cl(h), skips[-(i + 1)], ch,
                                            self.skip_conv_layers[-(i + 1)].out_channels))
            h = self.terminal_block.forward_cl(h)
            return F.from_channels_last_3d(h, self.terminal_block.out_filters)

    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda:
            return self._forward_sm100(x)
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
