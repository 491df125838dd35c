"""Restatement of the monai 1.5.2 pieces the reference composes (TEST INFRASTRUCTURE, see oracle/__init__.py).

Call sites: VM/components/blocks.py:7-8,118-146 (UpSample pixelshuffle; ResidualUnit/get_conv_layer only on the
known-broken deconv path), VM/components/heads.py:15-16,607-628 (Convolution, UpSample, normal_init).
Semantics: SURVEY.md Appendix B.2.
"""
from __future__ import annotations

import types

import torch
from torch import nn


class UpSample(nn.Module):
    """mode='pixelshuffle' only (SubpixelUpsample)."""

    def __init__(self, spatial_dims, in_channels=None, out_channels=None, scale_factor=2, mode="pixelshuffle",
                 pre_conv="default", apply_pad_pool=True, **kw):
        super().__init__()
        if str(mode) != "pixelshuffle" or spatial_dims != 2:
            raise NotImplementedError("oracle restates the 2-D pixelshuffle UpSample only")
        r = scale_factor
        if pre_conv == "default":
            self.conv_block = nn.Conv2d(in_channels, (out_channels or in_channels) * r * r, 3, 1, 1)
        elif pre_conv is None or pre_conv == "None":
            self.conv_block = nn.Identity()
        else:
            self.conv_block = pre_conv
        self.shuffle = nn.PixelShuffle(r)
        self.pad_pool = nn.Identity()
        if apply_pad_pool:
            self.pad_pool = nn.Sequential(nn.ConstantPad2d((r - 1, 0) * 2, 0.0), nn.AvgPool2d(kernel_size=r, stride=1))

    def forward(self, x):
        return self.pad_pool(self.shuffle(self.conv_block(x)))


class ADN(nn.Sequential):
    def __init__(self, channels):
        super().__init__()
        self.add_module("N", nn.InstanceNorm3d(channels))
        self.add_module("A", nn.PReLU())


class Convolution(nn.Sequential):
    """spatial_dims=3 Conv3d + ADN('NDA': InstanceNorm3d, no dropout, PReLU)."""

    def __init__(self, spatial_dims, in_channels, out_channels, strides=1, kernel_size=3, padding=None, **kw):
        super().__init__()
        if spatial_dims != 3:
            raise NotImplementedError
        self.add_module("conv", nn.Conv3d(in_channels, out_channels, kernel_size, strides, padding, bias=True))
        self.add_module("adn", ADN(out_channels))


def normal_init(m, std=0.02, normal_func=torch.nn.init.normal_):
    cname = m.__class__.__name__
    if getattr(m, "weight", None) is not None and (cname.find("Conv") != -1 or cname.find("Linear") != -1):
        normal_func(m.weight.data, 0.0, std)
        if getattr(m, "bias", None) is not None:
            nn.init.constant_(m.bias.data, 0.0)
    elif cname.find("BatchNorm") != -1:
        normal_func(m.weight.data, 1.0, std)
        nn.init.constant_(m.bias.data, 0)


def _unsupported(*a, **k):
    raise NotImplementedError("decoder_mode='deconv' is known-broken in the reference (test_unext2.py:45-57)")


def as_modules() -> dict[str, types.ModuleType]:
    monai = types.ModuleType("monai")
    networks = types.ModuleType("monai.networks")
    blocks = types.ModuleType("monai.networks.blocks")
    dyn = types.ModuleType("monai.networks.blocks.dynunet_block")
    utils = types.ModuleType("monai.networks.utils")
    blocks.UpSample, blocks.Convolution, blocks.ResidualUnit = UpSample, Convolution, _unsupported
    dyn.get_conv_layer = _unsupported
    utils.normal_init = normal_init
    monai.networks, networks.blocks, networks.utils, blocks.dynunet_block = networks, blocks, utils, dyn
    monai.__oracle_restatement__ = True
    return {"monai": monai, "monai.networks": networks, "monai.networks.blocks": blocks,
            "monai.networks.blocks.dynunet_block": dyn, "monai.networks.utils": utils}
