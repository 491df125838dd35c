"""Restatement of the timm 1.0.27 pieces the reference composes (TEST INFRASTRUCTURE, see oracle/__init__.py).

Call sites in the reference: VM/unet/unext2.py:40-49 (create_model features_only), VM/components/blocks.py:60-73
(ConvNeXtStage, LayerNorm2d, LayerNorm, _init_weights), VM/contrastive/encoder.py:93-124 (create_model,
encoder.stem, encoder.head.fc, encoder.num_features).  Semantics written down in SURVEY.md Appendix B.1.
"""
from __future__ import annotations

import types

import torch
import torch.nn.functional as F
from torch import nn


class LayerNorm(nn.LayerNorm):
    def __init__(self, num_channels, eps=1e-6, affine=True):
        super().__init__(num_channels, eps=eps, elementwise_affine=affine)


class LayerNorm2d(nn.LayerNorm):
    """LayerNorm over the channel dim of NCHW."""

    def __init__(self, num_channels, eps=1e-6, affine=True):
        super().__init__(num_channels, eps=eps, elementwise_affine=affine)

    def forward(self, x):
        x = x.permute(0, 2, 3, 1)
        x = F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        return x.permute(0, 3, 1, 2)


class GlobalResponseNorm(nn.Module):
    def __init__(self, dim, eps=1e-6, channels_last=True):
        super().__init__()
        self.eps = eps
        if channels_last:
            self.spatial_dim, self.channel_dim, self.wb_shape = (1, 2), -1, (1, 1, 1, -1)
        else:
            self.spatial_dim, self.channel_dim, self.wb_shape = (2, 3), 1, (1, -1, 1, 1)
        self.weight = nn.Parameter(torch.zeros(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def forward(self, x):
        x_g = x.norm(p=2, dim=self.spatial_dim, keepdim=True)
        x_n = x_g / (x_g.mean(dim=self.channel_dim, keepdim=True) + self.eps)
        return x + torch.addcmul(self.bias.view(self.wb_shape), self.weight.view(self.wb_shape), x * x_n)


class Mlp(nn.Module):
    def __init__(self, in_features, hidden_features, use_conv=False):
        super().__init__()
        lin = (lambda i, o: nn.Conv2d(i, o, 1)) if use_conv else nn.Linear
        self.fc1 = lin(in_features, hidden_features)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(0.0)
        self.norm = nn.Identity()
        self.fc2 = lin(hidden_features, in_features)
        self.drop2 = nn.Dropout(0.0)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class GlobalResponseNormMlp(nn.Module):
    def __init__(self, in_features, hidden_features, out_features=None, use_conv=False):
        super().__init__()
        lin = (lambda i, o: nn.Conv2d(i, o, 1)) if use_conv else nn.Linear
        self.fc1 = lin(in_features, hidden_features)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(0.0)
        self.grn = GlobalResponseNorm(hidden_features, channels_last=not use_conv)
        self.fc2 = lin(hidden_features, out_features or in_features)
        self.drop2 = nn.Dropout(0.0)

    def forward(self, x):
        return self.drop2(self.fc2(self.grn(self.drop1(self.act(self.fc1(x))))))


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


class ConvNeXtBlock(nn.Module):
    def __init__(self, in_chs, out_chs=None, kernel_size=7, stride=1, dilation=(1, 1), mlp_ratio=4, conv_mlp=False,
                 conv_bias=True, use_grn=False, ls_init_value=1e-6, act_layer="gelu", norm_layer=None, drop_path=0.0):
        super().__init__()
        out_chs = out_chs or in_chs
        if not norm_layer:
            norm_layer = LayerNorm2d if conv_mlp else LayerNorm
        mlp_layer = GlobalResponseNormMlp if use_grn else Mlp
        self.use_conv_mlp = conv_mlp
        self.conv_dw = nn.Conv2d(in_chs, out_chs, kernel_size, stride=stride, padding=kernel_size // 2,
                                 groups=in_chs, bias=conv_bias)
        self.norm = norm_layer(out_chs)
        self.mlp = mlp_layer(out_chs, int(mlp_ratio * out_chs), use_conv=conv_mlp)
        self.gamma = nn.Parameter(ls_init_value * torch.ones(out_chs)) if ls_init_value is not None else None
        self.shortcut = nn.Identity()
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()

    def forward(self, x):
        shortcut = x
        x = self.conv_dw(x)
        if self.use_conv_mlp:
            x = self.norm(x)
            x = self.mlp(x)
        else:
            x = x.permute(0, 2, 3, 1)
            x = self.norm(x)
            x = self.mlp(x)
            x = x.permute(0, 3, 1, 2)
        if self.gamma is not None:
            x = x.mul(self.gamma.reshape(1, -1, 1, 1))
        return self.drop_path(x) + self.shortcut(shortcut)


class ConvNeXtStage(nn.Module):
    def __init__(self, in_chs, out_chs, kernel_size=7, stride=2, depth=2, dilation=(1, 1), drop_path_rates=None,
                 ls_init_value=1.0, conv_mlp=False, conv_bias=True, use_grn=False, act_layer="gelu",
                 norm_layer=None, norm_layer_cl=None):
        super().__init__()
        self.grad_checkpointing = False
        if in_chs != out_chs or stride > 1:
            ds_ks = 2 if stride > 1 else 1
            self.downsample = nn.Sequential(
                norm_layer(in_chs),
                nn.Conv2d(in_chs, out_chs, kernel_size=ds_ks, stride=stride, padding=0, bias=conv_bias),
            )
            in_chs = out_chs
        else:
            self.downsample = nn.Identity()
        drop_path_rates = drop_path_rates or [0.0] * depth
        blocks = []
        for i in range(depth):
            blocks.append(ConvNeXtBlock(in_chs=in_chs, out_chs=out_chs, kernel_size=kernel_size,
                                        drop_path=drop_path_rates[i], ls_init_value=ls_init_value, conv_mlp=conv_mlp,
                                        conv_bias=conv_bias, use_grn=use_grn, act_layer=act_layer,
                                        norm_layer=norm_layer if conv_mlp else norm_layer_cl))
            in_chs = out_chs
        self.blocks = nn.Sequential(*blocks)

    def forward(self, x):
        return self.blocks(self.downsample(x))


def _init_weights(module, name=None, head_init_scale=1.0):
    if isinstance(module, nn.Conv2d):
        nn.init.trunc_normal_(module.weight, std=0.02)
        if module.bias is not None:
            nn.init.zeros_(module.bias)
    elif isinstance(module, nn.Linear):
        nn.init.trunc_normal_(module.weight, std=0.02)
        nn.init.zeros_(module.bias)
        if name and "head." in name:
            module.weight.data.mul_(head_init_scale)
            module.bias.data.mul_(head_init_scale)


_CFGS = {
    "convnext_tiny": dict(depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), use_grn=False, ls_init_value=1e-6, conv_mlp=False),
    "convnextv2_tiny": dict(depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), use_grn=True, ls_init_value=None, conv_mlp=False),
    "convnextv2_atto": dict(depths=(2, 2, 6, 2), dims=(40, 80, 160, 320), use_grn=True, ls_init_value=None, conv_mlp=True),
    "convnextv2_femto": dict(depths=(2, 2, 6, 2), dims=(48, 96, 192, 384), use_grn=True, ls_init_value=None, conv_mlp=True),
    "convnextv2_pico": dict(depths=(2, 2, 6, 2), dims=(64, 128, 256, 512), use_grn=True, ls_init_value=None, conv_mlp=True),
    "convnextv2_nano": dict(depths=(2, 2, 8, 2), dims=(80, 160, 320, 640), use_grn=True, ls_init_value=None, conv_mlp=True),
}


class _Head(nn.Module):
    """timm NormMlpClassifierHead (hidden_size=None)."""

    def __init__(self, in_features, num_classes):
        super().__init__()
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.norm = LayerNorm2d(in_features)
        self.flatten = nn.Flatten(1)
        self.pre_logits = nn.Identity()
        self.drop = nn.Dropout(0.0)
        self.fc = nn.Linear(in_features, num_classes) if num_classes > 0 else nn.Identity()

    def forward(self, x):
        return self.fc(self.drop(self.pre_logits(self.flatten(self.norm(self.global_pool(x))))))


class ConvNeXt(nn.Module):
    def __init__(self, in_chans=3, num_classes=1000, depths=(3, 3, 9, 3), dims=(96, 192, 384, 768), ls_init_value=1e-6,
                 conv_mlp=False, use_grn=False, drop_path_rate=0.0):
        super().__init__()
        norm_layer, norm_layer_cl = LayerNorm2d, (LayerNorm2d if conv_mlp else LayerNorm)
        self.stem = nn.Sequential(nn.Conv2d(in_chans, dims[0], kernel_size=4, stride=4), norm_layer(dims[0]))
        dp = [x.tolist() for x in torch.linspace(0, drop_path_rate, sum(depths)).split(depths)]
        stages, prev = [], dims[0]
        for i in range(4):
            stages.append(ConvNeXtStage(prev, dims[i], kernel_size=7, stride=2 if i > 0 else 1, depth=depths[i],
                                        drop_path_rates=dp[i], ls_init_value=ls_init_value, conv_mlp=conv_mlp,
                                        use_grn=use_grn, norm_layer=norm_layer, norm_layer_cl=norm_layer_cl))
            prev = dims[i]
        self.stages = nn.Sequential(*stages)
        self.num_features = prev
        self.norm_pre = nn.Identity()
        self.head = _Head(prev, num_classes)
        for n, m in self.named_modules():
            _init_weights(m, n)

    def forward(self, x):
        return self.head(self.norm_pre(self.stages(self.stem(x))))


class _FeatureInfo:
    def __init__(self, chans):
        self._c = list(chans)

    def channels(self):
        return list(self._c)


class FeatureListNet(nn.Module):
    """features_only=True wrapper with flatten_sequential=True naming: stem_0, stem_1, stages_0..3."""

    def __init__(self, model: ConvNeXt, dims):
        super().__init__()
        self.stem_0, self.stem_1 = model.stem[0], model.stem[1]
        for i, s in enumerate(model.stages):
            setattr(self, f"stages_{i}", s)
        self.feature_info = _FeatureInfo(dims)

    def forward(self, x):
        x = self.stem_1(self.stem_0(x))
        out = []
        for i in range(4):
            x = getattr(self, f"stages_{i}")(x)
            out.append(x)
        return out


def create_model(name, pretrained=False, features_only=False, drop_path_rate=0.0, num_classes=1000, **kw):
    if pretrained:
        raise RuntimeError("oracle restatement has no pretrained weights")
    cfg = _CFGS[name]
    m = ConvNeXt(num_classes=num_classes, depths=cfg["depths"], dims=cfg["dims"], ls_init_value=cfg["ls_init_value"],
                 conv_mlp=cfg["conv_mlp"], use_grn=cfg["use_grn"], drop_path_rate=drop_path_rate)
    return FeatureListNet(m, cfg["dims"]) if features_only else m


class Downsample(nn.Module):
    """timm.models.convnext.Downsample (used by VM/unet/fcmae.py:190-193 only when channels / stride change)."""

    def __init__(self, in_chs, out_chs, stride=1, dilation=1):
        super().__init__()
        avg_stride = stride if dilation == 1 else 1
        if stride > 1 or dilation > 1:
            self.pool = nn.AvgPool2d(2, avg_stride, ceil_mode=True, count_include_pad=False)
        else:
            self.pool = nn.Identity()
        self.conv = nn.Conv2d(in_chs, out_chs, 1, stride=1) if in_chs != out_chs else nn.Identity()

    def forward(self, x):
        return self.conv(self.pool(x))


def create_conv2d(in_channels, out_channels, kernel_size, **kwargs):
    """timm.layers.create_conv2d for the one call in VM/unet/fcmae.py:176-182: depthwise, padding '' (= symmetric 'same'
    padding for stride 1, dilation 1), bias True."""
    depthwise = kwargs.pop("depthwise", False)
    stride = kwargs.pop("stride", 1)
    dilation = kwargs.pop("dilation", 1)
    groups = in_channels if depthwise else kwargs.pop("groups", 1)
    padding = ((stride - 1) + dilation * (kernel_size - 1)) // 2
    return nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, dilation=dilation,
                     groups=groups, **kwargs)


def as_module() -> types.ModuleType:
    """A module object shaped like `timm` for the symbols the reference touches."""
    timm = types.ModuleType("timm")
    timm.create_model = create_model
    timm.layers = types.ModuleType("timm.layers")
    timm.layers.LayerNorm2d, timm.layers.LayerNorm = LayerNorm2d, LayerNorm
    timm.models = types.ModuleType("timm.models")
    timm.models.convnext = types.ModuleType("timm.models.convnext")
    timm.models.convnext.ConvNeXtStage = ConvNeXtStage
    timm.models.convnext.ConvNeXtBlock = ConvNeXtBlock
    timm.models.convnext._init_weights = _init_weights
    # VM/unet/fcmae.py:12-19
    timm.models.convnext.Downsample = Downsample
    timm.models.convnext.DropPath = DropPath
    timm.models.convnext.GlobalResponseNormMlp = GlobalResponseNormMlp
    timm.models.convnext.LayerNorm2d = LayerNorm2d
    timm.models.convnext.create_conv2d = create_conv2d
    timm.models.convnext.trunc_normal_ = nn.init.trunc_normal_
    timm.__oracle_restatement__ = True
    return timm
