"""ContrastiveEncoder (VM/contrastive/encoder.py:52-154) with the reference constructor / forward / state_dict surface.

timm `convnext_tiny` / `convnextv2_tiny` trunk with the 3-D-to-2-D `StemDepthtoChannels`, the pooled
LayerNorm head (fc = Identity) and the Linear-BN-ReLU-Linear-BN projection MLP.  CUDA tensors run through the
sm_100a kernels; CPU tensors run the same math in torch ops.
"""

from __future__ import annotations

from typing import Literal

import torch
from torch import Tensor, nn

from . import functional as F
from .components import LayerNorm2d, StemDepthtoChannels, _make_stages, init_convnext_weights
from .unext2 import resolve_compute_dtype


class _ConvNeXtHead(nn.Module):
    """timm NormMlpClassifierHead (hidden_size=None): global_pool -> norm -> flatten -> pre_logits -> drop -> fc."""

    def __init__(self, in_features: int, num_classes: int):
        super().__init__()
        self.global_pool = nn.AdaptiveAvgPool2d(1)
        self.norm = LayerNorm2d(in_features)
        self.flatten = nn.Flatten(1)
        self.pre_logits = nn.Identity()
        self.drop = nn.Dropout(0.0)
        self.fc = nn.Linear(in_features, num_classes) if num_classes > 0 else nn.Identity()

    def forward(self, x: Tensor) -> Tensor:
        return self.fc(self.drop(self.pre_logits(self.flatten(self.norm(self.global_pool(x))))))


class ConvNeXt(nn.Module):
    """`timm.create_model(backbone, features_only=False, num_classes=...)` for the ConvNeXt family."""

    def __init__(self, backbone: str, num_classes: int, drop_path_rate: float = 0.0, in_chans: int = 3):
        super().__init__()
        stages, dims = _make_stages(backbone, drop_path_rate)
        self.stem = nn.Sequential(nn.Conv2d(in_chans, dims[0], kernel_size=4, stride=4), LayerNorm2d(dims[0]))
        self.stages = nn.Sequential(*stages)
        self.num_features = dims[-1]
        self.norm_pre = nn.Identity()
        self.head = _ConvNeXtHead(dims[-1], num_classes)
        self.apply(init_convnext_weights)

    def forward(self, x: Tensor) -> Tensor:
        return self.head(self.norm_pre(self.stages(self.stem(x))))

    def forward_cl(self, x: Tensor) -> Tensor:
        if not isinstance(self.stem[0], nn.Identity):
            raise NotImplementedError("sm_100a path expects the 3-D stem in place of timm's stem conv")
        x = self.stem[1].forward_cl(x)
        for s in self.stages:
            x = s.forward_cl(x)
        emb = F.avgpool_ln(x, self.head.norm)
        if not isinstance(self.head.fc, nn.Identity):
            emb = F.LinearFn.apply(emb, self.head.fc.weight, self.head.fc.bias)
        return emb


class ContrastiveEncoder(nn.Module):
    """Contrastive encoder network using ConvNeXt backbones (reference: timm models) with a 3-D stem."""

    def __init__(
        self,
        backbone: Literal["convnext_tiny", "convnextv2_tiny", "resnet50"],
        in_channels: int,
        in_stack_depth: int,
        stem_kernel_size: tuple[int, int, int] = (5, 4, 4),
        stem_stride: tuple[int, int, int] = (5, 4, 4),
        embedding_dim: int = 768,
        projection_dim: int = 128,
        drop_path_rate: float = 0.0,
        pretrained: bool = False,
    ) -> None:
        super().__init__()
        self.backbone = backbone
        if "convnext" not in backbone:
            raise NotImplementedError(f"viscy_b200.ContrastiveEncoder provides the ConvNeXt backbones, not {backbone!r}")
        if pretrained:
            raise RuntimeError("pretrained=True needs timm's weight hub; load a checkpoint with load_state_dict instead")
        encoder = ConvNeXt(backbone, num_classes=embedding_dim, drop_path_rate=drop_path_rate)
        in_channels_encoder = encoder.stem[0].out_channels
        # Remove the convolution layer of stem, but keep the layernorm.
        encoder.stem[0] = nn.Identity()
        projection = nn.Sequential(
            nn.Linear(encoder.num_features, embedding_dim),
            nn.BatchNorm1d(embedding_dim),
            nn.ReLU(inplace=True),
            nn.Linear(embedding_dim, projection_dim),
            nn.BatchNorm1d(projection_dim),
        )
        encoder.head.fc = nn.Identity()
        self.stem = StemDepthtoChannels(
            in_channels=in_channels,
            in_stack_depth=in_stack_depth,
            in_channels_encoder=in_channels_encoder,
            stem_kernel_size=tuple(stem_kernel_size),
            stem_stride=tuple(stem_stride),
        )
        self.encoder = encoder
        self.projection = projection
        self.compute_dtype: torch.dtype | None = None
        self._packs = None

    def _weight_packs(self, dt: torch.dtype) -> None:
        """One-launch refresh of the 16-bit operand copies of every ConvNeXt block's fc1 / fc2 / conv_dw weights (instead
        of one cast launch per use: 166 launches, 0.9 ms per two-view step)."""
        from .components import ConvNeXtBlock
        if self._packs is None or self._packs.dtype != dt or self._packs.stale():
            blocks = [m for m in self.modules() if isinstance(m, ConvNeXtBlock)]
            lin = [w for b in blocks for w in (b.mlp.fc1.weight, b.mlp.fc2.weight)]
            self._packs = F.ops.WeightPacks(lin, [b.conv_dw.weight for b in blocks], dt)
        self._packs.refresh()
        F.ops.ACTIVE_PACKS = self._packs

    def forward(self, x: Tensor) -> tuple[Tensor, Tensor]:
        """Return (embedding, projection)."""
        if x.is_cuda:
            return self._forward_sm100(x)
        x = self.stem(x)
        embedding = self.encoder(x)
        projections = self.projection(embedding)
        return (embedding, projections)

    def _forward_sm100(self, x: Tensor) -> tuple[Tensor, Tensor]:
        dt = resolve_compute_dtype(x, self.compute_dtype)
        self._weight_packs(dt)
        F.ops.STEP.begin(x.device, torch.is_grad_enabled())  # one zero-filled allocation for the step's accumulators
        with torch.autocast("cuda", enabled=False):
            f = self.stem.forward_cl(x, dt)
            emb = self.encoder.forward_cl(f)
            p = self.projection
            h = F.LinearFn.apply(emb, p[0].weight, p[0].bias)
            h = F.batchnorm_rows(h, p[1], relu=True)
            h = F.LinearFn.apply(h, p[3].weight, p[3].bias)
            proj = F.batchnorm_rows(h, p[4], relu=False)
        return (emb, proj)
