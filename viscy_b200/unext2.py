"""UNeXt2 (VM/unet/unext2.py:13-82) with the reference constructor / forward / state_dict surface.

CUDA tensors run through the hand-written sm_100a kernels (channels-last 16-bit activations, fp32 parameters);
CPU tensors run the same math in plain torch ops.
"""

from __future__ import annotations

from typing import Literal

import torch
from torch import Tensor, nn

from .components import ConvNeXtFeatures, PixelToVoxelHead, UNeXt2Decoder, UNeXt2Stem


def resolve_compute_dtype(x: Tensor, override: torch.dtype | None) -> torch.dtype:
    """16-bit arithmetic type of the sm_100a path: explicit override > autocast dtype > the input's own 16-bit dtype."""
    if override is not None:
        dt = override
    elif torch.is_autocast_enabled("cuda"):
        dt = torch.get_autocast_dtype("cuda")
    else:
        dt = x.dtype
    if dt not in (torch.bfloat16, torch.float16):
        raise NotImplementedError(
            "the sm_100a path computes in bf16/fp16 with fp32 accumulation: run under torch.autocast('cuda', "
            "torch.bfloat16 | torch.float16), pass 16-bit input, or set `model.compute_dtype`; "
            f"got input dtype {x.dtype} without autocast"
        )
    return dt


class UNeXt2(nn.Module):
    """UNeXt2 model composing a ConvNeXt encoder with custom stem, decoder, and head."""

    def __init__(
        self,
        in_channels: int = 1,
        out_channels: int = 1,
        in_stack_depth: int = 5,
        out_stack_depth: int = None,
        backbone: str = "convnextv2_tiny",
        pretrained: bool = False,
        stem_kernel_size: tuple[int, int, int] = (5, 4, 4),
        decoder_mode: Literal["deconv", "pixelshuffle"] = "pixelshuffle",
        decoder_conv_blocks: int = 2,
        decoder_norm_layer: str = "instance",
        decoder_upsample_pre_conv: bool = False,
        head_pool: bool = False,
        head_expansion_ratio: int = 4,
        drop_path_rate: float = 0.0,
    ) -> None:
        super().__init__()
        if in_stack_depth % stem_kernel_size[0] != 0:
            raise ValueError(
                f"Input stack depth {in_stack_depth} is not divisible by stem kernel depth {stem_kernel_size[0]}."
            )
        if out_stack_depth is None:
            out_stack_depth = in_stack_depth
        if pretrained:
            raise RuntimeError("pretrained=True needs timm's weight hub; load a checkpoint with load_state_dict instead")
        multi_scale_encoder = ConvNeXtFeatures(backbone, drop_path_rate=drop_path_rate)
        num_channels = multi_scale_encoder.feature_info.channels()
        # replace first convolution layer with a projection tokenizer
        multi_scale_encoder.stem_0 = nn.Identity()
        self.encoder_stages = multi_scale_encoder
        self.stem = UNeXt2Stem(in_channels, num_channels[0], tuple(stem_kernel_size), in_stack_depth)
        decoder_channels = num_channels
        decoder_channels.reverse()
        decoder_channels[-1] = (out_stack_depth + 2) * out_channels * 2**2 * head_expansion_ratio
        self.decoder = UNeXt2Decoder(
            decoder_channels,
            norm_name=decoder_norm_layer,
            mode=decoder_mode,
            conv_blocks=decoder_conv_blocks,
            strides=[2] * (len(num_channels) - 1) + [stem_kernel_size[-1]],
            upsample_pre_conv="default" if decoder_upsample_pre_conv else None,
        )
        self.head = PixelToVoxelHead(
            decoder_channels[-1],
            out_channels,
            out_stack_depth,
            head_expansion_ratio,
            pool=head_pool,
        )
        self.out_stack_depth = out_stack_depth
        self.compute_dtype: torch.dtype | None = None
        # sm_100a path: split the batch into this many chunks that run on concurrent CUDA streams.  Every norm in
        # this model is per-sample (LayerNorm, GRN, InstanceNorm), so chunking is exact; it keeps the SMs busy on
        # the small, latency-bound feature maps of the deep stages.
        self.batch_streams: int = 1
        self._chunk_streams: list[torch.cuda.Stream] = []
        self._packs = None

    @property
    def num_blocks(self) -> int:
        """2-times downscaling factor of the smallest feature map."""
        return 6

    def forward(self, x: Tensor) -> Tensor:
        """Forward pass through the UNeXt2 model."""
        if x.is_cuda:
            return self._forward_sm100(x)
        x = self.stem(x)
        x: list = self.encoder_stages(x)
        x.reverse()
        x = self.decoder(x)
        return self.head(x)

    def _forward_chunk(self, x: Tensor, dt: torch.dtype) -> Tensor:
        f = self.stem.forward_cl(x, dt)
        feats = self.encoder_stages.forward_cl(f)
        feats.reverse()
        f = self.decoder.forward_cl(feats)
        return self.head.forward_cl(f)

    def _weight_packs(self, dt: torch.dtype):
        """One-launch refresh of the 16-bit operand copies of every ConvNeXt block's fc1 / fc2 / conv_dw weights."""
        from . import ops
        from .components import ConvNeXtBlock
        if self._packs is None or self._packs.dtype != dt or self._packs.stale():
            blocks = [m for m in self.modules() if isinstance(m, ConvNeXtBlock)]
            lin = [w for b in blocks for w in (b.mlp.fc1.weight, b.mlp.fc2.weight)]
            self._packs = ops.WeightPacks(lin, [b.conv_dw.weight for b in blocks], dt)
        self._packs.refresh()
        ops.ACTIVE_PACKS = self._packs

    def _forward_sm100(self, x: Tensor) -> Tensor:
        dt = resolve_compute_dtype(x, self.compute_dtype)
        self._weight_packs(dt)
        from . import ops
        ops.STEP.begin(x.device, torch.is_grad_enabled())  # one zero-filled allocation for the step's accumulators
        n = self.batch_streams
        with torch.autocast("cuda", enabled=False):
            if n <= 1 or x.shape[0] < n or x.shape[0] % n:
                return self._forward_chunk(x, dt)
            main = torch.cuda.current_stream(x.device)
            while len(self._chunk_streams) < n:
                self._chunk_streams.append(torch.cuda.Stream(device=x.device))
            outs = []
            for st, xc in zip(self._chunk_streams, x.chunk(n)):
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    outs.append(self._forward_chunk(xc.contiguous(), dt))
            for st, o in zip(self._chunk_streams, outs):
                main.wait_stream(st)
                o.record_stream(main)
            return torch.cat(outs)
