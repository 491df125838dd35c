"""Prediction path of the virtual-staining engine (CY/engine.py:48-101, 432-501, 618-805;
VU/callbacks/prediction_writer.py:74-111): divisible padding, centre crop, rotation test-time augmentation, Z sliding windows
with linear feathering - the same functions and the `AugmentedPredictionVSUNet` surface, without the Lightning base class.

The model forward is the sm_100a path of whichever viscy_b200 model is wrapped.  On CUDA the per-window epilogue (centre crop +
cast + `_blend_in` into the output volume) is one fused kernel (csrc/predict_sm100.cu) instead of crop, three elementwise passes
and a slice assignment; padding, rot90 and the TTA reduction are torch tensor plumbing.
"""

from __future__ import annotations

import ctypes as C
from functools import partial
from typing import Callable, Literal

import numpy as np
import torch
import torch.nn.functional as TF
from torch import Tensor, nn

from . import _lib as L

_DT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}


class DivisiblePad:
    """monai.transforms.DivisiblePad(k, method="symmetric", mode="constant") as cytoland uses it on (B,C,Z,Y,X) batches
    (engine.py:48-53: k = (0, 0, f, f) or (0, f, f, f) over the four trailing dims; the batch dim plays monai's channel)."""

    def __init__(self, k):
        self.k = tuple(k)

    def widths(self, shape):
        """[(before, after)] for the trailing len(k) dims."""
        out = []
        for size, k in zip(shape[-len(self.k):], self.k):
            new = size if k <= 0 else -(-size // k) * k
            tot = new - size
            out.append((tot // 2, tot - tot // 2))
        return out

    def __call__(self, x: Tensor) -> Tensor:
        w = self.widths(x.shape)
        if not any(a or b for a, b in w):
            return x
        flat = [v for ab in reversed(w) for v in ab]
        return TF.pad(x, flat, mode="constant", value=0.0)


def _make_divisible_pad(model: nn.Module) -> DivisiblePad:
    """engine.py:48-53"""
    down_factor = 2**model.num_blocks
    if getattr(model, "downsamples_z", False):
        return DivisiblePad((0, down_factor, down_factor, down_factor))
    return DivisiblePad((0, 0, down_factor, down_factor))


def _identity(x: Tensor) -> Tensor:
    return x


def _center_crop_to_shape(tensor: Tensor, spatial_shape: tuple[int, ...]) -> Tensor:
    """engine.py:61-71"""
    slices = [slice(None)] * tensor.ndim
    start_dim = tensor.ndim - len(spatial_shape)
    for dim, size in enumerate(spatial_shape, start=start_dim):
        current = tensor.shape[dim]
        if current < size:
            raise ValueError(f"Cannot crop dimension {dim} from {current} to {size}")
        start = (current - size) // 2
        slices[dim] = slice(start, start + size)
    return tensor[tuple(slices)]


def rotation_tta_transforms(n: int = 4):
    """Forward / inverse k*90-degree rotations of the YX plane (engine.py:74-101)."""
    if n < 1:
        raise ValueError(f"n must be >= 1, got {n}")
    forward = [partial(torch.rot90, k=k, dims=(-2, -1)) for k in range(n)]
    inverse = [partial(torch.rot90, k=-k, dims=(-2, -1)) for k in range(n)]
    return forward, inverse


def _blend_factors(z_slice: slice) -> list[int]:
    depth = z_slice.stop - z_slice.start
    samples = min(z_slice.start + 1, depth)
    return [min(i + 1, samples) for i in reversed(range(depth))]


def _blend_in(old_stack, new_stack, z_slice: slice):
    """Linear feathering of a new Z window into an old stack (prediction_writer.py:74-111); torch 5-D or numpy 4-D."""
    if z_slice.start == 0:
        return new_stack
    factors = _blend_factors(z_slice)
    if isinstance(old_stack, torch.Tensor):
        factors = torch.tensor(factors, dtype=old_stack.dtype, device=old_stack.device).view(1, 1, -1, 1, 1)
    else:
        factors = np.array(factors)[np.newaxis, :, np.newaxis, np.newaxis]
    return old_stack * (factors - 1) / factors + new_stack / factors


def blend_window_(out: Tensor, pred: Tensor, start: int) -> None:
    """out[:, :, start:start+d] = _blend_in(out[:, :, start:start+d], centre_crop(pred), slice(start, start+d)), in place.
    CUDA: one fused kernel; CPU: the reference ops."""
    d = pred.shape[2]
    z = slice(start, start + d)
    if not out.is_cuda:
        out[:, :, z] = _blend_in(out[:, :, z], _center_crop_to_shape(pred, out.shape[-2:]), z)
        return
    if out.dtype not in _DT or pred.dtype not in _DT:
        raise NotImplementedError(f"sm_100a blend: bf16 / fp16 / fp32 volumes, got {out.dtype} / {pred.dtype}")
    if not out.is_contiguous():
        raise ValueError("the output volume must be contiguous")
    pred = pred.contiguous()
    B, Cc, Z, H, W = out.shape
    if pred.shape[:2] != out.shape[:2]:
        raise ValueError(f"prediction {tuple(pred.shape)} does not match the output volume {tuple(out.shape)}")
    Hs, Ws = pred.shape[-2:]
    if Hs < H or Ws < W:
        raise ValueError(f"Cannot crop ({Hs}, {Ws}) to ({H}, {W})")
    L.check(L.lib().vb200_blend_window(L.ptr(out), L.ptr(pred), _DT[out.dtype], _DT[pred.dtype], C.c_int64(B * Cc), Z, H, W, d,
                                       Hs, Ws, start, (Hs - H) // 2, (Ws - W) // 2, L.stream_ptr()), "vb200_blend_window")


class AugmentedPredictionVSUNet(nn.Module):
    """Apply test-time augmentation and sliding-window inference around a virtual-staining model
    (CY/engine.py:566-805: same constructor, `with_rotation_tta`, `forward`, `predict_step`, `predict_sliding_windows`)."""

    def __init__(self, model: nn.Module, forward_transforms: list[Callable[[Tensor], Tensor]] | None = None,
                 inverse_transforms: list[Callable[[Tensor], Tensor]] | None = None,
                 reduction: Literal["mean", "median"] = "mean") -> None:
        super().__init__()
        self._predict_pad = _make_divisible_pad(model)
        self.model = model
        self._forward_transforms = forward_transforms or [_identity]
        self._inverse_transforms = inverse_transforms or [_identity]
        self._reduction = reduction

    @classmethod
    def with_rotation_tta(cls, model: nn.Module, n_rotations: int = 4,
                          reduction: Literal["mean", "median"] = "median") -> "AugmentedPredictionVSUNet":
        forward_transforms, inverse_transforms = rotation_tta_transforms(n_rotations)
        return cls(model=model, forward_transforms=forward_transforms, inverse_transforms=inverse_transforms,
                   reduction=reduction)

    def forward(self, x: Tensor) -> Tensor:
        return self.model(x)

    def setup(self, stage: str) -> None:
        if stage != "predict":
            raise NotImplementedError(f"Only the 'predict' stage is supported by {type(self)}")

    def _reduce_predictions(self, preds: list[Tensor]) -> Tensor:
        prediction = torch.stack(preds, dim=0)
        if self._reduction == "mean":
            prediction = prediction.mean(dim=0)
        elif self._reduction == "median":
            prediction = prediction.median(dim=0).values
        return prediction

    def _predict_with_tta(self, source: Tensor, crop: bool = True) -> Tensor:
        """crop=False (single transform only) leaves the centre crop to the fused blend kernel."""
        preds = []
        single = len(self._forward_transforms) == 1
        for fwd_t, inv_t in zip(self._forward_transforms, self._inverse_transforms):
            aug_source = fwd_t(source)
            aug_shape = aug_source.shape[2:]  # the prediction lives in the augmented frame until inv_t undoes it
            pred = self.forward(self._predict_pad(aug_source))
            if crop or not single:
                pred = _center_crop_to_shape(pred, aug_shape)
            preds.append(inv_t(pred))
        if len(preds) == 1:
            return preds[0]
        return self._reduce_predictions(preds)

    def predict_step(self, batch, batch_idx: int = 0, dataloader_idx: int = 0) -> Tensor:
        return self._predict_with_tta(batch["source"])

    def predict_sliding_windows(self, x: Tensor, out_channel: int = 2, step: int = 1) -> Tensor:
        """Sliding windows along Z with linear feathering (engine.py:757-805)."""
        if x.ndim != 5:
            raise ValueError(f"Expected input with 5 dimensions (B, C, Z, Y, X), got {x.shape}")
        batch_size, _, depth, height, width = x.shape
        in_stack_depth = getattr(self.model, "out_stack_depth", None)
        if in_stack_depth is None:
            raise ValueError(
                f"Model {type(self.model).__name__} does not support sliding window "
                "prediction (missing out_stack_depth attribute)."
            )
        if in_stack_depth > depth:
            raise ValueError(f"in_stack_depth {in_stack_depth} > input depth {depth}")
        out_tensor = x.new_zeros((batch_size, out_channel, depth, height, width))
        plain = len(self._forward_transforms) == 1 and self._inverse_transforms[0] is _identity
        for start in range(0, depth - in_stack_depth + 1, step):
            pred = self._predict_with_tta(x[:, :, start:start + in_stack_depth], crop=not (plain and x.is_cuda))
            blend_window_(out_tensor, pred, start)
        return out_tensor
