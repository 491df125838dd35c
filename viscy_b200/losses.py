"""MixedLoss and the 2.5-D (multi-scale) SSIM it is built on (VU/losses/mixed_loss.py:13-69,
VU/evaluation/metrics.py:174-349): same signatures, defaults and error texts.

CUDA tensors: every pyramid level is ONE fused sm_100a kernel pass over the two volumes (five bf16 box filters, SSIM and
contrast-sensitivity maps, per-sample means, L1 / L2 sums, the (1,2,2) average pooling that feeds the next level and its
data range), and one pass for the gradient (csrc/ssim_sm100.cu).  The handful of [levels, B] scalar operations that combine
the levels stay torch ops.  CPU tensors run the reference math in plain torch ops.
"""

from __future__ import annotations

import ctypes as C
from math import prod
from typing import Sequence, Union
from warnings import warn

import torch
import torch.nn.functional as TF
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib as L

_DT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}


def _dt(t: torch.Tensor) -> int:
    if t.dtype not in _DT:
        raise NotImplementedError(f"sm_100a loss kernels take bf16 / fp16 / fp32 volumes, got {t.dtype}")
    return _DT[t.dtype]


def _call(name, *args):
    L.check(getattr(L.lib(), name)(*args, L.stream_ptr()), name)


class _LevelFn(Function):
    """One pyramid level.  -> (ssim [B], cs [B], l1, l2, pooled preds, pooled target, data range of the next level)."""

    @staticmethod
    def forward(ctx, x, y, data_range, kh, kw, flags):
        B, Cc, D, H, W = x.shape
        x, y = x.contiguous(), y.contiguous()
        dev = x.device
        acc = torch.zeros((B, 4), device=dev, dtype=torch.float32)
        need_grad = x.requires_grad and bool(flags & 1)
        Ho, Wo = H - kh + 1, W - kw + 1
        mu = torch.empty((5, B * Cc, Ho, Wo), device=dev, dtype=torch.float32) if need_grad else None
        xp = yp = nxt = None
        if flags & 4:
            xp = torch.empty((B, Cc, D, H // 2, W // 2), device=dev, dtype=x.dtype)
            yp = torch.empty((B, Cc, D, H // 2, W // 2), device=dev, dtype=y.dtype)
            nxt = torch.full((), float("-inf"), device=dev, dtype=torch.float32)
        _call("vb200_ssim25d_level_fwd", L.ptr(x), L.ptr(y), _dt(x), _dt(y), B, Cc, D, H, W, kh, kw, L.ptr(data_range),
              L.ptr(acc), L.ptr(mu), L.ptr(xp), L.ptr(yp), L.ptr(nxt), flags)
        ctx.save_for_backward(x, y, data_range, mu)
        ctx.meta = (kh, kw, flags)
        npix, ntot = float(Cc * Ho * Wo), float(B * Cc * D * H * W)
        ssim, cs = acc[:, 0] / npix, acc[:, 1] / npix
        l1, l2 = acc[:, 2].sum() / ntot, acc[:, 3].sum() / ntot
        if flags & 4:
            ctx.mark_non_differentiable(yp, nxt)
        else:
            xp = x.new_empty(0)
            yp = nxt = x.new_empty(0)
            ctx.mark_non_differentiable(xp, yp, nxt)
        return ssim, cs, l1, l2, xp, yp, nxt

    @staticmethod
    @once_differentiable
    def backward(ctx, g_ssim, g_cs, g_l1, g_l2, g_xp, _gy, _gn):
        x, y, data_range, mu = ctx.saved_tensors
        kh, kw, flags = ctx.meta
        B, Cc, D, H, W = x.shape
        dx = torch.empty_like(x)
        f32 = lambda t: None if t is None else t.contiguous().float()  # noqa: E731
        g_ssim, g_cs = (f32(g_ssim), f32(g_cs)) if (flags & 1) else (None, None)
        g_l1, g_l2 = (f32(g_l1), f32(g_l2)) if (flags & 2) else (None, None)
        g_xp = g_xp.contiguous().to(x.dtype) if (flags & 4) and g_xp is not None else None
        _call("vb200_ssim25d_level_bwd", L.ptr(x), L.ptr(y), _dt(x), _dt(y), B, Cc, D, H, W, kh, kw, L.ptr(data_range),
              L.ptr(mu), L.ptr(g_ssim), L.ptr(g_cs), L.ptr(g_l1), L.ptr(g_l2), L.ptr(g_xp), L.ptr(dx))
        return dx, None, None, None, None, None


def _device_max(t: torch.Tensor) -> torch.Tensor:
    out = torch.full((), float("-inf"), device=t.device, dtype=torch.float32)
    t = t.contiguous()
    _call("vb200_max_f", L.ptr(t), _dt(t), C.c_int64(t.numel()), L.ptr(out))
    return out


# ------------------------------------------------------------------------------------------------ reference math (CPU)
def _compute_ssim_and_cs_bf16(y_pred, y, kernel_size, data_range=1.0, k1=0.01, k2=0.03):
    """metrics.py:174-262 in plain torch ops."""
    if y.shape != y_pred.shape:
        raise ValueError(f"y_pred and y must have same shape, got {y_pred.shape} and {y.shape}.")
    nc = y_pred.size(1)
    kernel = (torch.ones((nc, 1, *kernel_size), device=y_pred.device, dtype=torch.float32)
              / float(prod(kernel_size))).to(torch.bfloat16)
    xf, yf = y_pred.float(), y.float()
    terms = (y_pred.to(torch.bfloat16), y.to(torch.bfloat16), (xf * xf).to(torch.bfloat16), (yf * yf).to(torch.bfloat16),
             (xf * yf).to(torch.bfloat16))
    mu_x, mu_y, mu_xx, mu_yy, mu_xy = (TF.conv3d(t, kernel, groups=nc).float() for t in terms)
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    sigma_x, sigma_y, sigma_xy = mu_xx - mu_x * mu_x, mu_yy - mu_y * mu_y, mu_xy - mu_x * mu_y
    cs = (2 * sigma_xy + c2) / (sigma_x + sigma_y + c2)
    return ((2 * mu_x * mu_y + c1) / (mu_x * mu_x + mu_y * mu_y + c1)) * cs, cs


def ssim_25d(preds: torch.Tensor, target: torch.Tensor, in_plane_window_size: tuple[int, int] = (11, 11),
             return_contrast_sensitivity: bool = False) -> Union[torch.Tensor, tuple[torch.Tensor, torch.Tensor]]:
    """SSIM for 2.5D volumes: uniform window, depth window = depth (metrics.py:265-304)."""
    if preds.ndim != 5:
        raise ValueError(f"Input shape must be (B, C, D, W, H), got input shape {preds.shape}")
    depth = preds.shape[2]
    if depth > 15:
        warn(f"Input depth {depth} is potentially too large for 2.5D SSIM.")
    if preds.is_cuda:
        if target.shape != preds.shape:
            raise ValueError(f"y_pred and y must have same shape, got {preds.shape} and {target.shape}.")
        ssim, cs, *_ = _LevelFn.apply(preds, target.detach(), _device_max(target), *in_plane_window_size, 1)
    else:
        ssim_img, cs_img = _compute_ssim_and_cs_bf16(preds, target, (depth, *in_plane_window_size), data_range=target.max())
        ssim, cs = ssim_img.view(ssim_img.shape[0], -1).mean(1), cs_img.view(cs_img.shape[0], -1).mean(1)
    return (ssim, cs) if return_contrast_sensitivity else ssim


_BETAS: dict = {}


def _betas(betas, device) -> torch.Tensor:
    """The level exponents as a [levels, 1] device tensor, uploaded once per (values, device): no host-to-device copy in the
    step, so the loss can be captured into a CUDA graph."""
    key = (tuple(float(b) for b in betas), str(device))
    t = _BETAS.get(key)
    if t is None:
        t = _BETAS[key] = torch.tensor(key[0], device=device).view(-1, 1)
    return t


def _combine(ssim_last, cs_list, clamp, betas, base_min=1e-4):
    """metrics.py:337-349: clamp, replace the last level's cs by its ssim, weight by the betas, product, batch mean."""
    if clamp:
        cs_list = [c.clamp(min=base_min) for c in cs_list]
        ssim_last = ssim_last.clamp(min=base_min)
    stack = torch.stack(cs_list[:-1] + [ssim_last])
    if not stack.is_cuda:
        return torch.prod(stack ** torch.tensor(betas, device=stack.device).view(-1, 1), axis=0).mean()
    # CUDA: an explicit product chain - torch.prod's backward inspects the input for zeros on the host (a sync that a CUDA
    # graph capture of the training step cannot contain)
    w = stack ** _betas(betas, stack.device)
    r = w[0]
    for i in range(1, w.shape[0]):
        r = r * w[i]
    return r.mean()


def _ms_ssim_cuda(preds, target, window, clamp, betas, l1l2: bool):
    """-> (ms-ssim, l1, l2): all pyramid levels on the fused kernels; l1 / l2 come with the first level's pass."""
    if preds.ndim != 5:
        raise ValueError(f"Input shape must be (B, C, D, W, H), got input shape {preds.shape}")
    if target.shape != preds.shape:
        raise ValueError(f"y_pred and y must have same shape, got {preds.shape} and {target.shape}.")
    if preds.shape[2] > 15:
        warn(f"Input depth {preds.shape[2]} is potentially too large for 2.5D SSIM.")
    kh, kw = window
    x, y, dr = preds, target.detach(), _device_max(target)
    cs_list, l1, l2, ssim = [], None, None, None
    n = len(betas)
    for lvl in range(n):
        if x.shape[-2] < kh or x.shape[-1] < kw:
            raise RuntimeError(f"level {lvl}: a {x.shape[-2]}x{x.shape[-1]} plane is smaller than the {kh}x{kw} SSIM window")
        flags = 1 | (2 if (l1l2 and lvl == 0) else 0) | (4 if lvl + 1 < n else 0)
        ssim, cs, a, b, xp, yp, nxt = _LevelFn.apply(x, y, dr, kh, kw, flags)
        if lvl == 0:
            l1, l2 = a, b
        cs_list.append(cs)
        x, y, dr = xp, yp, nxt
    return _combine(ssim, cs_list, clamp, betas), l1, l2


def ms_ssim_25d(preds: torch.Tensor, target: torch.Tensor, in_plane_window_size: tuple[int, int] = (11, 11),
                clamp: bool = False, betas: Sequence[float] = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)) -> torch.Tensor:
    """Multi-scale SSIM for 2.5D volumes; depth is not downsampled (metrics.py:307-349)."""
    if preds.is_cuda:
        return _ms_ssim_cuda(preds, target, tuple(in_plane_window_size), clamp, betas, False)[0]
    cs_list = []
    for _ in range(len(betas)):
        ssim, cs = ssim_25d(preds, target, in_plane_window_size, return_contrast_sensitivity=True)
        cs_list.append(cs)
        preds = TF.avg_pool3d(preds, (1, 2, 2))
        target = TF.avg_pool3d(target, (1, 2, 2))
    return _combine(ssim, cs_list, clamp, betas)


class MixedLoss(nn.Module):
    """Mixed reconstruction loss: l1_alpha * L1 + l2_alpha * L2 + ms_dssim_alpha * (1 - MS-SSIM) (mixed_loss.py:13-69)."""

    def __init__(self, l1_alpha: float = 0.5, l2_alpha: float = 0.0, ms_dssim_alpha: float = 0.5):
        super().__init__()
        if not any([l1_alpha, l2_alpha, ms_dssim_alpha]):
            raise ValueError("Loss term weights cannot be all zero!")
        self.l1_alpha = l1_alpha
        self.l2_alpha = l2_alpha
        self.ms_dssim_alpha = ms_dssim_alpha

    def forward(self, preds, target):
        if preds.is_cuda:
            return self._forward_sm100(preds, target)
        loss = 0
        if self.l1_alpha:
            loss += TF.l1_loss(preds, target) * self.l1_alpha
        if self.l2_alpha:
            loss += TF.mse_loss(preds, target) * self.l2_alpha
        if self.ms_dssim_alpha:
            loss += (1 - ms_ssim_25d(preds, target, clamp=True)) * self.ms_dssim_alpha
        return loss

    def _forward_sm100(self, preds, target):
        want_l = bool(self.l1_alpha or self.l2_alpha)
        with torch.autocast("cuda", enabled=False):
            if self.ms_dssim_alpha:
                ms, l1, l2 = _ms_ssim_cuda(preds, target, (11, 11), True,
                                           (0.0448, 0.2856, 0.3001, 0.2363, 0.1333), want_l)
            else:
                if preds.shape != target.shape:
                    raise ValueError(f"preds and target must have same shape, got {preds.shape} and {target.shape}.")
                x5 = preds.reshape(1, 1, 1, 1, -1) if preds.ndim != 5 else preds
                y5 = target.reshape(1, 1, 1, 1, -1) if target.ndim != 5 else target
                _, _, l1, l2, *_ = _LevelFn.apply(x5, y5.detach(), None, 1, 1, 2)
                ms = None
            loss = 0
            if self.l1_alpha:
                loss = loss + l1 * self.l1_alpha
            if self.l2_alpha:
                loss = loss + l2 * self.l2_alpha
            if ms is not None:
                loss = loss + (1 - ms) * self.ms_dssim_alpha
            return loss


class _MSEFn(Function):
    """sum or mean of (pred - target)^2 in one pass over the two tensors, gradient w.r.t. pred in a second one."""

    @staticmethod
    def forward(ctx, pred, target, mean):
        n = pred.numel()
        out = torch.zeros((), device=pred.device, dtype=torch.float32)
        _call("vb200_mse_sum", L.ptr(pred), L.ptr(target), _dt(pred), _dt(target), C.c_int64(n),
              C.c_float(1.0 / n if mean else 1.0), L.ptr(out))
        ctx.save_for_backward(pred, target)
        ctx.scale = 2.0 / n if mean else 2.0
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("MSELoss on the sm_100a path: gradient with respect to the prediction only")
        dpred = torch.empty_like(pred)
        g = g.to(torch.float32).contiguous()
        _call("vb200_mse_bwd", L.ptr(pred), L.ptr(target), _dt(pred), _dt(target), C.c_int64(pred.numel()), L.ptr(g),
              C.c_float(ctx.scale), L.ptr(dpred))
        return dpred, None, None


class MSELoss(nn.Module):
    """``torch.nn.MSELoss`` (the reference's default ``loss_function``, CY/engine.py:197) for the training step: on CUDA the
    16-bit prediction and the fp32 target are read once per pass (no fp32 copy of the prediction, no separate reduction),
    the result is the fp32 scalar autocast's ``mse_loss`` returns.  reduction: "mean" (default) | "sum".  CPU tensors run
    ``torch.nn.functional.mse_loss``."""

    def __init__(self, reduction: str = "mean"):
        super().__init__()
        if reduction not in ("mean", "sum"):
            raise ValueError(f"reduction {reduction!r}: 'mean' or 'sum' (use torch.nn.MSELoss for 'none')")
        self.reduction = reduction

    def forward(self, preds: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if not preds.is_cuda:
            return TF.mse_loss(preds, target, reduction=self.reduction)
        if preds.shape != target.shape:
            raise ValueError(f"preds {tuple(preds.shape)} and target {tuple(target.shape)} must have the same shape")
        preds, target = preds.contiguous(), target.contiguous()
        if preds.data_ptr() % 32:  # views at odd offsets: the kernels read 32-byte vectors
            preds = preds.clone()
        if target.data_ptr() % 32:
            target = target.clone()
        return _MSEFn.apply(preds, target, self.reduction == "mean")


class MaskedMSELoss(nn.Module):
    """Masked MSE loss for FCMAE pre-training (CY/engine.py:104-125): a handful of elementwise / reduction ops on the
    reconstruction, kept as torch ops (FcmaeUNet itself insists on its own class, engine.py:877-878; this one is for callers
    outside that engine)."""

    def forward(self, preds, original, mask):
        loss = TF.mse_loss(preds, original, reduction="none")
        return (loss.mean(2) * mask).sum() / mask.sum()
