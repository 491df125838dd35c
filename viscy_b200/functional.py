"""Autograd wiring of the sm_100a kernels: one `torch.autograd.Function` per fused block of the hot path.

All activations between these functions are channels-last 16-bit tensors ([B,H,W,C] / rows [M,C]);
parameters stay fp32 in the reference's (PyTorch-native) layout and are re-packed to 16-bit GEMM
operands on use.  Gradients of parameters come back fp32 in the parameter's own shape.
"""

from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib as L
from . import ops

LN_EPS = 1e-6  # timm LayerNorm / LayerNorm2d
PAD = ops.ONES_PAD  # columns behind l / g that carry the ones column of the bias-gradient trick

# Backward runs the weight-gradient kernels (wgrad GEMMs, bias column sums, depthwise wgrad) on a side stream so
# that on small feature maps they overlap the latency-bound dgrad chain; every block joins the side stream before
# it returns.  Inside a CUDA graph this becomes a fork/join of parallel branches.
_SIDE: dict[int, torch.cuda.Stream] = {}
OVERLAP_WGRAD = True
FUSE_GRN_SUMSQ = True  # GRN sum of squares in the fc1 + GELU'/GELU epilogue (False, or feature maps under 32 pixels: column pass)


def _side_stream(device: torch.device) -> torch.cuda.Stream:
    """The helper stream paired with the CURRENT stream (one per main stream, so batch-chunk streams stay independent)."""
    key = torch.cuda.current_stream(device).cuda_stream
    st = _SIDE.get(key)
    if st is None:
        st = _SIDE[key] = torch.cuda.Stream(device=device)
    return st


# --------------------------------------------------------------------------------------------------
def _wgrad_splits(n_out: int, k_in: int, pixels: int) -> int:
    """K-split for the MN-major wgrad GEMM so that ~2 waves of tiles exist."""
    tiles = -(-n_out // 128) * -(-k_in // 256)
    want = max(1, (2 * 148 + tiles - 1) // tiles)
    kb = max(1, pixels // 64)
    return max(1, min(want, kb // 4 if kb >= 4 else 1, 64))


def linear_fwd(a2d, w, bias, *, residual=None, act=L.ACT_NONE, epilogue=L.EPI_STORE, out2=None, colsq=None,
               rows_per_sample=0):
    """a2d [M,K] 16-bit, w fp32 [N,K,...] -> [M,N] 16-bit."""
    wp = ops.packed(w, a2d.dtype)
    return ops.gemm(a2d, wp, bias=bias, residual=residual, act=act, epilogue=epilogue, out2=out2, colsq=colsq,
                    rows_per_sample=rows_per_sample)


def linear_bwd(dout2d, a2d, w, *, need_da=True, need_db=True):
    """-> (da [M,K] 16-bit or None, dW fp32 shaped like w, db fp32 [N] or None)"""
    N = w.shape[0]
    K = w.numel() // N
    M = dout2d.shape[0]
    dw = ops.gemm(dout2d, a2d, mn_major=True, epilogue=L.EPI_F32, k_splits=_wgrad_splits(N, K, M))
    db = ops.colreduce(dout2d.view(1, M, N), 0).view(N) if (need_db and N % 8 == 0) else (ops.colsum(dout2d) if need_db else None)
    da = None
    if need_da:
        wt = ops.packed(w, dout2d.dtype, transpose=True)  # [K, N]
        da = ops.gemm(dout2d, wt)
    return da, dw.view(w.shape), db


def _dw_taps(w):
    """conv_dw.weight [C,1,7,7] fp32 -> (tap-major [49,C], flipped tap-major [49,C])"""
    return ops.dw_taps(w)


# --------------------------------------------------------------------------------------------------
class ConvNeXtBlockFn(Function):
    """timm ConvNeXtBlock (V2: GRN, no layer scale / V1: layer-scale gamma, no GRN), NHWC in / NHWC out.

    forward:  d = dwconv7(x)+b ; l = LN(d) ; h = l W1^T + b1 ; y = GRN(GELU(h)) | GELU(h) ;
              out = keep[n] * ((y W2^T + b2) [* gamma]) + x        (keep: stochastic-depth scale per sample, or None)

    Bias gradients without column-sum passes (`onescol`, C % 16 == 0): LayerNorm writes l as [M, C + PAD] with a ones
    column and plants the same column behind the GELU output g [M, C4 + PAD]; the weight-gradient GEMMs dh^T [l | 1] and
    dout^T [g | 1] then carry d fc1.bias / d fc2.bias as their last column.
    """

    @staticmethod
    def forward(ctx, x, dw_w, dw_b, ln_w, ln_b, fc1_w, fc1_b, fc2_w, fc2_b, grn_w, grn_b, gamma, keep, eps=LN_EPS,
                dw=True):
        B, H, W, C = x.shape
        M = B * H * W
        ctx.dw = dw
        if dw:
            wt, wt_flip = _dw_taps(dw_w)
            ctx.wt_flip = wt_flip
            d = ops.dwconv7(x, wt, dw_b)
            res = x.view(M, C)
        else:
            # rows already behind the depthwise conv (FCMAE sparse path: x = the gathered unmasked rows [B, L, 1, C]);
            # the shortcut is added by the caller's scatter
            d, res = x, None
        C4 = fc1_w.shape[0]
        use_grn = grn_w is not None
        R = H * W
        fused = use_grn and R % 128 == 0 and B <= 16
        onescol = (fused or not use_grn) and C % 16 == 0 and C <= 2048 and C4 % 8 == 0
        if onescol:
            gbuf = torch.empty((M, C4 + PAD), device=x.device, dtype=x.dtype)  # [g | 1 0 ... 0]
            lbuf, mean, rstd = ops.layernorm_fwd(d, ln_w, ln_b, eps, ones=True, ones2=gbuf, ones2_col=C4)
            l2, y2 = lbuf[:, :C], gbuf[:, :C4]
        else:
            l, mean, rstd = ops.layernorm_fwd(d, ln_w, ln_b, eps)
            lbuf = l2 = l.view(M, C)
            gbuf = y2 = None
        if fused:
            # gp = gelu'(u) and g = gelu(u) from the fc1 epilogue; GRN scale folded into per-sample fc2 weights
            if FUSE_GRN_SUMSQ and R % 32 == 0:
                # the GRN statistic sum_rows g^2 accumulates in the fc1 epilogue (no pass over the hidden tensor)
                sumsq = ops.zeros((B, C4), x.device)
                h, y2 = linear_fwd(l2, fc1_w, fc1_b, epilogue=L.EPI_GELU_GP, out2=y2, colsq=sumsq, rows_per_sample=R)
                if not onescol:
                    gbuf = y2
            else:
                h, y2 = linear_fwd(l2, fc1_w, fc1_b, epilogue=L.EPI_GELU_GP, out2=y2)
                if onescol:
                    sumsq = ops.colreduce(gbuf.view(B, R, C4 + PAD), 1, width=C4)
                else:
                    sumsq = ops.colreduce(y2.view(B, R, C4), 1)
                    gbuf = y2
            w2 = fc2_w.detach().reshape(C, C4)
            s, w2s, b2e = ops.grn_prepare(sumsq, grn_w.detach(), grn_b.detach(), w2, fc2_b.detach(), x.dtype)
            out = ops.gemm(y2, w2s, bias=b2e, residual=res, b_batch_rows=R, rvec=keep, rvec_rows=R)
            ctx.save_for_backward(x, d, mean, rstd, lbuf, h, gbuf, sumsq, s, dw_w, ln_w, fc1_w, fc2_w, fc2_b, grn_w, grn_b, keep)
            ctx.use_grn, ctx.fused, ctx.onescol = True, True, onescol
            return out.view(B, H, W, C)
        if use_grn:
            h = linear_fwd(l2, fc1_w, fc1_b)
            y, sumsq, s = ops.gelu_grn_fwd(h.view(B, H * W, C4), grn_w, grn_b)
            gbuf = y2 = y.view(M, C4)
        else:
            h, y2 = linear_fwd(l2, fc1_w, fc1_b, epilogue=L.EPI_GELU_GP, out2=y2)  # h := gelu'(u)
            if not onescol:
                gbuf = y2
            sumsq = s = None
        # ConvNeXt-V1 layer scale: out = gamma * (y W2^T + b2) + x, gamma applied in fp32 in the epilogue
        out = ops.gemm(y2, ops.packed(fc2_w, x.dtype).view(C, C4), bias=fc2_b.detach(),
                       svec=None if gamma is None else gamma.detach(), residual=res, rvec=keep, rvec_rows=R)
        ctx.fused, ctx.onescol = False, onescol
        ctx.save_for_backward(x, d, mean, rstd, lbuf, h, gbuf, sumsq, s, dw_w, ln_w, fc1_w, fc2_w, fc2_b, grn_w, gamma, keep)
        ctx.use_grn = use_grn
        return out.view(B, H, W, C)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, d, mean, rstd, lbuf, h, gbuf, sumsq, s, dw_w, ln_w, fc1_w, fc2_w, fc2_b, grn_w, gamma, keep = ctx.saved_tensors
        B, H, W, C = x.shape
        M = B * H * W
        R = H * W
        C4 = fc1_w.shape[0]
        onescol = ctx.onescol
        dout = dout.contiguous()
        do2 = dout.view(M, C)
        # stochastic depth: the residual branch sees keep[n] * dout, the shortcut sees dout
        dob = do2 if keep is None else ops.scale_rows(dout, keep).view(M, C)
        l2 = lbuf[:, :C] if onescol else lbuf
        y2 = gbuf[:, :C4] if onescol else gbuf
        ar = ops.Arena(x.device, [2 * C, (B + 2) * C4, C4, 2 * C, 50 * C] + ([C4 * C, PAD * C4] if onescol else []))
        if ctx.fused:
            grn_b = gamma  # the 16th saved tensor is grn.bias on the fused path (V2 blocks carry no layer scale)
            gamma = None
            w2 = fc2_w.reshape(C, C4)
            # per-sample wgrad partials P[n] = dout_n^T [g_n | 1]: the ones column carries the fc2 bias gradient
            db2 = None if onescol else ops.colreduce(dob.view(1, M, C), 0, ar).view(C)
            P = ops.gemm(dob, gbuf, mn_major=True, epilogue=L.EPI_F32, k_splits=B, split_slabs=True)
            dw2, S1, dgb, db2 = ops.grn_wgrad_finish(P, w2, s, grn_b, db2, ar)
            t = torch.empty_like(S1)
            dgw = ar.take(C4)
            ops._call("vb200_grn_coef_bwd", ops._p(sumsq), ops._p(S1), ops._p(grn_w), ops._p(t), ops._p(dgw), B, C4,
                      ops.C.c_float(1e-6))
            w2t = ops.packed(fc2_w, x.dtype, transpose=True)  # [C4, C]
            dh = ops.gemm(dob, w2t, epilogue=L.EPI_DGELU_GRN, aux=y2, aux2=h, tvec=t, svec=s, rows_per_sample=R)
            dw2 = dw2.view(fc2_w.shape)
            dgamma = None
        else:
            w2 = fc2_w.reshape(C, C4)
            if ctx.use_grn:
                dy, dw2, db2 = linear_bwd(dob, y2, w2)
                dw2, dgamma = dw2.view(fc2_w.shape), None
                dh, dgw, dgb, db1 = ops.gelu_grn_bwd(h.view(B, H * W, C4), dy.view(B, H * W, C4), sumsq, s, grn_w)
            else:
                # ConvNeXt-V1: GELU backward rides in the fc2-dgrad epilogue (h holds gelu'(u))
                if onescol:
                    Gx = ops.gemm(dob, gbuf, mn_major=True, epilogue=L.EPI_F32, k_splits=_wgrad_splits(C, C4 + PAD, M))
                    G, db_raw = Gx[:, :C4], Gx[:, C4]  # G = dout^T y (without the layer scale), its ones column
                else:
                    _, G, db_raw = linear_bwd(dob, y2, w2, need_da=False)
                if gamma is not None:
                    # out = gamma * (y W2^T + b2): weights carry gamma / max|gamma| (16-bit safe), the epilogue the rest;
                    # dW2, db2, dgamma, the scaled transposed 16-bit operand and sv from one kernel
                    w2t, sv, dgamma, dw2, db2 = ops.layerscale_bwd(G, w2.contiguous(), gamma, fc2_b, db_raw.contiguous(),
                                                                   x.dtype, ar)
                    dw2 = dw2.view(fc2_w.shape)
                else:
                    sv, dgamma, dw2, db2 = None, None, G.reshape(fc2_w.shape), db_raw.contiguous()
                    w2t = ops.packed(fc2_w, x.dtype, transpose=True)
                dh = ops.gemm(dob, w2t, epilogue=L.EPI_DGELU_GRN, aux2=h, svec=sv,
                              rows_per_sample=128 * (-(-M // 128)) if sv is not None else 0)
                if not onescol:
                    db1 = ops.colreduce(dh.view(1, M, C4), 0, ar).view(C4)
                dgw = dgb = None
        dh2 = dh.view(M, C4)
        main = torch.cuda.current_stream()
        side = _side_stream(x.device) if OVERLAP_WGRAD else main
        if side is not main:
            side.wait_stream(main)
        with torch.cuda.stream(side):
            if onescol:
                # dW1 = dh^T [l | 1]: the weight gradient to its own buffer, the ones column (= d fc1.bias) to db1x
                dw1, db1x = ar.take(C4, C), ar.take(C4, PAD)
                ops.gemm(dh2, lbuf, mn_major=True, epilogue=L.EPI_F32, k_splits=_wgrad_splits(C4, C + PAD, M),
                         out=dw1, out2=db1x, n_split=C, accumulate=True)
                db1 = db1x[:, 0].contiguous()
                dw1 = dw1.view(fc1_w.shape)
            else:
                if ctx.fused:
                    db1 = ops.colreduce(dh.view(1, M, C4), 0, ar).view(C4)
                dw1 = ops.gemm(dh2, l2, mn_major=True, epilogue=L.EPI_F32,
                               k_splits=_wgrad_splits(C4, C, M)).view(fc1_w.shape)
        dl = ops.gemm(dh2, ops.packed(fc1_w, x.dtype, transpose=True))
        dd, dlnw, dlnb = ops.layernorm_bwd(dl.view(B, H, W, C), d, mean, rstd, ln_w, ar)
        if not ctx.dw:
            if side is not main:
                main.wait_stream(side)
            return dd, None, None, dlnw, dlnb, dw1, db1, dw2, db2, dgw, dgb, dgamma, None, None, None
        if side is not main:
            side.wait_stream(main)
        with torch.cuda.stream(side):
            dwt, ddb = ops.dwconv7_wgrad(x, dd, arena=ar)
        dx = ops.dwconv7(dd, ctx.wt_flip, None, add=dout)
        if side is not main:
            main.wait_stream(side)
        ddw = dwt.t().reshape(dw_w.shape)
        return dx, ddw, ddb, dlnw, dlnb, dw1, db1, dw2, db2, dgw, dgb, dgamma, None, None, None


def drop_path_scale(x: torch.Tensor, drop_prob: float, training: bool):
    """timm DropPath (scale_by_keep=True): per-sample Bernoulli(keep) / keep as an fp32 [B] vector, or None when the
    layer is the identity (VM/contrastive/encoder.py:79-90, VM/unet/unext2.py:40-46)."""
    if drop_prob == 0.0 or not training:
        return None
    keep = 1.0 - drop_prob
    mask = torch.empty((x.shape[0],), device=x.device, dtype=torch.float32).bernoulli_(keep)
    if keep > 0.0:
        mask.div_(keep)
    return mask


def convnext_block(x, blk, keep=None) -> torch.Tensor:
    """blk: module with conv_dw, norm, mlp.fc1, mlp.fc2, optional mlp.grn, optional gamma.  keep: stochastic-depth
    scale per sample (fp32 [B]) or None."""
    grn = getattr(blk.mlp, "grn", None)
    return ConvNeXtBlockFn.apply(
        x, blk.conv_dw.weight, blk.conv_dw.bias, blk.norm.weight, blk.norm.bias,
        blk.mlp.fc1.weight, blk.mlp.fc1.bias, blk.mlp.fc2.weight, blk.mlp.fc2.bias,
        None if grn is None else grn.weight, None if grn is None else grn.bias, getattr(blk, "gamma", None), keep,
    )


# --------------------------------------------------------------------------------------------------
# FCMAE (VM/unet/fcmae.py): dense blocks are ConvNeXtBlockFn with nn.LayerNorm's eps; the sparse (masked) path gathers
# the unmasked rows behind the depthwise conv, runs LN / MLP / GRN on them alone and scatters back onto the shortcut.
class DwConv7Fn(Function):
    """Depthwise 7x7 (padding 3) on NHWC rows."""

    @staticmethod
    def forward(ctx, x, dw_w, dw_b):
        wt, wt_flip = _dw_taps(dw_w)
        ctx.save_for_backward(x, dw_w)
        ctx.wt_flip = wt_flip
        return ops.dwconv7(x, wt, dw_b)

    @staticmethod
    @once_differentiable
    def backward(ctx, dd):
        x, dw_w = ctx.saved_tensors
        dd = dd.contiguous()
        dwt, ddb = ops.dwconv7_wgrad(x, dd)
        dx = ops.dwconv7(dd, ctx.wt_flip, None)
        return dx, dwt.t().reshape(dw_w.shape), ddb


class RowsSelectFn(Function):
    """dst[r] = (fwd_map[r] >= 0 ? src[fwd_map[r]] : 0) [+ base[r]];  bwd_map is the adjoint's map over the rows of src
    (gather <-> scatter: row index list <-> inverse index with -1 at the rows that are not selected)."""

    @staticmethod
    def forward(ctx, src, base, fwd_map, bwd_map):
        ctx.save_for_backward(bwd_map)
        ctx.has_base = base is not None
        return ops.rows_select(src, fwd_map, base=base)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        (bwd_map,) = ctx.saved_tensors
        dout = dout.contiguous()
        return ops.rows_select(dout, bwd_map), (dout if ctx.has_base else None), None, None


class MaskIndex:
    """Row maps of one boolean foreground mask [B, H, W] (True = kept; the same count L in every sample):
    idx [B*L] = kept rows in row-major order (masked_patchify's order), inv [B*H*W] = position in idx or -1,
    keep_self [B*H*W] = own row index or -1 (`x *= unmasked`)."""

    def __init__(self, unmasked: torch.Tensor, rows_per_sample: int):
        flat = unmasked.reshape(-1)
        M = flat.numel()
        n_keep = rows_per_sample * unmasked.shape[0]
        order = torch.argsort(flat.to(torch.int8), descending=True, stable=True)
        self.idx = order[:n_keep].to(torch.int32).contiguous()
        rows = torch.arange(M, device=flat.device, dtype=torch.int32)
        self.inv = torch.full((M,), -1, device=flat.device, dtype=torch.int32)
        self.inv[self.idx.long()] = torch.arange(n_keep, device=flat.device, dtype=torch.int32)
        self.keep_self = torch.where(flat, rows, torch.full_like(rows, -1)).contiguous()
        self.rows_per_sample = n_keep // unmasked.shape[0]


def fcmae_block(x, blk, mi: MaskIndex | None = None, keep=None):
    """MaskedConvNeXtV2Block (fcmae.py:144-227) on NHWC rows.  mi: row maps of the (upsampled) foreground mask or None."""
    mlp = blk.mlp
    args = (blk.layernorm.weight, blk.layernorm.bias, mlp.fc1.weight, mlp.fc1.bias, mlp.fc2.weight, mlp.fc2.bias,
            mlp.grn.weight, mlp.grn.bias, None, keep, blk.layernorm.eps)
    if mi is None:
        return ConvNeXtBlockFn.apply(x, blk.dwconv.weight, blk.dwconv.bias, *args, True)
    B, H, W, C = x.shape
    M = B * H * W
    xm = RowsSelectFn.apply(x.view(M, C), None, mi.keep_self, mi.keep_self)  # x *= unmasked (the shortcut is masked too)
    d = DwConv7Fn.apply(xm.view(B, H, W, C), blk.dwconv.weight, blk.dwconv.bias)
    rows = RowsSelectFn.apply(d.view(M, C), None, mi.idx, mi.inv)            # masked_patchify
    o = ConvNeXtBlockFn.apply(rows.view(B, mi.rows_per_sample, 1, C), None, None, *args, False)
    out = RowsSelectFn.apply(o.view(-1, C), xm, mi.inv, mi.idx)              # masked_unpatchify + shortcut
    return out.view(B, H, W, C)


class ShufflePoolFn(Function):
    """PixelToVoxelShuffleHead (heads.py:657-695): NHWC decoder rows -> (B, Cq, H r, W r)."""

    @staticmethod
    def forward(ctx, dec, r, pool):
        ctx.rp = (r, pool)
        return ops.shuffle_pool_fwd(dec, r, pool)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        r, pool = ctx.rp
        return ops.shuffle_pool_bwd(dout.contiguous(), r, pool), None, None


def shuffle_pool(dec, r, pool):
    return ShufflePoolFn.apply(dec, r, pool)


# --------------------------------------------------------------------------------------------------
class LayerNormFn(Function):
    """LayerNorm over the channel dim of channels-last rows (timm LayerNorm2d on NCHW == this on NHWC)."""

    @staticmethod
    def forward(ctx, x, w, b, eps):
        y, mean, rstd = ops.layernorm_fwd(x, w, b, eps)
        ctx.save_for_backward(x, mean, rstd, w)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, mean, rstd, w = ctx.saved_tensors
        dx, dw, db = ops.layernorm_bwd(dy.contiguous(), x, mean, rstd, w)
        return dx, dw, db, None


def layernorm(x, w, b, eps=LN_EPS):
    return LayerNormFn.apply(x, w, b, eps)


class LNConvFn(Function):
    """LayerNorm2d + Conv2d(k=1) or Conv2d(k=2,s=2): timm ConvNeXtStage.downsample (encoder k2s2, decoder k1)."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, w, b):
        B, H, W, C = x.shape
        k = w.shape[-1]
        l, mean, rstd = ops.layernorm_fwd(x, ln_w, ln_b, LN_EPS)
        if k == 2:
            a = ops.patchify2(l)
            wk = w.detach().permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()  # (kh,kw,c)
            Ho, Wo = H // 2, W // 2
        elif k == 1:
            a, wk, Ho, Wo = l, w.detach().reshape(w.shape[0], -1), H, W
        else:
            raise NotImplementedError(f"downsample kernel {k}")
        a2 = a.view(B * Ho * Wo, -1)
        Kc = a2.shape[1]
        if Kc % 8:  # TMA rows need a 16-byte pitch: zero-pad K (only odd channel counts, e.g. convnextv2_atto decoders)
            K8 = -(-Kc // 8) * 8
            a_p = a2.new_zeros((a2.shape[0], K8))
            a_p[:, :Kc] = a2
            a2 = a_p
            wk = torch.nn.functional.pad(wk, (0, K8 - Kc))
        out = ops.gemm(a2, ops.cast_pack(wk.contiguous(), x.dtype), bias=b)
        ctx.save_for_backward(x, mean, rstd, ln_w, a2, w)
        ctx.k = k
        return out.view(B, Ho, Wo, w.shape[0])

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, mean, rstd, ln_w, a2, w = ctx.saved_tensors
        B, H, W, C = x.shape
        k = ctx.k
        Co = w.shape[0]
        do2 = dout.contiguous().view(-1, Co)
        wk = w.permute(0, 2, 3, 1).reshape(Co, -1).contiguous() if k == 2 else w.reshape(Co, -1)
        Kc, K8 = wk.shape[1], a2.shape[1]
        if K8 != Kc:
            wk = torch.nn.functional.pad(wk, (0, K8 - Kc)).contiguous()
        da, dwk, db = linear_bwd(do2, a2, wk)
        if K8 != Kc:
            da, dwk = da[:, :Kc].contiguous(), dwk[:, :Kc]
        if k == 2:
            dl = ops.patchify2(da.view(B, H // 2, W // 2, 4 * C), inverse=True, shape=(B, H, W, C))
            dw = dwk.reshape(Co, 2, 2, C).permute(0, 3, 1, 2)
        else:
            dl = da.view(B, H, W, C)
            dw = dwk.reshape(w.shape)
        dx, dlnw, dlnb = ops.layernorm_bwd(dl, x, mean, rstd, ln_w)
        return dx, dlnw, dlnb, dw, db


def ln_conv(x, ln, conv):
    return LNConvFn.apply(x, ln.weight, ln.bias, conv.weight, conv.bias)


# --------------------------------------------------------------------------------------------------
class PixShufCatFn(Function):
    """monai SubpixelUpsample(scale 2, pre_conv=None) + torch.cat([up, skip], 1) on NHWC tensors."""

    @staticmethod
    def forward(ctx, prev, skip):
        ctx.Cp = prev.shape[-1]
        ctx.Cs = 0 if skip is None else skip.shape[-1]
        return ops.pixshuf_cat_fwd(prev, skip)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        dprev, dskip = ops.pixshuf_cat_bwd(dout.contiguous(), ctx.Cp, ctx.Cs)
        return dprev, dskip


def pixshuf_cat(prev, skip):
    return PixShufCatFn.apply(prev, skip)


# --------------------------------------------------------------------------------------------------
def _stem_full_weight(w, b, D, sD):
    """Conv3d(kernel (kD,kH,kW), stride (sD,kH,kW)) + (B,C,D',H,W)->(B,C*D',H,W) as one [C*D', Cin*D*kH*kW] matrix."""
    Co, Cin, kD, kH, kW = w.shape
    Dd = (D - kD) // sD + 1
    full = w.new_zeros((Co, Dd, Cin, D, kH, kW))
    for dd in range(Dd):
        full[:, dd, :, dd * sD: dd * sD + kD] = w
    bias = None if b is None else b.repeat_interleave(Dd)
    return full.view(Co * Dd, -1), bias, Dd


class StemFn(Function):
    """UNeXt2Stem / StemDepthtoChannels (VM/components/stems.py:33-50,117-134): NCDHW in, NHWC (C*D') out."""

    @staticmethod
    def forward(ctx, x, w, b, stride, dtype):
        Bn, Cin, D, H, W = x.shape
        Co, _, kD, kH, kW = w.shape
        sD, sH, sW = stride
        if (sH, sW) != (kH, kW):
            raise NotImplementedError("sm_100a stem needs in-plane stride == kernel (patchify)")
        if H % kH or W % kW:
            raise NotImplementedError("sm_100a stem needs H, W divisible by the stem kernel")
        A = ops.stem_patchify(x, kH, kW, dtype)
        wfull, bfull, Dd = _stem_full_weight(w.detach(), b.detach(), D, sD)
        out = ops.gemm(A, ops.cast_pack(wfull, dtype), bias=bfull.contiguous())
        ctx.save_for_backward(A, w)
        ctx.geom = (D, sD, Dd)
        return out.view(Bn, H // kH, W // kW, Co * Dd)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        A, w = ctx.saved_tensors
        D, sD, Dd = ctx.geom
        Co, Cin, kD, kH, kW = w.shape
        do2 = dout.contiguous().view(-1, Co * Dd)
        dfull = ops.gemm(do2, A, mn_major=True, epilogue=L.EPI_F32,
                         k_splits=_wgrad_splits(Co * Dd, A.shape[1], A.shape[0]))
        dfull = dfull.view(Co, Dd, Cin, D, kH, kW)
        dw = torch.zeros_like(w)
        for dd in range(Dd):
            dw += dfull[:, dd, :, dd * sD: dd * sD + kD]
        db = ops.colsum(do2).view(Co, Dd).sum(1)
        return None, dw, db, None, None


def stem(x, conv, dtype):
    if x.requires_grad:
        raise NotImplementedError("sm_100a stem does not produce an input gradient")
    return StemFn.apply(x, conv.weight, conv.bias, tuple(conv.stride), dtype)


# --------------------------------------------------------------------------------------------------
class HeadFn(Function):
    """PixelToVoxelHead (VM/components/heads.py:632-641): NHWC decoder features -> (B, Cout, D, 4h, 4w) NCDHW."""

    @staticmethod
    def forward(ctx, dec, conv_w, conv_b, alpha, w1, b1, out_depth, pool):
        B, h, w, Cd = dec.shape
        Dz = out_depth + 2
        Cm = Cd // 4
        Cc = Cm // Dz
        Cmid = conv_w.shape[0]
        Cu = -(-Cc // 8) * 8
        u = ops.head_shuffle_pool_fwd(dec, Dz, pool, Cu)
        implicit = Cu == 8 and Cmid in (16, 32)
        if implicit:
            # tcgen05 implicit GEMM: im2col folded into shifted 5-D TMA boxes (csrc/conv3d_sm100.cu)
            wp = ops.conv3d_pack_weights(conv_w, dec.dtype, 8, Cmid)
            z = ops.conv3d_k3(u, wp, conv_b, (0, 1, 1), Cmid, Cmid).view(-1, Cmid)
            col, geom = u, None
        else:
            geom = ops.conv3d_geom(tuple(u.shape), (3, 3, 3), (1, 1, 1), (0, 1, 1))
            col = ops.im2col3d(u, geom)
            wc = conv_w.detach().permute(0, 2, 3, 4, 1)  # [Cmid, 3,3,3, Cc]
            if Cu != Cc:
                wc = torch.nn.functional.pad(wc, (0, Cu - Cc))
            wc = wc.reshape(Cmid, -1).contiguous()
            z = ops.gemm(col, ops.cast_pack(wc, dec.dtype), bias=conv_b)
        H2, W2 = 2 * h, 2 * w
        R = out_depth * H2 * W2
        z3 = z.view(B, R, Cmid)
        mean, rstd = ops.instnorm_stats(z3, 1e-5)
        w1m = w1.detach().reshape(w1.shape[0], Cmid).contiguous()
        out = ops.head_tail_fwd(z3, mean, rstd, alpha.detach(), w1m, b1.detach(), out_depth, H2, W2)
        ctx.save_for_backward(col, z3, mean, rstd, alpha, w1, conv_w)
        ctx.meta = (geom, Cm, Cc, Cu, pool, out_depth, H2, W2, implicit)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        col, z3, mean, rstd, alpha, w1, conv_w = ctx.saved_tensors
        geom, Cm, Cc, Cu, pool, out_depth, H2, W2, implicit = ctx.meta
        B, R, Cmid = z3.shape
        w1m = w1.reshape(w1.shape[0], Cmid).contiguous()
        dz, dW1, db1, dalpha, dbz = ops.head_tail_bwd(z3, mean, rstd, alpha, w1m, dout.contiguous(), out_depth, H2, W2)
        if implicit:
            u = col
            dz5 = dz.view(B, out_depth, H2, W2, Cmid)
            main = torch.cuda.current_stream()
            side = _side_stream(dz.device) if OVERLAP_WGRAD else main
            if side is not main:
                side.wait_stream(main)
            with torch.cuda.stream(side):
                dconv_w = ops.conv3d_k3_wgrad(u, dz5, (0, 1, 1))[:, :Cc]
            if Cmid % 32 == 0 and ops.conv3d_igemm_supported(tuple(dz5.shape), 8, (3, 3, 3), (2, 1, 1)):
                # data gradient on the patch-form implicit GEMM (filter resident in shared memory)
                du = ops.conv3d_igemm(dz5, ops.conv_weight_rows(conv_w, 8, Cmid, dz.dtype, flipped=True), None,
                                      (3, 3, 3), (2, 1, 1))
            else:
                wpt = ops.conv3d_pack_weights(conv_w, dz.dtype, Cmid, 16, transpose_flip=True)
                du = ops.conv3d_k3(dz5, wpt, None, (2, 1, 1), 16, 8)
            if side is not main:
                main.wait_stream(side)
        else:
            dz2 = dz.view(B * R, Cmid)
            wc = conv_w.permute(0, 2, 3, 4, 1)
            if Cu != Cc:
                wc = torch.nn.functional.pad(wc, (0, Cu - Cc))
            wc = wc.reshape(Cmid, -1).contiguous()
            dcol, dwc, _ = linear_bwd(dz2, col, wc, need_db=False)
            du = ops.col2im3d(dcol, geom)
            dconv_w = dwc.view(Cmid, 3, 3, 3, Cu)[..., :Cc].permute(0, 4, 1, 2, 3)
        ddec = ops.head_shuffle_pool_bwd(du, Cm, pool)
        return ddec, dconv_w, dbz, dalpha.view(alpha.shape), dW1.view(w1.shape), db1, None, None


def pixel_to_voxel_head(dec, conv0, prelu, conv1, out_depth, pool):
    return HeadFn.apply(dec, conv0.weight, conv0.bias, prelu.weight, conv1.weight, conv1.bias, out_depth, pool)


# --------------------------------------------------------------------------------------------------
class AvgPoolLNFn(Function):
    """timm NormMlpClassifierHead up to `flatten`: global average pool over (H, W) then LayerNorm2d over C.
    NHWC [B,H,W,C] -> [B,C]."""

    @staticmethod
    def forward(ctx, x, ln_w, ln_b, eps):
        B, H, W, C = x.shape
        pooled = (ops.colreduce(x.view(B, H * W, C), 0) * (1.0 / (H * W))).to(x.dtype)
        y, mean, rstd = ops.layernorm_fwd(pooled, ln_w, ln_b, eps)
        ctx.save_for_backward(pooled, mean, rstd, ln_w)
        ctx.hw = (H, W)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        pooled, mean, rstd, ln_w = ctx.saved_tensors
        H, W = ctx.hw
        dp, dw, db = ops.layernorm_bwd(dy.contiguous().to(pooled.dtype), pooled, mean, rstd, ln_w)
        dx = ops.bcast_rows(dp, H * W, 1.0 / (H * W))
        return dx.view(pooled.shape[0], H, W, pooled.shape[1]), dw, db, None


def avgpool_ln(x, ln):
    return AvgPoolLNFn.apply(x, ln.weight, ln.bias, ln.eps)


class LinearFn(Function):
    """nn.Linear on 16-bit rows through the tcgen05 GEMM."""

    @staticmethod
    def forward(ctx, a, w, b):
        ctx.save_for_backward(a, w)
        return linear_fwd(a, w, b)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        a, w = ctx.saved_tensors
        da, dw, db = linear_bwd(dout.contiguous(), a, w)
        return da, dw, db


class BatchNormRowsFn(Function):
    """nn.BatchNorm1d (+ optional ReLU) over [B, C] rows; running statistics are updated by the caller."""

    @staticmethod
    def forward(ctx, x, weight, bias, run_mean, run_var, eps, training, relu):
        y, mean, rstd, var_unb = ops.bn_rows_fwd(x, weight, bias, run_mean, run_var, eps, training, relu)
        ctx.save_for_backward(x, y, weight, mean, rstd)
        ctx.flags = (training, relu)
        ctx.mark_non_differentiable(mean)
        if var_unb is None:
            var_unb = mean
        ctx.mark_non_differentiable(var_unb)
        return y, mean, var_unb

    @staticmethod
    @once_differentiable
    def backward(ctx, dy, _dm, _dv):
        x, y, weight, mean, rstd = ctx.saved_tensors
        training, relu = ctx.flags
        dx, dg, db = ops.bn_rows_bwd(dy.contiguous(), x, y, weight, mean, rstd, training, relu)
        return dx, dg, db, None, None, None, None, None


def batchnorm_rows(x, bn, relu=False):
    """Apply an nn.BatchNorm1d module (train: batch stats + running-stat update with its momentum; eval: running stats)."""
    training = bn.training or bn.running_mean is None
    y, mean, var_unb = BatchNormRowsFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, training, relu)
    if bn.training and bn.track_running_stats and bn.running_mean is not None:
        with torch.no_grad():
            bn.num_batches_tracked += 1
            mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            bn.running_mean.mul_(1 - mom).add_(mean, alpha=mom)
            bn.running_var.mul_(1 - mom).add_(var_unb, alpha=mom)
    return y


# --------------------------------------------------------------------------------------------------
# 3-D U-Net family (UNet3DBase / Unet3d): channels-last [N,D,H,W,C] 16-bit activations
def _pad8(c: int) -> int:
    return -(-c // 8) * 8


class ToChannelsLast3dFn(Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.c = x.shape[1]
        return ops.to_channels_last_3d(x, dtype, _pad8(x.shape[1]))

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        return ops.from_channels_last_3d(dy.contiguous(), ctx.c).float(), None


def to_channels_last_3d(x, dtype):
    if x.requires_grad:
        return ToChannelsLast3dFn.apply(x, dtype)
    return ops.to_channels_last_3d(x, dtype, _pad8(x.shape[1]))


class FromChannelsLast3dFn(Function):
    @staticmethod
    def forward(ctx, y, c):
        ctx.cpad = y.shape[-1]
        return ops.from_channels_last_3d(y, c)

    @staticmethod
    @once_differentiable
    def backward(ctx, dx):
        return ops.to_channels_last_3d(dx, dx.dtype, ctx.cpad), None


def from_channels_last_3d(y, c):
    return FromChannelsLast3dFn.apply(y, c)


def _conv_weight_rows(w, cin_pad, cout_pad):
    """nn.Conv3d weight [Co,Ci,kd,kh,kw] -> fp32 [cout_pad, (kd,kh,kw,cin_pad)] (zero padded)"""
    Co, Ci = w.shape[:2]
    wk = w.permute(0, 2, 3, 4, 1)
    if cin_pad != Ci or cout_pad != Co:
        wk = torch.nn.functional.pad(wk, (0, cin_pad - Ci, 0, 0, 0, 0, 0, 0, 0, cout_pad - Co))
    return wk.reshape(cout_pad, -1).contiguous()


def _conv_weight_rows_flipped(w, cin_pad, cout_pad):
    """nn.Conv3d weight [Co,Ci,kd,kh,kw] -> fp32 [cin_pad, (kd,kh,kw flipped, cout_pad)]: the filter of the data gradient"""
    Co, Ci = w.shape[:2]
    wk = w.flip(2, 3, 4).permute(1, 2, 3, 4, 0)
    if cin_pad != Ci or cout_pad != Co:
        wk = torch.nn.functional.pad(wk, (0, cout_pad - Co, 0, 0, 0, 0, 0, 0, 0, cin_pad - Ci))
    return wk.reshape(cin_pad, -1).contiguous()


def _small_k3(ks, stride, cin_rows, cout_rows) -> bool:
    """3x3x3 stride-1 conv with 8 / 16 input channels and <= 32 output channels (below the 32-channel granule of the
    implicit GEMM): served by the line kernel of conv3d_sm100.cu."""
    return ks == (3, 3, 3) and tuple(stride) == (1, 1, 1) and cin_rows in (8, 16) and cout_rows <= 32


class Conv3dFn(Function):
    """nn.Conv3d (groups=1) on channels-last rows.  Convs whose geometry tiles into TMA boxes run as implicit GEMMs
    (`ops.conv3d_igemm*`: no patch matrix in HBM): forward and weight gradient at any stride, data gradient at stride 1.
    Anything else (odd extents, < 32 channels) goes through im2col3d + the same tcgen05 GEMM, the patch matrix rebuilt in
    backward; the data gradient of a strided conv is the GEMM + col2im3d scatter."""

    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        N, D, H, W, Cp = x.shape
        Co = w.shape[0]
        Cop = _pad8(Co)
        if Cp != _pad8(w.shape[1]):
            raise NotImplementedError(
                f"sm_100a conv3d: rows carry {Cp} channels but the filter expects {w.shape[1]} (padded {_pad8(w.shape[1])}); "
                "channel-padded operands must be concatenated with cat_cl(a, b, ca, cb)")
        ks = tuple(w.shape[2:])
        stride, padding = tuple(stride), tuple(padding)
        geom = ops.conv3d_geom((N, D, H, W, Cp), ks, stride, padding)
        pointwise = ks == (1, 1, 1) and stride == (1, 1, 1) and padding == (0, 0, 0)
        igemm = not pointwise and ops.conv3d_igemm_supported((N, D, H, W, Cp), Cop, ks, padding, stride=stride)
        w16 = ops.conv_weight_rows(w, Cp, Cop, x.dtype) if w.numel() // (Co * w.shape[1]) <= 27 else \
            ops.cast_pack(_conv_weight_rows(w.detach(), Cp, Cop), x.dtype)
        bias = None
        if b is not None:
            bias = b.detach() if Cop == Co else torch.nn.functional.pad(b.detach(), (0, Cop - Co))
        if igemm:
            out = ops.conv3d_igemm(x, w16, bias, ks, padding, stride=stride)
        elif _small_k3(ks, stride, Cp, Cop):
            # model-boundary convs (3 -> 32 channels): the 130-voxel-line kernel, taps as views of resident lines
            cpad = 16 if Cop <= 16 else 32
            out = ops.conv3d_k3(x, ops.conv3d_pack_weights(w, x.dtype, Cp, cpad), bias, padding, cpad, Cop)
        else:
            col = x.view(-1, Cp) if pointwise else ops.im2col3d(x, geom)
            out = ops.gemm(col, w16, bias=bias)
        ctx.save_for_backward(x, w)
        ctx.meta = (geom, pointwise, stride, Cop, b is not None, padding)
        return out.view(N, geom[14], geom[15], geom[16], Cop)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        geom, pointwise, stride, Cop, has_bias, padding = ctx.meta
        N, D, H, W, Cp = x.shape
        Co, Ci = w.shape[:2]
        ks = tuple(w.shape[2:])
        dout = dout.contiguous()
        do2 = dout.view(-1, Cop)
        need_dx = ctx.needs_input_grad[0]
        unit = stride == (1, 1, 1) and not pointwise
        dx = dwk = None
        # data gradient at stride 1: the conv of dout with the flipped, transposed filter and padding k-1-p
        bpad = tuple(k - 1 - p for k, p in zip(ks, padding))
        if need_dx and unit and ops.conv3d_igemm_supported(tuple(dout.shape), Cp, ks, bpad):
            dx = ops.conv3d_igemm(dout, ops.conv_weight_rows(w, Cp, Cop, x.dtype, flipped=True), None, ks, bpad)
            need_dx = False
        if unit and Cp <= 128 and ops.conv3d_wgrad_kh3_supported((N, D, H, W, Cp), Cop, ks, padding):
            dwk = ops.conv3d_wgrad_kh3(x, dout, ks, padding)  # few channels: patch form, kh taps share one haloed box
        elif not pointwise and ops.conv3d_igemm_supported((N, D, H, W, Cp), Cop, ks, padding, wgrad=True, stride=stride):
            dwk = ops.conv3d_igemm_wgrad(x, dout, ks, padding, stride=stride)
        if need_dx and _small_k3(ks, stride, Cop, Cp):
            # few-channel data gradient (e.g. the 32 -> 3 output conv): line kernel with the flipped, transposed filter
            cpad = 16 if Cp <= 16 else 32
            wpt = ops.conv3d_pack_weights(w, x.dtype, Cop, cpad, transpose_flip=True)
            dx = ops.conv3d_k3(dout, wpt, None, bpad, cpad, Cp)
            need_dx = False
        if need_dx and not unit and not pointwise:
            # strided conv: dx is the transposed conv of dout = one stride-1 sub-convolution per parity class of dx
            rows = ops.cast_pack(_convT_weight_rows_fwd(w, Cop, Cp), x.dtype)
            dx = conv_transpose3d_igemm(dout, rows, None, ks, stride, padding, (D, H, W), Cp)
            need_dx = dx is None
        if need_dx:
            wt = ops.cast_pack(_conv_weight_rows(w, Cp, Cop), x.dtype, transpose=True)  # [K, Cop]
            dcol = ops.gemm(do2, wt)
            dx = dcol.view(x.shape) if pointwise else ops.col2im3d(dcol, geom)
        if dwk is None:
            col = x.view(-1, Cp) if pointwise else ops.im2col3d(x, geom)
            dwk = ops.gemm(do2, col, mn_major=True, epilogue=L.EPI_F32, k_splits=_wgrad_splits(Cop, col.shape[1], col.shape[0]))
        db = ops.colreduce(do2.view(1, -1, Cop), 0).view(Cop)[:Co] if has_bias else None
        dw = dwk.view(Cop, *ks, Cp)[:Co, ..., :Ci].permute(0, 4, 1, 2, 3)
        return dx, dw, db, None, None


def conv3d_cl(x, conv):
    if conv.groups != 1 or tuple(conv.dilation) != (1, 1, 1) or conv.padding_mode != "zeros":
        raise NotImplementedError("sm_100a conv3d: groups=1, dilation=1, zero padding only")
    return Conv3dFn.apply(x, conv.weight, conv.bias, tuple(conv.stride), tuple(conv.padding))


def _parity_classes(k: int, s: int, p: int, n_in: int, n_out: int):
    """Transposed conv along one dimension: output o = i*s - p + kk.  For each parity class r = o mod s returns
    (r, taps kk sorted by input offset, pad_lo, extra) such that the class is the stride-1 correlation
    out[s*j + r] = sum_t x[j + t - pad_lo] w[taps[t]], j in [0, J_r), J_r = n_in + 2*pad_lo - len(taps) + 1 + extra.
    None when a class cannot be expressed that way."""
    out = []
    for r in range(s):
        J = -(-(n_out - r) // s)
        if J <= 0:
            continue
        kks = [kk for kk in range(k) if (kk - r - p) % s == 0]
        if not kks:
            return None  # a class without taps is bias only: leave it to the scatter lowering
        deltas = sorted(((r + p - kk) // s, kk) for kk in kks)
        dmin, dmax = deltas[0][0], deltas[-1][0]
        if [d for d, _ in deltas] != list(range(dmin, dmax + 1)):
            return None
        pad_lo = -dmin
        K = len(deltas)
        if pad_lo < 0:
            return None
        extra = J - (n_in + 2 * pad_lo - K + 1)
        if extra < 0:
            return None
        out.append((r, [kk for _, kk in deltas], pad_lo, extra))
    return out


def conv_transpose3d_igemm(x, wrows16, bias, ks, stride, padding, out_sp, Cop):
    """ConvTranspose3d forward (= data gradient of the strided conv) without a column matrix: one implicit-GEMM launch
    per output parity class, each a stride-1 sub-convolution over its own taps (tapmap into the full filter `wrows16`
    [Cop, (kd,kh,kw,Cin)]), written straight to the class's sub-lattice of the output.  Returns None when the geometry
    does not fit (the caller then uses GEMM + col2im3d)."""
    N, d, h, w, Cip = x.shape
    if Cip % 32 != 0:
        return None
    per_dim = [_parity_classes(k, s, p, n, o) for k, s, p, n, o in zip(ks, stride, padding, (d, h, w), out_sp)]
    if any(c is None for c in per_dim):
        return None
    OD, OH, OW = out_sp
    plans = []
    for rz, tz, pz, ez in per_dim[0]:
        for ry, ty, py, ey in per_dim[1]:
            for rx, tx, px, ex in per_dim[2]:
                kern = (len(tz), len(ty), len(tx))
                if not ops.conv3d_igemm_supported((N, d, h, w, Cip), Cop, kern, (pz, py, px), extra=(ez, ey, ex)):
                    return None
                tapmap = [(a * ks[1] + b) * ks[2] + c for a in tz for b in ty for c in tx]
                plans.append(((rz, ry, rx), kern, (pz, py, px), (ez, ey, ex), tapmap))
    out = torch.empty((N, OD, OH, OW, Cop), device=x.device, dtype=x.dtype)
    sd, sh, sw = stride
    pitch = (sw * Cop, sh * OW * Cop, sd * OH * OW * Cop, OD * OH * OW * Cop)
    for (rz, ry, rx), kern, pad, extra, tapmap in plans:
        view = out[:, rz:, ry:, rx:]  # base pointer of the class's sub-lattice
        ops.conv3d_igemm(x, wrows16, bias, kern, pad, out=view, extra=extra, tapmap=tapmap, out_pitch=pitch)
    return out


def _convT_weight_rows_fwd(w, cin_pad, cout_pad):
    """transposed-conv filter [Cin, Cout, kd,kh,kw] (= nn.ConvTranspose3d.weight, or nn.Conv3d.weight read as the filter
    of its data gradient) -> fp32 [cout_pad, (kd,kh,kw,cin_pad)]: rows of the per-parity-class sub-convolutions"""
    Ci, Co = w.shape[:2]
    wk = w.permute(1, 2, 3, 4, 0)
    if cin_pad != Ci or cout_pad != Co:
        wk = torch.nn.functional.pad(wk, (0, cin_pad - Ci, 0, 0, 0, 0, 0, 0, 0, cout_pad - Co))
    return wk.reshape(cout_pad, -1).contiguous()


class ConvTranspose3dFn(Function):
    """nn.ConvTranspose3d on channels-last rows = the data gradient of the matching strided conv:
    rows @ W -> patch gradients -> col2im3d gather (+ bias)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, padding, output_padding):
        N, d, h, wd, Cip = x.shape
        Ci, Co = w.shape[:2]
        Cop = _pad8(Co)
        if Cip != _pad8(Ci):
            raise NotImplementedError(f"sm_100a conv_transpose3d: rows carry {Cip} channels, the filter expects {Ci}")
        ks = tuple(w.shape[2:])
        out_sp = tuple((i - 1) * s - 2 * p + k + op for i, s, p, k, op in zip((d, h, wd), stride, padding, ks, output_padding))
        geom = ops.conv3d_geom((N, *out_sp, Cop), ks, stride, padding)
        if (geom[14], geom[15], geom[16]) != (d, h, wd):
            raise NotImplementedError("sm_100a conv_transpose3d: geometry is not the adjoint of a strided conv")
        bias = None
        if b is not None:
            bias = (b.detach() if Cop == Co else torch.nn.functional.pad(b.detach(), (0, Cop - Co))).contiguous()
        # one implicit-GEMM sub-convolution per output parity class when the geometry tiles, else GEMM + col2im scatter
        out = conv_transpose3d_igemm(x, ops.cast_pack(_convT_weight_rows_fwd(w.detach(), Cip, Cop), x.dtype), bias, ks,
                                     stride, padding, out_sp, Cop)
        if out is None:
            wt = _convT_weight_rows(w.detach(), Cip, Cop)  # [(kd,kh,kw,co), ci]
            dcol = ops.gemm(x.view(-1, Cip), ops.cast_pack(wt, x.dtype))
            out = ops.col2im3d(dcol, geom)
            if bias is not None:
                out = ops.add_rows(out, None, bias)
        ctx.save_for_backward(x, w)
        ctx.meta = (geom, Cop, b is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        geom, Cop, has_bias = ctx.meta
        Cip = x.shape[-1]
        Ci, Co = w.shape[:2]
        ks = tuple(w.shape[2:])
        stride, padding = tuple(geom[8:11]), tuple(geom[11:14])
        dout = dout.contiguous()
        wt = _convT_weight_rows(w, Cip, Cop)  # [(kd,kh,kw,co), ci]
        x2 = x.view(-1, Cip)
        col = None
        # dx = the strided forward conv of dout (adjoint of this transposed conv): implicit GEMM when the geometry tiles
        if ops.conv3d_igemm_supported(tuple(dout.shape), Cip, ks, padding, stride=stride):
            dx = ops.conv3d_igemm(dout, ops.cast_pack(wt, x.dtype, transpose=True), None, ks, padding, stride=stride)
        else:
            col = ops.im2col3d(dout, geom)  # [M_in, (kd,kh,kw,co)]
            dx = ops.gemm(col, ops.cast_pack(wt, x.dtype, transpose=True)).view(x.shape)
        # dW[ci, (tap, co)] = sum_v x[v, ci] * dout[v*s + tap - p, co]: the strided weight gradient with x as "dout"
        if ops.conv3d_igemm_supported(tuple(dout.shape), Cip, ks, padding, wgrad=True, stride=stride):
            dwc = ops.conv3d_igemm_wgrad(dout, x, ks, padding, stride=stride)  # [Cip, (kd,kh,kw,Cop)]
            dw = dwc.view(Cip, *ks, Cop)[:Ci, ..., :Co].permute(0, 4, 1, 2, 3)
        else:
            if col is None:
                col = ops.im2col3d(dout, geom)
            dwt = ops.gemm(col, x2, mn_major=True, epilogue=L.EPI_F32, k_splits=_wgrad_splits(col.shape[1], Cip, col.shape[0]))
            dw = dwt.view(*ks, Cop, Cip)[..., :Co, :Ci].permute(4, 3, 0, 1, 2)
        db = ops.colreduce(dout.view(1, -1, Cop), 0).view(Cop)[:Co] if has_bias else None
        return dx, dw, db, None, None, None


def _convT_weight_rows(w, cin_pad, cout_pad):
    """nn.ConvTranspose3d weight [Ci,Co,kd,kh,kw] -> fp32 [(kd,kh,kw,cout_pad), cin_pad]"""
    Ci, Co = w.shape[:2]
    wt = w.permute(2, 3, 4, 1, 0)  # [kd,kh,kw,Co,Ci]
    if cin_pad != Ci or cout_pad != Co:
        wt = torch.nn.functional.pad(wt, (0, cin_pad - Ci, 0, cout_pad - Co))
    return wt.reshape(-1, cin_pad).contiguous()


def conv_transpose3d_cl(x, conv):
    if conv.groups != 1 or tuple(conv.dilation) != (1, 1, 1):
        raise NotImplementedError("sm_100a conv_transpose3d: groups=1, dilation=1 only")
    return ConvTranspose3dFn.apply(x, conv.weight, conv.bias, tuple(conv.stride), tuple(conv.padding),
                                   tuple(conv.output_padding))


class BatchNormActFn(Function):
    """nn.BatchNorm3d (+ ReLU) on channels-last rows: y = act(x * scale + shift) with the operands `batchnorm_act_cl`
    prepared (statistics, running-statistics update and the affine coefficients come from one column-reduction pass and
    one finalize launch); backward = the full batch-norm backward (through the batch statistics in training mode)."""

    @staticmethod
    def forward(ctx, x, weight, bias, scale, shift, mean, rstd, training, relu):
        y = ops.affine_act(x, scale, shift, relu)
        ctx.save_for_backward(x, y, weight, mean, rstd)
        ctx.flags = (training, relu)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, y, weight, mean, rstd = ctx.saved_tensors
        training, relu = ctx.flags
        Cc = x.shape[-1]
        gamma = weight.detach()
        if gamma.numel() != Cc:  # channel-padded rows
            gamma = torch.nn.functional.pad(gamma, (0, Cc - gamma.numel()))
        dx, dg, db = ops.bn_bwd(dy.contiguous(), x, y, mean, rstd, gamma, relu, training)
        n = weight.numel()
        return dx, dg[:n], db[:n], None, None, None, None, None, None


def batchnorm_act_cl(x, bn, relu):
    training = bn.training or bn.running_mean is None
    Cc, Cn = x.shape[-1], bn.num_features
    M = x.numel() // Cc
    rm, rv = bn.running_mean, bn.running_var
    with torch.no_grad():
        sums = pivot = None
        momentum = -1.0
        if training:
            # sums and sums of squares in one pass, taken about the running mean (a per-channel pivot close to the batch
            # mean): E[(x-p)^2] - E[x-p]^2 does not cancel catastrophically when |mean| >> std
            if rm is not None:
                pivot = rm if Cn == Cc else torch.nn.functional.pad(rm, (0, Cc - Cn))
                pivot = pivot.contiguous()
            sums = ops.colreduce(x.view(1, M, Cc), 2, pivot=pivot)
            if bn.training and bn.track_running_stats and rm is not None:
                bn.num_batches_tracked += 1
                momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
        scale, shift, mean, rstd = ops.bn_finalize(sums, pivot, bn.weight, bn.bias, rm, rv, Cc, M, bn.eps, momentum)
    return BatchNormActFn.apply(x, bn.weight, bn.bias, scale, shift, mean, rstd, training, relu)


class GroupNormActFn(Function):
    """nn.GroupNorm -> [* (scale + 1) + shift] -> SiLU | ReLU | identity on channels-last rows [N, D, H, W, Cp]
    (VM/unet/blocks.py:107-113).  Statistics per (sample, group) from one column-reduction pass; the elementwise
    passes take per-(sample, channel) coefficients (see csrc/groupnorm_sm100.cu).  Channel padding (Cp > C) stays zero."""

    @staticmethod
    def forward(ctx, x, weight, bias, groups, eps, scale, shift, act):
        N, Cp = x.shape[0], x.shape[-1]
        Cn = weight.numel()
        R = x.numel() // (N * Cp)
        cg = Cn // groups
        st = ops.colreduce(x.view(N, R, Cp), 2)[:, :, :Cn].reshape(2, N, groups, cg).sum(-1) / (R * cg)
        mean = st[0]
        rstd = torch.rsqrt((st[1] - mean * mean).clamp_min_(0.0) + eps)  # [N, G]
        mean_c = mean.repeat_interleave(cg, dim=1)  # [N, C]
        rstd_c = rstd.repeat_interleave(cg, dim=1)
        s1p = 1.0 if scale is None else (scale.detach().float() + 1.0)
        w = weight.detach() * s1p * torch.ones_like(rstd_c)  # gamma (1 + scale)   [N, C]
        a = rstd_c * w
        b = (bias.detach() * s1p - mean_c * a) + (0.0 if shift is None else shift.detach().float())
        if Cp != Cn:
            a, b = torch.nn.functional.pad(a, (0, Cp - Cn)), torch.nn.functional.pad(b, (0, Cp - Cn))
        a, b = a.contiguous(), b.contiguous()
        y = ops.affine_nc_act(x, a, b, act)
        ctx.save_for_backward(x, a, b, mean_c, rstd_c, w, weight, bias, scale)
        ctx.meta = (groups, act, Cn, shift is not None)
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, a, b, mean_c, rstd_c, w, weight, bias, scale = ctx.saved_tensors
        groups, act, Cn, has_shift = ctx.meta
        N, Cp = x.shape[0], x.shape[-1]
        R = x.numel() // (N * Cp)
        cg = Cn // groups
        dy = dy.contiguous()
        S1, S2 = ops.gn_bwd_reduce(dy, x, a, b, act)
        S1, S2 = S1[:, :Cn], S2[:, :Cn]
        S2h = rstd_c * (S2 - mean_c * S1)  # sum_r dv * xhat
        cnt = float(R * cg)
        m1 = (w * S1).view(N, groups, cg).sum(-1, keepdim=True).expand(N, groups, cg).reshape(N, Cn) / cnt
        m2 = (w * S2h).view(N, groups, cg).sum(-1, keepdim=True).expand(N, groups, cg).reshape(N, Cn) / cnt
        c1 = rstd_c * w
        c2 = -rstd_c * rstd_c * m2
        c3 = -rstd_c * m1 - c2 * mean_c
        coef = torch.stack([a[:, :Cn], b[:, :Cn], c1, c2, c3])
        if Cp != Cn:
            coef = torch.nn.functional.pad(coef, (0, Cp - Cn))
        dx = ops.gn_bwd_apply(dy, x, coef.contiguous(), act)
        s1p = 1.0 if scale is None else (scale.float() + 1.0)
        dgamma = (s1p * S2h).sum(0)
        dbeta = (s1p * S1).sum(0)
        dscale = (weight * S2h + bias * S1).to(scale.dtype) if scale is not None else None
        dshift = S1.clone() if has_shift else None
        return dx, dgamma, dbeta, None, None, dscale, dshift, None


def groupnorm_act_cl(x, gn, act: str, scale=None, shift=None):
    """gn: nn.GroupNorm; scale / shift: [N, C] timestep conditioning or None; act: "silu" | "relu" | "none"."""
    return GroupNormActFn.apply(x, gn.weight, gn.bias, gn.num_groups, gn.eps, scale, shift, act)


def instancenorm_act_cl(x, eps: float, act: str = "none"):
    """nn.InstanceNorm3d (affine=False, batch statistics) [+ activation] on channels-last rows: GroupNorm with one group
    per channel and unit affine parameters."""
    Cn = x.shape[-1]
    one = torch.ones((Cn,), device=x.device, dtype=torch.float32)
    return GroupNormActFn.apply(x, one, torch.zeros_like(one), Cn, eps, None, None, act)


class ScaleActFn(Function):
    """act(x * scale[n, c]): nn.Dropout3d (per sample-and-channel scale, or None) fused with any ConvBlock3D activation
    (relu | leakyrelu | elu | selu | linear) through the per-(sample, channel) affine + activation kernels."""

    @staticmethod
    def forward(ctx, x, scale, act):
        N, Cp = x.shape[0], x.shape[-1]
        a = torch.ones((N, Cp), device=x.device, dtype=torch.float32) if scale is None else scale.contiguous()
        b = torch.zeros_like(a)
        ctx.save_for_backward(x, a, b)
        ctx.act = act
        return ops.affine_nc_act(x, a, b, act)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        x, a, b = ctx.saved_tensors
        zero = torch.zeros_like(a)
        coef = torch.stack([a, b, a, zero, zero]).contiguous()  # dx = a * dy * act'(a x)
        return ops.gn_bwd_apply(dy.contiguous(), x, coef, ctx.act), None, None


def scale_act_cl(x, scale, act: str):
    return ScaleActFn.apply(x, scale, act)


class Cat2Fn(Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.c = (a.shape[-1], b.shape[-1])
        return ops.cat2(a, b)

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        return ops.split2(dout.contiguous(), *ctx.c)


def cat_cl(a, b, ca=None, cb=None):
    """torch.cat([a, b], channel dim) on channels-last rows.  ca / cb: the REAL channel counts when the rows are padded
    to a multiple of 8: the result must be [a | b | zero pad] (the layout the next conv's weight rows assume), so
    padded operands are compacted first (plain torch ops; only channel counts that are not multiples of 8 pay this)."""
    ca = a.shape[-1] if ca is None else ca
    cb = b.shape[-1] if cb is None else cb
    if ca == a.shape[-1] and cb == b.shape[-1]:
        return Cat2Fn.apply(a, b)
    out = torch.cat([a[..., :ca], b[..., :cb]], dim=-1)
    pad = _pad8(ca + cb) - (ca + cb)
    return torch.nn.functional.pad(out, (0, pad)) if pad else out.contiguous()


class AddFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        return ops.add_rows(a, b)

    @staticmethod
    def backward(ctx, dout):
        return dout, dout


def add_cl(a, b):
    return AddFn.apply(a, b)


# --------------------------------------------------------------------------------------------------
# 2.5-D U-Net (Unet25d / ConvBlock3D): Dropout3d + ReLU, (1,2,2) average pooling, (1,2,2) trilinear upsampling
class ScaleReluFn(Function):
    """relu?(x * scale[n, c]): nn.Dropout3d (scale = keep mask / (1 - p) per sample and channel, or None) fused with nn.ReLU."""

    @staticmethod
    def forward(ctx, x, scale, relu):
        y = ops.scale_relu(x, scale, relu)
        ctx.save_for_backward(y if relu else None, scale)
        ctx.relu = relu
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        y, scale = ctx.saved_tensors
        return ops.scale_relu(dy.contiguous(), scale, ctx.relu, gate=y), None, None


def scale_relu_cl(x, scale, relu):
    return ScaleReluFn.apply(x, scale, relu)


def dropout3d_scale(x, p: float, training: bool):
    """Per-(sample, channel) scale of nn.Dropout3d on channels-last rows, or None when it is the identity."""
    if not training or p <= 0.0:
        return None
    if p >= 1.0:
        return torch.zeros((x.shape[0], x.shape[-1]), device=x.device, dtype=torch.float32)
    keep = torch.full((x.shape[0], x.shape[-1]), 1.0 - p, device=x.device, dtype=torch.float32)
    return torch.bernoulli(keep) / (1.0 - p)


class AvgPoolHW2Fn(Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape = tuple(x.shape)
        return ops.avgpool_hw2(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        return ops.avgpool_hw2(dy.contiguous(), backward_shape=ctx.shape)


def avgpool_hw2_cl(x):
    return AvgPoolHW2Fn.apply(x)


class Upsample2xHWFn(Function):
    @staticmethod
    def forward(ctx, x):
        return ops.upsample2x_hw(x)

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        return ops.upsample2x_hw(dy.contiguous(), backward=True)


def upsample2x_hw_cl(x):
    return Upsample2xHWFn.apply(x)
