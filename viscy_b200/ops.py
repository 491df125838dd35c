"""Raw (non-autograd) launchers of the sm_100a kernels through the C ABI.

Every function takes CUDA tensors, launches on torch's current stream and returns torch tensors
allocated by the caller side (the library never allocates).  Autograd wiring lives in `functional.py`.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def _rowmajor2d(t: torch.Tensor, name: str) -> int:
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be a 2-D tensor with unit inner stride, got {tuple(t.shape)} / {t.stride()}")
    return t.stride(0)


def gemm(
    a: torch.Tensor,
    b: torch.Tensor,
    *,
    bias: torch.Tensor | None = None,
    residual: torch.Tensor | None = None,
    aux: torch.Tensor | None = None,
    act: int = L.ACT_NONE,
    epilogue: int = L.EPI_STORE,
    mn_major: bool = False,
    k_splits: int = 1,
    out: torch.Tensor | None = None,
    out2: torch.Tensor | None = None,
    accumulate: bool = False,
    b_batch_rows: int = 0,
    aux2: torch.Tensor | None = None,
    tvec: torch.Tensor | None = None,
    svec: torch.Tensor | None = None,
    rows_per_sample: int = 0,
    split_slabs: bool = False,
    n_split: int = 0,
    rvec: torch.Tensor | None = None,
    rvec_rows: int = 0,
    colsq: torch.Tensor | None = None,
) -> torch.Tensor | tuple[torch.Tensor, torch.Tensor]:
    """tcgen05 GEMM.  K-major form: a [M,K], b [N,K] -> [M,N].  MN-major (wgrad) form: a [P,M], b [P,N]."""
    lda = _rowmajor2d(a, "a")
    ldb = _rowmajor2d(b, "b")
    if a.dtype != b.dtype:
        raise ValueError("a and b must share a dtype")
    if mn_major:
        K, M = a.shape
        K2, N = b.shape
    else:
        M, K = a.shape
        N, K2 = b.shape
        if b_batch_rows:
            N = N // (-(-M // b_batch_rows))
    if K != K2:
        raise ValueError(f"reduction dims differ: {K} vs {K2}")
    d = L.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.dtype = L.dtype_code(a.dtype)
    d.mn_major = 1 if mn_major else 0
    d.epilogue = epilogue
    d.act = act
    d.k_splits = max(1, k_splits)
    d.b_batch_rows = b_batch_rows
    d.rows_per_sample = rows_per_sample
    if epilogue == L.EPI_F32 and split_slabs:
        # one fp32 slab per K-split (no atomics): out [k_splits, M, N]
        if out is None:
            out = torch.empty((d.k_splits, M, N), device=a.device, dtype=torch.float32)
        d.atomic_out = 0
        d.split_out_stride = M * N
        d.lda, d.ldb = lda, ldb
        d.ldo = N
        d.A, d.B, d.out = a.data_ptr(), b.data_ptr(), out.data_ptr()
        L.check(L.lib().vb200_gemm(C.byref(d), L.stream_ptr()), "vb200_gemm")
        return out
    if epilogue == L.EPI_F32 and n_split:
        # columns >= n_split leave through out2 (fp32 [M, ldo2]): dW to `out` [M, n_split], the ones-column sums to out2
        if out is None or out2 is None:
            raise ValueError("n_split needs caller-provided out [M, n_split] and out2 [M, >= N - n_split] (fp32)")
        d.n_split = n_split
        d.atomic_out = 1 if (k_splits > 1 or accumulate) else 0
        d.lda, d.ldb = lda, ldb
        d.ldo, d.ldo2 = _rowmajor2d(out, "out"), _rowmajor2d(out2, "out2")
        d.A, d.B, d.out, d.out2 = a.data_ptr(), b.data_ptr(), out.data_ptr(), out2.data_ptr()
        L.check(L.lib().vb200_gemm(C.byref(d), L.stream_ptr()), "vb200_gemm")
        return out, out2
    if epilogue == L.EPI_F32:
        if out is None:
            out = zeros((M, N), a.device) if (k_splits > 1 or accumulate) else torch.empty(
                (M, N), device=a.device, dtype=torch.float32)
        d.atomic_out = 1 if (k_splits > 1 or accumulate) else 0
    else:
        if out is None:
            out = torch.empty((M, N), device=a.device, dtype=a.dtype)
    d.lda, d.ldb = lda, ldb
    d.ldo = _rowmajor2d(out, "out")
    d.A, d.B, d.out = a.data_ptr(), b.data_ptr(), out.data_ptr()
    if epilogue in (L.EPI_GELU_DUAL, L.EPI_GELU_GP):
        if out2 is None:
            out2 = torch.empty((M, N), device=a.device, dtype=a.dtype)
        d.out2 = out2.data_ptr()
        d.ldo2 = _rowmajor2d(out2, "out2")
    if bias is not None:
        if bias.dtype != torch.float32 or not bias.is_contiguous() or bias.numel() != N:
            raise ValueError("bias must be contiguous fp32 [N]")
        d.bias = bias.data_ptr()
    if residual is not None:
        d.residual = residual.data_ptr()
        d.ldr = _rowmajor2d(residual, "residual")
    if aux is not None:
        d.aux = aux.data_ptr()
        d.ldaux = _rowmajor2d(aux, "aux")
    if aux2 is not None:
        d.aux2 = aux2.data_ptr()
        d.ldaux2 = _rowmajor2d(aux2, "aux2")
    if tvec is not None:
        if tvec.dtype != torch.float32 or not tvec.is_contiguous():
            raise ValueError("tvec must be contiguous fp32")
        d.tvec = tvec.data_ptr()
    if svec is not None:
        if svec.dtype != torch.float32 or not svec.is_contiguous():
            raise ValueError("svec must be contiguous fp32")
        d.svec = svec.data_ptr()
    if colsq is not None:
        if colsq.dtype != torch.float32 or not colsq.is_contiguous() or rows_per_sample <= 0:
            raise ValueError("colsq must be contiguous fp32 [M / rows_per_sample, N] with rows_per_sample set")
        d.colsq = colsq.data_ptr()
    if rvec is not None:
        if rvec.dtype != torch.float32 or not rvec.is_contiguous() or rvec_rows <= 0 or rvec.numel() * rvec_rows < M:
            raise ValueError("rvec must be contiguous fp32 with one entry per rvec_rows rows")
        d.rvec = rvec.data_ptr()
        d.rvec_rows = rvec_rows
    L.check(L.lib().vb200_gemm(C.byref(d), L.stream_ptr()), "vb200_gemm")
    if epilogue in (L.EPI_GELU_DUAL, L.EPI_GELU_GP):
        return out, out2
    return out


def gemm_uses_wide_tiles(M: int, N: int) -> bool:
    """True when vb200_gemm picks the 256-wide (TMA-store) tiles for a K-major [M, N] product: the same rule as the host
    code (N > 128 and at least one tile per SM)."""
    return N > 128 and (-(-M // 128)) * (-(-N // 256)) >= SM_COUNT


SM_COUNT = 148


# ----------------------------------------------------------------------------------------------
# helpers
def _p(t):
    return L.ptr(t)


def _call(name: str, *args) -> None:
    L.check(getattr(L.lib(), name)(*args, L.stream_ptr()), name)


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous fp32")
    return t


def _act(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous (channels-last rows)")
    L.dtype_code(t.dtype)
    return t


class StepArena:
    """Every zero-initialised fp32 accumulator of one step (split-K weight gradients, bias / norm-parameter gradient
    sums, GRN statistics) served from ONE allocation: one memset instead of ~100 fill launches per step.

    `begin()` (called by a model at the start of its forward) allocates a fresh zero buffer sized from the requests of
    the previous step with the same (device, grad mode) key; slices are handed out once and never recycled, so a slice
    is zero when handed out and stays valid for as long as any view of it lives (gradients returned to autograd keep the
    allocation alive; the next step gets a new one).  Requests beyond the buffer fall back to torch.zeros."""

    def __init__(self):
        self.buf = None
        self.off = 0
        self.need = 0
        self.key = None
        self.cap: dict = {}

    def begin(self, device, grad: bool) -> None:
        if self.key is not None:
            self.cap[self.key] = max(self.cap.get(self.key, 0), self.need)
        self.key = (device, grad)
        self.need = self.off = 0
        size = self.cap.get(self.key, 0)
        self.buf = torch.zeros((size,), device=device, dtype=torch.float32) if size else None

    def zeros(self, shape, device) -> torch.Tensor:
        n = 1
        for d in shape:
            n *= int(d)
        n_al = -(-n // 64) * 64  # 256-byte granules: every slice is aligned for 16-byte vector reductions
        self.need += n_al
        b = self.buf
        if b is not None and b.device == device and self.off + n_al <= b.numel():
            v = b[self.off:self.off + n].view(*shape)
            self.off += n_al
            return v
        return torch.zeros(tuple(shape), device=device, dtype=torch.float32)


STEP = StepArena()


def zeros(shape, device) -> torch.Tensor:
    """Zero-initialised fp32 accumulator (from the step arena when one is active)."""
    return STEP.zeros(tuple(shape) if not isinstance(shape, int) else (shape,), device)


def cast_pack(w: torch.Tensor, dtype: torch.dtype, transpose: bool = False) -> torch.Tensor:
    """fp32 [R, C] parameter -> 16-bit operand copy ([C, R] when transpose)."""
    w2 = _f32(w.detach().reshape(w.shape[0], -1), "w")
    R, Cc = w2.shape
    out = torch.empty((Cc, R) if transpose else (R, Cc), device=w.device, dtype=dtype)
    _call("vb200_cast_pack", _p(w2), _p(out), C.c_int64(R), C.c_int64(Cc), int(transpose), L.dtype_code(dtype))
    return out


def conv_weight_rows(w: torch.Tensor, cin_pad: int, cout_pad: int, dtype: torch.dtype, flipped: bool = False) -> torch.Tensor:
    """nn.Conv3d weight [Co,Ci,kd,kh,kw] fp32 -> 16-bit [cout_pad, (kd,kh,kw,cin_pad)], or with flipped=True the data
    gradient's filter [cin_pad, (kd,kh,kw flipped, cout_pad)]; zero padded; one launch."""
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():  # e.g. the flipped / transposed view a transposed conv passes
        w = w.float().contiguous()
    Co, Ci, KD, KH, KW = w.shape
    rows, inner = (cin_pad, cout_pad) if flipped else (cout_pad, cin_pad)
    out = torch.empty((rows, KD * KH * KW * inner), device=w.device, dtype=dtype)
    _call("vb200_conv_weight_rows", _p(w), _p(out), Co, Ci, KD, KH, KW, cin_pad, cout_pad, int(flipped), L.dtype_code(dtype))
    return out


class Arena:
    """One zero-filled fp32 allocation handed out in slices (accumulators that kernels add into)."""

    def __init__(self, device, sizes):
        self.buf = zeros((int(sum(sizes)),), device)
        self.off = 0

    def take(self, *shape):
        n = 1
        for d in shape:
            n *= d
        v = self.buf[self.off:self.off + n].view(*shape)
        self.off += n
        return v


def dw_pack(w):
    """conv_dw.weight [C,1,7,7] fp32 -> (tap-major [49,C], flipped tap-major [49,C])."""
    Cc = w.shape[0]
    out = torch.empty((2, 49, Cc), device=w.device, dtype=torch.float32)
    _call("vb200_dw_pack", _p(_f32(w.detach(), "conv_dw.weight")), _p(out[0]), _p(out[1]), Cc)
    return out[0], out[1]


# ----------------------------------------------------------------------------------------------
# ConvNeXt block pieces
def dwconv7(x, wt, bias=None, add=None):
    """x [B,H,W,C] 16-bit, wt fp32 [49,C] tap-major."""
    _act(x, "x")
    B, H, W, Cc = x.shape
    y = torch.empty_like(x)
    _call("vb200_dwconv7", _p(x), _p(_f32(wt, "wt")), _p(bias), _p(add), _p(y), B, H, W, Cc, L.dtype_code(x.dtype))
    return y


def dwconv7_wgrad(x, dy, want_bias=True, arena=None):
    B, H, W, Cc = x.shape
    if arena is not None:
        dwt, db = arena.take(49, Cc), (arena.take(Cc) if want_bias else None)
    else:
        dwt = zeros((49, Cc), x.device)
        db = zeros((Cc,), x.device) if want_bias else None
    _call("vb200_dwconv7_wgrad", _p(_act(x, "x")), _p(_act(dy, "dy")), _p(dwt), _p(db), B, H, W, Cc, L.dtype_code(x.dtype))
    return dwt, db


ONES_PAD = 16  # columns appended for the ones-column trick: {1, 0 x 15} keeps row pitches multiples of 32 bytes


def layernorm_fwd(x, gamma, beta, eps, ones=False, ones2=None, ones2_col=0):
    """x [..., C] 16-bit rows -> (y, mean, rstd).

    ones=True: y is [M, C + ONES_PAD] with y[:, C] = 1 and y[:, C+1:] = 0 (use y[:, :C] as the normalised rows): a weight
    gradient GEMM against it yields the bias gradient as an extra column.  ones2 ([M, ld2] 16-bit, ones2_col % 8 == 0):
    a second matrix that receives the same {1, 0, ...} columns per row."""
    _act(x, "x")
    Cc = x.shape[-1]
    M = x.numel() // Cc
    mean = torch.empty((M,), device=x.device, dtype=torch.float32)
    rstd = torch.empty((M,), device=x.device, dtype=torch.float32)
    if not ones and ones2 is None:
        y = torch.empty_like(x)
        _call("vb200_layernorm_fwd", _p(x), _p(_f32(gamma, "gamma")), _p(_f32(beta, "beta")), _p(y), _p(mean), _p(rstd),
              C.c_int64(M), Cc, C.c_float(eps), L.dtype_code(x.dtype))
        return y, mean, rstd
    ldy = Cc + ONES_PAD if ones else Cc
    y = torch.empty((M, ldy), device=x.device, dtype=x.dtype)
    ld2 = 8
    if ones2 is not None:
        if ones2.dtype != x.dtype or ones2.dim() != 2 or ones2.shape[0] != M or ones2.stride(1) != 1 or not ones:
            raise ValueError("ones2 must be a [M, ld] matrix of the activation dtype (and needs ones=True)")
        ld2 = ones2.stride(0)
    _call("vb200_layernorm_fwd_ld", _p(x), _p(_f32(gamma, "gamma")), _p(_f32(beta, "beta")), _p(y), C.c_int64(ldy),
          ONES_PAD // 8 if ones else 0, _p(ones2), C.c_int64(ld2), int(ones2_col), _p(mean), _p(rstd), C.c_int64(M), Cc,
          C.c_float(eps), L.dtype_code(x.dtype))
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, arena=None):
    Cc = x.shape[-1]
    M = x.numel() // Cc
    dx = torch.empty_like(x)
    if arena is not None:
        dgamma, dbeta = arena.take(Cc), arena.take(Cc)
    else:
        dgamma = zeros((Cc,), x.device)
        dbeta = zeros((Cc,), x.device)
    _call("vb200_layernorm_bwd", _p(_act(dy, "dy")), _p(_act(x, "x")), _p(mean), _p(rstd), _p(_f32(gamma, "gamma")),
          _p(dx), _p(dgamma), _p(dbeta), C.c_int64(M), Cc, L.dtype_code(x.dtype))
    return dx, dgamma, dbeta


def gelu_grn_fwd(h, w, b, eps=1e-6):
    """h [B, R, C] 16-bit -> (y, sumsq [B,C], s [B,C])."""
    _act(h, "h")
    B, R, Cc = h.shape
    dt = L.dtype_code(h.dtype)
    sumsq = zeros((B, Cc), h.device)
    _call("vb200_grn_sumsq", _p(h), _p(sumsq), B, R, Cc, dt)
    s = torch.empty_like(sumsq)
    _call("vb200_grn_coef_fwd", _p(sumsq), _p(_f32(w, "grn.weight")), _p(s), B, Cc, C.c_float(eps))
    y = torch.empty_like(h)
    _call("vb200_grn_apply_fwd", _p(h), _p(s), _p(_f32(b, "grn.bias")), _p(y), B, R, Cc, dt)
    return y, sumsq, s


def gelu_grn_bwd(h, dy, sumsq, s, w, eps=1e-6, want_dbias=True):
    """-> (dh, d grn.weight, d grn.bias, d fc1.bias)"""
    B, R, Cc = h.shape
    dt = L.dtype_code(h.dtype)
    dev = h.device
    S1 = zeros((B, Cc), dev)
    sdy = zeros((Cc,), dev)
    _call("vb200_grn_bwd_reduce", _p(_act(h, "h")), _p(_act(dy, "dy")), _p(S1), _p(sdy), B, R, Cc, dt)
    t = torch.empty_like(S1)
    dw = zeros((Cc,), dev)
    _call("vb200_grn_coef_bwd", _p(sumsq), _p(S1), _p(_f32(w, "grn.weight")), _p(t), _p(dw), B, Cc, C.c_float(eps))
    dh = torch.empty_like(h)
    dbias = zeros((Cc,), dev) if want_dbias else None
    _call("vb200_grn_apply_bwd", _p(h), _p(dy), _p(s), _p(t), _p(dh), _p(dbias), B, R, Cc, dt)
    return dh, dw, sdy, dbias


def colsum(x):
    Cc = x.shape[-1]
    M = x.numel() // Cc
    out = zeros((Cc,), x.device)
    _call("vb200_colsum", _p(_act(x, "x")), _p(out), C.c_int64(M), Cc, L.dtype_code(x.dtype))
    return out


# ----------------------------------------------------------------------------------------------
# layout
def pixshuf_cat_fwd(prev, skip):
    _act(prev, "prev")
    B, h, w, Cp = prev.shape
    Cs = 0 if skip is None else skip.shape[-1]
    if skip is not None:
        _act(skip, "skip")
        if skip.shape[:3] != (B, 2 * h, 2 * w):
            raise ValueError(f"skip {tuple(skip.shape)} does not match upsampled {(B, 2 * h, 2 * w)}")
    out = torch.empty((B, 2 * h, 2 * w, Cp // 4 + Cs), device=prev.device, dtype=prev.dtype)
    _call("vb200_pixshuf_cat_fwd", _p(prev), _p(skip), _p(out), B, h, w, Cp, Cs)
    return out


def pixshuf_cat_bwd(dout, Cp, Cs):
    _act(dout, "dout")
    B, H2, W2, _ = dout.shape
    h, w = H2 // 2, W2 // 2
    dprev = torch.empty((B, h, w, Cp), device=dout.device, dtype=dout.dtype)
    dskip = torch.empty((B, H2, W2, Cs), device=dout.device, dtype=dout.dtype) if Cs else None
    _call("vb200_pixshuf_cat_bwd", _p(dout), _p(dprev), _p(dskip), B, h, w, Cp, Cs)
    return dprev, dskip


def patchify2(x, inverse=False, shape=None):
    """forward: [B,H,W,C] -> [B,H/2,W/2,4C];  inverse: rows back to [B,H,W,C] (shape = (B,H,W,C))."""
    _act(x, "x")
    if not inverse:
        B, H, W, Cc = x.shape
        out = torch.empty((B, H // 2, W // 2, 4 * Cc), device=x.device, dtype=x.dtype)
    else:
        B, H, W, Cc = shape
        out = torch.empty((B, H, W, Cc), device=x.device, dtype=x.dtype)
    _call("vb200_patchify2", _p(x), _p(out), B, H, W, Cc, int(inverse))
    return out


_XDT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}


def stem_patchify(x, kH, kW, dtype):
    """x (B,Cin,D,H,W) NCDHW fp32/16-bit -> rows [B*H/kH*W/kW, Cin*D*kH*kW] in `dtype`."""
    if not x.is_contiguous():
        x = x.contiguous()
    B, Cin, D, H, W = x.shape
    K = Cin * D * kH * kW
    if K % 8:
        raise NotImplementedError(f"stem K={K} must be a multiple of 8")
    A = torch.empty((B * (H // kH) * (W // kW), K), device=x.device, dtype=dtype)
    _call("vb200_stem_patchify", _p(x), _XDT[x.dtype], _p(A), B, Cin, D, H, W, kH, kW, K, L.dtype_code(dtype))
    return A


def conv3d_geom(shape, kernel, stride, padding):
    N, D, H, W, Cc = shape
    kd, kh, kw = kernel
    sd, sh, sw = stride
    pd, ph, pw = padding
    OD = (D + 2 * pd - kd) // sd + 1
    OH = (H + 2 * ph - kh) // sh + 1
    OW = (W + 2 * pw - kw) // sw + 1
    return [N, D, H, W, Cc, kd, kh, kw, sd, sh, sw, pd, ph, pw, OD, OH, OW]


def _conv3d_desc(shape, cout, kernel, padding, dtype_code=L.BF16, stride=(1, 1, 1)):
    d = L.Conv3dDesc()
    d.N, d.D, d.H, d.W, d.cin = shape
    d.cout = cout
    d.kd, d.kh, d.kw = kernel
    d.pd, d.ph, d.pw = padding
    d.sd, d.sh, d.sw = stride
    d.dtype = dtype_code
    return d


def _conv3d_out(shape, kernel, padding, stride):
    return tuple((i + 2 * p - k) // s + 1 for i, k, p, s in zip(shape, kernel, padding, stride))


def conv3d_igemm_supported(shape, cout, kernel, padding, wgrad=False, stride=(1, 1, 1), extra=(0, 0, 0)) -> bool:
    """Host-only geometry query: does the TMA implicit-GEMM form cover this conv ([N,D,H,W,C] input)?"""
    fn = L.lib().vb200_conv3d_igemm_supported
    d = _conv3d_desc(shape, cout, kernel, padding, stride=stride)
    d.xd, d.xh, d.xw = extra
    return bool(fn(C.byref(d), int(wgrad)))


def conv3d_igemm(x, w16, bias, kernel, padding, act=L.ACT_NONE, residual=None, out=None, stride=(1, 1, 1),
                 extra=(0, 0, 0), tapmap=None, out_pitch=None):
    """Conv3d as an implicit GEMM on tcgen05: x [N,D,H,W,Cin] 16-bit, w16 [Cout, (kd,kh,kw,Cin)] 16-bit K-major.
    Returns [N,OD,OH,OW,Cout].  The patch matrix is never materialised (TMA boxes at tap-shifted coordinates).
    `extra`: additional output extent per dimension (asymmetric high-side padding).  `tapmap`: tap t of `kernel` uses
    the weight columns of tap tapmap[t] of a wider w16.  `out_pitch` (x, y, z, n element pitches): scatter the output
    voxels into `out` (required then), e.g. one parity class of a transposed conv."""
    _act(x, "x")
    N, D, H, W, Ci = x.shape
    Co = w16.shape[0]
    kd, kh, kw = kernel
    wtaps = w16.shape[1] // Ci
    if w16.dtype != x.dtype or w16.shape[1] != wtaps * Ci or not w16.is_contiguous() or \
            (tapmap is None and wtaps != kd * kh * kw):
        raise ValueError(f"w16 must be contiguous {x.dtype} [Cout, taps * {Ci}], got {tuple(w16.shape)} {w16.dtype}")
    d = _conv3d_desc((N, D, H, W, Ci), Co, kernel, padding, L.dtype_code(x.dtype), stride)
    d.xd, d.xh, d.xw = extra
    OD, OH, OW = (o + e for o, e in zip(_conv3d_out((D, H, W), kernel, padding, stride), extra))
    keep = None
    if tapmap is not None:
        if len(tapmap) != kd * kh * kw:
            raise ValueError("tapmap needs one entry per tap of `kernel`")
        keep = (C.c_int32 * len(tapmap))(*tapmap)
        d.tapmap = C.cast(keep, C.POINTER(C.c_int32))
        d.w_taps = wtaps
    if out_pitch is not None:
        if out is None:
            raise ValueError("out_pitch needs the destination tensor `out`")
        for i, v in enumerate(out_pitch):
            d.out_pitch[i] = v
    if out is None:
        out = torch.empty((N, OD, OH, OW, Co), device=x.device, dtype=x.dtype)
    d.act = act
    d.x, d.w, d.out = x.data_ptr(), w16.data_ptr(), out.data_ptr()
    if bias is not None:
        d.bias = _f32(bias, "bias").data_ptr()
    if residual is not None:
        d.residual = _act(residual, "residual").data_ptr()
    L.check(L.lib().vb200_conv3d_igemm(C.byref(d), L.stream_ptr()), "vb200_conv3d_igemm")
    return out


def conv3d_igemm_wgrad(x, dout, kernel, padding, k_splits=0, stride=(1, 1, 1)):
    """dw[co, (kd,kh,kw,ci)] (fp32) = sum over output voxels of dout[v, co] * x[v * stride + tap - pad, ci]."""
    _act(x, "x")
    _act(dout, "dout")
    N, D, H, W, Ci = x.shape
    Co = dout.shape[-1]
    kd, kh, kw = kernel
    d = _conv3d_desc((N, D, H, W, Ci), Co, kernel, padding, L.dtype_code(x.dtype), stride)
    if tuple(dout.shape[1:4]) != _conv3d_out((D, H, W), kernel, padding, stride):
        raise ValueError(f"dout extent {tuple(dout.shape[1:4])} does not match the conv geometry")
    d.k_splits = k_splits
    dw = zeros((Co, kd * kh * kw * Ci), x.device)
    d.x, d.dout, d.dw = x.data_ptr(), dout.data_ptr(), dw.data_ptr()
    L.check(L.lib().vb200_conv3d_igemm_wgrad(C.byref(d), L.stream_ptr()), "vb200_conv3d_igemm_wgrad")
    return dw


def conv3d_wgrad_kh3_supported(shape, cout, kernel, padding, stride=(1, 1, 1)) -> bool:
    fn = L.lib().vb200_conv3d_wgrad_kh3_supported
    return bool(fn(C.byref(_conv3d_desc(shape, cout, kernel, padding, stride=stride))))


def conv3d_wgrad_kh3(x, dout, kernel, padding, k_splits=0):
    """Few-channel stride-1 weight gradient (8 x 8 voxel patches, kh taps as views of one haloed box); same result
    layout as `conv3d_igemm_wgrad`: fp32 [Cout, (kd,kh,kw,Cin)]."""
    _act(x, "x")
    _act(dout, "dout")
    N, D, H, W, Ci = x.shape
    Co = dout.shape[-1]
    kd, kh, kw = kernel
    d = _conv3d_desc((N, D, H, W, Ci), Co, kernel, padding, L.dtype_code(x.dtype))
    if tuple(dout.shape[1:4]) != _conv3d_out((D, H, W), kernel, padding, (1, 1, 1)):
        raise ValueError(f"dout extent {tuple(dout.shape[1:4])} does not match the conv geometry")
    d.k_splits = k_splits
    dw = zeros((Co, kd * kh * kw * Ci), x.device)
    d.x, d.dout, d.dw = x.data_ptr(), dout.data_ptr(), dw.data_ptr()
    L.check(L.lib().vb200_conv3d_wgrad_kh3(C.byref(d), L.stream_ptr()), "vb200_conv3d_wgrad_kh3")
    return dw


def im2col3d(u, geom):
    _act(u, "u")
    g = (C.c_int32 * 17)(*geom)
    N, OD, OH, OW = geom[0], geom[14], geom[15], geom[16]
    K = geom[5] * geom[6] * geom[7] * geom[4]
    col = torch.empty((N * OD * OH * OW, K), device=u.device, dtype=u.dtype)
    _call("vb200_im2col3d", _p(u), _p(col), g)
    return col


def col2im3d(dcol, geom):
    _act(dcol, "dcol")
    g = (C.c_int32 * 17)(*geom)
    du = torch.empty(tuple(geom[0:5]), device=dcol.device, dtype=dcol.dtype)
    _call("vb200_col2im3d", _p(dcol), _p(du), g, L.dtype_code(dcol.dtype))
    return du


# ----------------------------------------------------------------------------------------------
# FCMAE pieces
def rows_select(src, rowmap, n_dst=None, base=None):
    """dst[r] = (rowmap[r] >= 0 ? src[rowmap[r]] : 0) + (base[r] if base is given); src [S, C] / rowmap int32 [n_dst]."""
    _act(src, "src")
    Cc = src.shape[-1]
    n_dst = rowmap.numel() if n_dst is None else n_dst
    if rowmap.dtype != torch.int32 or not rowmap.is_contiguous():
        raise ValueError("rowmap must be contiguous int32")
    if base is not None:
        _act(base, "base")
    dst = torch.empty((n_dst, Cc), device=src.device, dtype=src.dtype)
    _call("vb200_rows_select", _p(src), _p(rowmap), _p(base), _p(dst), C.c_int64(n_dst), Cc, L.dtype_code(src.dtype))
    return dst


def shuffle_pool_fwd(dec, r, pool):
    """dec [B,h,w,Cq*r*r] -> [B,Cq,h*r,w*r] (pixel shuffle + front pad + r x r average pool, stride 1)"""
    _act(dec, "dec")
    B, h, w, Cd = dec.shape
    Cq = Cd // (r * r)
    out = torch.empty((B, Cq, h * r, w * r), device=dec.device, dtype=dec.dtype)
    _call("vb200_shuffle_pool", _p(dec), _p(out), B, h, w, Cq, r, int(pool), 0, L.dtype_code(dec.dtype))
    return out


def shuffle_pool_bwd(dout, r, pool):
    _act(dout, "dout")
    B, Cq, Hs, Ws = dout.shape
    h, w = Hs // r, Ws // r
    ddec = torch.empty((B, h, w, Cq * r * r), device=dout.device, dtype=dout.dtype)
    _call("vb200_shuffle_pool", _p(dout), _p(ddec), B, h, w, Cq, r, int(pool), 1, L.dtype_code(dout.dtype))
    return ddec


# ----------------------------------------------------------------------------------------------
# PixelToVoxelHead pieces
def head_shuffle_pool_fwd(dec, Dz, pool, Cu):
    _act(dec, "dec")
    B, h, w, Cd = dec.shape
    Cm = Cd // 4
    u = torch.empty((B, Dz, 2 * h, 2 * w, Cu), device=dec.device, dtype=dec.dtype)
    _call("vb200_head_shuffle_pool", _p(dec), _p(u), B, h, w, Cm, Dz, Cu, int(pool), 0, L.dtype_code(dec.dtype))
    return u


def head_shuffle_pool_bwd(du, Cm, pool):
    _act(du, "du")
    B, Dz, H2, W2, Cu = du.shape
    h, w = H2 // 2, W2 // 2
    ddec = torch.empty((B, h, w, 4 * Cm), device=du.device, dtype=du.dtype)
    _call("vb200_head_shuffle_pool", _p(du), _p(ddec), B, h, w, Cm, Dz, Cu, int(pool), 1, L.dtype_code(du.dtype))
    return ddec


def instnorm_stats(z, eps=1e-5):
    """z [B, R, C] -> mean, rstd [B, C]"""
    _act(z, "z")
    B, R, Cc = z.shape
    scratch = torch.empty((4, B, Cc), device=z.device, dtype=torch.float32)
    _call("vb200_instnorm_stats", _p(z), _p(scratch[0]), _p(scratch[1]), _p(scratch[2]), _p(scratch[3]), B,
          C.c_int64(R), Cc, C.c_float(eps), L.dtype_code(z.dtype))
    return scratch[2], scratch[3]


def head_tail_fwd(z, mean, rstd, alpha, W1, b1, Dz, H, W):
    B, R, Cmid = z.shape
    Co4 = W1.shape[0]
    out = torch.empty((B, Co4 // 4, Dz, 2 * H, 2 * W), device=z.device, dtype=z.dtype)
    _call("vb200_head_tail_fwd", _p(_act(z, "z")), _p(mean), _p(rstd), _p(_f32(alpha, "alpha")), alpha.numel(),
          _p(_f32(W1, "W1")), _p(_f32(b1, "b1")), _p(out), B, Dz, H, W, Cmid, Co4, L.dtype_code(z.dtype))
    return out


HEAD_BWD_STREAM = True  # False: the generic two-phase kernels for every geometry (A/B timing, tests)


def head_tail_bwd(z, mean, rstd, alpha, W1, dout, Dz, H, W):
    """-> dz, dW1, db1, dalpha, dbz (bias gradient of the conv that produced z)"""
    B, R, Cmid = z.shape
    Co4 = W1.shape[0]
    ldt = -(-Co4 // 8) * 8
    dev = z.device
    dout = _act(dout, "dout")
    if HEAD_BWD_STREAM and Cmid == 32 and Co4 == 8 and W >= 4 and 256 % W == 0 and R % 256 == 0:
        # streaming form: nothing materialised, dW1 / db1 / dalpha / per-sample sums formed in registers
        ar = Arena(dev, [2 * B * Cmid, Co4, alpha.numel(), Cmid, Co4 * Cmid])
        sdp, sdpx = ar.take(B, Cmid), ar.take(B, Cmid)
        db1, dalpha, dbz, dW1 = ar.take(Co4), ar.take(alpha.numel()), ar.take(Cmid), ar.take(Co4, Cmid)
        dz = torch.empty_like(z)
        for phase in (0, 1):
            _call("vb200_head_tail_bwd_stream", phase, _p(z), _p(mean), _p(rstd), _p(alpha), alpha.numel(), _p(W1), _p(dout),
                  _p(sdp), _p(sdpx), _p(db1), _p(dalpha), _p(dW1), _p(dz), _p(dbz), B, Dz, H, W, Cmid, Co4,
                  L.dtype_code(z.dtype))
        return dz, dW1, db1, dalpha, dbz
    ar = Arena(dev, [2 * B * Cmid, Co4, alpha.numel(), Cmid])
    sdp, sdpx = ar.take(B, Cmid), ar.take(B, Cmid)
    db1, dalpha, dbz = ar.take(Co4), ar.take(alpha.numel()), ar.take(Cmid)
    act = torch.empty_like(z)
    dt_rows = torch.empty((B, R, ldt), device=dev, dtype=z.dtype)
    dz = torch.empty_like(z)
    dt = L.dtype_code(z.dtype)
    for phase in (0, 1):
        _call("vb200_head_tail_bwd", phase, _p(z), _p(mean), _p(rstd), _p(alpha), alpha.numel(), _p(W1), _p(dout),
              _p(sdp), _p(sdpx), _p(db1), _p(dalpha), _p(act), _p(dt_rows), _p(dz), _p(dbz), B, Dz, H, W, Cmid, Co4, dt)
        if phase == 0:
            # dW1[o, c] = sum_rows dt[row, o] * act[row, c] on the tensor cores (MN-major wgrad GEMM, K-split)
            dW1 = gemm(dt_rows.view(B * R, ldt), act.view(B * R, Cmid), mn_major=True, epilogue=L.EPI_F32,
                       k_splits=min(148, max(1, (B * R) // 4096)))[:Co4]
    return dz, dW1, db1, dalpha, dbz


# ----------------------------------------------------------------------------------------------
# fused GRN path helpers
def colreduce(x, mode, arena=None, width=None, pivot=None):
    """x [B,R,C] 16-bit -> fp32 [B,C]: mode 0 column sums, mode 1 column sums of squares; mode 2: both in one pass,
    fp32 [2,B,C].  width: reduce only the first `width` columns of rows whose pitch is x.shape[-1].  pivot: fp32 [C]
    subtracted from every element before summing."""
    _act(x, "x")
    B, R, ld = x.shape
    Cc = ld if width is None else width
    shape = (2, B, Cc) if mode == 2 else (B, Cc)
    out = arena.take(*shape) if arena is not None else zeros(shape, x.device)
    _call("vb200_colreduce_ld", _p(x), _p(out), B, C.c_int64(R), Cc, C.c_int64(ld), mode,
          _p(None if pivot is None else _f32(pivot, "pivot")), L.dtype_code(x.dtype))
    return out


def grn_pack_w2(w2, s, dtype):
    """w2 fp32 [C,C4], s fp32 [nb,C4] -> 16-bit [nb*C, C4]: per-sample GRN-scaled fc2 weights."""
    Cc, C4 = w2.shape
    nb = s.shape[0]
    out = torch.empty((nb * Cc, C4), device=w2.device, dtype=dtype)
    _call("vb200_grn_pack_w2", _p(_f32(w2, "w2")), _p(_f32(s, "s")), _p(out), nb, Cc, C4, L.dtype_code(dtype))
    return out


def grn_prepare(sumsq, gw, gb, w2, b2, dtype, eps=1e-6):
    """-> (s [nb,C4] fp32, w2s [nb*C, C4] 16-bit, b2eff [C] fp32): GRN coefficients (one short launch, grid = samples x
    column chunks), then per-sample scaled fc2 weights + effective bias in one launch.  (The single-launch variant
    `vb200_grn_prepare2` recomputes the coefficients in every block: measured 2-4x slower, kept for reference.)"""
    nb, C4 = sumsq.shape
    Cc = w2.shape[0]
    s = torch.empty_like(sumsq)
    _call("vb200_grn_coef_fwd", _p(_f32(sumsq, "sumsq")), _p(_f32(gw, "grn.weight")), _p(s), nb, C4, C.c_float(eps))
    w2s = torch.empty((nb * Cc, C4), device=w2.device, dtype=dtype)
    b2e = torch.empty((Cc,), device=w2.device, dtype=torch.float32)
    _call("vb200_grn_prepare", _p(s), _p(_f32(gb, "grn.bias")), _p(_f32(w2, "w2")), _p(_f32(b2, "b2")), _p(w2s), _p(b2e),
          nb, Cc, C4, L.dtype_code(dtype))
    return s, w2s, b2e


def layerscale_bwd(G, w2, gamma, b2, db_raw, dtype, arena=None):
    """ConvNeXt-V1 layer-scale backward pieces in one launch.  G fp32 [C, >= C4] (row pitch = its stride) ->
    (w2t16 [C4, C], sv [C4], dgamma [C], dW2 [C, C4], db2 [C])."""
    Cc, C4 = w2.shape
    dev = w2.device
    w2t = torch.empty((C4, Cc), device=dev, dtype=dtype)
    dgamma = arena.take(Cc) if arena is not None else zeros((Cc,), dev)
    dW2 = torch.empty((Cc, C4), device=dev, dtype=torch.float32)
    db2 = torch.empty((Cc,), device=dev, dtype=torch.float32)
    sv = torch.empty((C4,), device=dev, dtype=torch.float32)
    if G.dtype != torch.float32 or G.stride(1) != 1:
        raise ValueError("G must be fp32 with unit inner stride")
    _call("vb200_layerscale_bwd", _p(G), C.c_int64(G.stride(0)), _p(_f32(w2, "w2")), _p(_f32(gamma, "gamma")), _p(_f32(b2, "b2")),
          _p(_f32(db_raw, "db_raw")), _p(w2t), _p(dgamma), _p(dW2), _p(db2), _p(sv), Cc, C4, L.dtype_code(dtype))
    return w2t, sv, dgamma, dW2, db2


def grn_bias_eff(w2, bgrn, b2):
    Cc, C4 = w2.shape
    out = torch.empty((Cc,), device=w2.device, dtype=torch.float32)
    _call("vb200_grn_bias_eff", _p(_f32(w2, "w2")), _p(_f32(bgrn, "bgrn")), _p(_f32(b2, "b2")), _p(out), Cc, C4)
    return out


def grn_wgrad_finish(P, w2, s, bgrn, db2, arena=None):
    """P fp32 [nb,C,ldp] -> (dW2 [C,C4], S1 [nb,C4], dbgrn [C4], db2 [C]).  db2=None: P carries the ones column
    (ldp > C4, column C4 = per-sample sums of dout) and the fc2 bias gradient is taken from it."""
    nb, Cc, ldp = P.shape
    C4 = w2.shape[1]
    dW2 = torch.empty((Cc, C4), device=P.device, dtype=torch.float32)
    acc = arena.take(nb + 1, C4) if arena is not None else zeros((nb + 1, C4), P.device)
    db2_out = None
    if db2 is None:
        db2_out = torch.empty((Cc,), device=P.device, dtype=torch.float32)
    _call("vb200_grn_wgrad_finish_ld", _p(P), C.c_int64(ldp), _p(_f32(w2, "w2")), _p(_f32(s, "s")), _p(_f32(bgrn, "bgrn")),
          _p(db2), _p(dW2), _p(acc[:nb]), _p(acc[nb]), _p(db2_out), nb, Cc, C4)
    return dW2, acc[:nb], acc[nb], (db2 if db2 is not None else db2_out)


# ----------------------------------------------------------------------------------------------
# implicit-GEMM conv3d (k=3, stride 1, small channel counts)
def conv3d_pack_weights(w, dtype, cin_pad, cout_pad, transpose_flip=False):
    """w fp32 [Co, Ci, 3,3,3] -> 16-bit [KCH][cout_pad][8] in K order (kd, kh, channel chunk, kw).

    transpose_flip=True builds the data-gradient weights W'[ci, co, kd, kh, kw] = W[co, ci, 2-kd, 2-kh, 2-kw]."""
    w = w.detach()
    if transpose_flip:
        w = w.flip(2, 3, 4).transpose(0, 1)
    Co, Ci = w.shape[:2]
    wp = w.new_zeros((cout_pad, cin_pad, 3, 3, 3))
    wp[:Co, :Ci] = w
    cch = cin_pad // 8
    # [co, chunk, j, kd, kh, kw] -> [kd, kh, chunk, kw, co, j]
    wp = wp.view(cout_pad, cch, 8, 3, 3, 3).permute(3, 4, 1, 5, 0, 2).reshape(27 * cch, cout_pad, 8)
    if (27 * cch) % 2:
        wp = torch.cat([wp, wp.new_zeros((1, cout_pad, 8))], 0)
    return wp.contiguous().to(dtype)


def conv3d_k3(u, wpack, bias, padding, cout_pad, co_store):
    """u [N,D,H,W,cin] 16-bit channels-last -> [N,OD,OH,OW,co_store]"""
    _act(u, "u")
    N, D, H, W, cin = u.shape
    pd, ph, pw = padding
    out = torch.empty((N, D + 2 * pd - 2, H + 2 * ph - 2, W + 2 * pw - 2, co_store), device=u.device, dtype=u.dtype)
    g = (C.c_int32 * 7)(N, D, H, W, pd, ph, pw)
    _call("vb200_conv3d_k3", _p(u), _p(wpack), _p(bias), _p(out), g, cin, cout_pad, co_store, L.dtype_code(u.dtype))
    return out


def conv3d_k3_wgrad(u, dz, padding):
    """-> dW fp32 [Co, 8 (ci), 3,3,3] for cin == 8"""
    N, D, H, W, cin = u.shape
    Co = dz.shape[-1]
    pd, ph, pw = padding
    dw = zeros((Co, 9, 3, 8), u.device)
    g = (C.c_int32 * 7)(N, D, H, W, pd, ph, pw)
    _call("vb200_conv3d_k3_wgrad", _p(_act(u, "u")), _p(_act(dz, "dz")), _p(dw), g, cin, Co, L.dtype_code(u.dtype))
    # [co, (kd,kh), kw, ci] -> [co, ci, kd, kh, kw]
    return dw.view(Co, 3, 3, 3, 8).permute(0, 4, 1, 2, 3)


# ----------------------------------------------------------------------------------------------
# ContrastiveEncoder pooled head / projection MLP
def bn_rows_fwd(x, gamma, beta, run_mean, run_var, eps, training, relu):
    """x [B,C] 16-bit -> (y, mean [C], rstd [C], unbiased batch var [C] or None)"""
    _act(x, "x")
    B, Cc = x.shape
    y = torch.empty_like(x)
    st = torch.empty((3, Cc), device=x.device, dtype=torch.float32)
    _call("vb200_bn_rows_fwd", _p(x), _p(_f32(gamma, "weight")), _p(_f32(beta, "bias")), _p(run_mean), _p(run_var), _p(y),
          _p(st[0]), _p(st[1]), _p(st[2]) if training else _p(None), B, Cc, C.c_float(eps), int(training), int(relu),
          L.dtype_code(x.dtype))
    return y, st[0], st[1], (st[2] if training else None)


def bn_rows_bwd(dy, x, y, gamma, mean, rstd, training, relu):
    B, Cc = x.shape
    dx = torch.empty_like(x)
    g = torch.empty((2, Cc), device=x.device, dtype=torch.float32)
    _call("vb200_bn_rows_bwd", _p(_act(dy, "dy")), _p(x), _p(y), _p(_f32(gamma, "weight")), _p(mean), _p(rstd), _p(dx),
          _p(g[0]), _p(g[1]), B, Cc, int(training), int(relu), L.dtype_code(x.dtype))
    return dx, g[0], g[1]


def bcast_rows(src, R, scale):
    """src [B,C] 16-bit -> [B,R,C] = src * scale"""
    _act(src, "src")
    B, Cc = src.shape
    out = torch.empty((B, R, Cc), device=src.device, dtype=src.dtype)
    _call("vb200_bcast_rows", _p(src), _p(out), B, R, Cc, C.c_float(scale), L.dtype_code(src.dtype))
    return out


# ----------------------------------------------------------------------------------------------
# 3-D U-Net family helpers (channels-last rows)
def to_channels_last_3d(x, dtype, cpad):
    """x (N,C,D,H,W) fp32/16-bit -> [N,D,H,W,cpad] 16-bit (zero channel padding)."""
    if not x.is_contiguous():
        x = x.contiguous()
    N, Cc, D, H, W = x.shape
    y = torch.empty((N, D, H, W, cpad), device=x.device, dtype=dtype)
    _call("vb200_to_channels_last", _p(x), _XDT[x.dtype], _p(y), C.c_int64(N), Cc, cpad, C.c_int64(D * H * W),
          L.dtype_code(dtype))
    return y


def from_channels_last_3d(y, c):
    """[N,D,H,W,cpad] 16-bit -> (N,c,D,H,W) 16-bit"""
    _act(y, "y")
    N, D, H, W, cpad = y.shape
    x = torch.empty((N, c, D, H, W), device=y.device, dtype=y.dtype)
    _call("vb200_from_channels_last", _p(y), _p(x), C.c_int64(N), c, cpad, C.c_int64(D * H * W))
    return x


def affine_act(x, scale, shift, relu):
    _act(x, "x")
    Cc = x.shape[-1]
    y = torch.empty_like(x)
    _call("vb200_affine_act", _p(x), _p(_f32(scale, "scale")), _p(_f32(shift, "shift")), _p(y),
          C.c_int64(x.numel() // Cc), Cc, int(relu), L.dtype_code(x.dtype))
    return y


def bn_bwd(dy, x, y, mean, rstd, gamma, relu, training):
    """BatchNorm (+ReLU) backward on rows [..., C] -> (dx, dgamma, dbeta)"""
    Cc = x.shape[-1]
    M = x.numel() // Cc
    dt = L.dtype_code(x.dtype)
    s = zeros((2, Cc), x.device)
    _call("vb200_bn_bwd_reduce", _p(_act(dy, "dy")), _p(x), _p(y), _p(mean), _p(rstd), _p(s[0]), _p(s[1]),
          C.c_int64(M), Cc, int(relu), dt)
    dx = torch.empty_like(x)
    _call("vb200_bn_bwd_apply_raw", _p(dy), _p(x), _p(y), _p(mean), _p(rstd), _p(_f32(gamma.detach(), "gamma")), _p(s[0]), _p(s[1]),
          C.c_float(1.0 / M if training else 0.0), _p(dx), C.c_int64(M), Cc, int(relu), dt)
    return dx, s[1], s[0]


def bn_finalize(sums, pivot, weight, bias, run_mean, run_var, Cc, M, eps, momentum):
    """BatchNorm statistics -> (scale, shift, mean, rstd) fp32 [Cc] in one launch; updates run_mean / run_var in place when
    momentum >= 0 (training).  sums=None: eval mode (running statistics).  weight / bias / run_* may be shorter than Cc
    (channel-padded rows)."""
    dev = weight.device
    out = torch.empty((4, Cc), device=dev, dtype=torch.float32)
    _call("vb200_bn_finalize", _p(sums), _p(pivot), _p(_f32(weight.detach(), "weight")), _p(_f32(bias.detach(), "bias")),
          _p(run_mean), _p(run_var), weight.numel(), Cc, C.c_double(float(M)), C.c_float(eps), C.c_float(momentum),
          _p(out[0]), _p(out[1]), _p(out[2]), _p(out[3]))
    return out[0], out[1], out[2], out[3]


ACT_CODES = {"none": 0, "linear": 0, "relu": 1, "silu": 2, "leakyrelu": 3, "elu": 4, "selu": 5}


def affine_nc_act(x, a, b, act: str):
    """y = act(a[n,c] * x + b[n,c]) on rows [N, ..., C] (a, b fp32 [N, C])."""
    _act(x, "x")
    N, Cc = x.shape[0], x.shape[-1]
    R = x.numel() // (N * Cc)
    y = torch.empty_like(x)
    _call("vb200_affine_nc_act", _p(x), _p(_f32(a, "a")), _p(_f32(b, "b")), _p(y), C.c_int64(N), C.c_int64(R), Cc,
          ACT_CODES[act], L.dtype_code(x.dtype))
    return y


def gn_bwd_reduce(dy, x, a, b, act: str):
    """-> (S1, S2) fp32 [N, C]: sums over rows of dv and dv * x, dv = dy * act'(a x + b)."""
    N, Cc = x.shape[0], x.shape[-1]
    R = x.numel() // (N * Cc)
    s = zeros((2, N, Cc), x.device)
    _call("vb200_gn_bwd_reduce", _p(_act(dy, "dy")), _p(_act(x, "x")), _p(_f32(a, "a")), _p(_f32(b, "b")), _p(s[0]), _p(s[1]),
          C.c_int64(N), C.c_int64(R), Cc, ACT_CODES[act], L.dtype_code(x.dtype))
    return s[0], s[1]


def gn_bwd_apply(dy, x, coef, act: str):
    """dx = c1 * dv + c2 * x + c3; coef fp32 [5, N, C] = (a, b, c1, c2, c3)."""
    N, Cc = x.shape[0], x.shape[-1]
    R = x.numel() // (N * Cc)
    dx = torch.empty_like(x)
    _call("vb200_gn_bwd_apply", _p(_act(dy, "dy")), _p(_act(x, "x")), _p(_f32(coef, "coef")), _p(dx), C.c_int64(N),
          C.c_int64(R), Cc, ACT_CODES[act], L.dtype_code(x.dtype))
    return dx


def cat2(a, b):
    _act(a, "a"), _act(b, "b")
    Ca, Cb = a.shape[-1], b.shape[-1]
    out = torch.empty((*a.shape[:-1], Ca + Cb), device=a.device, dtype=a.dtype)
    _call("vb200_cat2", _p(a), _p(b), _p(out), C.c_int64(a.numel() // Ca), Ca, Cb, 0)
    return out


def split2(dout, Ca, Cb):
    _act(dout, "dout")
    M = dout.numel() // (Ca + Cb)
    a = torch.empty((*dout.shape[:-1], Ca), device=dout.device, dtype=dout.dtype)
    b = torch.empty((*dout.shape[:-1], Cb), device=dout.device, dtype=dout.dtype)
    _call("vb200_cat2", _p(a), _p(b), _p(dout), C.c_int64(M), Ca, Cb, 1)
    return a, b


def add_rows(x, other=None, bias=None):
    _act(x, "x")
    Cc = x.shape[-1]
    y = torch.empty_like(x)
    _call("vb200_add_rows", _p(x), _p(other), _p(bias), _p(y), C.c_int64(x.numel() // Cc), Cc, L.dtype_code(x.dtype))
    return y


# ----------------------------------------------------------------------------------------------
# 2.5-D U-Net helpers (channels-last rows [N,D,H,W,C])
def scale_relu(x, scale, relu, gate=None):
    """forward: relu?(x * scale[n,c]); backward (gate = forward output): x * scale[n,c] * (gate > 0).  scale [N,C] fp32 or None."""
    _act(x, "x")
    N, Cc = x.shape[0], x.shape[-1]
    rows = x.numel() // (N * Cc)
    y = torch.empty_like(x)
    _call("vb200_scale_relu", _p(x), _p(gate), _p(None if scale is None else _f32(scale, "scale")), _p(y), C.c_int64(N),
          C.c_int64(rows), Cc, int(relu), L.dtype_code(x.dtype))
    return y


def scale_rows(x, scale):
    """x [N, ..., C] 16-bit times scale[n] (fp32 [N]): stochastic-depth scaling of a gradient."""
    N, Cc = x.shape[0], x.shape[-1]
    return scale_relu(x, scale.view(N, 1).expand(N, Cc).contiguous(), False)


def avgpool_hw2(x, backward_shape=None):
    """[N,D,H,W,C] -> [N,D,H//2,W//2,C] (AvgPool3d (1,2,2)); backward_shape = input shape: the gradient of that."""
    _act(x, "x")
    if backward_shape is None:
        N, D, H, W, Cc = x.shape
        y = torch.empty((N, D, H // 2, W // 2, Cc), device=x.device, dtype=x.dtype)
        _call("vb200_avgpool_hw2", _p(x), _p(y), C.c_int64(N * D), H, W, Cc, 0, L.dtype_code(x.dtype))
        return y
    N, D, H, W, Cc = backward_shape
    dx = torch.empty(backward_shape, device=x.device, dtype=x.dtype)
    _call("vb200_avgpool_hw2", _p(x), _p(dx), C.c_int64(N * D), H, W, Cc, 1, L.dtype_code(x.dtype))
    return dx


def upsample2x_hw(x, backward=False):
    """bilinear x2 in H and W (trilinear, scale (1,2,2), align_corners=False) on [N,D,H,W,C]; backward: the adjoint."""
    _act(x, "x")
    N, D, H, W, Cc = x.shape
    if not backward:
        y = torch.empty((N, D, 2 * H, 2 * W, Cc), device=x.device, dtype=x.dtype)
        _call("vb200_upsample2x_hw", _p(x), _p(y), C.c_int64(N * D), H, W, Cc, 0, L.dtype_code(x.dtype))
        return y
    dx = torch.empty((N, D, H // 2, W // 2, Cc), device=x.device, dtype=x.dtype)
    _call("vb200_upsample2x_hw", _p(x), _p(dx), C.c_int64(N * D), H // 2, W // 2, Cc, 1, L.dtype_code(x.dtype))
    return dx


# ----------------------------------------------------------------------------------------------
# per-step weight packing in one launch
class WeightPacks:
    """16-bit GEMM operand copies ([N,K] and [K,N]) of Linear / 1x1-conv weights and tap-major depthwise filters of a
    set of modules, refreshed by ONE kernel launch per step (`refresh`), looked up by parameter (`get`)."""

    def __init__(self, linear_weights, dw_weights, dtype):
        self.dtype = dtype
        dev = linear_weights[0].device if linear_weights else dw_weights[0].device
        n16 = sum(2 * w.numel() + 16 for w in linear_weights)  # both copies, each 16-byte aligned
        n32 = sum(2 * 49 * w.shape[0] for w in dw_weights)
        self.buf16 = torch.empty((n16,), device=dev, dtype=dtype)
        self.buf32 = torch.empty((n32,), device=dev, dtype=torch.float32)
        self.views: dict[tuple[int, str], torch.Tensor] = {}
        import weakref
        self._refs = [(weakref.ref(w), w.data_ptr()) for w in list(linear_weights) + list(dw_weights)]
        self._alive = {ptr: ref for ref, ptr in self._refs}
        rows, off16, off32, blk = [], 0, 0, 0
        per = 1024
        for w in linear_weights:
            # one 64 x 64 tile per block, read once, written as both operand copies ([R,Cc] and [Cc,R])
            R, Cc = w.shape[0], w.numel() // w.shape[0]
            sz = -(-R * Cc // 8) * 8
            vn = self.buf16[off16:off16 + R * Cc].view(R, Cc)
            vt = self.buf16[off16 + sz:off16 + sz + R * Cc].view(Cc, R)
            off16 += 2 * sz
            self.views[(w.data_ptr(), "n")] = vn
            self.views[(w.data_ptr(), "t")] = vt
            rows.append([w.data_ptr(), vn.data_ptr(), vt.data_ptr(), R, Cc, 3, blk])
            blk += -(-R // 64) * -(-Cc // 64)
        for w in dw_weights:
            Cc = w.shape[0]
            a = self.buf32[off32:off32 + 49 * Cc].view(49, Cc)
            b = self.buf32[off32 + 49 * Cc:off32 + 98 * Cc].view(49, Cc)
            off32 += 98 * Cc
            self.views[(w.data_ptr(), "dw")] = a
            self.views[(w.data_ptr(), "dwf")] = b
            rows.append([w.data_ptr(), a.data_ptr(), b.data_ptr(), Cc, 49, 2, blk])
            blk += -(-49 * Cc // per)
        self.total_blocks = blk
        self.table = torch.tensor(rows, dtype=torch.int64, device=dev)
        self.n_items = len(rows)

    def stale(self) -> bool:
        """True when a registered parameter has been freed or re-allocated (the table holds raw pointers)."""
        for ref, ptr in self._refs:
            p = ref()
            if p is None or p.data_ptr() != ptr:
                return True
        return False

    def refresh(self) -> None:
        _call("vb200_pack_multi", _p(self.table), self.n_items, C.c_int64(self.total_blocks), L.dtype_code(self.dtype))

    def get(self, w, kind):
        ptr = w.data_ptr()
        ref = self._alive.get(ptr)
        if ref is None:
            return None
        owner = ref()
        if owner is None or owner.data_ptr() != ptr:  # the registered parameter is gone: never serve its pack
            return None
        return self.views.get((ptr, kind))


ACTIVE_PACKS: WeightPacks | None = None


def packed(w, dtype, transpose=False):
    """16-bit operand copy of fp32 weight `w`: from the step's WeightPacks when registered, else a direct cast."""
    if ACTIVE_PACKS is not None and ACTIVE_PACKS.dtype == dtype:
        v = ACTIVE_PACKS.get(w, "t" if transpose else "n")
        if v is not None:
            return v
    return cast_pack(w, dtype, transpose)


def dw_taps(w):
    if ACTIVE_PACKS is not None:
        a = ACTIVE_PACKS.get(w, "dw")
        if a is not None:
            return a, ACTIVE_PACKS.get(w, "dwf")
    return dw_pack(w)
