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
    if K != K2:
        raise ValueError(f"reduction dims differ: {K} vs {K2}")
    d = L.GemmDesc()
    d.M, d.N, d.K = M, N, K
    d.dtype = L.dtype_code(a.dtype)
    d.mn_major = 1 if mn_major else 0
    d.epilogue = epilogue
    d.act = act
    d.k_splits = max(1, k_splits)
    if epilogue == L.EPI_F32:
        if out is None:
            out = (torch.zeros if (k_splits > 1 or accumulate) else torch.empty)(
                (M, N), device=a.device, dtype=torch.float32
            )
        d.atomic_out = 1 if (k_splits > 1 or accumulate) else 0
    else:
        if out is None:
            out = torch.empty((M, N), device=a.device, dtype=a.dtype)
    d.lda, d.ldb = lda, ldb
    d.ldo = _rowmajor2d(out, "out")
    d.A, d.B, d.out = a.data_ptr(), b.data_ptr(), out.data_ptr()
    if epilogue == L.EPI_GELU_DUAL:
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
    L.check(L.lib().vb200_gemm(C.byref(d), L.stream_ptr()), "vb200_gemm")
    if epilogue == L.EPI_GELU_DUAL:
        return out, out2
    return out
