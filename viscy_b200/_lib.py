"""ctypes binding of libviscy_b200.so (C ABI declared in include/viscy_b200.h).

The product path never falls back: if the library is missing, `lib()` raises.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

import os

_PKG = Path(__file__).resolve().parent
# VB200_LIB: load another build of the same library (A/B timing of kernel variants); never a different implementation
LIB_PATH = Path(os.environ["VB200_LIB"]) if os.environ.get("VB200_LIB") else _PKG / "libviscy_b200.so"

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA = 0, 1, 2, 3
BF16, FP16 = 0, 1
EPI_STORE, EPI_GELU_DUAL, EPI_DGELU, EPI_F32, EPI_DGELU_GRN, EPI_GELU_GP = 0, 1, 2, 3, 4, 5
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
        ("dtype", C.c_int32), ("mn_major", C.c_int32), ("epilogue", C.c_int32),
        ("act", C.c_int32), ("k_splits", C.c_int32), ("atomic_out", C.c_int32),
        ("b_batch_rows", C.c_int32), ("rows_per_sample", C.c_int32),
        ("lda", C.c_int64), ("ldb", C.c_int64),
        ("ldo", C.c_int64), ("ldo2", C.c_int64), ("ldr", C.c_int64), ("ldaux", C.c_int64), ("ldaux2", C.c_int64),
        ("split_out_stride", C.c_int64),
        ("A", C.c_void_p), ("B", C.c_void_p), ("out", C.c_void_p), ("out2", C.c_void_p),
        ("bias", C.c_void_p), ("residual", C.c_void_p), ("aux", C.c_void_p), ("aux2", C.c_void_p),
        ("tvec", C.c_void_p), ("svec", C.c_void_p),
        ("n_split", C.c_int32), ("rvec_rows", C.c_int32), ("rvec", C.c_void_p), ("colsq", C.c_void_p),
    ]


class Conv3dDesc(C.Structure):
    """struct vb200_conv3d_desc (include/viscy_b200.h), field for field."""

    _fields_ = [
        ("N", C.c_int32), ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("cin", C.c_int32), ("cout", C.c_int32),
        ("kd", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
        ("pd", C.c_int32), ("ph", C.c_int32), ("pw", C.c_int32),
        ("sd", C.c_int32), ("sh", C.c_int32), ("sw", C.c_int32),
        ("dtype", C.c_int32), ("act", C.c_int32), ("k_splits", C.c_int32), ("w_taps", C.c_int32),
        ("xd", C.c_int32), ("xh", C.c_int32), ("xw", C.c_int32), ("reserved", C.c_int32),
        ("ldo", C.c_int64), ("ldr", C.c_int64), ("out_pitch", C.c_int64 * 4), ("tapmap", C.POINTER(C.c_int32)),
        ("x", C.c_void_p), ("w", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p), ("out", C.c_void_p),
        ("dout", C.c_void_p), ("dw", C.c_void_p),
    ]


_lib = None


def lib() -> C.CDLL:
    """Load the native library; raise loudly when it is absent (no CPU / library fallback)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} not built: run `python -m viscy_b200.build` "
                "(the sm_100a kernels are mandatory for CUDA tensors; there is no fallback)"
            )
        _lib = C.CDLL(str(LIB_PATH))
        _lib.vb200_last_error.argtypes = [C.c_char_p, C.c_size_t]
        _lib.vb200_launch_count.restype = C.c_int64
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(512)
    lib().vb200_last_error(buf, 512)
    return buf.value.decode()


def check(rc: int, what: str) -> None:
    if rc != OK:
        msg = last_error()
        if rc == ERR_UNSUPPORTED:
            raise NotImplementedError(f"{what}: unsupported on the sm_100a path: {msg}")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_code(dt: torch.dtype) -> int:
    if dt == torch.bfloat16:
        return BF16
    if dt == torch.float16:
        return FP16
    raise NotImplementedError(f"sm_100a kernels take 16-bit activations, got {dt}")


def ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def launch_count() -> int:
    return int(lib().vb200_launch_count())
