"""Thin torch-tensor wrappers over the C ABI: extract pointers / strides / stream, call, check.

torch is used only for device memory and streams; all arithmetic happens in libmodegpt_b200.so.
"""
from __future__ import annotations

import torch

from ._lib import check, lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _rowmajor_2d(x: torch.Tensor, name: str) -> tuple[int, int, int]:
    if x.dim() != 2 or x.stride(1) != 1:
        raise ValueError(f"{name} must be a 2-D tensor with unit column stride")
    return x.shape[0], x.shape[1], x.stride(0)


def _as_rows(x: torch.Tensor) -> torch.Tensor:
    """[..., n] activation -> [T, n] view without copying when the layout allows it."""
    if x.dim() == 2:
        return x if x.stride(1) == 1 else x.contiguous()
    x2 = x.reshape(-1, x.shape[-1])
    return x2 if x2.stride(1) == 1 else x2.contiguous()


def syrk_(C: torch.Tensor, X: torch.Tensor, alpha: float = 1.0, accumulate: bool = True) -> None:
    """C[n,n] (upper triangle) (+)= alpha * X^T X.  X bf16 [T, n]; C fp32."""
    X = _as_rows(X)
    if X.dtype != torch.bfloat16 or C.dtype != torch.float32:
        raise TypeError("syrk_: X must be bfloat16 and C float32")
    T, n, ldx = _rowmajor_2d(X, "X")
    cn, cm, ldc = _rowmajor_2d(C, "C")
    if cn != n or cm != n:
        raise ValueError(f"syrk_: C is {tuple(C.shape)}, expected ({n}, {n})")
    check("mg_syrk_bf16_f32",
          lib.mg_syrk_bf16_f32(X.data_ptr(), T, n, ldx, C.data_ptr(), ldc, alpha,
                               int(accumulate), _stream()))


def syrk_heads_(C: torch.Tensor, X: torch.Tensor, alpha: float = 1.0,
                accumulate: bool = True) -> None:
    """C[H,hd,hd] (+)= alpha * per-head Gram of X[T, H*hd] (bf16)."""
    X = _as_rows(X)
    if X.dtype != torch.bfloat16 or C.dtype != torch.float32:
        raise TypeError("syrk_heads_: X must be bfloat16 and C float32")
    if C.dim() != 3 or C.shape[1] != C.shape[2] or not C.is_contiguous():
        raise ValueError("syrk_heads_: C must be contiguous [H, hd, hd]")
    T, n, ldx = _rowmajor_2d(X, "X")
    H, hd = C.shape[0], C.shape[1]
    if H * hd != n:
        raise ValueError(f"syrk_heads_: X has {n} columns, C describes {H}x{hd}")
    check("mg_syrk_heads_bf16_f32",
          lib.mg_syrk_heads_bf16_f32(X.data_ptr(), T, n, ldx, hd, C.data_ptr(), alpha,
                                     int(accumulate), _stream()))


def bi_cosine_(acc: torch.Tensor, x_in: torch.Tensor, x_out: torch.Tensor) -> None:
    """acc[0] (fp64) += sum over rows of (1 - cos(x_in[row], x_out[row]))."""
    a, b = _as_rows(x_in), _as_rows(x_out)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or acc.dtype != torch.float64:
        raise TypeError("bi_cosine_: inputs must be bfloat16, acc float64")
    if a.shape != b.shape:
        raise ValueError("bi_cosine_: shape mismatch")
    rows, d, lda = _rowmajor_2d(a, "x_in")
    _, _, ldb = _rowmajor_2d(b, "x_out")
    check("mg_bi_cosine_bf16",
          lib.mg_bi_cosine_bf16(a.data_ptr(), lda, b.data_ptr(), ldb, rows, d, acc.data_ptr(),
                                _stream()))


def finalize_sym_(C: torch.Tensor, scale: float) -> None:
    """Scale the upper triangle of C by `scale` and mirror it into the lower triangle."""
    n, m, ldc = _rowmajor_2d(C, "C")
    if n != m or C.dtype != torch.float32:
        raise ValueError("finalize_sym_: C must be a square float32 matrix")
    check("mg_finalize_sym_f32", lib.mg_finalize_sym_f32(C.data_ptr(), n, ldc, scale, _stream()))


def scale_(x: torch.Tensor, scale: float) -> None:
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("scale_: x must be contiguous float32")
    check("mg_scale_f32", lib.mg_scale_f32(x.data_ptr(), x.numel(), scale, _stream()))
