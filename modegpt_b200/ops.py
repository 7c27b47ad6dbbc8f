"""Thin torch-tensor wrappers over the C ABI: extract pointers / strides / stream, call, check.

torch is used only for device memory and streams; all arithmetic happens in libmodegpt_b200.so.
"""
from __future__ import annotations

import os

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


_SCRATCH: dict = {}


def _scratch_square(n: int, device) -> torch.Tensor:
    """One reusable [n, n] fp32 scratch per (device, n, stream): stream-ordered reuse is safe."""
    key = (str(device), n, torch.cuda.current_stream(device).cuda_stream)
    buf = _SCRATCH.get(key)
    if buf is None:
        buf = _SCRATCH[key] = torch.empty(n, n, dtype=torch.float32, device=device)
    return buf


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
    if hd in (32, 64, 128):
        check("mg_syrk_heads_bf16_f32",
              lib.mg_syrk_heads_bf16_f32(X.data_ptr(), T, n, ldx, hd, C.data_ptr(), alpha,
                                         int(accumulate), _stream()))
        return
    # other head dims (OPT-2.7b: 80, 96, ...): Gram of the whole projection, then its diagonal blocks
    full = _scratch_square(n, X.device)
    check("mg_syrk_bf16_f32",
          lib.mg_syrk_bf16_f32(X.data_ptr(), T, n, ldx, full.data_ptr(), full.stride(0), 1.0, 0, _stream()))
    if not accumulate:
        C.zero_()
    check("mg_add_diag_blocks_f32",
          lib.mg_add_diag_blocks_f32(full.data_ptr(), n, full.stride(0), hd, alpha, C.data_ptr(), _stream()))


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


def packed_upper_numel(n: int) -> int:
    return n * (n + 1) // 2


def pack_upper_(packed: torch.Tensor, C: torch.Tensor) -> None:
    """packed[n(n+1)/2] <- the upper triangle of C (row-major rows i..n-1): reduction wire format."""
    n, m, ldc = _rowmajor_2d(C, "C")
    if n != m or C.dtype != torch.float32 or packed.dtype != torch.float32:
        raise ValueError("pack_upper_: C must be a square float32 matrix, packed float32")
    if packed.numel() < packed_upper_numel(n) or not packed.is_contiguous():
        raise ValueError("pack_upper_: packed buffer too small or not contiguous")
    check("mg_pack_upper_f32", lib.mg_pack_upper_f32(C.data_ptr(), n, ldc, packed.data_ptr(), _stream()))


def unpack_upper_(C: torch.Tensor, packed: torch.Tensor) -> None:
    """Upper triangle of C <- packed (inverse of `pack_upper_`; the lower triangle is untouched)."""
    n, m, ldc = _rowmajor_2d(C, "C")
    if n != m or C.dtype != torch.float32 or packed.dtype != torch.float32:
        raise ValueError("unpack_upper_: C must be a square float32 matrix, packed float32")
    if packed.numel() < packed_upper_numel(n) or not packed.is_contiguous():
        raise ValueError("unpack_upper_: packed buffer too small or not contiguous")
    check("mg_unpack_upper_f32", lib.mg_unpack_upper_f32(packed.data_ptr(), n, C.data_ptr(), ldc, _stream()))


# ------------------------------------------------------------------------------------------------
# decompositions
# ------------------------------------------------------------------------------------------------
class NotPositiveDefinite(RuntimeError):
    """A Cholesky pivot was <= 0 (the reference's torch.linalg.cholesky raises here too)."""

    def __init__(self, what: str, pivot: int):
        super().__init__(f"{what}: matrix not positive definite at pivot {pivot}")
        self.pivot = pivot


_POISON_WS = os.environ.get("MG_POISON_WS", "0") == "1"


def _workspace(nbytes: int, device) -> torch.Tensor:
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
    if _POISON_WS:      # tests: every byte 0xFF (NaN as bf16 / fp32 / fp64) — a kernel that reads
        ws.fill_(0xFF)  # workspace it did not write shows up as NaNs instead of passing by luck
    return ws


def _f32_square(C: torch.Tensor, name: str) -> tuple[int, int]:
    n, m, ld = _rowmajor_2d(C, name)
    if n != m or C.dtype != torch.float32:
        raise ValueError(f"{name} must be a square float32 matrix")
    return n, ld


def ridge_scores(C: torch.Tensor, ridge: float, what: str = "ridge_scores") -> torch.Tensor:
    """diag((C + ridge I)^-1) as float32 [n]; C float32 [n, n] (upper triangle is read)."""
    n, ldc = _f32_square(C, "C")
    scores = torch.empty(n, dtype=torch.float32, device=C.device)
    info = torch.zeros(1, dtype=torch.int32, device=C.device)
    nbytes = lib.mg_ridge_scores_ws_bytes(n)
    ws = _workspace(nbytes, C.device)
    check("mg_ridge_scores_f32",
          lib.mg_ridge_scores_f32(C.data_ptr(), n, ldc, ridge, scores.data_ptr(), ws.data_ptr(),
                                  nbytes, info.data_ptr(), _stream()))
    piv = int(info.item())
    if piv:
        raise NotPositiveDefinite(what, piv)
    return scores


def select_k(scores: torch.Tensor, k: int, largest: bool = False) -> torch.Tensor:
    """indices (int64, ascending) of the k smallest / largest entries of a float32 vector."""
    if scores.dtype != torch.float32 or scores.dim() != 1 or not scores.is_contiguous():
        raise ValueError("select_k: scores must be a contiguous float32 vector")
    idx = torch.empty(k, dtype=torch.int64, device=scores.device)
    check("mg_select_k_f32",
          lib.mg_select_k_f32(scores.data_ptr(), scores.numel(), k, int(largest), idx.data_ptr(),
                              _stream()))
    return idx


def gather_rows(W: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """W[idx, :] for a bf16 matrix and int64 device indices."""
    rows, d, ldw = _rowmajor_2d(W, "W")
    if W.dtype != torch.bfloat16 or idx.dtype != torch.int64:
        raise TypeError("gather_rows: W must be bfloat16 and idx int64")
    out = torch.empty(idx.numel(), d, dtype=torch.bfloat16, device=W.device)
    check("mg_gather_rows_bf16",
          lib.mg_gather_rows_bf16(W.data_ptr(), ldw, idx.data_ptr(), idx.numel(), d,
                                  out.data_ptr(), out.stride(0), _stream()))
    return out


# refine automatically when the fp32 solve is expected to miss ~1e-4 (error ~ 1e-7 / min_rel_pivot),
# and always when a sweep is nearly free (small models)
REFINE_BELOW_PIVOT = 2e-3
REFINE_FREE_FLOPS = 5e10


def nystrom_down(C: torch.Tensor, idx: torch.Tensor, Wd: torch.Tensor, jitter: float = 1e-6,
                 refine: bool | str = "auto", stats: dict | None = None) -> torch.Tensor:
    """((C[idx,idx] + jitter I)^-1 C[idx,:] Wd^T)^T as bf16 [d, k]; C full symmetric float32.

    refine: "auto" (default) runs fp64-residual refinement sweeps when the Cholesky factor's
    smallest relative pivot says the fp32 solve is not enough, or when a sweep costs next to
    nothing; True forces one sweep, False none.  `stats` (optional dict) receives the indicator and
    the number of sweeps."""
    n, ldc = _f32_square(C, "C")
    d, n2, ldwd = _rowmajor_2d(Wd, "Wd")
    if n2 != n or Wd.dtype != torch.bfloat16:
        raise ValueError("nystrom_down: Wd must be bfloat16 [d, n]")
    k = idx.numel()
    out = torch.empty(d, k, dtype=torch.bfloat16, device=C.device)
    info = torch.zeros(1, dtype=torch.int32, device=C.device)
    pivot = torch.zeros(1, dtype=torch.float32, device=C.device)
    nbytes = lib.mg_nystrom_down_ws_bytes(n, k, d)
    ws = _workspace(nbytes, C.device)
    check("mg_nystrom_down_f32",
          lib.mg_nystrom_down_f32(C.data_ptr(), n, ldc, idx.data_ptr(), k, Wd.data_ptr(), d, ldwd,
                                  jitter, out.data_ptr(), out.stride(0), ws.data_ptr(), nbytes,
                                  info.data_ptr(), pivot.data_ptr(), _stream()))
    piv = int(info.item())
    if piv:
        raise NotPositiveDefinite("nystrom_down", piv)
    rel_pivot = float(pivot.item())
    if refine == "auto":
        sweeps = 0
        if rel_pivot < REFINE_BELOW_PIVOT or 2.0 * k * n * d < REFINE_FREE_FLOPS:
            sweeps = 2 if rel_pivot < 2e-5 else 1
    else:
        sweeps = 1 if refine else 0
    for _ in range(sweeps):
        check("mg_nystrom_refine_f32",
              lib.mg_nystrom_refine_f32(C.data_ptr(), n, ldc, idx.data_ptr(), k, d, jitter, out.data_ptr(),
                                        out.stride(0), ws.data_ptr(), nbytes, _stream()))
    if stats is not None:
        stats.update(min_rel_pivot=rel_pivot, refine_sweeps=sweeps)
    return out


def qk_select(Cq: torch.Tensor, Ck: torch.Tensor, r: int, mode: int, ridge_q: float,
              ridge_k: float) -> torch.Tensor:
    """Rotary mask [KV, r] (int64).  mode 0 = RoPE-paired (llama/qwen), 1 = OPT."""
    if Cq.dtype != torch.float32 or Ck.dtype != torch.float32:
        raise TypeError("qk_select: statistics must be float32")
    if not (Cq.is_contiguous() and Ck.is_contiguous()) or Cq.dim() != 3 or Ck.dim() != 3:
        raise ValueError("qk_select: Cq/Ck must be contiguous [heads, hd, hd]")
    H, hd = Cq.shape[0], Cq.shape[1]
    KV = Ck.shape[0]
    mask = torch.empty(KV, r, dtype=torch.int64, device=Cq.device)
    check("mg_qk_select_f32",
          lib.mg_qk_select_f32(Cq.data_ptr(), Ck.data_ptr(), H, KV, hd, mode, ridge_q, ridge_k, r,
                               mask.data_ptr(), _stream()))
    return mask


def gather_head_rows(W: torch.Tensor, mask: torch.Tensor, n_heads: int, group: int,
                     hd: int) -> torch.Tensor:
    """out[q*r + t] = W[q*hd + mask[q // group, t]] for every head q."""
    rows, d, ldw = _rowmajor_2d(W, "W")
    if W.dtype != torch.bfloat16 or mask.dtype != torch.int64 or not mask.is_contiguous():
        raise TypeError("gather_head_rows: W bfloat16, mask contiguous int64")
    r = mask.shape[1]
    out = torch.empty(n_heads * r, d, dtype=torch.bfloat16, device=W.device)
    check("mg_gather_head_rows_bf16",
          lib.mg_gather_head_rows_bf16(W.data_ptr(), ldw, mask.data_ptr(), n_heads, group, hd, r,
                                       d, out.data_ptr(), out.stride(0), _stream()))
    return out


VO_FACTOR, VO_GRAM = 0, 1     # MG_VO_FACTOR / MG_VO_GRAM (include/modegpt_b200.h)


def _vo_check(Cx, Wv, Wo, n_heads, n_kv_heads, hd, who):
    d, ldc = _f32_square(Cx, "Cx")
    _, _, ldwv = _rowmajor_2d(Wv, "Wv")
    _, _, ldwo = _rowmajor_2d(Wo, "Wo")
    if Wv.dtype != torch.bfloat16 or Wo.dtype != torch.bfloat16:
        raise TypeError(f"{who}: weights must be bfloat16")
    if Wv.shape != (n_kv_heads * hd, d) or Wo.shape != (d, n_heads * hd):
        raise ValueError(f"{who}: weight shapes do not match the head layout")
    return d, ldc, ldwv, ldwo


def vo_compress(Cx: torch.Tensor, ridge: float, Wv: torch.Tensor, Wo: torch.Tensor, n_heads: int,
                n_kv_heads: int, hd: int, r: int, method: int = VO_GRAM,
                out_dtype: torch.dtype = torch.bfloat16) -> tuple[torch.Tensor, torch.Tensor]:
    """Type-III factors: (v_proj [KV*r, d], o_proj [d, H*r]) in bf16 (or the unrounded fp32 factors
    with out_dtype=torch.float32).  Cx full symmetric.  method VO_GRAM (default): tensor-core
    P = (Cx + ridge I) W_v^T, per-head Grams accumulated in fp64; VO_FACTOR: through the Cholesky
    factor of Cx + ridge I (raises NotPositiveDefinite if that is singular in fp32)."""
    d, ldc, ldwv, ldwo = _vo_check(Cx, Wv, Wo, n_heads, n_kv_heads, hd, "vo_compress")
    if out_dtype not in (torch.bfloat16, torch.float32):
        raise TypeError("vo_compress: out_dtype must be bfloat16 or float32")
    v_out = torch.empty(n_kv_heads * r, d, dtype=out_dtype, device=Cx.device)
    o_out = torch.empty(d, n_heads * r, dtype=out_dtype, device=Cx.device)
    nbytes = lib.mg_vo_ws_bytes(d, n_heads, n_kv_heads, hd)
    ws = _workspace(nbytes, Cx.device)
    info = torch.zeros(1, dtype=torch.int32, device=Cx.device)
    check("mg_vo_compress",
          lib.mg_vo_compress(Cx.data_ptr(), ldc, ridge, Wv.data_ptr(), ldwv, Wo.data_ptr(), ldwo,
                             n_heads, n_kv_heads, hd, d, r, method, v_out.data_ptr(), v_out.stride(0),
                             o_out.data_ptr(), o_out.stride(0), int(out_dtype == torch.float32),
                             info.data_ptr(), ws.data_ptr(), nbytes, _stream()))
    if method == VO_FACTOR:        # the Gram route has no pivots: no read-back, no sync
        piv = int(info.item())
        if piv:
            raise NotPositiveDefinite("vo_compress", piv)
    return v_out, o_out


def vo_prepare(Cx: torch.Tensor, ridge: float, Wv: torch.Tensor, Wo: torch.Tensor, n_heads: int,
               n_kv_heads: int, hd: int, method: int = VO_GRAM
               ) -> tuple[torch.Tensor, torch.Tensor]:
    """First half of `vo_compress` (tensor-core products + fp64 Grams); returns (workspace, info)
    for `vo_finish`.  With method=VO_FACTOR, info[0] != 0 reports a non-positive Cholesky pivot
    (Cx + ridge I singular in fp32); the Gram route never sets it."""
    d, ldc, ldwv, ldwo = _vo_check(Cx, Wv, Wo, n_heads, n_kv_heads, hd, "vo_prepare")
    nbytes = lib.mg_vo_ws_bytes(d, n_heads, n_kv_heads, hd)
    ws = _workspace(nbytes, Cx.device)
    info = torch.zeros(1, dtype=torch.int32, device=Cx.device)
    check("mg_vo_prepare",
          lib.mg_vo_prepare(Cx.data_ptr(), ldc, ridge, Wv.data_ptr(), ldwv, Wo.data_ptr(), ldwo,
                            n_heads, n_kv_heads, hd, d, method, info.data_ptr(), ws.data_ptr(), nbytes,
                            _stream()))
    return ws, info


def vo_prepare_into(ws: torch.Tensor, info: torch.Tensor, Cx, ridge, Wv, Wo, n_heads, n_kv_heads, hd,
                    method: int) -> None:
    """`vo_prepare` into an existing workspace (the retry with the other method)."""
    d, ldc, ldwv, ldwo = _vo_check(Cx, Wv, Wo, n_heads, n_kv_heads, hd, "vo_prepare")
    check("mg_vo_prepare",
          lib.mg_vo_prepare(Cx.data_ptr(), ldc, ridge, Wv.data_ptr(), ldwv, Wo.data_ptr(), ldwo,
                            n_heads, n_kv_heads, hd, d, method, info.data_ptr(), ws.data_ptr(),
                            ws.numel(), _stream()))


def vo_outputs(Wv: torch.Tensor, n_heads: int, n_kv_heads: int, r: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Output buffers of `vo_finish` (v_proj [KV*r, d], o_proj [d, H*r]).  Allocating them up front,
    on the caller's stream, keeps `vo_finish` free of allocations: a first allocation on a fresh
    side stream is a cudaMalloc, which synchronises the device and would serialise the streams."""
    d = Wv.shape[1]
    return (torch.empty(n_kv_heads * r, d, dtype=torch.bfloat16, device=Wv.device),
            torch.empty(d, n_heads * r, dtype=torch.bfloat16, device=Wv.device))


def vo_finish(ws: torch.Tensor, Wv: torch.Tensor, Wo: torch.Tensor, n_heads: int, n_kv_heads: int,
              hd: int, r: int, out: tuple[torch.Tensor, torch.Tensor] | None = None
              ) -> tuple[torch.Tensor, torch.Tensor]:
    """Second half of `vo_compress`: per-head eigensolves + recombination, on the current stream."""
    d = Wv.shape[1]
    v_out, o_out = out if out is not None else vo_outputs(Wv, n_heads, n_kv_heads, r)
    check("mg_vo_finish",
          lib.mg_vo_finish(Wv.data_ptr(), Wv.stride(0), Wo.data_ptr(), Wo.stride(0), n_heads, n_kv_heads,
                           hd, d, r, v_out.data_ptr(), v_out.stride(0), o_out.data_ptr(),
                           o_out.stride(0), int(v_out.dtype == torch.float32), ws.data_ptr(),
                           ws.numel(), _stream()))
    return v_out, o_out


# ------------------------------------------------------------------------------------------------
# calibration forward: fused elementwise kernels
# ------------------------------------------------------------------------------------------------
def rmsnorm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """HF *RMSNorm.forward in one kernel (bf16, last dim contiguous)."""
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    y = torch.empty(x2.shape, dtype=x.dtype, device=x.device)
    check("mg_rmsnorm_bf16",
          lib.mg_rmsnorm_bf16(x2.data_ptr(), x2.stride(0), x2.shape[0], x2.shape[1], weight.data_ptr(),
                              eps, y.data_ptr(), y.stride(0), _stream()))
    return y.view(shape)


def swiglu(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """bf16(silu(gate)) * up for contiguous bf16 tensors of equal shape."""
    out = torch.empty_like(gate)
    check("mg_swiglu_bf16",
          lib.mg_swiglu_bf16(gate.data_ptr(), up.data_ptr(), out.data_ptr(), gate.numel(), _stream()))
    return out


def rope_bthd(x_bthd: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """Rotary embedding on a contiguous [B, T, H, hd] bf16 tensor; cos / sin [1 or B, T, hd]."""
    B, T, H, hd = x_bthd.shape
    out = torch.empty_like(x_bthd)
    stride = 0 if cos.shape[0] == 1 else cos.stride(0)
    check("mg_rope_bf16",
          lib.mg_rope_bf16(x_bthd.data_ptr(), out.data_ptr(), cos.data_ptr(), sin.data_ptr(), B, T, H,
                           hd, stride, _stream()))
    return out


# ------------------------------------------------------------------------------------------------
# rebuilt-model ops (masked RoPE / masked q-k norm) and the perplexity kernel
# ------------------------------------------------------------------------------------------------
def rope_masked_bthr(x_bthr: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor, mask: torch.Tensor,
                     group: int) -> torch.Tensor:
    """Masked rotary embedding on a contiguous [B, T, H, r] bf16 tensor; cos / sin [1 or B, T, hd]
    of the original head dim; mask [H / group, r] int64 (contiguous)."""
    B, T, H, r = x_bthr.shape
    out = torch.empty_like(x_bthr)
    stride = 0 if cos.shape[0] == 1 else cos.stride(0)
    check("mg_rope_masked_bf16",
          lib.mg_rope_masked_bf16(x_bthr.data_ptr(), out.data_ptr(), cos.data_ptr(), sin.data_ptr(),
                                  mask.data_ptr(), B, T, H, group, r, cos.shape[-1], stride, _stream()))
    return out


def rmsnorm_masked(x: torch.Tensor, weight: torch.Tensor, mask: torch.Tensor, group: int,
                   eps: float) -> torch.Tensor:
    """x contiguous [..., H, r] bf16; weight [hd] bf16; mask [H / group, r] int64."""
    H, r = x.shape[-2], x.shape[-1]
    rows = x.numel() // (H * r)
    out = torch.empty_like(x)
    check("mg_rmsnorm_masked_bf16",
          lib.mg_rmsnorm_masked_bf16(x.data_ptr(), rows, H, group, r, weight.data_ptr(), mask.data_ptr(),
                                     eps, out.data_ptr(), _stream()))
    return out


def ce_rows(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """Per-row negative log-likelihood (fp32 [rows]) of bf16 logits [rows, vocab]."""
    rows, vocab, ld = _rowmajor_2d(logits, "logits")
    if logits.dtype != torch.bfloat16 or labels.dtype != torch.int64 or not labels.is_contiguous():
        raise TypeError("ce_rows: logits bfloat16 [rows, vocab], labels contiguous int64 [rows]")
    if labels.numel() != rows:
        raise ValueError("ce_rows: one label per row")
    nll = torch.empty(rows, dtype=torch.float32, device=logits.device)
    check("mg_ce_rows_bf16",
          lib.mg_ce_rows_bf16(logits.data_ptr(), ld, rows, vocab, labels.data_ptr(), nll.data_ptr(), _stream()))
    return nll
