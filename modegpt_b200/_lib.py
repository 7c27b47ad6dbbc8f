"""ctypes binding of libmodegpt_b200.so (the C ABI declared in include/modegpt_b200.h).

There is deliberately no fallback: if the shared library is missing or a symbol is absent the
import fails loudly, and every wrapper raises on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libmodegpt_b200.so"

i64 = C.c_int64
vp = C.c_void_p
f32 = C.c_float
f64 = C.c_double
i32 = C.c_int

# name -> (restype, argtypes); must list every symbol of include/modegpt_b200.h
SIGNATURES: dict[str, tuple] = {
    "mg_version": (i32, []),
    "mg_device_sm_count": (i32, []),
    "mg_error_string": (C.c_char_p, [i32]),
    "mg_syrk_bf16_f32": (i32, [vp, i64, i64, i64, vp, i64, f32, i32, vp]),
    "mg_syrk_heads_bf16_f32": (i32, [vp, i64, i64, i64, i32, vp, f32, i32, vp]),
    "mg_add_diag_blocks_f32": (i32, [vp, i64, i64, i32, f32, vp, vp]),
    "mg_bi_cosine_bf16": (i32, [vp, i64, vp, i64, i64, i64, vp, vp]),
    "mg_finalize_sym_f32": (i32, [vp, i64, i64, f32, vp]),
    "mg_scale_f32": (i32, [vp, i64, f32, vp]),
    "mg_pack_upper_f32": (i32, [vp, i64, i64, vp, vp]),
    "mg_unpack_upper_f32": (i32, [vp, i64, vp, i64, vp]),
    "mg_set_concurrent_factorizations": (i32, [i32]),
    "mg_ridge_scores_ws_bytes": (C.c_size_t, [i64]),
    "mg_ridge_scores_f32": (i32, [vp, i64, i64, f32, vp, vp, C.c_size_t, vp, vp]),
    "mg_select_k_f32": (i32, [vp, i64, i64, i32, vp, vp]),
    "mg_gather_rows_bf16": (i32, [vp, i64, vp, i64, i64, vp, i64, vp]),
    "mg_nystrom_down_ws_bytes": (C.c_size_t, [i64, i64, i64]),
    "mg_nystrom_down_f32": (i32, [vp, i64, i64, vp, i64, vp, i64, i64, f32, vp, i64, vp,
                                  C.c_size_t, vp, vp, vp]),
    "mg_nystrom_refine_f32": (i32, [vp, i64, i64, vp, i64, i64, f32, vp, i64, vp, C.c_size_t, vp]),
    "mg_qk_select_f32": (i32, [vp, vp, i32, i32, i32, i32, f32, f32, i32, vp, vp]),
    "mg_gather_head_rows_bf16": (i32, [vp, i64, vp, i32, i32, i64, i64, i64, vp, i64, vp]),
    "mg_vo_ws_bytes": (C.c_size_t, [i64, i32, i32, i32]),
    "mg_vo_compress": (i32, [vp, i64, f32, vp, i64, vp, i64, i32, i32, i32, i64, i32, i32, vp, i64,
                             vp, i64, i32, vp, vp, C.c_size_t, vp]),
    "mg_vo_prepare": (i32, [vp, i64, f32, vp, i64, vp, i64, i32, i32, i32, i64, i32, vp, vp,
                            C.c_size_t, vp]),
    "mg_vo_finish": (i32, [vp, i64, vp, i64, i32, i32, i32, i64, i32, vp, i64, vp, i64, i32, vp,
                           C.c_size_t, vp]),
    "mg_rmsnorm_bf16": (i32, [vp, i64, i64, i64, vp, f32, vp, i64, vp]),
    "mg_swiglu_bf16": (i32, [vp, vp, vp, i64, vp]),
    "mg_rope_bf16": (i32, [vp, vp, vp, vp, i64, i64, i32, i32, i64, vp]),
    "mg_rope_masked_bf16": (i32, [vp, vp, vp, vp, vp, i64, i64, i32, i32, i32, i32, i64, vp]),
    "mg_rmsnorm_masked_bf16": (i32, [vp, i64, i32, i32, i32, vp, vp, f32, vp, vp]),
    "mg_ce_rows_bf16": (i32, [vp, i64, i64, i64, vp, vp, vp]),
}


class MgError(RuntimeError):
    def __init__(self, fn: str, code: int, msg: str):
        super().__init__(f"{fn} failed with status {code}: {msg}")
        self.fn = fn
        self.code = code


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m modegpt_b200.build` "
            "(there is no CPU or PyTorch fallback for the compression hot path)."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(fn: str, code: int) -> int:
    """Raise on argument / CUDA errors (code < 0); numerical info (code > 0) is returned."""
    if code < 0:
        raise MgError(fn, code, lib.mg_error_string(code).decode())
    return code
