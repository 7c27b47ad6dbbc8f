"""Rank allocation and the PSD square root (reference: src/compression_utils.py).

`allocate_global_sparsity` decides every rank in the model, so it reproduces the reference's
arithmetic bit for bit — fp64 on the host, including its two quirks: Block-Influence scores pass
through float32 (`torch.tensor(list)`), and capped layers are re-opened every round of the
redistribution loop.  `sqrt_M` is kept for API parity only: the type-II / type-III kernels never
form a matrix square root (||sqrt(C + rho I)[:, j]||^2 == C_jj + rho, and any factor F with
F^T F = C + rho I yields the same V/O factors — SURVEY Appendix B).
"""
from __future__ import annotations

import logging

import torch
from torch import Tensor

logger = logging.getLogger("MoDeGPT")

_MAX_REDISTRIBUTION_ROUNDS = 1_000_000


def allocate_global_sparsity(bi_scores: list[float], compression_ratio: float,
                             smoothing: float = 0.015, max_sparsity: float = 0.8, adapter=None,
                             invert: bool = False) -> list[float]:
    """softmax(-BI / smoothing) * (L * ratio), capped at `max_sparsity` with the excess handed to
    the other layers in proportion to their softmax weights; returns keep ratios = 1 - sparsity
    (src/compression_utils.py:79-124)."""
    if adapter is not None:
        adapter.metrics["smoothing"] = smoothing
    s = torch.tensor(bi_scores, dtype=torch.float32).to(torch.float64)   # fp32 round trip on purpose
    if invert:
        s = -s
    weights = torch.softmax(-s / smoothing, dim=0)
    sparsity = weights * (len(bi_scores) * compression_ratio)
    logger.info(f"Max Layer Sparsity: {sparsity.max().item()}, Avg = {sparsity.mean().item()}")
    if adapter is not None:
        adapter.metrics["max_layer_sparsity"] = sparsity.max().item()
    for _ in range(_MAX_REDISTRIBUTION_ROUNDS):
        over = sparsity > max_sparsity
        if not bool(over.any()):
            break
        excess = (sparsity[over] - max_sparsity).sum()
        sparsity[over] = max_sparsity
        free = ~over
        if bool(free.any()):
            sparsity[free] += excess * (weights[free] / weights[free].sum())
    else:
        raise RuntimeError("allocate_global_sparsity: cap redistribution did not converge "
                           "(the reference's loop would not terminate on these scores either)")
    return (1 - sparsity).tolist()


def head_rank(head_dim: int, keep_ratio: float, rope: bool, clamp_to_head: bool = True) -> int:
    """int(head_dim * keep) clamped to >= 1 (and <= head_dim for Q/K), rounded down to even and
    >= 2 for RoPE architectures (compress_qk.py:176-182, compress_vo.py:35-41).  Q/K and V/O must
    agree: the rebuilt attention uses one per-layer head dim for q, k and v."""
    r = int(head_dim * keep_ratio)
    r = max(1, min(r, head_dim)) if clamp_to_head else max(1, r)
    if rope:
        r -= r % 2
        r = max(2, min(r, head_dim)) if clamp_to_head else max(2, r)
    return r


@torch.no_grad()
def sqrt_M(M: Tensor, ridge_lambda: float = 1e-4, scaled: bool = False, debug: str = "",
           inverse_sqrt: bool = False):
    """V diag(sqrt(lambda + ridge * scale)) V^T (src/compression_utils.py:15-55).  Compatibility
    helper for callers of the reference API; NOT used by the compression kernels."""
    lam, vec = torch.linalg.eigh(M)
    scale = lam.max() if scaled else 1.0
    root = torch.sqrt((lam + ridge_lambda * scale).clamp(min=0))
    out = (vec * root) @ vec.T
    if not inverse_sqrt:
        return out.to(M.dtype)
    inv = (vec * (1.0 / root.clamp(min=1e-12))) @ vec.T
    return out.to(M.dtype), inv.to(M.dtype)
