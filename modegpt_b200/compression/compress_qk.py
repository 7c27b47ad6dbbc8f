"""Type-II: CR compression of Q/K (reference: src/compression/compress_qk.py:152-476).

Per kv head a CR score ranks the head dimensions; RoPE architectures rank (j, j + hd/2) PAIRS so
the rotary structure survives, and the kept indices become the layer's rotary mask.  The score is
the product of column norms of sqrt(C_q + rho_q I) and sqrt(C_k + rho_k I), i.e. products of
(C_jj + rho) — one kernel reads the two diagonals, ranks and writes the mask; a second gathers
the weight rows (mg_qk_select_f32 / mg_gather_head_rows_bf16).

Ridges follow the reference exactly: the K side uses `config.ridge_qk` only on the
GQA path; the Q side, and both sides of the MHA and OPT paths, use sqrt_M's default 1e-4.
"""
from __future__ import annotations

import logging

import torch
from torch import Tensor

from .. import distributed as D
from .. import ops
from ..adapters.model_adapter import ModelAdapter
from ..compression_utils import head_rank

logger = logging.getLogger("MoDeGPT")

SQRT_M_DEFAULT_RIDGE = 1e-4  # src/compression_utils.py:17


@torch.no_grad()
def compress_qk(adapter: ModelAdapter, cov, keep_ratios, rank=None, slice_dims=True,
                target_layers: list[int] | None = None, local_masks: dict | None = None):
    """Returns the rotary masks ([KV, r] int64, one per target layer, in layer order).

    `local_masks` (additive): a dict that receives this rank's {layer: mask} INSTEAD of the
    cross-rank gather — the layer-streamed flow calls this once per layer and would otherwise put a
    collective (a synchronisation of all ranks on the owner's decomposition) into every layer; it
    gathers once at the end with `gather_rotary_masks`."""
    if target_layers is None:
        target_layers = list(range(adapter.n_layers))
    cov_q_list, cov_k_list = cov
    hd = adapter.head_dim
    local: dict[int, Tensor] = {}
    for i in D.owned_layers(target_layers):
        rank_i = head_rank(hd, keep_ratios[i], adapter.uses_rope) if rank is None else rank
        mask = compress_layer(adapter, i, rank_i, cov_q_list=cov_q_list[i], cov_k_list=cov_k_list[i])
        if adapter.uses_rope:
            local[i] = mask
        logger.info(f"[QK] Layer {i}: compressed to rank {rank_i} per head (CR score)")
    if not slice_dims:
        return None
    if local_masks is not None:
        local_masks.update(local)
        return None
    if D.is_distributed():
        device = next(adapter.model.parameters()).device
        merged = D.gather_masks(local, adapter.n_layers, adapter.n_kv_heads, hd, device)
        return [merged[i] for i in target_layers if merged[i] is not None]
    return [local[i] for i in target_layers if i in local]


def gather_rotary_masks(adapter: ModelAdapter, local_masks: dict, target_layers: list[int]) -> list:
    """Every rank's {layer: mask} -> the full list in layer order on every rank (one collective)."""
    if D.is_distributed():
        device = next(adapter.model.parameters()).device
        merged = D.gather_masks(local_masks, adapter.n_layers, adapter.n_kv_heads, adapter.head_dim, device)
        return [merged[i] for i in target_layers if merged[i] is not None]
    return [local_masks[i] for i in target_layers if i in local_masks]


@torch.no_grad()
def compress_layer(adapter: ModelAdapter, layer_idx: int, rank: int, cov_q_list: Tensor,
                   cov_k_list: Tensor, rotary_masks: list | None = None, slice_dims=True,
                   bias=True) -> Tensor:
    """One layer (compress_qk.py:208-308): select per kv head, gather W_q / W_k rows (and the
    OPT biases), write `layer_{i}_qk`.  Output head order: every query head keeps its position."""
    H, KV, hd = adapter.n_heads, adapter.n_kv_heads, adapter.head_dim
    comps = adapter.get_qk_components(layer_idx)
    wq, wk = comps.query_proj.weight.detach(), comps.key_proj.weight.detach()
    grouped = KV != H
    if adapter.uses_rope:
        ridge_k = adapter.config.ridge_qk if grouped else SQRT_M_DEFAULT_RIDGE
        mask = ops.qk_select(cov_q_list, cov_k_list, rank, 0, SQRT_M_DEFAULT_RIDGE, ridge_k)
    elif adapter.arch == "opt":
        mask = ops.qk_select(cov_q_list, cov_k_list, rank, 1, SQRT_M_DEFAULT_RIDGE,
                             SQRT_M_DEFAULT_RIDGE)
    else:
        raise NotImplementedError(f"type-II compression is not defined for arch {adapter.arch!r}")
    weights = {"q_proj": ops.gather_head_rows(wq, mask, H, H // KV, hd),
               "k_proj": ops.gather_head_rows(wk, mask, KV, 1, hd)}
    bq, bk = getattr(comps.query_proj, "bias", None), getattr(comps.key_proj, "bias", None)
    if bq is not None and bk is not None:   # OPT, Qwen2: the bias entries follow their rows
        rows_q = (torch.arange(H, device=mask.device) * hd)[:, None] + mask.repeat_interleave(H // KV, 0)
        rows_k = (torch.arange(KV, device=mask.device) * hd)[:, None] + mask
        weights["q_bias"] = bq.detach()[rows_q.reshape(-1)]
        weights["k_bias"] = bk.detach()[rows_k.reshape(-1)]
    if rotary_masks is not None and adapter.uses_rope:
        rotary_masks.append(mask)
    adapter.save_layer(output_dir=adapter.config.temp_storage_dir, suffix="qk", weights=weights,
                       layer_idx=layer_idx)
    return mask
