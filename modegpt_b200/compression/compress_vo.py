"""Type-III: SVD compression of V/O (reference: src/compression/compress_vo.py:13-223).

The reference forms sqrt(C_x) by a d x d eigendecomposition, inverts it by LU and runs one (GQA)
or two (MHA, the second one FULL d x d) SVDs per head.  Algebraically the new heads are hd x r
recombinations of the old ones determined by hd x hd symmetric eigenproblems (SURVEY §3.5,
Appendix B4/B5): the right singular pairs of sqrt(C) W_v,h^T are the eigenpairs of
G1[h] = W_v,h (C_x + ridge I) W_v,h^T.  `mg_vo_compress` forms P = (C_x + ridge I) W_v^T on the
tensor cores, accumulates the hd x hd Grams W_v,h P_h in FP64 (the fp32 tensor-core accumulator
would lose 2^-24 of the largest eigenvalue, which is what the small singular values divide by),
solves the eigenproblems with a batched fp64 Jacobi in shared memory and applies the recombination.
(`method=VO_FACTOR` goes through the blocked Cholesky factor of C_x + ridge I instead; measured
equally accurate up to cond(G1) = 3e5 and 3x slower, so it is not the default.)
Singular vectors are defined up to sign, so V' / O' match the reference up to a per-component
sign; the products O'V' match.
"""
from __future__ import annotations

import logging

import torch

from .. import distributed as D
from .. import ops
from ..adapters.model_adapter import ModelAdapter
from ..compression_utils import head_rank

logger = logging.getLogger("MoDeGPT")


@torch.no_grad()
def compress_vo(adapter: ModelAdapter, cov, keep_ratios=None, slice_dims=True,
                target_layers: list[int] | None = None):
    """Layer loop + `save_layer(suffix="vo")` (compress_vo.py:13-109).

    The per-head eigensolver runs one CTA per kv head for milliseconds (32 of 148 SMs for an MHA
    layer, 8 for a GQA one).  Layers are therefore taken in groups: the tensor-core halves
    (`mg_vo_prepare`) of a group run back to back on the caller's stream, then the eigensolver halves
    (`mg_vo_finish`) are issued on one stream per layer so they run side by side; the group is
    joined before the next one starts, so a persistent GEMM grid never has to share the GPU with
    long-running eigensolver CTAs.  `--vo_group 1` keeps one `mg_vo_compress` per layer."""
    if target_layers is None:
        target_layers = list(range(adapter.n_layers))
    H, KV, hd = adapter.n_heads, adapter.n_kv_heads, adapter.head_dim
    layers = list(D.owned_layers(target_layers))
    group = getattr(adapter.config, "vo_group", 0) or max(1, min(8, 128 // max(KV, 1)))
    if group <= 1 or not layers or not cov[layers[0]].is_cuda:
        for layer in layers:
            _compress_layer(adapter, cov, keep_ratios, layer, H, KV, hd)
        return
    device = cov[layers[0]].device
    caller = torch.cuda.current_stream(device)
    streams = _side_streams(device, group)
    for g0 in range(0, len(layers), group):
        chunk = layers[g0:g0 + group]
        prepared = [_prepare_layer(adapter, cov, keep_ratios, layer, H, KV, hd) for layer in chunk]
        ready = torch.cuda.Event()
        ready.record(caller)
        for st, item in zip(streams, prepared):
            if item is None:
                continue
            with torch.cuda.stream(st):
                st.wait_event(ready)
                _finish_layer(adapter, item, H, KV, hd)
        for st in streams:
            caller.wait_stream(st)


_STREAMS: dict = {}


def _side_streams(device, n: int) -> list:
    pool = _STREAMS.setdefault(torch.device(device), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]


def _layer_inputs(adapter, keep_ratios, layer, hd):
    # same rule as Q/K so the rebuilt attention has ONE head dim per layer;
    # the reference does not clamp the V/O rank to head_dim (compress_vo.py:36-41).
    rank_i = min(head_rank(hd, keep_ratios[layer], adapter.uses_rope, clamp_to_head=False), hd)
    try:
        comps = adapter.get_attn_components(layer)
        wv, wo = comps.v_proj.weight.detach(), comps.o_proj.weight.detach()
    except AttributeError as e:     # the reference skips such layers too (compress_vo.py:47-53)
        logger.warning(f"[VO] Layer {layer}: cannot access v_proj/o_proj: {e}")
        return None
    return rank_i, comps, wv.contiguous(), wo.contiguous()


def _prepare_layer(adapter, cov, keep_ratios, layer, H, KV, hd):
    got = _layer_inputs(adapter, keep_ratios, layer, hd)
    if got is None:
        return None
    rank_i, comps, wv, wo = got
    ws, info = ops.vo_prepare(cov[layer], adapter.config.ridge_vo, wv, wo, H, KV, hd)
    return layer, rank_i, comps, wv, wo, ws, ops.vo_outputs(wv, H, KV, rank_i), info


def _finish_layer(adapter, item, H, KV, hd):
    layer, rank_i, comps, wv, wo, ws, out, _ = item
    here = torch.cuda.current_stream(ws.device)
    for t in (ws, *out):                                     # allocated on the caller's stream
        t.record_stream(here)
    v_new, o_new = ops.vo_finish(ws, wv, wo, H, KV, hd, rank_i, out=out)
    _save(adapter, layer, rank_i, comps, wo, v_new, o_new, H, KV, hd)


def _compress_layer(adapter, cov, keep_ratios, layer, H, KV, hd):
    got = _layer_inputs(adapter, keep_ratios, layer, hd)
    if got is None:
        return
    rank_i, comps, wv, wo = got
    v_new, o_new = ops.vo_compress(cov[layer], adapter.config.ridge_vo, wv, wo, H, KV, hd, rank_i)
    _save(adapter, layer, rank_i, comps, wo, v_new, o_new, H, KV, hd)


def _save(adapter, layer, rank_i, comps, wo, v_new, o_new, H, KV, hd):
    weights = {"v_proj": v_new, "o_proj": o_new}
    bv, bo = getattr(comps.v_proj, "bias", None), getattr(comps.o_proj, "bias", None)
    if bo is not None or bv is not None:
        # attention weights sum to one, so a v bias reaches the output as the constant W_o b_v:
        # fold it into the (kept) output bias instead of dropping it
        fold = torch.zeros(wo.shape[0], device=wo.device, dtype=torch.float32)
        if bv is not None:   # each kv head's bias block serves its whole query group
            b_full = bv.detach().float().view(KV, hd).repeat_interleave(H // KV, dim=0).reshape(-1)
            fold += wo.float() @ b_full
        if bo is not None:
            fold += bo.detach().float()
        weights["o_bias"] = fold.to(torch.bfloat16)
    adapter.save_layer(output_dir=adapter.config.temp_storage_dir, suffix="vo", weights=weights,
                       layer_idx=layer)
    logger.info(f"[VO] Compressed layer {layer} to rank {rank_i} per head")
