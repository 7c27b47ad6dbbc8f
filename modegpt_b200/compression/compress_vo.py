"""Type-III: SVD compression of V/O (reference: src/compression/compress_vo.py:13-223).

The reference forms sqrt(C_x) by a d x d eigendecomposition, inverts it by LU and runs one (GQA)
or two (MHA, the second one FULL d x d) SVDs per head.  Algebraically the new heads are hd x r
recombinations of the old ones determined by hd x hd symmetric eigenproblems (SURVEY §3.5,
Appendix B4/B5); `mg_vo_compress` builds those Gram matrices on the tensor cores, solves the
eigenproblems with a batched fp64 Jacobi in shared memory and applies the recombination.
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
    if target_layers is None:
        target_layers = list(range(adapter.n_layers))
    H, KV, hd = adapter.n_heads, adapter.n_kv_heads, adapter.head_dim
    layers = list(D.owned_layers(target_layers))
    # The per-head eigensolver runs one CTA per kv head (32 of 148 SMs for an MHA layer, 8 for a
    # GQA one) and nothing in this function synchronises with the host, so with `--vo_streams N`
    # consecutive layers are issued round-robin on N CUDA streams and their eigensolves can run
    # side by side.  Opt-in: measured 0.30 -> 0.14 s for 32 Llama-2-7B layers in one run and 9.4 ->
    # 16.9 ms/layer in another — a persistent GEMM grid of the next layer that cannot be fully
    # resident next to the eigensolver CTAs leaves its unplaced CTAs' tiles waiting (the engine
    # strides tiles statically); it needs a dynamic tile scheduler to be dependable.
    streams, caller = [None], None
    if layers and cov[layers[0]].is_cuda and getattr(adapter.config, "vo_streams", 1) != 1:
        device = cov[layers[0]].device
        n_streams = getattr(adapter.config, "vo_streams", 1) or max(2, min(8, 128 // max(KV, 1)))
        caller = torch.cuda.current_stream(device)
        ready = torch.cuda.Event()
        ready.record(caller)
        streams = [torch.cuda.Stream(device=device) for _ in range(min(n_streams, len(layers)))]
        for st in streams:
            st.wait_event(ready)       # statistics / weights were produced on the caller's stream
    try:
        _compress_layers(adapter, cov, keep_ratios, layers, streams, H, KV, hd)
    finally:
        if caller is not None:
            for st in streams:
                caller.wait_stream(st)


def _compress_layers(adapter, cov, keep_ratios, layers, streams, H, KV, hd):
    import contextlib

    for n, layer in enumerate(layers):
        st = streams[n % len(streams)]
        with (torch.cuda.stream(st) if st is not None else contextlib.nullcontext()):
            _compress_layer(adapter, cov, keep_ratios, layer, H, KV, hd)


def _compress_layer(adapter, cov, keep_ratios, layer, H, KV, hd):
    # same rule as Q/K so the rebuilt attention has ONE head dim per layer;
    # the reference does not clamp the V/O rank to head_dim (compress_vo.py:36-41).
    rank_i = min(head_rank(hd, keep_ratios[layer], adapter.uses_rope, clamp_to_head=False), hd)
    try:
        comps = adapter.get_attn_components(layer)
        wv, wo = comps.v_proj.weight.detach(), comps.o_proj.weight.detach()
    except AttributeError as e:     # the reference skips such layers too (compress_vo.py:47-53)
        logger.warning(f"[VO] Layer {layer}: cannot access v_proj/o_proj: {e}")
        return
    v_new, o_new = ops.vo_compress(cov[layer], adapter.config.ridge_vo, wv.contiguous(),
                                   wo.contiguous(), H, KV, hd, rank_i)
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
