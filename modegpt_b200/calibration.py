"""Calibration: hooked forwards -> C_mlp, C_q, C_k, C_x (fp32, normalised, full symmetric) and
Block-Influence scores.  Interface of the reference's `load_calibs` (src/calibration.py:18-36).

Differences underneath: statistics are accumulated by the tcgen05 SYRK kernels
straight from the live bf16 activations in fp32; BI comes from hooks on the blocks (no
`output_hidden_states`, no per-layer `.item()`); with `torch.distributed` initialised the batches
are dealt to ranks and each layer's sums are reduced once to the layer's owner.
"""
from __future__ import annotations

import logging

import torch

from . import distributed as D
from . import ops
from .adapters.model_adapter import ModelAdapter
from .fused_forward import fused_elementwise

logger = logging.getLogger("MoDeGPT")

NORMALISER_SEQ_LEN = 2048  # src/calibration.py:141 divides by n_texts * 2048 whatever the length


def load_calibs(adapter: ModelAdapter, n_samples: int, batch_size: int, dataset: str = "wikitext",
                load_calibs_from: str = "", calibs_save_path: str = "",
                target_layers: list[int] | None = None):
    return _calibrate_model(adapter, n_samples=n_samples, batch_size=batch_size, dataset=dataset,
                            target_layers=target_layers or [])


@torch.no_grad()
def _calibrate_model(adapter: ModelAdapter, n_samples: int, batch_size: int,
                     target_layers: list[int], dataset: str = "wikitext"):
    model = adapter.model
    n_layers = adapter.n_layers
    blocks = adapter.get_transformer_blocks()
    if not target_layers:
        target_layers = list(range(n_layers))
    device = next(model.parameters()).device

    if adapter.calibs is None:
        from .eval import load_calibration_texts

        adapter.calibs = load_calibration_texts(calib_size=n_samples, model=model,
                                                tokenizer=adapter.tokenizer, batch_size=batch_size,
                                                dataset=dataset, seq_len=adapter.config.seq_len,
                                                seed=adapter.config.seed)
    logger.info(f"Detected architecture: {adapter.arch}")
    logger.info(f"target_layers = {target_layers}")

    n_inner, d = adapter.get_n_inner(), adapter.d_model
    H, KV, hd = adapter.n_heads, adapter.n_kv_heads, adapter.head_dim
    f32 = dict(dtype=torch.float32, device=device)
    cov_mlp = [None] * n_layers
    cov_q = [None] * n_layers
    cov_k = [None] * n_layers
    cov_x = [None] * n_layers
    for i in target_layers:
        cov_mlp[i] = torch.zeros(n_inner, n_inner, **f32)
        cov_q[i] = torch.zeros(H, hd, hd, **f32)
        cov_k[i] = torch.zeros(KV, hd, hd, **f32)
        cov_x[i] = torch.zeros(d, d, **f32)

    handles: list = []
    for i in target_layers:
        adapter.register_hooks(i, blocks[i], cov_mlp_list=cov_mlp, cov_q_list=cov_q,
                               cov_k_list=cov_k, cov_x_list=cov_x, handles=handles, logger=logger)
    bi_batch = torch.zeros(n_layers, dtype=torch.float64, device=device)
    bi_total = torch.zeros(n_layers, dtype=torch.float64, device=device)
    adapter.register_bi_hooks(bi_batch, handles)

    model.eval()
    # the LM head is not on the statistics path: run the decoder stack only (its final norm, whose
    # output closes the last Block-Influence pair, is part of it)
    body = getattr(model, "model", model)
    n_texts = 0
    import contextlib

    fuse = contextlib.nullcontext() if adapter.config.eager_forward else fused_elementwise(model)
    try:
        with fuse:
            for batch in D.shard_batches(adapter.calibs):
                batch = batch.to(device, non_blocking=True)
                n_texts += len(batch)
                body(batch, use_cache=False)
                # sum over the batch, mean over positions (src/calibration.py:122-124)
                bi_total += bi_batch / batch.shape[1]
                bi_batch.zero_()
    finally:
        for h in handles:
            h.remove()

    # ---- cross-rank exchange: one reduction per statistic per layer, to the layer's owner
    count = torch.tensor([float(n_texts)], dtype=torch.float64, device=device)
    D.all_reduce_sum_(count)
    D.all_reduce_sum_(bi_total)
    n_texts_global = int(count.item())
    bi_scores = (bi_total / n_texts_global).tolist()
    adapter.bi_scores = bi_scores

    scale = 1.0 / float(n_texts_global * NORMALISER_SEQ_LEN)
    for i in target_layers:
        mine = True
        for lst in (cov_mlp, cov_x, cov_q, cov_k):
            mine = D.reduce_to_owner(lst[i], i)
        if not mine:
            cov_mlp[i] = cov_q[i] = cov_k[i] = cov_x[i] = None   # only the owner decomposes layer i
            continue
        ops.finalize_sym_(cov_mlp[i], scale)
        ops.finalize_sym_(cov_x[i], scale)
        ops.scale_(cov_q[i], scale)
        ops.scale_(cov_k[i], scale)
    logger.info("Finished calibration and computed BI scores.")
    return cov_mlp, cov_q, cov_k, cov_x, bi_scores


# ================================================================================================
# Layer-streamed calibration (70B-class models: SURVEY §7 hard part 3, §8e)
# ================================================================================================
# All-layer statistics do not fit beside a large model (Llama-2-70B: 80 x 3.3 GB of C_mlp next to
# 140 GB of weights).  Statistics come from the UNMODIFIED model, so the calibration can run layer
# by layer over resident hidden states: one layer's accumulators exist at a time, they are reduced
# to the layer's owner, handed to the caller (who decomposes and frees them) and the hidden states
# advance.  Numerically identical to `load_calibs`: same kernels, same per-(layer, batch) order.
# Pass 1 (`block_influence`) produces the BI scores every rank needs before any rank is known.


def _forward_mode(adapter: ModelAdapter):
    import contextlib

    return contextlib.nullcontext() if adapter.config.eager_forward else fused_elementwise(adapter.model)


class _LayerStepper:
    """Runs Llama-style decoder layers one at a time over resident per-batch hidden states."""

    def __init__(self, adapter: ModelAdapter, dataset: str):
        model = adapter.model
        self.adapter = adapter
        self.body = getattr(model, "model", model)
        if not (hasattr(self.body, "rotary_emb") and hasattr(self.body, "embed_tokens")):
            raise NotImplementedError("layer-streamed calibration needs a Llama/Qwen3-style decoder")
        if any(t != "full_attention" for t in (getattr(model.config, "layer_types", None) or [])):
            raise NotImplementedError("layer-streamed calibration supports full-attention layers only")
        self.device = next(model.parameters()).device
        if adapter.calibs is None:
            from .eval import load_calibration_texts

            cfg = adapter.config
            adapter.calibs = load_calibration_texts(calib_size=cfg.calib_size, model=model,
                                                    tokenizer=adapter.tokenizer,
                                                    batch_size=cfg.calibs_batch_size, dataset=dataset,
                                                    seq_len=cfg.seq_len, seed=cfg.seed)
        self.batches = [b.to(self.device) for b in D.shard_batches(adapter.calibs)]
        self.n_texts = sum(len(b) for b in self.batches)
        self.blocks = adapter.get_transformer_blocks()

    def embed(self) -> list[torch.Tensor]:
        self.hidden = [self.body.embed_tokens(b) for b in self.batches]
        self.pos = []
        for h in self.hidden:
            ids = torch.arange(h.shape[1], device=h.device).unsqueeze(0)
            self.pos.append((ids, self.body.rotary_emb(h, position_ids=ids)))
        return self.hidden

    def run_layer(self, layer_idx: int, b: int) -> torch.Tensor:
        ids, pe = self.pos[b]
        out = self.blocks[layer_idx](self.hidden[b], attention_mask=None, position_ids=ids,
                                     position_embeddings=pe, past_key_values=None, use_cache=False)
        return out[0] if isinstance(out, (tuple, list)) else out


@torch.no_grad()
def block_influence(adapter: ModelAdapter, dataset: str = "synthetic") -> list[float]:
    """Pass 1 of the streamed flow: BI scores of every layer (src/calibration.py:118-136 semantics,
    including the post-final-norm quirk of the last pair)."""
    st = _LayerStepper(adapter, dataset)
    st.embed()
    L = adapter.n_layers
    final_norm = None
    from .adapters.model_adapter import _get

    final_norm = _get(adapter.model, adapter.module_map.final_norm)
    total = torch.zeros(L, dtype=torch.float64, device=st.device)
    acc = torch.zeros(1, dtype=torch.float64, device=st.device)
    with _forward_mode(adapter):
        for l in range(L):
            for b in range(len(st.batches)):
                out = st.run_layer(l, b)
                acc.zero_()
                ops.bi_cosine_(acc, st.hidden[b], final_norm(out) if l == L - 1 else out)
                total[l] += acc[0] / out.shape[1]
                st.hidden[b] = out
    count = torch.tensor([float(st.n_texts)], dtype=torch.float64, device=st.device)
    D.all_reduce_sum_(count)
    D.all_reduce_sum_(total)
    adapter.bi_scores = (total / count).tolist()
    return adapter.bi_scores


@torch.no_grad()
def iter_layer_statistics(adapter: ModelAdapter, target_layers: list[int] | None = None,
                          dataset: str = "synthetic"):
    """Pass 2: yields (layer_idx, cov_mlp, cov_q, cov_k, cov_x) one layer at a time — normalised,
    full symmetric fp32 on the layer's owner, four Nones elsewhere.  The caller must drop the
    tensors before asking for the next layer if memory matters."""
    st = _LayerStepper(adapter, dataset)
    st.embed()
    L = adapter.n_layers
    targets = set(target_layers) if target_layers else set(range(L))
    n_inner, d = adapter.get_n_inner(), adapter.d_model
    H, KV, hd = adapter.n_heads, adapter.n_kv_heads, adapter.head_dim
    f32 = dict(dtype=torch.float32, device=st.device)
    count = torch.tensor([float(st.n_texts)], dtype=torch.float64, device=st.device)
    D.all_reduce_sum_(count)
    scale = 1.0 / float(int(count.item()) * NORMALISER_SEQ_LEN)
    for l in range(L):
        handles: list = []
        cov = None
        if l in targets:
            cov = ([None] * L, [None] * L, [None] * L, [None] * L)
            cov[0][l] = torch.zeros(n_inner, n_inner, **f32)
            cov[1][l] = torch.zeros(H, hd, hd, **f32)
            cov[2][l] = torch.zeros(KV, hd, hd, **f32)
            cov[3][l] = torch.zeros(d, d, **f32)
            adapter.register_hooks(l, st.blocks[l], cov_mlp_list=cov[0], cov_q_list=cov[1],
                                   cov_k_list=cov[2], cov_x_list=cov[3], handles=handles, logger=logger)
        try:
            with _forward_mode(adapter):
                for b in range(len(st.batches)):
                    st.hidden[b] = st.run_layer(l, b)
        finally:
            for h in handles:
                h.remove()
        if cov is None:
            continue
        mine = True
        for lst in cov:
            mine = D.reduce_to_owner(lst[l], l)
        if not mine:
            del cov
            yield l, None, None, None, None
            continue
        ops.finalize_sym_(cov[0][l], scale)
        ops.finalize_sym_(cov[3][l], scale)
        ops.scale_(cov[1][l], scale)
        ops.scale_(cov[2][l], scale)
        out = (l, cov[0][l], cov[1][l], cov[2][l], cov[3][l])
        del cov
        yield out
