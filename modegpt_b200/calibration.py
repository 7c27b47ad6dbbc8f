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
    try:
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
