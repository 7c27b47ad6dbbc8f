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


class Calibrator:
    """The state of one calibration pass (src/calibration.py:39-150): fp32 accumulators for the
    target layers, the statistics / Block-Influence hooks, and the cross-rank exchange.

        cal = Calibrator(adapter, target_layers)
        for i, batch in enumerate(batches):
            cal.run_batch(batch, last=(i == len(batches) - 1))
        cov_mlp, cov_q, cov_k, cov_x, bi_scores = cal.finish()

    `last=True` arms the reduce-as-you-go exchange: in that batch, as soon as block l has run
    (all four of its accumulations are enqueued), layer l's sums are packed and reduced to the
    layer's owner on a side stream (`distributed.LayerReducer`) while the forward continues with
    block l+1.  Without an armed batch `finish()` submits every layer itself.  `finish()` joins
    the exchange, normalises by the GLOBAL n_texts * 2048 (src/calibration.py:141), mirrors the
    symmetric matrices and returns the reference's five lists; non-owners hold None for a layer."""

    def __init__(self, adapter: ModelAdapter, target_layers: list[int] | None = None):
        import contextlib

        self.adapter = adapter
        model = adapter.model
        self.n_layers = adapter.n_layers
        blocks = adapter.get_transformer_blocks()
        self.targets = list(target_layers) if target_layers else list(range(self.n_layers))
        self.device = next(model.parameters()).device
        n_inner, d = adapter.get_n_inner(), adapter.d_model
        H, KV, hd = adapter.n_heads, adapter.n_kv_heads, adapter.head_dim
        f32 = dict(dtype=torch.float32, device=self.device)
        L = self.n_layers
        self.cov_mlp, self.cov_q, self.cov_k, self.cov_x = [None] * L, [None] * L, [None] * L, [None] * L
        for i in self.targets:
            self.cov_mlp[i] = torch.zeros(n_inner, n_inner, **f32)
            self.cov_q[i] = torch.zeros(H, hd, hd, **f32)
            self.cov_k[i] = torch.zeros(KV, hd, hd, **f32)
            self.cov_x[i] = torch.zeros(d, d, **f32)
        self.handles: list = []
        for i in self.targets:
            adapter.register_hooks(i, blocks[i], cov_mlp_list=self.cov_mlp, cov_q_list=self.cov_q,
                                   cov_k_list=self.cov_k, cov_x_list=self.cov_x, handles=self.handles,
                                   logger=logger)
            self.handles.append(blocks[i].register_forward_hook(self._layer_done(i)))
        self.bi_batch = torch.zeros(L, dtype=torch.float64, device=self.device)
        self.bi_total = torch.zeros(L, dtype=torch.float64, device=self.device)
        adapter.register_bi_hooks(self.bi_batch, self.handles)
        model.eval()
        # the LM head is not on the statistics path: run the decoder stack only (its final norm,
        # whose output closes the last Block-Influence pair, is part of it)
        self.body = getattr(model, "model", model)
        self.n_texts = 0
        self._armed = False
        self._submitted: set[int] = set()
        self._mine: dict[int, bool] = {}
        self.reducer = D.LayerReducer()
        self._stack = contextlib.ExitStack()
        if not adapter.config.eager_forward:
            self._stack.enter_context(fused_elementwise(model))

    def _submit(self, i: int) -> None:
        if i not in self._submitted:
            self._submitted.add(i)
            self._mine[i] = self.reducer.submit(i, self.cov_mlp[i], self.cov_x[i], self.cov_q[i],
                                                self.cov_k[i])

    def _layer_done(self, i: int):
        def hook(_mod, _inp, _out):
            if self._armed:
                self._submit(i)
        return hook

    @torch.no_grad()
    def run_batch(self, batch: torch.Tensor, last: bool = False) -> None:
        if self._submitted:
            raise RuntimeError("Calibrator: a batch after the armed (last) one would be lost")
        self._armed = bool(last)
        batch = batch.to(self.device, non_blocking=True)
        self.n_texts += len(batch)
        self.body(batch, use_cache=False)
        # sum over the batch, mean over positions (src/calibration.py:122-124)
        self.bi_total += self.bi_batch / batch.shape[1]
        self.bi_batch.zero_()
        self._armed = False

    def close(self) -> None:
        for h in self.handles:
            h.remove()
        self.handles = []
        self._stack.close()

    @torch.no_grad()
    def finish(self):
        self.close()
        for i in self.targets:          # layers not exchanged during the last batch
            self._submit(i)
        self.reducer.wait()
        count = torch.tensor([float(self.n_texts)], dtype=torch.float64, device=self.device)
        D.all_reduce_sum_(count)
        D.all_reduce_sum_(self.bi_total)
        n_texts_global = int(count.item())
        bi_scores = (self.bi_total / n_texts_global).tolist()
        self.adapter.bi_scores = bi_scores
        scale = 1.0 / float(n_texts_global * NORMALISER_SEQ_LEN)
        for i in self.targets:
            if not self._mine[i]:       # only the owner decomposes layer i
                self.cov_mlp[i] = self.cov_q[i] = self.cov_k[i] = self.cov_x[i] = None
                continue
            ops.finalize_sym_(self.cov_mlp[i], scale)
            ops.finalize_sym_(self.cov_x[i], scale)
            ops.scale_(self.cov_q[i], scale)
            ops.scale_(self.cov_k[i], scale)
        return self.cov_mlp, self.cov_q, self.cov_k, self.cov_x, bi_scores


@torch.no_grad()
def warm_up(adapter: ModelAdapter, target_layers: list[int] | None = None, tokens: int = 64) -> None:
    """One-time process start-up, kept out of the calibration timing: a hooked forward of a few
    tokens through every target layer and a complete (tiny) exchange.  It loads every kernel
    module, creates the cuBLAS / NCCL state and — the bulk of it — makes torch's allocator obtain
    the accumulators' memory from the driver (17.8 GB for Llama-2-7B), which the real pass then
    reuses.  Measured on 2 GPUs: the first `load_calibs` of a process took 4.4 s, a warm one 2.05 s."""
    model = adapter.model
    device = next(model.parameters()).device
    cal = Calibrator(adapter, target_layers)
    try:
        ids = torch.zeros(1, tokens, dtype=torch.int64, device=device)
        cal.run_batch(ids, last=True)
        cal.finish()
    finally:
        cal.close()
    adapter.bi_scores = None
    if device.type == "cuda":
        torch.cuda.synchronize(device)


@torch.no_grad()
def _calibrate_model(adapter: ModelAdapter, n_samples: int, batch_size: int,
                     target_layers: list[int], dataset: str = "wikitext"):
    model = adapter.model
    if adapter.calibs is None:
        from .eval import load_calibration_texts

        adapter.calibs = load_calibration_texts(calib_size=n_samples, model=model,
                                                tokenizer=adapter.tokenizer, batch_size=batch_size,
                                                dataset=dataset, seq_len=adapter.config.seq_len,
                                                seed=adapter.config.seed)
    logger.info(f"Detected architecture: {adapter.arch}")
    logger.info(f"target_layers = {target_layers or list(range(adapter.n_layers))}")
    cal = Calibrator(adapter, target_layers)
    try:
        mine = D.shard_batches(adapter.calibs)
        for i, batch in enumerate(mine):
            cal.run_batch(batch, last=(i == len(mine) - 1))
        out = cal.finish()
    finally:
        cal.close()
    logger.info("Finished calibration and computed BI scores.")
    return out


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
        # the layers are called with attention_mask=None: sdpa / flash kernels are then causal
        # (is_causal), HF's eager attention would be BIDIRECTIONAL and silently change C and BI
        impl = getattr(model.config, "_attn_implementation", None) or "sdpa"
        if impl not in ("sdpa", "flash_attention_2", "flash_attention_3"):
            raise NotImplementedError(
                f"layer-streamed calibration needs attn_implementation sdpa or flash_attention_* "
                f"(got {impl!r}: without an explicit mask it would not be causal)")
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
    reducer = D.LayerReducer()
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
        mine = reducer.submit(l, cov[0][l], cov[3][l], cov[1][l], cov[2][l])
        reducer.wait()
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
