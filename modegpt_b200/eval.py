"""Calibration batches and perplexity (reference: src/eval.py).

Contracts kept: calibration data is a list of `[batch, seq]` int64 token batches
(src/eval.py:33-68), perplexity is exp(sum_b CE_b * (T-1) * B_b / (N * (T-1))) over fixed-length
sequences (src/eval.py:192-220).  The reference only knows hub datasets; with no network this
build adds `dataset="synthetic"`: seeded uniform token ids (calibration seed 1234 like the
reference's RNG seeds, held-out seed 4321).
"""
from __future__ import annotations

import logging
import time

import torch

logger = logging.getLogger("MoDeGPT")

HELD_OUT_SEED = 4321


def synthetic_tokens(n: int, seq_len: int, vocab: int, seed: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, vocab, (n, seq_len), generator=g, dtype=torch.int64)


def load_calibration_texts(calib_size, model, tokenizer, batch_size: int, dataset="wikitext",
                           seq_len: int = 2048, seed: int = 1234) -> list[torch.Tensor]:
    device = next(model.parameters()).device
    if dataset == "synthetic":
        tokens = synthetic_tokens(int(calib_size), seq_len, model.config.vocab_size, seed)
    else:
        tokens = _hub_tokens(dataset, tokenizer, model, split="train", n=int(calib_size), seed=seed)
    return [tokens[i:i + batch_size].to(device) for i in range(0, tokens.shape[0], batch_size)]


def _hub_tokens(dataset: str, tokenizer, model, split: str, n: int, seed: int) -> torch.Tensor:
    """wikitext / c4 through `datasets` (needs network and a tokenizer; src/eval.py:40-68)."""
    if tokenizer is None:
        raise ValueError(f"dataset={dataset!r} needs a tokenizer; use dataset='synthetic' offline")
    from datasets import load_dataset  # deferred: optional dependency

    if dataset == "wikitext":
        text = "\n\n".join(load_dataset("wikitext", "wikitext-2-raw-v1", split=split)["text"])
    elif dataset == "c4":
        name = "train" if split == "train" else "validation"
        shard = "c4-train.00000-of-01024" if split == "train" else "c4-validation.00000-of-00008"
        ds = load_dataset("json", data_files={name: "https://huggingface.co/datasets/allenai/c4/"
                                              f"resolve/main/en/{shard}.json.gz"})
        text = "\n\n".join([t for t in ds[name]["text"] if t.strip()][:10000])
    else:
        raise ValueError(f"Unknown dataset: {dataset}")
    ids = tokenizer(text, truncation=False, return_tensors="pt", add_special_tokens=False)["input_ids"][0]
    width = min(2048, model.config.max_position_embeddings)
    chunks = ids[: ids.numel() // width * width].view(-1, width)
    g = torch.Generator().manual_seed(seed)
    pick = torch.randperm(chunks.shape[0], generator=g)[: min(n, chunks.shape[0])]
    return chunks[pick]


def _ce_chunk_rows() -> int:
    return 8192          # rows of logits alive at once: 8192 x vocab bf16 (0.5 GB at vocab 32000)


@torch.no_grad()
def sequence_nll(model, x: torch.Tensor) -> torch.Tensor:
    """Sum over the batch of next-token negative log-likelihoods (fp64 scalar on the device) for
    token ids x [B, T], without materialising [B, T, vocab] logits: the decoder body runs once,
    then row chunks of `hidden @ W_lm^T` (bf16) go through `mg_ce_rows_bf16`.  Equal to
    CrossEntropyLoss(logits[:, :-1].float(), x[:, 1:]) * (T - 1) * B  (src/eval.py:205-212)."""
    from . import ops

    body = getattr(model, model.base_model_prefix, None)
    head = model.get_output_embeddings()
    fast = (body is not None and head is not None and x.is_cuda
            and head.weight.dtype == torch.bfloat16)
    if not fast:   # CPU / non-bf16 models (CPU tests): the plain formula
        logits = model(x, use_cache=False).logits
        loss = torch.nn.functional.cross_entropy(
            logits[:, :-1, :].reshape(-1, logits.size(-1)).float(), x[:, 1:].reshape(-1), reduction="sum")
        return loss.double()
    hidden = body(x, use_cache=False).last_hidden_state            # [B, T, d], final norm applied
    B, T, d = hidden.shape
    h = hidden[:, :-1, :].reshape(-1, d)
    labels = x[:, 1:].reshape(-1).contiguous()
    total = torch.zeros((), dtype=torch.float64, device=x.device)
    step = _ce_chunk_rows()
    for r0 in range(0, h.shape[0], step):
        logits = torch.nn.functional.linear(h[r0:r0 + step], head.weight, head.bias)
        total += ops.ce_rows(logits, labels[r0:r0 + step]).double().sum()
        del logits
    return total


@torch.no_grad()
def compute_perplexity(model, tokenizer, bs: int = 16, device="cuda", dataset="wikitext",
                       adapter=None, n_samples: int | None = None, seq_len: int | None = None,
                       shard: bool = True) -> float:
    """exp(sum of token NLLs / (N * (T - 1))) over fixed-length sequences (src/eval.py:134-225).
    Differences underneath: chunked cross-entropy from bf16 logits (`sequence_nll`), the fused
    elementwise / rebuilt-model kernels during the forward, and — with torch.distributed
    initialised and shard=True — the sequences are dealt to ranks and the NLL sum is all-reduced
    (every rank returns the same perplexity; all ranks must call)."""
    import contextlib

    from . import distributed as D
    from .fused_forward import fused_elementwise, fused_rebuilt

    model.eval()
    dev = next(model.parameters()).device
    cfg = adapter.config if adapter is not None else None
    seq_len = seq_len or (cfg.seq_len if cfg else 2048)
    if dataset == "synthetic":
        n = n_samples or (cfg.eval_samples if cfg else 16)
        tokens = synthetic_tokens(n, seq_len, model.config.vocab_size, HELD_OUT_SEED)
    else:
        tokens = _hub_tokens(dataset, tokenizer, model, split="test", n=n_samples or 512, seed=0)
        seq_len = tokens.shape[1]
    n = tokens.shape[0]
    world, rank = (D.world_size(), D.rank()) if shard else (1, 0)
    mine = tokens[rank::world]
    nll = torch.zeros((), dtype=torch.float64, device=dev)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    fused = dev.type == "cuda" and not (cfg is not None and cfg.eager_forward)
    with (fused_elementwise(model) if fused else contextlib.nullcontext()), \
            (fused_rebuilt(model) if fused else contextlib.nullcontext()):
        for i in range(0, mine.shape[0], bs):
            nll += sequence_nll(model, mine[i:i + bs].to(dev))
    if world > 1:
        D.all_reduce_sum_(nll)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    elapsed = time.perf_counter() - t0
    if adapter is not None:
        adapter.metrics["throughput_tok/s"] = n * seq_len / max(elapsed, 1e-9)
    return float(torch.exp(nll / (n * (seq_len - 1))).item())
