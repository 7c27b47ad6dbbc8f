"""Model I/O for the compression flow (reference: src/model_utils.py).

`reload_compressed_model` / `save_compressed_model` keep the reference's call shapes.  Additions:
`synthetic:<preset>[:<n_layers>]` model names build random-init models of the real shapes (no hub access), the
tokenizer is optional, and the rotary-mask path is stored relative to the checkpoint so a saved
model can be moved (the reference stores an absolute path).
"""
from __future__ import annotations

import logging
import os
import shutil
from pathlib import Path

import torch

logger = logging.getLogger("MoDeGPT")

# (model_type, hidden, heads, kv_heads, head_dim, intermediate, layers, vocab)
PRESETS = {
    "opt-125m": ("opt", 768, 12, 12, 64, 3072, 12, 50272),
    "llama-2-7b": ("llama", 4096, 32, 32, 128, 11008, 32, 32000),
    "llama-3-8b": ("llama", 4096, 32, 8, 128, 14336, 32, 128256),
    "qwen3-8b": ("qwen3", 4096, 32, 8, 128, 12288, 36, 151936),
    "qwen2.5-7b": ("qwen2", 3584, 28, 4, 128, 18944, 28, 152064),
    "llama-2-70b": ("llama", 8192, 64, 8, 128, 28672, 80, 32000),
    "tiny-llama": ("llama", 256, 4, 4, 64, 512, 3, 512),
    "tiny-llama-gqa": ("llama", 256, 4, 2, 64, 512, 3, 512),
    "tiny-qwen3": ("qwen3", 256, 4, 2, 64, 512, 3, 512),
    "tiny-qwen2": ("qwen2", 256, 4, 2, 64, 512, 3, 512),
    "tiny-opt": ("opt", 256, 4, 4, 64, 512, 3, 512),
    "tiny-llama-hd80": ("llama", 320, 4, 2, 80, 512, 3, 512),     # a head dim the 32/64/128 tile sets do not cover
    "tiny-opt-hd80": ("opt", 320, 4, 4, 80, 512, 3, 512),         # OPT-2.7b's head dim
}


def synthetic_config(preset: str, n_layers: int | None = None, max_positions: int = 2048):
    from transformers import LlamaConfig, OPTConfig, Qwen2Config, Qwen3Config

    kind, d, H, KV, hd, d_int, L, vocab = PRESETS[preset]
    L = n_layers or L
    if kind == "opt":
        return OPTConfig(hidden_size=d, num_attention_heads=H, ffn_dim=d_int, num_hidden_layers=L,
                         vocab_size=vocab, max_position_embeddings=max_positions,
                         word_embed_proj_dim=d, do_layer_norm_before=True)
    if kind == "qwen2":     # Qwen2Config has no head_dim: hidden / heads
        return Qwen2Config(hidden_size=d, num_attention_heads=H, num_key_value_heads=KV,
                           intermediate_size=d_int, num_hidden_layers=L, vocab_size=vocab,
                           max_position_embeddings=max_positions, tie_word_embeddings=False,
                           use_sliding_window=False)
    cls = Qwen3Config if kind == "qwen3" else LlamaConfig
    return cls(hidden_size=d, num_attention_heads=H, num_key_value_heads=KV, head_dim=hd,
               intermediate_size=d_int, num_hidden_layers=L, vocab_size=vocab,
               max_position_embeddings=max_positions, tie_word_embeddings=False)


def build_synthetic_model(preset: str, device="cuda:0", seed: int = 0, n_layers: int | None = None,
                          max_positions: int = 2048):
    """Random-init model of a preset's shape, HF default init, cast to bf16."""
    from transformers import AutoModelForCausalLM

    torch.manual_seed(seed)
    cfg = synthetic_config(preset, n_layers, max_positions)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            model = AutoModelForCausalLM.from_config(cfg)
    finally:
        torch.set_default_dtype(prev)
    return model.eval()


def reload_compressed_model(model_dir: str, device="cuda:0", tokenizer_source: str = ""):
    """(model, tokenizer-or-None).  `model_dir` is a hub id, a local checkpoint (plain or written by
    `save_compressed_model`) or `synthetic:<preset>`."""
    from transformers import AutoModelForCausalLM, AutoTokenizer

    logger.info(f"Loading model from: {model_dir}")
    if model_dir.startswith("synthetic:"):
        # synthetic:<preset>[:<n_layers>] — real widths, optionally fewer layers (70B-class shapes
        # on one GPU)
        parts = model_dir.split(":")
        n_layers = int(parts[2]) if len(parts) > 2 and parts[2] else None
        return build_synthetic_model(parts[1], device=device, n_layers=n_layers), None
    src = tokenizer_source
    if not src:
        marker = os.path.join(model_dir, "tokenizer_source.txt")
        src = Path(marker).read_text().strip() if os.path.exists(marker) else model_dir
    tokenizer = None
    if src and not src.startswith("synthetic:"):
        try:
            tokenizer = AutoTokenizer.from_pretrained(src)
            if tokenizer.pad_token is None:
                tokenizer.pad_token = tokenizer.eos_token
        except Exception as e:  # offline / tokenizer-less synthetic checkpoints
            logger.warning(f"no tokenizer loaded from {src!r}: {e}")
    # The kernels take bf16 activations and weights (north_star: "bf16 activations, fp32
    # accumulate") and the flow saves bf16 (the reference's convert_model creates bf16 Linears,
    # model_adapter.py:199-208), so fp16 / fp32 checkpoints (facebook/opt-*, Llama-2-*-hf) are cast
    # once at load instead of failing at the first calibration hook.
    model = AutoModelForCausalLM.from_pretrained(model_dir, trust_remote_code=True, dtype=torch.bfloat16)
    model.to(device)
    bad = {p.dtype for p in model.parameters() if p.is_floating_point() and p.dtype != torch.bfloat16}
    if bad:
        logger.info(f"casting {sorted(map(str, bad))} parameters to bfloat16")
        model.to(torch.bfloat16)
    model.eval()
    return model, tokenizer


def save_compressed_model(adapter, rotary_masks, save_dir: str, source_model_name: str):
    """save_pretrained + rotary_masks.pt + the arch's Rebuild module + tokenizer_source.txt
    (src/model_utils.py:83-126)."""
    model, tokenizer = adapter.model, adapter.tokenizer
    rebuild = Path(__file__).resolve().parent / "patchers" / f"{adapter.rebuild_module}.py"
    if not rebuild.exists():
        raise RuntimeError(f"no compressed model definition for arch {adapter.arch!r}")
    os.makedirs(save_dir, exist_ok=True)
    have_masks = rotary_masks is not None and len(rotary_masks) > 0
    model.config.mask_path = "rotary_masks.pt" if have_masks else None
    model.config.dtype = "bfloat16"
    logger.info(f"Saving compressed model to {save_dir}")
    model.save_pretrained(save_dir)
    if tokenizer is not None:
        tokenizer.save_pretrained(save_dir)
    if have_masks:
        torch.save([m.cpu() for m in rotary_masks], os.path.join(save_dir, "rotary_masks.pt"))
    shutil.copy(rebuild, save_dir)
    Path(save_dir, "tokenizer_source.txt").write_text(source_model_name.strip())
