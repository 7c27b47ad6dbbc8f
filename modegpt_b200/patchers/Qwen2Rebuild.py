"""Rank-aware Qwen2 / Qwen2.5 for checkpoints written by the MoDeGPT flow (self-contained; copied
next to the checkpoint and loaded through `auto_map`).

Same contract as LlamaRebuild (per-layer ranks, one compressed head dim for q/k/v, scale =
head_dim ** -0.5, masked RoPE).  Qwen2 specifics: q_proj / k_proj keep their (gathered) biases,
v_proj loses its bias and o_proj gains one (the v bias folded through W_o by the compression).
"""
import os
from typing import Callable, Optional

import torch
import torch.nn as nn
from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
from transformers.models.qwen2.modeling_qwen2 import (
    Qwen2Attention,
    Qwen2ForCausalLM as _StockQwen2ForCausalLM,
    Qwen2MLP,
    eager_attention_forward,
    rotate_half,
)


def masked_rope(q, k, cos, sin, mask, groups: int):
    if mask is None:
        cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
        return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin
    cos_k = cos[:, :, mask].permute(0, 2, 1, 3)
    sin_k = sin[:, :, mask].permute(0, 2, 1, 3)
    cos_q = cos_k.repeat_interleave(groups, dim=1)
    sin_q = sin_k.repeat_interleave(groups, dim=1)
    return q * cos_q + rotate_half(q) * sin_q, k * cos_k + rotate_half(k) * sin_k


def load_rotary_masks(config):
    path = getattr(config, "mask_path", None)
    if not path:
        return None
    if not os.path.isabs(path):
        path = os.path.join(getattr(config, "_name_or_path", "") or ".", path)
    return torch.load(path, map_location="cpu")


class CompressedQwen2MLP(Qwen2MLP):
    def __init__(self, config, layer_idx: int):
        super().__init__(config)
        r, d = config.gate_ranks[layer_idx], config.hidden_size
        self.intermediate_size = r
        self.gate_proj = nn.Linear(d, r, bias=False)
        self.up_proj = nn.Linear(d, r, bias=False)
        self.down_proj = nn.Linear(r, d, bias=False)


class CompressedQwen2Attention(Qwen2Attention):
    def __init__(self, config, layer_idx: int, rotary_mask: Optional[torch.Tensor] = None):
        super().__init__(config, layer_idx)
        d = config.hidden_size
        self.head_dim = config.q_ranks[layer_idx] // config.num_attention_heads
        self.scaling = self.head_dim ** -0.5
        self.q_proj = nn.Linear(d, config.q_ranks[layer_idx], bias=True)
        self.k_proj = nn.Linear(d, config.k_ranks[layer_idx], bias=True)
        self.v_proj = nn.Linear(d, config.v_ranks[layer_idx], bias=False)
        self.o_proj = nn.Linear(config.o_ranks[layer_idx], d, bias=True)
        self._mask_cpu = rotary_mask
        self._mask_dev = None

    def rotary_mask(self, device):
        if self._mask_cpu is None:
            return None
        if self._mask_dev is None or self._mask_dev.device != device:
            self._mask_dev = self._mask_cpu.to(device=device, dtype=torch.long)
        return self._mask_dev

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None,
                past_key_values=None, **kwargs):
        input_shape = hidden_states.shape[:-1]
        hidden_shape = (*input_shape, -1, self.head_dim)
        q = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        k = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        v = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        cos, sin = position_embeddings
        q, k = masked_rope(q, k, cos, sin, self.rotary_mask(hidden_states.device),
                           self.num_key_value_groups)
        if past_key_values is not None:
            k, v = past_key_values.update(k, v, self.layer_idx)
        attend: Callable = ALL_ATTENTION_FUNCTIONS.get_interface(self.config._attn_implementation,
                                                                 eager_attention_forward)
        out, weights = attend(self, q, k, v, attention_mask,
                              dropout=0.0 if not self.training else self.attention_dropout,
                              scaling=self.scaling, sliding_window=getattr(self, "sliding_window", None),
                              **kwargs)
        out = out.reshape(*input_shape, -1).contiguous()
        return self.o_proj(out), weights


class Qwen2ForCausalLM(_StockQwen2ForCausalLM):
    def __init__(self, config):
        super().__init__(config)
        masks = load_rotary_masks(config)
        for i, layer in enumerate(self.model.layers):
            layer.self_attn = CompressedQwen2Attention(config, i, None if masks is None else masks[i])
            layer.mlp = CompressedQwen2MLP(config, i)
        self.post_init()
