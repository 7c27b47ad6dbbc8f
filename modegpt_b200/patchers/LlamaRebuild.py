"""Rank-aware Llama for checkpoints written by the MoDeGPT flow (loaded through `auto_map` +
`trust_remote_code`; this file is copied next to the checkpoint and must stay self-contained).

Re-authored against transformers >= 5 as thin subclasses of the stock modules — the reference
ships a full copy of the 4.5x modeling file (src/patchers/LlamaRebuild.py) that no longer loads
(SURVEY §2 row 12).  Behaviour kept from it:
  * per-layer widths from config.{q,k,v,o,gate}_ranks;
  * ONE head dim per layer, `q_rank // n_heads`, used for q, k and v (LlamaRebuild.py:266,326-328);
  * softmax scale = that compressed head dim ** -0.5 (LlamaRebuild.py:282);
  * RoPE applied on the kept dimensions only: cos / sin gathered through the layer's rotary mask
    [n_kv_heads, r], query heads using their kv head's mask (LlamaRebuild.py:155-180).
The masks are read from `config.mask_path` (a file name is resolved against the checkpoint dir).
"""
import os
from typing import Callable, Optional

import torch
import torch.nn as nn
from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
from transformers.models.llama.modeling_llama import (
    LlamaAttention,
    LlamaForCausalLM as _StockLlamaForCausalLM,
    LlamaMLP,
    eager_attention_forward,
    rotate_half,
)


def masked_rope(q, k, cos, sin, mask: Optional[torch.Tensor], groups: int):
    """q [B,H,T,r], k [B,KV,T,r]; cos/sin [B,T,hd]; mask [KV,r] indices into hd."""
    if mask is None:
        cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
        return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin
    cos_k = cos[:, :, mask].permute(0, 2, 1, 3)       # [B, KV, T, r]
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


class CompressedLlamaMLP(LlamaMLP):
    def __init__(self, config, layer_idx: int):
        super().__init__(config)
        r, d, bias = config.gate_ranks[layer_idx], config.hidden_size, config.mlp_bias
        self.intermediate_size = r
        self.gate_proj = nn.Linear(d, r, bias=bias)
        self.up_proj = nn.Linear(d, r, bias=bias)
        self.down_proj = nn.Linear(r, d, bias=bias)


class CompressedLlamaAttention(LlamaAttention):
    def __init__(self, config, layer_idx: int, rotary_mask: Optional[torch.Tensor] = None):
        super().__init__(config, layer_idx)
        d, bias = config.hidden_size, config.attention_bias
        self.head_dim = config.q_ranks[layer_idx] // config.num_attention_heads
        self.scaling = self.head_dim ** -0.5
        self.q_proj = nn.Linear(d, config.q_ranks[layer_idx], bias=bias)
        self.k_proj = nn.Linear(d, config.k_ranks[layer_idx], bias=bias)
        self.v_proj = nn.Linear(d, config.v_ranks[layer_idx], bias=bias)
        self.o_proj = nn.Linear(config.o_ranks[layer_idx], d, bias=bias)
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
                              scaling=self.scaling, **kwargs)
        out = out.reshape(*input_shape, -1).contiguous()
        return self.o_proj(out), weights


class LlamaForCausalLM(_StockLlamaForCausalLM):
    def __init__(self, config):
        super().__init__(config)
        masks = load_rotary_masks(config)
        for i, layer in enumerate(self.model.layers):
            layer.self_attn = CompressedLlamaAttention(config, i, None if masks is None else masks[i])
            layer.mlp = CompressedLlamaMLP(config, i)
        self.post_init()
