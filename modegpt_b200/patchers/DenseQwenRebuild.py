"""Rank-aware dense Qwen3 for checkpoints written by the MoDeGPT flow (self-contained; copied next
to the checkpoint and loaded through `auto_map`).

Same contract as LlamaRebuild (per-layer ranks, one compressed head dim for q/k/v, scale =
head_dim ** -0.5, masked RoPE) plus Qwen3's q_norm / k_norm, which act on the head dimension:
their weights keep the ORIGINAL head_dim and are gathered through the rotary mask, the RMS being
taken over the kept dimensions (reference: src/patchers/DenseQwenRebuild.py:262-286).
"""
import os
from typing import Callable, Optional

import torch
import torch.nn as nn
from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
from transformers.models.qwen3.modeling_qwen3 import (
    Qwen3Attention,
    Qwen3ForCausalLM as _StockQwen3ForCausalLM,
    Qwen3MLP,
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


class CompressedQwen3MLP(Qwen3MLP):
    def __init__(self, config, layer_idx: int):
        super().__init__(config)
        r, d = config.gate_ranks[layer_idx], config.hidden_size
        self.intermediate_size = r
        self.gate_proj = nn.Linear(d, r, bias=False)
        self.up_proj = nn.Linear(d, r, bias=False)
        self.down_proj = nn.Linear(r, d, bias=False)


class CompressedQwen3Attention(Qwen3Attention):
    def __init__(self, config, layer_idx: int, rotary_mask: Optional[torch.Tensor] = None):
        super().__init__(config, layer_idx)       # q_norm / k_norm keep the original head_dim
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

    @staticmethod
    def _masked_rms_norm(x, norm, mask):
        """x [B, T, heads, r]; mask [heads, r] into the norm weight's original head_dim."""
        xf = x.float()
        xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + norm.variance_epsilon)
        return (norm.weight[mask][None, None] * xf).to(x.dtype)

    def forward(self, hidden_states, position_embeddings=None, attention_mask=None,
                past_key_values=None, **kwargs):
        input_shape = hidden_states.shape[:-1]
        hidden_shape = (*input_shape, -1, self.head_dim)
        mask = self.rotary_mask(hidden_states.device)
        q = self.q_proj(hidden_states).view(hidden_shape)
        k = self.k_proj(hidden_states).view(hidden_shape)
        if mask is None:
            q, k = self.q_norm(q), self.k_norm(k)
        else:
            q = self._masked_rms_norm(q, self.q_norm, mask.repeat_interleave(self.num_key_value_groups, 0))
            k = self._masked_rms_norm(k, self.k_norm, mask)
        q, k = q.transpose(1, 2), k.transpose(1, 2)
        v = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        cos, sin = position_embeddings
        q, k = masked_rope(q, k, cos, sin, mask, self.num_key_value_groups)
        if past_key_values is not None:
            k, v = past_key_values.update(k, v, self.layer_idx)
        attend: Callable = ALL_ATTENTION_FUNCTIONS.get_interface(self.config._attn_implementation,
                                                                 eager_attention_forward)
        out, weights = attend(self, q, k, v, attention_mask,
                              dropout=0.0 if not self.training else self.attention_dropout,
                              scaling=self.scaling, sliding_window=self.sliding_window, **kwargs)
        out = out.reshape(*input_shape, -1).contiguous()
        return self.o_proj(out), weights


class Qwen3ForCausalLM(_StockQwen3ForCausalLM):
    def __init__(self, config):
        super().__init__(config)
        masks = load_rotary_masks(config)
        for i, layer in enumerate(self.model.layers):
            layer.self_attn = CompressedQwen3Attention(config, i, None if masks is None else masks[i])
            layer.mlp = CompressedQwen3MLP(config, i)
        self.post_init()
