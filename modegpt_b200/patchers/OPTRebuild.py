"""Rank-aware OPT for checkpoints written by the MoDeGPT flow (self-contained; loaded through
`auto_map`).

The reference's OPTRebuild is stale — it reads `config.qk_ranks` / `vo_ranks`, which `patch_config`
no longer writes — so this one is derived from the ranks that ARE written
(`q/k/v/o_ranks`, `gate_ranks` for fc1).  OPT has no RoPE: no mask, and the softmax scale stays the
ORIGINAL head_dim ** -0.5 (the CR step approximates Q K^T, it does not renormalise it).
"""
import torch.nn as nn
from transformers.models.opt.modeling_opt import OPTAttention, OPTForCausalLM as _StockOPTForCausalLM


class CompressedOPTAttention(OPTAttention):
    def __init__(self, config, layer_idx: int):
        super().__init__(config, layer_idx=layer_idx)
        d, bias = config.hidden_size, config.enable_bias
        original_scaling = self.scaling
        self.head_dim = config.q_ranks[layer_idx] // config.num_attention_heads
        self.scaling = original_scaling
        self.q_proj = nn.Linear(d, config.q_ranks[layer_idx], bias=bias)
        self.k_proj = nn.Linear(d, config.k_ranks[layer_idx], bias=bias)
        self.v_proj = nn.Linear(d, config.v_ranks[layer_idx], bias=False)   # folded into out_proj bias
        self.out_proj = nn.Linear(config.o_ranks[layer_idx], d, bias=bias)


class OPTForCausalLM(_StockOPTForCausalLM):
    def __init__(self, config):
        super().__init__(config)
        d, bias = config.hidden_size, config.enable_bias
        for i, layer in enumerate(self.model.decoder.layers):
            layer.self_attn = CompressedOPTAttention(config, i)
            layer.fc1 = nn.Linear(d, config.gate_ranks[i], bias=bias)
            layer.fc2 = nn.Linear(config.gate_ranks[i], d, bias=bias)
        self.post_init()
