"""Llama-family adapter (Llama-2/3, and anything `from_model` routes here by default).

Module table for HF `LlamaForCausalLM` (reference: src/adapters/LlamaAdapter.py:35-248).  All
behaviour lives in `ModelAdapter`; the statistics hooks it registers sit where the reference's do:
pre-hook on `mlp.down_proj`, forward hooks on `input_layernorm`, `self_attn.q_proj`, `self_attn.k_proj`
(src/adapters/LlamaAdapter.py:71-100).
"""
from __future__ import annotations

import torch.nn as nn

from .model_adapter import ModelAdapter, ModuleMap


class LlamaAdapter(ModelAdapter):
    rebuild_module = "LlamaRebuild"
    rebuild_class = "LlamaForCausalLM"

    _MAP = ModuleMap(up="mlp.up_proj", down="mlp.down_proj", gate="mlp.gate_proj",
                     q="self_attn.q_proj", k="self_attn.k_proj", v="self_attn.v_proj",
                     o="self_attn.o_proj", attn_in_norm="input_layernorm", final_norm="model.norm")

    @property
    def arch(self) -> str:
        return "llama"

    @property
    def module_map(self) -> ModuleMap:
        return self._MAP

    def get_transformer_blocks(self) -> nn.ModuleList:
        return self.model.model.layers
