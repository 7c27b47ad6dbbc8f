"""OPT adapter (BASELINE config #1).

The reference's OPTAdapter cannot be instantiated in HEAD (missing abstract methods, signature
mismatch, C_x hook never fired, no patch_config — SURVEY Appendix A.3).  This adapter implements
the semantics its surviving pieces describe:
  * C_mlp from relu(fc1(x))            (src/adapters/model_adapter.py:546-554) — that tensor IS the
    input of fc2, so the same down-projection pre-hook serves;
  * per-head C_q / C_k from the raw q_proj / k_proj outputs, biases included
    (src/adapters/model_adapter.py:556-567);
  * C_x from the attention input, i.e. the output of `self_attn_layer_norm`
    (the role `on_batch_end_step` was meant to play, src/adapters/OPTAdapter.py:45-46);
  * biases: fc1 / q / k biases are gathered with their rows, fc2 / out_proj biases are kept
    (src/adapters/model_adapter.py:442-452,529-538).
"""
from __future__ import annotations

import torch.nn as nn

from .model_adapter import ModelAdapter, ModuleMap


class OPTAdapter(ModelAdapter):
    rebuild_module = "OPTRebuild"
    rebuild_class = "OPTForCausalLM"

    _MAP = ModuleMap(up="fc1", down="fc2", gate=None, q="self_attn.q_proj", k="self_attn.k_proj",
                     v="self_attn.v_proj", o="self_attn.out_proj",
                     attn_in_norm="self_attn_layer_norm",
                     final_norm="model.decoder.final_layer_norm")

    @property
    def arch(self) -> str:
        return "opt"

    @property
    def module_map(self) -> ModuleMap:
        return self._MAP

    def get_transformer_blocks(self) -> nn.ModuleList:
        return self.model.model.decoder.layers
