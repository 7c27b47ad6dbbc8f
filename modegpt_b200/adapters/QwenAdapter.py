"""Dense Qwen3 adapter: Llama layout, `arch == "qwen3"` (reference: src/adapters/QwenAdapter.py:1-9).
The q_norm / k_norm weights are not touched by compression; the rebuilt attention gathers them
through the rotary mask (patchers/DenseQwenRebuild.py)."""
from __future__ import annotations

from .LlamaAdapter import LlamaAdapter


class QwenAdapter(LlamaAdapter):
    rebuild_module = "DenseQwenRebuild"
    rebuild_class = "Qwen3ForCausalLM"

    @property
    def arch(self) -> str:
        return "qwen3"
