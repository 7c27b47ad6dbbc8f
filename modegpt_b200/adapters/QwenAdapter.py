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


class Qwen2Adapter(LlamaAdapter):
    """Qwen2 / Qwen2.5 (BASELINE config #4).  The reference routes `model_type == "qwen2"` to its
    Llama adapter, which then fails on `config.head_dim` and would drop the q/k/v biases
    (SURVEY A.4).  Here the q/k biases are gathered with their rows (they are part of the raw
    projections the statistics see) and the v bias — which reaches the output as the constant
    W_o b_v because attention weights sum to one — is folded exactly into an o_proj bias."""
    rebuild_module = "Qwen2Rebuild"
    rebuild_class = "Qwen2ForCausalLM"

    @property
    def arch(self) -> str:
        return "qwen2"
