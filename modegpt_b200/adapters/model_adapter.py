"""ModelAdapter: the seam between the HF model and the compression hot path.

Interface parity with the reference ABC (src/adapters/model_adapter.py:96-392): same property and
method names, same argument meaning, `register_hooks` mutates the statistics lists in place and
appends removable handles, `save_layer` / `convert_model` keep the `layer_{i}_{suffix}` file format.

What differs underneath:
  * module lookup is table driven (`ModuleMap`) — one concrete implementation serves Llama, Qwen3
    and OPT; the per-architecture subclasses only supply the table;
  * the hooks hand the live bf16 activation to the tcgen05 SYRK kernels (no fp64 copy, no
    permute, no per-sample [B, d, d] temporaries);
  * Block-Influence is accumulated by hooks on the blocks themselves (one fused cosine kernel per
    block per batch, a single device->host read at the end) instead of `output_hidden_states`.
"""
from __future__ import annotations

import json
import logging
import os
from abc import ABC, abstractmethod
from dataclasses import dataclass
from datetime import datetime
from typing import Any, List, Optional, Tuple

import torch
import torch.nn as nn
from torch import Tensor

from .. import ops
from .CompressionConfig import CompressionConfig


# ---- value objects (names and fields as in the reference, model_adapter.py:19-82) --------------
@dataclass
class MLPTensors:
    up_proj: Tensor
    down_proj: Tensor
    gate_proj: Optional[Tensor]

    def to(self, dtype):
        self.up_proj = self.up_proj.to(dtype=dtype)
        self.down_proj = self.down_proj.to(dtype=dtype)
        if self.gate_proj is not None:
            self.gate_proj = self.gate_proj.to(dtype=dtype)
        return self


@dataclass
class VOTensors:
    v_proj: Tensor
    o_proj: Tensor

    def to(self, dtype):
        self.v_proj, self.o_proj = self.v_proj.to(dtype=dtype), self.o_proj.to(dtype=dtype)
        return self


@dataclass
class QKTensors:
    query_proj: Tensor
    key_proj: Tensor

    def to(self, dtype):
        self.query_proj = self.query_proj.to(dtype=dtype)
        self.key_proj = self.key_proj.to(dtype=dtype)
        return self


@dataclass
class MLPComponents:
    block: Optional[nn.Module]
    up_proj: nn.Module
    down_proj: nn.Module
    gate_proj: Optional[nn.Module] = None


@dataclass
class QKComponents:
    block: Optional[nn.Module]
    query_proj: nn.Module
    key_proj: nn.Module


@dataclass
class VOComponents:
    block: Optional[nn.Module]
    v_proj: nn.Module
    o_proj: nn.Module


@dataclass
class AttentionComponents:
    block: nn.Module
    q_proj: nn.Module
    k_proj: nn.Module
    v_proj: Optional[nn.Module] = None
    o_proj: Optional[nn.Module] = None


@dataclass(frozen=True)
class ModuleMap:
    """Dotted paths, relative to a transformer block, of the modules the path touches."""
    up: str
    down: str
    gate: Optional[str]
    q: str
    k: str
    v: str
    o: str
    attn_in_norm: str   # its OUTPUT is the attention input whose Gram is C_x
    final_norm: str     # relative to the MODEL: output of this module ends the last block's BI pair


def _get(root: nn.Module, path: str) -> nn.Module:
    for part in path.split("."):
        root = getattr(root, part)
    return root


def _set(root: nn.Module, path: str, value: nn.Module) -> None:
    *parents, leaf = path.split(".")
    for part in parents:
        root = getattr(root, part)
    setattr(root, leaf, value)


class ModelAdapter(ABC):
    _metrics: dict = {}

    def __init__(self, model: nn.Module, tokenizer=None):
        self.model = model
        self.model_config = model.config
        self.config: CompressionConfig = CompressionConfig()
        self.tokenizer = tokenizer
        self.calibs = None            # pre-seeded list of [B, T] int64 batches (synthetic seam)
        self.bi_scores = None
        self._layer_store: dict[tuple[int, str], dict[str, Tensor]] = {}
        self._layer_cache: dict[tuple[int, str], dict[str, Tensor]] = {}   # device copies of written layers
        self._writer = None           # handoff.LayerWriter, created on the first asynchronous save
        self.metrics = {
            "RunName": datetime.now().strftime("%Y_%m_%d--%H_%M_%S"),
            "RunDate": datetime.now().strftime("%b %d, %Y %I:%M %p"),
        }
        ModelAdapter._metrics[self.metrics["RunName"]] = self.metrics

    # ---- factory (model_adapter.py:118-135) ---------------------------------------------------
    @staticmethod
    def from_model(model: nn.Module, tokenizer=None) -> "ModelAdapter":
        from .LlamaAdapter import LlamaAdapter
        from .OPTAdapter import OPTAdapter
        from .QwenAdapter import Qwen2Adapter, QwenAdapter

        inner = getattr(model, "model", None)
        if inner is not None and hasattr(inner, "decoder"):
            return OPTAdapter(model, tokenizer=tokenizer)
        if inner is not None and hasattr(inner, "layers"):
            model_type = getattr(model.config, "model_type", "") or ""
            if "qwen3" in model_type:
                return QwenAdapter(model, tokenizer=tokenizer)
            if "qwen2" in model_type:
                return Qwen2Adapter(model, tokenizer=tokenizer)
            return LlamaAdapter(model, tokenizer=tokenizer)
        raise RuntimeError("Unsupported model architecture")

    # ---- architecture table -------------------------------------------------------------------
    @property
    @abstractmethod
    def module_map(self) -> ModuleMap:
        ...

    @abstractmethod
    def get_transformer_blocks(self) -> nn.ModuleList:
        ...

    # ---- shape properties (model_adapter.py:253-294) ------------------------------------------
    @property
    def arch(self) -> str:
        return self.model_config.model_type

    @property
    def n_layers(self) -> int:
        c = self.model_config
        return getattr(c, "n_layer", None) or getattr(c, "num_hidden_layers", None) or c.num_layers

    @property
    def n_heads(self) -> int:
        c = self.model_config
        return getattr(c, "n_head", None) or c.num_attention_heads

    @property
    def n_kv_heads(self) -> int:
        return getattr(self.model_config, "num_key_value_heads", None) or self.n_heads

    @property
    def d_model(self) -> int:
        c = self.model_config
        return getattr(c, "hidden_size", None) or c.dim

    @property
    def d_int(self) -> int:
        c = self.model_config
        return getattr(c, "intermediate_size", None) or getattr(c, "ffn_dim", None)

    @property
    def head_dim(self) -> int:
        # the reference reads config.head_dim only, which OPTConfig / Qwen2Config lack
        return getattr(self.model_config, "head_dim", None) or self.d_model // self.n_heads

    @property
    def n_experts(self) -> int:
        return getattr(self.model_config, "num_local_experts", 0) or 0

    def get_n_inner(self) -> int:
        return self.get_mlp_components(0).up_proj.out_features

    @property
    def uses_rope(self) -> bool:
        return self.arch == "llama" or "qwen" in self.arch

    # ---- component getters (same names as the reference's abstract methods) ------------------
    def validate_for_kernels(self) -> None:
        """Fail before calibration, with a readable message, on shapes the kernels do not take
        (the C ABI would otherwise return MG status -6 / -11 in the middle of a run)."""
        hd = self.head_dim
        if hd > 128 or hd % 4 != 0:
            raise NotImplementedError(
                f"head_dim = {hd}: the type-II / type-III kernels support head dims that are "
                f"multiples of 4 up to 128 (per-head eigenproblems live in one SM's shared memory)")
        if self.n_heads % self.n_kv_heads:
            raise NotImplementedError("n_heads must be a multiple of n_kv_heads")
        dt = next(self.model.parameters()).dtype
        if dt != torch.bfloat16:
            raise TypeError(f"model parameters are {dt}: load through reload_compressed_model "
                            f"(casts to bfloat16) or call model.to(torch.bfloat16)")

    def _block(self, layer_idx: int) -> nn.Module:
        return self.get_transformer_blocks()[layer_idx]

    def get_mlp_components(self, layer_idx: int, expert_idx: int | None = None) -> MLPComponents:
        b, m = self._block(layer_idx), self.module_map
        return MLPComponents(block=b, up_proj=_get(b, m.up), down_proj=_get(b, m.down),
                             gate_proj=_get(b, m.gate) if m.gate else None)

    def get_mlp_tensors(self, layer_idx: int, expert_idx: int | None = None) -> MLPTensors:
        c = self.get_mlp_components(layer_idx)
        return MLPTensors(up_proj=c.up_proj.weight, down_proj=c.down_proj.weight,
                          gate_proj=c.gate_proj.weight if c.gate_proj is not None else None)

    def get_qk_components(self, layer_idx: int, expert_idx: int | None = None) -> QKComponents:
        b, m = self._block(layer_idx), self.module_map
        return QKComponents(block=b, query_proj=_get(b, m.q), key_proj=_get(b, m.k))

    def get_qk_tensors(self, layer_idx: int, expert_idx: int | None = None) -> QKTensors:
        c = self.get_qk_components(layer_idx)
        return QKTensors(query_proj=c.query_proj.weight, key_proj=c.key_proj.weight)

    def get_vo_components(self, layer_idx: int, expert_idx: int | None = None) -> VOComponents:
        b, m = self._block(layer_idx), self.module_map
        return VOComponents(block=b, v_proj=_get(b, m.v), o_proj=_get(b, m.o))

    def get_vo_tensors(self, layer_idx: int, expert_idx: int | None = None) -> VOTensors:
        c = self.get_vo_components(layer_idx)
        return VOTensors(v_proj=c.v_proj.weight, o_proj=c.o_proj.weight)

    def get_attn_components(self, layer_idx: int) -> AttentionComponents:
        b, m = self._block(layer_idx), self.module_map
        return AttentionComponents(block=b, q_proj=_get(b, m.q), k_proj=_get(b, m.k),
                                   v_proj=_get(b, m.v), o_proj=_get(b, m.o))

    def get_qk_weights(self, layer_idx: int) -> Tuple[Tensor, Tensor]:
        t = self.get_qk_tensors(layer_idx)
        return t.query_proj, t.key_proj

    def get_vo_weights(self, layer_idx: int) -> Tuple[Tensor, Tensor]:
        t = self.get_vo_tensors(layer_idx)
        return t.v_proj, t.o_proj

    def replace_mlp_layers(self, layer_idx: int, new_up: nn.Module, new_down: nn.Module,
                           new_gate: Optional[nn.Module] = None,
                           expert_idx: Optional[int] = None) -> None:
        b, m = self._block(layer_idx), self.module_map
        _set(b, m.up, new_up)
        _set(b, m.down, new_down)
        if new_gate is not None and m.gate:
            _set(b, m.gate, new_gate)

    def replace_attn_layers(self, layer_idx: int, new_q: Optional[nn.Module],
                            new_k: Optional[nn.Module], new_v: Optional[nn.Module],
                            new_o: Optional[nn.Module]) -> None:
        b, m = self._block(layer_idx), self.module_map
        for path, mod in ((m.q, new_q), (m.k, new_k), (m.v, new_v), (m.o, new_o)):
            if mod is not None:
                _set(b, path, mod)

    # ---- never called by main() in the reference either (model_adapter.py:249-251,296-300) ----
    def compute_layer_energy(self, layer_idx: int, Ca: Tensor | None = None):
        raise NotImplementedError("compute_layer_energy is unused by the MoDeGPT flow")

    def calibrate_model(self, n_samples: int, batch_size: int, target_layers: list[int],
                        dataset="wikitext"):
        from ..calibration import load_calibs

        return load_calibs(self, n_samples, batch_size, dataset=dataset, target_layers=target_layers)

    # ---- statistics hooks ---------------------------------------------------------------------
    def register_hooks(self, layer_idx: int, block: nn.Module, cov_mlp_list: List[Tensor],
                       cov_q_list: List[Tensor], cov_k_list: List[Tensor], cov_x_list: List[Tensor],
                       handles: List[Any], logger: logging.Logger | None = None) -> None:
        """Attach the four statistics hooks of one block.  Each hook accumulates into
        `cov_*_list[layer_idx]` in place (fp32, raw sums; `calibration` normalises at the end):

          down_proj / fc2 INPUT    -> C_mlp += H^T H         (LlamaAdapter.py:127-136; for OPT the
                                      input of fc2 is relu(fc1(x)), model_adapter.py:546-554)
          attention-input norm OUT -> C_x += X^T X            (LlamaAdapter.py:138-147)
          q_proj / k_proj OUTPUT   -> per-head C_q, C_k       (LlamaAdapter.py:115-125; raw
                                      projections: pre-RoPE, pre-q/k-norm, bias included)
        """
        m = self.module_map

        def mlp_pre_hook(_mod, inputs):
            ops.syrk_(cov_mlp_list[layer_idx], inputs[0].detach())

        def x_hook(_mod, _inp, out):
            ops.syrk_(cov_x_list[layer_idx], out.detach())

        def head_hook(cov_list):
            def hook(_mod, _inp, out):
                ops.syrk_heads_(cov_list[layer_idx], out.detach())
            return hook

        handles.append(_get(block, m.down).register_forward_pre_hook(mlp_pre_hook))
        handles.append(_get(block, m.attn_in_norm).register_forward_hook(x_hook))
        handles.append(_get(block, m.k).register_forward_hook(head_hook(cov_k_list)))
        handles.append(_get(block, m.q).register_forward_hook(head_hook(cov_q_list)))

    def register_bi_hooks(self, bi_acc: Tensor, handles: List[Any]) -> None:
        """Block-Influence partial sums: bi_acc[l] += sum over tokens of 1 - cos(h_l, h_{l+1}).
        h_l is the input of block l; h_{l+1} its output, except for the last block where the
        reference compares with `hidden_states[L]`, which HF returns AFTER the final norm
        (src/calibration.py:118-124) — reproduced by reading the final norm's output."""
        blocks = self.get_transformer_blocks()
        n = len(blocks)
        state: dict[str, Tensor] = {}

        def pre(idx):
            def hook(_mod, args, kwargs):
                state[f"in{idx}"] = (args[0] if args else kwargs["hidden_states"]).detach()
            return hook

        def post(idx):
            def hook(_mod, _args, _kwargs, out):
                y = out[0] if isinstance(out, (tuple, list)) else out
                ops.bi_cosine_(bi_acc[idx:idx + 1], state.pop(f"in{idx}"), y.detach())
            return hook

        for i, blk in enumerate(blocks):
            handles.append(blk.register_forward_pre_hook(pre(i), with_kwargs=True))
            if i < n - 1:
                handles.append(blk.register_forward_hook(post(i), with_kwargs=True))
        final_norm = _get(self.model, self.module_map.final_norm)

        def last(_mod, _inp, out):
            if f"in{n - 1}" in state:
                ops.bi_cosine_(bi_acc[n - 1:n], state.pop(f"in{n - 1}"), out.detach())

        handles.append(final_norm.register_forward_hook(last))

    # ---- stage hand-off (model_adapter.py:184-237) --------------------------------------------
    def save_layer(self, output_dir: str, suffix: str, weights: dict[str, Tensor], layer_idx) -> None:
        """`layer_{i}_{suffix}` hand-off (model_adapter.py:184-191).  Same file, same dict; by default
        it is written by the asynchronous `handoff.LayerWriter` (files are complete after
        `flush_saves()`), and while device memory is plentiful the tensors also stay resident so
        `convert_model` on this rank does not read them back.  `--sync_save` restores the reference's
        blocking `torch.save`; `--keep_layers_in_memory` skips the files altogether."""
        if self.config.keep_layers_in_memory:
            self._layer_store[(int(layer_idx), suffix)] = weights
            return
        output_dir = os.path.expandvars(output_dir)
        path = os.path.join(output_dir, f"layer_{layer_idx}_{suffix}")
        if self.config.sync_save:
            os.makedirs(output_dir, exist_ok=True)
            torch.save(weights, path)
            return
        self.prepare_writer()
        self._writer.submit(path, weights)
        first = next(iter(weights.values()))
        if first.is_cuda:
            # cudaMemGetInfo takes the driver's context lock — behind the writer threads' copies it
            # cost ~10 ms per call (96 calls on a 7B run): ask once every 16 saves
            self._cache_probe = getattr(self, "_cache_probe", 0)
            if self._cache_probe % 16 == 0:
                free, total = torch.cuda.mem_get_info(first.device)
                self._cache_ok = free > 0.3 * total
            self._cache_probe += 1
            if self._cache_ok:
                self._layer_cache[(int(layer_idx), suffix)] = weights

    def prepare_writer(self) -> None:
        """Create the asynchronous layer writer (handoff.LayerWriter) on first use."""
        if self._writer is not None or self.config.keep_layers_in_memory or self.config.sync_save:
            return
        from ..handoff import LayerWriter

        # one pool buffer holds the largest tensor a layer can produce (an uncompressed MLP matrix)
        largest = 2 * self.d_model * max(self.get_n_inner(), self.n_heads * self.head_dim)
        self._writer = LayerWriter(device=next(self.model.parameters()).device,
                                   pool_buffer_bytes=(largest + 4095) // 4096 * 4096,
                                   n_threads=int(os.environ.get("MG_WRITER_SAVERS", "8")),
                                   n_stagers=int(os.environ.get("MG_WRITER_STAGERS", "3")))

    def flush_saves(self) -> None:
        """Block until every submitted layer file is on disk (raises if a write failed)."""
        if self._writer is not None:
            self._writer.flush()

    def load_layer(self, saved_layers_dir: str, suffix: str, layer_idx: int, device) -> dict:
        if (layer_idx, suffix) in self._layer_store:
            return self._layer_store[(layer_idx, suffix)]
        if (layer_idx, suffix) in self._layer_cache:
            return self._layer_cache.pop((layer_idx, suffix))
        self.flush_saves()
        path = os.path.join(os.path.expandvars(saved_layers_dir), f"layer_{layer_idx}_{suffix}")
        return torch.load(path, map_location=device)

    @torch.no_grad()
    def convert_model(self, saved_layers_dir: str = "./compressed_output/layers/",
                      suffixes=("mlp", "qk", "vo")) -> None:
        device = next(self.model.parameters()).device

        def linear(weight: Tensor, bias: Tensor | None = None) -> nn.Linear:
            # built on the meta device and given its storage directly: no random init, no extra copy
            # when the hand-off tensor is already a contiguous bf16 tensor on the device
            lin = nn.Linear(weight.shape[1], weight.shape[0], bias=bias is not None, device="meta",
                            dtype=torch.bfloat16)
            lin.weight = nn.Parameter(weight.detach().to(device=device, dtype=torch.bfloat16).contiguous())
            if bias is not None:
                lin.bias = nn.Parameter(bias.detach().to(device=device, dtype=torch.bfloat16).contiguous())
            return lin

        for suffix in suffixes:
            for i in range(self.n_layers):
                w = self.load_layer(saved_layers_dir, suffix, i, device)
                if suffix == "mlp":
                    self.replace_mlp_layers(
                        i, new_up=linear(w["up"], w.get("up_bias")),
                        new_down=linear(w["down"], w.get("down_bias")),
                        new_gate=linear(w["gate"]) if "gate" in w else None)
                elif suffix == "qk":
                    self.replace_attn_layers(i, new_q=linear(w["q_proj"], w.get("q_bias")),
                                             new_k=linear(w["k_proj"], w.get("k_bias")),
                                             new_v=None, new_o=None)
                elif suffix == "vo":
                    self.replace_attn_layers(i, new_q=None, new_k=None,
                                             new_v=linear(w["v_proj"], w.get("v_bias")),
                                             new_o=linear(w["o_proj"], w.get("o_bias")))

    # ---- config patch (LlamaAdapter.py:250-302) -----------------------------------------------
    rebuild_module = "LlamaRebuild"
    rebuild_class = "LlamaForCausalLM"

    def patch_config(self):
        import copy

        cfg = self.model.config
        original = copy.deepcopy(cfg)
        ranks = {"q_ranks": [], "k_ranks": [], "v_ranks": [], "o_ranks": [], "gate_ranks": []}
        for i in range(self.n_layers):
            qk, vo, mlp = self.get_qk_tensors(i), self.get_vo_tensors(i), self.get_mlp_tensors(i)
            ranks["q_ranks"].append(qk.query_proj.shape[0])
            ranks["k_ranks"].append(qk.key_proj.shape[0])
            ranks["v_ranks"].append(vo.v_proj.shape[0])
            ranks["o_ranks"].append(vo.o_proj.shape[1])
            ranks["gate_ranks"].append(mlp.up_proj.shape[0])
        for k, v in ranks.items():
            setattr(cfg, k, v)
        cfg.auto_map = {"AutoModelForCausalLM": f"{self.rebuild_module}.{self.rebuild_class}"}
        return original

    # ---- metrics (model_adapter.py:137-182) ---------------------------------------------------
    def save_metrics(self, path: str = "./metrics/metrics.json", jsons_path: str = "./metrics/jsons/"):
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        os.makedirs(jsons_path, exist_ok=True)
        with open(path, "w") as f:
            json.dump(ModelAdapter._metrics, f, indent=4, default=str)
        note = (self.metrics.get("note") or "")[:15]
        with open(os.path.join(jsons_path, f"{self.metrics['RunName']}--{note}.json"), "w") as f:
            json.dump(self.metrics, f, indent=4, default=str)
