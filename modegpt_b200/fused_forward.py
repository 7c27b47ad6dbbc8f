"""Fused elementwise kernels for the calibration forward.

`fused_elementwise(model)` is a context manager that, for Llama / Qwen2 / Qwen3 style models in
bf16 on CUDA, swaps three eager elementwise code paths of the HF modeling module for one-kernel
equivalents from libmodegpt_b200 while the calibration forward runs:

    *RMSNorm.forward            7 kernels -> mg_rmsnorm_bf16
    *MLP.forward                silu + mul -> mg_swiglu_bf16   (gate/up/down projections untouched,
                                                                so the down_proj pre-hook still fires)
    apply_rotary_pos_emb        8 kernels per tensor -> mg_rope_bf16

Every intermediate bf16 rounding of the HF code is reproduced (SwiGLU and RoPE are bit-exact,
RMSNorm differs only in the summation order of the variance), so the activations the statistics
hooks observe are the HF activations.  Anything the kernels do not cover (other dtypes, CPU,
non-contiguous layouts, activations other than SiLU, OPT's LayerNorm) runs the original code.
"""
from __future__ import annotations

import contextlib
import sys

import torch

from . import ops


def _is_fast(t: torch.Tensor) -> bool:
    return t.is_cuda and t.dtype == torch.bfloat16


@contextlib.contextmanager
def fused_elementwise(model):
    module = sys.modules.get(type(model).__module__)
    cfg = model.config
    saved: list[tuple[object, str, object]] = []

    def patch(obj, name, new):
        saved.append((obj, name, getattr(obj, name)))
        setattr(obj, name, new)

    try:
        if module is not None and getattr(cfg, "model_type", "") in ("llama", "qwen2", "qwen3"):
            for name in dir(module):
                cls = getattr(module, name)
                if not isinstance(cls, type):
                    continue
                if name.endswith("RMSNorm") and hasattr(cls, "forward"):
                    orig = cls.forward

                    def norm_forward(self, hidden_states, _orig=orig):
                        if (_is_fast(hidden_states) and self.weight.dtype == torch.bfloat16
                                and hidden_states.shape[-1] % 8 == 0):
                            return ops.rmsnorm(hidden_states, self.weight, float(self.variance_epsilon))
                        return _orig(self, hidden_states)

                    patch(cls, "forward", norm_forward)
                elif name.endswith("MLP") and getattr(cfg, "hidden_act", None) == "silu":
                    orig = cls.forward

                    def mlp_forward(self, x, _orig=orig):
                        if not _is_fast(x):
                            return _orig(self, x)
                        gate, up = self.gate_proj(x), self.up_proj(x)
                        if not (gate.is_contiguous() and up.is_contiguous() and gate.numel() % 8 == 0):
                            return self.down_proj(self.act_fn(gate) * up)
                        return self.down_proj(ops.swiglu(gate, up))

                    patch(cls, "forward", mlp_forward)
            if hasattr(module, "apply_rotary_pos_emb"):
                orig_rope = module.apply_rotary_pos_emb

                def rope(q, k, cos, sin, unsqueeze_dim=1, _orig=orig_rope):
                    # HF passes [B, H, T, hd] VIEWS of contiguous [B, T, H, hd] projections
                    ok = (unsqueeze_dim == 1 and _is_fast(q) and _is_fast(k) and cos.dtype == q.dtype
                          and q.dim() == 4 and q.shape[-1] % 16 == 0 and cos.dim() == 3
                          and cos.is_contiguous() and sin.is_contiguous()
                          and q.transpose(1, 2).is_contiguous() and k.transpose(1, 2).is_contiguous()
                          and cos.shape[0] in (1, q.shape[0]))
                    if not ok:
                        return _orig(q, k, cos, sin, unsqueeze_dim)
                    qo = ops.rope_bthd(q.transpose(1, 2), cos, sin).transpose(1, 2)
                    ko = ops.rope_bthd(k.transpose(1, 2), cos, sin).transpose(1, 2)
                    return qo, ko

                patch(module, "apply_rotary_pos_emb", rope)
        yield
    finally:
        for obj, name, old in reversed(saved):
            setattr(obj, name, old)


@contextlib.contextmanager
def fused_rebuilt(model):
    """Kernel-backed `masked_rope` / `_masked_rms_norm` for a model rebuilt by the MoDeGPT flow
    (patchers/*Rebuild.py, loaded through `auto_map`; those files stay self-contained torch code so
    a checkpoint loads anywhere).  The reference gathers `cos[:, :, mask]` / `sin[:, :, mask]` on
    every forward of every layer (src/patchers/LlamaRebuild.py:155-180) and runs the eager RoPE
    sequence on the result; `mg_rope_masked_bf16` does the gather inside the rotation.  Same
    arithmetic incl. the bf16 roundings (bit-exact RoPE; the masked RMS norm differs only in the
    summation order)."""
    module = sys.modules.get(type(model).__module__)
    saved: list[tuple[object, str, object]] = []
    try:
        if module is not None and hasattr(module, "masked_rope"):
            orig = module.masked_rope

            def masked_rope(q, k, cos, sin, mask, groups, _orig=orig):
                ok = (mask is not None and _is_fast(q) and _is_fast(k) and cos.dtype == q.dtype
                      and q.dim() == 4 and cos.dim() == 3 and cos.is_contiguous() and sin.is_contiguous()
                      and q.transpose(1, 2).is_contiguous() and k.transpose(1, 2).is_contiguous()
                      and cos.shape[0] in (1, q.shape[0]) and q.shape[-1] % 2 == 0
                      and mask.dtype == torch.int64 and mask.is_contiguous())
                if not ok:
                    return _orig(q, k, cos, sin, mask, groups)
                qo = ops.rope_masked_bthr(q.transpose(1, 2), cos, sin, mask, groups).transpose(1, 2)
                ko = ops.rope_masked_bthr(k.transpose(1, 2), cos, sin, mask, 1).transpose(1, 2)
                return qo, ko

            saved.append((module, "masked_rope", orig))
            module.masked_rope = masked_rope
            for name in dir(module):
                cls = getattr(module, name)
                if isinstance(cls, type) and "_masked_rms_norm" in cls.__dict__:
                    orig_norm = cls.__dict__["_masked_rms_norm"]
                    fn = orig_norm.__func__ if isinstance(orig_norm, staticmethod) else orig_norm

                    def masked_norm(x, norm, mask, _fn=fn):
                        # the Rebuild code hands the mask already repeated per query head
                        if (_is_fast(x) and x.is_contiguous() and norm.weight.dtype == torch.bfloat16
                                and x.shape[-1] <= 128 and mask.is_contiguous()
                                and mask.shape[0] == x.shape[-2]):
                            return ops.rmsnorm_masked(x, norm.weight, mask, 1, float(norm.variance_epsilon))
                        return _fn(x, norm, mask)

                    saved.append((cls, "_masked_rms_norm", orig_norm))
                    setattr(cls, "_masked_rms_norm", staticmethod(masked_norm))
        yield
    finally:
        for obj, name, old in reversed(saved):
            setattr(obj, name, old)
