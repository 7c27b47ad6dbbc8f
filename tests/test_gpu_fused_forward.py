"""The fused calibration-forward kernels must reproduce HF's eager arithmetic: SwiGLU and RoPE bit
for bit, RMSNorm up to the summation order of the variance (a bf16 ulp on a vanishing fraction of
elements).  Then the statistics of a hooked forward must not move."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from modegpt_b200 import ops as _ops

    return _ops


@pytest.mark.parametrize("rows,d", [(1000, 4096), (4096, 128), (777, 3584), (64, 8192), (300, 256), (50, 64)])
def test_rmsnorm_matches_hf(ops, rows, d):
    from transformers.models.llama.modeling_llama import LlamaRMSNorm

    g = torch.Generator(device=DEV).manual_seed(d)
    x = (torch.randn(rows, d, device=DEV, generator=g) * 3).bfloat16()
    norm = LlamaRMSNorm(d, eps=1e-5).to(DEV).bfloat16()
    norm.weight.data = (1 + 0.3 * torch.randn(d, device=DEV, generator=g)).bfloat16()
    want = norm(x)
    got = ops.rmsnorm(x, norm.weight, 1e-5)
    diff = (got.float() - want.float()).abs()
    ulp = want.float().abs() * 2 ** -7 + 1e-30
    # a last-bit difference in the variance can flip the rounding of x*rs by one bf16 ulp, which the
    # weight multiply + second rounding turns into at most two ulps — on a vanishing fraction
    assert (diff <= 2.01 * ulp).all()
    assert (got != want).float().mean().item() < 1e-3
    # 3-D input and a strided row view
    x3 = x.view(-1, 10 if rows % 10 == 0 else 1, d) if rows % 10 == 0 else x.view(1, rows, d)
    assert torch.equal(ops.rmsnorm(x3, norm.weight, 1e-5).view(rows, d), got)


def test_swiglu_is_bit_exact(ops):
    g = torch.Generator(device=DEV).manual_seed(1)
    gate = (torch.randn(4096, 11008, device=DEV, generator=g) * 2).bfloat16()
    up = torch.randn(4096, 11008, device=DEV, generator=g).bfloat16()
    want = torch.nn.functional.silu(gate) * up
    assert torch.equal(ops.swiglu(gate, up), want)


@pytest.mark.parametrize("B,T,H,hd,bc", [(2, 256, 8, 128, 1), (3, 100, 4, 64, 1), (2, 64, 2, 32, 2)])
def test_rope_is_bit_exact(ops, B, T, H, hd, bc):
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb

    g = torch.Generator(device=DEV).manual_seed(hd)
    q = torch.randn(B, T, H, hd, device=DEV, generator=g).bfloat16()
    k = torch.randn(B, T, max(H // 2, 1), hd, device=DEV, generator=g).bfloat16()
    ang = torch.rand(bc, T, hd // 2, device=DEV, generator=g) * 6.28
    cos = torch.cat([ang.cos(), ang.cos()], -1).bfloat16()
    sin = torch.cat([ang.sin(), ang.sin()], -1).bfloat16()
    wq, wk = apply_rotary_pos_emb(q.transpose(1, 2), k.transpose(1, 2), cos, sin)
    assert torch.equal(ops.rope_bthd(q, cos, sin).transpose(1, 2), wq)
    assert torch.equal(ops.rope_bthd(k, cos, sin).transpose(1, 2), wk)


@pytest.mark.parametrize("preset", ["tiny-llama-gqa", "tiny-qwen3", "tiny-qwen2"])
def test_fused_forward_leaves_model_and_statistics_unchanged(preset):
    from modegpt_b200.adapters.CompressionConfig import CompressionConfig
    from modegpt_b200.adapters.model_adapter import ModelAdapter
    from modegpt_b200.calibration import load_calibs
    from modegpt_b200.eval import synthetic_tokens
    from modegpt_b200.fused_forward import fused_elementwise
    from modegpt_b200.model_utils import build_synthetic_model

    model = build_synthetic_model(preset, device=DEV, seed=3, max_positions=512)
    # make the norms and projections non-trivial so a wrong kernel cannot hide
    g = torch.Generator(device=DEV).manual_seed(0)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.mul_((1 + 0.3 * torch.randn(p.shape, device=DEV, generator=g)).to(p.dtype))
            else:
                p.mul_(1.5)
    tokens = synthetic_tokens(4, 256, model.config.vocab_size, 5).to(DEV)
    # layer-0 activations see identical inputs in both modes: differences there are the kernels' own
    cap = {}
    blk = model.model.layers[0]
    hooks = [blk.input_layernorm.register_forward_hook(lambda m, i, o: cap.setdefault("ln", []).append(o)),
             blk.self_attn.o_proj.register_forward_pre_hook(lambda m, i: cap.setdefault("attn", []).append(i[0])),
             blk.mlp.down_proj.register_forward_pre_hook(lambda m, i: cap.setdefault("mlp", []).append(i[0]))]
    with torch.no_grad():
        eager = model(tokens, use_cache=False).logits
        with fused_elementwise(model):
            fused = model(tokens, use_cache=False).logits
        again = model(tokens, use_cache=False).logits
    for h in hooks:
        h.remove()
    assert torch.equal(again, eager)                                   # patches are removed
    assert (cap["ln"][0] != cap["ln"][1]).float().mean().item() < 1e-3          # rare 1-ulp flips only
    for key in ("attn", "mlp"):
        a, b = cap[key][0].float(), cap[key][1].float()
        assert ((a - b).norm() / a.norm()).item() < 1e-3, key
    rel = ((fused.float() - eager.float()).norm() / eager.float().norm()).item()
    assert rel < 1e-2, rel      # a deep random net amplifies single-ulp flips; bounded, not exact
    adapter = ModelAdapter.from_model(model, None)
    stats = {}
    for eager_flag in (True, False):
        adapter.config = CompressionConfig(model=preset, dataset="synthetic", seq_len=256,
                                           eager_forward=eager_flag)
        adapter.calibs = [tokens[:2], tokens[2:]]
        stats[eager_flag] = load_calibs(adapter, 4, 2, dataset="synthetic")
    for a, b in zip(stats[True][:4], stats[False][:4]):
        for x, y in zip(a, b):
            assert ((x - y).norm() / x.norm()).item() < 1e-3
    np.testing.assert_allclose(stats[True][4], stats[False][4], rtol=1e-3)
