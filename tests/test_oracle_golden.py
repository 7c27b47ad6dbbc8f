"""The CPU oracle against vectors produced by the unmodified reference (oracle/make_golden.py).

This is what pins the oracle: every restated function must reproduce what the reference's own
function returned on the recorded inputs.
"""
import numpy as np
import pytest

from oracle import modegpt_oracle as O


def rel(a, b):
    return np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-300)


def test_sqrt_psd(golden):
    g = golden("utils")
    assert rel(O.sqrt_psd(g["sqrt_in"], 1e-4), g["sqrt_out"]) < 1e-12
    s, si = O.sqrt_psd(g["sqrt_in"], 1e-2, inverse=True)
    assert rel(s, g["sqrt2_out"]) < 1e-12
    assert rel(si, g["sqrt2_inv"]) < 1e-10


@pytest.mark.parametrize("case", range(4))
def test_allocate_global_sparsity(golden, case):
    g = golden("utils")
    ratio, smooth, cap = g[f"alloc{case}_params"]
    keep = O.allocate_global_sparsity(g[f"alloc{case}_bi"], ratio, smooth, cap)
    np.testing.assert_allclose(keep, g[f"alloc{case}_keep"], rtol=0, atol=1e-14)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_nystrom_mlp(golden, tag):
    g = golden("mlp")
    keep, ridge = g[f"{tag}_params"]
    assert rel(O.ridge_scores(g[f"{tag}_c"], ridge), g[f"{tag}_scores"]) < 1e-9
    out, idx, rank = O.nystrom_mlp(g[f"{tag}_wu"], g[f"{tag}_wg"], g[f"{tag}_wd"], g[f"{tag}_c"],
                                   keep, ridge)
    assert rank == int(g[f"{tag}_rank"])
    # gathers are exact; the solve is compared after bf16 rounding (ulp flips allowed)
    np.testing.assert_array_equal(out["up"], g[f"{tag}_up"])
    np.testing.assert_array_equal(out["gate"], g[f"{tag}_gate"])
    assert rel(out["down"], g[f"{tag}_down"]) < 1e-3
    assert np.mean(out["down"] == g[f"{tag}_down"]) > 0.99


def test_qk_heads(golden):
    g = golden("qk")
    m = O.qk_head_gqa(g["gqa_cq"], g["gqa_ck"], int(g["gqa_rank"]), float(g["gqa_ridge"]))
    np.testing.assert_array_equal(m, g["gqa_mask"])
    np.testing.assert_array_equal(g["gqa_wq"][:, m, :], g["gqa_q"])
    np.testing.assert_array_equal(g["gqa_wk"][0][m], g["gqa_k"])
    m = O.qk_head_mha(g["mha_cq"], g["mha_ck"], int(g["mha_rank"]))
    np.testing.assert_array_equal(m, g["mha_mask"])
    np.testing.assert_array_equal(g["gqa_wq"][0][m], g["mha_q"])
    m = O.qk_head_opt(g["mha_cq"], g["mha_ck"], int(g["opt_rank"]))
    np.testing.assert_array_equal(g["opt_wq"][m], g["opt_q"])
    np.testing.assert_array_equal(g["opt_bq"][m], g["opt_bq_out"])
    np.testing.assert_array_equal(g["opt_bk"][m], g["opt_bk_out"])


def _align(v_ours, o_ours, v_ref):
    """singular vectors are defined up to sign: align each component to the reference."""
    sgn = np.sign(np.sum(v_ours * v_ref, axis=1))
    sgn[sgn == 0] = 1.0
    return v_ours * sgn[:, None], o_ours * sgn[None, :]


def test_vo_heads(golden):
    g = golden("vo")
    hd, rank = int(g["hd"]), int(g["rank"])
    root, root_inv = O.vo_roots(g["c"], float(g["ridge"]))
    for h in range(2):
        v, o = O.vo_head_mha(g["mha_wv"][h * hd:(h + 1) * hd], g["mha_wo"][:, h * hd:(h + 1) * hd],
                             root, root_inv, rank)
        vr, orf = g["mha_v"][h * rank:(h + 1) * rank], g["mha_o"][:, h * rank:(h + 1) * rank]
        v, o = _align(v, o, vr)
        assert rel(v, vr) < 1e-8 and rel(o, orf) < 1e-8
    v, os_ = O.vo_head_gqa(g["gqa_wv"], [g["mha_wo"][:, :hd], g["mha_wo"][:, hd:]], root, root_inv,
                           rank)
    o = np.concatenate(os_, 1)
    sgn = np.sign(np.sum(v * g["gqa_v"], axis=1))
    assert rel(v * sgn[:, None], g["gqa_v"]) < 1e-8
    assert rel(o * np.tile(sgn, 2)[None, :], g["gqa_o"]) < 1e-8


def test_vo_heads_ill_conditioned(golden):
    """channel spread 1.5: cond(C) ~ 7e6, sigma_1 / sigma_r ~ 27 — the fixture that separates an
    fp32 Gram-squaring type-III from a factor-based one (tests/test_gpu_kernels.py)."""
    g = golden("vo_illcond")
    assert float(g["cond_c"]) > 1e6
    hd, rank, heads = int(g["hd"]), int(g["rank"]), int(g["heads"])
    c = g["c"].astype(np.float64)
    _, v64, o64 = O.vo_layer(g["wv"], g["wo"], c, heads, heads, hd, rank, float(g["ridge"]))
    sgn = np.sign(np.sum(v64 * g["mha_v"], axis=1))
    assert rel(v64 * sgn[:, None], g["mha_v"]) < 1e-7
    assert rel(o64 * sgn[None, :], g["mha_o"]) < 1e-7
    _, v64, o64 = O.vo_layer(g["wv"][:2 * hd], g["wo"], c, heads, 2, hd, rank, float(g["ridge"]))
    sgn = np.sign(np.sum(v64 * g["gqa_v"], axis=1))
    assert rel(v64 * sgn[:, None], g["gqa_v"]) < 1e-7
    sgn_o = np.repeat(sgn.reshape(2, rank), 2, axis=0).reshape(-1)
    assert rel(o64 * sgn_o[None, :], g["gqa_o"]) < 1e-7


@pytest.mark.parametrize("tag", ["llama_mha", "llama_gqa", "qwen3_gqa"])
def test_pipeline(golden, tag):
    """Layer-level functions and the statistics, on the recorded hook inputs of a tiny model."""
    g = golden(f"pipeline_{tag}")
    d, d_int, L, H, KV, hd, _, qwen = (int(x) for x in g["cfg"])
    arch = "qwen3" if qwen else "llama"
    n_texts = g["tokens"].shape[0]
    bi = np.zeros(L)
    for b in range(2):
        hs = g[f"hidden{b}"]
        for l in range(L):
            bi[l] += O.bi_batch(hs[l], hs[l + 1])
    bi /= n_texts
    np.testing.assert_allclose(bi, g["bi"], rtol=1e-10)
    keep = O.allocate_global_sparsity(bi, 0.3, 0.04948, 0.95)
    np.testing.assert_allclose(keep, g["keep"], rtol=0, atol=1e-12)
    for l in range(L):
        c_mlp = O.normalise_stats(O.gram_rows(g[f"mlp_in{l}"]), n_texts)
        c_x = O.normalise_stats(O.gram_rows(g[f"ln_out{l}"]), n_texts)
        c_q = O.normalise_stats(O.gram_heads(g[f"q_out{l}"], H, hd), n_texts)
        c_k = O.normalise_stats(O.gram_heads(g[f"k_out{l}"], KV, hd), n_texts)
        assert rel(c_mlp, g[f"cov_mlp{l}"]) < 1e-12
        assert rel(c_x, g[f"cov_x{l}"]) < 1e-12
        assert rel(c_q, g[f"cov_q{l}"]) < 1e-12
        assert rel(c_k, g[f"cov_k{l}"]) < 1e-12
        pre = f"w:model.layers.{l}."
        out, idx, rank = O.nystrom_mlp(g[pre + "mlp.up_proj.weight"], g[pre + "mlp.gate_proj.weight"],
                                       g[pre + "mlp.down_proj.weight"], g[f"cov_mlp{l}"],
                                       g["keep"][l], 1e-4)
        np.testing.assert_array_equal(out["up"], g[f"L{l}_mlp_up"])
        np.testing.assert_array_equal(out["gate"], g[f"L{l}_mlp_gate"])
        assert rel(out["down"], g[f"L{l}_mlp_down"]) < 2e-3
        r = O.head_rank(hd, g["keep"][l], rope=True)
        qk, mask = O.qk_layer(g[pre + "self_attn.q_proj.weight"], g[pre + "self_attn.k_proj.weight"],
                              g[f"cov_q{l}"], g[f"cov_k{l}"], H, KV, hd, r, arch, 1e-2)
        np.testing.assert_array_equal(mask, g[f"mask{l}"])
        np.testing.assert_array_equal(qk["q_proj"], g[f"L{l}_qk_q_proj"])
        np.testing.assert_array_equal(qk["k_proj"], g[f"L{l}_qk_k_proj"])
        rv = O.head_rank(hd, g["keep"][l], rope=True, clamp_to_head=False)
        vo, v64, o64 = O.vo_layer(g[pre + "self_attn.v_proj.weight"], g[pre + "self_attn.o_proj.weight"],
                                  g[f"cov_x{l}"], H, KV, hd, rv, 1e-5)
        assert vo["v_proj"].shape == g[f"L{l}_vo_v_proj"].shape
        assert vo["o_proj"].shape == g[f"L{l}_vo_o_proj"].shape
        # sign-free comparison: per head the product O'V' (d x d, rank r) is unique
        grp = H // KV
        for h in range(KV):
            vr = g[f"L{l}_vo_v_proj"][h * rv:(h + 1) * rv]
            vo_h = v64[h * rv:(h + 1) * rv]
            for j in range(grp):
                q = h * grp + j
                orf = g[f"L{l}_vo_o_proj"][:, q * rv:(q + 1) * rv]
                prod_ref = orf.astype(np.float64) @ vr.astype(np.float64)
                prod = o64[:, q * rv:(q + 1) * rv] @ vo_h
                assert rel(prod, prod_ref) < 2e-2   # reference side is bf16-rounded


def test_pipeline_opt(golden):
    """BASELINE config #1 in miniature: the oracle-level OPT pipeline (oracle/opt_pipeline.py) on
    the recorded tiny OPT reproduces the fixture (which make_golden.py cross-checked, head by head,
    against the reference's surviving compress_head_opt / compress_head / get_ridge_scores)."""
    import torch
    from transformers import AutoModelForCausalLM, OPTConfig

    from oracle import opt_pipeline

    g = golden("pipeline_opt")
    d, ffn, L, H, _, hd, vocab = (int(x) for x in g["cfg"])
    cfg = OPTConfig(hidden_size=d, num_attention_heads=H, ffn_dim=ffn, num_hidden_layers=L, vocab_size=vocab,
                    max_position_embeddings=256, word_embed_proj_dim=d, do_layer_norm_before=True)
    model = AutoModelForCausalLM.from_config(cfg)
    model.load_state_dict({k[2:]: torch.tensor(v) for k, v in g.items() if k.startswith("w:")})
    model = model.to(torch.bfloat16).eval()
    tokens = torch.tensor(g["tokens"])
    ratio, ridge, ridge_vo, smooth, cap = (float(x) for x in g["hyper"])
    res = opt_pipeline.run(model, [tokens[0:2], tokens[2:4]], compression_ratio=ratio, nystrom_ridge=ridge,
                           ridge_vo=ridge_vo, smoothing=smooth, max_sparsity=cap)
    np.testing.assert_allclose(res["bi"], g["bi"], rtol=1e-9)
    np.testing.assert_allclose(res["keep"], g["keep"], rtol=0, atol=1e-12)
    for l in range(L):
        assert rel(res[f"cov_mlp{l}"], g[f"cov_mlp{l}"]) < 1e-12
        np.testing.assert_array_equal(res[f"L{l}_mlp_idx"], g[f"L{l}_mlp_idx"])
        np.testing.assert_array_equal(res[f"L{l}_qk_mask"], g[f"L{l}_qk_mask"])
        np.testing.assert_array_equal(res[f"L{l}_qk_q_bias"], g[f"L{l}_qk_q_bias"])
        assert (res[f"L{l}_mlp_down"] == g[f"L{l}_mlp_down"]).mean() > 0.99
        sgn = np.sign(np.sum(res[f"L{l}_vo_v64"] * g[f"L{l}_vo_v64"], axis=1))
        assert rel(res[f"L{l}_vo_v64"] * sgn[:, None], g[f"L{l}_vo_v64"]) < 1e-7
