"""BASELINE config #1 at full size: OPT-125M, 30 %, 32 x 2048 synthetic calibration tokens.

"CPU reference run (plumbing, no GPU)": the reference's own OPT adapter cannot be instantiated
(SURVEY A.3), so the CPU side is the oracle-level OPT pipeline (oracle/opt_pipeline.py — pinned
oracle functions, cross-checked against the reference's surviving OPT functions by
oracle/make_golden.py) on the host cores, fp64.  The GPU side is this build.  Same seeded
random-init weights (initialised on the CPU, then copied), same tokens.

  1. statistics + Block-Influence: GPU calibration vs the oracle's (different forward arithmetic:
     fp32 on the CPU, bf16 on the GPU — agreement to bf16 noise);
  2. decompositions of every layer on the ORACLE's statistics: kept MLP rows, Q/K masks and biases
     must be identical; W_down / V' / O' within 1e-3 of the oracle's bf16-rounded tensors.

Runs as a GPU test (`pytest -m gpu`, ~75 s of which 60 s are the CPU oracle) and as a script:

    python tests/test_gpu_config1_opt125m.py [out.json]

The report goes to $MG_REPORT_DIR/config1_opt125m.json when that variable is set
(profiles/r2_config1_opt125m.json is such a report).
"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import modegpt_oracle as O  # noqa: E402
from oracle import opt_pipeline  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def run_config1() -> dict:
    from modegpt_b200.adapters.CompressionConfig import CompressionConfig
    from modegpt_b200.adapters.model_adapter import ModelAdapter
    from modegpt_b200.calibration import load_calibs
    from modegpt_b200.compression.compress_mlp import compress_nystrom
    from modegpt_b200.compression.compress_qk import compress_qk
    from modegpt_b200.compression.compress_vo import compress_vo
    from modegpt_b200.compression_utils import allocate_global_sparsity
    from modegpt_b200.eval import synthetic_tokens
    from modegpt_b200.model_utils import build_synthetic_model
    from modegpt_b200 import ops as ops_mod

    torch.set_num_threads(max(1, torch.get_num_threads()))
    hyper = dict(compression_ratio=0.30, nystrom_ridge=1e-4, ridge_vo=1e-5, smoothing=0.04948, max_sparsity=0.95)
    n_seq, bs, T = 32, 4, 2048
    cpu_model = build_synthetic_model("opt-125m", device="cpu", seed=0)          # bf16 weights
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():     # HF zero-initialises biases; give them values so the bias paths count
        for blk in cpu_model.model.decoder.layers:
            for lin in (blk.fc1, blk.fc2, blk.self_attn.q_proj, blk.self_attn.k_proj, blk.self_attn.v_proj,
                        blk.self_attn.out_proj):
                lin.bias.copy_((0.02 * torch.randn(lin.bias.shape, generator=g)).to(torch.bfloat16))
    tokens = synthetic_tokens(n_seq, T, cpu_model.config.vocab_size, 1234)
    batches = [tokens[i:i + bs] for i in range(0, n_seq, bs)]
    L, H, d = cpu_model.config.num_hidden_layers, cpu_model.config.num_attention_heads, cpu_model.config.hidden_size

    # ---- CPU: oracle-level pipeline (fp32 forward of the SAME bf16 weights, fp64 everything else)
    t0 = time.time()
    ref_model = build_synthetic_model("opt-125m", device="cpu", seed=0)
    ref_model.load_state_dict(cpu_model.state_dict())
    ref = opt_pipeline.run(ref_model.float(), batches, **hyper)
    cpu_s = time.time() - t0
    res = {"what": "BASELINE config #1: OPT-125M 30 %, 32x2048 synthetic tokens; oracle-level CPU pipeline vs GPU",
           "cpu_pipeline_seconds": round(cpu_s, 1), "cpu_threads": torch.get_num_threads()}

    # ---- GPU: this build, same weights
    dev = "cuda:0"
    model = build_synthetic_model("opt-125m", device="cpu", seed=0)
    model.load_state_dict(cpu_model.state_dict())
    model = model.to(dev)
    adapter = ModelAdapter.from_model(model, tokenizer=None)
    adapter.config = CompressionConfig(model="synthetic:opt-125m", order="mlp,qk,vo", dataset="synthetic",
                                       calib_size=n_seq, calibs_batch_size=bs, compression_ratio=0.30,
                                       nystrom_ridge=1e-4, ridge_vo=1e-5, ridge_qk=1e-2, sparsity_smoothing=0.04948,
                                       max_sparsity=0.95, keep_layers_in_memory=True)
    adapter.calibs = [b.to(dev) for b in batches]
    layers = list(range(L))
    torch.cuda.synchronize()
    t0 = time.time()
    cov_mlp, cov_q, cov_k, cov_x, bi = load_calibs(adapter, n_seq, bs, dataset="synthetic", target_layers=layers)
    torch.cuda.synchronize()
    res["gpu_calibration_seconds"] = round(time.time() - t0, 3)
    res["statistics_rel_vs_oracle"] = {
        "C_mlp": max(rel(cov_mlp[l].cpu().numpy(), ref[f"cov_mlp{l}"]) for l in layers),
        "C_x": max(rel(cov_x[l].cpu().numpy(), ref[f"cov_x{l}"]) for l in layers),
        "C_q": max(rel(cov_q[l].cpu().numpy(), ref[f"cov_q{l}"]) for l in layers),
        "C_k": max(rel(cov_k[l].cpu().numpy(), ref[f"cov_k{l}"]) for l in layers),
        "BI": float(np.max(np.abs(np.array(bi) - ref["bi"]) / np.abs(ref["bi"])))}
    keep_gpu = allocate_global_sparsity(bi, 0.30, 0.04948, 0.95, adapter=adapter)
    res["keep_ratio_max_abs_diff_from_gpu_statistics"] = float(np.max(np.abs(np.array(keep_gpu) - ref["keep"])))

    # ---- decompositions on the oracle's statistics
    f32 = lambda a: torch.tensor(np.asarray(a), device=dev, dtype=torch.float32)
    keep = allocate_global_sparsity(list(map(float, ref["bi"])), 0.30, 0.04948, 0.95, adapter=adapter)
    assert np.allclose(keep, ref["keep"], rtol=0, atol=1e-12)
    t0 = time.time()
    compress_nystrom(adapter, [f32(ref[f"cov_mlp{l}"]) for l in layers], keep, layers)
    compress_qk(adapter, ([f32(ref[f"cov_q{l}"]) for l in layers], [f32(ref[f"cov_k{l}"]) for l in layers]), keep,
                target_layers=layers)
    compress_vo(adapter, [f32(ref[f"cov_x{l}"]) for l in layers], keep, target_layers=layers)
    torch.cuda.synchronize()
    res["gpu_decomposition_seconds"] = round(time.time() - t0, 3)
    ours = adapter._layer_store
    bf = lambda t: t.float().cpu().numpy()
    per_layer = []
    for l in layers:
        mlp, qk, vo = ours[(l, "mlp")], ours[(l, "qk")], ours[(l, "vo")]
        v64, o64 = ref[f"L{l}_vo_v64"], ref[f"L{l}_vo_o64"]
        vb, ob = bf(vo["v_proj"]).astype(np.float64), bf(vo["o_proj"]).astype(np.float64)
        sgn = np.sign(np.sum(vb * v64, axis=1))
        per_layer.append({
            "layer": l, "ranks_mlp_qk_vo": [int(x) for x in ref[f"L{l}_ranks"]],
            "mlp_rows_identical": bool(np.array_equal(bf(mlp["up"]), ref[f"L{l}_mlp_up"])),
            "mlp_up_bias_identical": bool(np.array_equal(bf(mlp["up_bias"]), O.to_bf16(ref[f"L{l}_mlp_up_bias"]))),
            "mlp_down_rel": rel(bf(mlp["down"]), ref[f"L{l}_mlp_down"]),
            "mlp_down_identical": float(np.mean(bf(mlp["down"]) == ref[f"L{l}_mlp_down"])),
            "qk_q_identical": bool(np.array_equal(bf(qk["q_proj"]), ref[f"L{l}_qk_q_proj"])),
            "qk_k_identical": bool(np.array_equal(bf(qk["k_proj"]), ref[f"L{l}_qk_k_proj"])),
            "qk_bias_identical": bool(np.array_equal(bf(qk["q_bias"]), O.to_bf16(ref[f"L{l}_qk_q_bias"]))
                                      and np.array_equal(bf(qk["k_bias"]), O.to_bf16(ref[f"L{l}_qk_k_bias"]))),
            "vo_v_rel": rel(vb * sgn[:, None], O.to_bf16(v64)), "vo_o_rel": rel(ob * sgn[None, :], O.to_bf16(o64)),
            "vo_v_identical": float(np.mean(vb * sgn[:, None] == O.to_bf16(v64))),
            "vo_o_bias_max_abs_diff": float(np.max(np.abs(bf(vo["o_bias"]) - ref[f"L{l}_vo_o_bias"]))),
        })
    # kept-row mismatches: are they ties within fp32 noise of the selection threshold?
    for p in per_layer:
        l = p["layer"]
        if not p["mlp_rows_identical"]:
            scores = O.ridge_scores(ref[f"cov_mlp{l}"], 1e-4)
            rank = int(ref[f"L{l}_ranks"][0])
            order = np.sort(scores)
            thr = 0.5 * (order[rank - 1] + order[rank])
            ref_idx = set(int(i) for i in ref[f"L{l}_mlp_idx"])
            gpu_scores = ops_mod.ridge_scores(f32(ref[f"cov_mlp{l}"]), float(np.float32(1e-4))).cpu().numpy()
            gpu_idx = set(int(i) for i in ops_mod.select_k(torch.tensor(gpu_scores, device=dev), rank).cpu().numpy())
            diff = sorted(ref_idx ^ gpu_idx)
            p["mlp_rows_differing"] = len(diff)
            p["mlp_rows_max_rel_gap_to_threshold"] = float(max(abs(scores[i] - thr) / abs(thr) for i in diff))
            p["mlp_scores_rel_err"] = rel(gpu_scores, scores)
    res["layers"] = per_layer
    res["all_index_sets_identical"] = all(p["mlp_rows_identical"] and p["qk_q_identical"] and p["qk_k_identical"]
                                          and p["qk_bias_identical"] for p in per_layer)
    res["worst_rel"] = {k: max(p[k] for p in per_layer) for k in ("mlp_down_rel", "vo_v_rel", "vo_o_rel")}
    return res


try:
    import pytest

    @pytest.mark.gpu
    def test_config1_opt125m_cpu_oracle_pipeline_vs_gpu():
        import os

        res = run_config1()
        out = os.environ.get("MG_REPORT_DIR")
        if out:
            Path(out, "config1_opt125m.json").write_text(json.dumps(res, indent=1))
        st = res["statistics_rel_vs_oracle"]
        assert max(st.values()) < 3e-3, st           # bf16 GPU forward vs fp32 CPU forward
        for p in res["layers"]:
            assert p["qk_q_identical"] and p["qk_k_identical"] and p["qk_bias_identical"], p
            assert p["vo_v_rel"] < 1e-3 and p["vo_o_rel"] < 1e-3 and p["vo_v_identical"] > 0.9, p
            assert p["vo_o_bias_max_abs_diff"] < 1e-3, p
            if p["mlp_rows_identical"]:
                assert p["mlp_up_bias_identical"], p
                assert p["mlp_down_rel"] < 1e-3 and p["mlp_down_identical"] > 0.95, p
            else:
                # a different selection is admissible only between candidates whose scores sit
                # inside the fp32 error of the scores at the threshold (north_star's noise rule)
                assert p["mlp_rows_differing"] <= 4, p
                assert p["mlp_rows_max_rel_gap_to_threshold"] < 2.0 * p["mlp_scores_rel_err"] + 1e-5, p
except ImportError:      # script use without pytest
    pass


if __name__ == "__main__":
    report = json.dumps(run_config1(), indent=1)
    print(report)
    if len(sys.argv) > 1:
        Path(sys.argv[1]).write_text(report)
