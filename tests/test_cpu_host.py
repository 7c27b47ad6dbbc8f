"""CPU-only tests: C ABI surface, configuration surface, rank allocation, distributed plumbing."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    from modegpt_b200 import _lib

    header = (ROOT / "include" / "modegpt_b200.h").read_text()
    declared = set(re.findall(r"\b(mg_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(raw, name), name
    assert _lib.lib.mg_version() == 1


def test_argument_errors_are_reported_without_touching_the_gpu():
    from modegpt_b200 import _lib

    lib = _lib.lib
    assert lib.mg_syrk_bf16_f32(None, 8, 8, 8, None, 8, 1.0, 1, None) == -1
    assert lib.mg_bi_cosine_bf16(None, 8, None, 8, 4, 8, None, None) == -1
    assert lib.mg_select_k_f32(ctypes.c_void_p(16), 8, 9, 0, ctypes.c_void_p(16), None) == -2
    assert lib.mg_qk_select_f32(ctypes.c_void_p(16), ctypes.c_void_p(16), 4, 3, 64, 0, 1e-4, 1e-4, 8,
                                ctypes.c_void_p(16), None) == -2
    assert lib.mg_ridge_scores_ws_bytes(11008) > 3 * 11008 * 11008 * 2
    assert b"workspace" in lib.mg_error_string(-10)
    with pytest.raises(_lib.MgError):
        _lib.check("mg_x", -7)
    assert _lib.check("mg_x", 5) == 5     # numerical info is returned, not raised


def test_missing_library_fails_loudly(tmp_path):
    code = ("import modegpt_b200._lib as L, pathlib; L.LIB_PATH = pathlib.Path('/nonexistent/x.so');"
            "L._load()")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU or PyTorch fallback" in r.stderr


REFERENCE_FIELDS = {   # src/adapters/CompressionConfig.py:8-35
    "model": "facebook/opt-6.7b", "device": 0, "factorize_src_model": "", "nystrom_src_model": "",
    "tokenizer_src": "mistralai/Mixtral-8x7B-v0.1", "output_dir": "compressed_output",
    "temp_storage_dir": "./compressed_output/layers/", "dataset": "wikitext", "nystrom_ridge": 1e-2,
    "order": None, "calib_size": 32, "calibs_batch_size": 4, "compression_ratio": 0.5, "note": "NA",
    "max_sparsity": 0.8, "sparsity_smoothing": 0.15, "ridge_vo": 1e-4, "ridge_qk": 1e-6, "debug": False,
}


def test_compression_config_keeps_the_reference_surface():
    from modegpt_b200.adapters.CompressionConfig import CompressionConfig

    c = CompressionConfig()
    for k, v in REFERENCE_FIELDS.items():
        assert getattr(c, k) == v, k
    # the tests.sh recipe parses (tests.sh:116-133)
    c = CompressionConfig.from_args(
        "--model Qwen/Qwen3-8B --device 0 --compression_ratio 0.30 --calib_size 128 "
        "--calibs_batch_size 16 --output_dir out --note x --order mlp,qk,vo --max_sparsity 0.95 "
        "--ridge_vo 1e-5 --ridge_qk 1e-2 --sparsity_smoothing 0.04948 --nystrom_ridge 1e-4 "
        "--dataset alpaca".split())
    assert c.order == "mlp,qk,vo" and c.ridge_qk == 1e-2 and c["calib_size"] == 128
    assert "order" in c and c.get("missing", 3) == 3 and c.to_dict()["dataset"] == "alpaca"


@pytest.mark.parametrize("case", range(4))
def test_allocate_global_sparsity_matches_reference(golden, case):
    from modegpt_b200.compression_utils import allocate_global_sparsity

    g = golden("utils")
    ratio, smooth, cap = g[f"alloc{case}_params"]
    keep = allocate_global_sparsity(list(map(float, g[f"alloc{case}_bi"])), ratio, smooth, cap)
    np.testing.assert_allclose(keep, g[f"alloc{case}_keep"], rtol=0, atol=1e-15)


def test_ranks_of_the_golden_pipelines(golden):
    """keep ratio -> per-layer ranks exactly as the reference's saved tensors have them."""
    from modegpt_b200.compression_utils import head_rank

    for tag in ("llama_mha", "llama_gqa", "qwen3_gqa"):
        g = golden(f"pipeline_{tag}")
        d, d_int, L, H, KV, hd, _, _ = (int(x) for x in g["cfg"])
        for l in range(L):
            keep = g["keep"][l]
            assert g[f"L{l}_mlp_up"].shape[0] == int(d_int * keep)
            r = head_rank(hd, keep, rope=True)
            assert g[f"L{l}_qk_q_proj"].shape[0] == H * r and g[f"mask{l}"].shape == (KV, r)
            assert g[f"L{l}_vo_v_proj"].shape[0] == KV * head_rank(hd, keep, True, clamp_to_head=False)


def test_sqrt_m_compat(golden):
    from modegpt_b200.compression_utils import sqrt_M

    g = golden("utils")
    out = sqrt_M(torch.tensor(g["sqrt_in"]), ridge_lambda=1e-4)
    np.testing.assert_allclose(out.numpy(), g["sqrt_out"], rtol=1e-10, atol=1e-12)


def _gloo_worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from modegpt_b200 import distributed as D

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batches = list(range(7))
    mine = D.shard_batches(batches)
    assert mine == [b for b in batches if b % world == rank]
    # per-layer sums reduced to the owner only
    n_layers = 5
    for layer in range(n_layers):
        t = torch.full((4, 4), float(rank + 1))
        owner = D.reduce_to_owner(t, layer)
        assert owner == (layer % world == rank)
        if owner:
            assert torch.all(t == sum(range(1, world + 1)))
    # one layer's four statistics through the reducer (CPU tensors: plain reductions; on CUDA the
    # same call packs the upper triangles into one buffer — tests/test_gpu_kernels.py, dist_check)
    red = D.LayerReducer()
    for layer in range(n_layers):
        ts = [torch.full(shape, float(rank + 1 + layer)) for shape in ((6, 6), (4, 4), (2, 3, 3), (1, 3, 3))]
        mine = red.submit(layer, *ts)
        red.wait()
        assert mine == (layer % world == rank)
        if mine:
            for t in ts:
                assert torch.all(t == sum(r + 1 + layer for r in range(world)))
    assert D.owned_layers(range(n_layers)) == [l for l in range(n_layers) if l % world == rank]
    full = D.gather_by_layer({l: torch.tensor([l, rank]) for l in D.owned_layers(range(n_layers))}, n_layers)
    assert [int(x[0]) for x in full] == list(range(n_layers))
    assert [int(x[1]) for x in full] == [l % world for l in range(n_layers)]
    # rotary masks of ragged rank (r differs per layer) through the padded all-reduce
    local = {l: torch.arange(2 * (3 + l)).reshape(2, 3 + l) + 100 * l for l in D.owned_layers(range(n_layers))}
    masks = D.gather_masks(local, n_layers + 1, rows=2, width=8, device="cpu")
    assert masks[n_layers] is None
    for l in range(n_layers):
        assert torch.equal(masks[l], torch.arange(2 * (3 + l)).reshape(2, 3 + l) + 100 * l)
    s = D.all_reduce_sum_(torch.tensor([1.0 + rank]))
    assert s.item() == sum(1.0 + r for r in range(world))
    # token-sharded perplexity (eval.compute_perplexity): every rank evaluates its share of the
    # held-out sequences, the NLL sum is all-reduced, every rank returns the single-process value
    from modegpt_b200.eval import compute_perplexity
    from modegpt_b200.model_utils import build_synthetic_model

    model = build_synthetic_model("tiny-llama-gqa", device="cpu", seed=3, max_positions=64).float()
    sharded = compute_perplexity(model, None, bs=2, dataset="synthetic", n_samples=5, seq_len=48)
    D._force_single = True
    single = compute_perplexity(model, None, bs=2, dataset="synthetic", n_samples=5, seq_len=48)
    D._force_single = False
    assert abs(sharded - single) / single < 1e-6, (sharded, single)
    D.barrier()
    Path(out_dir, f"ok{rank}").write_text("ok")
    dist.destroy_process_group()


def test_distributed_plumbing_world2_gloo(tmp_path):
    import torch.multiprocessing as mp

    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_adapter_tables_resolve_on_tiny_models():
    from modegpt_b200.adapters.model_adapter import ModelAdapter
    from modegpt_b200.model_utils import build_synthetic_model

    for preset, arch, has_gate in (("tiny-llama-gqa", "llama", True), ("tiny-qwen3", "qwen3", True),
                                   ("tiny-qwen2", "qwen2", True), ("tiny-opt", "opt", False)):
        model = build_synthetic_model(preset, device="cpu")
        a = ModelAdapter.from_model(model, tokenizer=None)
        assert a.arch == arch and a.n_layers == 3 and a.head_dim == 64 and a.d_model == 256
        mlp = a.get_mlp_components(1)
        assert (mlp.gate_proj is not None) == has_gate and a.get_n_inner() == 512
        assert a.get_qk_tensors(0).query_proj.shape == (256, 256)
        assert a.get_vo_tensors(2).o_proj.shape[0] == 256
        assert a.n_kv_heads == (2 if "gqa" in preset or "qwen" in preset else 4)


def test_layer_writer_files_match_blocking_save(tmp_path):
    """handoff.LayerWriter writes the same `torch.load`-able dicts the reference's blocking
    torch.save produces (file names and keys unchanged), including non-contiguous views, and
    surfaces writer errors in flush()."""
    import torch

    from modegpt_b200.handoff import LayerWriter

    g = torch.Generator().manual_seed(3)
    w = LayerWriter(n_threads=2, max_in_flight=2)
    expect = {}
    for i in range(6):
        d = {"up": torch.randn(8 + i, 16, generator=g).bfloat16(),
             "down": torch.randn(16, 8 + i, generator=g).bfloat16().T,     # transposed view
             "mask": torch.arange(5 + i)}
        path = tmp_path / "layers" / f"layer_{i}_mlp"
        w.submit(str(path), d)
        expect[path] = d
    w.flush()
    for path, d in expect.items():
        got = torch.load(path)
        assert set(got) == set(d)
        for k in d:
            assert got[k].dtype == d[k].dtype and torch.equal(got[k], d[k])
    bad = tmp_path / "file_not_dir"
    bad.write_text("x")
    w.submit(str(bad / "layer_0_mlp"), {"up": torch.zeros(2)})
    with pytest.raises(RuntimeError):
        w.flush()
    w.close()


def test_fp16_checkpoint_is_cast_to_bf16_at_load(tmp_path):
    """facebook/opt-* and Llama-2-*-hf ship fp16 weights; the kernels take bf16 only.  The loader
    casts once instead of failing at the first calibration hook."""
    import torch
    from transformers import AutoModelForCausalLM, LlamaConfig

    from modegpt_b200.adapters.model_adapter import ModelAdapter
    from modegpt_b200.model_utils import reload_compressed_model

    cfg = LlamaConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4,
                      num_key_value_heads=2, head_dim=16, vocab_size=96, max_position_embeddings=64,
                      tie_word_embeddings=False)
    model = AutoModelForCausalLM.from_config(cfg).to(torch.float16)
    model.save_pretrained(tmp_path / "fp16")
    loaded, _ = reload_compressed_model(str(tmp_path / "fp16"), device="cpu", tokenizer_source="synthetic:none")
    assert {p.dtype for p in loaded.parameters()} == {torch.bfloat16}
    adapter = ModelAdapter.from_model(loaded, tokenizer=None)
    adapter.validate_for_kernels()
    bad = AutoModelForCausalLM.from_config(cfg).to(torch.float16)
    with pytest.raises(TypeError):
        ModelAdapter.from_model(bad, tokenizer=None).validate_for_kernels()


def test_layer_streamed_flow_refuses_eager_attention():
    """With attention_mask=None HF's eager attention is bidirectional: the streamed calibration
    must not silently compute different statistics."""
    from modegpt_b200.adapters.CompressionConfig import CompressionConfig
    from modegpt_b200.adapters.model_adapter import ModelAdapter
    from modegpt_b200.calibration import _LayerStepper
    from modegpt_b200.model_utils import build_synthetic_model

    model = build_synthetic_model("tiny-llama", device="cpu", max_positions=64)
    adapter = ModelAdapter.from_model(model, tokenizer=None)
    adapter.config = CompressionConfig(model="tiny", dataset="synthetic", calib_size=2, calibs_batch_size=2,
                                       seq_len=32)
    model.config._attn_implementation = "eager"
    with pytest.raises(NotImplementedError):
        _LayerStepper(adapter, "synthetic")
    model.config._attn_implementation = "sdpa"
    _LayerStepper(adapter, "synthetic")


def test_timeline_summary_reads_a_lane_timeline(tmp_path):
    """tools/timeline_summary.py on a hand-made MG_PROFILE_TIMELINE file: lane totals and the
    hindsight path (the launch that ends last, then whatever ended closest before each start)."""
    root = Path(__file__).resolve().parents[1]
    src = tmp_path / "tl.csv"
    src.write_text("#rep 0\ncall,old,0x1,0.0,9.0\n#rep 1\n"
                   "call,potrf,0x1,0.0,1.0\ncall,bulk,0x2,0.5,2.0\ncall,trsm,0x1,2.1,3.0\n")
    out = tmp_path / "out.txt"
    subprocess.run([sys.executable, str(root / "tools" / "timeline_summary.py"), str(src), str(out)], check=True)
    text = out.read_text()
    assert "3 launches, 3.000 ms" in text            # only the last repetition is read
    assert "lane 0:    2 launches, busy   1.900 ms" in text
    assert "hindsight path: 2 launches, 2.400 ms in kernels + 0.100 ms of gaps" in text
