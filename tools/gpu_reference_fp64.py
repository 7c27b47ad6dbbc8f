"""the UNMODIFIED reference's hot path on the same B200, in fp64 through
cuBLAS / cuSOLVER, for context next to our numbers.

The reference is installed (not copied into the repo) by the base contract's one offline install:
    cp -r /root/reference /tmp/refcopy
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref /tmp/refcopy
`baseline/_ref` is git-ignored and travels to the GPU box with the snapshot.  Its modules import
each other as `src.*`; pip flattened the package, so an alias module maps `src` onto the install
directory.  Nothing of ours is on the timed path: reference adapter, hooks, load_calibs and
compress_* functions on a 2-layer random-init model of Llama-2-7B widths.

    python tools/gpu_reference_fp64.py [out.json]
"""
import json
import os
import sys
import tempfile
import time
import types
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "baseline" / "_ref"


def main():
    out_path = str(Path(sys.argv[1]).resolve()) if len(sys.argv) > 1 else None   # before the chdir below
    res = {"what": "unmodified reference (fp64, cuBLAS/cuSOLVER) on this GPU, 2 layers of Llama-2-7B widths"}
    try:
        if not (REF / "calibration.py").exists():
            raise RuntimeError("baseline/_ref is missing (see the module docstring for the install line)")
        import torch

        alias = types.ModuleType("src")
        alias.__path__ = [str(REF)]
        sys.modules["src"] = alias
        os.chdir(tempfile.mkdtemp(prefix="mg_ref_cwd_"))
        import src.adapters.CompressionConfig as cc
        import src.adapters.model_adapter as ma
        import src.calibration as cal
        import src.compression.compress_mlp as mlp
        import src.compression.compress_qk as qk
        import src.compression.compress_vo as vo
        import src.compression_utils as cu
        from transformers import AutoModelForCausalLM, LlamaConfig

        L, B, T = 2, 4, 2048
        torch.manual_seed(0)
        cfg = LlamaConfig(hidden_size=4096, intermediate_size=11008, num_hidden_layers=L,
                          num_attention_heads=32, num_key_value_heads=32, head_dim=128, vocab_size=32000,
                          max_position_embeddings=2048, tie_word_embeddings=False)
        model = AutoModelForCausalLM.from_config(cfg).to(torch.bfloat16).to("cuda:0").eval()
        adapter = ma.ModelAdapter.from_model(model, tokenizer=None)
        tmp = tempfile.mkdtemp(prefix="mg_ref_layers_")
        adapter.config = cc.CompressionConfig(
            model="synthetic", temp_storage_dir=tmp, order="mlp,qk,vo", calib_size=2 * B, calibs_batch_size=B,
            compression_ratio=0.25, max_sparsity=0.95, sparsity_smoothing=0.04948, ridge_vo=1e-5,
            ridge_qk=1e-2, nystrom_ridge=1e-4)
        g = torch.Generator().manual_seed(1234)
        tokens = torch.randint(0, 32000, (2 * B, T), generator=g).to("cuda:0")
        layers = list(range(L))

        def timed(fn):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            return time.perf_counter() - t0, r

        adapter.calibs = [tokens[:B]]
        timed(lambda: cal.load_calibs(adapter, B, B, dataset="synthetic", target_layers=layers))   # warm-up
        adapter.calibs = [tokens[:B], tokens[B:]]
        t_cal, (cov_mlp, cov_q, cov_k, cov_x, bi) = timed(
            lambda: cal.load_calibs(adapter, 2 * B, B, dataset="synthetic", target_layers=layers))
        n_tok = 2 * B * T
        res["calibration"] = {
            "seconds": t_cal, "tokens": n_tok, "layers": L,
            "tokens_per_s_these_layers": n_tok / t_cal,
            "tokens_per_s_scaled_to_32_layers": n_tok / t_cal * L / 32,
            "note": "hooked bf16 forward + fp64 statistics of 2 layers; scaled by 2/32 for the whole model"}
        keep = [0.75] * L
        t_mlp, _ = timed(lambda: mlp.compress_nystrom(adapter, cov_mlp, keep, layers))
        t_qk, _ = timed(lambda: qk.compress_qk(adapter, (cov_q, cov_k), keep, target_layers=layers))
        t_vo, _ = timed(lambda: vo.compress_vo(adapter, cov_x, keep, target_layers=layers))
        res["compress"] = {"s_per_layer": (t_mlp + t_qk + t_vo) / L,
                           "ms_per_layer": {"mlp": 1e3 * t_mlp / L, "qk": 1e3 * t_qk / L, "vo": 1e3 * t_vo / L},
                           "note": "incl. the reference's blocking torch.save of each layer file"}
        res["gpu"] = torch.cuda.get_device_name(0)
    except BaseException as e:   # report, never fail the calling script
        import traceback

        res["unavailable"] = f"{type(e).__name__}: {e}"
        res["traceback"] = traceback.format_exc()[-2000:]
    text = json.dumps(res, indent=1)
    print(text)
    if out_path:
        Path(out_path).write_text(text)


if __name__ == "__main__":
    main()
