#!/usr/bin/env python
"""bench.py — calibration tokens/s (+ compression s/layer) for Llama-2-7B, BASELINE.json config #2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = one calibration batch (calibs_batch_size 16 x 2048 synthetic tokens) through the hooked
bf16 forward of a random-init Llama-2-7B with every statistics kernel firing (C_mlp, C_x, per-head
C_q / C_k, Block-Influence) — the hot loop of src/calibration.py:114-127.  `value` times it with the
token batch already in HBM; `e2e` re-times it with the tokens in pinned host memory (H2D inside)
and the BI scores read back (D2H inside).  After the timed steps the statistics are finalised and
`--compress-layers` layers go through type I/II/III to report seconds per layer.

`--impl reference` (and the `cpu_baseline` object of the default arm) time the reference's algorithm
for the same path — the fp64 restatement in oracle/ — on the host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

PRESET = "llama-2-7b"
BATCH, SEQ = 16, 2048            # tests.sh calibs_batch_size, fixed sequence length
D, D_INT, HEADS, KV, HD, LAYERS = 4096, 11008, 32, 32, 128, 32
HYPER = dict(compression_ratio=0.25, nystrom_ridge=1e-4, ridge_vo=1e-5, ridge_qk=1e-2,
             sparsity_smoothing=0.04948, max_sparsity=0.95)


def ncu_traffic_bytes():
    """DRAM bytes (read + write) of one C_mlp SYRK launch from the committed ncu --set full capture."""
    p = ROOT / "profiles" / "r1_syrk_pair_banded_ncu_full.json"
    if not p.exists():
        return None
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    best = None
    for k in json.loads(p.read_text())["kernels"]:
        if "gemm_tn" not in k["kernel"]:
            continue
        tot = sum(float(k[m]["value"]) * mult[k[m]["unit"]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        best = max(best or 0.0, tot)      # the n = 11008 launch is the largest <256> launch captured
    return best


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        m = json.loads(p.read_text())
        return float(m["bf16_tflops_sustained"]), float(m["hbm_gbs"]), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        busy = [x for x in sm if x >= 0.5 * max(sm)]
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle (fp64 port of the reference's algorithm) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_sample(t_slice: int = 2048, seed: int = 0) -> dict:
    """Statistics of ONE real-shape layer on `t_slice` tokens, fp64, all host threads:
    H^T H (11008^2), X^T X (4096^2), per-head Q/K Grams and the BI cosine — what the reference's
    four hooks + BI loop do per layer per batch (LlamaAdapter.py:115-147, calibration.py:118-124).
    Whole-model calibration tokens/s = t_slice / (wall * LAYERS); the bf16 forward is excluded."""
    from threadpoolctl import threadpool_limits

    from oracle import modegpt_oracle as O

    cores = os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    h = rng.standard_normal((t_slice, D_INT)).astype(np.float32)
    x = rng.standard_normal((t_slice, D)).astype(np.float32)
    q = rng.standard_normal((t_slice, HEADS * HD)).astype(np.float32)
    k = rng.standard_normal((t_slice, KV * HD)).astype(np.float32)
    y = x + 0.3 * rng.standard_normal((t_slice, D)).astype(np.float32)
    # torchrun exports OMP_NUM_THREADS=1; the baseline is entitled to every host core
    with threadpool_limits(limits=cores):
        t0 = time.perf_counter()
        O.gram_rows(h)
        O.gram_rows(x)
        O.gram_heads(q, HEADS, HD)
        O.gram_heads(k, KV, HD)
        O.bi_batch(x[None], y[None])
        wall = time.perf_counter() - t0
    return {"value": t_slice / (wall * LAYERS), "unit": "tokens/s", "cores": cores, "kind": "port",
            "sample": f"fp64 oracle statistics (C_mlp, C_x, C_q, C_k, BI) of one Llama-2-7B layer on "
                      f"{t_slice} tokens in {wall:.2f} s, scaled by 1/{LAYERS} layers; model forward excluded"}


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup):
        cpu_reference_sample(512)
    t0 = time.perf_counter()
    for i in range(args.steps):
        vals.append(cpu_reference_sample(2048, seed=i))
    ms = (time.perf_counter() - t0) * 1e3 / max(args.steps, 1)
    v = float(np.mean([x["value"] for x in vals]))
    base = dict(vals[-1], value=v)
    print(json.dumps({
        "impl": "reference", "metric": "calib_tokens_per_s", "value": v, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.gpus), "cpu_baseline": base,
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(n_gpus: int) -> dict:
    return {"workload": "Llama-2-7B 25% MoDeGPT calibration, 16x2048 synthetic tokens per step "
                        "(BASELINE configs[1]); random-init weights",
            "calibs_batch_size": BATCH, "seq_len": SEQ, "layers": LAYERS,
            "parallelism": f"token-sharded x{n_gpus}, one reduce-to-owner per layer at the end",
            "l2_policy": "inputs larger than L2 (each statistics operand is 268-721 MB)",
            "forward": "HF modules, fused RMSNorm/SwiGLU/RoPE kernels (bit-compatible)", **HYPER}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args) -> None:
    import torch.distributed as dist

    from modegpt_b200 import distributed as Dm
    from modegpt_b200 import ops
    from modegpt_b200.adapters.CompressionConfig import CompressionConfig
    from modegpt_b200.adapters.model_adapter import ModelAdapter
    from modegpt_b200.fused_forward import fused_elementwise
    from modegpt_b200.model_utils import build_synthetic_model

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tf_peak, hbm_peak, peak_src = peaks()

    model = build_synthetic_model(PRESET, device=str(dev), seed=0)
    adapter = ModelAdapter.from_model(model, tokenizer=None)
    adapter.config = CompressionConfig(model=f"synthetic:{PRESET}", dataset="synthetic", order="mlp,qk,vo",
                                       calib_size=128, calibs_batch_size=BATCH, keep_layers_in_memory=True,
                                       **HYPER)
    body = model.model                      # the LM head is not on the calibration path
    blocks = adapter.get_transformer_blocks()
    f32 = dict(dtype=torch.float32, device=dev)
    cov_mlp = [torch.zeros(D_INT, D_INT, **f32) for _ in range(LAYERS)]
    cov_x = [torch.zeros(D, D, **f32) for _ in range(LAYERS)]
    cov_q = [torch.zeros(HEADS, HD, HD, **f32) for _ in range(LAYERS)]
    cov_k = [torch.zeros(KV, HD, HD, **f32) for _ in range(LAYERS)]
    handles: list = []
    for i in range(LAYERS):
        adapter.register_hooks(i, blocks[i], cov_mlp, cov_q, cov_k, cov_x, handles, None)
    bi = torch.zeros(LAYERS, dtype=torch.float64, device=dev)
    adapter.register_bi_hooks(bi, handles)

    # time the dominant kernel (the 11008-wide SYRK) live: events around every C_mlp launch
    syrk_events: list = []
    real_syrk = ops.syrk_

    def timed_syrk(C, X, alpha=1.0, accumulate=True):
        if C.shape[0] == D_INT and recording[0]:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            real_syrk(C, X, alpha, accumulate)
            e.record()
            syrk_events.append((s, e))
        else:
            real_syrk(C, X, alpha, accumulate)

    recording = [False]
    ops.syrk_ = timed_syrk
    import modegpt_b200.adapters.model_adapter as MA
    MA.ops.syrk_ = timed_syrk

    steps_total = args.warmup + args.steps
    vocab = model.config.vocab_size
    g = torch.Generator().manual_seed(1234 + rank)
    host_tokens = [torch.randint(0, vocab, (BATCH, SEQ), generator=g).pin_memory() for _ in range(steps_total)]
    dev_tokens = [t.to(dev) for t in host_tokens]

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import contextlib

    fuse = contextlib.nullcontext if args.eager_forward else (lambda: fused_elementwise(model))

    @torch.no_grad()
    def step(tokens):
        with fuse():
            body(tokens, use_cache=False)

    def reduce_all():
        for i in range(LAYERS):
            for lst in (cov_mlp, cov_x, cov_q, cov_k):
                Dm.reduce_to_owner(lst[i], i)

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing
    with torch.no_grad():
        for w in range(args.warmup):
            step(dev_tokens[w])
        sync_all()
        recording[0] = True
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
            s.record()
            for k in range(args.steps):
                step(dev_tokens[args.warmup + k])
            if world > 1:
                reduce_all()        # the one exchange step of token-sharded calibration
            e.record()
            sync_all()
        recording[0] = False
    ms_total = max_over_ranks(s.elapsed_time(e))
    tokens_per_step = BATCH * SEQ * world
    value = tokens_per_step * args.steps / (ms_total * 1e-3)
    syrk_ms = float(np.mean([a.elapsed_time(b) for a, b in syrk_events])) if syrk_events else float("nan")
    syrk_share = float(np.sum([a.elapsed_time(b) for a, b in syrk_events])) / s.elapsed_time(e)
    flops_per_launch = BATCH * SEQ * D_INT * (D_INT + 1)      # upper triangle, 2 flop / MAC
    achieved = flops_per_launch / (syrk_ms * 1e-3) / 1e12

    # ---- end to end: pinned host tokens in, BI scores out, every step
    host_bi = torch.empty(LAYERS, dtype=torch.float64).pin_memory()
    with torch.no_grad():
        sync_all()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for k in range(args.steps):
            tok = host_tokens[args.warmup + k].to(dev, non_blocking=True)
            step(tok)
            host_bi.copy_(bi, non_blocking=True)
        if world > 1:
            reduce_all()
        e2.record()
        sync_all()
    e2e_ms = max_over_ranks(s2.elapsed_time(e2))
    e2e_value = tokens_per_step * args.steps / (e2e_ms * 1e-3)
    for h in handles:
        h.remove()
    ops.syrk_ = real_syrk
    MA.ops.syrk_ = real_syrk

    # ---- compression seconds per layer (type I / II / III on this rank's first layers)
    compress = None
    if rank == 0 and args.compress_layers > 0:
        from modegpt_b200.compression.compress_mlp import compress_nystrom
        from modegpt_b200.compression.compress_qk import compress_qk
        from modegpt_b200.compression.compress_vo import compress_vo

        n_texts = BATCH * (args.steps * 2 + args.warmup)
        scale = 1.0 / (n_texts * 2048)
        layers = list(range(min(args.compress_layers, LAYERS)))
        for i in layers:
            ops.finalize_sym_(cov_mlp[i], scale)
            ops.finalize_sym_(cov_x[i], scale)
            ops.scale_(cov_q[i], scale)
            ops.scale_(cov_k[i], scale)
        keep = [1.0 - HYPER["compression_ratio"]] * LAYERS
        stage_ms = {}
        for name, fn in (("mlp", lambda: compress_nystrom(adapter, cov_mlp, keep, layers)),
                         ("qk", lambda: compress_qk(adapter, (cov_q, cov_k), keep, target_layers=layers)),
                         ("vo", lambda: compress_vo(adapter, cov_x, keep, target_layers=layers))):
            if world > 1:
                break   # ownership would skip layers on rank 0; reported at N = 1 only
            fn()        # warm (workspace allocation, attribute setup)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            stage_ms[name] = (time.perf_counter() - t0) * 1e3 / len(layers)
        if stage_ms:
            compress = {"s_per_layer": sum(stage_ms.values()) / 1e3,
                        "ms_per_layer": stage_ms, "layers_timed": len(layers),
                        "note": "type I (n=11008, r=8256) + II + III (MHA, 32 heads) per layer, "
                                "in-memory hand-off"}
            # the same stages writing the reference's layer_{i}_{mlp,qk,vo} files through the
            # asynchronous writer, final flush included (a full run hides the writes of all but
            # the last layers behind the next layers' kernels; with so few layers it cannot)
            import shutil
            import tempfile

            tmp = tempfile.mkdtemp(prefix="mg_layers_")
            adapter.config.keep_layers_in_memory = False
            adapter.config.temp_storage_dir = tmp
            try:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                compress_nystrom(adapter, cov_mlp, keep, layers)
                compress_qk(adapter, (cov_q, cov_k), keep, target_layers=layers)
                compress_vo(adapter, cov_x, keep, target_layers=layers)
                adapter.flush_saves()
                torch.cuda.synchronize()
                compress["s_per_layer_with_files"] = (time.perf_counter() - t0) / len(layers)
                compress["file_bytes_per_layer"] = sum(
                    os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp)) / len(layers)
            finally:
                adapter.config.keep_layers_in_memory = True
                adapter._layer_cache.clear()
                shutil.rmtree(tmp, ignore_errors=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    base = cpu_reference_sample(2048) if world == 1 else None
    out = {
        "metric": "calib_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(world),
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": BATCH * SEQ * 8,
                "d2h_bytes_per_step": LAYERS * 8},
        "gpu_launches": args.steps * LAYERS * (5 if args.eager_forward else 10),
        "roofline": {"bound": "tensor", "kernel": "gemm_tn_pair_kernel, cta_group::2 256x256 tiles (C_mlp SYRK, n=11008, T=32768)",
                     "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                     "peak_source": f"{peak_src} bf16_tflops_sustained", "traffic": ncu_traffic_bytes(),
                     "algorithmic_bytes": BATCH * SEQ * D_INT * 2 + D_INT * (D_INT + 1) * 4,
                     "ms_per_launch": syrk_ms, "share_of_step": syrk_share},
        "cpu_baseline": base,
        "compress": compress,
        "clocks": clocks.summary(),
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--compress-layers", type=int, default=4)
    ap.add_argument("--eager-forward", action="store_true",
                    help="keep HF's eager elementwise kernels in the calibration forward")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the compression hot path has no CPU fallback "
                         "(use --impl reference for the host-core baseline)")
    run_gpu_arm(args)


if __name__ == "__main__":
    main()
