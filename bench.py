#!/usr/bin/env python
"""bench.py — calibration tokens/s (+ compression s/layer) for Llama-2-7B, BASELINE.json config #2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A step = one calibration batch (calibs_batch_size 16 x 2048 synthetic tokens) through the hooked
bf16 forward of a random-init Llama-2-7B with every statistics kernel firing (C_mlp, C_x, per-head
C_q / C_k, Block-Influence) — the hot loop of src/calibration.py:114-127.  `value` times K such steps with the
token batches already in HBM, plus the cross-rank exchange and the final normalise/mirror; `e2e` is
the same work through the public API `load_calibs` on pinned host batches (H2D inside, BI scores
read back).  `strong_scaling` runs the real configuration (128 x 2048 tokens in total) through
`load_calibs`, and its statistics feed `compress`: every layer through type I/II/III by its owner
rank, layer files written, max-over-ranks wall per layer.

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
    """DRAM bytes (read + write) of one C_mlp SYRK launch from the newest committed ncu --set full
    capture of the kernel (tools/summarize_ncu.py output), and the file it came from."""
    cands = sorted((ROOT / "profiles").glob("r*_syrk*_ncu_full.json"), reverse=True)
    if not cands:
        return None, None
    p = cands[0]
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    best = None
    for k in json.loads(p.read_text())["kernels"]:
        if "gemm_tn" not in k["kernel"]:
            continue
        tot = sum(float(k[m]["value"]) * mult[k[m]["unit"]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        best = max(best or 0.0, tot)      # the n = 11008 launch is the largest launch captured
    return best, f"profiles/{p.name}"


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


def cpu_compress_sample(seed: int = 0) -> dict:
    """The compress half of the metric on the host cores: the oracle's `nystrom_mlp`
    (compress_mlp.py:28-64), `qk_layer` (compress_qk.py:208-308) and the type-III layer body
    (compress_vo.py:43-99) once each at the full Llama-2-7B shape, fp64, all host threads.
    Type III is sampled: sqrt_M + inverse of the 4096^2 statistic once, then 2 of the 32 heads,
    scaled to 32 (every head costs the same: two SVDs, the second a full 128 x 4096 one)."""
    from threadpoolctl import threadpool_limits

    from oracle import modegpt_oracle as O

    cores = os.cpu_count() or 1
    rng = np.random.default_rng(seed)
    keep = 1.0 - HYPER["compression_ratio"]

    def spd(n, rank=256):
        a = rng.standard_normal((n, rank))
        return a @ a.T / rank + np.diag(np.exp(rng.standard_normal(n)))

    with threadpool_limits(limits=cores):
        c_mlp = spd(D_INT)
        w_up = rng.standard_normal((D_INT, D)) * 0.02
        w_gate = rng.standard_normal((D_INT, D)) * 0.02
        w_down = rng.standard_normal((D, D_INT)) * 0.02
        t0 = time.perf_counter()
        O.nystrom_mlp(w_up, w_gate, w_down, c_mlp, keep, HYPER["nystrom_ridge"])
        t_mlp = time.perf_counter() - t0
        del c_mlp, w_up, w_gate, w_down
        c_q = np.stack([spd(HD, 32) for _ in range(HEADS)])
        c_k = np.stack([spd(HD, 32) for _ in range(KV)])
        w_q = rng.standard_normal((HEADS * HD, D)) * 0.02
        w_k = rng.standard_normal((KV * HD, D)) * 0.02
        r = O.head_rank(HD, keep, True)
        t0 = time.perf_counter()
        O.qk_layer(w_q, w_k, c_q, c_k, HEADS, KV, HD, r, "llama", HYPER["ridge_qk"])
        t_qk = time.perf_counter() - t0
        c_x = spd(D)
        w_v = rng.standard_normal((KV * HD, D)) * 0.02
        w_o = rng.standard_normal((D, HEADS * HD)) * 0.02
        t0 = time.perf_counter()
        root, root_inv = O.vo_roots(c_x, HYPER["ridge_vo"])
        t_root = time.perf_counter() - t0
        sample_heads = 2
        t0 = time.perf_counter()
        for h in range(sample_heads):
            O.vo_head_mha(w_v[h * HD:(h + 1) * HD], w_o[:, h * HD:(h + 1) * HD], root, root_inv, r)
        t_head = (time.perf_counter() - t0) / sample_heads
        t_vo = t_root + t_head * HEADS
    return {"s_per_layer": t_mlp + t_qk + t_vo, "unit": "s/layer", "cores": cores, "kind": "port",
            "ms_per_layer": {"mlp": 1e3 * t_mlp, "qk": 1e3 * t_qk, "vo": 1e3 * t_vo},
            "sample": f"fp64 oracle, one Llama-2-7B layer: nystrom_mlp n=11008 r={int(D_INT * keep)} "
                      f"({t_mlp:.1f} s), qk_layer 32 heads ({t_qk:.2f} s), type III = sqrt_M+inv of 4096^2 "
                      f"({t_root:.1f} s) + {sample_heads} of {HEADS} heads timed ({t_head:.2f} s each) "
                      f"scaled to {HEADS}; no file write"}


def fs_write_gbs(directory: str, threads: int = 8, mb_each: int = 128) -> float:
    """Aggregate write bandwidth of the file system the layer files go to (`threads` writers, raw
    bytes, no serialisation): the floor under "compress s/layer incl. file write" on this box."""
    buf = np.random.default_rng(0).integers(0, 255, mb_each << 20, dtype=np.uint8).tobytes()

    def one(i):
        with open(os.path.join(directory, f"_probe{i}"), "wb") as f:
            f.write(buf)

    ths = [threading.Thread(target=one, args=(i,)) for i in range(threads)]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    for i in range(threads):
        os.remove(os.path.join(directory, f"_probe{i}"))
    return threads * len(buf) / dt / 1e9


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
    base["compress"] = cpu_compress_sample()
    print(json.dumps({
        "impl": "reference", "metric": "calib_tokens_per_s", "value": v, "unit": "tokens/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args.gpus), "cpu_baseline": base,
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "compress": {"s_per_layer": base["compress"]["s_per_layer"], "ms_per_layer": base["compress"]["ms_per_layer"]},
    }))


def workload_config(n_gpus: int) -> dict:
    return {"workload": "Llama-2-7B 25% MoDeGPT calibration, 16x2048 synthetic tokens per step "
                        "(BASELINE configs[1]); random-init weights",
            "calibs_batch_size": BATCH, "seq_len": SEQ, "layers": LAYERS,
            "parallelism": f"token-sharded x{n_gpus}; each layer's sums packed (upper triangles) and "
                           f"reduced to its owner during the last local batch; layers decomposed by owner",
            "l2_policy": "inputs larger than L2 (each statistics operand is 268-721 MB)",
            "forward": "HF modules, fused RMSNorm/SwiGLU/RoPE kernels (bit-compatible)", **HYPER}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args) -> None:
    import gc
    import shutil
    import tempfile

    import torch.distributed as dist

    import modegpt_b200.adapters.model_adapter as MA
    from modegpt_b200 import ops
    from modegpt_b200.adapters.CompressionConfig import CompressionConfig
    from modegpt_b200.adapters.model_adapter import ModelAdapter
    from modegpt_b200.calibration import Calibrator, load_calibs
    from modegpt_b200.compression.compress_mlp import compress_nystrom
    from modegpt_b200.compression.compress_qk import compress_qk
    from modegpt_b200.compression.compress_vo import compress_vo
    from modegpt_b200.compression_utils import allocate_global_sparsity
    from modegpt_b200.model_utils import build_synthetic_model

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    tf_peak, hbm_peak, peak_src = peaks()
    sys.setswitchinterval(5e-4)     # as run_modegpt.main does (launch threads vs writer threads)

    model = build_synthetic_model(PRESET, device=str(dev), seed=0)
    adapter = ModelAdapter.from_model(model, tokenizer=None)
    adapter.config = CompressionConfig(model=f"synthetic:{PRESET}", dataset="synthetic", order="mlp,qk,vo",
                                       calib_size=128, calibs_batch_size=BATCH, eager_forward=args.eager_forward,
                                       **HYPER)
    all_layers = list(range(LAYERS))

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def release():
        gc.collect()      # blocks stay in torch's caching allocator: the next phase reuses them

    # time the dominant kernel (the 11008-wide SYRK) live: events around every C_mlp launch
    syrk_events: list = []
    real_syrk = ops.syrk_
    recording = [False]

    def timed_syrk(C, X, alpha=1.0, accumulate=True):
        if C.shape[0] == D_INT and recording[0]:
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            real_syrk(C, X, alpha, accumulate)
            e.record()
            syrk_events.append((s, e))
        else:
            real_syrk(C, X, alpha, accumulate)

    steps_total = args.warmup + args.steps
    vocab = model.config.vocab_size
    g = torch.Generator().manual_seed(1234 + rank)
    dev_tokens = [torch.randint(0, vocab, (BATCH, SEQ), generator=g).to(dev) for _ in range(steps_total)]

    # ---- A. device-resident: K hooked forwards + the cross-rank exchange + normalise/mirror
    ops.syrk_ = MA.ops.syrk_ = timed_syrk
    try:
        cal = Calibrator(adapter, all_layers)
        for w in range(args.warmup):
            cal.run_batch(dev_tokens[w])
        sync_all()
        recording[0] = True
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_fwd = torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clocks:
            s.record()
            for k in range(args.steps):
                cal.run_batch(dev_tokens[args.warmup + k], last=(k == args.steps - 1))
            s_fwd.record()
            recording[0] = False
            stats_a = cal.finish()        # join the reductions, BI all-reduce, scale + mirror
            e.record()
            sync_all()
    finally:
        ops.syrk_ = MA.ops.syrk_ = real_syrk
    ms_total = max_over_ranks(s.elapsed_time(e))
    ms_tail = max_over_ranks(s_fwd.elapsed_time(e))
    tokens_per_step = BATCH * SEQ * world
    value = tokens_per_step * args.steps / (ms_total * 1e-3)
    syrk_ms = float(np.mean([a.elapsed_time(b) for a, b in syrk_events])) if syrk_events else float("nan")
    syrk_share = float(np.sum([a.elapsed_time(b) for a, b in syrk_events])) / s.elapsed_time(e)
    flops_per_launch = BATCH * SEQ * D_INT * (D_INT + 1)      # upper triangle, 2 flop / MAC
    achieved = flops_per_launch / (syrk_ms * 1e-3) / 1e12
    del stats_a, cal, dev_tokens
    release()

    # ---- B. end to end through the public API: load_calibs() on PINNED HOST token batches
    # (H2D of every batch, hooked forwards, exchange, normalise/mirror, BI scores read back)
    def host_batches(n_batches, seed):
        gg = torch.Generator().manual_seed(seed)
        return [torch.randint(0, vocab, (BATCH, SEQ), generator=gg).pin_memory() for _ in range(n_batches)]

    def timed_load_calibs(batches):
        adapter.calibs = batches
        sync_all()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        out = load_calibs(adapter, n_samples=len(batches) * BATCH, batch_size=BATCH, dataset="synthetic",
                          target_layers=all_layers)
        e2.record()
        sync_all()
        return max_over_ranks(s2.elapsed_time(e2)), out

    e2e_ms, stats_b = timed_load_calibs(host_batches(args.steps * world, 4321))
    e2e_value = tokens_per_step * args.steps / (e2e_ms * 1e-3)
    del stats_b
    release()

    # ---- C. strong scaling at the real configuration: BASELINE configs[1]'s 128 x 2048 tokens in
    # total, split over the ranks, again through load_calibs; its statistics feed the compress stage
    n_real = 128 // BATCH
    strong_ms, (cov_mlp, cov_q, cov_k, cov_x, bi_scores) = timed_load_calibs(host_batches(n_real, 1234))
    strong = {"tokens": 128 * SEQ, "seconds": strong_ms * 1e-3, "tokens_per_s": 128 * SEQ / (strong_ms * 1e-3),
              "batches_per_rank": n_real / world,
              "note": "load_calibs on 128x2048 tokens in total (strong scaling; weak scaling is `value`)"}

    # ---- D. compress seconds per layer: type I / II / III of every layer, each by its owner,
    # layer files written (the metric says "incl. file write"); wall = max over ranks
    compress = None
    if args.compress_layers > 0:
        layers = all_layers[:args.compress_layers]
        keep = allocate_global_sparsity(bi_scores, compression_ratio=HYPER["compression_ratio"],
                                        smoothing=HYPER["sparsity_smoothing"],
                                        max_sparsity=HYPER["max_sparsity"], adapter=adapter)
        owned = [l for l in layers if l % world == rank]

        def stages(which):
            t = {}
            for name, fn in (("mlp", lambda: compress_nystrom(adapter, cov_mlp, keep, which)),
                             ("qk", lambda: compress_qk(adapter, (cov_q, cov_k), keep, target_layers=which)),
                             ("vo", lambda: compress_vo(adapter, cov_x, keep, target_layers=which))):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                t[name] = time.perf_counter() - t0
            return t

        adapter.config.keep_layers_in_memory = True
        stages(owned[:1])                        # warm: workspaces, attributes, lanes
        adapter._layer_store.clear()
        tmp = tempfile.mkdtemp(prefix="mg_layers_")
        adapter.config.keep_layers_in_memory = False
        adapter.config.temp_storage_dir = tmp
        adapter.prepare_writer()
        try:
            sync_all()
            t0 = time.perf_counter()
            st_files = stages(layers)            # each stage takes the layers this rank owns
            tf0 = time.perf_counter()
            adapter.flush_saves()
            torch.cuda.synchronize()
            t_flush = time.perf_counter() - tf0
            wall_files = max_over_ranks(time.perf_counter() - t0)
            file_bytes = sum(os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp))
            fs_gbs = fs_write_gbs(tmp) if rank == 0 else 0.0
        finally:
            adapter._layer_cache.clear()
            shutil.rmtree(tmp, ignore_errors=True)
        adapter.config.keep_layers_in_memory = True
        sync_all()
        t0 = time.perf_counter()
        st_mem = stages(layers)
        wall_mem = max_over_ranks(time.perf_counter() - t0)
        adapter._layer_store.clear()
        n_own = max(len(owned), 1)
        compress = {
            "s_per_layer": wall_files / len(layers), "unit": "s/layer",
            "definition": "max over ranks of wall(type I + II + III of the rank's layers + layer-file "
                          "writes + final flush) / layers",
            "layers_timed": len(layers), "layers_per_rank": len(owned),
            "ms_per_layer_rank0": {k: 1e3 * v / n_own for k, v in st_files.items()},
            "flush_s_rank0": t_flush, "file_bytes_per_layer": file_bytes / n_own,
            "fs_write_gbs_8_threads": fs_gbs,
            "writer_thread_seconds": dict(adapter._writer.stats) if adapter._writer is not None else None,
            "fs_floor_s_per_layer": (file_bytes / n_own) / (fs_gbs * 1e9) if fs_gbs else None,
            "in_memory": {"s_per_layer": wall_mem / len(layers),
                          "ms_per_layer_rank0": {k: 1e3 * v / n_own for k, v in st_mem.items()}},
            "note": "type I (n=11008, r~8256) + II + III (MHA, 32 heads) per layer; keep ratios from the "
                    "strong-scaling calibration's BI scores"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    base = None
    if world == 1:
        base = cpu_reference_sample(2048)
        base["compress"] = cpu_compress_sample()
    traffic, traffic_src = ncu_traffic_bytes()
    out = {
        "metric": "calib_tokens_per_s", "value": value, "unit": "tokens/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic", "config": workload_config(world),
        "timed_region": "K hooked forwards + cross-rank exchange + BI all-reduce + normalise/mirror "
                        f"(tail after the last forward: {ms_tail:.1f} ms)",
        "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": BATCH * SEQ * 8 * world,
                "d2h_bytes_per_step": (LAYERS * 8 + 8) * world / args.steps,
                "api": "modegpt_b200.calibration.load_calibs(adapter, n_samples, batch_size, 'synthetic', "
                       "target_layers) with adapter.calibs = pinned host batches; BI scores read back once per call"},
        "strong_scaling": strong,
        "gpu_launches": args.steps * LAYERS * (5 if args.eager_forward else 10) + 4 * LAYERS,
        "roofline": {"bound": "tensor", "kernel": "gemm_tn_pair_kernel, cta_group::2 256x256 tiles (C_mlp SYRK, n=11008, T=32768)",
                     "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                     "peak_source": f"{peak_src} bf16_tflops_sustained", "traffic": traffic,
                     "traffic_source": traffic_src,
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
    ap.add_argument("--compress-layers", type=int, default=LAYERS,
                    help="layers put through type I/II/III for the s/layer figure (0 = skip)")
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
