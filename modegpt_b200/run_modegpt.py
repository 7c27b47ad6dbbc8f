"""CLI / orchestrator: `python -m modegpt_b200.run_modegpt ...` (alias `python -m src.run_modegpt`).

Same flow and flags as the reference's `main` (src/run_modegpt.py:72-196): load -> baseline ppl ->
[calibrate -> allocate -> type I / II / III] per window of <= 48 layers -> convert_model ->
patch_config -> save -> reload through the Rebuild class -> compressed ppl -> metrics.
Under `torchrun` the calibration tokens are sharded over ranks and each layer is decomposed by its
owner; rank 0 converts, saves and evaluates.
"""
from __future__ import annotations

import gc
import logging
import os
import time

import torch

from . import distributed as D
from .adapters.CompressionConfig import CompressionConfig
from .adapters.model_adapter import ModelAdapter
from .calibration import block_influence, iter_layer_statistics, load_calibs, warm_up
from .compression.compress_mlp import compress_nystrom
from .compression.compress_qk import compress_qk
from .compression.compress_vo import compress_vo
from .compression_utils import allocate_global_sparsity
from .eval import compute_perplexity
from .model_utils import reload_compressed_model, save_compressed_model

logger = logging.getLogger("MoDeGPT")
LAYERS_PER_STEP = 48  # src/run_modegpt.py:107


def _setup_logging():
    logger.setLevel(logging.INFO)
    if not logger.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter("%(asctime)s - %(levelname)s - %(message)s"))
        logger.addHandler(h)


def _init_distributed(cfg: CompressionConfig) -> str:
    if "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        if not torch.distributed.is_initialized():
            torch.distributed.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        return f"cuda:{local}"
    torch.cuda.set_device(cfg.device)
    return f"cuda:{cfg.device}"


def _run_stages_serial(stages: dict, timings: dict, device: str) -> dict:
    """mlp, then qk, then vo — the reference's order (src/run_modegpt.py:128-151)."""
    out = {}
    t_all = time.perf_counter()
    for key, fn in stages.items():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out[key] = fn()
        torch.cuda.synchronize()
        timings[key] += time.perf_counter() - t0
    timings["compress_wall_s"] += time.perf_counter() - t_all
    return out


@torch.no_grad()
def main(trial=None, config: CompressionConfig | None = None):
    _setup_logging()
    gc.collect()
    import sys

    # kernel-launching threads (type-I workers) share the interpreter with the layer-writer
    # threads; a 5 ms switch interval lets a writer's pickling delay a launch sequence by that much
    sys.setswitchinterval(5e-4)
    config = config or CompressionConfig.from_args()
    device = _init_distributed(config)
    is_root = D.rank() == 0
    if is_root:
        print(config.to_dict())
    if not config.order:
        raise SystemExit('--order is required, e.g. --order "mlp,qk,vo"')

    model, tokenizer = reload_compressed_model(config.model, device=device)
    adapter = ModelAdapter.from_model(model=model, tokenizer=tokenizer)
    adapter.config = config
    adapter.validate_for_kernels()     # fail here, not after calibration
    adapter.metrics["note"] = config.note

    # held-out sequences are dealt to the ranks, the NLL sum is all-reduced (eval.compute_perplexity)
    if not config.skip_baseline_ppl:
        baseline = compute_perplexity(model, tokenizer, dataset=config.dataset, adapter=adapter)
        if is_root:
            logger.info(f"Baseline ppl: {baseline}")
        adapter.metrics["baseline-ppl"] = baseline

    adapter.prepare_writer()      # writer threads + 384 MB of pinned bounce buffers, before the timed stages
    n_layers = adapter.n_layers
    save_dir = os.path.join(config.output_dir, "model")
    rotary_masks: list = []
    timings = {"startup_s": 0.0, "calibration_s": 0.0, "mlp_s": 0.0, "qk_s": 0.0, "vo_s": 0.0,
               "file_flush_s": 0.0, "compress_wall_s": 0.0}

    def timed(key, fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        timings[key] += time.perf_counter() - t0
        return out

    if config.stream_layers:
        # two passes over the (resident) hidden states: BI for every layer first — every rank's
        # keep ratio depends on all of them — then one layer's statistics at a time, decomposed by
        # the layer's owner and freed before the next layer is touched
        bi_scores = timed("calibration_s", lambda: block_influence(adapter, dataset=config.dataset))
        keep = allocate_global_sparsity(bi_scores, compression_ratio=config.compression_ratio,
                                        smoothing=config.sparsity_smoothing,
                                        max_sparsity=config.max_sparsity, adapter=adapter)
        stats = iter_layer_statistics(adapter, dataset=config.dataset)
        streamed_masks: dict = {}       # gathered once after the loop: no collective per layer
        while True:
            item = timed("calibration_s", lambda: next(stats, None))
            if item is None:
                break
            l, c_mlp, c_q, c_k, c_x = item
            one = lambda t: [t if i == l else None for i in range(n_layers)]
            if "mlp" in config.order:
                timed("mlp_s", lambda: compress_nystrom(adapter=adapter, cov=one(c_mlp), keep_ratios=keep,
                                                        target_layers=[l]))
            if "qk" in config.order:
                timed("qk_s", lambda: compress_qk(adapter=adapter, cov=(one(c_q), one(c_k)), keep_ratios=keep,
                                                  target_layers=[l], local_masks=streamed_masks))
            if "vo" in config.order:
                timed("vo_s", lambda: compress_vo(adapter=adapter, cov=one(c_x), keep_ratios=keep,
                                                  target_layers=[l]))
            del item, c_mlp, c_q, c_k, c_x
        if "qk" in config.order:
            from .compression.compress_qk import gather_rotary_masks

            rotary_masks.extend(gather_rotary_masks(adapter, streamed_masks, list(range(n_layers))))
        torch.cuda.empty_cache()

    for start in (() if config.stream_layers else range(0, n_layers, LAYERS_PER_STEP)):
        target = list(range(start, min(n_layers, start + LAYERS_PER_STEP)))
        if start == 0:   # one-time start-up (kernel modules, allocator, NCCL), reported on its own
            timed("startup_s", lambda: warm_up(adapter, target))
        cov_mlp, cov_q, cov_k, cov_x, bi_scores = timed("calibration_s", lambda: load_calibs(
            adapter=adapter, n_samples=config.calib_size, batch_size=config.calibs_batch_size,
            dataset=config.dataset, target_layers=target))
        keep = allocate_global_sparsity(bi_scores, compression_ratio=config.compression_ratio,
                                        smoothing=config.sparsity_smoothing,
                                        max_sparsity=config.max_sparsity, adapter=adapter)
        stages = {}
        if "mlp" in config.order:
            stages["mlp_s"] = lambda: compress_nystrom(adapter=adapter, cov=cov_mlp, keep_ratios=keep,
                                                       target_layers=target)
        if "qk" in config.order:
            stages["qk_s"] = lambda: compress_qk(adapter=adapter, cov=(cov_q, cov_k), keep_ratios=keep,
                                                 target_layers=target)
        if "vo" in config.order:
            stages["vo_s"] = lambda: compress_vo(adapter=adapter, cov=cov_x, keep_ratios=keep,
                                                 target_layers=target)
        results = _run_stages_serial(stages, timings, device)
        rotary_masks.extend(results.get("qk_s") or [])
        del cov_mlp, cov_q, cov_k, cov_x, stages
        gc.collect()
        torch.cuda.empty_cache()

    # the layer files are written behind the decompositions (handoff.LayerWriter); the wait for the
    # last of them belongs to the compression stage ("compress s/layer incl. file write")
    timed("file_flush_s", adapter.flush_saves)
    if D.world_size() > 1:
        logger.info(f"[rank {D.rank()}] calibration {timings['calibration_s']:.2f}s mlp {timings['mlp_s']:.2f}s "
                    f"qk {timings['qk_s']:.2f}s vo {timings['vo_s']:.2f}s flush {timings['file_flush_s']:.2f}s")
    D.barrier()   # every owner has written its layer files
    tokens = config.calib_size * config.seq_len
    # stages may overlap: the wall time of the whole decomposition phase is what counts
    compress_wall = timings["compress_wall_s"] or (timings["mlp_s"] + timings["qk_s"] + timings["vo_s"])
    adapter.metrics.update({**timings, "calib_tokens_per_s": tokens / max(timings["calibration_s"], 1e-9),
                            "compress_s_per_layer": (compress_wall + timings["file_flush_s"]) / n_layers,
                            "world_size": D.world_size()})
    if is_root:
        logger.info(f"start-up {timings['startup_s']:.2f}s, calibration {timings['calibration_s']:.2f}s "
                    f"({adapter.metrics['calib_tokens_per_s']:.0f} tok/s), stages: mlp {timings['mlp_s']:.2f}s "
                    f"qk {timings['qk_s']:.2f}s vo {timings['vo_s']:.2f}s + file flush "
                    f"{timings['file_flush_s']:.2f}s -> compress {adapter.metrics['compress_s_per_layer']:.3f} "
                    f"s/layer (max over ranks is the slowest rank's line above)")
    if config.skip_rebuild:
        if is_root:
            adapter.save_metrics()
        return None
    if is_root:
        suffixes = [s for s in ("mlp", "qk", "vo") if s in config.order]
        adapter.convert_model(saved_layers_dir=config.temp_storage_dir, suffixes=suffixes)
        adapter.patch_config()
        save_compressed_model(adapter, rotary_masks=rotary_masks if "qk" in config.order else None,
                              save_dir=save_dir, source_model_name=config.model)
    del model
    adapter.model = None
    gc.collect()
    torch.cuda.empty_cache()
    D.barrier()   # the checkpoint is on disk: every rank reloads it and takes its share of the
                  # held-out sequences (token-sharded perplexity)
    model, tokenizer = reload_compressed_model(save_dir, device=device)
    adapter.model, adapter.tokenizer = model, tokenizer
    ppl = compute_perplexity(model, tokenizer, dataset=config.dataset, adapter=adapter)
    if not is_root:
        return None
    adapter.metrics[f"ppl-{config.dataset}"] = ppl
    adapter.save_metrics()
    logger.info(f"Compressed (PPL): {ppl}")
    logger.info(f"stages: mlp {timings['mlp_s']:.2f}s qk {timings['qk_s']:.2f}s vo {timings['vo_s']:.2f}s "
                f"(wall {compress_wall:.2f}s) + file flush {timings['file_flush_s']:.2f}s")
    logger.info(f"calibration {timings['calibration_s']:.2f}s "
                f"({adapter.metrics['calib_tokens_per_s']:.0f} tok/s), compress "
                f"{adapter.metrics['compress_s_per_layer']:.3f} s/layer")
    return ppl


if __name__ == "__main__":
    main()
