"""Type-I: Nystrom compression of the MLP (reference: src/compression/compress_mlp.py).

    scores = diag((C + lambda I)^-1)            blocked Cholesky + blocked triangular inverse
    idx    = the `rank` smallest scores, ascending index
    W_up' = W_up[idx], W_gate' = W_gate[idx]    row gathers
    W_down' = ((C[idx,idx] + 1e-6 I)^-1 C[idx,:] W_down^T)^T

All four steps are kernels of libmodegpt_b200 (see include/modegpt_b200.h); this module only
sequences them per layer and writes the `layer_{i}_mlp` hand-off file.
"""
from __future__ import annotations

import logging

import numpy as np
import torch
from torch import Tensor

from .. import distributed as D
from .. import ops
from ..adapters.model_adapter import MLPComponents, ModelAdapter

logger = logging.getLogger("MoDeGPT")

NYSTROM_JITTER = 1e-6  # compress_mlp.py:56


def get_ridge_scores(C: Tensor, layer_idx: int, ridge_lambda: float = 1e-2) -> Tensor:
    """diag((C + ridge I)^-1), fp32 [n] (compress_mlp.py:13-25).  The reference's ridge reaches its
    fp64 sum as float32(ridge) (`ridge * torch.eye(n)` is fp32); same value here."""
    return ops.ridge_scores(C, float(np.float32(ridge_lambda)), what=f"layer {layer_idx} ridge scores")


@torch.no_grad()
def compress_weights(comps: MLPComponents, C: Tensor, keep_ratio: float, layer_idx: int,
                     ridge_lambda: float):
    """Returns (W_up'^T [d, r], W_down'^T [r, d], W_gate'^T [d, r] | None, rank, idx) — the first
    four as in the reference (compress_mlp.py:28-64; it returns transposes and `compress_nystrom`
    transposes back), plus the kept indices (needed for the OPT fc1 bias)."""
    n = C.shape[0]
    rank = int(n * keep_ratio)
    scores = get_ridge_scores(C, layer_idx, ridge_lambda)
    idx = ops.select_k(scores, rank, largest=False)
    up = ops.gather_rows(comps.up_proj.weight.detach(), idx)
    gate = None
    if comps.gate_proj is not None:
        gate = ops.gather_rows(comps.gate_proj.weight.detach(), idx)
    stats: dict = {}
    down = ops.nystrom_down(C, idx, comps.down_proj.weight.detach().contiguous(), NYSTROM_JITTER, stats=stats)
    if stats.get("refine_sweeps"):
        logger.info(f"[MLP] Layer {layer_idx}: C_kk min relative pivot {stats['min_rel_pivot']:.1e} -> "
                    f"{stats['refine_sweeps']} fp64-residual refinement sweep(s)")
    return up.T, down.T, (gate.T if gate is not None else None), rank, idx


def _compress_layer(adapter: ModelAdapter, cov, keep_ratios, layer_idx: int) -> None:
    comps = adapter.get_mlp_components(layer_idx)
    up_t, down_t, gate_t, rank, idx = compress_weights(
        comps, cov[layer_idx], keep_ratios[layer_idx], layer_idx=layer_idx,
        ridge_lambda=adapter.config.nystrom_ridge)
    logger.info(f"[MLP] Layer {layer_idx} compressed to rank {rank}")
    weights = {"up": up_t.T, "down": down_t.T}
    if gate_t is not None:
        weights["gate"] = gate_t.T
    if getattr(comps.up_proj, "bias", None) is not None:      # OPT: fc1 bias follows its rows
        weights["up_bias"] = comps.up_proj.bias.detach()[idx]
    if getattr(comps.down_proj, "bias", None) is not None:    # fc2 bias is kept
        weights["down_bias"] = comps.down_proj.bias.detach()
    adapter.save_layer(output_dir=adapter.config.temp_storage_dir, suffix="mlp",
                       weights=weights, layer_idx=layer_idx)


@torch.no_grad()
def compress_nystrom(adapter: ModelAdapter, cov, keep_ratios, target_layers, ridge_lambda=1e-4):
    """Layer loop + `save_layer(suffix="mlp")` (compress_mlp.py:67-117).  As in the reference the
    ridge actually used is `adapter.config.nystrom_ridge` (:93).  Each rank takes the layers it
    owns (`distributed.owned_layers`).

    A factorisation is bound by its panel chain (about 12 ms of tensor work spread over 21 ms), so
    layers — which are independent — are decomposed `config.mlp_workers` at a time: that many host
    threads, each with its own CUDA stream and, inside the library, its own set of lanes, take
    layers from a shared list.  `mg_set_concurrent_factorizations` divides the SMs left after the
    chain reserve between their bulk GEMMs: in round 1 two workers gained nothing because one
    factorisation's persistent trailing updates filled the GPU and the other's chain kernels
    queued behind them.  Results do not depend on the number of workers."""
    from .._lib import lib

    layers = list(D.owned_layers(target_layers))
    workers = max(1, min(int(getattr(adapter.config, "mlp_workers", 1)), len(layers)))
    if workers == 1 or not layers or not cov[layers[0]].is_cuda:
        for layer_idx in layers:
            _compress_layer(adapter, cov, keep_ratios, layer_idx)
        return

    import threading

    adapter.prepare_writer()                # created once, before the workers race for it
    device = cov[layers[0]].device
    caller = torch.cuda.current_stream(device)
    ready = torch.cuda.Event()
    ready.record(caller)
    pending = list(reversed(layers))        # pop() hands layers out in ascending order
    lock = threading.Lock()
    errors: list[BaseException] = []
    streams = _worker_streams(device, workers)

    def work(stream):
        try:
            torch.cuda.set_device(device)
            with torch.no_grad(), torch.cuda.stream(stream):
                stream.wait_event(ready)     # statistics / weights produced on the caller's stream
                while not errors:
                    with lock:
                        if not pending:
                            break
                        layer_idx = pending.pop()
                    _compress_layer(adapter, cov, keep_ratios, layer_idx)
        except BaseException as e:           # re-raised on the calling thread
            errors.append(e)

    lib.mg_set_concurrent_factorizations(workers)
    try:
        threads = [threading.Thread(target=work, args=(s,), name=f"mg-mlp-{i}") for i, s in enumerate(streams)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for s in streams:
            caller.wait_stream(s)
    finally:
        lib.mg_set_concurrent_factorizations(1)
    if errors:
        raise errors[0]


_WORKER_STREAMS: dict = {}


def _worker_streams(device, n: int) -> list:
    pool = _WORKER_STREAMS.setdefault(torch.device(device), [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=device))
    return pool[:n]
