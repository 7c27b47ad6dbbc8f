"""One-process-per-GPU plumbing for the two ways the path shards:

  * calibration batches are dealt round-robin to ranks; every rank accumulates its own fp32
    statistics and ONE reduction per layer sends the sum to the layer's owner;
  * decompositions are independent per layer: layer l is owned by rank l % world.

`torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is the only transport; nothing here
touches tensors' contents.  The reference has no distributed code at all.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.distributed as dist


_force_single = False   # tests: run the single-process semantics inside an initialised group


def is_distributed() -> bool:
    return (not _force_single and dist.is_available() and dist.is_initialized()
            and dist.get_world_size() > 1)


def rank() -> int:
    return dist.get_rank() if is_distributed() else 0


def world_size() -> int:
    return dist.get_world_size() if is_distributed() else 1


def owner_of(layer_idx: int) -> int:
    return layer_idx % world_size()


def owns(layer_idx: int) -> bool:
    return owner_of(layer_idx) == rank()


def owned_layers(layers: Iterable[int]) -> list[int]:
    return [l for l in layers if owns(l)]


def shard_batches(batches: Sequence) -> list:
    """Batches b with b % world == rank, order preserved (fp32 sums differ from the single-GPU run
    in summation order only)."""
    return [b for i, b in enumerate(batches) if i % world_size() == rank()]


def reduce_to_owner(t: torch.Tensor, layer_idx: int) -> bool:
    """Sum `t` over ranks into the owner's copy.  Returns True on the owner."""
    if is_distributed():
        dist.reduce(t, dst=owner_of(layer_idx), op=dist.ReduceOp.SUM)
    return owns(layer_idx)


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if is_distributed():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def gather_by_layer(local: dict, n_layers: int) -> list:
    """Every rank contributes {layer_idx: object}; every rank receives the full list in layer
    order (objects are small: rotary masks, ranks)."""
    if not is_distributed():
        return [local.get(i) for i in range(n_layers)]
    parts = [None] * world_size()
    dist.all_gather_object(parts, local)
    merged = {}
    for p in parts:
        merged.update(p)
    return [merged.get(i) for i in range(n_layers)]


def gather_masks(local: dict, n_layers: int, rows: int, width: int, device) -> list:
    """Rotary masks ([rows, r] int64 with r <= width, one per layer, each owned by one rank) to every
    rank in layer order: one all-reduce of a padded [n_layers, 2 + rows * width] int64 tensor —
    each layer has exactly one owner, so the sum is the gather (an object all-gather costs ~0.7 s
    of pickling and extra collectives per call)."""
    if not is_distributed():
        return [local.get(i) for i in range(n_layers)]
    buf = torch.zeros(n_layers, 2 + rows * width, dtype=torch.int64, device=device)
    for i, m in local.items():
        r = m.shape[1]
        buf[i, 0] = 1
        buf[i, 1] = r
        buf[i, 2:2 + rows * r] = m.reshape(-1).to(device)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    head = buf[:, :2].cpu()
    out = []
    for i in range(n_layers):
        if int(head[i, 0]) == 0:
            out.append(None)
        else:
            r = int(head[i, 1])
            out.append(buf[i, 2:2 + rows * r].reshape(rows, r).clone())
    return out


def barrier() -> None:
    if is_distributed():
        dist.barrier()
