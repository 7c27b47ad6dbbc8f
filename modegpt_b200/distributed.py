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


class LayerReducer:
    """Reduce-as-you-go exchange of one layer's statistics.

    `submit(layer, c_mlp, c_x, c_q, c_k)` is called on the stream that produced the sums (in the
    last local batch, as soon as the layer's hooks have fired).  On a side stream it packs the two
    symmetric accumulators' upper triangles and the per-head blocks into ONE contiguous fp32
    buffer (half the bytes of the square matrices, one collective instead of four), reduces it to
    the layer's owner over NCCL and, on the owner, scatters the sum back — all of it overlapped
    with the forward of the following layers.  `wait()` joins the side stream.  Every rank submits
    the layers in the same order, so the collectives match up.

    CPU tensors (the gloo tests of this plumbing) take four plain reductions: the pack kernels
    are CUDA-only and the product path never holds statistics on the host."""

    def __init__(self):
        self._side = None
        self._buf = None

    def submit(self, layer_idx: int, c_mlp, c_x, c_q, c_k) -> bool:
        mine = owns(layer_idx)
        if not is_distributed():
            return mine
        if not c_mlp.is_cuda:
            for t in (c_mlp, c_x, c_q, c_k):
                dist.reduce(t, dst=owner_of(layer_idx), op=dist.ReduceOp.SUM)
            return mine
        from . import ops

        dev = c_mlp.device
        n1, n2 = c_mlp.shape[0], c_x.shape[0]
        sizes = [ops.packed_upper_numel(n1), ops.packed_upper_numel(n2), c_q.numel(), c_k.numel()]
        total = sum(sizes)
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        if self._buf is None or self._buf.numel() < total:
            self._buf = torch.empty(total, dtype=torch.float32, device=dev)
            self._buf.record_stream(self._side)
        ready = torch.cuda.Event()
        ready.record(main)
        o1, o2, o3 = sizes[0], sizes[0] + sizes[1], sizes[0] + sizes[1] + sizes[2]
        buf = self._buf[:total]
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            ops.pack_upper_(buf[:o1], c_mlp)
            ops.pack_upper_(buf[o1:o2], c_x)
            buf[o2:o3].copy_(c_q.reshape(-1))
            buf[o3:].copy_(c_k.reshape(-1))
            dist.reduce(buf, dst=owner_of(layer_idx), op=dist.ReduceOp.SUM)
            if mine:
                ops.unpack_upper_(c_mlp, buf[:o1])
                ops.unpack_upper_(c_x, buf[o1:o2])
                c_q.copy_(buf[o2:o3].view_as(c_q))
                c_k.copy_(buf[o3:].view_as(c_k))
        return mine

    def wait(self) -> None:
        if self._side is not None:
            torch.cuda.current_stream(self._side.device).wait_stream(self._side)


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
