"""Stage hand-off: asynchronous `layer_{i}_{suffix}` files.

The reference writes every decomposed layer with a blocking `torch.save` of CUDA tensors
(src/adapters/model_adapter.py:184-191) and `convert_model` reads all of them back (:193-237);
on Llama-2-7B that is 10 GB through one thread's pageable copies and pickling — several times the
GPU time of the decompositions themselves.  `LayerWriter` keeps the files (same names, same
`torch.load`-able dicts) but takes them off the critical path:

    submit():  records an event on the producing stream and queues the job (no copy, no sync);
    stagers:   three threads, each with its own CUDA stream and two small pinned bounce buffers:
               wait for the event on that stream, stream each tensor to host memory in 64 MB
               slices (device -> pinned asynchronously, pinned -> pageable by memcpy, double-
               buffered; one thread tops out near 5 GB/s on first-touch page faults) and pass the
               host copies on;
    savers:    a small pool of threads that `torch.save` them;
    flush():   drains both queues and re-raises the first error.

Measured alternatives (Llama-2-7B, 10 GB of layer files): per-layer pinned staging buffers pin
memory at ~1 GB/s and stall every other CUDA call meanwhile (+3 s on the calibration it overlapped);
pageable `tensor.cpu()` from the saver threads slowed the kernel-launching threads 2.5x.  The
bounce buffers cost 3 x 128 MB of pinned memory once.  `submit` blocks when `max_in_flight` jobs are
pending, which bounds the device and host memory held by results waiting to be written.
"""
from __future__ import annotations

import os
import queue
import threading

import torch
from torch import Tensor

_SLICE = 64 << 20


class LayerWriter:
    def __init__(self, n_threads: int = 8, max_in_flight: int = 16, device=None, n_stagers: int = 3):
        self._stage_q: queue.Queue = queue.Queue()
        self._save_q: queue.Queue = queue.Queue()
        self._slots = threading.Semaphore(max_in_flight)
        self._errors: list[BaseException] = []
        self.bytes_written = 0
        # the zip writer's CRC-32 costs about as much as the write itself; torch.load does not check it
        self._crc_prev = None
        ser = torch.serialization
        if hasattr(ser, "set_crc32_options") and hasattr(ser, "get_crc32_options"):
            self._crc_prev = ser.get_crc32_options()
            ser.set_crc32_options(False)
        self._warm_device = torch.device(device) if device is not None else None   # pin the bounce buffers now
        self._n_stagers = n_stagers
        self._threads = [threading.Thread(target=self._stager, daemon=True, name=f"mg-stager-{i}")
                         for i in range(n_stagers)]
        self._threads += [threading.Thread(target=self._saver, daemon=True, name=f"mg-saver-{i}")
                          for i in range(n_threads)]
        for t in self._threads:
            t.start()

    # ------------------------------------------------------------------------------------------
    def submit(self, path: str, weights: dict[str, Tensor]) -> None:
        self._slots.acquire()                           # blocks while too many results are pending
        ready = None
        cuda = [w for w in weights.values() if w.is_cuda]
        if cuda:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(cuda[0].device))
        self._stage_q.put((path, dict(weights), ready))

    def flush(self) -> None:
        self._stage_q.join()
        self._save_q.join()
        if self._errors:
            err = self._errors[0]
            self._errors.clear()
            raise RuntimeError(f"layer writer failed: {err!r}") from err

    def close(self) -> None:
        self.flush()
        for _ in range(self._n_stagers):
            self._stage_q.put(None)
        for _ in self._threads[self._n_stagers:]:
            self._save_q.put(None)
        for t in self._threads:
            t.join()
        self._threads = []
        if self._crc_prev is not None:
            torch.serialization.set_crc32_options(self._crc_prev)
            self._crc_prev = None

    # ------------------------------------------------------------------------------------------
    def _stager(self) -> None:
        stream, bounce, events = None, None, None
        if self._warm_device is not None and self._warm_device.type == "cuda":
            try:
                torch.cuda.set_device(self._warm_device)   # threads start on device 0: keep rank r on GPU r
                stream = torch.cuda.Stream(device=self._warm_device)
                bounce = [torch.empty(_SLICE, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
                events = [torch.cuda.Event() for _ in range(2)]
            except BaseException as e:
                self._errors.append(e)
                stream = None
        while True:
            job = self._stage_q.get()
            if job is None:
                self._stage_q.task_done()
                return
            path, weights, ready = job
            try:
                if ready is not None:
                    dev = next(w.device for w in weights.values() if w.is_cuda)
                    if stream is None or stream.device != dev:
                        torch.cuda.set_device(dev)
                        stream = torch.cuda.Stream(device=dev)
                        bounce = [torch.empty(_SLICE, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
                        events = [torch.cuda.Event() for _ in range(2)]
                    with torch.cuda.stream(stream):
                        stream.wait_event(ready)
                        weights = {k: self._to_host(w, stream, bounce, events) for k, w in weights.items()}
                self._save_q.put((path, weights))
            except BaseException as e:                  # surfaced by flush()
                self._errors.append(e)
                self._slots.release()
            finally:
                del weights, job
                self._stage_q.task_done()

    @staticmethod
    def _to_host(w: Tensor, stream, bounce, events) -> Tensor:
        if not w.is_cuda:
            return w
        # dense source bytes: the tensor itself, or its transpose for [r, d] views of [d, r] results
        transposed = w.dim() == 2 and not w.is_contiguous() and w.T.is_contiguous()
        src = w.T if transposed else w.contiguous()
        dst = torch.empty(src.shape, dtype=src.dtype)
        s8, d8 = src.reshape(-1).view(torch.uint8), dst.reshape(-1).view(torch.uint8)
        n, pos, turn, pending = s8.numel(), 0, 0, []
        while pos < n or pending:
            if pos < n and len(pending) < 2:
                b, size = turn & 1, min(_SLICE, n - pos)
                bounce[b][:size].copy_(s8[pos:pos + size], non_blocking=True)
                events[b].record(stream)
                pending.append((b, pos, size))
                pos, turn = pos + size, turn + 1
                continue
            b, off, size = pending.pop(0)
            events[b].synchronize()
            d8[off:off + size].copy_(bounce[b][:size])
        return dst.T if transposed else dst

    def _saver(self) -> None:
        while True:
            job = self._save_q.get()
            if job is None:
                self._save_q.task_done()
                return
            path, weights = job
            try:
                os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
                torch.save(weights, path)
                self.bytes_written += sum(v.numel() * v.element_size() for v in weights.values())
            except BaseException as e:                  # surfaced by flush()
                self._errors.append(e)
            finally:
                del weights, job
                self._slots.release()
                self._save_q.task_done()
