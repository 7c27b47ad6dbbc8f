"""Stage hand-off: asynchronous `layer_{i}_{suffix}` files.

The reference writes every decomposed layer with a blocking `torch.save` of CUDA tensors
(src/adapters/model_adapter.py:184-191) and `convert_model` reads all of them back (:193-237);
on Llama-2-7B that is 10 GB through pageable copies and one pickling thread — several times the
GPU time of the decompositions themselves.  `LayerWriter` keeps the files (same names, same
`torch.load`-able dicts) but takes them off the critical path:

    submit():  device -> pinned staging copy on a side stream (ordered after the producing stream),
               then the job goes to a small pool of writer threads;
    worker:    waits for the copy event, detaches the staged views into private tensors and
               `torch.save`s them;
    flush():   drains the queue and re-raises the first writer error.

Staging buffers are a bounded pool of pinned allocations that only grow (pinning is slow, so it is
done a handful of times per run, not per layer); `submit` blocks when every buffer is in flight,
which bounds host memory.
"""
from __future__ import annotations

import os
import queue
import threading

import torch
from torch import Tensor


class LayerWriter:
    def __init__(self, n_threads: int = 4, n_buffers: int = 6):
        self._jobs: queue.Queue = queue.Queue()
        self._free: queue.Queue = queue.Queue()
        for _ in range(n_buffers):
            self._free.put(None)            # slots; the pinned tensor is allocated on first use
        self._errors: list[BaseException] = []
        self._threads = [threading.Thread(target=self._worker, daemon=True, name=f"mg-writer-{i}")
                         for i in range(n_threads)]
        for t in self._threads:
            t.start()
        self._copy_stream: torch.cuda.Stream | None = None
        self.bytes_written = 0

    # ------------------------------------------------------------------------------------------
    def submit(self, path: str, weights: dict[str, Tensor]) -> None:
        cuda = [w for w in weights.values() if w.is_cuda]
        if not cuda:                                    # CPU tensors (tests): nothing to stage
            self._jobs.put((path, dict(weights), None, None, None))
            return
        dev = cuda[0].device
        sizes = {k: w.numel() * w.element_size() for k, w in weights.items()}
        total = sum((s + 255) // 256 * 256 for s in sizes.values())
        buf = self._free.get()                          # blocks while every buffer is in flight
        if buf is None or buf.numel() < total:
            buf = torch.empty(max(total, 1), dtype=torch.uint8, pin_memory=True)
        if self._copy_stream is None or self._copy_stream.device != dev:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        cs.wait_stream(torch.cuda.current_stream(dev))
        staged, off = {}, 0
        with torch.cuda.stream(cs):
            for k, w in weights.items():
                view = buf[off:off + sizes[k]].view(w.dtype).view(w.shape)
                view.copy_(w, non_blocking=True)        # strided sources (transposed views) are fine
                staged[k] = view
                off += (sizes[k] + 255) // 256 * 256
            done = torch.cuda.Event()
            done.record(cs)
        # `weights` rides along so the device tensors outlive the copy
        self._jobs.put((path, staged, done, buf, weights))

    def flush(self) -> None:
        self._jobs.join()
        if self._errors:
            err = self._errors[0]
            self._errors.clear()
            raise RuntimeError(f"layer writer failed: {err!r}") from err

    def close(self) -> None:
        self.flush()
        for _ in self._threads:
            self._jobs.put(None)
        for t in self._threads:
            t.join()
        self._threads = []

    # ------------------------------------------------------------------------------------------
    def _worker(self) -> None:
        while True:
            job = self._jobs.get()
            if job is None:
                self._jobs.task_done()
                return
            path, staged, done, buf, keep = job
            try:
                if done is not None:
                    done.synchronize()
                del keep
                # private copies: torch.save would otherwise serialise the whole staging buffer
                out = {k: v.clone() for k, v in staged.items()} if buf is not None else staged
                if buf is not None:
                    self._free.put(buf)
                    buf = None
                os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
                torch.save(out, path)
                self.bytes_written += sum(v.numel() * v.element_size() for v in out.values())
            except BaseException as e:                  # surfaced by flush()
                self._errors.append(e)
            finally:
                if buf is not None:
                    self._free.put(buf)
                self._jobs.task_done()
