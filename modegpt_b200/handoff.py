"""Stage hand-off: asynchronous `layer_{i}_{suffix}` files.

The reference writes every decomposed layer with a blocking `torch.save` of CUDA tensors
(src/adapters/model_adapter.py:184-191) and `convert_model` reads all of them back (:193-237);
on Llama-2-7B that is 10 GB through one thread's pageable copies and pickling — several times the
GPU time of the decompositions themselves.  `LayerWriter` keeps the files (same names, same
`torch.load`-able dicts) but takes them off the critical path:

    submit():  records an event on the producing stream and queues the job (no copy, no sync);
    worker:    one of a small pool of threads, each with its own CUDA stream: waits for the event
               on that stream, copies the tensors to host memory and `torch.save`s them;
    flush():   drains the queue and re-raises the first writer error.

The device->host copies land in ordinary pageable memory: pinning staging buffers was measured at
~1 GB/s and stalls every other CUDA call of the process while it runs, which cost more than the
copies save.  `submit` blocks when `max_in_flight` jobs are queued, which bounds the device memory
held by results that are waiting to be written.
"""
from __future__ import annotations

import os
import queue
import threading

import torch
from torch import Tensor


class LayerWriter:
    def __init__(self, n_threads: int = 8, max_in_flight: int = 16):
        self._jobs: queue.Queue = queue.Queue()
        self._slots = threading.Semaphore(max_in_flight)
        self._errors: list[BaseException] = []
        self.bytes_written = 0
        # the zip writer's CRC-32 costs about as much as the write itself; torch.load does not check it
        self._crc_prev = None
        ser = torch.serialization
        if hasattr(ser, "set_crc32_options") and hasattr(ser, "get_crc32_options"):
            self._crc_prev = ser.get_crc32_options()
            ser.set_crc32_options(False)
        self._threads = [threading.Thread(target=self._worker, daemon=True, name=f"mg-writer-{i}")
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
        self._jobs.put((path, dict(weights), ready))

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
        if self._crc_prev is not None:
            torch.serialization.set_crc32_options(self._crc_prev)
            self._crc_prev = None

    # ------------------------------------------------------------------------------------------
    def _worker(self) -> None:
        stream = None
        while True:
            job = self._jobs.get()
            if job is None:
                self._jobs.task_done()
                return
            path, weights, ready = job
            try:
                if ready is not None:
                    dev = next(w.device for w in weights.values() if w.is_cuda)
                    if stream is None or stream.device != dev:
                        stream = torch.cuda.Stream(device=dev)
                    with torch.cuda.stream(stream):
                        stream.wait_event(ready)
                        # pageable destination: the copy blocks this thread (only) until it is done
                        weights = {k: w.to("cpu") for k, w in weights.items()}
                os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
                torch.save(weights, path)
                self.bytes_written += sum(v.numel() * v.element_size() for v in weights.values())
            except BaseException as e:                  # surfaced by flush()
                self._errors.append(e)
            finally:
                del weights, job
                self._slots.release()
                self._jobs.task_done()
