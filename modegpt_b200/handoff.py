"""Stage hand-off: asynchronous `layer_{i}_{suffix}` files.

The reference writes every decomposed layer with a blocking `torch.save` of CUDA tensors
(src/adapters/model_adapter.py:184-191) and `convert_model` reads all of them back (:193-237);
on Llama-2-7B that is 10 GB through one thread's pageable copies and pickling — several times the
GPU time of the decompositions themselves.  `LayerWriter` keeps the files (same names, same
`torch.load`-able dicts) but takes them off the critical path:

    submit():  records an event on the producing stream and queues the job (no copy, no sync);
    stagers:   threads with their own CUDA stream: wait for the event on that stream and copy each
               tensor device -> host.  Destination is a POOL of pinned buffers allocated once
               (`pool_bytes`, sized from the largest tensor the model can produce): one
               cudaMemcpyAsync per tensor at PCIe rate, no intermediate copy; the buffer travels
               with the job and returns to the pool when the file is written.  Tensors larger than
               a pool buffer (or a job needing more buffers than the pool has) take the older
               route: 64 MB slices through two pinned bounce buffers into pageable memory, which
               tops out near 2 GB/s per thread on first-touch page faults;
    savers:    a small pool of threads that `torch.save` them;
    flush():   drains both queues and re-raises the first error.
Round 2, Llama-2-7B (9.7 GB of layer files): the bounce route sustained 5.9 GB/s end to end and its
back-pressure (`max_in_flight`) throttled the decompositions to 0.051 s/layer against 0.026 in
memory, while 8 raw writer threads reach 26 GB/s on the same file system.

Measured alternatives (Llama-2-7B, 10 GB of layer files): per-layer pinned staging buffers pin
memory at ~1 GB/s and stall every other CUDA call meanwhile (+3 s on the calibration it overlapped);
pageable `tensor.cpu()` from the saver threads slowed the kernel-launching threads 2.5x.  The
`submit` blocks when `max_in_flight` jobs are pending; host memory in flight is bounded by the pinned
pool anyway (a job waits for its pool buffers before it is staged), so the limit is generous (64):
at 16 the short Q/K stage of a 7B run spent 2.4 ms per layer waiting for slots.
"""
from __future__ import annotations

import os
import queue
import threading
import time

import torch
from torch import Tensor

_SLICE = 64 << 20


class LayerWriter:
    def __init__(self, n_threads: int = 8, max_in_flight: int = 64, device=None, n_stagers: int = 3,
                 pool_buffer_bytes: int = 0, pool_bytes: int = 2 << 30):
        # pinned staging pool (CUDA only): `pool_buffer_bytes` = size of one buffer (0 = no pool)
        self._pool_free: list = []
        self._pool_cv = threading.Condition()
        self._pool_buf_bytes = 0
        if device is not None and torch.device(device).type == "cuda" and pool_buffer_bytes > 0:
            count = max(4, min(24, pool_bytes // pool_buffer_bytes))
            try:
                self._pool_free = [torch.empty(pool_buffer_bytes, dtype=torch.uint8, pin_memory=True)
                                   for _ in range(count)]
                self._pool_buf_bytes = pool_buffer_bytes
            except RuntimeError:          # not enough lockable memory: bounce route only
                self._pool_free = []
        self._pool_total = len(self._pool_free)
        self._stage_q: queue.Queue = queue.Queue()
        self._save_q: queue.Queue = queue.Queue()
        self._slots = threading.Semaphore(max_in_flight)
        self._errors: list[BaseException] = []
        self.bytes_written = 0
        # thread-seconds per phase (diagnostics: which side of the pipeline is the limit)
        self.stats = {"submit_wait_s": 0.0, "pool_wait_s": 0.0, "stage_s": 0.0, "save_s": 0.0, "jobs": 0}
        self._stats_lock = threading.Lock()
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
    def _add(self, key: str, dt: float) -> None:
        with self._stats_lock:
            self.stats[key] += dt

    def submit(self, path: str, weights: dict[str, Tensor]) -> None:
        t0 = time.perf_counter()
        self._slots.acquire()                           # blocks while too many results are pending
        self._add("submit_wait_s", time.perf_counter() - t0)
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
        # The stream is created up front; the two 64 MB bounce buffers of the fallback route only
        # when a tensor actually needs them: pinning memory (cudaHostAlloc) stalls every other CUDA
        # call of the process while it runs, and three stager threads pinning 384 MB right after
        # `prepare_writer()` returned cost the FIRST decomposition stage of a run ~10 ms per layer.
        stream, bounce, events = None, None, None
        if self._warm_device is not None and self._warm_device.type == "cuda":
            try:
                torch.cuda.set_device(self._warm_device)   # threads start on device 0: keep rank r on GPU r
                stream = torch.cuda.Stream(device=self._warm_device)
            except BaseException as e:
                self._errors.append(e)
                stream = None
        while True:
            job = self._stage_q.get()
            if job is None:
                self._stage_q.task_done()
                return
            path, weights, ready = job
            held: list = []
            try:
                if ready is not None:
                    dev = next(w.device for w in weights.values() if w.is_cuda)
                    if stream is None or stream.device != dev:
                        torch.cuda.set_device(dev)
                        stream = torch.cuda.Stream(device=dev)
                    with torch.cuda.stream(stream):
                        stream.wait_event(ready)
                        t0 = time.perf_counter()
                        bufs = self._take_buffers(weights)
                        t1 = time.perf_counter()
                        self._add("pool_wait_s", t1 - t0)
                        held = list(bufs.values())
                        staged = {}
                        for k, w in weights.items():
                            if k in bufs:
                                staged[k] = self._to_pinned(w, bufs[k], stream)
                            else:
                                if w.is_cuda and bounce is None:
                                    bounce = [torch.empty(_SLICE, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
                                    events = [torch.cuda.Event() for _ in range(2)]
                                staged[k] = self._to_host(w, stream, bounce, events)
                        weights = staged
                        if held:
                            stream.synchronize()
                        self._add("stage_s", time.perf_counter() - t1)
                self._save_q.put((path, weights, held))
                held = []
            except BaseException as e:                  # surfaced by flush()
                self._errors.append(e)
                self._give_back(held)
                self._slots.release()
            finally:
                del weights, job
                self._stage_q.task_done()

    # ---- pinned pool ---------------------------------------------------------------------------
    def _take_buffers(self, weights: dict) -> dict:
        """One pool buffer per CUDA tensor that fits, taken ATOMICALLY (a job holding part of its
        buffers while waiting for the rest could deadlock against another stager)."""
        if not self._pool_total:
            return {}
        want = [k for k, w in weights.items()
                if w.is_cuda and 0 < w.numel() * w.element_size() <= self._pool_buf_bytes]
        want = want[: self._pool_total]
        if not want:
            return {}
        with self._pool_cv:
            while len(self._pool_free) < len(want):
                self._pool_cv.wait()
            return {k: self._pool_free.pop() for k in want}

    def _give_back(self, bufs: list) -> None:
        if bufs:
            with self._pool_cv:
                self._pool_free.extend(bufs)
                self._pool_cv.notify_all()

    @staticmethod
    def _to_pinned(w: Tensor, buf: Tensor, stream) -> Tensor:
        """One asynchronous device -> pinned copy (the caller synchronises the stream once per
        job).  The tensor handed to `torch.save` aliases the pool buffer through a numpy slice, so
        its STORAGE is exactly the tensor's bytes — a plain view of the pool buffer would make
        torch.save write the whole buffer."""
        transposed = w.dim() == 2 and not w.is_contiguous() and w.T.is_contiguous()
        src = w.T if transposed else w.contiguous()
        nbytes = src.numel() * src.element_size()
        buf[:nbytes].view(src.dtype).view(src.shape).copy_(src, non_blocking=True)
        dst = torch.from_numpy(buf.numpy()[:nbytes]).view(src.dtype).view(src.shape)
        return dst.T if transposed else dst

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
            path, weights, held = job
            try:
                os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
                t0 = time.perf_counter()
                torch.save(weights, path)
                self._add("save_s", time.perf_counter() - t0)
                self._add("jobs", 1)
                self.bytes_written += sum(v.numel() * v.element_size() for v in weights.values())
            except BaseException as e:                  # surfaced by flush()
                self._errors.append(e)
            finally:
                del weights, job
                self._give_back(held)
                self._slots.release()
                self._save_q.task_done()
