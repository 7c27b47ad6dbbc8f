"""Host-side probe: how fast can this box write layer files?  torch.save (zip / legacy format) vs raw
writes, 1..16 threads, from pageable and pinned memory."""
import os, sys, tempfile, threading, time
import torch
torch.serialization.set_crc32_options(False)
d = tempfile.mkdtemp(prefix="mg_probe_")
t = torch.randn(8256, 4096).bfloat16()
tp = t.pin_memory() if torch.cuda.is_available() else t
nb = t.numel() * 2


def run(fn, n):
    ths = [threading.Thread(target=fn, args=(i,)) for i in range(n)]
    t0 = time.perf_counter()
    [x.start() for x in ths]; [x.join() for x in ths]
    dt = time.perf_counter() - t0
    for i in range(n):
        p = os.path.join(d, f"f{i}")
        if os.path.exists(p):
            os.remove(p)
    return n * nb / dt / 1e9


for name, src in (("pageable", t), ("pinned", tp)):
    for n in (1, 4, 8, 16):
        z = run(lambda i: torch.save({"a": src}, os.path.join(d, f"f{i}")), n)
        l = run(lambda i: torch.save({"a": src}, os.path.join(d, f"f{i}"), _use_new_zipfile_serialization=False), n)
        def raw(i):
            with open(os.path.join(d, f"f{i}"), "wb") as f:
                f.write(memoryview(src.view(torch.uint8).numpy()).cast("B"))
        r = run(raw, n)
        print(f"{name:9s} threads={n:2d}: torch.save zip {z:6.2f} GB/s   legacy {l:6.2f} GB/s   raw {r:6.2f} GB/s", flush=True)
print("cores", os.cpu_count())
