"""Stress the type-I drivers for timing-dependent results: the same 8 layers decomposed again and
again — with 1 or 2 worker threads, optionally with a background stream hammering the GPU — must
reproduce the first single-worker result exactly (selected indices) / to rounding (W_down)."""
import os, sys, threading, time, torch
sys.path.insert(0, ".")
from modegpt_b200 import ops
from modegpt_b200.adapters.CompressionConfig import CompressionConfig
from modegpt_b200.adapters.model_adapter import ModelAdapter
from modegpt_b200.compression.compress_mlp import compress_nystrom
from modegpt_b200.model_utils import build_synthetic_model

dev = "cuda:0"
L = 8
iters = int(os.environ.get("ITERS", "12"))
model = build_synthetic_model("llama-2-7b", device=dev, n_layers=L)
adapter = ModelAdapter.from_model(model, None)
torch.manual_seed(0)
n, T = 11008, 16384
x = (torch.randn(T, n, device=dev) * torch.exp(0.5 * torch.randn(n, device=dev))).bfloat16()
c = torch.zeros(n, n, device=dev); ops.syrk_(c, x); ops.finalize_sym_(c, 1.0 / T)
del x
cov, keep = [c] * L, [0.75] * L


def run(workers):
    adapter.config = CompressionConfig(model="x", order="mlp", nystrom_ridge=1e-4, keep_layers_in_memory=True,
                                       mlp_workers=workers)
    adapter._layer_store = {}
    compress_nystrom(adapter, cov, keep, list(range(L)))
    torch.cuda.synchronize()
    return {l: adapter._layer_store[(l, "mlp")] for l in range(L)}


ref = run(1)
stop = [False]


def noise():
    torch.cuda.set_device(dev)
    s = torch.cuda.Stream()
    a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    with torch.cuda.stream(s):
        while not stop[0]:
            for _ in range(4):
                a @ a
            s.synchronize()
            time.sleep(0.0005)


for mode, workers, noisy in (("1 worker", 1, False), ("1 worker + background GEMMs", 1, True),
                             ("2 workers", 2, False), ("2 workers + background GEMMs", 2, True)):
    stop[0] = False
    th = threading.Thread(target=noise) if noisy else None
    if th:
        th.start()
    bad = 0
    worst = 0.0
    for it in range(iters):
        out = run(workers)
        for l in range(L):
            same_rows = torch.equal(out[l]["up"], ref[l]["up"])
            d = ((out[l]["down"].float() - ref[l]["down"].float()).norm() / ref[l]["down"].float().norm()).item()
            worst = max(worst, d)
            if not same_rows or d > 1e-3:
                bad += 1
                print(f"  MISMATCH {mode}: iter {it} layer {l} rows_equal={same_rows} down rel diff {d:.3e}", flush=True)
    stop[0] = True
    if th:
        th.join()
    print(f"{mode}: {iters * L} layer decompositions, {bad} mismatches, worst W_down rel diff {worst:.2e}", flush=True)
