"""compress_nystrom over 8 Llama-2-7B-shaped layers with 1 / 2 / 3 worker threads (in-memory)."""
import sys, time, torch
sys.path.insert(0, ".")
from modegpt_b200 import ops
from modegpt_b200.adapters.CompressionConfig import CompressionConfig
from modegpt_b200.adapters.model_adapter import ModelAdapter
from modegpt_b200.compression.compress_mlp import compress_nystrom
from modegpt_b200.model_utils import build_synthetic_model

dev = "cuda:0"
L = 8
model = build_synthetic_model("llama-2-7b", device=dev, n_layers=L)
adapter = ModelAdapter.from_model(model, None)
torch.manual_seed(0)
n, T = 11008, 16384
x = (torch.randn(T, n, device=dev) * torch.exp(0.5 * torch.randn(n, device=dev))).bfloat16()
c = torch.zeros(n, n, device=dev); ops.syrk_(c, x); ops.finalize_sym_(c, 1.0 / T)
del x
cov = [c] * L
keep = [0.75] * L
ref = None
import os
for workers in tuple(int(x) for x in os.environ.get('WORKERS', '1,2,3,1,2').split(',')):
    adapter.config = CompressionConfig(model="x", order="mlp", nystrom_ridge=1e-4, keep_layers_in_memory=True,
                                       mlp_workers=workers)
    adapter._layer_store = {}
    compress_nystrom(adapter, cov, keep, list(range(2)))      # warm
    torch.cuda.synchronize(); t0 = time.perf_counter()
    adapter._layer_store = {}
    compress_nystrom(adapter, cov, keep, list(range(L)))
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    down = adapter._layer_store[(L - 1, "mlp")]["down"].float()
    if ref is None:
        ref = down
    print(f"workers={workers}: {1e3 * dt / L:.2f} ms/layer   rel diff vs workers=1: "
          f"{((down - ref).norm() / ref.norm()).item():.2e}", flush=True)
