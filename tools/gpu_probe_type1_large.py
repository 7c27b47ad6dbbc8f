"""Type-I at Llama-2-70B MLP width (n = 28672, d = 8192, keep 0.7): timing + in-situ profile
(run with MG_PROFILE=1 for the per-kernel breakdown on stderr)."""
import sys
import time
import torch
sys.path.insert(0, ".")
from modegpt_b200 import ops

n, d, T = 28672, 8192, 32768
torch.manual_seed(0)
c = torch.zeros(n, n, device="cuda")
for _ in range(2):
    x = (torch.randn(T // 2, n, device="cuda") * torch.exp(0.5 * torch.randn(n, device="cuda"))).bfloat16()
    ops.syrk_(c, x)
    del x
ops.finalize_sym_(c, 1.0 / T)
wd = (torch.randn(d, n, device="cuda") * 0.02).bfloat16()
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    s = ops.ridge_scores(c, 1e-4)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    idx = ops.select_k(s, int(n * 0.7))
    out = ops.nystrom_down(c, idx, wd)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"rep {rep}: ridge {1e3 * (t1 - t0):.1f} ms, select+nystrom {1e3 * (t2 - t1):.1f} ms", flush=True)
