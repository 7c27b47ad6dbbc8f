"""Small ncu driver: type-II + type-III at Llama-2-7B attention shape (d = 4096, 32 MHA heads of
128, rank 96), and the GQA variant (8 kv heads)."""
import sys
import time
import torch
sys.path.insert(0, ".")
from modegpt_b200 import ops

dev = "cuda:0"
d, H, hd, r = 4096, 32, 128, 96
torch.manual_seed(0)
x = (torch.randn(16384, d, device=dev) * torch.exp(0.5 * torch.randn(d, device=dev))).bfloat16()
cx = torch.zeros(d, d, device=dev)
ops.syrk_(cx, x)
ops.finalize_sym_(cx, 1.0 / 16384)
for KV in (32, 8):
    wv = (torch.randn(KV * hd, d, device=dev) * 0.02).bfloat16()
    wo = (torch.randn(d, H * hd, device=dev) * 0.02).bfloat16()
    v, o = ops.vo_compress(cx, 1e-5, wv, wo, H, KV, hd, r)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    v, o = ops.vo_compress(cx, 1e-5, wv, wo, H, KV, hd, r)
    torch.cuda.synchronize()
    print(f"KV={KV}: vo_compress {1e3 * (time.perf_counter() - t0):.2f} ms", float(v.float().abs().sum()))
