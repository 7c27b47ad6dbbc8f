"""Small driver for ncu: one ridge_scores + one nystrom_down at Llama-2-7B MLP size."""
import sys
import torch
sys.path.insert(0, ".")
from modegpt_b200 import ops

dev = "cuda:0"
n, d = 11008, 4096
torch.manual_seed(0)
x = torch.randn(16384, n, device=dev).bfloat16()
c = torch.zeros(n, n, device=dev)
ops.syrk_(c, x)
ops.finalize_sym_(c, 1.0 / 16384)
del x
s = ops.ridge_scores(c, 1e-4)
idx = ops.select_k(s, int(n * 0.75))
wd = (torch.randn(d, n, device=dev) * 0.02).bfloat16()
out = ops.nystrom_down(c, idx, wd)
torch.cuda.synchronize()
print("ok", float(s.sum()), float(out.float().abs().sum()))
