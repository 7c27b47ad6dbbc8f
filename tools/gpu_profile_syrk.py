"""Small ncu driver: the C_mlp SYRK of the bench (n = 11008, T = 32768), 3 launches."""
import sys
import torch
sys.path.insert(0, ".")
from modegpt_b200 import ops
x = torch.randn(32768, 11008, device="cuda").bfloat16()
c = torch.zeros(11008, 11008, device="cuda")
for _ in range(3):
    ops.syrk_(c, x)
torch.cuda.synchronize()
print("ok", float(c[0, 0]))
