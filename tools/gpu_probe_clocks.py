import sys, ctypes
import torch
sys.path.insert(0, ".")
from modegpt_b200 import ops
from modegpt_b200._lib import LIB_PATH
raw = ctypes.CDLL(str(LIB_PATH))
dev = "cuda:0"
n = 2048
x = torch.randn(8192, n, device=dev).bfloat16()
c = torch.zeros(n, n, device=dev)
ops.syrk_(c, x); ops.finalize_sym_(c, 1.0 / 8192)
for _ in range(2):
    s = ops.ridge_scores(c, 1e-4)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 64)()
print("rc", raw.mg_debug_clocks(buf))
v = list(buf)
print("potrf128: global load", v[0]-v[21], " load->sync", v[1]-v[0])
# stamps of thread 0 (warp 0 = the chain): end of the diagonal block, own block-row solve + wait for
# warps 1..3, next diagonal block updated (the last sub-panel has neither)
for kb in range(4):
    line = f"  kb{kb}: diag {v[2+3*kb]-(v[1] if kb==0 else v[4+3*(kb-1)])}"
    if kb < 3:
        line += f"  solve {v[3+3*kb]-v[2+3*kb]}  next-diag update {v[4+3*kb]-v[3+3*kb]}"
    print(line)
print("  outputs fwd", v[19]-v[11], " bwd", v[20]-v[19], " total", v[20]-v[0])
print("trsm128: Tload", v[33]-v[32])
for rb in range(4):
    prev = v[33] if rb == 0 else v[36+4*(rb-1)]
    print(f"  rb{rb}: Bload+out_prev {v[34+4*rb]-prev}  offdiag {v[35+4*rb]-v[34+4*rb]}  diag {v[36+4*rb]-v[35+4*rb]}")
print("  last outputs", v[50]-v[48], " total", v[50]-v[32])
