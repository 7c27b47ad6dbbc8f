"""SYRK rasterisation probe: time the C_mlp SYRK (n = 11008, T = 32768) for several MG_SYRK_BAND
values (0 = column-major order).  The band is read once per process, so each value runs in a child."""
import os
import subprocess
import sys

CHILD = r"""
import sys, torch
sys.path.insert(0, ".")
from modegpt_b200 import ops
n, T = int(sys.argv[1]), int(sys.argv[2])
x = torch.randn(T, n, device="cuda").bfloat16()
c = torch.zeros(n, n, device="cuda")
for _ in range(3):
    ops.syrk_(c, x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.syrk_(c, x)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print(f"ms={ms:.4f} tflops={T * n * (n + 1) / ms / 1e9:.1f} checksum={float(c.double().sum()):.6e}")
"""

if __name__ == "__main__":
    for n, T in ((11008, 32768), (4096, 32768), (14336, 16384)):
        for band in (0, 4, 6, 8, 12, 16):
            env = dict(os.environ, MG_SYRK_BAND=str(band))
            r = subprocess.run([sys.executable, "-c", CHILD, str(n), str(T)], env=env,
                               capture_output=True, text=True)
            print(f"n={n} T={T} band={band}: {r.stdout.strip()} {r.stderr.strip()[-200:]}", flush=True)
