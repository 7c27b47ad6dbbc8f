"""Type-I timing probe at Llama-2-7B MLP size: ridge_scores + nystrom_down with the lanes on and off
(MG_SERIAL=1), results compared between the two modes.  MG_PROFILE=1 in the environment adds the
per-kernel in-situ breakdown (stderr)."""
import os
import subprocess
import sys

CHILD = r"""
import sys, time, torch
sys.path.insert(0, ".")
from modegpt_b200 import ops
n, d, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
torch.manual_seed(0)
x = (torch.randn(T, n, device="cuda") * torch.exp(0.5 * torch.randn(n, device="cuda"))).bfloat16()
c = torch.zeros(n, n, device="cuda")
ops.syrk_(c, x); ops.finalize_sym_(c, 1.0 / T); del x
wd = (torch.randn(d, n, device="cuda") * 0.02).bfloat16()
def run():
    s = ops.ridge_scores(c, 1e-4)
    idx = ops.select_k(s, int(n * 0.75))
    out = ops.nystrom_down(c, idx, wd)
    return s, idx, out
run(); torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
reps = int(sys.argv[4])
tr = tn = 0.0
for _ in range(reps):
    ev[0].record(); s = ops.ridge_scores(c, 1e-4); ev[1].record()
    idx = ops.select_k(s, int(n * 0.75))
    e2 = torch.cuda.Event(enable_timing=True); e2.record()
    out = ops.nystrom_down(c, idx, wd); ev[2].record()
    torch.cuda.synchronize()
    tr += ev[0].elapsed_time(ev[1]); tn += e2.elapsed_time(ev[2])
print(f"ridge_ms={tr/reps:.2f} nystrom_ms={tn/reps:.2f}")
# host side: how long the library call itself takes to enqueue everything (returns before the GPU is done)
from modegpt_b200._lib import lib
import modegpt_b200.ops as _ops
acc = {"mg_ridge_scores_f32": 0.0, "mg_nystrom_down_f32": 0.0}
for name in acc:
    real = getattr(lib, name)
    def timed(*a, _real=real, _name=name):
        t0 = time.perf_counter(); rc = _real(*a); acc[_name] += time.perf_counter() - t0; return rc
    setattr(_ops.lib, name, timed)
torch.cuda.synchronize()
for _ in range(reps):
    run(); torch.cuda.synchronize()
print("host_enqueue_ms " + " ".join(f"{k}={1e3*v/reps:.2f}" for k, v in acc.items()))
torch.save({"s": s.cpu(), "idx": idx.cpu(), "out": out.cpu()}, sys.argv[5])
"""

if __name__ == "__main__":
    import torch
    n, d, T = 11008, 4096, 16384
    outs = {}
    for serial in ("1", "0"):
        env = dict(os.environ, MG_SERIAL=serial)
        path = f"/tmp/type1_serial{serial}.pt"
        r = subprocess.run([sys.executable, "-c", CHILD, str(n), str(d), str(T), "5", path], env=env,
                           capture_output=True, text=True)
        print(f"MG_SERIAL={serial}: {r.stdout.strip()}", flush=True)
        if r.returncode != 0 or os.environ.get("MG_PROFILE") == "1":
            print(r.stderr[-6000:], flush=True)
        if os.path.exists(path):
            outs[serial] = torch.load(path)
    if len(outs) == 2:
        a, b = outs["1"], outs["0"]
        ds = ((a["s"].double() - b["s"].double()).norm() / a["s"].double().norm()).item()
        same = bool((a["idx"] == b["idx"]).all())
        do = ((a["out"].double() - b["out"].double()).norm() / a["out"].double().norm()).item()
        print(f"serial vs lanes: scores rel diff {ds:.3e}, indices identical {same}, W_down rel diff {do:.3e}")
