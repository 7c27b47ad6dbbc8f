"""Per-launch timeline of one mg_ridge_scores_f32 + one mg_nystrom_down_f32 at Llama-2-7B MLP size
(MG_PROFILE=1 + MG_PROFILE_TIMELINE): every launch with its lane, start and end.  The in-situ events
cost a few percent; the point is the STRUCTURE — which lane the call is waiting for.
    python tools/gpu_timeline_type1.py gpurun_out/timeline.csv ; python tools/timeline_summary.py ..."""
import os
import sys

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/type1_timeline.csv"
if os.path.exists(out):
    os.remove(out)
os.environ["MG_PROFILE"] = "1"
os.environ["MG_PROFILE_TIMELINE"] = out
sys.path.insert(0, ".")
import torch  # noqa: E402

from modegpt_b200 import ops  # noqa: E402

n, d, T = 11008, 4096, 16384
torch.manual_seed(0)
x = (torch.randn(T, n, device="cuda") * torch.exp(0.5 * torch.randn(n, device="cuda"))).bfloat16()
c = torch.zeros(n, n, device="cuda")
ops.syrk_(c, x)
ops.finalize_sym_(c, 1.0 / T)
del x
wd = (torch.randn(d, n, device="cuda") * 0.02).bfloat16()
for rep in range(3):   # the last repetition is the one to read
    with open(out, "a") as f:
        f.write(f"#rep {rep}\n")
    s = ops.ridge_scores(c, 1e-4)
    idx = ops.select_k(s, int(n * 0.75))
    w = ops.nystrom_down(c, idx, wd)
    torch.cuda.synchronize()
print("timeline written to", out)
