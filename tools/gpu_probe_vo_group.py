"""Do the eigensolver halves of several layers overlap when issued on different streams?"""
import sys, time, torch
sys.path.insert(0, ".")
from modegpt_b200 import ops
dev = "cuda:0"
d, H, KV, hd, r = 4096, 32, 32, 128, 96
torch.manual_seed(0)
x = (torch.randn(16384, d, device=dev) * torch.exp(0.5 * torch.randn(d, device=dev))).bfloat16()
cx = torch.zeros(d, d, device=dev); ops.syrk_(cx, x); ops.finalize_sym_(cx, 1.0 / 16384)
wv = [(torch.randn(KV * hd, d, device=dev) * 0.02).bfloat16() for _ in range(4)]
wo = [(torch.randn(d, H * hd, device=dev) * 0.02).bfloat16() for _ in range(4)]
streams = [torch.cuda.Stream() for _ in range(4)]
def run(parallel):
    ws = [ops.vo_prepare(cx, 1e-5, wv[i], wo[i], H, KV, hd)[0] for i in range(4)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(4):
        if parallel:
            with torch.cuda.stream(streams[i]):
                ops.vo_finish(ws[i], wv[i], wo[i], H, KV, hd, r)
        else:
            ops.vo_finish(ws[i], wv[i], wo[i], H, KV, hd, r)
    torch.cuda.synchronize(); return 1e3 * (time.perf_counter() - t0)
for _ in range(2):
    print(f"sequential {run(False):.2f} ms   4 streams {run(True):.2f} ms", flush=True)
