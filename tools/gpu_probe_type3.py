"""Type-III timing at the 7B shapes: prepare (per method) / finish split, one layer and groups.
MG_PROF=1 adds the in-library per-kernel profile of the Cholesky lanes."""
import sys, time, torch
sys.path.insert(0, ".")
from modegpt_b200 import ops
dev = "cuda:0"
torch.manual_seed(0)


def ev_ms(fn, iters=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


for (d, H, KV, hd, r) in [(4096, 32, 32, 128, 96), (4096, 32, 8, 128, 96), (3584, 28, 4, 128, 88),
                          (8192, 64, 8, 128, 88)]:
    T = 2 * d
    x = (torch.randn(T, d, device=dev) * torch.exp(0.5 * torch.randn(d, device=dev))).bfloat16()
    cx = torch.zeros(d, d, device=dev); ops.syrk_(cx, x); ops.finalize_sym_(cx, 1.0 / T)
    wv = (torch.randn(KV * hd, d, device=dev) * 0.02).bfloat16()
    wo = (torch.randn(d, H * hd, device=dev) * 0.02).bfloat16()
    ws, info = ops.vo_prepare(cx, 1e-5, wv, wo, H, KV, hd)
    out = ops.vo_outputs(wv, H, KV, r)
    for m, name in ((ops.VO_FACTOR, "factor"), (ops.VO_GRAM, "gram")):
        t = ev_ms(lambda: ops.vo_prepare_into(ws, info, cx, 1e-5, wv, wo, H, KV, hd, m))
        print(f"d={d} H={H} KV={KV}: prepare[{name}] {t:.3f} ms  info={int(info.item())}", flush=True)
    ops.vo_prepare_into(ws, info, cx, 1e-5, wv, wo, H, KV, hd, ops.VO_FACTOR)
    t = ev_ms(lambda: ops.vo_finish(ws, wv, wo, H, KV, hd, r, out=out))
    print(f"d={d} H={H} KV={KV}: finish {t:.3f} ms", flush=True)
    t = ev_ms(lambda: ops.vo_compress(cx, 1e-5, wv, wo, H, KV, hd, r, method=ops.VO_FACTOR))
    print(f"d={d} H={H} KV={KV}: vo_compress (incl. info read-back) {t:.3f} ms", flush=True)
    del x, cx, ws
