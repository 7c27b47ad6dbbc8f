"""Is type-I host-bound?  Time the C call (enqueue only) vs the GPU completion."""
import sys, time
import torch
sys.path.insert(0, ".")
from modegpt_b200 import ops
from modegpt_b200._lib import lib

dev = "cuda:0"
n = 11008
x = torch.randn(16384, n, device=dev).bfloat16()
c = torch.zeros(n, n, device=dev)
ops.syrk_(c, x); ops.finalize_sym_(c, 1.0 / 16384); del x
scores = torch.empty(n, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
nbytes = lib.mg_ridge_scores_ws_bytes(n)
ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for it in range(3):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    rc = lib.mg_ridge_scores_f32(c.data_ptr(), n, c.stride(0), 1e-4, scores.data_ptr(), ws.data_ptr(), nbytes, info.data_ptr(), st)
    e.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"rc={rc} host enqueue {1e3*(t1-t0):.2f} ms, until done {1e3*(t2-t0):.2f} ms, gpu events {s.elapsed_time(e):.2f} ms", flush=True)
