"""Achieved HBM GB/s of the memory-bound kernels at Llama-2-7B shapes (north_star: "achieved HBM GB/s
against peak for the score/select/gather kernels").

Each op is timed alone with CUDA events, L2 flushed (a 512 MB write) before every timed launch,
median of 7; achieved = ALGORITHMIC bytes (each operand element read once, each result element
written once) / time; peak = MEASURED_PEAKS.json hbm_gbs (6539.5 on this pool).  Kernels that are
only reachable inside a driver (gather_sym, gather_cols_planes, split_planes, copy_ridge,
vo_apply_*) are listed from the ncu launch list instead (tools/summarize_ncu.py hbm ...).

    python tools/gpu_hbm_kernels.py [out.txt]
"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, ".")
from modegpt_b200 import ops  # noqa: E402

dev = "cuda:0"
ROOT = Path(__file__).resolve().parent.parent
peak = 6539.5
p = ROOT / "MEASURED_PEAKS.json"
if p.exists():
    peak = float(json.loads(p.read_text())["hbm_gbs"])
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=7):
    fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


rows = []


def report(name, kernel, nbytes, ms):
    gbs = nbytes / (ms * 1e-3) / 1e9
    rows.append((name, kernel, nbytes / 1e6, ms * 1e3, gbs, gbs / peak))


torch.manual_seed(0)
n, d, T, H, hd, r = 11008, 4096, 32768, 32, 128, 96
k = int(n * 0.75)
c = torch.randn(n, n, device=dev)
cx = torch.randn(d, d, device=dev)
report("finalize_sym C_mlp (n=11008): r+w upper, w lower", "finalize_sym_kernel",
       (n * (n + 1) // 2 * 2 + n * (n - 1) // 2) * 4, timed(lambda: ops.finalize_sym_(c, 1.0)))
report("finalize_sym C_x (n=4096)", "finalize_sym_kernel",
       (d * (d + 1) // 2 * 2 + d * (d - 1) // 2) * 4, timed(lambda: ops.finalize_sym_(cx, 1.0)))
cq = torch.randn(H, hd, hd, device=dev)
report("scale C_q [32,128,128]", "scale_kernel", cq.numel() * 8, timed(lambda: ops.scale_(cq, 1.0)))
packed = torch.empty(ops.packed_upper_numel(n), device=dev)
report("pack_upper C_mlp", "pack_upper_kernel<true>", packed.numel() * 8, timed(lambda: ops.pack_upper_(packed, c)))
report("unpack_upper C_mlp", "pack_upper_kernel<false>", packed.numel() * 8, timed(lambda: ops.unpack_upper_(c, packed)))
x_in = torch.randn(T, d, device=dev).bfloat16()
x_out = torch.randn(T, d, device=dev).bfloat16()
acc = torch.zeros(1, dtype=torch.float64, device=dev)
report("bi_cosine [32768, 4096] x 2", "bi_cosine_kernel", 2 * T * d * 2, timed(lambda: ops.bi_cosine_(acc, x_in, x_out)))
w_up = torch.randn(n, d, device=dev).bfloat16()
scores = torch.rand(n, device=dev)
idx = ops.select_k(scores, k)
report("select_k n=11008 k=8256 (one CTA, latency-bound)", "select_k_kernel", n * 4 + k * 8,
       timed(lambda: ops.select_k(scores, k)))
report("gather_rows W_up[idx] (8256 x 4096 bf16)", "gather_rows_kernel", 2 * k * d * 2,
       timed(lambda: ops.gather_rows(w_up, idx)))
wq = torch.randn(H * hd, d, device=dev).bfloat16()
mask = torch.stack([torch.randperm(hd, device=dev)[:r] for _ in range(H)]).contiguous()
report("gather_head_rows W_q (32 heads x 96 x 4096)", "gather_rows_kernel", 2 * H * r * d * 2,
       timed(lambda: ops.gather_head_rows(wq, mask, H, 1, hd)))
w = torch.ones(d, device=dev).bfloat16()
report("rmsnorm [32768, 4096]", "rmsnorm_kernel", 2 * T * d * 2, timed(lambda: ops.rmsnorm(x_in, w, 1e-5)))
gate = torch.randn(T, n, device=dev).bfloat16()
up = torch.randn(T, n, device=dev).bfloat16()
report("swiglu [32768, 11008]", "swiglu_kernel", 3 * T * n * 2, timed(lambda: ops.swiglu(gate, up)))
del gate, up
xq = torch.randn(16, 2048, H, hd, device=dev).bfloat16()
cos = torch.randn(1, 2048, hd, device=dev).bfloat16()
sin = torch.randn(1, 2048, hd, device=dev).bfloat16()
report("rope [16, 2048, 32, 128]", "rope_kernel", 2 * xq.numel() * 2, timed(lambda: ops.rope_bthd(xq, cos, sin)))

lines = [f"# achieved HBM bandwidth of the memory-bound kernels, Llama-2-7B shapes; peak = {peak} GB/s "
         f"(MEASURED_PEAKS.json hbm_gbs); L2 flushed before each timed launch; median of 7",
         f"{'op':58s} {'kernel':26s} {'MB (algorithmic)':>17s} {'us':>9s} {'GB/s':>8s} {'frac':>6s}"]
for name, kernel, mb, us, gbs, frac in rows:
    lines.append(f"{name:58s} {kernel:26s} {mb:17.1f} {us:9.1f} {gbs:8.0f} {frac:6.2f}")
text = "\n".join(lines)
print(text)
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text(text + "\n")
