"""GPU probe for the statistics kernels: correctness vs torch fp64 and first timings.

Run on the GPU box:  timeout 600 python tools/gpu_probe_stats.py
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from modegpt_b200 import ops  # noqa: E402

torch.manual_seed(0)
dev = "cuda:0"


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def check_syrk(T, n, accumulate_twice=False):
    X = (torch.randn(T, n, device=dev) * torch.exp(0.5 * torch.randn(n, device=dev))).bfloat16()
    C = torch.zeros(n, n, device=dev, dtype=torch.float32)
    ops.syrk_(C, X)
    if accumulate_twice:
        ops.syrk_(C, X)
    torch.cuda.synchronize()
    ref = X.double().T @ X.double()
    if accumulate_twice:
        ref = ref * 2
    up = torch.triu(torch.ones(n, n, device=dev, dtype=torch.bool))
    e = rel(C.double()[up], ref[up])
    low = C[~up].abs().max().item() if n > 1 else 0.0
    print(f"syrk T={T} n={n} twice={accumulate_twice}: rel_fro_upper={e:.3e} lower_max={low:.3e}",
          flush=True)
    if e > 1e-3:
        # diagnose: which 64x64 blocks are right?
        nb = min(n // 64, 8)
        for bi in range(nb):
            row = []
            for bj in range(nb):
                a = C.double()[bi * 64:(bi + 1) * 64, bj * 64:(bj + 1) * 64]
                b = ref[bi * 64:(bi + 1) * 64, bj * 64:(bj + 1) * 64]
                row.append(f"{rel(a, b):8.1e}")
            print("   ", " ".join(row))
        print("    C[0,:8]  ", C[0, :8].tolist())
        print("    ref[0,:8]", ref[0, :8].float().tolist())
    return e


def check_heads(T, H, hd):
    n = H * hd
    X = torch.randn(T, n, device=dev).bfloat16()
    C = torch.zeros(H, hd, hd, device=dev, dtype=torch.float32)
    ops.syrk_heads_(C, X)
    torch.cuda.synchronize()
    Xh = X.double().view(T, H, hd).permute(1, 0, 2)
    ref = torch.bmm(Xh.transpose(1, 2), Xh)
    e = rel(C.double(), ref)
    print(f"heads T={T} H={H} hd={hd}: rel_fro={e:.3e}", flush=True)
    return e


def check_bi(rows, d):
    a = torch.randn(rows, d, device=dev).bfloat16()
    b = (a.float() + 0.3 * torch.randn(rows, d, device=dev)).bfloat16()
    acc = torch.zeros(1, device=dev, dtype=torch.float64)
    ops.bi_cosine_(acc, a, b)
    torch.cuda.synchronize()
    ref = (1 - torch.cosine_similarity(a.double(), b.double(), dim=1)).sum().item()
    print(f"bi rows={rows} d={d}: ours={acc.item():.12f} ref={ref:.12f} "
          f"rel={abs(acc.item() - ref) / abs(ref):.3e}", flush=True)


def check_finalize(n):
    C = torch.randn(n, n, device=dev)
    ref = torch.triu(C) * 0.5
    ref = ref + torch.triu(ref, 1).T
    ops.finalize_sym_(C, 0.5)
    torch.cuda.synchronize()
    print(f"finalize n={n}: max_abs_err={(C - ref).abs().max().item():.3e}", flush=True)


def bench_syrk(T, n, iters=5):
    X = torch.randn(T, n, device=dev).bfloat16()
    C = torch.zeros(n, n, device=dev, dtype=torch.float32)
    for _ in range(2):
        ops.syrk_(C, X)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        ops.syrk_(C, X)
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    flops = T * n * (n + 1)
    print(f"bench syrk T={T} n={n}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s (upper-tri flops)",
          flush=True)
    # torch bf16 full GEMM for scale
    Y = torch.empty(n, n, device=dev, dtype=torch.bfloat16)
    for _ in range(2):
        torch.matmul(X.T, X, out=Y)
    torch.cuda.synchronize()
    s.record()
    for _ in range(iters):
        torch.matmul(X.T, X, out=Y)
    e.record()
    torch.cuda.synchronize()
    ms2 = s.elapsed_time(e) / iters
    print(f"      torch bf16 X^T X (full, bf16 out): {ms2:.3f} ms  {2 * T * n * n / ms2 / 1e9:.1f} "
          f"TFLOP/s", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    t0 = time.time()
    check_syrk(64, 256)
    check_syrk(128, 256)
    check_syrk(1000, 512)
    check_syrk(4096, 1024, accumulate_twice=True)
    check_syrk(300, 328)       # ragged n (multiple of 8)
    check_syrk(2048, 4096)
    check_heads(4096, 32, 128)
    check_heads(1000, 12, 64)
    check_heads(512, 8, 32)
    check_bi(4096, 4096)
    check_bi(1000, 768)
    check_finalize(1000)
    bench_syrk(8192, 4096)
    bench_syrk(32768, 4096)
    bench_syrk(8192, 11008)
    bench_syrk(32768, 11008)
    print(f"done in {time.time() - t0:.1f}s", flush=True)
