"""GPU probe for the decomposition kernels (type I/II/III) against fp64 torch / the CPU oracle.

Run on the GPU box:  timeout 900 python tools/gpu_probe_decomp.py [--big]
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from modegpt_b200 import ops  # noqa: E402
from oracle import modegpt_oracle as O  # noqa: E402

torch.manual_seed(0)
dev = "cuda:0"


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-300)).item()


def shaped_c(n, t=None, spread=1.0, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    t = t or 4 * n
    x = torch.randn(t, n, device=dev, generator=g, dtype=torch.float64)
    x = x * torch.exp(spread * torch.randn(n, device=dev, generator=g, dtype=torch.float64))
    c = (x.T @ x) / t
    return c


def timed(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        out = fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters, out


def check_scores(n, ridge, spread=1.0):
    c64 = shaped_c(n, spread=spread)
    c32 = c64.float()
    s = ops.ridge_scores(c32, ridge)
    ref = torch.linalg.inv(c32.double() + ridge * torch.eye(n, device=dev, dtype=torch.float64)).diagonal()
    e = rel(s.double(), ref)
    k = int(n * 0.75)
    idx = ops.select_k(s, k)
    ref_idx = torch.sort(torch.topk(ref, k, largest=False).indices).values
    same = torch.equal(idx, ref_idx)
    inter = len(set(idx.tolist()) & set(ref_idx.tolist()))
    print(f"scores n={n} ridge={ridge} spread={spread}: rel={e:.3e} topk_exact={same} overlap={inter}/{k}",
          flush=True)
    # select on the fp64 reference scores must be exact
    idx2 = ops.select_k(ref.float(), k)
    ref2 = torch.sort(torch.topk(ref.float(), k, largest=False).indices).values
    print(f"   select_k exact on same fp32 scores: {torch.equal(idx2, ref2)}  "
          f"largest: {torch.equal(ops.select_k(ref.float(), 7, True), torch.sort(torch.topk(ref.float(), 7).indices).values)}",
          flush=True)
    return c32, ref_idx


def check_nystrom(n, d, ridge=1e-2, spread=1.0):
    c32, idx = check_scores(n, ridge, spread)
    # full symmetric fp32 C
    wd = (torch.randn(d, n, device=dev) * 0.05).bfloat16()
    out = ops.nystrom_down(c32, idx, wd)
    c = c32.double()
    k = idx.numel()
    ckk = c[idx][:, idx] + 1e-6 * torch.eye(k, device=dev, dtype=torch.float64)
    cross = c[idx, :] @ wd.double().T
    ref = torch.cholesky_solve(cross, torch.linalg.cholesky(ckk)).T
    print(f"nystrom n={n} d={d} k={k}: rel(bf16 out vs fp64)={rel(out.double(), ref):.3e} "
          f"(bf16 rounding alone: {rel(ref.bfloat16().double(), ref):.3e})", flush=True)
    wu = torch.randn(n, d, device=dev).bfloat16()
    print(f"   gather_rows exact: {torch.equal(ops.gather_rows(wu, idx), wu[idx])}", flush=True)


def check_qk():
    rng = np.random.default_rng(0)
    for (H, KV, hd, r, mode, arch) in [(8, 2, 128, 96, 0, "llama"), (4, 4, 64, 40, 0, "llama"),
                                        (12, 12, 64, 44, 1, "opt"), (6, 2, 32, 20, 0, "qwen3")]:
        cq = np.stack([np.array(shaped_c(hd, 300, seed=i).cpu()) for i in range(H)])
        ck = np.stack([np.array(shaped_c(hd, 300, seed=100 + i).cpu()) for i in range(KV)])
        d = 256
        wq = torch.randn(H * hd, d, device=dev).bfloat16()
        wk = torch.randn(KV * hd, d, device=dev).bfloat16()
        ridge_k = 1e-2 if (H != KV) else 1e-4
        mask = ops.qk_select(torch.tensor(cq, device=dev, dtype=torch.float32),
                             torch.tensor(ck, device=dev, dtype=torch.float32), r, mode, 1e-4, ridge_k)
        out, ref_mask = O.qk_layer(wq.float().cpu().numpy(), wk.float().cpu().numpy(), cq, ck, H, KV,
                                   hd, r, arch, 1e-2)
        qn = ops.gather_head_rows(wq, mask, H, H // KV, hd)
        kn = ops.gather_head_rows(wk, mask, KV, 1, hd)
        ok_m = np.array_equal(mask.cpu().numpy(), ref_mask)
        ok_q = np.array_equal(qn.float().cpu().numpy(), out["q_proj"])
        ok_k = np.array_equal(kn.float().cpu().numpy(), out["k_proj"])
        print(f"qk H={H} KV={KV} hd={hd} r={r} mode={mode}: mask={ok_m} q={ok_q} k={ok_k}", flush=True)


def check_vo():
    for (d, H, KV, hd, r) in [(256, 4, 4, 64, 40), (512, 8, 2, 64, 48), (512, 4, 4, 128, 96),
                              (384, 6, 3, 32, 20)]:
        c64 = shaped_c(d, spread=0.6, seed=5)
        wv = (torch.randn(KV * hd, d, device=dev) * 0.05).bfloat16()
        wo = (torch.randn(d, H * hd, device=dev) * 0.05).bfloat16()
        v, o = ops.vo_compress(c64.float(), 1e-5, wv, wo, H, KV, hd, r)
        _, v64, o64 = O.vo_layer(wv.float().cpu().numpy(), wo.float().cpu().numpy(),
                                 c64.float().double().cpu().numpy(), H, KV, hd, r, 1e-5)
        grp = H // KV
        worst = 0.0
        for h in range(KV):
            for j in range(grp):
                q = h * grp + j
                pr = o64[:, q * r:(q + 1) * r] @ v64[h * r:(h + 1) * r]
                po = (o[:, q * r:(q + 1) * r].double() @ v[h * r:(h + 1) * r].double()).cpu().numpy()
                worst = max(worst, np.linalg.norm(po - pr) / np.linalg.norm(pr))
        print(f"vo d={d} H={H} KV={KV} hd={hd} r={r}: worst rel(O'V' product)={worst:.3e}", flush=True)


def big():
    n, d = 11008, 4096
    x = torch.randn(32768, n, device=dev).bfloat16()
    c = torch.zeros(n, n, device=dev)
    ops.syrk_(c, x)
    ops.finalize_sym_(c, 1.0 / 32768)
    del x
    ms, s = timed(lambda: ops.ridge_scores(c, 1e-4), iters=2)
    ref = torch.linalg.inv(c.double() + 1e-4 * torch.eye(n, device=dev, dtype=torch.float64)).diagonal()
    print(f"BIG ridge_scores n={n}: {ms:.1f} ms  rel={rel(s.double(), ref):.3e}", flush=True)
    k = int(n * 0.75)
    idx = ops.select_k(s, k)
    ref_idx = torch.sort(torch.topk(ref, k, largest=False).indices).values
    print(f"    topk overlap {len(set(idx.tolist()) & set(ref_idx.tolist()))}/{k}", flush=True)
    wd = (torch.randn(d, n, device=dev) * 0.02).bfloat16()
    ms, out = timed(lambda: ops.nystrom_down(c, idx, wd), iters=2)
    cd = c.double()
    ckk = cd[idx][:, idx] + 1e-6 * torch.eye(k, device=dev, dtype=torch.float64)
    t0 = time.time()
    ref_out = torch.cholesky_solve(cd[idx, :] @ wd.double().T, torch.linalg.cholesky(ckk)).T
    torch.cuda.synchronize()
    print(f"BIG nystrom_down k={k}: {ms:.1f} ms rel={rel(out.double(), ref_out):.3e} "
          f"(torch fp64 GPU same step: {1e3 * (time.time() - t0):.0f} ms)", flush=True)
    del cd, ckk, ref_out, c
    dm, H, hd, r = 4096, 32, 128, 96
    cx = shaped_c(dm, 8192, spread=0.5).float()
    wv = (torch.randn(H * hd, dm, device=dev) * 0.02).bfloat16()
    wo = (torch.randn(dm, H * hd, device=dev) * 0.02).bfloat16()
    ms, _ = timed(lambda: ops.vo_compress(cx, 1e-5, wv, wo, H, H, hd, r), iters=2)
    print(f"BIG vo MHA d={dm} H={H}: {ms:.1f} ms", flush=True)
    ms, _ = timed(lambda: ops.vo_compress(cx, 1e-5, wv[:8 * hd], wo, H, 8, hd, r), iters=2)
    print(f"BIG vo GQA d={dm} H={H} KV=8: {ms:.1f} ms", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    t0 = time.time()
    check_qk()
    check_scores(128, 1e-2)
    check_scores(200, 1e-2)
    check_nystrom(512, 128)
    check_nystrom(1000, 192, ridge=1e-4)
    check_nystrom(1536, 256, ridge=1e-4, spread=0.2)
    check_vo()
    if "--big" in sys.argv:
        big()
    print(f"done in {time.time() - t0:.1f}s", flush=True)
