"""GPU tests at BASELINE.json's full sizes, through size-independent properties (the CPU oracle
cannot finish these sizes in seconds): sub-block spot checks, linearity, symmetry, residuals of the
linear systems, order statistics.  The checker is plain torch fp64 on the same device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def ops():
    from modegpt_b200 import ops as _ops

    return _ops


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


def activations(T, n, seed, spread=0.5):
    g = torch.Generator(device=DEV).manual_seed(seed)
    x = torch.randn(T, n, device=DEV, generator=g)
    return (x * torch.exp(spread * torch.randn(n, device=DEV, generator=g))).bfloat16()


@pytest.fixture(scope="module")
def c_mlp(ops):
    """C_mlp of Llama-2-7B shape (n = 11008) from 16384 synthetic tokens, normalised + mirrored."""
    n, T = 11008, 16384
    x = activations(T, n, 0)
    c = torch.zeros(n, n, device=DEV)
    ops.syrk_(c, x[: T // 2])
    ops.syrk_(c, x[T // 2:])
    ops.finalize_sym_(c, 1.0 / T)
    return c, x


def test_syrk_7b_mlp_shape_blocks_linearity_symmetry(ops, c_mlp):
    c, x = c_mlp
    n, T = c.shape[0], x.shape[0]
    xd = x.double()
    for (r0, c0) in [(0, 0), (128, 10880), (5000, 5100), (10900, 10900), (3, 7777)]:
        ref = xd[:, r0:r0 + 100].T @ xd[:, c0:c0 + 100] / T
        # bar is 1e-3; the measured 1e-5..3e-5 is the tensor core's truncating fp32 accumulation
        assert rel(c[r0:r0 + 100, c0:c0 + 100], ref) < 1e-4
    assert torch.equal(c, c.T)                       # finalize mirrors exactly
    whole = torch.zeros(n, n, device=DEV)            # two half-batches == one batch (fp32 order only)
    ops.syrk_(whole, x)
    ops.finalize_sym_(whole, 1.0 / T)
    assert rel(whole, c) < 1e-4


@pytest.mark.parametrize("n,T,H,hd", [(14336, 4096, 32, 128), (28672, 2048, 64, 128), (3072, 8192, 12, 64),
                                      (18944, 2048, 28, 128)])
def test_syrk_other_baseline_shapes(ops, n, T, H, hd):
    """Llama-3-8B (d_int 14336), Llama-2-70B (d_int 28672), OPT-125M (3072), Qwen2.5-7B (d_int 18944,
    28 heads: d = 3584) operand widths."""
    x = activations(T, n, n)
    c = torch.zeros(n, n, device=DEV)
    ops.syrk_(c, x)
    xd = x.double()
    for (r0, c0) in [(0, 0), (n - 200, n - 100), (n // 3, n // 2)]:
        ref = xd[:, r0:r0 + 64].T @ xd[:, c0:c0 + 64]
        got = c[r0:r0 + 64, c0:c0 + 64]
        if r0 == c0:        # not finalised: only the upper triangle is defined
            ref, got = torch.triu(ref), torch.triu(got)
        assert rel(got, ref) < 1e-4
    d = H * hd
    ch = torch.zeros(H, hd, hd, device=DEV)
    ops.syrk_heads_(ch, x[:, :d])
    xh = xd[:, :d].view(T, H, hd)
    for h in (0, H // 2, H - 1):
        assert rel(ch[h], xh[:, h].T @ xh[:, h]) < 1e-4


def test_type1_7b_scores_selection_and_solve(ops, c_mlp):
    c, _ = c_mlp
    n, d, ridge = c.shape[0], 4096, 1e-4
    scores = ops.ridge_scores(c, ridge)
    # scores are diag((C + ridge I)^-1): check against fp64 solves of (C + ridge I) x = e_j
    a = c.double() + ridge * torch.eye(n, device=DEV, dtype=torch.float64)
    chol = torch.linalg.cholesky(a)
    js = torch.tensor([0, 17, 5503, 9999, n - 1], device=DEV)
    e = torch.zeros(n, js.numel(), device=DEV, dtype=torch.float64)
    e[js, torch.arange(js.numel())] = 1.0
    xs = torch.cholesky_solve(e, chol)
    want = xs[js, torch.arange(js.numel())]
    assert rel(scores[js], want) < 1e-4
    # selection: exactly k indices, strictly ascending, every kept score <= every dropped score
    k = int(n * 0.75)
    idx = ops.select_k(scores, k)
    assert idx.numel() == k and bool((idx[1:] > idx[:-1]).all())
    keep = torch.zeros(n, dtype=torch.bool, device=DEV)
    keep[idx] = True
    assert scores[keep].max() <= scores[~keep].min()
    ref_idx = torch.sort(torch.topk(torch.linalg.inv(a).diagonal(), k, largest=False).indices).values
    assert torch.equal(idx, ref_idx)
    # Nystrom solve: residual of (C_kk + 1e-6 I) X = C_k: Wd^T in fp64
    g = torch.Generator(device=DEV).manual_seed(3)
    wd = (torch.randn(d, n, device=DEV, generator=g) * 0.02).bfloat16()
    down = ops.nystrom_down(c, idx, wd)                   # [d, k] bf16
    assert down.shape == (d, k)
    cd = c.double()
    ckk = cd[idx][:, idx] + 1e-6 * torch.eye(k, device=DEV, dtype=torch.float64)
    rhs = cd[idx, :] @ wd.double().T
    ref = torch.cholesky_solve(rhs, torch.linalg.cholesky(ckk)).T
    assert rel(down, ref.bfloat16()) < 1e-3               # vs the reference's bf16-rounded result
    assert (down == ref.bfloat16()).float().mean() > 0.97
    up = ops.gather_rows(wd.T.contiguous(), idx)          # any [n, d] bf16 matrix
    assert torch.equal(up, wd.T.contiguous()[idx])


# type III at the BASELINE shapes, against the CPU oracle: tests/test_gpu_type3.py


def test_type2_7b_shape(ops):
    H, hd, r = 32, 128, 96
    x = activations(4096, H * hd, 21)
    cq = torch.zeros(H, hd, hd, device=DEV)
    ck = torch.zeros(H, hd, hd, device=DEV)
    ops.syrk_heads_(cq, x)
    ops.syrk_heads_(ck, activations(4096, H * hd, 22))
    mask = ops.qk_select(cq, ck, r, 0, 1e-4, 1e-4)
    dq = torch.diagonal(cq, dim1=1, dim2=2).double() + 1e-4
    dk = torch.diagonal(ck, dim1=1, dim2=2).double() + 1e-4
    score = dq[:, :64] * dk[:, :64] + dq[:, 64:] * dk[:, 64:]
    want = torch.topk(score, r // 2, dim=1).indices
    assert torch.equal(mask[:, : r // 2], want) and torch.equal(mask[:, r // 2:], want + 64)
    w = activations(H * hd, 4096, 23)
    out = ops.gather_head_rows(w, mask, H, 1, hd)
    rows = (torch.arange(H, device=DEV) * hd)[:, None] + mask
    assert torch.equal(out, w[rows.reshape(-1)])


def test_type1_70b_mlp_width(ops):
    """Llama-2-70B MLP width (n = 28672, d = 8192, keep 0.7): scores, selection and the Nystrom
    solve at the largest size BASELINE.json names, checked through fp64 solves / residuals."""
    n, d, T, ridge = 28672, 8192, 32768, 1e-4
    x = activations(T, n, 70, spread=0.2)
    c = torch.zeros(n, n, device=DEV)
    ops.syrk_(c, x)
    ops.finalize_sym_(c, 1.0 / T)
    del x
    # T / n = 1.14: cond(C + ridge I) ~ 1e3-1e4.  The factorisation is fp32 (bf16x3 tensor-core
    # products, fp64 diagonal blocks), so the score error scales like cond * 2^-24; with T < n the
    # matrix is rank-deficient (cond ~ lambda_max / ridge ~ 1e7) and 1e-2 relative is the physics.
    scores = ops.ridge_scores(c, ridge)
    a = c.double()
    a.diagonal().add_(ridge)
    chol = torch.linalg.cholesky(a)
    del a
    js = torch.tensor([1, 9000, 20000, n - 2], device=DEV)
    e = torch.zeros(n, js.numel(), device=DEV, dtype=torch.float64)
    e[js, torch.arange(js.numel())] = 1.0
    want = torch.cholesky_solve(e, chol)[js, torch.arange(js.numel())]
    del chol, e
    assert rel(scores[js], want) < 2e-3
    k = int(n * 0.7)
    idx = ops.select_k(scores, k)
    keep = torch.zeros(n, dtype=torch.bool, device=DEV)
    keep[idx] = True
    assert idx.numel() == k and scores[keep].max() <= scores[~keep].min()
    g = torch.Generator(device=DEV).manual_seed(7)
    wd = (torch.randn(d, n, device=DEV, generator=g) * 0.02).bfloat16()
    down = ops.nystrom_down(c, idx, wd)                    # [d, k]
    assert down.shape == (d, k) and bool(torch.isfinite(down.float()).all())
    # residual of (C_kk + 1e-6 I) X = C_k: Wd^T on 64 right-hand sides, in fp64
    cols = torch.arange(0, d, d // 64, device=DEV)[:64]
    ckk = c[idx][:, idx].double()
    ckk.diagonal().add_(1e-6)
    rhs = c[idx, :].double() @ wd[cols].double().T         # [k, 64]
    ref = torch.cholesky_solve(rhs, torch.linalg.cholesky(ckk))
    assert rel(down[cols].T, ref.bfloat16()) < 5e-3
