"""CPU oracle for the MoDeGPT compression hot path — TEST INFRASTRUCTURE ONLY.

A plain numpy / fp64 restatement of what the reference (cbacary/MoDeGPT) computes on the path
SURVEY.md §8 scopes.  Nothing in the product (`modegpt_b200/`) may import this module; only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py`.

Pinning: the reference ships no golden vectors or known-answer tests, so this
oracle is pinned against the reference ITSELF: `oracle/make_golden.py` imports the unmodified
reference functions from /root/reference (devices rebound to CPU), runs them on seeded inputs and
writes `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below against
those vectors.

Every function cites the reference lines it restates (paths relative to the reference tree).
All matrices are numpy float64 unless stated; weights follow torch's [out, in] convention.
"""
from __future__ import annotations

import numpy as np

SEQ_LEN = 2048  # hard-coded normaliser, src/calibration.py:141

# --------------------------------------------------------------------------------------------
# dtype helper
# --------------------------------------------------------------------------------------------


def to_bf16(x: np.ndarray) -> np.ndarray:
    """Round like torch's `.to(torch.bfloat16)` on an fp64 tensor (fp64 -> fp32 -> bf16, both
    round-to-nearest-even) and return the values as float32."""
    f = np.ascontiguousarray(x, dtype=np.float64).astype(np.float32)
    u = f.view(np.uint32).astype(np.uint64)
    nan = np.isnan(f)
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    out = rounded.astype(np.uint32).view(np.float32).copy()
    out[nan] = np.nan
    return out.reshape(np.shape(x))


# --------------------------------------------------------------------------------------------
# calibration statistics  (src/calibration.py:39-150, src/adapters/LlamaAdapter.py:115-147)
# --------------------------------------------------------------------------------------------


def gram_rows(act: np.ndarray) -> np.ndarray:
    """`H.T @ H` over all rows of a [..., n] activation.
    MLP hook: src/adapters/LlamaAdapter.py:127-136 (input of down_proj);
    input hook: src/adapters/LlamaAdapter.py:138-147 (per-sample X^T X summed over the batch —
    the same thing as one Gram over the flattened rows);
    OPT fc hook applies relu first: src/adapters/model_adapter.py:546-554."""
    h = np.asarray(act, dtype=np.float64).reshape(-1, act.shape[-1])
    return h.T @ h


def gram_heads(proj_out: np.ndarray, n_heads: int, head_dim: int) -> np.ndarray:
    """Per-head Gram of a projection output [..., n_heads*head_dim] -> [n_heads, hd, hd].
    src/adapters/LlamaAdapter.py:115-125 (raw q_proj / k_proj outputs: pre-RoPE, pre-norm)."""
    p = np.asarray(proj_out, dtype=np.float64).reshape(-1, n_heads, head_dim)
    p = np.transpose(p, (1, 0, 2))
    return np.matmul(np.transpose(p, (0, 2, 1)), p)


def bi_batch(x_in: np.ndarray, x_out: np.ndarray, eps: float = 1e-8) -> float:
    """One batch's Block-Influence increment for one layer: sum over the batch of
    (1 - cos) per position, then the mean over positions.  src/calibration.py:118-124.
    x_*: [B, T, D]."""
    a = np.asarray(x_in, dtype=np.float64)
    b = np.asarray(x_out, dtype=np.float64)
    na = np.maximum(np.linalg.norm(a, axis=2), eps)
    nb = np.maximum(np.linalg.norm(b, axis=2), eps)
    cos = np.sum(a * b, axis=2) / (na * nb)
    return float(np.mean(np.sum(1.0 - cos, axis=0)))


def normalise_stats(c: np.ndarray, n_texts: int) -> np.ndarray:
    """`cov /= n_texts * 2048`, src/calibration.py:141-146."""
    return c / float(n_texts * SEQ_LEN)


# --------------------------------------------------------------------------------------------
# rank allocation  (src/compression_utils.py:79-124)
# --------------------------------------------------------------------------------------------


def allocate_global_sparsity(bi_scores, compression_ratio: float, smoothing: float = 0.015,
                             max_sparsity: float = 0.8) -> list[float]:
    """softmax(-BI/eps) spreads a budget of L*ratio over the layers; layers above the cap are
    clamped and their excess handed to the uncapped ones in proportion to their softmax weight,
    repeated until nothing exceeds the cap.  Returns keep ratios (1 - sparsity)."""
    # `torch.tensor(list_of_floats)` is float32: the scores are rounded to fp32 BEFORE the
    # fp64 softmax (compression_utils.py:96) — a quirk that moves keep ratios by ~1e-8.
    s = np.asarray(bi_scores, dtype=np.float64).astype(np.float32).astype(np.float64)
    z = -s / smoothing
    z = z - z.max()
    w = np.exp(z)
    w = w / w.sum()
    sp = w * (len(s) * compression_ratio)
    while True:
        over = sp > max_sparsity
        if not over.any():
            break
        excess = (sp[over] - max_sparsity).sum()
        sp[over] = max_sparsity
        free = ~over
        if free.any():
            sp[free] = sp[free] + excess * (w[free] / w[free].sum())
    return (1.0 - sp).tolist()


# --------------------------------------------------------------------------------------------
# matrix square root  (src/compression_utils.py:15-55)
# --------------------------------------------------------------------------------------------


def sqrt_psd(m: np.ndarray, ridge: float = 1e-4, inverse: bool = False):
    """V diag(sqrt(max(lambda + ridge, 0))) V^T from a symmetric eigendecomposition; the ridge is
    added to the eigenvalues unscaled (scaled=False path)."""
    lam, vec = np.linalg.eigh(np.asarray(m, dtype=np.float64))
    root = np.sqrt(np.clip(lam + ridge, 0.0, None))
    s = (vec * root) @ vec.T
    if not inverse:
        return s
    inv_root = 1.0 / np.clip(root, 1e-12, None)
    return s, (vec * inv_root) @ vec.T


# --------------------------------------------------------------------------------------------
# type-I  Nystrom MLP  (src/compression/compress_mlp.py:13-64)
# --------------------------------------------------------------------------------------------


def _solve_lower(low: np.ndarray, rhs: np.ndarray, trans: bool = False) -> np.ndarray:
    """low^-1 rhs (or low^-T rhs) by a triangular solve — the flop count of the reference's
    cholesky_inverse / cholesky_solve, so the CPU baseline is not charged an LU it never does."""
    try:
        from scipy.linalg import solve_triangular
    except ImportError:                                # numpy only: LU of a triangular matrix
        return np.linalg.solve(low.T if trans else low, rhs)
    return solve_triangular(low, rhs, lower=True, trans="T" if trans else "N", check_finite=False)


def ridge_scores(c: np.ndarray, ridge: float) -> np.ndarray:
    """diag((C + ridge*I)^-1) through a Cholesky factorisation (compress_mlp.py:13-25)."""
    c = np.asarray(c, dtype=np.float64)
    n = c.shape[0]
    # `ridge * torch.eye(n)` is a float32 tensor in the reference (no dtype given), so the ridge
    # that reaches the fp64 sum is float32(ridge).
    ridge = float(np.float32(ridge))
    low = np.linalg.cholesky(c + ridge * np.eye(n))
    low_inv = _solve_lower(low, np.eye(n))     # any exact route to the inverse diagonal
    return np.sum(low_inv * low_inv, axis=0)


def topk_smallest_sorted(scores: np.ndarray, k: int) -> np.ndarray:
    """indices of the k smallest scores, ascending index order (compress_mlp.py:45-47)."""
    order = np.argsort(scores, kind="stable")[:k]
    return np.sort(order)


def nystrom_mlp(w_up, w_gate, w_down, c, keep_ratio: float, ridge: float):
    """compress_weights (compress_mlp.py:28-64) followed by the transposes of
    compress_nystrom (:97).  Returns dict(up [r,d], gate [r,d], down [d,r]) as bf16-rounded
    float32, the kept indices and the rank.  w_gate may be None (OPT-style MLP)."""
    c = np.asarray(c, dtype=np.float64)
    n = c.shape[0]
    rank = int(n * keep_ratio)
    idx = topk_smallest_sorted(ridge_scores(c, ridge), rank)
    w_up = np.asarray(w_up, dtype=np.float64)
    w_down = np.asarray(w_down, dtype=np.float64)
    up = w_up[idx, :]
    gate = None if w_gate is None else np.asarray(w_gate, dtype=np.float64)[idx, :]
    c_kk = c[np.ix_(idx, idx)]
    cross = c[idx, :] @ w_down.T                      # [r, d]
    low = np.linalg.cholesky(c_kk + 1e-6 * np.eye(rank))
    y = _solve_lower(low, cross)
    down_t = _solve_lower(low, y, trans=True)         # [r, d] == cholesky_solve(cross, L)
    out = {"up": to_bf16(up), "down": to_bf16(down_t.T)}
    if gate is not None:
        out["gate"] = to_bf16(gate)
    return out, idx, rank


# --------------------------------------------------------------------------------------------
# type-II  CR  Q/K  (src/compression/compress_qk.py:152-476)
# --------------------------------------------------------------------------------------------


def head_rank(head_dim: int, keep_ratio: float, rope: bool, clamp_to_head: bool = True) -> int:
    """Per-head rank rule shared by Q/K (compress_qk.py:176-182) and V/O (compress_vo.py:35-41;
    the V/O variant does not clamp to head_dim)."""
    r = int(head_dim * keep_ratio)
    r = max(1, min(r, head_dim)) if clamp_to_head else max(1, r)
    if rope:
        r = r - (r % 2)
        r = max(2, min(r, head_dim)) if clamp_to_head else max(2, r)
    return r


def _col_norms(m: np.ndarray) -> np.ndarray:
    return np.linalg.norm(m, axis=0)


def _topk_desc(score: np.ndarray, k: int) -> np.ndarray:
    return np.argsort(-score, kind="stable")[:k]


def qk_head_gqa(c_q_group: np.ndarray, c_k: np.ndarray, rank: int, ridge_k: float,
                ridge_q: float = 1e-4) -> np.ndarray:
    """compress_head_llama_grouped (compress_qk.py:320-382): RoPE-paired CR score summed over the
    query heads of the group, square-rooted, top rank/2 pairs; returns the row mask
    cat(idx, idx + hd/2) in top-k (descending score) order."""
    hd = c_k.shape[0]
    half = hd // 2
    rk = sqrt_psd(c_k, ridge_k)
    nk1, nk2 = _col_norms(rk[:, :half]), _col_norms(rk[:, half:])
    score = np.zeros(half)
    for c_q in c_q_group:
        rq = sqrt_psd(c_q, ridge_q)
        nq1, nq2 = _col_norms(rq[:, :half]), _col_norms(rq[:, half:])
        score += nq1 ** 2 * nk1 ** 2 + nq2 ** 2 * nk2 ** 2
    score = np.sqrt(score)
    top = _topk_desc(score, rank // 2)
    return np.concatenate([top, top + half]).astype(np.int64)


def qk_head_mha(c_q: np.ndarray, c_k: np.ndarray, rank: int, ridge: float = 1e-4) -> np.ndarray:
    """compress_head_llama (compress_qk.py:387-436): same pairing, ridge 1e-4 on both sides, no
    square root on the score."""
    hd = c_k.shape[0]
    half = hd // 2
    rq, rk = sqrt_psd(c_q, ridge), sqrt_psd(c_k, ridge)
    score = (_col_norms(rq[:, :half]) ** 2 * _col_norms(rk[:, :half]) ** 2
             + _col_norms(rq[:, half:]) ** 2 * _col_norms(rk[:, half:]) ** 2)
    top = _topk_desc(score, rank // 2)
    return np.concatenate([top, top + half]).astype(np.int64)


def qk_head_opt(c_q: np.ndarray, c_k: np.ndarray, rank: int, ridge: float = 1e-4) -> np.ndarray:
    """compress_head_opt (compress_qk.py:439-476): score = ||sqrt(Cq)[:,j]|| * ||sqrt(Ck)[:,j]||,
    top `rank` columns, no RoPE pairing."""
    score = _col_norms(sqrt_psd(c_q, ridge)) * _col_norms(sqrt_psd(c_k, ridge))
    return _topk_desc(score, rank).astype(np.int64)


def qk_layer(w_q, w_k, c_q, c_k, n_heads: int, n_kv_heads: int, head_dim: int, rank: int,
             arch: str, ridge_qk: float, b_q=None, b_k=None):
    """compress_layer (compress_qk.py:208-308).  Returns dict(q_proj [H*r, d], k_proj [KV*r, d])
    bf16-rounded, the [KV, r] int64 mask (None for OPT) and, for OPT, the gathered biases.
    Output head order follows the reference: for each kv head, its K rows, and its group's Q heads
    in order."""
    w_q = np.asarray(w_q)
    w_k = np.asarray(w_k)
    wq = w_q.reshape(n_heads, head_dim, -1)
    wk = w_k.reshape(n_kv_heads, head_dim, -1)
    group = n_heads // n_kv_heads
    rope = (arch == "llama") or ("qwen" in arch)
    q_out, k_out, masks, bq_out, bk_out = [], [], [], [], []
    for h in range(n_kv_heads):
        if rope and n_kv_heads != n_heads:
            m = qk_head_gqa(c_q[h * group:(h + 1) * group], c_k[h], rank, ridge_qk)
            for j in range(group):
                q_out.append(wq[h * group + j][m])
            k_out.append(wk[h][m])
            masks.append(m)
        elif arch == "llama":
            m = qk_head_mha(c_q[h], c_k[h], rank)
            q_out.append(wq[h][m])
            k_out.append(wk[h][m])
            masks.append(m)
        elif arch == "opt":
            m = qk_head_opt(c_q[h], c_k[h], rank)
            q_out.append(wq[h][m])
            k_out.append(wk[h][m])
            if b_q is not None:
                bq_out.append(np.asarray(b_q)[h * head_dim:(h + 1) * head_dim][m])
                bk_out.append(np.asarray(b_k)[h * head_dim:(h + 1) * head_dim][m])
            masks.append(m)
        else:
            raise NotImplementedError(arch)
    out = {"q_proj": to_bf16(np.concatenate(q_out, 0)), "k_proj": to_bf16(np.concatenate(k_out, 0))}
    mask = np.stack(masks, 0).astype(np.int64)
    if arch == "opt" and bq_out:
        out["q_bias"] = np.concatenate(bq_out)
        out["k_bias"] = np.concatenate(bk_out)
    return out, mask


# --------------------------------------------------------------------------------------------
# type-III  SVD  V/O  (src/compression/compress_vo.py:13-223)
# --------------------------------------------------------------------------------------------


def vo_roots(c_x: np.ndarray, ridge_vo: float):
    """sqrt_M(C, ridge_vo) and its LU inverse, compress_vo.py:43-45."""
    root = sqrt_psd(c_x, ridge_vo)
    return root, np.linalg.inv(root)


def vo_head_gqa(w_v_head, w_o_group, root, root_inv, rank: int):
    """compress_head_grouped (compress_vo.py:112-159).  w_v_head [hd, d]; w_o_group list of
    [d, hd].  Returns V' [r, d] and the list of O'_j [d, r]."""
    u, s, vt = np.linalg.svd(root @ np.asarray(w_v_head, dtype=np.float64).T, full_matrices=False)
    v_new = (root_inv @ u[:, :rank]).T
    o_new = []
    for w_o in w_o_group:
        o_new.append((np.diag(s)[:rank, :rank] @ vt[:rank, :] @ np.asarray(w_o, np.float64).T).T)
    return v_new, o_new


def vo_head_mha(w_v_head, w_o_head, root, root_inv, rank: int):
    """compress_head (compress_vo.py:162-223): a thin SVD of sqrt(C) Wv^T, then a FULL SVD of
    S V^T Wo^T whose leading `rank` components give the new factors."""
    u, s, vt = np.linalg.svd(root @ np.asarray(w_v_head, dtype=np.float64).T, full_matrices=False)
    a = np.diag(s) @ vt @ np.asarray(w_o_head, dtype=np.float64).T       # [hd, d]
    up, sp, vpt = np.linalg.svd(a, full_matrices=True)
    v_new = (root_inv @ u @ up)[:, :rank].T                               # [r, d]
    o_new = (np.diag(sp)[:rank, :rank] @ vpt[:rank, :]).T                 # [d, r]
    return v_new, o_new


def vo_layer(w_v, w_o, c_x, n_heads: int, n_kv_heads: int, head_dim: int, rank: int,
             ridge_vo: float):
    """compress_vo body for one layer (compress_vo.py:43-99).  Returns dict(v_proj [KV*r, d],
    o_proj [d, H*r]) bf16-rounded plus the unrounded fp64 factors for sign-free comparisons."""
    w_v = np.asarray(w_v, dtype=np.float64)
    w_o = np.asarray(w_o, dtype=np.float64)
    root, root_inv = vo_roots(c_x, ridge_vo)
    group = n_heads // n_kv_heads
    v_heads, o_heads = [], []
    for h in range(n_kv_heads):
        wv = w_v[h * head_dim:(h + 1) * head_dim, :]
        if n_kv_heads != n_heads:
            wo = [w_o[:, (h * group + j) * head_dim:(h * group + j + 1) * head_dim]
                  for j in range(group)]
            v_new, o_new = vo_head_gqa(wv, wo, root, root_inv, rank)
            v_heads.append(v_new)
            o_heads.extend(o_new)
        else:
            v_new, o_new = vo_head_mha(wv, w_o[:, h * head_dim:(h + 1) * head_dim], root,
                                       root_inv, rank)
            v_heads.append(v_new)
            o_heads.append(o_new)
    v64 = np.concatenate(v_heads, 0)
    o64 = np.concatenate(o_heads, 1)
    return {"v_proj": to_bf16(v64), "o_proj": to_bf16(o64)}, v64, o64


# --------------------------------------------------------------------------------------------
# perplexity formula  (src/eval.py:192-220)
# --------------------------------------------------------------------------------------------


def perplexity_from_batch_losses(mean_losses, batch_sizes, seqlen: int = SEQ_LEN) -> float:
    """Each batch contributes mean-CE * (seqlen-1) * batch; ppl = exp(sum / (nsamples*(seqlen-1)))."""
    nll = sum(l * (seqlen - 1) * b for l, b in zip(mean_losses, batch_sizes))
    return float(np.exp(nll / (sum(batch_sizes) * (seqlen - 1))))
