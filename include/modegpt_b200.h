/* modegpt_b200.h — C ABI of libmodegpt_b200.so (sm_100a).
 *
 * The reference (cbacary/MoDeGPT) has no FFI layer: its hot path is Python calling ATen in fp64.
 * Each entry point below replaces one group of those ATen call sites; the reference location is
 * cited per function (paths relative to the reference tree).  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: raw device pointers, explicit sizes / leading dimensions (in ELEMENTS), explicit
 *     stream (`void*` = cudaStream_t, NULL = legacy default stream); no torch types.
 *   - the caller owns every buffer; no hidden device allocation.  Functions that need scratch take
 *     (ws, ws_bytes) and have a `*_ws_bytes` query.
 *   - return value: 0 = OK; < 0 = argument error (-1000 - cudaError_t for CUDA runtime errors);
 *     > 0 = numerical info (1-based index of the first non-positive pivot).
 *   - asynchronous and stream-ordered: to the caller every entry point behaves as one operation
 *     enqueued on `stream`; no host synchronisation unless stated.  The two blocked type-I entry
 *     points (mg_ridge_scores_f32, mg_nystrom_down_f32) fork `stream` into internal streams
 *     ("lanes", created once per device; events only, no memory) and join them before returning,
 *     so their panel chain, trailing updates and triangular-inverse / substitution work overlap.
 *     Two lane sets exist per device: two host threads may call them concurrently; a third
 *     concurrent call waits for a set.  MG_SERIAL=1 keeps everything on `stream`.
 *   - matrices are row-major.  "upper" = elements with col >= row are defined; the strict lower
 *     triangle is unspecified unless the function says it mirrors.
 */
#ifndef MODEGPT_B200_H
#define MODEGPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* library / device ----------------------------------------------------------------------------- */
int mg_version(void);              /* ABI version, currently 1 */
int mg_device_sm_count(void);      /* SM count of the current device */
const char* mg_error_string(int code);

/* ---- statistics (calibration hooks) ---------------------------------------------------------- */

/* C[n,n] (upper) = accumulate ? C + alpha*X^T X : alpha*X^T X.   X: bf16 [T, n], ld = ldx.
 * Replaces `cov_mlp += H.T @ H` (src/adapters/LlamaAdapter.py:127-136) and
 * `cov_x += sum_b X_b^T X_b` (src/adapters/LlamaAdapter.py:138-147); fp32 accumulate on tcgen05. */
int mg_syrk_bf16_f32(const void* X, int64_t T, int64_t n, int64_t ldx, float* C, int64_t ldc,
                     float alpha, int accumulate, void* stream);

/* Per-head Gram: C[h] (+)= alpha * X_h^T X_h, X_h = X[:, h*hd:(h+1)*hd]; C is [n/hd, hd, hd]
 * (full blocks, both triangles).  hd in {32, 64, 128}.
 * Replaces the permute + fp64 bmm of src/adapters/LlamaAdapter.py:115-125 and
 * src/adapters/model_adapter.py:556-567. */
int mg_syrk_heads_bf16_f32(const void* X, int64_t T, int64_t n, int64_t ldx, int hd, float* C,
                           float alpha, int accumulate, void* stream);

/* heads[h] += alpha * (diagonal hd x hd block h of the upper-triangular Gram `full` [n, n], mirrored
 * to a full block).  With mg_syrk_bf16_f32 on the whole projection this gives the per-head Grams
 * for head dims mg_syrk_heads_bf16_f32 does not take (e.g. OPT-2.7b's 80): H times the flops of
 * the block-diagonal tile set on a matrix that is small anyway. */
int mg_add_diag_blocks_f32(const float* full, int64_t n, int64_t ld, int hd, float alpha, float* heads,
                           void* stream);

/* acc[0] += sum_rows (1 - cos(x_in[row], x_out[row])), accumulated in fp64.
 * Replaces the Block-Influence loop body, src/calibration.py:118-124. */
int mg_bi_cosine_bf16(const void* x_in, int64_t ld_in, const void* x_out, int64_t ld_out,
                      int64_t rows, int64_t d, double* acc, void* stream);

/* C = scale * C on the upper triangle, then mirror it into the lower triangle.
 * Replaces the normalisation `cov /= total_tokens`, src/calibration.py:141-146 (the reference
 * hands full symmetric matrices to the decompositions). */
int mg_finalize_sym_f32(float* C, int64_t n, int64_t ldc, float scale, void* stream);

/* scale a dense fp32 buffer in place (per-head blocks are already full). */
int mg_scale_f32(float* x, int64_t count, float scale, void* stream);

/* Wire format of the cross-rank reduction of a symmetric accumulator ("one exchange
 * step per layer", proposed there as mg_comm_reduce_to_owner).  The SYRK kernels accumulate the
 * upper triangle of a square fp32 matrix; only n(n+1)/2 values carry information:
 *   packed[i*n - i*(i-1)/2 + (c - i)] = C[i, c]  for c >= i   (row-major, row i = columns i..n-1).
 * mg_pack_upper_f32 gathers them into one contiguous buffer (which torch.distributed reduces over
 * NCCL/NVLink to the layer's owner, one collective per layer, overlapped with the forward of the
 * following layers), mg_unpack_upper_f32 scatters the sum back into the owner's accumulator; the
 * strict lower triangle is untouched (mg_finalize_sym_f32 mirrors afterwards).
 * The reference has no distributed code; this replaces nothing there. */
int mg_pack_upper_f32(const float* C, int64_t n, int64_t ldc, float* packed, void* stream);
int mg_unpack_upper_f32(const float* packed, int64_t n, float* C, int64_t ldc, void* stream);

/* ---- type-I: Nystrom MLP ---------------------------------------------------------------------- */

/* Tell the library how many type-I factorisations the caller runs side by side (host threads with
 * one stream each; the library has one set of internal lanes per concurrent call, up to 2).  The
 * persistent bulk GEMMs of each factorisation are then capped at 1/n of the SMs left after the
 * chain reserve, so a second factorisation's panel chain is never starved by the first one's
 * trailing updates.  Process-wide; default 1.  Results do not depend on it. */
int mg_set_concurrent_factorizations(int n);

/* scores[j] = diag((C + ridge I)^-1)_j.  C: fp32 [n,n], upper triangle read.  Two-level blocked
 * Cholesky (fp64 128-wide diagonal blocks, fused triangular-solve panels, tcgen05 updates on bf16x3
 * planes: K = 128 inside an outer block of 4 panels, K = 512 below it) and a blocked triangular
 * inverse that runs concurrently, one panel behind.  *info (device int, caller zero-initialises) receives the 1-based index of
 * the first non-positive pivot, else stays 0.
 * Replaces get_ridge_scores, src/compression/compress_mlp.py:13-25. */
size_t mg_ridge_scores_ws_bytes(int64_t n);
int mg_ridge_scores_f32(const float* C, int64_t n, int64_t ldc, float ridge, float* scores,
                        void* ws, size_t ws_bytes, int* info, void* stream);

/* idx_out[0..k) = indices of the k smallest (largest != 0: largest) scores in ASCENDING INDEX
 * order; ties at the threshold resolve to the lower index.
 * Replaces topk(largest=False) + sort, src/compression/compress_mlp.py:45-47. */
int mg_select_k_f32(const float* scores, int64_t n, int64_t k, int largest, int64_t* idx_out,
                    void* stream);

/* out[i, :] = W[idx[i], :], bf16 rows of length d.
 * Replaces W_u[topk, :], W_g[topk, :], src/compression/compress_mlp.py:49-50. */
int mg_gather_rows_bf16(const void* W, int64_t ldw, const int64_t* idx, int64_t k, int64_t d,
                        void* out, int64_t ldo, void* stream);

/* Wd_out[d, k] (bf16) = ((C[idx,idx] + jitter I)^-1 (C[idx, :] Wd^T))^T.
 * C: fp32 [n,n] FULL symmetric (both triangles valid); Wd: bf16 [d, n].
 * Solved in correction form (only the dropped channels' contribution passes through the fp32
 * factor).  *min_rel_pivot (device, may be NULL) receives min_i U_ii^2 / A_ii of the Cholesky factor
 * of A = C[idx,idx] + jitter I — a conditioning indicator of the EQUILIBRATED system: the fp32
 * solve is accurate to about 1e-7 / min_rel_pivot.
 * Replaces src/compression/compress_mlp.py:52-57 (+ the transpose at :97). */
size_t mg_nystrom_down_ws_bytes(int64_t n, int64_t k, int64_t d);
int mg_nystrom_down_f32(const float* C, int64_t n, int64_t ldc, const int64_t* idx, int64_t k,
                        const void* Wd, int64_t d, int64_t ldwd, float jitter, void* Wd_out,
                        int64_t ld_out, void* ws, size_t ws_bytes, int* info, float* min_rel_pivot,
                        void* stream);

/* One sweep of iterative refinement of the solution mg_nystrom_down_f32 left in `ws` (same C, idx,
 * k, d, jitter, ws; call right after it on the same stream, after checking *info == 0): the
 * residual C[idx,:] Wd^T - (C[idx,idx] + jitter I) X is formed in FP64 (CUDA cores, 2 k n d flop),
 * the correction is solved with the fp32 factor, Wd_out is rewritten.  Each sweep multiplies the
 * error by about 1e-7 / min_rel_pivot; the reference solves in fp64 (compress_mlp.py:56-57), and
 * this is how an ill-conditioned C_kk (ReLU-sparse OPT activations, heavy compression) reaches
 * the same bf16 result. */
int mg_nystrom_refine_f32(const float* C, int64_t n, int64_t ldc, const int64_t* idx, int64_t k,
                          int64_t d, float jitter, void* Wd_out, int64_t ld_out, void* ws,
                          size_t ws_bytes, void* stream);

/* ---- type-II: CR Q/K ------------------------------------------------------------------------- */

/* mask[h, :] for every kv head.  mode 0 (llama / qwen, RoPE-paired): score_j = sum over the
 * group's query heads of (Cq_jj + ridge_q)(Ck_jj + ridge_k) + same at j + hd/2; the top r/2 pairs
 * in descending-score order give mask = cat(idx, idx + hd/2).  mode 1 (OPT): score_j =
 * sqrt((Cq_jj + ridge_q)(Ck_jj + ridge_k)), top r.  Cq [H,hd,hd], Ck [KV,hd,hd] fp32; mask [KV,r].
 * Replaces compress_head_llama_grouped / compress_head_llama / compress_head_opt,
 * src/compression/compress_qk.py:320-476 (sqrt_M column norms == diagonal + ridge). */
int mg_qk_select_f32(const float* Cq, const float* Ck, int n_heads, int n_kv_heads, int hd,
                     int mode, float ridge_q, float ridge_k, int r, int64_t* mask, void* stream);

/* out[q*r + t, :] = W[q*hd + mask[(q/group)*r + t], :] for q < n_heads, t < r.
 * Replaces Q_heads[:, Sk_mask, :], K_head[Sk_mask, :], src/compression/compress_qk.py:369-376. */
int mg_gather_head_rows_bf16(const void* W, int64_t ldw, const int64_t* mask, int n_heads,
                             int group, int64_t hd, int64_t r, int64_t d, void* out, int64_t ldo,
                             void* stream);

/* ---- type-III: SVD V/O ----------------------------------------------------------------------- */

/* Wv_out[KV*r, d], Wo_out[d, H*r] (bf16, or fp32 when out_f32 != 0: the unrounded factors, for
 * accuracy checks) from Cx [d,d] fp32 FULL symmetric, Wv [KV*hd, d], Wo [d, H*hd] (bf16).
 * GQA when n_heads != n_kv_heads, else the MHA two-stage form.  hd: any multiple of 4 up to 128.
 * Replaces compress_vo's per-layer body, src/compression/compress_vo.py:43-99,112-223.
 *
 * method selects how the hd x hd Gram G1[h] = W_v,h (Cx + ridge I) W_v,h^T (whose eigenpairs are
 * the right singular pairs of the reference's sqrt(C) W_v,h^T) is obtained:
 *   MG_VO_FACTOR  Cx + ridge I = U^T U by the blocked Cholesky, M^T = W_v U^T on the tensor cores,
 *                 G1[h] = M_h^T M_h accumulated in fp64.  Small singular values keep a relative
 *                 accuracy of ~2^-24 sigma_1/sigma_r.  *info receives the 1-based index of a
 *                 non-positive pivot (Cx + ridge I numerically singular in fp32): the outputs are
 *                 then meaningless and the caller should repeat the layer with MG_VO_GRAM.
 *   MG_VO_GRAM    P = (Cx + ridge I) W_v^T on the tensor cores (fp32), G1[h] = W_v,h P_h in fp64.
 *                 No factorisation, works for any symmetric Cx; accuracy ~2^-24 (sigma_1/sigma_r)^2.
 * info is a device pointer (cleared by the call, stream-ordered). */
#define MG_VO_FACTOR 0
#define MG_VO_GRAM 1
size_t mg_vo_ws_bytes(int64_t d, int n_heads, int n_kv_heads, int hd);
int mg_vo_compress(const float* Cx, int64_t ldc, float ridge, const void* Wv, int64_t ldwv,
                   const void* Wo, int64_t ldwo, int n_heads, int n_kv_heads, int hd, int64_t d,
                   int r, int method, void* Wv_out, int64_t ldv_out, void* Wo_out, int64_t ldo_out,
                   int out_f32, int* info, void* ws, size_t ws_bytes, void* stream);

/* The same operation in two stream-ordered halves that share `ws` (mg_vo_ws_bytes):
 *   mg_vo_prepare : factorisation / tensor-core products and the fp64 Grams G1[h] and, for MHA,
 *                   G2[h] = W_o,h^T W_o,h — left in the workspace;
 *   mg_vo_finish  : the per-head eigensolves (one CTA per kv head, milliseconds) and the
 *                   recombination of the old heads into Wv_out / Wo_out.
 * A caller that prepares several layers back to back and then finishes them on different streams
 * overlaps their eigensolves (each uses n_kv_heads of the SMs); mg_vo_compress is prepare + finish. */
int mg_vo_prepare(const float* Cx, int64_t ldc, float ridge, const void* Wv, int64_t ldwv,
                  const void* Wo, int64_t ldwo, int n_heads, int n_kv_heads, int hd, int64_t d,
                  int method, int* info, void* ws, size_t ws_bytes, void* stream);
int mg_vo_finish(const void* Wv, int64_t ldwv, const void* Wo, int64_t ldwo, int n_heads,
                 int n_kv_heads, int hd, int64_t d, int r, void* Wv_out, int64_t ldv_out,
                 void* Wo_out, int64_t ldo_out, int out_f32, void* ws, size_t ws_bytes, void* stream);

/* ---- calibration forward: fused elementwise kernels ------------------------ */

/* y[r, :] = weight * bf16(x32[r, :] * rsqrt(mean(x32[r, :]^2) + eps)), every intermediate rounded
 * like HF's LlamaRMSNorm / Qwen2RMSNorm / Qwen3RMSNorm.forward (7 eager kernels -> 1). */
int mg_rmsnorm_bf16(const void* x, int64_t ldx, int64_t rows, int64_t d, const void* weight, float eps,
                    void* y, int64_t ldy, void* stream);

/* out = bf16(silu(gate)) * up, elementwise over `count` bf16 values
 * (LlamaMLP.forward: act_fn(gate_proj(x)) * up_proj(x); 2 eager kernels -> 1, bit-exact). */
int mg_swiglu_bf16(const void* gate, const void* up, void* out, int64_t count, void* stream);

/* out = bf16(x*cos) + bf16(rotate_half(x)*sin) on a contiguous [batch, seq, n_heads, head_dim]
 * buffer; cos / sin are [*, seq, head_dim] with batch stride cos_batch_stride (0 = broadcast)
 * (HF apply_rotary_pos_emb; 8 eager kernels per tensor -> 1, bit-exact). */
int mg_rope_bf16(const void* x, void* out, const void* cos, const void* sin, int64_t batch,
                 int64_t seq, int n_heads, int head_dim, int64_t cos_batch_stride, void* stream);

/* ---- rebuilt-model ops and perplexity -------------------------------- */

/* Masked RoPE of a compressed attention layer: x / out contiguous [batch, seq, n_heads, r] bf16,
 * cos / sin [*, seq, head_dim] of the ORIGINAL head dim, mask [n_heads / group, r] int64 (the
 * type-II selection: entries j < r/2 from the first half of the head, j + r/2 their partners).
 * out[j] = bf16(bf16(x[j] cos[mask[j]]) + bf16(rot(x)[j] sin[mask[j]])), rot = cat(-x[r/2:], x[:r/2]).
 * Replaces the per-forward `cos[:, :, mask]` gathers + 8 eager kernels of
 * src/patchers/LlamaRebuild.py:155-180 (bit-exact with that sequence). */
int mg_rope_masked_bf16(const void* x, void* out, const void* cos, const void* sin,
                        const int64_t* mask, int64_t batch, int64_t seq, int n_heads, int group, int r,
                        int head_dim, int64_t cos_batch_stride, void* stream);

/* Qwen3 q_norm / k_norm on compressed heads: x / out contiguous [rows, n_heads, r] bf16, weight
 * [head_dim] bf16 gathered through mask [n_heads / group, r]; RMS over the r kept dimensions.
 * Replaces src/patchers/DenseQwenRebuild.py:262-286. */
int mg_rmsnorm_masked_bf16(const void* x, int64_t rows, int n_heads, int group, int r,
                           const void* weight, const int64_t* mask, float eps, void* out,
                           void* stream);

/* nll[row] = logsumexp(logits[row, :vocab]) - logits[row, labels[row]] (fp32) from bf16 logits,
 * one pass.  The chunked evaluator (eval.compute_perplexity) calls it on row chunks of
 * hidden @ W_lm^T so [B, T, vocab] fp32 logits never exist (src/eval.py:192-220 formula). */
int mg_ce_rows_bf16(const void* logits, int64_t ld, int64_t rows, int64_t vocab,
                    const int64_t* labels, float* nll, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MODEGPT_B200_H */
