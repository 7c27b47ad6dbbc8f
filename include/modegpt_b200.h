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
 *   - asynchronous: work is enqueued on `stream`; no host synchronisation unless stated.
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

#ifdef __cplusplus
}
#endif
#endif /* MODEGPT_B200_H */
