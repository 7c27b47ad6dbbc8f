// mg_linalg.cuh — internal interface of the dense linear-algebra building blocks (mg_linalg.cu):
// bf16 plane splitting, the 128-wide diagonal-block Cholesky kernel, the fused triangular-solve
// panel kernel and the blocked Cholesky driver.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "mg_lanes.cuh"

namespace mg {

constexpr int kNB = 128;      // panel width of every blocked algorithm
constexpr int kPlanes = 3;    // fp32 = hi + mid + lo bf16 planes (24 mantissa bits)
constexpr int kTLd = 132;     // leading dimension of the compact triangular blocks (float4 rows)
constexpr int kTBlock = kNB * kTLd;

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// dst_p[r, c] (or dst_p[c, r] when transpose) = p-th bf16 plane of src[r, c];
// optional colsumsq[c] += sum_r src[r, c]^2; diag_add is added to src[i, i] before splitting;
// upper_only treats src[r, c] with c < r as zero (the strict lower triangle of an in-place
// Cholesky factor is scratch).
int split_planes(const float* src, int64_t ld_src, int64_t rows, int64_t cols, __nv_bfloat16* dst,
                 int64_t ld_dst, int64_t plane_stride, bool transpose, float* colsumsq,
                 cudaStream_t s, float diag_add = 0.f, bool upper_only = false);

// Cholesky of the nb x nb (nb <= 128) diagonal block at (j0, j0) of the fp32 matrix A: upper
// factor U11 written back in place (upper triangle), fp64 arithmetic, hierarchical 32-blocking
// (register-resident 32x32 factorisation by one warp, substitution + rank-32 update by the CTA).
// Also emits
//   t_fwd [128 x 132] fp32 : T[m][i] = U11[m][i]                (solves U11^T x = b)
//   t_bwd [128 x 132] fp32 : T[m'][i'] = U11[nb-1-i'][nb-1-m']  (solves U11 x = b, reversed rows)
// both identity-padded beyond nb.  (No bf16 planes: every tensor-core product of the blocked
// drivers reads block rows strictly outside the diagonal blocks.)
// A non-positive pivot records its 1-based global index in *info (first failure wins).
int potrf128(float* A, int64_t ld, int64_t j0, int nb, float* t_fwd, float* t_bwd, int* info,
             cudaStream_t s);

// Triangular solve with one compact 128-block against many right-hand-side columns, one thread
// per column:  X = alpha * T^-1 B  where T is t_fwd (rows in natural order) or t_bwd (reversed).
//   B    [nb x ncols] fp32 (ldb)            right-hand sides
//   X    [nb x ncols] fp32 (ldx), optional  (may alias B)
//   planes  [3][nb x ncols] bf16 (ldp, pstride), optional: planes of X at [row][col]
//   tplanes [3][ncols x nb] bf16 (ldtp, tpstride), optional: planes of X^T at [col][row]
//   colsumsq[ncols], optional: += sum_rows X^2
int trsm128(const float* tblock, bool reversed, int nb, const float* B, int64_t ldb, int64_t ncols,
            float alpha, float* X, int64_t ldx, __nv_bfloat16* planes, int64_t ldp, int64_t pstride,
            __nv_bfloat16* tplanes, int64_t ldtp, int64_t tpstride, float* colsumsq,
            cudaStream_t s);

struct CholWorkspace {
  // all device pointers, carved from the caller's workspace
  __nv_bfloat16* u_planes;   // [3][n_pad x n_pad]  planes of U (upper), ld = n_pad
  __nv_bfloat16* l_planes;   // [3][n_pad x n_pad]  planes of U^T (optional, may be null)
  float* t_fwd;              // [npanels][128 x 132]
  float* t_bwd;              // [npanels][128 x 132]
  int64_t n_pad;
};

// A (fp32, n x n, upper triangle) -> U with U^T U = A, in place; fills ws.  Right-looking, panel
// 128, with one panel of look-ahead over the lanes (mg_lanes.cuh).  `step(pj)` enqueues panel pj:
//   chain : potrf128(pj), trsm128(pj) [records lanes.trsm], then the update of block row pj+1
//           (after the upd lane has finished panel pj-1, which keeps the L2 reduce-adds ordered
//           and the result deterministic);
//   upd   : waits for lanes.trsm, updates the rows below block row pj+1, records upd_done[pj&1].
// Callers interleave their own per-panel work on the tri lane between steps (it may wait on
// lanes.trsm right after step(pj) to consume block row pj).
struct CholStepper {
  float* A;
  int64_t n, ld;
  CholWorkspace ws;
  int* info;
  const Lanes* lanes;
  int64_t panels() const { return (n + kNB - 1) / kNB; }
  int step(int64_t pj) const;
  int step_panelwise(int64_t pj) const;   // one stream / MG_CHOL_OUTER=1: panel-wise right-looking
};

int cholesky_upper(float* A, int64_t n, int64_t ld, const CholWorkspace& ws, int* info,
                   const Lanes& lanes);

}  // namespace mg
