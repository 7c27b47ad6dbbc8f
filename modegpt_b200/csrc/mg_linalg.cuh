// mg_linalg.cuh — internal interface of the dense linear-algebra building blocks (mg_linalg.cu):
// bf16 plane splitting, the 128-wide diagonal-block factor/invert kernel, the blocked Cholesky
// driver, gathers, transposes and the radix select.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

constexpr int kNB = 128;      // panel width of every blocked algorithm
constexpr int kPlanes = 3;    // fp32 = hi + mid + lo bf16 planes (24 mantissa bits)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// dst_p[r, c] (or dst_p[c, r] when transpose) = p-th bf16 plane of src[r, c];
// optional colsumsq[c] += sum_r src[r, c]^2; diag_add is added to src[i, i] before splitting.
int split_planes(const float* src, int64_t ld_src, int64_t rows, int64_t cols, __nv_bfloat16* dst,
                 int64_t ld_dst, int64_t plane_stride, bool transpose, float* colsumsq,
                 cudaStream_t s, float diag_add = 0.f);

// Factor the nb x nb (nb <= 128) diagonal block at (j0, j0) of the fp32 matrix A (upper, in
// place, fp64 arithmetic in shared memory), and emit W = U11^-1 as bf16 planes: w_planes holds
// W row-major [128 x 128] (x3), wt_planes holds W^T.  On a non-positive pivot the 1-based global
// index is recorded in *info (first failure wins) and the pivot is clamped so nothing is NaN.
// u_planes (optional) receives the bf16 planes of U11 (zeros below the diagonal) at the same
// coordinates as in A, leading dimension ld_up; l_planes (optional) the transpose.
int diag_block_factor(float* A, int64_t ld, int64_t j0, int nb, __nv_bfloat16* w_planes,
                      __nv_bfloat16* wt_planes, __nv_bfloat16* u_planes, __nv_bfloat16* l_planes,
                      int64_t ld_up, int64_t up_plane_stride, int* info, cudaStream_t s);

struct CholWorkspace {
  // all device pointers, carved from the caller's workspace
  __nv_bfloat16* u_planes;   // [3][n_pad x n_pad]  planes of U (upper), ld = n_pad
  __nv_bfloat16* l_planes;   // [3][n_pad x n_pad]  planes of U^T (optional, may be null)
  __nv_bfloat16* w_planes;   // [npanels][3][128 x 128]  W_j = U_jj^-1
  __nv_bfloat16* wt_planes;  // [npanels][3][128 x 128]  W_j^T
  __nv_bfloat16* row_planes; // [3][128 x n_pad] scratch
  int64_t n_pad;
};

// A (fp32, n x n, upper triangle) -> U with U^T U = A, in place; fills ws planes.
int cholesky_upper(float* A, int64_t n, int64_t ld, const CholWorkspace& ws, int* info,
                   cudaStream_t s);

}  // namespace mg
