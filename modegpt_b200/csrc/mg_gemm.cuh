// mg_gemm.cuh — internal interface of the tensor-core engine (mg_gemm.cu).
//
// One persistent, warp-specialised tcgen05 kernel computes
//     D[M,N] (op)= alpha * sum_p sum_k A_{pa(p)}[k, m] * B_{pb(p)}[k, n]
// where A and B are row-major bf16 "planes" of shape [K, M] / [K, N] (both operands MN-major:
// the reduction index is the slow one, exactly the layout of an activation matrix X[T, n] in
// C = X^T X and of a Cholesky block row U12[nb, n]).  Planes let an fp32 matrix be fed as its
// bf16 hi/mid/lo split so tensor-core products reach fp32 accuracy.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

enum TileSet : int {
  TILES_FULL = 0,   // every 128 x BN tile of D
  TILES_UPPER = 1,  // only tiles that intersect col >= row (M == N); elements col < row untouched
  TILES_DIAG = 2,   // 128 x 128 diagonal tiles; D is [M/hd, hd, hd] packed per-head blocks
};

enum EpiOp : int {
  EPI_STORE = 0,  // D  = alpha * acc
  EPI_ADD = 1,    // D += alpha * acc   (atomic when ksplit > 1)
};

constexpr int kMaxPairs = 8;

struct GemmArgs {
  const __nv_bfloat16* A;  // plane 0 of A, [K, M] row-major
  int64_t lda;             // elements
  int64_t a_plane_stride;  // elements between planes (0 if a single plane)
  int a_planes;
  const __nv_bfloat16* B;
  int64_t ldb;
  int64_t b_plane_stride;
  int b_planes;
  int npairs;  // >= 1
  int pair_a[kMaxPairs];
  int pair_b[kMaxPairs];
  int64_t M, N, K;
  float* D;
  int64_t ldd;
  float alpha;
  int tiles;   // TileSet
  int epi;     // EpiOp
  int hd;      // TILES_DIAG: block size (divides 128)
  int ksplit;  // >= 1; > 1 requires EPI_ADD; 0 = choose automatically; < 0 = automatic, at least -ksplit
  // B is lower-triangular in (k, n): B[k, n] == 0 for n > k, so an N tile starting at column c0
  // only needs k >= c0 (used by the blocked triangular inverse).
  int klo_from_n;
  // Upper bound on the CTAs of the persistent grid (0 = one per SM).  The blocked drivers keep a
  // few SMs free so the latency-critical panel kernels of the look-ahead lane start at once.
  int max_ctas;
};

// Returns 0, a negative argument error, or -1000 - cudaError_t.
int gemm_tn_launch(const GemmArgs& a, cudaStream_t stream);

int device_sm_count();

// The tensor core aligns and TRUNCATES when it accumulates in fp32, which biases long sums of
// same-sign products (a Gram diagonal) by about 2^-24 per 16 accumulated rows.  Statistics
// therefore accumulate at most kMaxAccumRows rows in TMEM per run and add the runs in L2 with
// round-to-nearest (the reduce-add epilogue): a K segmentation, not extra work.  Measured on
// B200 at n = 11008: one 16384-row run 2.7e-5 relative on the diagonal, 4096-row runs 1.2e-5 but
// 10 % slower (8x the epilogues) — so the cap only bounds the bias of very long calls.
constexpr int64_t kMaxAccumRows = 32768;
inline int segments_for(int64_t rows) {
  return static_cast<int>((rows + kMaxAccumRows - 1) / kMaxAccumRows);
}

}  // namespace mg
