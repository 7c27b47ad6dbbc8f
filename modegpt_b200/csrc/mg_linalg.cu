// mg_linalg.cu — building blocks of the blocked factorizations:
//   * fp32 -> 3 x bf16 plane splitting (so tensor-core products are fp32-accurate),
//   * the 128-wide diagonal block kernel (fp64 Cholesky + triangular inverse in shared memory),
//   * the right-looking blocked Cholesky driver whose TRSM and trailing SYRK run on tcgen05.
#include "mg_linalg.cuh"

#include "mg_gemm.cuh"

namespace mg {

namespace {

inline int cuda_rc() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - static_cast<int>(e);
}

__device__ __forceinline__ void split3(float x, __nv_bfloat16& hi, __nv_bfloat16& mid,
                                       __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(hi);  // exact
  mid = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mid);  // exact
  lo = __float2bfloat16_rn(r2);
}

// ---------------------------------------------------------------------------------------------
// split kernel: 32 x 64 tile per block, 256 threads (64 columns x 4 row lanes)
// ---------------------------------------------------------------------------------------------
template <bool TRANSPOSE>
__global__ void __launch_bounds__(256) split_planes_kernel(const float* __restrict__ src,
                                                           int64_t ld_src, int64_t rows,
                                                           int64_t cols,
                                                           __nv_bfloat16* __restrict__ dst,
                                                           int64_t ld_dst, int64_t plane_stride,
                                                           float* __restrict__ colsumsq,
                                                           float diag_add) {
  __shared__ __nv_bfloat16 tile[kPlanes][64][34];
  __shared__ float csum[4][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 32;
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 64;
  const int64_t c = c0 + tx;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = ty + 4 * i;
    const int64_t r = r0 + rr;
    float x = 0.f;
    if (r < rows && c < cols) x = src[r * ld_src + c] + (r == c ? diag_add : 0.f);
    acc = fmaf(x, x, acc);
    __nv_bfloat16 h, m, l;
    split3(x, h, m, l);
    if (TRANSPOSE) {
      tile[0][tx][rr] = h;
      tile[1][tx][rr] = m;
      tile[2][tx][rr] = l;
    } else if (r < rows && c < cols) {
      dst[r * ld_dst + c] = h;
      dst[plane_stride + r * ld_dst + c] = m;
      dst[2 * plane_stride + r * ld_dst + c] = l;
    }
  }
  if (colsumsq) {
    csum[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < cols)
      atomicAdd(colsumsq + c, csum[0][tx] + csum[1][tx] + csum[2][tx] + csum[3][tx]);
  }
  if (TRANSPOSE) {
    __syncthreads();
    // dst row = source column; 32 source rows are contiguous in dst
    const int wr = threadIdx.x & 31, wc0 = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int cc = wc0 + 8 * i;
      const int64_t dr = c0 + cc, dc = r0 + wr;
      if (dr < cols && dc < rows) {
#pragma unroll
        for (int p = 0; p < kPlanes; ++p) dst[p * plane_stride + dr * ld_dst + dc] = tile[p][cc][wr];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// diagonal block: fp64 Cholesky (upper) + in-place triangular inverse, one CTA of 1024 threads
// ---------------------------------------------------------------------------------------------
constexpr int kDiagLd = kNB + 1;
constexpr size_t kDiagSmem = sizeof(double) * (kNB * kDiagLd + 2 * kNB);

__global__ void __launch_bounds__(1024, 1)
    diag_block_kernel(float* __restrict__ A, int64_t ld, int64_t j0, int nb,
                      __nv_bfloat16* __restrict__ w_planes, __nv_bfloat16* __restrict__ wt_planes,
                      __nv_bfloat16* __restrict__ u_planes, __nv_bfloat16* __restrict__ l_planes,
                      int64_t ld_up, int64_t up_plane_stride, int* __restrict__ info) {
  extern __shared__ double sm[];
  double* a = sm;                       // [128][129]
  double* u = sm + kNB * kDiagLd;       // [128] scaled pivot row / copied column
  const int t = threadIdx.x;

  // load the upper triangle; pad a short last block with the identity
  for (int e = t; e < kNB * kNB; e += 1024) {
    const int i = e >> 7, k = e & 127;
    double v = (i == k) ? 1.0 : 0.0;
    if (i < nb && k < nb && i <= k) v = static_cast<double>(A[(j0 + i) * ld + (j0 + k)]);
    a[i * kDiagLd + k] = v;
  }

  // ---- Cholesky, right-looking: a = U^T U, U upper
  const int ti = t >> 5, tk = t & 31;
  for (int j = 0; j < nb; ++j) {
    __syncthreads();
    double piv = a[j * kDiagLd + j];
    if (!(piv > 0.0)) {
      if (t == 0) atomicCAS(info, 0, static_cast<int>(j0 + j + 1));
      piv = 1e-30;
    }
    const double d = sqrt(piv);
    const double inv = 1.0 / d;
    if (t < kNB) u[t] = (t > j && t < nb) ? a[j * kDiagLd + t] * inv : 0.0;
    __syncthreads();
    if (t < kNB) {
      if (t > j && t < nb) a[j * kDiagLd + t] = u[t];
      if (t == j) a[j * kDiagLd + j] = d;
    }
    for (int i = j + 1 + ti; i < nb; i += 32) {
      const double ui = u[i];
      for (int k = i + tk; k < nb; k += 32) a[i * kDiagLd + k] -= ui * u[k];
    }
  }
  __syncthreads();

  // ---- write U11 (fp32, upper triangle only) and its planes
  for (int e = t; e < kNB * kNB; e += 1024) {
    const int i = e >> 7, k = e & 127;
    if (i < nb && k < nb) {
      const float x = (i <= k) ? static_cast<float>(a[i * kDiagLd + k]) : 0.f;
      if (i <= k) A[(j0 + i) * ld + (j0 + k)] = x;
      if (u_planes || l_planes) {
        __nv_bfloat16 h, m, l;
        split3(x, h, m, l);
        if (u_planes) {
          const int64_t o = (j0 + i) * ld_up + (j0 + k);
          u_planes[o] = h;
          u_planes[up_plane_stride + o] = m;
          u_planes[2 * up_plane_stride + o] = l;
        }
        if (l_planes) {
          const int64_t o = (j0 + k) * ld_up + (j0 + i);
          l_planes[o] = h;
          l_planes[up_plane_stride + o] = m;
          l_planes[2 * up_plane_stride + o] = l;
        }
      }
    }
  }
  __syncthreads();

  // ---- in-place inverse of the upper-triangular factor, column by column:
  //      W[0:j, j] = -W[0:j, 0:j] * U[0:j, j] / U[j, j];  8 threads share each row's dot product
  const int row = t >> 3, part = t & 7;
  for (int j = 0; j < nb; ++j) {
    if (t < j) u[t] = a[t * kDiagLd + j];
    __syncthreads();
    const double wjj = 1.0 / a[j * kDiagLd + j];
    double s = 0.0;
    if (row < j)
      for (int k = row + part; k < j; k += 8) s += a[row * kDiagLd + k] * u[k];
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    __syncthreads();
    if (part == 0 && row < j) a[row * kDiagLd + j] = -s * wjj;
    if (t == 0) a[j * kDiagLd + j] = wjj;
  }
  __syncthreads();

  // ---- emit W and W^T planes ([128 x 128], zero outside the upper triangle / beyond nb)
  for (int e = t; e < kNB * kNB; e += 1024) {
    const int i = e >> 7, k = e & 127;
    float x = 0.f;
    if (i < nb && k < nb && i <= k) x = static_cast<float>(a[i * kDiagLd + k]);
    __nv_bfloat16 h, m, l;
    split3(x, h, m, l);
    w_planes[e] = h;
    w_planes[kNB * kNB + e] = m;
    w_planes[2 * kNB * kNB + e] = l;
    const int et = k * kNB + i;
    wt_planes[et] = h;
    wt_planes[kNB * kNB + et] = m;
    wt_planes[2 * kNB * kNB + et] = l;
  }
}

}  // namespace

int split_planes(const float* src, int64_t ld_src, int64_t rows, int64_t cols, __nv_bfloat16* dst,
                 int64_t ld_dst, int64_t plane_stride, bool transpose, float* colsumsq,
                 cudaStream_t s, float diag_add) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid(static_cast<unsigned>((cols + 63) / 64), static_cast<unsigned>((rows + 31) / 32));
  if (transpose)
    split_planes_kernel<true><<<grid, 256, 0, s>>>(src, ld_src, rows, cols, dst, ld_dst,
                                                   plane_stride, colsumsq, diag_add);
  else
    split_planes_kernel<false><<<grid, 256, 0, s>>>(src, ld_src, rows, cols, dst, ld_dst,
                                                    plane_stride, colsumsq, diag_add);
  return cuda_rc();
}

int diag_block_factor(float* A, int64_t ld, int64_t j0, int nb, __nv_bfloat16* w_planes,
                      __nv_bfloat16* wt_planes, __nv_bfloat16* u_planes, __nv_bfloat16* l_planes,
                      int64_t ld_up, int64_t up_plane_stride, int* info, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(diag_block_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kDiagSmem));
    if (e != cudaSuccess) return -1000 - static_cast<int>(e);
    attr_set = true;
  }
  diag_block_kernel<<<1, 1024, kDiagSmem, s>>>(A, ld, j0, nb, w_planes, wt_planes, u_planes,
                                               l_planes, ld_up, up_plane_stride, info);
  return cuda_rc();
}

namespace {
void set_pairs6(GemmArgs& g) {
  static const int pa[6] = {0, 0, 1, 1, 0, 2};
  static const int pb[6] = {0, 1, 0, 1, 2, 0};
  g.npairs = 6;
  for (int i = 0; i < 6; ++i) {
    g.pair_a[i] = pa[i];
    g.pair_b[i] = pb[i];
  }
}
}  // namespace

int cholesky_upper(float* A, int64_t n, int64_t ld, const CholWorkspace& ws, int* info,
                   cudaStream_t s) {
  const int64_t np = ws.n_pad;
  const int64_t pstride = np * np;
  const int64_t wstride = static_cast<int64_t>(kPlanes) * kNB * kNB;
  int rc;
  for (int64_t j0 = 0, pj = 0; j0 < n; j0 += kNB, ++pj) {
    const int nb = static_cast<int>(n - j0 < kNB ? n - j0 : kNB);
    __nv_bfloat16* wj = ws.w_planes + pj * wstride;
    __nv_bfloat16* wtj = ws.wt_planes + pj * wstride;
    rc = diag_block_factor(A, ld, j0, nb, wj, wtj, ws.u_planes, ws.l_planes, np, pstride, info, s);
    if (rc) return rc;
    const int64_t rest = n - j0 - nb;
    if (rest <= 0) break;
    float* a12 = A + j0 * ld + (j0 + nb);
    // TRSM as a GEMM with the inverted diagonal block: U12 = W^T A12
    rc = split_planes(a12, ld, nb, rest, ws.row_planes, np, kNB * np, false, nullptr, s);
    if (rc) return rc;
    GemmArgs g{};
    g.A = wj;
    g.lda = kNB;
    g.a_plane_stride = kNB * kNB;
    g.a_planes = kPlanes;
    g.B = ws.row_planes;
    g.ldb = np;
    g.b_plane_stride = kNB * np;
    g.b_planes = kPlanes;
    set_pairs6(g);
    g.M = nb;
    g.N = rest;
    g.K = nb;
    g.D = a12;
    g.ldd = ld;
    g.alpha = 1.f;
    g.tiles = TILES_FULL;
    g.epi = EPI_STORE;
    g.ksplit = 1;
    rc = gemm_tn_launch(g, s);
    if (rc) return rc;
    // planes of the finished block row (and of its transpose when the caller wants L = U^T)
    __nv_bfloat16* u12 = ws.u_planes + j0 * np + (j0 + nb);
    rc = split_planes(a12, ld, nb, rest, u12, np, pstride, false, nullptr, s);
    if (rc) return rc;
    if (ws.l_planes) {
      rc = split_planes(a12, ld, nb, rest, ws.l_planes + (j0 + nb) * np + j0, np, pstride, true,
                        nullptr, s);
      if (rc) return rc;
    }
    // trailing update on the upper triangle: A22 -= U12^T U12
    GemmArgs t{};
    t.A = t.B = u12;
    t.lda = t.ldb = np;
    t.a_plane_stride = t.b_plane_stride = pstride;
    t.a_planes = t.b_planes = kPlanes;
    set_pairs6(t);
    t.M = t.N = rest;
    t.K = nb;
    t.D = A + (j0 + nb) * ld + (j0 + nb);
    t.ldd = ld;
    t.alpha = -1.f;
    t.tiles = TILES_UPPER;
    t.epi = EPI_ADD;
    t.ksplit = 1;
    rc = gemm_tn_launch(t, s);
    if (rc) return rc;
  }
  return 0;
}

}  // namespace mg
