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
// potrf128: Cholesky of one 128 x 128 diagonal block, fp64, one CTA of 512 threads.
// Four 32-wide sub-panels; per sub-panel
//   (1) warp 0 factors the 32 x 32 diagonal block entirely in registers (lane k owns column k,
//       pivot row broadcast by shuffles),
//   (2) one thread per remaining column does the 32-step forward substitution,
//   (3) all threads apply the rank-32 update to the trailing part of the block.
// ---------------------------------------------------------------------------------------------
constexpr int kDiagLd = kNB + 1;
constexpr int kPotrfThreads = 512;
constexpr size_t kPotrfSmem = sizeof(double) * (kNB * kDiagLd + 32);

__global__ void __launch_bounds__(kPotrfThreads, 1)
    potrf128_kernel(float* __restrict__ A, int64_t ld, int64_t j0, int nb,
                    float* __restrict__ t_fwd, float* __restrict__ t_bwd,
                    __nv_bfloat16* __restrict__ u_planes, __nv_bfloat16* __restrict__ l_planes,
                    int64_t ld_up, int64_t up_plane_stride, int* __restrict__ info) {
  extern __shared__ double sm[];
  double* a = sm;                     // [128][129], upper triangle live
  double* invd = sm + kNB * kDiagLd;  // [32] reciprocal pivots of the current sub-panel
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;

  for (int e = t; e < kNB * kNB; e += kPotrfThreads) {
    const int i = e >> 7, k = e & 127;
    double v = (i == k) ? 1.0 : 0.0;  // identity padding for a short last block
    if (i < nb && k < nb && i <= k) v = static_cast<double>(A[(j0 + i) * ld + (j0 + k)]);
    a[i * kDiagLd + k] = v;
  }
  __syncthreads();

  const int nsub = (nb + 31) / 32;
  for (int kb = 0; kb < nsub; ++kb) {
    const int c0 = kb * 32;
    if (warp == 0) {
      double col[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) col[i] = (i <= lane) ? a[(c0 + i) * kDiagLd + c0 + lane] : 0.0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        double piv = __shfl_sync(0xffffffffu, col[j], j);
        if (!(piv > 0.0)) {
          if (lane == 0 && c0 + j < nb) atomicCAS(info, 0, static_cast<int>(j0 + c0 + j + 1));
          piv = 1e-30;
        }
        const double d = sqrt(piv);
        const double inv = 1.0 / d;
        const double ujk = (lane > j) ? col[j] * inv : (lane == j ? d : 0.0);
        col[j] = ujk;
        if (lane == j) invd[j] = inv;
#pragma unroll
        for (int i = j + 1; i < 32; ++i) {
          const double uji = __shfl_sync(0xffffffffu, ujk, i);
          col[i] -= uji * ujk;  // entries with i > lane are never read
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i <= lane) a[(c0 + i) * kDiagLd + c0 + lane] = col[i];
    }
    __syncthreads();
    const int rest = kNB - c0 - 32;  // columns right of the sub-panel (padding included)
    if (t < rest) {
      const int k = c0 + 32 + t;
      double x[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        double s = a[(c0 + i) * kDiagLd + k];
#pragma unroll
        for (int m = 0; m < i; ++m) s -= a[(c0 + m) * kDiagLd + c0 + i] * x[m];
        x[i] = s * invd[i];
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) a[(c0 + i) * kDiagLd + k] = x[i];
    }
    __syncthreads();
    // rank-32 update of the trailing upper triangle
    for (int e = t; e < rest * rest; e += kPotrfThreads) {
      const int ii = e / rest, kk = e - ii * rest;
      if (ii > kk) continue;
      const int i = c0 + 32 + ii, k = c0 + 32 + kk;
      double s = 0.0;
#pragma unroll 8
      for (int m = 0; m < 32; ++m) s += a[(c0 + m) * kDiagLd + i] * a[(c0 + m) * kDiagLd + k];
      a[i * kDiagLd + k] -= s;
    }
    __syncthreads();
  }

  // ---- outputs
  for (int e = t; e < kNB * kNB; e += kPotrfThreads) {
    const int i = e >> 7, k = e & 127;
    const bool in = i < nb && k < nb;
    const float x = (in && i <= k) ? static_cast<float>(a[i * kDiagLd + k]) : 0.f;
    if (in && i <= k) A[(j0 + i) * ld + (j0 + k)] = x;
    // forward block: T[m][i] = U[m][i]; identity beyond nb
    t_fwd[i * kTLd + k] = in ? x : (i == k ? 1.f : 0.f);
    if (in && (u_planes || l_planes)) {
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
  // backward block: T'[m'][i'] = U[nb-1-i'][nb-1-m'] for m' <= i' < nb; identity beyond
  for (int e = t; e < kNB * kNB; e += kPotrfThreads) {
    const int mp = e >> 7, ip = e & 127;
    float x = (mp == ip) ? 1.f : 0.f;
    if (mp < nb && ip < nb)
      x = (mp <= ip) ? static_cast<float>(a[(nb - 1 - ip) * kDiagLd + (nb - 1 - mp)]) : 0.f;
    t_bwd[mp * kTLd + ip] = x;
  }
}

// ---------------------------------------------------------------------------------------------
// trsm128: forward substitution with a compact lower-in-(m,i) block T (x_i = (b_i - sum_{m<i}
// T[m][i] x_m) / T[i][i]), one thread per right-hand-side column, 32-row chunks held in
// registers, earlier chunks' solutions parked in shared memory.  `reversed` maps chunk row i' to
// matrix row nb-1-i' (that is how U x = b becomes a forward substitution).
// ---------------------------------------------------------------------------------------------
constexpr int kTrsmThreads = 128;
constexpr size_t kTrsmSmem = sizeof(float) * (kTBlock + kNB * kTrsmThreads + kNB);

__global__ void __launch_bounds__(kTrsmThreads)
    trsm128_kernel(const float* __restrict__ tblock, int reversed, int nb,
                   const float* __restrict__ B, int64_t ldb, int64_t ncols, float alpha,
                   float* __restrict__ X, int64_t ldx, __nv_bfloat16* __restrict__ planes,
                   int64_t ldp, int64_t pstride, __nv_bfloat16* __restrict__ tplanes,
                   int64_t ldtp, int64_t tpstride, float* __restrict__ colsumsq) {
  extern __shared__ __align__(16) float smf[];
  float* T = smf;                        // [128][132]
  float* xs = smf + kTBlock;             // [128][kTrsmThreads]
  float* invd = xs + kNB * kTrsmThreads; // [128]
  const int tid = threadIdx.x;
  {
    const float4* src = reinterpret_cast<const float4*>(tblock);
    float4* dst = reinterpret_cast<float4*>(T);
    for (int e = tid; e < kTBlock / 4; e += kTrsmThreads) dst[e] = __ldg(src + e);
  }
  __syncthreads();
  if (tid < kNB) invd[tid] = 1.f / T[tid * kTLd + tid];
  __syncthreads();

  const int64_t c = static_cast<int64_t>(blockIdx.x) * kTrsmThreads + tid;
  const bool valid = c < ncols;
  const int nchunks = (nb + 31) / 32;
  float sumsq = 0.f;
  for (int rb = 0; rb < nchunks; ++rb) {
    float r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int ip = rb * 32 + i;
      const int row = reversed ? nb - 1 - ip : ip;
      r[i] = (valid && ip < nb) ? B[static_cast<int64_t>(row) * ldb + c] : 0.f;
    }
    for (int pb = 0; pb < rb; ++pb) {
#pragma unroll 4
      for (int m = 0; m < 32; ++m) {
        const float xm = xs[(pb * 32 + m) * kTrsmThreads + tid];
        const float4* t4 = reinterpret_cast<const float4*>(T + (pb * 32 + m) * kTLd + rb * 32);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 tv = t4[q];
          r[4 * q + 0] = fmaf(-tv.x, xm, r[4 * q + 0]);
          r[4 * q + 1] = fmaf(-tv.y, xm, r[4 * q + 1]);
          r[4 * q + 2] = fmaf(-tv.z, xm, r[4 * q + 2]);
          r[4 * q + 3] = fmaf(-tv.w, xm, r[4 * q + 3]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int ip = rb * 32 + i;
      const float x = r[i] * invd[ip];
      r[i] = x;
      xs[ip * kTrsmThreads + tid] = x;
      const float* trow = T + ip * kTLd + rb * 32;
#pragma unroll
      for (int i2 = i + 1; i2 < 32; ++i2) r[i2] = fmaf(-trow[i2], x, r[i2]);
    }
    if (!valid) continue;
    // ---- outputs of this chunk
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int ip = rb * 32 + i;
      if (ip >= nb) break;
      const int row = reversed ? nb - 1 - ip : ip;
      const float x = alpha * r[i];
      sumsq = fmaf(x, x, sumsq);
      if (X) X[static_cast<int64_t>(row) * ldx + c] = x;
      if (planes || tplanes) {
        __nv_bfloat16 h, m, l;
        split3(x, h, m, l);
        if (planes) {
          const int64_t o = static_cast<int64_t>(row) * ldp + c;
          planes[o] = h;
          planes[pstride + o] = m;
          planes[2 * pstride + o] = l;
        }
        if (tplanes) {
          const int64_t o = c * ldtp + row;
          tplanes[o] = h;
          tplanes[tpstride + o] = m;
          tplanes[2 * tpstride + o] = l;
        }
      }
    }
  }
  if (colsumsq && valid) atomicAdd(colsumsq + c, sumsq);
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

int potrf128(float* A, int64_t ld, int64_t j0, int nb, float* t_fwd, float* t_bwd,
             __nv_bfloat16* u_planes, __nv_bfloat16* l_planes, int64_t ld_up,
             int64_t up_plane_stride, int* info, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(potrf128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kPotrfSmem));
    if (e != cudaSuccess) return -1000 - static_cast<int>(e);
    attr_set = true;
  }
  potrf128_kernel<<<1, kPotrfThreads, kPotrfSmem, s>>>(A, ld, j0, nb, t_fwd, t_bwd, u_planes,
                                                       l_planes, ld_up, up_plane_stride, info);
  return cuda_rc();
}

int trsm128(const float* tblock, bool reversed, int nb, const float* B, int64_t ldb, int64_t ncols,
            float alpha, float* X, int64_t ldx, __nv_bfloat16* planes, int64_t ldp, int64_t pstride,
            __nv_bfloat16* tplanes, int64_t ldtp, int64_t tpstride, float* colsumsq,
            cudaStream_t s) {
  if (ncols <= 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(trsm128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kTrsmSmem));
    if (e != cudaSuccess) return -1000 - static_cast<int>(e);
    attr_set = true;
  }
  const unsigned grid = static_cast<unsigned>((ncols + kTrsmThreads - 1) / kTrsmThreads);
  trsm128_kernel<<<grid, kTrsmThreads, kTrsmSmem, s>>>(tblock, reversed ? 1 : 0, nb, B, ldb, ncols,
                                                       alpha, X, ldx, planes, ldp, pstride, tplanes,
                                                       ldtp, tpstride, colsumsq);
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
  int rc;
  for (int64_t j0 = 0, pj = 0; j0 < n; j0 += kNB, ++pj) {
    const int nb = static_cast<int>(n - j0 < kNB ? n - j0 : kNB);
    float* tf = ws.t_fwd + pj * kTBlock;
    float* tb = ws.t_bwd + pj * kTBlock;
    rc = potrf128(A, ld, j0, nb, tf, tb, ws.u_planes, ws.l_planes, np, pstride, info, s);
    if (rc) return rc;
    const int64_t rest = n - j0 - nb;
    if (rest <= 0) break;
    // block row: U12 = U11^-T A12 (in place) + its bf16 planes (and those of U12^T)
    float* a12 = A + j0 * ld + (j0 + nb);
    __nv_bfloat16* u12 = ws.u_planes + j0 * np + (j0 + nb);
    __nv_bfloat16* l21 = ws.l_planes ? ws.l_planes + (j0 + nb) * np + j0 : nullptr;
    rc = trsm128(tf, false, nb, a12, ld, rest, 1.f, a12, ld, u12, np, pstride, l21, np, pstride,
                 nullptr, s);
    if (rc) return rc;
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
