// mg_type1.cu — type-I (Nystrom MLP) path, SURVEY §8 rows a7-a8:
//   ridge-leverage scores diag((C + lambda I)^-1)      src/compression/compress_mlp.py:13-25
//   k-smallest selection, ascending indices             src/compression/compress_mlp.py:45-47
//   row gathers of W_up / W_gate                         src/compression/compress_mlp.py:49-50
//   W_down' = (C_kk + 1e-6 I)^-1 C_k: W_down^T          src/compression/compress_mlp.py:52-57
// Every O(n^3) step is a blocked algorithm whose block operations run on the tcgen05 engine with
// fp32 operands split into three bf16 planes; the 128-wide diagonal blocks are done in fp64.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/modegpt_b200.h"
#include "mg_gemm.cuh"
#include "mg_linalg.cuh"
#include "mg_prof.cuh"

namespace {

using mg::kNB;
using mg::kPlanes;
using bf16 = __nv_bfloat16;

inline int cuda_rc() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - static_cast<int>(e);
}

// ------------------------------------------------------------------------------- workspace carving
struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<uint8_t*>(p)) {}
  template <class T>
  T* take(size_t count) {
    off = (off + 255) & ~static_cast<size_t>(255);
    T* r = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return r;
  }
};

struct ScoresWs {
  float* a;          // [n x np] working copy C + ridge I -> U
  float* tt;         // [2][128 x np]
  float* ident;      // [128 x 128] identity (right-hand side for the diagonal blocks of U^-T)
  bf16* y_planes;    // [3][np x np]
  mg::CholWorkspace chol;
  size_t bytes;
};

ScoresWs carve_scores(void* p, int64_t n) {
  const int64_t np = mg::round_up(n, 64);
  const int64_t panels = (n + kNB - 1) / kNB;
  Carver c(p);
  ScoresWs w{};
  w.a = c.take<float>(n * np);
  w.tt = c.take<float>(2 * kNB * np);   // double-buffered (look-ahead on the tri lanes)
  w.ident = c.take<float>(kNB * kNB);
  w.y_planes = c.take<bf16>(kPlanes * np * np);
  w.chol.u_planes = c.take<bf16>(kPlanes * np * np);
  w.chol.l_planes = nullptr;
  w.chol.t_fwd = c.take<float>(panels * mg::kTBlock);
  w.chol.t_bwd = nullptr;   // no back-substitution on this path
  w.chol.n_pad = np;
  w.bytes = c.off + 256;
  return w;
}

struct NystromWs {
  bf16* g_planes;   // [3][n x kp]     planes of C[:, idx]
  bf16* wdt;        // [n x dp]        W_down^T
  float* rhs;       // [k x dp]        cross term -> Z -> X (in place)
  float* ckk;       // [k x kp]
  bf16* z_planes;   // [3][128 x dp]      X planes of one panel (single-stream back-substitution)
  bf16* zb_planes;  // [3][4*128 x dp]    forward substitution: Z planes of one outer block
  bf16* xb_planes;  // [2][3][4*128 x dp] back-substitution: X planes of an outer block (double-buffered)
  uint8_t* kept;    // [n]                1 where the channel is in idx
  int32_t* pos;     // [n]                position of channel l inside idx, -1 if dropped
  float* xacc;      // [k x dp]           the full solution W_down'^T (fp32), refined in place
  mg::CholWorkspace chol;
  size_t bytes;
};

NystromWs carve_nystrom(void* p, int64_t n, int64_t k, int64_t d) {
  const int64_t kp = mg::round_up(k, 64), dp = mg::round_up(d, 64);
  const int64_t panels = (k + kNB - 1) / kNB;
  Carver c(p);
  NystromWs w{};
  w.g_planes = c.take<bf16>(kPlanes * n * kp);
  w.wdt = c.take<bf16>(n * dp);
  w.rhs = c.take<float>(k * dp);
  w.ckk = c.take<float>(k * kp);
  w.z_planes = c.take<bf16>(kPlanes * kNB * dp);        // single-stream path only
  w.zb_planes = c.take<bf16>(kPlanes * 4 * kNB * dp);
  w.xb_planes = c.take<bf16>(2 * kPlanes * 4 * kNB * dp);
  w.kept = c.take<uint8_t>(n);
  w.pos = c.take<int32_t>(n);
  w.xacc = c.take<float>(k * dp);
  w.chol.u_planes = c.take<bf16>(kPlanes * kp * kp);
  w.chol.l_planes = c.take<bf16>(kPlanes * kp * kp);
  w.chol.t_fwd = c.take<float>(panels * mg::kTBlock);
  w.chol.t_bwd = c.take<float>(panels * mg::kTBlock);
  w.chol.n_pad = kp;
  w.bytes = c.off + 256;
  return w;
}

void pairs6(mg::GemmArgs& g) {
  static const int pa[6] = {0, 0, 1, 1, 0, 2};
  static const int pb[6] = {0, 1, 0, 1, 2, 0};
  g.npairs = 6;
  for (int i = 0; i < 6; ++i) {
    g.pair_a[i] = pa[i];
    g.pair_b[i] = pb[i];
  }
}

// ------------------------------------------------------------------------------- small kernels
__device__ __forceinline__ void split3(float x, bf16& hi, bf16& mid, bf16& lo) {
  hi = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(hi);
  mid = __float2bfloat16_rn(r1);
  lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
}

// dst (upper triangle) = src + ridge * I
__global__ void copy_ridge_kernel(const float* __restrict__ src, int64_t lds,
                                  float* __restrict__ dst, int64_t ldd, int64_t n, float ridge) {
  const int64_t r = blockIdx.y;
  for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < n;
       c += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (c >= r) dst[r * ldd + c] = src[r * lds + c] + (c == r ? ridge : 0.f);
  }
}

__global__ void identity_kernel(float* __restrict__ m, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n * n) m[e] = (e / n == e % n) ? 1.f : 0.f;
}

// out row (q*r + t) = W[q*hd + mask[(q/group)*r + t], :]   (plain row gather: hd = n rows, 1 head)
__global__ void __launch_bounds__(256) gather_rows_kernel(const bf16* __restrict__ W, int64_t ldw,
                                                          const int64_t* __restrict__ mask,
                                                          int group, int64_t hd, int64_t r,
                                                          int64_t d, bf16* __restrict__ out,
                                                          int64_t ldo, int vec) {
  const int64_t orow = blockIdx.x;
  const int64_t q = orow / r, t = orow - q * r;
  const int64_t srow = q * hd + mask[(q / group) * r + t];
  const bf16* s = W + srow * ldw;
  bf16* o = out + orow * ldo;
  if (vec) {
    const uint4* s4 = reinterpret_cast<const uint4*>(s);
    uint4* o4 = reinterpret_cast<uint4*>(o);
    for (int64_t i = threadIdx.x; i < (d >> 3); i += blockDim.x) o4[i] = __ldg(s4 + i);
  } else {
    for (int64_t i = threadIdx.x; i < d; i += blockDim.x) o[i] = s[i];
  }
}

// out[i, j] = C[idx_i, idx_j] + (i == j) * jitter, upper triangle only
__global__ void __launch_bounds__(256) gather_sym_kernel(const float* __restrict__ C, int64_t ldc,
                                                         const int64_t* __restrict__ idx,
                                                         int64_t k, float* __restrict__ out,
                                                         int64_t ldo, float jitter) {
  const int64_t i = blockIdx.y;
  const float* row = C + idx[i] * ldc;
  for (int64_t j = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; j < k;
       j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (j >= i) out[i * ldo + j] = __ldg(row + idx[j]) + (i == j ? jitter : 0.f);
  }
}

__global__ void mark_kept_kernel(const int64_t* __restrict__ idx, int64_t k, uint8_t* __restrict__ kept) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < k) kept[idx[i]] = 1;
}

// planes_p[l, i] = p-th plane of C[l, idx_i]   for l in [0, n), i in [0, k); rows l that are
// themselves kept are written as ZERO (the correction form of the Nystrom solve, see the driver)
__global__ void __launch_bounds__(256) gather_cols_planes_kernel(const float* __restrict__ C,
                                                                 int64_t ldc, int64_t n,
                                                                 const int64_t* __restrict__ idx,
                                                                 int64_t k,
                                                                 const uint8_t* __restrict__ kept,
                                                                 bf16* __restrict__ planes,
                                                                 int64_t ldp, int64_t pstride) {
  const int64_t l = blockIdx.y;
  const float* row = C + l * ldc;
  const bool skip = kept[l] != 0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < k;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    bf16 h, m, lo;
    split3(skip ? 0.f : __ldg(row + idx[i]), h, m, lo);
    const int64_t o = l * ldp + i;
    planes[o] = h;
    planes[pstride + o] = m;
    planes[2 * pstride + o] = lo;
  }
}

// out[c, r] = in[r, c]; IN is bf16 or float, OUT is bf16
template <class IN>
__global__ void __launch_bounds__(256) transpose_to_bf16_kernel(const IN* __restrict__ in,
                                                                int64_t ld_in, int64_t rows,
                                                                int64_t cols,
                                                                bf16* __restrict__ out,
                                                                int64_t ld_out) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 32, c0 = static_cast<int64_t>(blockIdx.x) * 32;
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? static_cast<float>(in[r * ld_in + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t orow = c0 + i, ocol = r0 + tx;
    if (orow < cols && ocol < rows) out[orow * ld_out + ocol] = __float2bfloat16_rn(tile[tx][i]);
  }
}

// rhs[i, c] -= jitter * Wd^T[idx_i, c]   (wdt = W_down^T, bf16 [n x ldw])
__global__ void __launch_bounds__(256) sub_jitter_rows_kernel(float* __restrict__ rhs, int64_t ldr,
                                                              const bf16* __restrict__ wdt, int64_t ldw,
                                                              const int64_t* __restrict__ idx, int64_t d,
                                                              float jitter) {
  const int64_t i = blockIdx.x;
  const bf16* src = wdt + idx[i] * ldw;
  float* dst = rhs + i * ldr;
  for (int64_t c = threadIdx.x; c < d; c += blockDim.x)
    dst[c] = fmaf(-jitter, __bfloat162float(src[c]), dst[c]);
}

// out[c, i] = bf16(X[i, c] + Wd^T[idx_i, c]):  W_down' = W_down[:, idx] + correction, transposed store
__global__ void __launch_bounds__(256) add_kept_transpose_kernel(const float* __restrict__ x, int64_t ldx,
                                                                 const bf16* __restrict__ wdt, int64_t ldw,
                                                                 const int64_t* __restrict__ idx,
                                                                 int64_t k, int64_t d,
                                                                 bf16* __restrict__ out, int64_t ld_out) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 32, c0 = static_cast<int64_t>(blockIdx.x) * 32;
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + tx;      // r: kept-channel index i, c: output feature
    tile[i][tx] = (r < k && c < d) ? x[r * ldx + c] + __bfloat162float(wdt[idx[r] * ldw + c]) : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t orow = c0 + i, ocol = r0 + tx;
    if (orow < d && ocol < k) out[orow * ld_out + ocol] = __float2bfloat16_rn(tile[tx][i]);
  }
}

// kept[l] = 1 and pos[l] = i for l = idx_i (pos = -1 elsewhere: cleared by the caller)
__global__ void mark_kept_pos_kernel(const int64_t* __restrict__ idx, int64_t k, uint8_t* __restrict__ kept,
                                     int32_t* __restrict__ pos) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < k) {
    kept[idx[i]] = 1;
    pos[idx[i]] = static_cast<int32_t>(i);
  }
}

__global__ void set_float_kernel(float* p, float v) { *p = v; }

// *out = min(*out, min_i U_ii^2 / (C[idx_i, idx_i] + jitter)); positive floats order like their bits
__global__ void __launch_bounds__(256) min_rel_pivot_kernel(const float* __restrict__ C, int64_t ldc,
                                                            const int64_t* __restrict__ idx, int64_t k,
                                                            const float* __restrict__ U, int64_t ldu,
                                                            float jitter, float* __restrict__ out) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  float r = 3.0e38f;
  if (i < k) {
    const float u = U[i * ldu + i];
    const float a = C[idx[i] * ldc + idx[i]] + jitter;
    r = a > 0.f ? u * u / a : 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r = fminf(r, __shfl_xor_sync(0xffffffffu, r, o));
  if ((threadIdx.x & 31) == 0) atomicMin(reinterpret_cast<int*>(out), __float_as_int(fmaxf(r, 0.f)));
}

// xacc[i, c] = X[i, c] + Wd^T[idx_i, c]   (the full solution W_down'^T = W_down[:, idx]^T + correction)
__global__ void __launch_bounds__(256) add_kept_rows_kernel(const float* __restrict__ x, int64_t ldx,
                                                            const bf16* __restrict__ wdt, int64_t ldw,
                                                            const int64_t* __restrict__ idx, int64_t d,
                                                            float* __restrict__ xacc, int accumulate) {
  const int64_t i = blockIdx.x;
  const bf16* src = wdt + idx[i] * ldw;
  for (int64_t c = threadIdx.x; c < d; c += blockDim.x) {
    const float base = accumulate ? xacc[i * ldx + c] : __bfloat162float(src[c]);
    xacc[i * ldx + c] = base + x[i * ldx + c];
  }
}

// FP64 residual of the Nystrom system for the current solution X (fp32, xacc [k x ldx]):
//     R = C[idx, :] W_d^T - (C_kk + jitter I) X  =  C[idx, :] E - jitter X,
//     E[l, c] = W_d^T[l, c] - (l kept ? X[pos_l, c] : 0)
// — one GEMM over ALL n channels whose kept rows carry the (small) difference W_d^T - X formed
// exactly in fp64, so the cancellation that defeats an fp32 right-hand side never happens.  C is
// fp32, W_d bf16, X fp32: every product is exact in fp64 and the accumulation is fp64 (CUDA cores,
// 8 x 8 register tiles; 2 k n d flop).  The result is rounded once, to the fp32 right-hand side of
// the correction solve.
constexpr int kResKC = 16;
__global__ void __launch_bounds__(256)
    nystrom_resid64_kernel(const float* __restrict__ C, int64_t ldc, int64_t n,
                           const int64_t* __restrict__ idx, int64_t k, const bf16* __restrict__ wdt,
                           int64_t ldw, const int32_t* __restrict__ pos, const float* __restrict__ xacc,
                           int64_t ldx, int64_t d, float jitter, float* __restrict__ rhs) {
  __shared__ double As[kResKC][128 + 1];
  __shared__ double Es[kResKC][128 + 1];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int64_t i0 = static_cast<int64_t>(blockIdx.x) * 128, c0 = static_cast<int64_t>(blockIdx.y) * 128;
  double acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.0;
  for (int64_t l0 = 0; l0 < n; l0 += kResKC) {
    __syncthreads();
    // As[ll][ii] = C[idx[i0 + ii], l0 + ll]: consecutive threads walk l (contiguous in C's row)
    for (int e = t; e < 128 * kResKC; e += 256) {
      const int ll = e & (kResKC - 1), ii = e / kResKC;
      const int64_t i = i0 + ii, l = l0 + ll;
      As[ll][ii] = (i < k && l < n) ? static_cast<double>(C[idx[i] * ldc + l]) : 0.0;
    }
    // Es[ll][cc] = E[l0 + ll, c0 + cc]: consecutive threads walk c (contiguous in wdt and in xacc)
    for (int e = t; e < 128 * kResKC; e += 256) {
      const int cc = e & 127, ll = e >> 7;
      const int64_t c = c0 + cc, l = l0 + ll;
      double v = 0.0;
      if (c < d && l < n) {
        v = static_cast<double>(__bfloat162float(wdt[l * ldw + c]));
        const int32_t pl = pos[l];
        if (pl >= 0) v -= static_cast<double>(xacc[static_cast<int64_t>(pl) * ldx + c]);
      }
      Es[ll][cc] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int ll = 0; ll < kResKC; ++ll) {
      double av[8], ev[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) av[a] = As[ll][ty + 16 * a];
#pragma unroll
      for (int b = 0; b < 8; ++b) ev[b] = Es[ll][tx + 16 * b];
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fma(av[a], ev[b], acc[a][b]);
    }
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int64_t i = i0 + ty + 16 * a;
    if (i >= k) continue;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
      const int64_t c = c0 + tx + 16 * b;
      if (c < d)
        rhs[i * ldx + c] = static_cast<float>(acc[a][b] - static_cast<double>(jitter) *
                                                               static_cast<double>(xacc[i * ldx + c]));
    }
  }
}

// ------------------------------------------------------------------------------- radix select
__device__ __forceinline__ uint32_t order_key(float x, int largest) {
  uint32_t u = __float_as_uint(x);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending float order -> ascending uint
  return largest ? ~u : u;
}

// idx_out[0..k) = indices of the k smallest (largest) scores, in ascending INDEX order; ties at the
// threshold go to the lower indices.  Single CTA: n is a vector length (<= a few 1e5).
__global__ void __launch_bounds__(1024, 1) select_k_kernel(const float* __restrict__ scores,
                                                           int64_t n, int64_t k, int largest,
                                                           int64_t* __restrict__ idx_out) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_mask;
  __shared__ int64_t s_need;
  __shared__ uint32_t warp_tot[32];
  __shared__ int64_t base_less, base_eq;
  const int t = threadIdx.x;
  if (t == 0) {
    s_prefix = 0;
    s_mask = 0;
    s_need = k;
    base_less = 0;
    base_eq = 0;
  }
  for (int shift = 24; shift >= 0; shift -= 8) {
    if (t < 256) hist[t] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix, mask = s_mask;
    for (int64_t i = t; i < n; i += 1024) {
      const uint32_t u = order_key(scores[i], largest);
      if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (t == 0) {
      int64_t need = s_need, cum = 0;
      int b = 0;
      for (; b < 255; ++b) {
        if (cum + hist[b] >= need) break;
        cum += hist[b];
      }
      s_need = need - cum;
      s_prefix = prefix | (static_cast<uint32_t>(b) << shift);
      s_mask = mask | (255u << shift);
    }
    __syncthreads();
  }
  const uint32_t thr = s_prefix;
  const int64_t need_eq = s_need;
  const int lane = t & 31, warp = t >> 5;
  for (int64_t c0 = 0; c0 < n; c0 += 1024) {
    const int64_t i = c0 + t;
    uint32_t less = 0, eq = 0;
    if (i < n) {
      const uint32_t u = order_key(scores[i], largest);
      less = u < thr;
      eq = u == thr;
    }
    uint32_t v = less | (eq << 16);  // both counts fit in 16 bits per 1024-chunk
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      uint32_t w = warp_tot[lane];
      uint32_t winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += y;
      }
      warp_tot[lane] = winc - w;  // exclusive
    }
    __syncthreads();
    const uint32_t excl = inc - v + warp_tot[warp];
    const int64_t less_before = base_less + (excl & 0xFFFFu);
    const int64_t eq_before = base_eq + (excl >> 16);
    const bool sel = less || (eq && eq_before < need_eq);
    if (sel) idx_out[less_before + (eq_before < need_eq ? eq_before : need_eq)] = i;
    __syncthreads();
    if (t == 1023) {
      const uint32_t tot = excl + v;
      base_less += tot & 0xFFFFu;
      base_eq += tot >> 16;
    }
    __syncthreads();
  }
}

// Forward + backward substitution of w.rhs [k x dp] with the factor in w.chol / w.ckk, in place
// (rhs -> Z -> X).  With `chol` the factorisation steps ride along (first solve: step pi needs
// block row pi of U and nothing later); without it the factor is complete (refinement solves).
int nystrom_solves(const NystromWs& w, const mg::Lanes& L, int64_t k, int64_t d, const mg::CholStepper* chol) {
  const int64_t kp = w.chol.n_pad, dp = mg::round_up(d, 64);
  const int64_t pstride = kp * kp;
  const int64_t panels = (k + kNB - 1) / kNB;
  int rc;
  // ---- forward solve  U^T Z = rhs  (right-looking, in place) rides one panel behind the
  //      factorisation on the tri lane: step pi needs block row pi of U and nothing later
  //      Two-level blocking like the factorisation: inside an outer block of kFwdOuter panels a
  //      panel only updates the block's remaining rows (K = 128); the rows below the block get one
  //      update per block with K = kFwdOuter * 128 from the block's stacked Z planes — a quarter
  //      of the read-modify-write passes over rhs.
  constexpr int64_t kFwdOuter = 4;
  const int64_t zb_stride = kFwdOuter * kNB * dp;      // plane stride of the stacked Z block
  for (int64_t pi = 0; pi < panels; ++pi) {
    if (chol) {
      if ((rc = chol->step(pi))) return rc;
      L.wait(L.tri, L.trsm);
    }
    const int64_t i0 = pi * kNB;
    const int nb = static_cast<int>(k - i0 < kNB ? k - i0 : kNB);
    float* bi = w.rhs + i0 * dp;
    const int64_t rest = k - i0 - nb;
    const int64_t q = pi % kFwdOuter, o0 = (pi - q) * kNB;
    const int64_t o_end = (o0 + kFwdOuter * kNB < k) ? o0 + kFwdOuter * kNB : k;
    bf16* zq = w.zb_planes + q * kNB * dp;              // this panel's rows inside the Z block
    // Z_i = U_ii^-T B_i (+ planes of Z_i for the updates below)
    MG_TIMED(L.tri, "nystrom.fwd_trsm",
             rc = mg::trsm128(w.chol.t_fwd + pi * mg::kTBlock, false, nb, bi, dp, d, 1.f, bi, dp,
                              rest > 0 ? zq : nullptr, dp, zb_stride, nullptr, 0, 0, nullptr, L.tri));
    if (rc) return rc;
    if (rest <= 0) break;
    mg::GemmArgs t{};
    t.lda = kp;
    t.a_plane_stride = pstride;
    t.a_planes = kPlanes;
    t.ldb = dp;
    t.b_plane_stride = zb_stride;
    t.b_planes = kPlanes;
    pairs6(t);
    t.N = d;
    t.ldd = dp;
    t.alpha = -1.f;
    t.tiles = mg::TILES_FULL;
    t.epi = mg::EPI_ADD;
    t.ksplit = 1;
    t.max_ctas = L.bulk_cta_cap();
    const int64_t in_block = o_end - (i0 + nb);
    if (in_block > 0) {            // B[block rows below] -= U[ib, those rows]^T Z_i
      t.A = w.chol.u_planes + i0 * kp + (i0 + nb);
      t.B = zq;
      t.M = in_block;
      t.K = nb;
      t.D = w.rhs + (i0 + nb) * dp;
      MG_TIMED(L.tri, "nystrom.fwd_update_in_block", rc = mg::gemm_tn_launch(t, L.tri));
    } else {                       // B[rows below the block] -= U[block rows, those rows]^T Z_block
      t.A = w.chol.u_planes + o0 * kp + o_end;
      t.B = w.zb_planes;
      t.M = rest;
      t.K = o_end - o0;
      t.D = w.rhs + o_end * dp;
      MG_TIMED(L.tri, "nystrom.fwd_update", rc = mg::gemm_tn_launch(t, L.tri));
    }
    if (rc) return rc;
  }
  // ---- backward solve  U X = Z  on the chain lane, once the forward pass and every trailing
  //      update have landed
  L.record(L.misc[0], L.tri);
  L.wait(L.chain, L.misc[0]);
  L.record(L.misc[1], L.upd);
  L.wait(L.chain, L.misc[1]);
  L.record(L.row_rest[0], L.chain2);
  L.wait(L.chain, L.row_rest[0]);
  //      Two-level blocking + look-ahead, mirrored from the factorisation (panels run from the
  //      last to the first): inside an outer block a panel only updates the block's rows above it
  //      (K = 128, chain lane); when the block's first panel is solved, the rows above the block
  //      get the whole block at once (K = 512) — the next block's last block row on the chain
  //      lane, its other rows and everything above on the upd lane.
  if (L.serial) {
    for (int64_t pi = panels - 1; pi >= 0; --pi) {
      const int64_t i0 = pi * kNB;
      const int nb = static_cast<int>(k - i0 < kNB ? k - i0 : kNB);
      float* zi = w.rhs + i0 * dp;
      MG_TIMED(L.chain, "nystrom.bwd_trsm",
               rc = mg::trsm128(w.chol.t_bwd + pi * mg::kTBlock, true, nb, zi, dp, d, 1.f, zi, dp,
                                i0 > 0 ? w.z_planes : nullptr, dp, kNB * dp, nullptr, 0, 0, nullptr,
                                L.chain));
      if (rc) return rc;
      if (i0 == 0) break;
      mg::GemmArgs t{};
      t.A = w.chol.l_planes + i0 * kp;   // Z[0:i0] -= U[0:i0, ib] X_i, A[k, m] = L[i0 + k, m]
      t.lda = kp;
      t.a_plane_stride = pstride;
      t.a_planes = kPlanes;
      t.B = w.z_planes;
      t.ldb = dp;
      t.b_plane_stride = kNB * dp;
      t.b_planes = kPlanes;
      pairs6(t);
      t.M = i0;
      t.N = d;
      t.K = nb;
      t.D = w.rhs;
      t.ldd = dp;
      t.alpha = -1.f;
      t.tiles = mg::TILES_FULL;
      t.epi = mg::EPI_ADD;
      t.ksplit = 1;
      MG_TIMED(L.chain, "nystrom.bwd_update", rc = mg::gemm_tn_launch(t, L.chain));
      if (rc) return rc;
    }
  } else {
    const int64_t nblocks = (panels + kFwdOuter - 1) / kFwdOuter;
    for (int64_t pi = panels - 1; pi >= 0; --pi) {
      const int64_t i0 = pi * kNB;
      const int nb = static_cast<int>(k - i0 < kNB ? k - i0 : kNB);
      const int64_t ob = pi / kFwdOuter, o0 = ob * kFwdOuter * kNB;
      const int64_t o_end = (o0 + kFwdOuter * kNB < k) ? o0 + kFwdOuter * kNB : k;
      // X planes of the block, stacked by row (double-buffered by block parity: the upd lane may
      // still read block ob while the chain lane fills block ob-1)
      bf16* xb = w.xb_planes + (ob & 1) * kPlanes * zb_stride;
      bf16* xq = xb + (i0 - o0) * dp;
      float* zi = w.rhs + i0 * dp;
      MG_TIMED(L.chain, "nystrom.bwd_trsm",
               rc = mg::trsm128(w.chol.t_bwd + pi * mg::kTBlock, true, nb, zi, dp, d, 1.f, zi, dp,
                                i0 > 0 ? xq : nullptr, dp, zb_stride, nullptr, 0, 0, nullptr, L.chain));
      if (rc) return rc;
      if (i0 == 0) break;
      mg::GemmArgs t{};
      t.lda = kp;
      t.a_plane_stride = pstride;
      t.a_planes = kPlanes;
      t.ldb = dp;
      t.b_plane_stride = zb_stride;
      t.b_planes = kPlanes;
      pairs6(t);
      t.N = d;
      t.ldd = dp;
      t.alpha = -1.f;
      t.tiles = mg::TILES_FULL;
      t.epi = mg::EPI_ADD;
      t.ksplit = 1;
      const int64_t above = i0 - o0;                  // rows of this block above panel pi
      if (above > 0) {
        // (ordered adds) the block below updated these rows from the upd lane
        if (i0 + nb >= o_end && ob + 1 < nblocks) L.wait(L.chain, L.next_done[(ob + 1) & 1]);
        t.A = w.chol.l_planes + i0 * kp + o0;         // A[k, m] = L[i0 + k, o0 + m] = U[o0 + m, i0 + k]
        t.B = xq;
        t.M = above;
        t.K = nb;
        t.D = w.rhs + o0 * dp;
        MG_TIMED(L.chain, "nystrom.bwd_update_in_block", rc = mg::gemm_tn_launch(t, L.chain));
        if (rc) return rc;
        continue;
      }
      // pi is the block's first panel: rows [0, o0) get the whole block (K = o_end - o0)
      t.K = o_end - o0;
      t.B = xb;
      const int64_t first = kNB;                      // o0 is a multiple of 4 * 128
      const int64_t next_rows = kFwdOuter * kNB;      // rows of the block above (all of them exist)
      L.record(L.trsm, L.chain);
      L.wait(L.upd, L.trsm);
      {
        mg::GemmArgs r = t;                           // block above, except its last block row
        r.A = w.chol.l_planes + o0 * kp + (o0 - next_rows);
        r.M = next_rows - first;
        r.D = w.rhs + (o0 - next_rows) * dp;
        r.max_ctas = L.bulk_cta_cap();
        MG_TIMED(L.upd, "nystrom.bwd_update_next", rc = mg::gemm_tn_launch(r, L.upd));
        if (rc) return rc;
        L.record(L.next_done[ob & 1], L.upd);
      }
      if (o0 - next_rows > 0) {                       // everything above that block
        mg::GemmArgs r = t;
        r.A = w.chol.l_planes + o0 * kp;
        r.M = o0 - next_rows;
        r.D = w.rhs;
        r.max_ctas = L.bulk_cta_cap();
        MG_TIMED(L.upd, "nystrom.bwd_update", rc = mg::gemm_tn_launch(r, L.upd));
        if (rc) return rc;
      }
      L.record(L.upd_done[ob & 1], L.upd);
      // chain: the last block row of the block above (the next panel to be solved)
      if (ob + 1 < nblocks) L.wait(L.chain, L.upd_done[(ob + 1) & 1]);
      t.A = w.chol.l_planes + o0 * kp + (o0 - first);
      t.M = first;
      t.D = w.rhs + (o0 - first) * dp;
      MG_TIMED(L.chain, "nystrom.bwd_update_first", rc = mg::gemm_tn_launch(t, L.chain));
      if (rc) return rc;
    }
  }
  return 0;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int mg_set_concurrent_factorizations(int n) {
  mg::set_lane_share(n);
  return 0;
}

size_t mg_ridge_scores_ws_bytes(int64_t n) { return carve_scores(nullptr, n).bytes; }

int mg_ridge_scores_f32(const float* C, int64_t n, int64_t ldc, float ridge, float* scores,
                        void* ws, size_t ws_bytes, int* info, void* stream) {
  if (!C || !scores || !ws || !info) return -1;
  if (n <= 0) return -2;
  if (ldc < n) return -7;
  ScoresWs w = carve_scores(ws, n);
  if (ws_bytes < w.bytes) return -10;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t np = w.chol.n_pad;
  int rc;

  mg::LaneScope scope(s, n);
  const mg::Lanes& L = scope.lanes();

  copy_ridge_kernel<<<dim3(static_cast<unsigned>((n + 1023) / 1024 < 8 ? (n + 1023) / 1024 : 8),
                           static_cast<unsigned>(n)),
                      256, 0, L.chain>>>(C, ldc, w.a, np, n, ridge);
  if ((rc = cuda_rc())) return rc;

  // ---- blocked inverse of U, stored transposed (Y = U^-T, lower) as bf16 planes; the scores are
  //      the column sums of squares of Y.  Row block pj of Y only needs block rows 0..pj of U, so
  //      it is enqueued on the tri lane right behind Cholesky panel pj and runs concurrently with
  //      the factorisation of the later panels.
  cudaMemsetAsync(w.y_planes, 0, sizeof(bf16) * kPlanes * np * np, L.tri);
  cudaMemsetAsync(scores, 0, sizeof(float) * n, L.tri);
  const int64_t pstride = np * np;
  identity_kernel<<<(kNB * kNB + 255) / 256, 256, 0, L.tri>>>(w.ident, kNB);
  if ((rc = cuda_rc())) return rc;
  L.record(L.misc[0], L.tri);

  mg::CholStepper chol{w.a, n, np, w.chol, info, &L};
  for (int64_t j0 = 0, pj = 0; j0 < n; j0 += kNB, ++pj) {
    if ((rc = chol.step(pj))) return rc;
    const int nb = static_cast<int>(n - j0 < kNB ? n - j0 : kNB);
    const float* tf = w.chol.t_fwd + pj * mg::kTBlock;
    // diagonal block Y[jb, jb] = U_jj^-T (solve U_jj^T X = I): two CTAs of pure latency, so it
    // rides on the upd lane (which has slack) instead of lengthening the tri lane's own chain
    if (pj == 0) L.wait(L.upd, L.misc[0]);   // scores / y_planes are cleared on the tri lane
    L.wait(L.upd, L.trsm);
    MG_TIMED(L.upd, "trtri.diag_trsm", rc = mg::trsm128(tf, false, nb, w.ident, kNB, nb, 1.f, nullptr, 0,
                                                        w.y_planes + j0 * np + j0, np, pstride, nullptr, 0, 0,
                                                        scores + j0, L.upd));
    if (rc) return rc;
    L.record(L.diag_done[pj & 1], L.upd);
    if (j0 == 0) {
      L.record(L.row_done[0], L.tri);
      continue;
    }
    // Tt[nb, j0] = U[0:j0, jb]^T * Y[0:j0, 0:j0], split by the rows of Y it needs:
    //   bulk (tri2 lane): block rows <= pj-2 of Y — everything except the row finished last, so it
    //                     overlaps the previous row's triangular solve;
    //   tail (tri lane) : block row pj-1 of Y (K = 128), added on top once bulk has landed.
    float* tt = w.tt + (pj & 1) * kNB * np;
    mg::GemmArgs g{};
    g.lda = np;
    g.a_plane_stride = pstride;
    g.a_planes = kPlanes;
    g.ldb = np;
    g.b_plane_stride = pstride;
    g.b_planes = kPlanes;
    pairs6(g);
    g.M = nb;
    g.N = j0;
    g.D = tt;
    g.ldd = np;
    g.alpha = 1.f;
    g.tiles = mg::TILES_FULL;
    g.epi = mg::EPI_ADD;
    g.max_ctas = L.bulk_cta_cap();
    const int64_t kbulk = L.serial ? j0 : j0 - kNB;     // j0 is a multiple of 128
    cudaStream_t sb = kbulk > 0 ? L.tri2 : L.tri;
    if (kbulk > 0) {
      L.wait(L.tri2, L.trsm);                            // block row pj of U
      if (pj >= 2) L.wait(L.tri2, L.row_done[(pj - 2) & 1]);   // rows <= pj-2 of Y; tt[pj&1] is free
      L.wait(L.tri2, L.diag_done[(pj - 1) & 1]);        // (serial: no-op) diagonal blocks < pj
    }
    cudaMemsetAsync(tt, 0, sizeof(float) * kNB * np, sb);
    if (kbulk > 0) {
      mg::GemmArgs b = g;
      b.A = w.chol.u_planes + j0;
      b.B = w.y_planes;
      b.K = kbulk;
      b.ksplit = 0;
      b.klo_from_n = 1;
      MG_TIMED(L.tri2, "trtri.gemm_bulk", rc = mg::gemm_tn_launch(b, L.tri2));
      if (rc) return rc;
      L.record(L.bulk_done[pj & 1], L.tri2);
    }
    L.wait(L.tri, L.trsm);
    L.wait(L.tri, L.diag_done[(pj - 1) & 1]);
    if (kbulk > 0) L.wait(L.tri, L.bulk_done[pj & 1]);
    if (!L.serial) {
      mg::GemmArgs tl = g;
      tl.A = w.chol.u_planes + (j0 - kNB) * np + j0;
      tl.B = w.y_planes + (j0 - kNB) * np;
      tl.K = kNB;
      tl.ksplit = 1;
      tl.klo_from_n = 0;
      tl.max_ctas = 0;
      MG_TIMED(L.tri, "trtri.gemm_tail", rc = mg::gemm_tn_launch(tl, L.tri));
      if (rc) return rc;
    }
    // Y[jb, 0:j0] = -U_jj^-T Tt : planes + column sums of squares in one pass
    MG_TIMED(L.tri, "trtri.row_trsm", rc = mg::trsm128(tf, false, nb, tt, np, j0, -1.f, nullptr, 0,
                                                       w.y_planes + j0 * np, np, pstride, nullptr, 0, 0,
                                                       scores, L.tri));
    if (rc) return rc;
    L.record(L.row_done[pj & 1], L.tri);
  }
  mg::Prof::get().report(s, "mg_ridge_scores_f32");
  return 0;
}

int mg_select_k_f32(const float* scores, int64_t n, int64_t k, int largest, int64_t* idx_out,
                    void* stream) {
  if (!scores || !idx_out) return -1;
  if (n <= 0 || k < 0 || k > n) return -2;
  if (k == 0) return 0;
  select_k_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(scores, n, k, largest,
                                                                      idx_out);
  return cuda_rc();
}

int mg_gather_rows_bf16(const void* W, int64_t ldw, const int64_t* idx, int64_t k, int64_t d,
                        void* out, int64_t ldo, void* stream) {
  if (!W || !idx || !out) return -1;
  if (k <= 0 || d <= 0) return -2;
  if (ldw < d || ldo < d) return -7;
  const int vec = (d % 8 == 0) && (ldw % 8 == 0) && (ldo % 8 == 0) &&
                  ((reinterpret_cast<uintptr_t>(W) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  // hd is irrelevant with a single "head": q = 0 for every output row
  gather_rows_kernel<<<static_cast<unsigned>(k), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(W), ldw, idx, 1, 0, k, d, static_cast<bf16*>(out), ldo, vec);
  return cuda_rc();
}

int mg_gather_head_rows_bf16(const void* W, int64_t ldw, const int64_t* mask, int n_heads,
                             int group, int64_t hd, int64_t r, int64_t d, void* out, int64_t ldo,
                             void* stream) {
  if (!W || !mask || !out) return -1;
  if (n_heads <= 0 || group <= 0 || hd <= 0 || r <= 0 || d <= 0) return -2;
  if (ldw < d || ldo < d) return -7;
  const int vec = (d % 8 == 0) && (ldw % 8 == 0) && (ldo % 8 == 0) &&
                  ((reinterpret_cast<uintptr_t>(W) & 15) == 0) &&
                  ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  gather_rows_kernel<<<static_cast<unsigned>(n_heads * r), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(W), ldw, mask, group, hd, r, d, static_cast<bf16*>(out), ldo, vec);
  return cuda_rc();
}

size_t mg_nystrom_down_ws_bytes(int64_t n, int64_t k, int64_t d) {
  return carve_nystrom(nullptr, n, k, d).bytes;
}

int mg_nystrom_down_f32(const float* C, int64_t n, int64_t ldc, const int64_t* idx, int64_t k,
                        const void* Wd, int64_t d, int64_t ldwd, float jitter, void* Wd_out,
                        int64_t ld_out, void* ws, size_t ws_bytes, int* info, float* min_rel_pivot,
                        void* stream) {
  if (!C || !idx || !Wd || !Wd_out || !ws || !info) return -1;
  if (n <= 0 || k <= 0 || k > n || d <= 0) return -2;
  if (ldc < n || ldwd < n || ld_out < k) return -7;
  NystromWs w = carve_nystrom(ws, n, k, d);
  if (ws_bytes < w.bytes) return -10;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t kp = w.chol.n_pad, dp = mg::round_up(d, 64);
  int rc;

  mg::LaneScope scope(s, k);
  const mg::Lanes& L = scope.lanes();

  // Correction form.  With A = C_kk + jitter I and the columns of C split into kept / dropped,
  //     A^-1 C[idx, :] W_d^T = W_d[:, idx]^T + A^-1 (C[idx, dropped] W_d[:, dropped]^T - jitter W_d[:, idx]^T)
  // because A^-1 C_kk = I - jitter A^-1.  The kept channels' own contribution — the large,
  // exactly-cancelling part when "massive activation" channels are kept — never passes through
  // the fp32 right-hand side or the fp32 solve: only the correction does (measured on a statistic
  // with a 1e6 diagonal spread and equilibrated cond 3.5e4: 2e-2 -> 1e-5 relative error).
  // tri lane: operands of the cross term  rhs[k, d] = C[idx, dropped] W_down[:, dropped]^T
  // — independent of the factorisation of C_kk, which starts at once on the chain lane
  cudaMemsetAsync(w.kept, 0, static_cast<size_t>(n), L.tri);
  cudaMemsetAsync(w.pos, 0xFF, sizeof(int32_t) * static_cast<size_t>(n), L.tri);     // -1
  mark_kept_pos_kernel<<<static_cast<unsigned>((k + 255) / 256), 256, 0, L.tri>>>(idx, k, w.kept, w.pos);
  if ((rc = cuda_rc())) return rc;
  gather_cols_planes_kernel<<<dim3(static_cast<unsigned>((k + 255) / 256 < 16 ? (k + 255) / 256 : 16),
                                   static_cast<unsigned>(n)),
                              256, 0, L.tri>>>(C, ldc, n, idx, k, w.kept, w.g_planes, kp, n * kp);
  if ((rc = cuda_rc())) return rc;
  transpose_to_bf16_kernel<bf16><<<dim3(static_cast<unsigned>((n + 31) / 32),
                                        static_cast<unsigned>((d + 31) / 32)),
                                   256, 0, L.tri>>>(static_cast<const bf16*>(Wd), ldwd, d, n, w.wdt, dp);
  if ((rc = cuda_rc())) return rc;
  {
    mg::GemmArgs g{};
    g.A = w.g_planes;
    g.lda = kp;
    g.a_plane_stride = n * kp;
    g.a_planes = kPlanes;
    g.B = w.wdt;
    g.ldb = dp;
    g.b_planes = 1;
    g.npairs = 3;
    for (int i = 0; i < 3; ++i) {
      g.pair_a[i] = i;
      g.pair_b[i] = 0;
    }
    g.M = k;
    g.N = d;
    g.K = n;
    g.D = w.rhs;
    g.ldd = dp;
    g.alpha = 1.f;
    g.tiles = mg::TILES_FULL;
    g.epi = mg::EPI_STORE;
    g.ksplit = 1;
    g.max_ctas = L.bulk_cta_cap();
    MG_TIMED(L.tri, "nystrom.cross_term", rc = mg::gemm_tn_launch(g, L.tri));
    if (rc) return rc;
    sub_jitter_rows_kernel<<<static_cast<unsigned>(k), 256, 0, L.tri>>>(w.rhs, dp, w.wdt, dp, idx, d, jitter);
    if ((rc = cuda_rc())) return rc;
  }
  // chain lane: C_kk + jitter I and its Cholesky factor (planes of U and of L = U^T)
  gather_sym_kernel<<<dim3(static_cast<unsigned>((k + 255) / 256 < 16 ? (k + 255) / 256 : 16),
                           static_cast<unsigned>(k)),
                      256, 0, L.chain>>>(C, ldc, idx, k, w.ckk, kp, jitter);
  if ((rc = cuda_rc())) return rc;

  mg::CholStepper chol{w.ckk, k, kp, w.chol, info, &L};
  if ((rc = nystrom_solves(w, L, k, d, &chol))) return rc;
  // conditioning indicator: min_i U_ii^2 / A_ii (1 / (A_ii (A^-1)_ii) for the last pivot; small
  // values announce an ill-conditioned EQUILIBRATED system — the raw diagonal spread is harmless)
  if (min_rel_pivot) {
    set_float_kernel<<<1, 1, 0, L.chain>>>(min_rel_pivot, 3.0e38f);
    min_rel_pivot_kernel<<<static_cast<unsigned>((k + 255) / 256), 256, 0, L.chain>>>(
        C, ldc, idx, k, w.ckk, kp, jitter, min_rel_pivot);
    if ((rc = cuda_rc())) return rc;
  }
  // ---- xacc = W_down[:, idx]^T + X (fp32, kept for mg_nystrom_refine_f32);  W_down' [d, k] = xacc^T, bf16
  L.record(L.misc[0], L.tri);      // wdt was produced on the tri lane
  L.wait(L.chain, L.misc[0]);
  add_kept_rows_kernel<<<static_cast<unsigned>(k), 256, 0, L.chain>>>(w.rhs, dp, w.wdt, dp, idx, d, w.xacc, 0);
  if ((rc = cuda_rc())) return rc;
  transpose_to_bf16_kernel<float><<<dim3(static_cast<unsigned>((d + 31) / 32),
                                         static_cast<unsigned>((k + 31) / 32)),
                                    256, 0, L.chain>>>(w.xacc, dp, k, d, static_cast<bf16*>(Wd_out), ld_out);
  rc = cuda_rc();
  mg::Prof::get().report(s, "mg_nystrom_down_f32");
  return rc;
}

int mg_nystrom_refine_f32(const float* C, int64_t n, int64_t ldc, const int64_t* idx, int64_t k,
                          int64_t d, float jitter, void* Wd_out, int64_t ld_out, void* ws,
                          size_t ws_bytes, void* stream) {
  if (!C || !idx || !Wd_out || !ws) return -1;
  if (n <= 0 || k <= 0 || k > n || d <= 0) return -2;
  if (ldc < n || ld_out < k) return -7;
  NystromWs w = carve_nystrom(ws, n, k, d);
  if (ws_bytes < w.bytes) return -10;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t dp = mg::round_up(d, 64);
  int rc;
  // fp64 residual of the current solution -> fp32 right-hand side of the correction
  nystrom_resid64_kernel<<<dim3(static_cast<unsigned>((k + 127) / 128), static_cast<unsigned>((d + 127) / 128)),
                           256, 0, s>>>(C, ldc, n, idx, k, w.wdt, dp, w.pos, w.xacc, dp, d, jitter, w.rhs);
  if ((rc = cuda_rc())) return rc;
  {
    mg::LaneScope scope(s, k);
    if ((rc = nystrom_solves(w, scope.lanes(), k, d, nullptr))) return rc;
  }
  add_kept_rows_kernel<<<static_cast<unsigned>(k), 256, 0, s>>>(w.rhs, dp, w.wdt, dp, idx, d, w.xacc, 1);
  if ((rc = cuda_rc())) return rc;
  transpose_to_bf16_kernel<float><<<dim3(static_cast<unsigned>((d + 31) / 32),
                                         static_cast<unsigned>((k + 31) / 32)),
                                    256, 0, s>>>(w.xacc, dp, k, d, static_cast<bf16*>(Wd_out), ld_out);
  return cuda_rc();
}

}  // extern "C"
