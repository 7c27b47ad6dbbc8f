// mg_type23.cu — type-II (CR, Q/K) and type-III (SVD, V/O) paths, SURVEY §8 rows a10-a17.
//
// type-II  src/compression/compress_qk.py:320-476.  ||sqrt(C + rho I)[:, j]||^2 == C_jj + rho
//          (SURVEY Appendix B1), so the per-head eigendecompositions of the reference reduce to
//          reading two diagonals; one CTA per kv head scores, ranks and writes the rotary mask.
// type-III src/compression/compress_vo.py:112-223.  With G1 = W_v,h (C + rho I) W_v,h^T = V S^2 V^T
//          the reference's thin SVD of sqrt(C) W_v,h^T has right vectors V and singular values S
//          (Appendix B4/B5), and the new heads are hd x r recombinations of the old ones:
//            GQA: V' = (V_r S_r^-1)^T W_v,h        O'_j = W_o,j (V_r S_r)
//            MHA: B = S V^T (W_o,h^T W_o,h) V S = U_p S_p^2 U_p^T
//                 V' = (V S^-1 U_p[:, :r])^T W_v,h  O' = W_o,h (V S U_p[:, :r])
//          G1 and W_o^T W_o come from the tensor-core engine; the hd x hd symmetric eigenproblems
//          are solved per head by a parallel-ordering two-sided Jacobi in fp64 shared memory.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/modegpt_b200.h"
#include "mg_gemm.cuh"
#include "mg_linalg.cuh"
#include "mg_once.cuh"

namespace {

using bf16 = __nv_bfloat16;
using mg::kPlanes;

inline int cuda_rc() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - static_cast<int>(e);
}

// =================================================================================================
// type-II
// =================================================================================================
// mode 0: RoPE-paired (llama / qwen; GQA sums the group's query heads), top r/2 pairs,
//         mask = cat(idx, idx + hd/2) in descending-score order.
// mode 1: OPT, unpaired product of column norms, top r columns.
__global__ void __launch_bounds__(128) qk_select_kernel(const float* __restrict__ Cq,
                                                        const float* __restrict__ Ck, int group,
                                                        int hd, int mode, double ridge_q,
                                                        double ridge_k, int r,
                                                        int64_t* __restrict__ mask) {
  __shared__ double score[128];
  const int h = blockIdx.x, j = threadIdx.x;
  const int count = mode == 0 ? hd / 2 : hd;
  const int take = mode == 0 ? r / 2 : r;
  const float* ck = Ck + static_cast<int64_t>(h) * hd * hd;
  double sc = -1.0;
  if (j < count) {
    if (mode == 0) {
      const int j2 = j + hd / 2;
      const double k1 = static_cast<double>(ck[j * hd + j]) + ridge_k;
      const double k2 = static_cast<double>(ck[j2 * hd + j2]) + ridge_k;
      sc = 0.0;
      for (int g = 0; g < group; ++g) {
        const float* cq = Cq + (static_cast<int64_t>(h) * group + g) * hd * hd;
        sc += (static_cast<double>(cq[j * hd + j]) + ridge_q) * k1 +
              (static_cast<double>(cq[j2 * hd + j2]) + ridge_q) * k2;
      }
    } else {
      const float* cq = Cq + static_cast<int64_t>(h) * hd * hd;
      sc = sqrt(static_cast<double>(cq[j * hd + j]) + ridge_q) *
           sqrt(static_cast<double>(ck[j * hd + j]) + ridge_k);
    }
  }
  score[j] = sc;
  __syncthreads();
  if (j < count) {
    int rank = 0;
    for (int i = 0; i < count; ++i) rank += (score[i] > sc) || (score[i] == sc && i < j);
    if (rank < take) {
      int64_t* m = mask + static_cast<int64_t>(h) * r;
      m[rank] = j;
      if (mode == 0) m[take + rank] = j + hd / 2;
    }
  }
}

// =================================================================================================
// type-III: batched symmetric eigensolver + recombination factors
// =================================================================================================
// 1/sqrt(x), x > 0: single-precision seed + one third-order correction inside the float range
// (error 5 e^3 / 16 with e ~ 2^-22), the library routine outside it.
__device__ __forceinline__ double rsqrt_pos(double x) {
  if (!(x > 1e-30 && x < 1e30)) return rsqrt(x);
  const double y0 = static_cast<double>(rsqrtf(static_cast<float>(x)));
  const double e = fma(-x * y0, y0, 1.0);
  return fma(y0 * e, fma(0.375, e, 0.5), y0);
}

constexpr int kMaxHd = 128;

struct EigSmem {
  double* g;     // [hd][hd+1]   matrix being diagonalised (fp64)
  float* v;      // [hd][hd+4]   accumulated eigenvectors (fp32); transposed while the sweeps run
  double* cs;    // [hd]         (c, s) per pair
  int* pq;       // [hd]         (p, q) per pair
  double* lam;   // [hd]
  int* perm;     // [hd]
  double* red;   // [64]
};

__device__ __forceinline__ EigSmem carve_eig(uint8_t* base, int hd) {
  EigSmem s;
  s.g = reinterpret_cast<double*>(base);
  base += sizeof(double) * hd * (hd + 1);
  s.cs = reinterpret_cast<double*>(base);
  base += sizeof(double) * hd;
  s.lam = reinterpret_cast<double*>(base);
  base += sizeof(double) * hd;
  s.red = reinterpret_cast<double*>(base);
  base += sizeof(double) * 64;
  s.v = reinterpret_cast<float*>(base);
  base += sizeof(float) * hd * (hd + 4);
  s.pq = reinterpret_cast<int*>(base);
  base += sizeof(int) * hd;
  s.perm = reinterpret_cast<int*>(base);
  return s;
}

size_t eig_smem_bytes(int hd) {
  return sizeof(double) * hd * (hd + 1) + sizeof(double) * (2 * hd + 64) +
         sizeof(float) * hd * (hd + 4) + sizeof(int) * 2 * hd + 64;
}

__device__ double block_sum(double x, double* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = x;
  __syncthreads();
  double t = 0.0;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}

// Cyclic two-sided Jacobi with the round-robin (tournament) ordering: hd/2 disjoint rotations per
// step, hd-1 steps per sweep.  On exit lam[] holds the eigenvalues sorted descending and the
// columns of v[] the matching eigenvectors (v is rewritten in sorted order via perm).
__device__ void jacobi_eig(EigSmem& s, int hd) {
  const int t = threadIdx.x, nt = blockDim.x;
  const int ld = hd + 1, ldv = hd + 4;
  const int half = hd / 2, m = hd - 1;
  // the (k <= l) pair blocks this thread owns in every step (the assignment does not depend on
  // the step: only the index pairs behind k and l move)
  constexpr int kMaxBlocksPerThread = 3;   // 64 * 65 / 2 = 2080 blocks over 1024 threads
  int my_blocks[kMaxBlocksPerThread];
#pragma unroll
  for (int b = 0; b < kMaxBlocksPerThread; ++b) {
    const int e = t + b * nt;
    my_blocks[b] = -1;
    if (e < half * (half + 1) / 2) {
      int l = static_cast<int>((sqrtf(8.f * static_cast<float>(e) + 1.f) - 1.f) * 0.5f);
      while (l * (l + 1) / 2 > e) --l;
      while ((l + 1) * (l + 2) / 2 <= e) ++l;
      my_blocks[b] = ((e - l * (l + 1) / 2) << 8) | l;
    }
  }
  // V is kept TRANSPOSED during the sweeps (vt[col][row], rows 16-byte aligned): rotating columns
  // p, q of V is then a rotation of two contiguous rows, four components per 16-byte access
  for (int e = t; e < hd * hd; e += nt) {
    const int i = e / hd, k = e - i * hd;
    s.v[i * ldv + k] = (i == k) ? 1.f : 0.f;
  }
  __syncthreads();
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, dg = 0.0;
    for (int e = t; e < hd * hd; e += nt) {
      const int i = e / hd, k = e - i * hd;
      if (i > k) continue;                     // the strict lower triangle is stale
      const double x = s.g[i * ld + k];
      if (i == k) dg += x * x;
      else off += 2.0 * x * x;
    }
    off = block_sum(off, s.red);
    dg = block_sum(dg, s.red);
    // ||off|| <= 1e-11 ||diag||: one more (quadratically converging) sweep would take it to 1e-22,
    // far below the 1e-7 the fp32 eigenvector store and the fp32 factors downstream can resolve
    if (off <= 1e-22 * dg || off == 0.0) break;
    for (int step = 0; step < m; ++step) {
      if (t < half) {
        int p, q;
        if (t == 0) {
          p = m;
          q = step;
        } else {
          p = (step + t) % m;
          q = (step - t + m) % m;
        }
        if (p > q) {
          const int x = p;
          p = q;
          q = x;
        }
        const double gpq = s.g[p * ld + q];
        double c = 1.0, sn = 0.0;
        if (gpq != 0.0) {
          // t = sgn(tau) / (|tau| + sqrt(1 + tau^2)), tau = a / b  ==  sgn(a) b / (|a| + hypot(a, b)):
          // no division, and the three reciprocal square roots come from a single-precision seed
          // (this scalar chain runs on 64 threads while the other 960 wait at the barrier)
          const double a = s.g[q * ld + q] - s.g[p * ld + p], b = 2.0 * gpq;
          const double r2 = fma(a, a, b * b);
          const double den = fabs(a) + r2 * rsqrt_pos(r2);
          const double yd = rsqrt_pos(den);
          double tt = b * (yd * yd);
          tt = a < 0.0 ? -tt : tt;
          c = rsqrt_pos(fma(tt, tt, 1.0));
          sn = tt * c;
        }
        s.cs[2 * t] = c;
        s.cs[2 * t + 1] = sn;
        s.pq[2 * t] = p;
        s.pq[2 * t + 1] = q;
      }
      __syncthreads();
      // G <- J^T G J in one pass over the pair blocks {p_k,q_k} x {p_l,q_l} with k <= l: the
      // blocks are disjoint, G stays symmetric, and only its upper triangle (row <= column) is
      // stored and touched — each 2 x 2 block is read, rotated on both sides and written back
      // by one thread (block (l, k) is the mirror image and is never formed)
      // (all loads of the thread's blocks first, then the arithmetic and the stores: the blocks are
      //  disjoint, but the compiler cannot hoist a load of s.g above a store to s.g by itself)
      double gv[kMaxBlocksPerThread][4];
      int go[kMaxBlocksPerThread][4];      // element offsets inside s.g (upper-triangle addressing)
#pragma unroll
      for (int b = 0; b < kMaxBlocksPerThread; ++b) {
        const int kl = my_blocks[b];
        if (kl < 0) continue;
        const int k = kl >> 8, l = kl & 255;
        const int pk = s.pq[2 * k], qk = s.pq[2 * k + 1];
        const int pl = s.pq[2 * l], ql = s.pq[2 * l + 1];
        go[b][0] = pk <= pl ? pk * ld + pl : pl * ld + pk;
        go[b][1] = pk <= ql ? pk * ld + ql : ql * ld + pk;
        go[b][2] = qk <= pl ? qk * ld + pl : pl * ld + qk;
        go[b][3] = qk <= ql ? qk * ld + ql : ql * ld + qk;
#pragma unroll
        for (int x = 0; x < 4; ++x) gv[b][x] = s.g[go[b][x]];
      }
#pragma unroll
      for (int b = 0; b < kMaxBlocksPerThread; ++b) {
        const int kl = my_blocks[b];
        if (kl < 0) continue;
        const int k = kl >> 8, l = kl & 255;
        const double ck = s.cs[2 * k], sk = s.cs[2 * k + 1];
        const double cl = s.cs[2 * l], sl = s.cs[2 * l + 1];
        const double gpp = gv[b][0], gpq = gv[b][1], gqp = gv[b][2], gqq = gv[b][3];
        // right rotation (columns pl, ql)
        const double a_pp = cl * gpp - sl * gpq, a_pq = sl * gpp + cl * gpq;
        const double a_qp = cl * gqp - sl * gqq, a_qq = sl * gqp + cl * gqq;
        // left rotation (rows pk, qk)
        s.g[go[b][0]] = ck * a_pp - sk * a_qp;
        s.g[go[b][1]] = ck * a_pq - sk * a_qq;
        s.g[go[b][2]] = sk * a_pp + ck * a_qp;
        s.g[go[b][3]] = sk * a_pq + ck * a_qq;
      }
      // V <- V J  (rows p, q of the transposed store)
      const int quads = hd / 4;
      for (int e = t; e < quads * half; e += nt) {
        const int k = e / quads, i4 = e - k * quads;
        const int p = s.pq[2 * k], q = s.pq[2 * k + 1];
        const double c = s.cs[2 * k], sn = s.cs[2 * k + 1];
        float4* rp = reinterpret_cast<float4*>(s.v + p * ldv) + i4;
        float4* rq = reinterpret_cast<float4*>(s.v + q * ldv) + i4;
        // V is stored in fp32; rotating it in fp32 too (instead of widening every component) costs
        // one extra rounding per update and removes 16 conversions per pair from the slow pipe
        const float cf = static_cast<float>(c), sf = static_cast<float>(sn);
        const float4 vp = *rp, vq = *rq;
        *rp = make_float4(fmaf(cf, vp.x, -sf * vq.x), fmaf(cf, vp.y, -sf * vq.y),
                          fmaf(cf, vp.z, -sf * vq.z), fmaf(cf, vp.w, -sf * vq.w));
        *rq = make_float4(fmaf(sf, vp.x, cf * vq.x), fmaf(sf, vp.y, cf * vq.y),
                          fmaf(sf, vp.z, cf * vq.z), fmaf(sf, vp.w, cf * vq.w));
      }
      __syncthreads();
    }
  }
  // back to v[row][col] for the callers
  for (int e = t; e < hd * hd; e += nt) {
    const int i = e / hd, k = e - i * hd;
    if (i < k) {
      const float x = s.v[i * ldv + k];
      s.v[i * ldv + k] = s.v[k * ldv + i];
      s.v[k * ldv + i] = x;
    }
  }
  __syncthreads();
  // sort descending (rank by counting), stable on index
  if (t < hd) s.lam[t] = s.g[t * ld + t];
  __syncthreads();
  if (t < hd) {
    const double x = s.lam[t];
    int rank = 0;
    for (int i = 0; i < hd; ++i) rank += (s.lam[i] > x) || (s.lam[i] == x && i < t);
    s.perm[rank] = t;
  }
  __syncthreads();
}

// G[a, b] = sum_c X(a, c) * Y(b, c) in fp64 for one head and one slab of the reduction index:
// X(a, c) = X[head * xh + a * xsa + c * xsc] (likewise Y).  The fp32 tensor-core accumulators lose
// ~2^-24 of the LARGEST eigenvalue, which is what limits the small singular values the type-III
// factors divide by; the hd x hd Grams are therefore accumulated in fp64 on the CUDA cores
// (2 * d * hd^2 flop per head: 4.3 GFLOP for a 7B layer).  Partial sums per slab are written to
// out[head][slab][hd * hd] and added in a fixed order by the consumer (no atomics: results do not
// depend on timing).
constexpr int kGramKC = 32;   // reduction indices per shared-memory tile
template <class TX, class TY>
__global__ void __launch_bounds__(256)
    gram64_kernel(const TX* __restrict__ X, int64_t xh, int64_t xsa, int64_t xsc,
                  const TY* __restrict__ Y, int64_t yh, int64_t ysa, int64_t ysc, int same, int hd,
                  int64_t K, int64_t slab, double* __restrict__ out) {
  extern __shared__ __align__(16) double gsm[];
  double* xs = gsm;                              // [kGramKC][128]
  double* ys = same ? gsm : gsm + kGramKC * kMaxHd;
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int head = blockIdx.x, part = blockIdx.y, nparts = gridDim.y;
  const TX* x = X + static_cast<int64_t>(head) * xh;
  const TY* y = Y + static_cast<int64_t>(head) * yh;
  const int64_t c_begin = static_cast<int64_t>(part) * slab;
  const int64_t c_end = c_begin + slab < K ? c_begin + slab : K;
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  for (int64_t c0 = c_begin; c0 < c_end; c0 += kGramKC) {
    __syncthreads();
    // the fastest-varying thread index follows the contiguous direction of the operand
    for (int e = t; e < kGramKC * kMaxHd; e += 256) {
      int a, kk;
      if (xsc == 1) {
        kk = e & (kGramKC - 1);
        a = e / kGramKC;
      } else {
        a = e & (kMaxHd - 1);
        kk = e / kMaxHd;
      }
      const int64_t c = c0 + kk;
      xs[kk * kMaxHd + a] =
          (a < hd && c < c_end) ? static_cast<double>(static_cast<float>(x[a * xsa + c * xsc])) : 0.0;
    }
    if (!same) {
      for (int e = t; e < kGramKC * kMaxHd; e += 256) {
        int a, kk;
        if (ysc == 1) {
          kk = e & (kGramKC - 1);
          a = e / kGramKC;
        } else {
          a = e & (kMaxHd - 1);
          kk = e / kMaxHd;
        }
        const int64_t c = c0 + kk;
        ys[kk * kMaxHd + a] =
            (a < hd && c < c_end) ? static_cast<double>(static_cast<float>(y[a * ysa + c * ysc])) : 0.0;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int kk = 0; kk < kGramKC; ++kk) {
      double xa[8], yb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xa[i] = xs[kk * kMaxHd + ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) yb[j] = ys[kk * kMaxHd + tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fma(xa[i], yb[j], acc[i][j]);
    }
  }
  double* o = out + (static_cast<int64_t>(head) * nparts + part) * hd * hd;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int a = ty + 16 * i;
    if (a >= hd) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int b = tx + 16 * j;
      if (b < hd) o[a * hd + b] = acc[i][j];
    }
  }
}

__global__ void f32_to_f64_kernel(const float* __restrict__ in, double* __restrict__ out, int64_t n) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<double>(in[i]);
}

// dst (upper triangle, ld = ldd) = src + ridge * I
__global__ void copy_upper_ridge_kernel(const float* __restrict__ src, int64_t lds,
                                        float* __restrict__ dst, int64_t ldd, int64_t n, float ridge) {
  const int64_t r = blockIdx.y;
  for (int64_t c = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < n;
       c += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (c >= r) dst[r * ldd + c] = src[r * lds + c] + (c == r ? ridge : 0.f);
  }
}

// In-place Cholesky of the symmetric positive definite matrix in s.g (fp64, [hd][hd+1]): on success
// the LOWER triangle holds L (G = L L^T), the strict upper triangle is stale.  Right-looking, one
// pivot per round, all threads update the trailing block (hd <= 128: ~40 us).  Returns false —
// leaving s.g unusable — if a pivot is not safely positive (G numerically singular in fp64).
__device__ bool chol_lower_inplace(double* g, int ld, int hd, double* red) {
  const int t = threadIdx.x, nt = blockDim.x;
  __shared__ int s_ok;
  __shared__ double s_scale;
  if (t == 0) {
    double mx = 0.0;
    for (int i = 0; i < hd; ++i) mx = fmax(mx, g[i * ld + i]);
    s_scale = mx;
    s_ok = 1;
  }
  __syncthreads();
  const double floor_piv = 1e-13 * s_scale;
  for (int j = 0; j < hd; ++j) {
    if (t == 0) {
      const double piv = g[j * ld + j];
      if (!(piv > floor_piv)) s_ok = 0;
      red[0] = piv > floor_piv ? rsqrt(piv) : 0.0;
    }
    __syncthreads();
    if (!s_ok) return false;
    const double inv = red[0];
    // column j of L (rows j..hd-1): g[i][j] = g[i][j] / sqrt(piv); the lower triangle is kept current
    for (int i = j + t; i < hd; i += nt) g[i * ld + j] *= inv;
    __syncthreads();
    // trailing update of the lower triangle: g[i][k] -= L[i][j] L[k][j] for j < k <= i
    const int m = hd - j - 1;
    for (int e = t; e < m * m; e += nt) {
      const int ii = e / m, kk = e - ii * m;
      if (kk <= ii) {
        const int i = j + 1 + ii, k = j + 1 + kk;
        g[i * ld + k] = fma(-g[i * ld + j], g[k * ld + j], g[i * ld + k]);
      }
    }
    __syncthreads();
  }
  return true;
}

// One CTA per kv head.
//   G1 [KV][n1][hd*hd] fp64 partial sums, G2 [H][n2][hd*hd] fp64 (MHA only, may be null),
//   scratch [KV][4][hd*hd] fp32 (per head: D in fp32, T in fp64), Rv, Ro [KV][hd, r] fp32.
__global__ void __launch_bounds__(1024, 1)
    vo_factor_kernel(const double* __restrict__ G1, int n1, const double* __restrict__ G2, int n2,
                     int hd, int r, float* __restrict__ scratch, float* __restrict__ Rv,
                     float* __restrict__ Ro) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  EigSmem s = carve_eig(smem_raw, hd);
  const int t = threadIdx.x, nt = blockDim.x;
  const int ld = hd + 1, ldv = hd + 4;
  const int h = blockIdx.x;
  const int hh = hd * hd;
  const double* g1 = G1 + static_cast<int64_t>(h) * n1 * hh;
  float* rv = Rv + static_cast<int64_t>(h) * hd * r;
  float* ro = Ro + static_cast<int64_t>(h) * hd * r;

  for (int e = t; e < hh; e += nt) {
    const int i = e / hd, k = e - i * hd;
    double x = 0.0, y = 0.0;
    for (int p = 0; p < n1; ++p) {
      x += g1[p * hh + i * hd + k];
      y += g1[p * hh + k * hd + i];
    }
    // a failed factorisation (reported through *info by mg_vo_prepare) leaves NaNs behind: keep
    // the eigensolver's index arithmetic well defined, the caller discards the result anyway
    const double gsym = 0.5 * (x + y);
    s.g[i * ld + k] = isfinite(gsym) ? gsym : 0.0;
  }
  __syncthreads();

  if (G2 == nullptr) {
    // GQA: Rv = V_r S_r^-1, Ro = V_r S_r
    jacobi_eig(s, hd);
    for (int e = t; e < hd * r; e += nt) {
      const int i = e / r, a = e - i * r;
      const int col = s.perm[a];
      const double sv = sqrt(fmax(s.lam[col], 0.0));
      const double v = s.v[i * ldv + col];
      rv[e] = static_cast<float>(v / fmax(sv, 1e-30));
      ro[e] = static_cast<float>(v * sv);
    }
    return;
  }

  // ---- MHA.  The reference's second SVD acts on A = S V^T W_o^T, i.e. on B = D^T G2 D with
  // D = V S.  ANY D with D D^T = G1 serves: D' = D Q (Q orthogonal) turns B into Q^T B Q, its
  // eigenvectors into Q^T U_p, and leaves Ro = D U_p[:, :r] unchanged.  Taking the Cholesky factor
  // D = L (G1 = L L^T, ~40 us) replaces the FIRST of the two Jacobi eigensolves (~4 ms); the
  // eigensolve of G1 stays as the fallback for a G1 that is singular in fp64.
  const double* g2 = G2 + static_cast<int64_t>(h) * n2 * hh;
  float* vg = scratch + static_cast<int64_t>(h) * 4 * hh;       // D in fp32 (survives the eigensolve)
  double* tg = reinterpret_cast<double*>(vg + hh);              // T = G2 D in fp64 (global, L2-resident)
  if (chol_lower_inplace(s.g, ld, hd, s.red)) {
    for (int e = t; e < hh; e += nt) {       // D = L: clear the stale strict upper triangle
      const int i = e / hd, a = e - i * hd;
      if (a > i) s.g[i * ld + a] = 0.0;
    }
  } else {
    for (int e = t; e < hh; e += nt) {       // reload G1 (the failed factorisation overwrote it)
      const int i = e / hd, k = e - i * hd;
      double x = 0.0, y = 0.0;
      for (int p = 0; p < n1; ++p) {
        x += g1[p * hh + i * hd + k];
        y += g1[p * hh + k * hd + i];
      }
      const double gsym = 0.5 * (x + y);
      s.g[i * ld + k] = isfinite(gsym) ? gsym : 0.0;
    }
    __syncthreads();
    jacobi_eig(s, hd);
    double* sval = s.cs;  // singular values S (sorted), reuse the rotation buffer
    if (t < hd) sval[t] = sqrt(fmax(s.lam[s.perm[t]], 0.0));
    __syncthreads();
    for (int e = t; e < hh; e += nt) {       // D = V S, columns in descending order (s.g is free now)
      const int i = e / hd, a = e - i * hd;
      s.g[i * ld + a] = static_cast<double>(s.v[i * ldv + s.perm[a]]) * sval[a];
    }
  }
  __syncthreads();
  for (int e = t; e < hh; e += nt) {
    const int i = e / hd, a = e - i * hd;
    vg[e] = static_cast<float>(s.g[i * ld + a]);
  }
  // T = G2 D   (fp64, through global memory: shared memory holds D and has no room for T)
  for (int e = t; e < hh; e += nt) {
    const int i = e / hd, a = e - i * hd;
    double acc = 0.0;
    for (int k = 0; k < hd; ++k) {
      double x = 0.0, y = 0.0;
      for (int p = 0; p < n2; ++p) {
        x += g2[p * hh + i * hd + k];
        y += g2[p * hh + k * hd + i];
      }
      acc += 0.5 * (x + y) * s.g[k * ld + a];
    }
    tg[e] = acc;
  }
  __syncthreads();
  // B = D^T T  (symmetric): computed into registers first (s.g holds D), then written back
  {
    constexpr int kPer = kMaxHd * kMaxHd / 1024;   // elements per thread at hd = 128
    double bacc[kPer];
#pragma unroll
    for (int c = 0; c < kPer; ++c) {
      const int e = t + c * nt;
      double acc = 0.0;
      if (e < hh) {
        const int a = e / hd, b = e - a * hd;
        for (int k = 0; k < hd; ++k) acc += s.g[k * ld + a] * tg[k * hd + b];
      }
      bacc[c] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kPer; ++c) {
      const int e = t + c * nt;
      if (e < hh) {
        const int a = e / hd, b = e - a * hd;
        s.g[a * ld + b] = bacc[c];
      }
    }
  }
  __syncthreads();
  for (int e = t; e < hh; e += nt) {  // symmetrise
    const int a = e / hd, b = e - a * hd;
    if (a < b) {
      const double x = 0.5 * (s.g[a * ld + b] + s.g[b * ld + a]);
      s.g[a * ld + b] = x;
      s.g[b * ld + a] = x;
    }
  }
  __syncthreads();
  jacobi_eig(s, hd);  // s.v now holds U_p (unsorted columns), perm the descending order
  // Ro = D U_p[:, :r].  Rv = D^-T U_p[:, :r] (= V S^-1 U_p[:, :r] for D = V S) is NOT formed through
  // D^-1: the components of the leading eigenvectors along the small singular directions are tiny
  // and carry only the eigensolver's absolute accuracy.  From B u = lambda u with B = D^T G2 D:
  //     D^-T u = G2 (D u) / lambda     =>     Rv = G2 Ro Lambda_r^-1,
  // which only ever divides by the r LARGEST eigenvalues of B.
  for (int e = t; e < hd * r; e += nt) {
    const int i = e / r, a = e - i * r;
    const int col = s.perm[a];
    double ao = 0.0;
    for (int k = 0; k < hd; ++k)
      ao += static_cast<double>(vg[i * hd + k]) * static_cast<double>(s.v[k * ldv + col]);
    ro[e] = static_cast<float>(ao);
  }
  __syncthreads();     // ro (global, written by this CTA) is read back below
  for (int e = t; e < hd * r; e += nt) {
    const int i = e / r, a = e - i * r;
    const double lam_a = s.lam[s.perm[a]];
    double av = 0.0;
    for (int j = 0; j < hd; ++j) {
      double x = 0.0, y = 0.0;
      for (int p = 0; p < n2; ++p) {
        x += g2[p * hh + i * hd + j];
        y += g2[p * hh + j * hd + i];
      }
      av += 0.5 * (x + y) * static_cast<double>(ro[j * r + a]);
    }
    rv[e] = lam_a > 0.0 ? static_cast<float>(av / lam_a) : 0.f;
  }
}

template <class OUT>
__device__ __forceinline__ OUT to_out(float x);
template <>
__device__ __forceinline__ bf16 to_out<bf16>(float x) { return __float2bfloat16_rn(x); }
template <>
__device__ __forceinline__ float to_out<float>(float x) { return x; }

template <class OUT>
__global__ void __launch_bounds__(256) vo_apply_v_kernel(const bf16* __restrict__ Wv, int64_t ldwv,
                                                         const float* __restrict__ Rv, int hd,
                                                         int r, int64_t d, OUT* __restrict__ out,
                                                         int64_t ldo) {
  extern __shared__ float sh[];  // Rv[h]: [hd][r]
  const int h = blockIdx.y;
  const float* rv = Rv + static_cast<int64_t>(h) * hd * r;
  for (int e = threadIdx.x; e < hd * r; e += 256) sh[e] = rv[e];
  __syncthreads();
  const int cx = threadIdx.x & 63, ay = threadIdx.x >> 6;
  const int64_t c = static_cast<int64_t>(blockIdx.x) * 64 + cx;
  if (c >= d) return;
  constexpr int kMaxA = kMaxHd / 4;
  float acc[kMaxA];
#pragma unroll
  for (int i = 0; i < kMaxA; ++i) acc[i] = 0.f;
  const bf16* w = Wv + static_cast<int64_t>(h) * hd * ldwv + c;
  for (int k = 0; k < hd; ++k) {
    const float x = __bfloat162float(w[k * ldwv]);
    const float* rr = sh + k * r;
#pragma unroll
    for (int i = 0; i < kMaxA; ++i) {
      const int a = ay + 4 * i;
      if (a < r) acc[i] = fmaf(rr[a], x, acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < kMaxA; ++i) {
    const int a = ay + 4 * i;
    if (a < r) out[(static_cast<int64_t>(h) * r + a) * ldo + c] = to_out<OUT>(acc[i]);
  }
}

// O'[c, q*r + a] = sum_k Wo[c, q*hd + k] * Ro[q / group][k, a]
template <class OUT>
__global__ void __launch_bounds__(256) vo_apply_o_kernel(const bf16* __restrict__ Wo, int64_t ldwo,
                                                         const float* __restrict__ Ro, int group,
                                                         int hd, int r, int64_t d,
                                                         OUT* __restrict__ out, int64_t ldo) {
  extern __shared__ float sh[];
  float* ro_s = sh;                 // [hd][r]
  float* w_s = sh + hd * r;         // [64][hd + 1]
  const int q = blockIdx.y;
  const float* ro = Ro + static_cast<int64_t>(q / group) * hd * r;
  for (int e = threadIdx.x; e < hd * r; e += 256) ro_s[e] = ro[e];
  const int64_t c0 = static_cast<int64_t>(blockIdx.x) * 64;
  for (int e = threadIdx.x; e < 64 * hd; e += 256) {
    const int i = e / hd, k = e - i * hd;
    const int64_t c = c0 + i;
    w_s[i * (hd + 1) + k] =
        c < d ? __bfloat162float(Wo[c * ldwo + static_cast<int64_t>(q) * hd + k]) : 0.f;
  }
  __syncthreads();
  // thread -> (row ci, column strip): consecutive threads take consecutive output columns
  const int nthr_a = 32;            // 32 threads across a, 8 rows at a time
  const int ax = threadIdx.x & 31, cy = threadIdx.x >> 5;
  for (int ci = cy; ci < 64; ci += 8) {
    const int64_t c = c0 + ci;
    if (c >= d) break;
    for (int a = ax; a < r; a += nthr_a) {
      float acc = 0.f;
      for (int k = 0; k < hd; ++k) acc = fmaf(w_s[ci * (hd + 1) + k], ro_s[k * r + a], acc);
      out[c * ldo + static_cast<int64_t>(q) * r + a] = to_out<OUT>(acc);
    }
  }
}

__global__ void __launch_bounds__(256) transpose_bf16_kernel(const bf16* __restrict__ in,
                                                             int64_t ld_in, int64_t rows,
                                                             int64_t cols, bf16* __restrict__ out,
                                                             int64_t ld_out) {
  __shared__ bf16 tile[32][34];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 32, c0 = static_cast<int64_t>(blockIdx.x) * 32;
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? in[r * ld_in + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t orow = c0 + i, ocol = r0 + tx;
    if (orow < cols && ocol < rows) out[orow * ld_out + ocol] = tile[tx][i];
  }
}

// ---- workspace ----------------------------------------------------------------------------------
constexpr int kMaxGramParts = 32;

inline int gram_parts(int heads) {
  int n = (2 * 148 + heads - 1) / heads;
  return n < 1 ? 1 : (n > kMaxGramParts ? kMaxGramParts : n);
}

struct VoWs {
  // factor route (method 0): C + rho I = U^T U, M^T = W_v U^T, G1 = M^T M (fp64)
  float* a;         // [d x dp]       C + rho I (upper) -> U in place
  mg::CholWorkspace chol;
  bf16* ut_planes;  // [3][dp x dp]   planes of U^T (lower, zero above); Gram route: planes of C + rho I
  bf16* wvt;        // [d x vp]       W_v^T
  float* mt;        // [vp x dp]      M^T = W_v U^T;  Gram route: P = (C + rho I) W_v^T as [d x vp]
  double* g1;       // [KV][parts][hd*hd]
  float* g2f;       // [H, hd, hd]    tensor-core W_o,h^T W_o,h (hd in {32, 64, 128})
  double* g2;       // [H][parts][hd*hd]
  float* scratch;   // [KV][4][hd*hd]  per head: D fp32, T fp64
  float* rv;        // [KV][hd, r<=hd]
  float* ro;
  size_t bytes;
};

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

VoWs carve_vo(void* ptr, int64_t d, int H, int KV, int hd) {
  const int64_t dp = mg::round_up(d, 64), vp = mg::round_up(static_cast<int64_t>(KV) * hd, 64);
  const int64_t panels = (d + mg::kNB - 1) / mg::kNB;
  const size_t hh = static_cast<size_t>(hd) * hd;
  Carver c(ptr);
  VoWs w{};
  w.a = c.take<float>(d * dp);
  w.chol.u_planes = c.take<bf16>(kPlanes * dp * dp);
  w.chol.l_planes = nullptr;
  w.chol.t_fwd = c.take<float>(panels * mg::kTBlock);
  w.chol.t_bwd = nullptr;
  w.chol.n_pad = dp;
  w.ut_planes = c.take<bf16>(kPlanes * dp * dp);
  w.wvt = c.take<bf16>(d * vp);
  w.mt = c.take<float>((vp > d ? vp : d) * (dp > vp ? dp : vp));
  w.g1 = c.take<double>(static_cast<size_t>(KV) * gram_parts(KV) * hh);
  w.g2f = c.take<float>(static_cast<size_t>(H) * hh);
  w.g2 = c.take<double>(static_cast<size_t>(H) * gram_parts(H) * hh);
  w.scratch = c.take<float>(static_cast<size_t>(KV) * 4 * hh);
  w.rv = c.take<float>(static_cast<size_t>(KV) * hh);
  w.ro = c.take<float>(static_cast<size_t>(KV) * hh);
  w.bytes = c.off + 256;
  return w;
}

inline bool tensor_hd(int hd) { return hd == 32 || hd == 64 || hd == 128; }
inline bool valid_hd(int hd) { return hd >= 4 && hd <= kMaxHd && hd % 4 == 0; }

template <class TX, class TY>
int launch_gram64(const TX* X, int64_t xh, int64_t xsa, int64_t xsc, const TY* Y, int64_t yh,
                  int64_t ysa, int64_t ysc, bool same, int heads, int hd, int64_t K, int parts,
                  double* out, cudaStream_t s) {
  static mg::PerDeviceOnce once;
  const int smem = static_cast<int>(sizeof(double) * kGramKC * kMaxHd * 2);
  if (int rc = once.run([smem] {
        return cudaFuncSetAttribute(gram64_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    smem);
      }))
    return rc;
  int64_t slab = (K + parts - 1) / parts;
  slab = mg::round_up(slab, kGramKC);
  gram64_kernel<TX, TY><<<dim3(heads, parts), 256, same ? smem / 2 : smem, s>>>(
      X, xh, xsa, xsc, Y, yh, ysa, ysc, same ? 1 : 0, hd, K, slab, out);
  return cuda_rc();
}

}  // namespace

extern "C" {

int mg_qk_select_f32(const float* Cq, const float* Ck, int n_heads, int n_kv_heads, int hd,
                     int mode, float ridge_q, float ridge_k, int r, int64_t* mask, void* stream) {
  if (!Cq || !Ck || !mask) return -1;
  if (n_heads <= 0 || n_kv_heads <= 0 || n_heads % n_kv_heads) return -2;
  if (hd <= 0 || hd > 128 || (mode == 0 && (hd % 2 || r % 2))) return -11;
  if (r <= 0 || r > hd) return -11;
  if (mode != 0 && mode != 1) return -11;
  if (mode == 1 && n_heads != n_kv_heads) return -11;
  qk_select_kernel<<<n_kv_heads, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      Cq, Ck, n_heads / n_kv_heads, hd, mode, static_cast<double>(ridge_q),
      static_cast<double>(ridge_k), r, mask);
  return cuda_rc();
}

size_t mg_vo_ws_bytes(int64_t d, int n_heads, int n_kv_heads, int hd) {
  return carve_vo(nullptr, d, n_heads, n_kv_heads, hd).bytes;
}

int mg_vo_prepare(const float* Cx, int64_t ldc, float ridge, const void* Wv, int64_t ldwv,
                  const void* Wo, int64_t ldwo, int n_heads, int n_kv_heads, int hd, int64_t d,
                  int method, int* info, void* ws, size_t ws_bytes, void* stream) {
  if (!Cx || !Wv || !Wo || !ws || !info) return -1;
  if (d <= 0 || n_heads <= 0 || n_kv_heads <= 0 || n_heads % n_kv_heads) return -2;
  if (!valid_hd(hd)) return -6;
  if (method != MG_VO_FACTOR && method != MG_VO_GRAM) return -11;
  if (ldc < d || ldwv < d || ldwo < static_cast<int64_t>(n_heads) * hd) return -7;
  VoWs w = carve_vo(ws, d, n_heads, n_kv_heads, hd);
  if (ws_bytes < w.bytes) return -10;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t dp = mg::round_up(d, 64);
  const int64_t nv = static_cast<int64_t>(n_kv_heads) * hd, vp = mg::round_up(nv, 64);
  const bool mha = n_heads == n_kv_heads;
  const bf16* wv = static_cast<const bf16*>(Wv);
  const bf16* wo = static_cast<const bf16*>(Wo);
  const int parts1 = gram_parts(n_kv_heads);
  int rc;

  cudaMemsetAsync(info, 0, sizeof(int), s);
  transpose_bf16_kernel<<<dim3(static_cast<unsigned>((d + 31) / 32),
                               static_cast<unsigned>((nv + 31) / 32)),
                          256, 0, s>>>(wv, ldwv, nv, d, w.wvt, vp);
  if ((rc = cuda_rc())) return rc;

  if (method == MG_VO_FACTOR) {
    // ---- C + rho I = U^T U  (blocked Cholesky, fp32 storage, fp64 diagonal blocks)
    copy_upper_ridge_kernel<<<dim3(static_cast<unsigned>((d + 1023) / 1024 < 8 ? (d + 1023) / 1024 : 8),
                                   static_cast<unsigned>(d)),
                              256, 0, s>>>(Cx, ldc, w.a, dp, d, ridge);
    if ((rc = cuda_rc())) return rc;
    {
      mg::LaneScope scope(s, d);
      if ((rc = mg::cholesky_upper(w.a, d, dp, w.chol, info, scope.lanes()))) return rc;
    }
    // planes of U^T (lower triangular, explicit zeros above the diagonal)
    if ((rc = mg::split_planes(w.a, dp, d, d, w.ut_planes, dp, dp * dp, true, nullptr, s, 0.f, true)))
      return rc;
    // M^T[nv, d] = W_v U^T :  D[m, n] = sum_k wvt[k, m] * Ut[k, n],  Ut[k, n] = U[n, k] = 0 for n > k
    {
      mg::GemmArgs g{};
      g.A = w.wvt;
      g.lda = vp;
      g.a_planes = 1;
      g.B = w.ut_planes;
      g.ldb = dp;
      g.b_plane_stride = dp * dp;
      g.b_planes = kPlanes;
      g.npairs = 3;
      for (int i = 0; i < 3; ++i) {
        g.pair_a[i] = 0;
        g.pair_b[i] = i;
      }
      g.M = nv;
      g.N = d;
      g.K = d;
      g.D = w.mt;
      g.ldd = dp;
      g.alpha = 1.f;
      g.tiles = mg::TILES_FULL;
      g.epi = mg::EPI_STORE;
      g.ksplit = 1;
      g.klo_from_n = 1;
      if ((rc = mg::gemm_tn_launch(g, s))) return rc;
    }
    // G1[h] = M_h^T M_h in fp64: rows h*hd .. of M^T, reduction over its d columns
    rc = launch_gram64<float, float>(w.mt, static_cast<int64_t>(hd) * dp, dp, 1, w.mt,
                                     static_cast<int64_t>(hd) * dp, dp, 1, true, n_kv_heads, hd, d,
                                     parts1, w.g1, s);
    if (rc) return rc;
  } else {
    // ---- Gram route: P = (C + rho I) W_v^T on the tensor cores (fp32), G1[h] = W_v,h P_h in fp64.
    //      Works for any symmetric C (no factorisation to break down) but the fp32 rounding of P
    //      costs ~2^-24 (sigma_1 / sigma_r)^2 in the small singular values.
    if ((rc = mg::split_planes(Cx, ldc, d, d, w.ut_planes, dp, dp * dp, false, nullptr, s, ridge)))
      return rc;
    {
      mg::GemmArgs g{};
      g.A = w.ut_planes;
      g.lda = dp;
      g.a_plane_stride = dp * dp;
      g.a_planes = kPlanes;
      g.B = w.wvt;
      g.ldb = vp;
      g.b_planes = 1;
      g.npairs = 3;
      for (int i = 0; i < 3; ++i) {
        g.pair_a[i] = i;
        g.pair_b[i] = 0;
      }
      g.M = d;
      g.N = nv;
      g.K = d;
      g.D = w.mt;
      g.ldd = vp;
      g.alpha = 1.f;
      g.tiles = mg::TILES_FULL;
      g.epi = mg::EPI_STORE;     // one accumulation run per tile: deterministic
      g.ksplit = 1;
      if ((rc = mg::gemm_tn_launch(g, s))) return rc;
    }
    // X(a, c) = W_v[h*hd + a, c] (bf16),  Y(b, c) = P[c, h*hd + b] (fp32)
    rc = launch_gram64<bf16, float>(wv, static_cast<int64_t>(hd) * ldwv, ldwv, 1, w.mt, hd, 1, vp, false,
                                    n_kv_heads, hd, d, parts1, w.g1, s);
    if (rc) return rc;
  }
  if (mha) {
    if (tensor_hd(hd)) {
      // overwrite mode = one accumulation run per tile, no split-K: results do not depend on timing
      rc = mg_syrk_heads_bf16_f32(wo, d, static_cast<int64_t>(n_heads) * hd, ldwo, hd, w.g2f, 1.f, 0,
                                  stream);
      if (rc) return rc;
      const int64_t cnt = static_cast<int64_t>(n_heads) * hd * hd;
      f32_to_f64_kernel<<<static_cast<unsigned>((cnt + 255) / 256), 256, 0, s>>>(w.g2f, w.g2, cnt);
      if ((rc = cuda_rc())) return rc;
    } else {
      // X(a, c) = W_o[c, h*hd + a]
      rc = launch_gram64<bf16, bf16>(wo, hd, 1, ldwo, wo, hd, 1, ldwo, true, n_heads, hd, d,
                                     gram_parts(n_heads), w.g2, s);
      if (rc) return rc;
    }
  }
  return 0;
}

int mg_vo_finish(const void* Wv, int64_t ldwv, const void* Wo, int64_t ldwo, int n_heads,
                 int n_kv_heads, int hd, int64_t d, int r, void* Wv_out, int64_t ldv_out,
                 void* Wo_out, int64_t ldo_out, int out_f32, void* ws, size_t ws_bytes,
                 void* stream) {
  if (!Wv || !Wo || !Wv_out || !Wo_out || !ws) return -1;
  if (d <= 0 || n_heads <= 0 || n_kv_heads <= 0 || n_heads % n_kv_heads) return -2;
  if (!valid_hd(hd)) return -6;
  if (r <= 0 || r > hd) return -11;
  if (ldwv < d || ldwo < static_cast<int64_t>(n_heads) * hd || ldv_out < d ||
      ldo_out < static_cast<int64_t>(n_heads) * r)
    return -7;
  VoWs w = carve_vo(ws, d, n_heads, n_kv_heads, hd);
  if (ws_bytes < w.bytes) return -10;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int group = n_heads / n_kv_heads;
  const bool mha = group == 1;
  const bf16* wv = static_cast<const bf16*>(Wv);
  const bf16* wo = static_cast<const bf16*>(Wo);
  int rc;
  const size_t esm = eig_smem_bytes(hd);
  const size_t osm = sizeof(float) * (hd * r + 64 * (hd + 1));
  static mg::PerDeviceOnce once;
  rc = once.run([] {
    const int o_max = static_cast<int>(sizeof(float) * (kMaxHd * kMaxHd + 64 * (kMaxHd + 1)));
    const int v_max = static_cast<int>(sizeof(float) * kMaxHd * kMaxHd);
    cudaError_t e = cudaFuncSetAttribute(vo_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(eig_smem_bytes(kMaxHd)));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(vo_apply_o_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, o_max);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(vo_apply_o_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, o_max);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(vo_apply_v_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, v_max);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(vo_apply_v_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, v_max);
    return e;
  });
  if (rc) return rc;
  const int parts2 = tensor_hd(hd) ? 1 : gram_parts(n_heads);
  vo_factor_kernel<<<n_kv_heads, 1024, esm, s>>>(w.g1, gram_parts(n_kv_heads), mha ? w.g2 : nullptr,
                                                 parts2, hd, r, w.scratch, w.rv, w.ro);
  if ((rc = cuda_rc())) return rc;
  const dim3 gv(static_cast<unsigned>((d + 63) / 64), n_kv_heads);
  const dim3 go(static_cast<unsigned>((d + 63) / 64), n_heads);
  const size_t vsm = sizeof(float) * hd * r;
  if (out_f32) {
    vo_apply_v_kernel<float><<<gv, 256, vsm, s>>>(wv, ldwv, w.rv, hd, r, d,
                                                  static_cast<float*>(Wv_out), ldv_out);
    if ((rc = cuda_rc())) return rc;
    vo_apply_o_kernel<float><<<go, 256, osm, s>>>(wo, ldwo, w.ro, group, hd, r, d,
                                                  static_cast<float*>(Wo_out), ldo_out);
  } else {
    vo_apply_v_kernel<bf16><<<gv, 256, vsm, s>>>(wv, ldwv, w.rv, hd, r, d,
                                                 static_cast<bf16*>(Wv_out), ldv_out);
    if ((rc = cuda_rc())) return rc;
    vo_apply_o_kernel<bf16><<<go, 256, osm, s>>>(wo, ldwo, w.ro, group, hd, r, d,
                                                 static_cast<bf16*>(Wo_out), ldo_out);
  }
  return cuda_rc();
}

int mg_vo_compress(const float* Cx, int64_t ldc, float ridge, const void* Wv, int64_t ldwv,
                   const void* Wo, int64_t ldwo, int n_heads, int n_kv_heads, int hd, int64_t d,
                   int r, int method, void* Wv_out, int64_t ldv_out, void* Wo_out, int64_t ldo_out,
                   int out_f32, int* info, void* ws, size_t ws_bytes, void* stream) {
  if (!Wv_out || !Wo_out) return -1;
  if (r <= 0 || r > hd) return -11;
  int rc = mg_vo_prepare(Cx, ldc, ridge, Wv, ldwv, Wo, ldwo, n_heads, n_kv_heads, hd, d, method, info,
                         ws, ws_bytes, stream);
  if (rc) return rc;
  return mg_vo_finish(Wv, ldwv, Wo, ldwo, n_heads, n_kv_heads, hd, d, r, Wv_out, ldv_out, Wo_out,
                      ldo_out, out_f32, ws, ws_bytes, stream);
}

}  // extern "C"
