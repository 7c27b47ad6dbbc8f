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
    if (off <= 1e-26 * dg || off == 0.0) break;
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

// One CTA per kv head.
//   G1 [KV, hd, hd] fp32, G2 [H, hd, hd] fp32 (MHA only, may be null), scratch [KV][2][hd*hd] fp32,
//   Rv, Ro [KV][hd, r] fp32.
__global__ void __launch_bounds__(1024, 1)
    vo_factor_kernel(const float* __restrict__ G1, const float* __restrict__ G2, int hd, int r,
                     float* __restrict__ scratch, float* __restrict__ Rv, float* __restrict__ Ro) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  EigSmem s = carve_eig(smem_raw, hd);
  const int t = threadIdx.x, nt = blockDim.x;
  const int ld = hd + 1, ldv = hd + 4;
  const int h = blockIdx.x;
  const float* g1 = G1 + static_cast<int64_t>(h) * hd * hd;
  float* rv = Rv + static_cast<int64_t>(h) * hd * r;
  float* ro = Ro + static_cast<int64_t>(h) * hd * r;

  for (int e = t; e < hd * hd; e += nt) {
    const int i = e / hd, k = e - i * hd;
    s.g[i * ld + k] = 0.5 * (static_cast<double>(g1[i * hd + k]) + static_cast<double>(g1[k * hd + i]));
  }
  __syncthreads();
  jacobi_eig(s, hd);

  if (G2 == nullptr) {
    // GQA: Rv = V_r S_r^-1, Ro = V_r S_r
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

  // ---- MHA second stage
  const float* g2 = G2 + static_cast<int64_t>(h) * hd * hd;
  float* vg = scratch + static_cast<int64_t>(h) * 2 * hd * hd;  // V (sorted columns)
  float* tg = vg + hd * hd;                                     // T = G2 * D
  double* sval = s.cs;  // singular values S (sorted), reuse the rotation buffer
  if (t < hd) sval[t] = sqrt(fmax(s.lam[s.perm[t]], 0.0));
  for (int e = t; e < hd * hd; e += nt) {
    const int i = e / hd, a = e - i * hd;
    vg[e] = s.v[i * ldv + s.perm[a]];
  }
  __syncthreads();
  // D = V S in shared memory (overwrites v, sorted order)
  for (int e = t; e < hd * hd; e += nt) {
    const int i = e / hd, a = e - i * hd;
    s.v[i * ldv + a] = static_cast<float>(static_cast<double>(vg[e]) * sval[a]);
  }
  __syncthreads();
  // T = G2 D
  for (int e = t; e < hd * hd; e += nt) {
    const int i = e / hd, a = e - i * hd;
    double acc = 0.0;
    for (int k = 0; k < hd; ++k)
      acc += 0.5 * (static_cast<double>(g2[i * hd + k]) + static_cast<double>(g2[k * hd + i])) *
             static_cast<double>(s.v[k * ldv + a]);
    tg[e] = static_cast<float>(acc);
  }
  __syncthreads();
  // B = D^T T  (symmetric), into the fp64 buffer
  for (int e = t; e < hd * hd; e += nt) {
    const int a = e / hd, b = e - a * hd;
    double acc = 0.0;
    for (int k = 0; k < hd; ++k)
      acc += static_cast<double>(s.v[k * ldv + a]) * static_cast<double>(tg[k * hd + b]);
    s.g[a * ld + b] = acc;
  }
  __syncthreads();
  for (int e = t; e < hd * hd; e += nt) {  // symmetrise
    const int a = e / hd, b = e - a * hd;
    if (a < b) {
      const double x = 0.5 * (s.g[a * ld + b] + s.g[b * ld + a]);
      s.g[a * ld + b] = x;
      s.g[b * ld + a] = x;
    }
  }
  // keep S: jacobi_eig overwrites cs; stash S in red-free space = lam after copying
  __shared__ double s_keep[kMaxHd];
  if (t < hd) s_keep[t] = sval[t];
  __syncthreads();
  jacobi_eig(s, hd);  // s.v now holds U_p (unsorted columns), perm the descending order
  // Rv = V S^-1 U_p[:, :r],  Ro = V S U_p[:, :r]
  for (int e = t; e < hd * r; e += nt) {
    const int i = e / r, a = e - i * r;
    const int col = s.perm[a];
    double av = 0.0, ao = 0.0;
    for (int k = 0; k < hd; ++k) {
      const double v = vg[i * hd + k];
      const double up = s.v[k * ldv + col];
      const double sv = s_keep[k];
      av += v / fmax(sv, 1e-30) * up;
      ao += v * sv * up;
    }
    rv[e] = static_cast<float>(av);
    ro[e] = static_cast<float>(ao);
  }
}

// V'[h*r + a, c] = sum_k Rv[h][k, a] * Wv[h*hd + k, c]
__global__ void __launch_bounds__(256) vo_apply_v_kernel(const bf16* __restrict__ Wv, int64_t ldwv,
                                                         const float* __restrict__ Rv, int hd,
                                                         int r, int64_t d, bf16* __restrict__ out,
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
    if (a < r) out[(static_cast<int64_t>(h) * r + a) * ldo + c] = __float2bfloat16_rn(acc[i]);
  }
}

// O'[c, q*r + a] = sum_k Wo[c, q*hd + k] * Ro[q / group][k, a]
__global__ void __launch_bounds__(256) vo_apply_o_kernel(const bf16* __restrict__ Wo, int64_t ldwo,
                                                         const float* __restrict__ Ro, int group,
                                                         int hd, int r, int64_t d,
                                                         bf16* __restrict__ out, int64_t ldo) {
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
      out[c * ldo + static_cast<int64_t>(q) * r + a] = __float2bfloat16_rn(acc);
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

struct VoWs {
  bf16* c_planes;   // [3][d x dp]
  bf16* wvt;        // [d x vp]       W_v^T
  float* p;         // [d x vp]       (C + rho I) W_v^T
  bf16* p_planes;   // [3][d x vp]
  float* g1;        // [KV, hd, hd]
  float* g2;        // [H, hd, hd]
  float* scratch;   // [KV][2][hd*hd]
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
  Carver c(ptr);
  VoWs w{};
  w.c_planes = c.take<bf16>(kPlanes * d * dp);
  w.wvt = c.take<bf16>(d * vp);
  w.p = c.take<float>(d * vp);
  w.p_planes = c.take<bf16>(kPlanes * d * vp);
  w.g1 = c.take<float>(static_cast<size_t>(KV) * hd * hd);
  w.g2 = c.take<float>(static_cast<size_t>(H) * hd * hd);
  w.scratch = c.take<float>(static_cast<size_t>(KV) * 2 * hd * hd);
  w.rv = c.take<float>(static_cast<size_t>(KV) * hd * hd);
  w.ro = c.take<float>(static_cast<size_t>(KV) * hd * hd);
  w.bytes = c.off + 256;
  return w;
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
                  void* ws, size_t ws_bytes, void* stream) {
  if (!Cx || !Wv || !Wo || !ws) return -1;
  if (d <= 0 || n_heads <= 0 || n_kv_heads <= 0 || n_heads % n_kv_heads) return -2;
  if (hd != 32 && hd != 64 && hd != 128) return -6;
  if (ldc < d || ldwv < d || ldwo < static_cast<int64_t>(n_heads) * hd) return -7;
  VoWs w = carve_vo(ws, d, n_heads, n_kv_heads, hd);
  if (ws_bytes < w.bytes) return -10;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t dp = mg::round_up(d, 64);
  const int64_t nv = static_cast<int64_t>(n_kv_heads) * hd, vp = mg::round_up(nv, 64);
  const int group = n_heads / n_kv_heads;
  const bool mha = group == 1;
  const bf16* wv = static_cast<const bf16*>(Wv);
  const bf16* wo = static_cast<const bf16*>(Wo);
  int rc;

  // planes of C + rho I (full symmetric matrix) and W_v^T
  if ((rc = mg::split_planes(Cx, ldc, d, d, w.c_planes, dp, d * dp, false, nullptr, s, ridge)))
    return rc;
  transpose_bf16_kernel<<<dim3(static_cast<unsigned>((d + 31) / 32),
                               static_cast<unsigned>((nv + 31) / 32)),
                          256, 0, s>>>(wv, ldwv, nv, d, w.wvt, vp);
  if ((rc = cuda_rc())) return rc;
  // P = (C + rho I) W_v^T   [d, nv]
  {
    mg::GemmArgs g{};
    g.A = w.c_planes;
    g.lda = dp;
    g.a_plane_stride = d * dp;
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
    g.D = w.p;
    g.ldd = vp;
    g.alpha = 1.f;
    g.tiles = mg::TILES_FULL;
    g.epi = mg::EPI_STORE;
    g.ksplit = 1;
    if ((rc = mg::gemm_tn_launch(g, s))) return rc;
  }
  if ((rc = mg::split_planes(w.p, vp, d, nv, w.p_planes, vp, d * vp, false, nullptr, s))) return rc;
  // G1[h] = W_v,h P_h  (diagonal hd x hd blocks of W_v P)
  cudaMemsetAsync(w.g1, 0, sizeof(float) * n_kv_heads * hd * hd, s);
  {
    mg::GemmArgs g{};
    g.A = w.wvt;
    g.lda = vp;
    g.a_planes = 1;
    g.B = w.p_planes;
    g.ldb = vp;
    g.b_plane_stride = d * vp;
    g.b_planes = kPlanes;
    g.npairs = 3;
    for (int i = 0; i < 3; ++i) {
      g.pair_a[i] = 0;
      g.pair_b[i] = i;
    }
    g.M = g.N = nv;
    g.K = d;
    g.D = w.g1;
    g.ldd = hd;
    g.alpha = 1.f;
    g.tiles = mg::TILES_DIAG;
    g.epi = mg::EPI_ADD;
    g.hd = hd;
    g.ksplit = 0;
    if ((rc = mg::gemm_tn_launch(g, s))) return rc;
  }
  if (mha) {
    cudaMemsetAsync(w.g2, 0, sizeof(float) * n_heads * hd * hd, s);
    rc = mg_syrk_heads_bf16_f32(wo, d, static_cast<int64_t>(n_heads) * hd, ldwo, hd, w.g2, 1.f, 1,
                                stream);
    if (rc) return rc;
  }
  return 0;
}

int mg_vo_finish(const void* Wv, int64_t ldwv, const void* Wo, int64_t ldwo, int n_heads,
                 int n_kv_heads, int hd, int64_t d, int r, void* Wv_out, int64_t ldv_out,
                 void* Wo_out, int64_t ldo_out, void* ws, size_t ws_bytes, void* stream) {
  if (!Wv || !Wo || !Wv_out || !Wo_out || !ws) return -1;
  if (d <= 0 || n_heads <= 0 || n_kv_heads <= 0 || n_heads % n_kv_heads) return -2;
  if (hd != 32 && hd != 64 && hd != 128) return -6;
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
  static bool attr_set = false;
  const size_t esm = eig_smem_bytes(hd);
  const size_t osm = sizeof(float) * (hd * r + 64 * (hd + 1));
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(vo_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(eig_smem_bytes(kMaxHd)));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(vo_apply_o_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(sizeof(float) * (kMaxHd * kMaxHd + 64 * (kMaxHd + 1))));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(vo_apply_v_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(sizeof(float) * kMaxHd * kMaxHd));
    if (e != cudaSuccess) return -1000 - static_cast<int>(e);
    attr_set = true;
  }
  vo_factor_kernel<<<n_kv_heads, 1024, esm, s>>>(w.g1, mha ? w.g2 : nullptr, hd, r, w.scratch, w.rv,
                                                 w.ro);
  if ((rc = cuda_rc())) return rc;
  vo_apply_v_kernel<<<dim3(static_cast<unsigned>((d + 63) / 64), n_kv_heads), 256,
                      sizeof(float) * hd * r, s>>>(wv, ldwv, w.rv, hd, r, d,
                                                   static_cast<bf16*>(Wv_out), ldv_out);
  if ((rc = cuda_rc())) return rc;
  vo_apply_o_kernel<<<dim3(static_cast<unsigned>((d + 63) / 64), n_heads), 256, osm, s>>>(
      wo, ldwo, w.ro, group, hd, r, d, static_cast<bf16*>(Wo_out), ldo_out);
  return cuda_rc();
}

int mg_vo_compress(const float* Cx, int64_t ldc, float ridge, const void* Wv, int64_t ldwv,
                   const void* Wo, int64_t ldwo, int n_heads, int n_kv_heads, int hd, int64_t d,
                   int r, void* Wv_out, int64_t ldv_out, void* Wo_out, int64_t ldo_out, void* ws,
                   size_t ws_bytes, void* stream) {
  if (!Wv_out || !Wo_out) return -1;
  if (r <= 0 || r > hd) return -11;
  int rc = mg_vo_prepare(Cx, ldc, ridge, Wv, ldwv, Wo, ldwo, n_heads, n_kv_heads, hd, d, ws, ws_bytes,
                         stream);
  if (rc) return rc;
  return mg_vo_finish(Wv, ldwv, Wo, ldwo, n_heads, n_kv_heads, hd, d, r, Wv_out, ldv_out, Wo_out,
                      ldo_out, ws, ws_bytes, stream);
}

}  // extern "C"
