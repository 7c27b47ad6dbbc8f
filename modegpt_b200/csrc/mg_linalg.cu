// mg_linalg.cu — building blocks of the blocked factorizations:
//   * fp32 -> 3 x bf16 plane splitting (so tensor-core products are fp32-accurate),
//   * the 128-wide diagonal block kernel (fp64 Cholesky + triangular inverse in shared memory),
//   * the right-looking blocked Cholesky driver whose TRSM and trailing SYRK run on tcgen05.
#include "mg_linalg.cuh"

#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <utility>
#include <vector>

#include "mg_gemm.cuh"
#include "mg_once.cuh"
#include "mg_prof.cuh"
#include "mg_ptx.cuh"

namespace mg {

__device__ long long g_dbg_clk[64];
#define MG_CLK(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_dbg_clk[slot] = clock64(); } while (0)

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

// 1/sqrt(x): the 64-bit MUFU seed (computed from the upper word, relative error e below 2^-17)
// and one second-order correction y0 (1 + e/2), e = 1 - x y0^2.  The neglected term 3 e^2 / 8 is
// below 3e-11 — the panels are stored in fp32 (6e-8) — and the chain is MUFU + three dependent fp64
// operations with no float conversions; the pivot loops sit on this latency.  Non-positive or
// non-finite x gives NaN / inf, which the caller detects after the loop.
__device__ __forceinline__ double rsqrt_fast(double x) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
  const double h = 0.5 * y0;
  const double e = fma(-(x * y0), y0, 1.0);
  return fma(h, e, y0);
}

// named barriers (ids 1..15; 0 is __syncthreads): `count` threads in total arrive or wait
__device__ __forceinline__ void named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// D (8 x 8, fp64) += A (8 x 4, row) * B (4 x 8, col) on the FP64 tensor cores; per-thread fragments:
// a = A[lane >> 2][lane & 3], b = B[lane & 3][lane >> 2], (d0, d1) = D[lane >> 2][2 (lane & 3) + {0, 1}].
__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
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
                                                           float diag_add, int upper_only) {
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
    if (r < rows && c < cols && (!upper_only || c >= r))
      x = src[r * ld_src + c] + (r == c ? diag_add : 0.f);
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
//       pivot row broadcast by shuffles) and leaves a pre-scaled copy U'[m][i] = U[m][i] / U[i][i],
//   (2) one thread per remaining column does the 32-step forward substitution (one fused
//       multiply-add per step on the chain, thanks to U'),
//   (3) the rank-32 update of the trailing part of the block on the FP64 tensor cores.
// The kernel sits on the critical path of every blocked factorisation, so the sub-panels are
// software-pipelined around warp 0: it solves the 32 columns of the NEXT diagonal block alone,
// updates that block together with warps 1..3 and goes straight on to (1) of the next sub-panel;
// the other fifteen warps solve the remaining columns, apply the rest of the update and write
// finished rows out behind it.  Chain per sub-panel: (1) + (2) of one warp + ten 8 x 8 tiles on
// four warps (tools/gpu_probe_clocks.py prints the stamps; tools/ubench_fp64*.cu the instruction
// latencies and rates the layout follows from).
// Named barriers: 4 = "(2) done by warp 0" and 1 = "(2) done by everyone" (warp 0 only arrives),
// 3 = "next diagonal block updated" (warps 1..3 arrive, warp 0 waits).
// ---------------------------------------------------------------------------------------------
constexpr int kDiagLd = kNB + 4;   // even: rows stay 16-byte aligned for double2 accesses; 4 rows x
                                   // 4 consecutive doubles (an mma fragment per half-warp) hit 16 banks
constexpr int kPotrfThreads = 512;
// a[128][132] | invd[2][32] | urow[2][32] | uscaled[2][32][32]
constexpr size_t kPotrfSmem = sizeof(double) * (kNB * kDiagLd + 128 + 2 * 32 * 32);

__global__ void __launch_bounds__(kPotrfThreads, 1)
    potrf128_kernel(float* __restrict__ A, int64_t ld, int64_t j0, int nb,
                    float* __restrict__ t_fwd, float* __restrict__ t_bwd, int* __restrict__ info) {
  extern __shared__ __align__(16) double sm[];
  double* a = sm;                     // [128][132], upper triangle live
  double* invd = sm + kNB * kDiagLd;  // [2][32] reciprocal pivots, by sub-panel parity
  double* urow = invd + 64;           // [2][32] pivot row exchange of warp 0 (16-byte aligned)
  double* usc = invd + 128;           // [2][32][32] pre-scaled diagonal blocks, by sub-panel parity
  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  const int fm = lane & 3, fr = lane >> 2;   // mma.m8n8k4 fragment coordinates

  MG_CLK(21);
  // All loads of a thread are in flight at once: under the bulk GEMMs of the other lanes a round
  // trip to L2 / HBM costs a microsecond, and this kernel is the head of the chain.  A full block
  // with 16-byte aligned rows is read as eight 16-byte loads per thread (the lower triangle comes
  // along and is dropped); anything else — the short last block, odd leading dimensions — as 32
  // predicated scalar loads.
  if (nb == kNB && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0) {
    constexpr int kPer = kNB * kNB / 4 / kPotrfThreads;   // 8
    float4 v[kPer];
    const int k = 4 * (t & 31), i0 = t >> 5;              // rows i0, i0 + 16, ...
    const float* src = A + (j0 + i0) * ld + j0 + k;
#pragma unroll
    for (int q = 0; q < kPer; ++q) v[q] = *reinterpret_cast<const float4*>(src + q * 16 * ld);
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int i = i0 + 16 * q;
      double* dst = a + i * kDiagLd + k;
      *reinterpret_cast<double2*>(dst) =
          make_double2(i <= k ? static_cast<double>(v[q].x) : 0.0, i <= k + 1 ? static_cast<double>(v[q].y) : 0.0);
      *reinterpret_cast<double2*>(dst + 2) = make_double2(i <= k + 2 ? static_cast<double>(v[q].z) : 0.0,
                                                          i <= k + 3 ? static_cast<double>(v[q].w) : 0.0);
    }
  } else {
    constexpr int kPer = kNB * kNB / kPotrfThreads;
    float v[kPer];
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int e = t + q * kPotrfThreads, i = e >> 7, k = e & 127;
      v[q] = (i == k) ? 1.f : 0.f;    // identity padding for a short last block
      if (i < nb && k < nb && i <= k) v[q] = A[(j0 + i) * ld + (j0 + k)];
    }
#pragma unroll
    for (int q = 0; q < kPer; ++q) {
      const int e = t + q * kPotrfThreads;
      a[(e >> 7) * kDiagLd + (e & 127)] = static_cast<double>(v[q]);
    }
  }
  MG_CLK(0);
  __syncthreads();
  MG_CLK(1);

  // Outputs for rows [r0, r1) of U by `nth` threads (`tid` = 0..nth-1): the fp32 factor back into A,
  //   forward block   T[m][i] = U[m][i], identity beyond nb;
  //   backward block  T'[m'][i'] = U[nb-1-i'][nb-1-m'] for m' <= i' < nb, identity beyond: row r of
  //                   U is column nb-1-r of T' (a padding row r >= nb is column r of the identity).
  auto write_rows = [&](int r0, int r1, int tid, int nth) {
    for (int e = tid; e < (r1 - r0) * kNB; e += nth) {
      const int i = r0 + (e >> 7), k = e & 127;
      const bool in = i < nb && k < nb;
      const float x = (in && i <= k) ? static_cast<float>(a[i * kDiagLd + k]) : 0.f;
      if (in && i <= k) A[(j0 + i) * ld + (j0 + k)] = x;
      t_fwd[i * kTLd + k] = in ? x : (i == k ? 1.f : 0.f);
    }
    if (t_bwd == nullptr) return;
    const int nr = r1 - r0;          // a multiple of 32: consecutive threads -> consecutive columns
    for (int e = tid; e < nr * kNB; e += nth) {
      const int mp = e / nr, r = r0 + (e - mp * nr);
      float x;
      int ip;
      if (r < nb) {
        ip = nb - 1 - r;
        x = (mp <= ip) ? static_cast<float>(a[r * kDiagLd + (nb - 1 - mp)]) : 0.f;
      } else {
        ip = r;
        x = (mp == ip) ? 1.f : 0.f;
      }
      t_bwd[mp * kTLd + ip] = x;
    }
  };

  const int nsub = (nb + 31) / 32;
  for (int kb = 0; kb < nsub; ++kb) {
    const int c0 = kb * 32;
    double* inv_s = invd + (kb & 1) * 32;
    double* us = usc + (kb & 1) * 1024;
    // (1) 32 x 32 diagonal block, one warp, lane k owns COLUMN k in registers (rows 0..k).
    //     Step j: lane j's reciprocal pivot is broadcast by shuffle, every lane scales its own
    //     entry of pivot row j (u_jk), folds it into its own diagonal and starts the rsqrt of that
    //     diagonal at once — lane j+1's is the next pivot, so the long-latency rsqrt overlaps the
    //     off-diagonal updates below — then the pivot row is exchanged through a double-buffered
    //     32-entry shared array (one parallel store, broadcast loads).  Nothing on the chain looks
    //     at the sign of a pivot: a non-positive one turns into NaNs, found after the loop.
    if (warp == 0) {
      double col[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) col[i] = a[(c0 + i) * kDiagLd + c0 + lane];   // rows > lane: don't-care
      double diag = col[0];
#pragma unroll
      for (int i = 1; i < 32; ++i) diag = (i == lane) ? col[i] : diag;
      double inv_own = rsqrt_fast(diag);
      double inv_mine = 1.0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const double inv = __shfl_sync(0xffffffffu, inv_own, j);
        // u_jk: own entry of pivot row j (lane j: the diagonal sqrt(d_j); lanes < j: unused)
        const double u = (lane == j ? diag : col[j]) * inv;
        diag = fma(-u, u, diag);         // lanes <= j: dead value from here on
        if (lane == j) {
          inv_s[j] = inv;
          inv_mine = inv;
        }
        double* ur = urow + (j & 1) * 32;
        ur[lane] = u;
        __syncwarp();
        // lane j+1's value is the next pivot's: issued here so that its dependent chain interleaves
        // with the (independent) row updates below
        inv_own = rsqrt_fast(diag);
        // rows j+1 .. 31 of the own column; entries at or below the own diagonal are don't-care,
        // so the pivot row is loaded and applied without predicates (batched 16-byte loads)
#pragma unroll
        for (int i2 = (j + 1) / 2; i2 < 16; ++i2) {
          const double2 pr = *reinterpret_cast<const double2*>(ur + 2 * i2);
          if (2 * i2 > j) col[2 * i2] = fma(-pr.x, u, col[2 * i2]);
          col[2 * i2 + 1] = fma(-pr.y, u, col[2 * i2 + 1]);
        }
        col[j] = u;                      // U[j][k]
      }
      // a pivot that was not a positive finite number poisons every later one (NaN / inf), so the
      // last diagonal entry tells whether the whole sub-panel is good
      const double ulast = __shfl_sync(0xffffffffu, col[31], 31);
      if (!(ulast > 0.0 && ulast < 1e150)) {
        // rare: report the first bad pivot; the sub-panel becomes the identity so that everything
        // downstream stays finite (the caller raises on `info`)
        double ujj = col[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) ujj = (i == lane) ? col[i] : ujj;
        const unsigned bad = __ballot_sync(0xffffffffu, !(ujj > 0.0 && ujj < 1e150));
        const int first = __ffs(bad) - 1;
        if (lane == 0 && c0 + first < nb) atomicCAS(info, 0, static_cast<int>(j0 + c0 + first + 1));
#pragma unroll
        for (int i = 0; i < 32; ++i) col[i] = (i == lane) ? 1.0 : 0.0;
        inv_mine = 1.0;
        inv_s[lane] = 1.0;
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if (i <= lane) a[(c0 + i) * kDiagLd + c0 + lane] = col[i];
        us[i * 32 + lane] = col[i] * inv_mine;   // U'[i][k]; rows > k are never read
      }
    }
    __syncthreads();   // (1) of this sub-panel visible; the other warps' update of the previous one done
    MG_CLK(2 + 3 * kb);
    // (2) block row: forward substitution, one thread per column with the 32 rows in registers:
    //     x_i = b_i / u_ii - sum_{m < i} U'[m][i] x_m.  The coefficients are the same for every
    //     column (broadcast 16-byte loads, independent of the running solution), so per column the
    //     chain is 32 fused multiply-adds; no shuffles, no intermediate barrier.  The phase is bound
    //     by shared-memory bandwidth (256 16-byte loads per warp), so warp 0 — whose 32 columns are
    //     the next diagonal block — runs it alone and the other warps follow behind barrier 4.
    const int rest = kNB - c0 - 32;  // columns right of the sub-panel (padding included)
    if (rest == 0) break;            // last sub-panel of a full block
    auto solve_columns = [&]() {
      if (t >= rest) return;
      const int k = c0 + 32 + t;
      double r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = a[(c0 + i) * kDiagLd + k] * inv_s[i];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const double x = r[i];
        const double* trow = us + i * 32;
#pragma unroll
        for (int l2 = (i + 1) / 2; l2 < 16; ++l2) {
          const double2 tv = *reinterpret_cast<const double2*>(trow + 2 * l2);
          if (2 * l2 > i) r[2 * l2] = fma(-tv.x, x, r[2 * l2]);
          r[2 * l2 + 1] = fma(-tv.y, x, r[2 * l2 + 1]);
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) a[(c0 + i) * kDiagLd + k] = r[i];
    };
    // (3) rank-32 update of the trailing upper triangle on the FP64 tensor cores, A22 -= U12^T U12 in
    //     8 x 8 tiles, eight mma.m8n8k4 per tile (the only fp64 shape sm_100a has; 64 FMA/clk/SM,
    //     the rate of the FMA pipe, for a quarter of its shared-memory traffic).  A warp takes one
    //     tile row: tiles (ti, tk0 .. tk0+cnt-1), the A fragment shared by up to MAXC accumulation
    //     chains.  Fragment coordinates: A[i][m] and B[m][k] are both read from row c0+m of the
    //     block (thread: m = lane & 3, i or k = lane >> 2); C rows lane >> 2, columns 2 (lane & 3).
    auto tile_row = [&](auto maxc, int ti, int tk0, int cnt) {
      constexpr int MAXC = decltype(maxc)::value;
      double acc[MAXC][2];
#pragma unroll
      for (int q = 0; q < MAXC; ++q) acc[q][0] = acc[q][1] = 0.0;
      const double* frag = a + (c0 + fm) * kDiagLd + c0 + 32 + fr;
#pragma unroll
      for (int m0 = 0; m0 < 32; m0 += 4) {
        const double* row = frag + m0 * kDiagLd;
        const double av = row[8 * ti];
#pragma unroll
        for (int q = 0; q < MAXC; ++q)
          if (q < cnt) dmma_m8n8k4(acc[q][0], acc[q][1], av, row[8 * (tk0 + q)]);
      }
      const int r = c0 + 32 + 8 * ti + fr;
#pragma unroll
      for (int q = 0; q < MAXC; ++q) {
        const int colb = c0 + 32 + 8 * (tk0 + q) + 2 * fm;
        if (q < cnt && r <= colb) a[r * kDiagLd + colb] -= acc[q][0];
        if (q < cnt && r <= colb + 1) a[r * kDiagLd + colb + 1] -= acc[q][1];
      }
    };
    const int nt = rest / 8;         // 12, 8 or 4 tile rows; the first 4 are the next diagonal block
    if (warp == 0) {
      solve_columns();
      __syncwarp();
      named_arrive(4, kPotrfThreads);   // "the next diagonal block's columns are solved"
      named_arrive(1, kPotrfThreads);
      MG_CLK(3 + 3 * kb);
      tile_row(std::integral_constant<int, 4>{}, 3, 3, 1);
      named_sync(3, 128);
      MG_CLK(4 + 3 * kb);
      continue;                      // on to (1) of the next sub-panel
    }
    named_sync(4, kPotrfThreads);
    if (warp < 4) {                  // the other three tile rows of the next diagonal block
      tile_row(std::integral_constant<int, 4>{}, warp - 1, warp - 1, 5 - warp);
      named_arrive(3, 128);
    }
    solve_columns();
    named_sync(1, kPotrfThreads);    // every column of the block row is solved
    if (warp <= nt && nt > 4) {      // tile row warp-1, the tiles right of the next diagonal block
      const int ti = warp - 1, tk0 = ti > 4 ? ti : 4;
      tile_row(std::integral_constant<int, 8>{}, ti, tk0, nt - tk0);
    }
    // rows that are final and not yet written go out behind warp 0's next diagonal block; the
    // longest update (first sub-panel) leaves no room for it, the last one has room for two
    if (kb == 1) write_rows(0, 32, t - 32, kPotrfThreads - 32);
    if (kb == 2) write_rows(32, 96, t - 32, kPotrfThreads - 32);
  }
  __syncthreads();

  // ---- the rows not written inside the loop (the last sub-panel; padding rows of a short block)
  MG_CLK(19);
  write_rows(nsub >= 3 ? 96 : (nsub == 2 ? 32 : 0), kNB, t, kPotrfThreads);
  MG_CLK(20);
}

// ---------------------------------------------------------------------------------------------
// trsm128: forward substitution with a compact lower-in-(m,i) block T (x_i = (b_i - sum_{m<i}
// T[m][i] x_m) / T[i][i]).  Four adjacent lanes (a quad) share one right-hand-side column: within
// every 32-row chunk lane `part` owns rows part*8 .. part*8+7 in registers; solved rows are parked
// in shared memory (xs[col][row], float4-readable) for the later chunks, and inside a chunk the
// freshly solved value is broadcast through the quad by shuffle.  `reversed` maps chunk row i' to
// matrix row nb-1-i' (that is how U x = b becomes a forward substitution).
// ---------------------------------------------------------------------------------------------
constexpr int kTrsmLanes = 8;                       // lanes sharing one column pair
constexpr int kTrsmRows = 32 / kTrsmLanes;          // chunk rows owned by a lane (4)
constexpr int kTrsmCpl = 2;                         // columns per lane: every T load feeds 8 FMAs
constexpr int kXsLd = kNB + 4;                      // xs row stride (floats), 16-byte aligned rows
// THREADS per CTA = 128 (32 columns) or 256 (64 columns).  The substitution is bound by instruction
// issue per scheduler, so 4-warp CTAs (one warp per scheduler, twice as many SMs for the same
// columns) halve the time of the short launches on the panel chain; long launches are indifferent.
constexpr int trsm_cols(int threads) { return threads / kTrsmLanes * kTrsmCpl; }
constexpr size_t trsm_smem(int threads) {
  return sizeof(float) * (kTBlock + trsm_cols(threads) * kXsLd + kNB) + 16;
}

template <int kTrsmThreads>
__global__ void __launch_bounds__(kTrsmThreads, 2)
    trsm128_kernel(const float* __restrict__ tblock, int reversed, int nb,
                   const float* __restrict__ B, int64_t ldb, int64_t ncols, float alpha,
                   float* __restrict__ X, int64_t ldx, __nv_bfloat16* __restrict__ planes,
                   int64_t ldp, int64_t pstride, __nv_bfloat16* __restrict__ tplanes,
                   int64_t ldtp, int64_t tpstride, float* __restrict__ colsumsq) {
  constexpr int kTrsmCols = trsm_cols(kTrsmThreads);
  extern __shared__ __align__(16) float smf[];
  float* T = smf;                          // [128][132]
  float* xs = smf + kTBlock;               // [64 cols][132]: xs[col][row]
  float* invd = xs + kTrsmCols * kXsLd;    // [128]
  const int tid = threadIdx.x;
  const int pair_local = tid / kTrsmLanes, part = tid % kTrsmLanes;
  const int lane = tid & 31;
  const int group_base = lane & ~(kTrsmLanes - 1);
  // the two columns of a lane are 32 apart so that a warp's 4 column pairs stay coalesced
  const int cl0 = (pair_local / 4) * 8 + (pair_local % 4);
  const int col_local[kTrsmCpl] = {cl0, cl0 + 4};
  int64_t c[kTrsmCpl];
  bool valid[kTrsmCpl];
#pragma unroll
  for (int q = 0; q < kTrsmCpl; ++q) {
    c[q] = static_cast<int64_t>(blockIdx.x) * kTrsmCols + col_local[q];
    valid[q] = c[q] < ncols;
  }
  const int nchunks = (nb + 31) / 32;

  auto load_chunk = [&](int rb, float (&dst)[kTrsmCpl][kTrsmRows]) {
#pragma unroll
    for (int q = 0; q < kTrsmCpl; ++q)
#pragma unroll
      for (int i = 0; i < kTrsmRows; ++i) {
        const int ip = rb * 32 + part * kTrsmRows + i;
        const int row = reversed ? nb - 1 - ip : ip;
        dst[q][i] = (valid[q] && ip < nb) ? B[static_cast<int64_t>(row) * ldb + c[q]] : 0.f;
      }
  };

  // the triangular block (one contiguous 66 KB record) arrives by a single bulk copy while the
  // threads fetch their first right-hand-side chunk
  uint64_t* tbar = reinterpret_cast<uint64_t*>(invd + kNB);
  if (tid == 0) {
    mbar_init(tbar, 1);
    fence_barrier_init();
    mbar_expect_tx(tbar, static_cast<uint32_t>(sizeof(float) * kTBlock));
    bulk_load_1d(T, tblock, static_cast<uint32_t>(sizeof(float) * kTBlock), tbar);
  }
  float rn[kTrsmCpl][kTrsmRows];
  load_chunk(0, rn);
  MG_CLK(32);
  __syncthreads();                         // barrier initialised before anyone polls it
  mbar_wait(tbar, 0);
  if (tid < kNB) {
    invd[tid] = 1.f / T[tid * kTLd + tid];
    // with a zero diagonal, T[i][ip] vanishes for every ip <= i: the substitution below can apply
    // row i of T to all rows a lane owns without selecting the ones still unsolved
    T[tid * kTLd + tid] = 0.f;
  }
  __syncthreads();
  MG_CLK(33);

  for (int rb = 0; rb < nchunks; ++rb) {
    float r[kTrsmCpl][kTrsmRows];
#pragma unroll
    for (int q = 0; q < kTrsmCpl; ++q)
#pragma unroll
      for (int i = 0; i < kTrsmRows; ++i) r[q][i] = rn[q][i];
    if (rb + 1 < nchunks) load_chunk(rb + 1, rn);   // prefetch the next chunk's right-hand sides
    const int ip0 = rb * 32 + part * kTrsmRows;     // first chunk-order row owned by this lane
    MG_CLK(34 + 4 * rb);
    // contributions of the chunks already solved
    for (int pb = 0; pb < rb; ++pb) {
#pragma unroll 2
      for (int m4 = 0; m4 < 32; m4 += 4) {
        const float4 xa = *reinterpret_cast<const float4*>(xs + col_local[0] * kXsLd + pb * 32 + m4);
        const float4 xb = *reinterpret_cast<const float4*>(xs + col_local[1] * kXsLd + pb * 32 + m4);
        const float xm[kTrsmCpl][4] = {{xa.x, xa.y, xa.z, xa.w}, {xb.x, xb.y, xb.z, xb.w}};
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) {
          const float4 ta = *reinterpret_cast<const float4*>(T + (pb * 32 + m4 + mm) * kTLd + ip0);
#pragma unroll
          for (int q = 0; q < kTrsmCpl; ++q) {
            r[q][0] = fmaf(-ta.x, xm[q][mm], r[q][0]);
            r[q][1] = fmaf(-ta.y, xm[q][mm], r[q][1]);
            r[q][2] = fmaf(-ta.z, xm[q][mm], r[q][2]);
            r[q][3] = fmaf(-ta.w, xm[q][mm], r[q][3]);
          }
        }
      }
    }
    MG_CLK(35 + 4 * rb);
    // the 32 x 32 diagonal chunk: row i is solved by its owner lane and broadcast in the group
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int owner = i / kTrsmRows, li = i % kTrsmRows;
      const int ip = rb * 32 + i;
      const float inv = invd[ip];
      float x[kTrsmCpl];
#pragma unroll
      for (int q = 0; q < kTrsmCpl; ++q) {
        x[q] = __shfl_sync(0xffffffffu, r[q][li] * inv, group_base | owner);
        if (part == owner) {
          r[q][li] = x[q];
          xs[col_local[q] * kXsLd + ip] = x[q];
        }
      }
      {
        // T row loaded unconditionally (it does not depend on x, so later steps' loads run ahead
        // of the substitution chain); its entries at or before column i are zero
        const float4 ta = *reinterpret_cast<const float4*>(T + ip * kTLd + ip0);
        const float tv[4] = {ta.x, ta.y, ta.z, ta.w};
#pragma unroll
        for (int l = 0; l < kTrsmRows; ++l)
#pragma unroll
          for (int q = 0; q < kTrsmCpl; ++q) r[q][l] = fmaf(-tv[l], x[q], r[q][l]);
      }
    }
    __syncwarp();
    MG_CLK(36 + 4 * rb);
  }
  __syncthreads();   // xs[col][ip] now holds every solved (unscaled) entry of this CTA's 64 columns

  // ---- outputs, staged through xs so that every global store is a full-width row segment
  const int64_t cbase = static_cast<int64_t>(blockIdx.x) * kTrsmCols;
  {
    // row-major targets (X, planes): a warp covers 32 consecutive columns of one row
    const int col = tid & (kTrsmCols - 1), rl = tid / kTrsmCols;
    const int64_t cg = cbase + col;
    float ssq = 0.f;
    if (cg < ncols) {
      for (int ip = rl; ip < nb; ip += kTrsmThreads / kTrsmCols) {
        const float xo = alpha * xs[col * kXsLd + ip];
        ssq = fmaf(xo, xo, ssq);
        const int row = reversed ? nb - 1 - ip : ip;
        if (X) X[static_cast<int64_t>(row) * ldx + cg] = xo;
        if (planes) {
          __nv_bfloat16 h, m, l;
          split3(xo, h, m, l);
          const int64_t o = static_cast<int64_t>(row) * ldp + cg;
          planes[o] = h;
          planes[pstride + o] = m;
          planes[2 * pstride + o] = l;
        }
      }
    }
    if (colsumsq) {
      float* red = T;                      // the triangular block is dead: reuse it
      red[rl * kTrsmCols + col] = ssq;
      __syncthreads();
      if (rl == 0 && cg < ncols) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < kTrsmThreads / kTrsmCols; ++q) v += red[q * kTrsmCols + col];
        atomicAdd(colsumsq + cg, v);
      }
    }
  }
  if (tplanes) {
    // transposed targets ([col][row]): a warp covers 32 consecutive rows of one column
    const int ip = tid & (kNB - 1), ch = tid / kNB;
    if (ip < nb) {
      const int row = reversed ? nb - 1 - ip : ip;
      for (int col = ch; col < kTrsmCols; col += kTrsmThreads / kNB) {
        const int64_t cg = cbase + col;
        if (cg >= ncols) break;
        __nv_bfloat16 h, m, l;
        split3(alpha * xs[col * kXsLd + ip], h, m, l);
        const int64_t o = cg * ldtp + row;
        tplanes[o] = h;
        tplanes[tpstride + o] = m;
        tplanes[2 * tpstride + o] = l;
      }
    }
  }
  MG_CLK(50);
}

}  // namespace

int split_planes(const float* src, int64_t ld_src, int64_t rows, int64_t cols, __nv_bfloat16* dst,
                 int64_t ld_dst, int64_t plane_stride, bool transpose, float* colsumsq,
                 cudaStream_t s, float diag_add, bool upper_only) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid(static_cast<unsigned>((cols + 63) / 64), static_cast<unsigned>((rows + 31) / 32));
  if (transpose)
    split_planes_kernel<true><<<grid, 256, 0, s>>>(src, ld_src, rows, cols, dst, ld_dst,
                                                   plane_stride, colsumsq, diag_add, upper_only);
  else
    split_planes_kernel<false><<<grid, 256, 0, s>>>(src, ld_src, rows, cols, dst, ld_dst,
                                                    plane_stride, colsumsq, diag_add, upper_only);
  return cuda_rc();
}

int potrf128(float* A, int64_t ld, int64_t j0, int nb, float* t_fwd, float* t_bwd, int* info,
             cudaStream_t s) {
  static PerDeviceOnce once;
  if (int rc = once.run([] {
        return cudaFuncSetAttribute(potrf128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(kPotrfSmem));
      }))
    return rc;
  potrf128_kernel<<<1, kPotrfThreads, kPotrfSmem, s>>>(A, ld, j0, nb, t_fwd, t_bwd, info);
  return cuda_rc();
}

template <int THREADS>
static int trsm128_launch(const float* tblock, bool reversed, int nb, const float* B, int64_t ldb,
                          int64_t ncols, float alpha, float* X, int64_t ldx, __nv_bfloat16* planes,
                          int64_t ldp, int64_t pstride, __nv_bfloat16* tplanes, int64_t ldtp,
                          int64_t tpstride, float* colsumsq, cudaStream_t s) {
  static PerDeviceOnce once;
  if (int rc = once.run([] {
        return cudaFuncSetAttribute(trsm128_kernel<THREADS>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(trsm_smem(THREADS)));
      }))
    return rc;
  constexpr int cols = trsm_cols(THREADS);
  const unsigned grid = static_cast<unsigned>((ncols + cols - 1) / cols);
  trsm128_kernel<THREADS><<<grid, THREADS, trsm_smem(THREADS), s>>>(
      tblock, reversed ? 1 : 0, nb, B, ldb, ncols, alpha, X, ldx, planes, ldp, pstride, tplanes, ldtp,
      tpstride, colsumsq);
  return cuda_rc();
}

int trsm128(const float* tblock, bool reversed, int nb, const float* B, int64_t ldb, int64_t ncols,
            float alpha, float* X, int64_t ldx, __nv_bfloat16* planes, int64_t ldp, int64_t pstride,
            __nv_bfloat16* tplanes, int64_t ldtp, int64_t tpstride, float* colsumsq,
            cudaStream_t s) {
  if (ncols <= 0) return 0;
  // MG_TRSM_THREADS=256 selects the 8-warp CTAs (A/B measurements)
  static const bool wide = [] {
    const char* e = std::getenv("MG_TRSM_THREADS");
    return e != nullptr && std::atoi(e) == 256;
  }();
  return wide ? trsm128_launch<256>(tblock, reversed, nb, B, ldb, ncols, alpha, X, ldx, planes, ldp,
                                    pstride, tplanes, ldtp, tpstride, colsumsq, s)
              : trsm128_launch<128>(tblock, reversed, nb, B, ldb, ncols, alpha, X, ldx, planes, ldp,
                                    pstride, tplanes, ldtp, tpstride, colsumsq, s);
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

// ---------------------------------------------------------------------------------------------
// lanes
// ---------------------------------------------------------------------------------------------
namespace {
struct DeviceLanes {
  std::mutex mu;
  bool tried = false, ok = false;
  Lanes proto;
};

// Two independent lane sets per device: two host threads can run two factorisations side by side
// (different layers; each chain is latency-bound and leaves most SMs idle).
constexpr int kLaneSets = 3;

int lane_slot_for(cudaStream_t user) {
  static std::mutex mu;
  static std::vector<std::pair<cudaStream_t, int>> seen;
  static int next = 0;
  std::lock_guard<std::mutex> g(mu);
  for (const auto& e : seen)
    if (e.first == user) return e.second;
  const int slot = next++ % kLaneSets;
  if (seen.size() > 64) seen.clear();     // streams come and go: forget, reassign
  seen.emplace_back(user, slot);
  return slot;
}

DeviceLanes& device_lanes(int slot) {
  static DeviceLanes per_dev[64][kLaneSets];
  int dev = 0;
  cudaGetDevice(&dev);
  return per_dev[dev & 63][slot];
}

int chol_outer_panels() {
  static const int g = [] {
    const char* e = std::getenv("MG_CHOL_OUTER");
    const int v = e ? std::atoi(e) : 4;
    return v < 1 ? 1 : (v > 8 ? 8 : v);
  }();
  return g;
}

bool lanes_disabled() {
  static const bool off = [] {
    const char* e = std::getenv("MG_SERIAL");
    return e && e[0] == '1';
  }();
  return off;
}
}  // namespace

// Number of factorisations the caller runs side by side (mg_set_concurrent_factorizations): the
// SMs left after the chain reserve are divided between their bulk GEMMs, so that no persistent
// grid can fill the GPU and starve the OTHER factorisation's latency-bound chain kernels.
std::atomic<int> g_lane_share{1};

int Lanes::bulk_cta_cap() const {
  if (serial) return 0;
  const int share = g_lane_share.load(std::memory_order_relaxed);
  // Side by side, the factorisations are bound by their bulk GEMMs on a share of the SMs, not by
  // one chain: measured with two workers at n = 11008, 16.9 / 17.4 / 17.9 / 18.6 / 20.2 ms per
  // layer for 16 / 32 / 48 / 64 / 80 reserved SMs (one worker: 19.8 / 19.4 / 19.0 / 19.4 / 20.8);
  // 8 and 0 are within noise of 16 (profiles/r2_mlp_workers.txt).
  const int reserve = (share >= 2 && !reserve_from_env && reserve_sms > 16) ? 16 : reserve_sms;
  int cap = (device_sm_count() - reserve) / (share < 1 ? 1 : share);
  cap &= ~1;                       // CTA pairs
  return cap < 16 ? 16 : cap;
}

void set_lane_share(int n) { g_lane_share.store(n < 1 ? 1 : (n > 4 ? 4 : n), std::memory_order_relaxed); }

LaneScope::LaneScope(cudaStream_t user, int64_t n) {
  // measured at n = 11008: Nystrom 12.8 -> 11.7 ms going from 4 to 52 reserved SMs (chain-bound);
  // at n = 28672 the bulk GEMMs are the bound (574 TFLOP/s on 100 CTAs) and want the SMs back
  static const int reserve_env = [] {
    const char* e = std::getenv("MG_BULK_RESERVE_SMS");
    return e ? std::atoi(e) : -1;
  }();
  const int reserve = reserve_env >= 0 ? reserve_env : (n > 16384 ? 8 : 48);
  lanes_.user = lanes_.chain = lanes_.chain2 = lanes_.upd = lanes_.tri = lanes_.tri2 = user;
  lanes_.serial = true;
  if (lanes_disabled()) return;
  // One lane set per CALLER STREAM (assigned round-robin the first time a stream is seen): two host
  // threads that decompose different layers on different streams then never share internal streams.
  // (Picking whichever set happened to be unlocked at enqueue time — round 1 — put both threads on
  // set 0 most of the time: a factorisation is enqueued in 3 ms and runs for 10, so the lock was
  // usually free, and the second factorisation queued behind the first on the same streams.)
  DeviceLanes& d = device_lanes(lane_slot_for(user));
  d.mu.lock();
  lock_ = &d;
  if (!d.tried) {
    d.tried = true;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // lo = least urgent, hi = most urgent
    bool ok = cudaStreamCreateWithPriority(&d.proto.chain, cudaStreamNonBlocking, hi) == cudaSuccess;
    ok = ok && cudaStreamCreateWithPriority(&d.proto.chain2, cudaStreamNonBlocking, hi) == cudaSuccess;
    ok = ok && cudaStreamCreateWithPriority(&d.proto.upd, cudaStreamNonBlocking, lo) == cudaSuccess;
    ok = ok && cudaStreamCreateWithPriority(&d.proto.tri, cudaStreamNonBlocking, lo) == cudaSuccess;
    ok = ok && cudaStreamCreateWithPriority(&d.proto.tri2, cudaStreamNonBlocking, lo) == cudaSuccess;
    cudaEvent_t* evs[] = {&d.proto.fork, &d.proto.join[0], &d.proto.join[1], &d.proto.join[2],
                          &d.proto.trsm, &d.proto.upd_done[0], &d.proto.upd_done[1],
                          &d.proto.misc[0], &d.proto.misc[1], &d.proto.diag_done[0],
                          &d.proto.diag_done[1], &d.proto.join[3], &d.proto.row_done[0],
                          &d.proto.row_done[1], &d.proto.bulk_done[0], &d.proto.bulk_done[1],
                          &d.proto.next_done[0], &d.proto.next_done[1], &d.proto.join[4], &d.proto.potrf,
                          &d.proto.first[0], &d.proto.first[1], &d.proto.row_rest[0], &d.proto.row_rest[1]};
    for (cudaEvent_t* e : evs)
      ok = ok && cudaEventCreateWithFlags(e, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) cudaGetLastError();
    d.ok = ok;
  }
  if (!d.ok) return;
  lanes_ = d.proto;
  lanes_.user = user;
  lanes_.serial = false;
  lanes_.reserve_sms = reserve;
  lanes_.reserve_from_env = reserve_env >= 0;
  cudaEventRecord(lanes_.fork, user);
  cudaStreamWaitEvent(lanes_.chain, lanes_.fork, 0);
  cudaStreamWaitEvent(lanes_.upd, lanes_.fork, 0);
  cudaStreamWaitEvent(lanes_.tri, lanes_.fork, 0);
  cudaStreamWaitEvent(lanes_.tri2, lanes_.fork, 0);
  cudaStreamWaitEvent(lanes_.chain2, lanes_.fork, 0);
}

LaneScope::~LaneScope() {
  if (!lanes_.serial) {
    cudaStream_t ls[5] = {lanes_.chain, lanes_.upd, lanes_.tri, lanes_.tri2, lanes_.chain2};
    for (int i = 0; i < 5; ++i) {
      cudaEventRecord(lanes_.join[i], ls[i]);
      cudaStreamWaitEvent(lanes_.user, lanes_.join[i], 0);
    }
  }
  if (lock_) static_cast<DeviceLanes*>(lock_)->mu.unlock();
}

// ---------------------------------------------------------------------------------------------
// blocked Cholesky
// ---------------------------------------------------------------------------------------------
int CholStepper::step_panelwise(int64_t pj) const {
  const Lanes& L = *lanes;
  const int64_t np = ws.n_pad;
  const int64_t pstride = np * np;
  const int64_t j0 = pj * kNB;
  const int nb = static_cast<int>(n - j0 < kNB ? n - j0 : kNB);
  float* tf = ws.t_fwd + pj * kTBlock;
  float* tb = ws.t_bwd ? ws.t_bwd + pj * kTBlock : nullptr;   // only back-substitutions need it
  int rc;
  MG_TIMED(L.chain, "chol.potrf128",
           rc = potrf128(A, ld, j0, nb, tf, tb, info, L.chain));
  if (rc) return rc;
  const int64_t rest = n - j0 - nb;
  if (rest <= 0) {
    L.record(L.trsm, L.chain);
    return 0;
  }
  // block row: U12 = U11^-T A12 (in place) + its bf16 planes (and those of U12^T)
  float* a12 = A + j0 * ld + (j0 + nb);
  __nv_bfloat16* u12 = ws.u_planes + j0 * np + (j0 + nb);
  __nv_bfloat16* l21 = ws.l_planes ? ws.l_planes + (j0 + nb) * np + j0 : nullptr;
  MG_TIMED(L.chain, "chol.trsm128",
           rc = trsm128(tf, false, nb, a12, ld, rest, 1.f, a12, ld, u12, np, pstride, l21, np, pstride,
                        nullptr, L.chain));
  if (rc) return rc;
  L.record(L.trsm, L.chain);

  // trailing update A22 -= U12^T U12, split into the next block row (chain lane: it is all the
  // next panel needs) and the rows below it (upd lane, hidden behind the next panel)
  GemmArgs t{};
  t.lda = t.ldb = np;
  t.a_plane_stride = t.b_plane_stride = pstride;
  t.a_planes = t.b_planes = kPlanes;
  set_pairs6(t);
  t.K = nb;
  t.ldd = ld;
  t.alpha = -1.f;
  t.epi = EPI_ADD;
  t.ksplit = 1;
  if (L.serial) {
    t.A = t.B = u12;
    t.M = t.N = rest;
    t.D = A + (j0 + nb) * ld + (j0 + nb);
    t.tiles = TILES_UPPER;
    MG_TIMED(L.chain, "chol.trailing_syrk", rc = gemm_tn_launch(t, L.chain));
    return rc;
  }
  const int64_t m1 = rest < kNB ? rest : kNB;
  const int64_t rest2 = rest - m1;
  if (rest2 > 0) {
    GemmArgs r = t;
    r.A = r.B = u12 + m1;
    r.M = r.N = rest2;
    r.D = A + (j0 + nb + m1) * ld + (j0 + nb + m1);
    r.tiles = TILES_UPPER;
    r.max_ctas = L.bulk_cta_cap();
    L.wait(L.upd, L.trsm);
    MG_TIMED(L.upd, "chol.trailing_syrk", rc = gemm_tn_launch(r, L.upd));
    if (rc) return rc;
    L.record(L.upd_done[pj & 1], L.upd);
  }
  // the update of panel pj-1 reaches block row pj+1 too: let it land first (ordered adds)
  if (pj >= 1) L.wait(L.chain, L.upd_done[(pj - 1) & 1]);
  t.A = u12;
  t.B = u12;
  t.M = m1;
  t.N = rest;
  t.D = A + (j0 + nb) * ld + (j0 + nb);
  t.tiles = TILES_FULL;   // 128 x rest: the strict lower part of its diagonal block is scratch
  MG_TIMED(L.chain, "chol.row_update", rc = gemm_tn_launch(t, L.chain));
  return rc;
}

// Two-level blocking: panels are grouped into outer blocks of G (MG_CHOL_OUTER, default 4).  Inside
// a block a panel only updates the block's remaining rows (K = 128, small); the rows below the
// block get ONE update per block with K = G * 128 — a quarter of the read-modify-write passes over
// the trailing matrix and four times the MMA work per epilogue (the K = 128 trailing update is
// paced by its L2 reduce-adds: 44 % tensor pipe in ncu, 75 % at K = 512).
//
// Split-chain step (lanes on, outer blocks of G > 1 panels).  What the NEXT panel's potrf128 needs
// from panel pj is tiny: the first 384 solved columns of block row pj and the update of one
// 128 x 384 tile.  Only those stay on the chain lane; the rest of the block row's solve and of the
// row updates run on chain2, concurrently with the next potrf128:
//   chain : potrf128(pj) -> trsm128(first 384 cols) -> tile update (next block row x 384 cols)
//   chain2: trsm128(remaining cols) -> [block row pj final: records lanes.trsm] -> remaining row
//           updates of the outer block (K = 128)            [records row_rest[pj & 1]]
//   upd   : at the end of an outer block, the K = G*128 update of everything below it
// Ordered adds: chain waits for chain2's row updates of panel pj-1 before its own tile update
// (they overlap on block row pj+1), both wait for the upd lane's block update of their rows.
int CholStepper::step(int64_t pj) const {
  const Lanes& L = *lanes;
  const int G = chol_outer_panels();
  if (L.serial || G <= 1) return step_panelwise(pj);
  const int64_t np = ws.n_pad;
  const int64_t pstride = np * np;
  const int64_t j0 = pj * kNB;
  const int nb = static_cast<int>(n - j0 < kNB ? n - j0 : kNB);
  float* tf = ws.t_fwd + pj * kTBlock;
  float* tb = ws.t_bwd ? ws.t_bwd + pj * kTBlock : nullptr;
  int rc;
  MG_TIMED(L.chain, "chol.potrf128", rc = potrf128(A, ld, j0, nb, tf, tb, info, L.chain));
  if (rc) return rc;
  const int64_t rest = n - j0 - nb;
  if (rest <= 0) {
    L.record(L.trsm, L.chain);
    return 0;
  }
  L.record(L.potrf, L.chain);
  const int64_t c0 = j0 + nb;                       // first column / row below the panel
  float* a12 = A + j0 * ld + c0;
  __nv_bfloat16* u12 = ws.u_planes + j0 * np + c0;
  __nv_bfloat16* l21 = ws.l_planes ? ws.l_planes + c0 * np + j0 : nullptr;
  // the chain's update tile spans the next diagonal block plus the 256 columns the next panel
  // solves first; its operands must all come from the chain's own solve, so that solve covers
  // the same 384 columns
  const int64_t tw = rest < 384 ? rest : 384;
  const int64_t sw = tw;
  const int64_t fr = rest < kNB ? rest : kNB;       // rows of the next block row
  // Block row pj, columns [c0, c0 + 384): the first 256 received panel pj-1's update from the
  // chain's own tile update, the last 128 from chain2's row update of panel pj-1 — a cross-lane
  // dependency.  (Round 1 waited for it only before the tile update below; the solve ran ahead of
  // it, protected by nothing but the 70 us of potrf128 in between — it broke once two
  // factorisations shared the GPU.)
  if (pj >= 1) L.wait(L.chain, L.row_rest[(pj - 1) & 1]);
  MG_TIMED(L.chain, "chol.trsm128_first",
           rc = trsm128(tf, false, nb, a12, ld, sw, 1.f, a12, ld, u12, np, pstride, l21, np, pstride,
                        nullptr, L.chain));
  if (rc) return rc;
  L.record(L.first[pj & 1], L.chain);
  L.wait(L.chain2, L.potrf);
  if (rest > sw) {
    MG_TIMED(L.chain2, "chol.trsm128_rest",
             rc = trsm128(tf, false, nb, a12 + sw, ld, rest - sw, 1.f, a12 + sw, ld, u12 + sw, np, pstride,
                          l21 ? l21 + sw * np : nullptr, np, pstride, nullptr, L.chain2));
    if (rc) return rc;
  }
  L.wait(L.chain2, L.first[pj & 1]);
  L.record(L.trsm, L.chain2);                       // block row pj of U is final

  GemmArgs t{};
  t.lda = t.ldb = np;
  t.a_plane_stride = t.b_plane_stride = pstride;
  t.a_planes = t.b_planes = kPlanes;
  set_pairs6(t);
  t.ldd = ld;
  t.alpha = -1.f;
  t.epi = EPI_ADD;
  t.ksplit = 1;
  t.tiles = TILES_FULL;                              // below-diagonal parts of a block row are scratch

  const int64_t q = pj % G, ob = pj / G;
  const int64_t o0 = ob * G * kNB;
  const int64_t o_end = (o0 + G * kNB < n) ? o0 + G * kNB : n;
  const int64_t in_block = o_end - c0;               // rows of this outer block below panel pj
  float* d0 = A + c0 * ld + c0;
  if (in_block > 0) {
    // ---- inside the block: K = 128 updates of the block's remaining rows
    t.K = nb;
    if (q == 0 && ob >= 1) {
      L.wait(L.chain, L.next_done[(ob - 1) & 1]);
      L.wait(L.chain2, L.next_done[(ob - 1) & 1]);
    }
    if (pj >= 1) L.wait(L.chain, L.row_rest[(pj - 1) & 1]);
    t.A = u12;
    t.B = u12;
    t.M = fr;
    t.N = tw;
    t.D = d0;
    MG_TIMED(L.chain, "chol.tile_update", rc = gemm_tn_launch(t, L.chain));
    if (rc) return rc;
    if (rest > tw) {                                  // next block row, columns right of the tile
      GemmArgs r = t;
      r.B = u12 + tw;
      r.N = rest - tw;
      r.D = d0 + tw;
      MG_TIMED(L.chain2, "chol.row_update", rc = gemm_tn_launch(r, L.chain2));
      if (rc) return rc;
    }
    if (in_block > fr) {                              // the block's rows below the next block row
      GemmArgs r = t;
      r.A = u12 + fr;
      r.B = u12 + fr;
      r.M = in_block - fr;
      r.N = rest - fr;
      r.D = d0 + fr * ld + fr;
      MG_TIMED(L.chain2, "chol.row_update", rc = gemm_tn_launch(r, L.chain2));
      if (rc) return rc;
    }
    L.record(L.row_rest[pj & 1], L.chain2);
    return 0;
  }
  // ---- panel pj closes its block: rows >= o_end (= c0) get the whole block at once
  const __nv_bfloat16* blk = ws.u_planes + o0 * np;  // block rows o0 .. o_end of U
  t.K = o_end - o0;
  const int64_t next_rows = rest < G * kNB ? rest : G * kNB;
  L.wait(L.upd, L.trsm);
  if (next_rows > fr) {                               // rows 2..G of the next block
    GemmArgs r = t;
    r.A = blk + (c0 + fr);
    r.B = blk + (c0 + fr);
    r.M = next_rows - fr;
    r.N = rest - fr;
    r.D = d0 + fr * ld + fr;
    r.max_ctas = L.bulk_cta_cap();
    MG_TIMED(L.upd, "chol.block_update_next", rc = gemm_tn_launch(r, L.upd));
    if (rc) return rc;
  }
  L.record(L.next_done[ob & 1], L.upd);
  if (rest > next_rows) {                             // everything below the next block
    GemmArgs r = t;
    r.A = r.B = blk + (c0 + next_rows);
    r.M = r.N = rest - next_rows;
    r.D = d0 + next_rows * ld + next_rows;
    r.tiles = TILES_UPPER;
    r.max_ctas = L.bulk_cta_cap();
    MG_TIMED(L.upd, "chol.trailing_syrk", rc = gemm_tn_launch(r, L.upd));
    if (rc) return rc;
  }
  L.record(L.upd_done[ob & 1], L.upd);
  // the next block's first row: tile on the chain, the rest of the row on chain2.  Both read every
  // block row of this outer block: the earlier ones were finished by chain2 (row_rest of pj-1
  // follows their solves in stream order), this one by the solves above.
  if (ob >= 1) {
    L.wait(L.chain, L.upd_done[(ob - 1) & 1]);
    L.wait(L.chain2, L.upd_done[(ob - 1) & 1]);
  }
  if (pj >= 1) L.wait(L.chain, L.row_rest[(pj - 1) & 1]);
  t.A = blk + c0;
  t.B = blk + c0;
  t.M = fr;
  t.N = tw;
  t.D = d0;
  MG_TIMED(L.chain, "chol.block_tile_update", rc = gemm_tn_launch(t, L.chain));
  if (rc) return rc;
  if (rest > tw) {
    GemmArgs r = t;
    r.B = blk + c0 + tw;
    r.N = rest - tw;
    r.D = d0 + tw;
    MG_TIMED(L.chain2, "chol.block_update_first", rc = gemm_tn_launch(r, L.chain2));
    if (rc) return rc;
  }
  L.record(L.row_rest[pj & 1], L.chain2);
  return 0;
}

int cholesky_upper(float* A, int64_t n, int64_t ld, const CholWorkspace& ws, int* info,
                   const Lanes& lanes) {
  CholStepper st{A, n, ld, ws, info, &lanes};
  for (int64_t pj = 0; pj < st.panels(); ++pj) {
    const int rc = st.step(pj);
    if (rc) return rc;
  }
  return 0;
}

// debug: copy the phase timestamps written by block 0 / thread 0 of the last potrf128 (slots 0..20)
// and trsm128 (slots 32..50) launches
extern "C" int mg_debug_clocks(long long* out64) {
  return cudaMemcpyFromSymbol(out64, g_dbg_clk, sizeof(long long) * 64) == cudaSuccess ? 0 : -1;
}

}  // namespace mg
