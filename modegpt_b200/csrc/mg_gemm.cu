// mg_gemm.cu — the tcgen05/TMEM/TMA tensor-core engine.
//
// Replaces the reference's fp64 cuBLAS products on the statistics path
//   src/adapters/LlamaAdapter.py:115-147  (H^T H, per-head bmm, X^T X)
// and supplies the bf16-split fp32-accurate GEMM the blocked Cholesky / TRSM / Nystrom kernels use.
//
// Kernel shape (one CTA per SM, persistent over work items):
//   warp 0      : TMA producer   — 64-column x BK-row boxes, SWIZZLE_128B, 3D maps (col,row,plane)
//   warp 1      : MMA issuer     — tcgen05.mma 128 x BN x 16, both operands MN-major, fp32 in TMEM
//   warps 2..9  : epilogue       — tcgen05.ld 32x32b, fused accumulate / triangular predication
//                 (two warps per TMEM lane quadrant, each draining half of the columns)
// TMEM holds two accumulators (2 x BN columns) so the epilogue of tile i overlaps tile i+1.
#include "mg_gemm.cuh"

#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "mg_once.cuh"
#include "mg_ptx.cuh"

namespace mg {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kChunkCols = 64;                    // one TMA box = 64 bf16 = 128 B wide
constexpr uint32_t kBoxBytes = BK * 128;          // 8 KB
constexpr uint32_t kABytes = (BM / 64) * kBoxBytes;  // 16 KB
constexpr int kEpiWarps = 8;                       // two per TMEM lane quadrant, half the columns each
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;

template <int BN>
struct Cfg {
  static constexpr uint32_t b_bytes = (BN / 64) * kBoxBytes;
  static constexpr uint32_t stage_bytes = kABytes + b_bytes;
  static constexpr int stages = (BN == 256) ? 4 : 6;
  static constexpr uint32_t epi_bytes = kEpiWarps * 4096;  // per epilogue warp: one 32 x 32 fp32 tile
  static constexpr uint32_t smem_bytes = stages * stage_bytes + epi_bytes + 1024 + 512;
};

struct KParams {
  int64_t M, N, K;
  float* D;
  int64_t ldd;
  float alpha;
  int tiles, epi, hd;
  int ksplit, kb_per_split, kblocks;
  int MT, NT, ntiles, total_work;
  int npairs;
  int pair_a[kMaxPairs];
  int pair_b[kMaxPairs];
  int atomic;
  int vec_ok;  // D and ldd allow float4 accesses
  int klo_from_n;
  int tma_epi;  // epilogue through swizzled smem + TMA store / reduce-add
  int band;     // TILES_UPPER with square tiles: rows per raster band (0 = column-major order)
};

struct Work {
  int mi, nj, kb0, kb1;
};

template <int BN, int BMT = BM>
__device__ __forceinline__ Work decode_work(const KParams& p, int w) {
  constexpr int R = BN / BMT;
  const int tile = w / p.ksplit;
  const int ks = w - tile * p.ksplit;
  Work o;
  if (p.tiles == TILES_FULL) {
    o.nj = tile / p.MT;
    o.mi = tile - o.nj * p.MT;
  } else if (p.tiles == TILES_DIAG) {
    o.mi = tile;
    o.nj = tile;
  } else if (R == 1 && p.band > 0) {
    // Banded rasterisation of the upper triangle (square tiles): bands of `band` tile rows, each
    // swept column by column with the row index fastest.  The ~74 tiles in flight at any moment
    // then cover about band x (74/band) tiles and read band + 74/band distinct operand panels
    // instead of ~NT + 2, which is what cuts the DRAM re-reads of X (every cluster walks K in step,
    // so the shared panels hit in L2).
    int t = tile, r0 = 0;
    for (;;) {
      const int g = min(p.band, p.MT - r0);
      const int cnt = g * (p.NT - r0) - g * (g - 1) / 2;   // sum_{i<g} (NT - r0 - i)
      if (t < cnt || r0 + g >= p.MT) break;
      t -= cnt;
      r0 += g;
    }
    const int g = min(p.band, p.MT - r0);
    const int tri = g * (g + 1) / 2;
    if (t < tri) {
      int c = static_cast<int>((sqrtf(1.f + 8.f * static_cast<float>(t)) - 1.f) * 0.5f);
      while (c > 0 && c * (c + 1) / 2 > t) --c;
      while ((c + 1) * (c + 2) / 2 <= t) ++c;
      o.nj = r0 + c;
      o.mi = r0 + t - c * (c + 1) / 2;
    } else {
      const int u = t - tri;
      o.nj = r0 + g + u / g;
      o.mi = r0 + u % g;
    }
  } else {
    // column nj of the tile grid owns rows mi in [0, min(R*(nj+1), MT)); only the last column can
    // be clipped, so prefix(nj) = R*nj*(nj+1)/2 for every valid nj.
    int nj = static_cast<int>((sqrtf(1.f + 8.f * static_cast<float>(tile) / R) - 1.f) * 0.5f);
    if (nj >= p.NT) nj = p.NT - 1;
    while (nj > 0 && R * nj * (nj + 1) / 2 > tile) --nj;
    while (nj + 1 < p.NT && R * (nj + 1) * (nj + 2) / 2 <= tile) ++nj;
    o.nj = nj;
    o.mi = tile - R * nj * (nj + 1) / 2;
  }
  if (p.klo_from_n) {
    // the tile's own k range is [lo, kblocks); split it evenly (a split may come out empty)
    const int lo = min(p.kblocks, (o.nj * BN) / BK);
    const int per = (p.kblocks - lo + p.ksplit - 1) / p.ksplit;
    o.kb0 = lo + ks * per;
    o.kb1 = min(p.kblocks, o.kb0 + per);
  } else {
    o.kb0 = ks * p.kb_per_split;
    o.kb1 = min(p.kblocks, o.kb0 + p.kb_per_split);
  }
  return o;
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmD, const KParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* epi_smem = smem + C::stages * C::stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_smem + C::epi_bytes);
  uint64_t* empty = full + C::stages;
  uint64_t* tmem_full = empty + C::stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.tma_epi) tma_prefetch_desc(&tmD);
    for (int s = 0; s < C::stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiThreads);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        const Work wk = decode_work<BN>(p, w);
        if (wk.kb0 >= wk.kb1) continue;
        const int a_col = wk.mi * BM;
        const int b_col = wk.nj * BN;
        for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
          for (int pr = 0; pr < p.npairs; ++pr) {
            mbar_wait(&empty[stage], phase ^ 1u);
            mbar_expect_tx(&full[stage], C::stage_bytes);
            uint8_t* sa = smem + stage * C::stage_bytes;
            uint8_t* sb = sa + kABytes;
#pragma unroll
            for (int c = 0; c < BM / kChunkCols; ++c)
              tma_load_3d(sa + c * kBoxBytes, &tmA, &full[stage], a_col + c * kChunkCols, kb * BK,
                          p.pair_a[pr]);
#pragma unroll
            for (int c = 0; c < BN / kChunkCols; ++c)
              tma_load_3d(sb + c * kBoxBytes, &tmB, &full[stage], b_col + c * kChunkCols, kb * BK,
                          p.pair_b[pr]);
            if (++stage == C::stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, true, true);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
        const Work wk = decode_work<BN>(p, w);
        if (wk.kb0 >= wk.kb1) continue;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1u;
        ++it;
        mbar_wait(&tmem_empty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * BN);
        uint32_t accum = 0;
        for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
          for (int pr = 0; pr < p.npairs; ++pr) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * C::stage_bytes);
            const uint32_t sb = sa + kABytes;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t adesc = make_smem_desc_sw128(sa + k * 2048, kBoxBytes, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(sb + k * 2048, kBoxBytes, 1024);
              umma_bf16(d_tmem, adesc, bdesc, idesc, accum);
              accum = 1;
            }
            umma_commit(&empty[stage]);
            if (++stage == C::stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int q = warp & 3;  // TMEM lane quadrant this warp may read
    const int chalf = (warp - 2) >> 2;   // which half of the accumulator columns this warp drains
    constexpr int kChunksPerWarp = BN / 64;
    const int cc_lo = chalf * kChunksPerWarp, cc_hi = cc_lo + kChunksPerWarp;
    const int row_in_tile = q * 32 + lane;
    int it = 0;
    for (int w = blockIdx.x; w < p.total_work; w += gridDim.x) {
      const Work wk = decode_work<BN>(p, w);
      if (wk.kb0 >= wk.kb1) continue;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1u;
      ++it;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int64_t gr = static_cast<int64_t>(wk.mi) * BM + row_in_tile;
      const int64_t warp_row0 = static_cast<int64_t>(wk.mi) * BM + q * 32;
      const bool row_ok = gr < p.M;
      if (p.tma_epi) {
        // TMEM -> registers -> 128B-swizzled staging tile -> TMA store / L2 reduce-add.
        // TMA clips at the tensor bounds; in TILES_UPPER chunks that straddle the diagonal are
        // written whole (the strict lower triangle is unspecified by contract).
        uint8_t* buf = epi_smem + (warp - 2) * 4096;
        // batch this warp's chunks into registers, hand the accumulator back, then stage / store
        float vv[kChunksPerWarp][32];
        bool live[kChunksPerWarp];
#pragma unroll
        for (int j = 0; j < kChunksPerWarp; ++j) {
          const int cc = cc_lo + j;
          const int64_t gc0 = static_cast<int64_t>(wk.nj) * BN + cc * 32;
          live[j] = warp_row0 < p.M && gc0 < p.N && !(p.tiles == TILES_UPPER && gc0 + 31 < warp_row0);
          if (live[j]) {
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                   static_cast<uint32_t>(as * BN + cc * 32);
            tmem_ld_32x32(taddr, vv[j]);
          }
        }
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&tmem_empty[as]);
#pragma unroll
        for (int j = 0; j < kChunksPerWarp; ++j) {
          if (!live[j]) continue;
          const int cc = cc_lo + j;
          const int64_t gc0 = static_cast<int64_t>(wk.nj) * BN + cc * 32;
          if (lane == 0) bulk_wait_read<0>();  // the previous store has read `buf`
          __syncwarp();
          uint8_t* rowp = buf + lane * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(rowp + ((c ^ (lane & 7)) << 4)) =
                make_float4(p.alpha * vv[j][4 * c], p.alpha * vv[j][4 * c + 1], p.alpha * vv[j][4 * c + 2],
                            p.alpha * vv[j][4 * c + 3]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            const int c0 = (p.tiles == TILES_DIAG) ? cc * 32 : static_cast<int>(gc0);
            if (p.epi == EPI_STORE) tma_store_2d(&tmD, buf, c0, static_cast<int>(warp_row0));
            else tma_reduce_add_2d(&tmD, buf, c0, static_cast<int>(warp_row0));
            bulk_commit();
          }
        }
        continue;
      }
#pragma unroll 1
      for (int cc = cc_lo; cc < cc_hi; ++cc) {
        const int64_t gc0 = static_cast<int64_t>(wk.nj) * BN + cc * 32;
        // warp-uniform skips: chunk entirely out of range / strictly below the diagonal /
        // outside this warp's head block.
        if (gc0 >= p.N) break;
        if (p.tiles == TILES_UPPER && gc0 + 31 < warp_row0) continue;
        if (p.tiles == TILES_DIAG && p.hd < 128) {
          const int64_t hb0 = (warp_row0 / p.hd) * p.hd;  // 32 | hd so a warp sits in one head
          if (gc0 + 31 < hb0 || gc0 >= hb0 + p.hd) continue;
        }
        float v[32];
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(as * BN + cc * 32);
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        if (!row_ok) continue;
        float* dst;
        bool full_chunk;
        if (p.tiles == TILES_DIAG) {
          const int64_t head = gr / p.hd;
          const int64_t hb0 = head * p.hd;
          dst = p.D + head * p.hd * p.hd + (gr - hb0) * p.hd + (gc0 - hb0);
          full_chunk = (gc0 >= hb0) && (gc0 + 32 <= hb0 + p.hd) && (gc0 + 32 <= p.N);
          if (!full_chunk) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int64_t gc = gc0 + j;
              if (gc >= hb0 && gc < hb0 + p.hd && gc < p.N) {
                const float x = p.alpha * v[j];
                if (p.epi == EPI_STORE) dst[j] = x;
                else if (p.atomic) atomicAdd(dst + j, x);
                else dst[j] += x;
              }
            }
            continue;
          }
        } else {
          dst = p.D + gr * p.ldd + gc0;
          full_chunk = (gc0 + 32 <= p.N) && (p.tiles != TILES_UPPER || gc0 >= gr);
          if (!full_chunk) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int64_t gc = gc0 + j;
              if (gc < p.N && (p.tiles != TILES_UPPER || gc >= gr)) {
                const float x = p.alpha * v[j];
                if (p.epi == EPI_STORE) dst[j] = x;
                else if (p.atomic) atomicAdd(dst + j, x);
                else dst[j] += x;
              }
            }
            continue;
          }
        }
        // full 32-wide chunk
        if (p.atomic) {
#pragma unroll
          for (int j = 0; j < 32; ++j) atomicAdd(dst + j, p.alpha * v[j]);
        } else if (p.vec_ok) {
          float4* d4 = reinterpret_cast<float4*>(dst);
          if (p.epi == EPI_STORE) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              d4[j] = make_float4(p.alpha * v[4 * j], p.alpha * v[4 * j + 1],
                                  p.alpha * v[4 * j + 2], p.alpha * v[4 * j + 3]);
          } else {
            float4 o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = d4[j];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              o[j].x += p.alpha * v[4 * j];
              o[j].y += p.alpha * v[4 * j + 1];
              o[j].z += p.alpha * v[4 * j + 2];
              o[j].w += p.alpha * v[4 * j + 3];
              d4[j] = o[j];
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = p.alpha * v[j];
            if (p.epi == EPI_STORE) dst[j] = x;
            else dst[j] += x;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
    }
  }

  if (warp >= 2 && lane == 0 && p.tma_epi) bulk_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair engine: one 256 x 256 output tile per cluster of two CTAs (tcgen05 cta_group::2).
//   * each CTA stages only ITS half of both operands (128 M-columns of A, 128 N-columns of B) —
//     half the shared-memory fill and read traffic of the single-CTA tile per MMA flop;
//   * both CTAs' TMA loads credit the LEADER's `full` barrier; the leader's elected thread issues
//     tcgen05.mma.cta_group::2 (M = 256: rows 0-127 accumulate in the leader's TMEM, 128-255 in the
//     peer's), and its commits are multicast to the `empty` / `tmem_full` barriers of both CTAs;
//   * each CTA drains its own 128 TMEM lanes through the same TMA store / reduce-add epilogue and
//     releases the accumulator by a remote arrive on the leader's `tmem_empty` barrier.
// ------------------------------------------------------------------------------------------------
//   * eight epilogue warps per CTA (two per TMEM lane quadrant, 128 accumulator columns each): with
//     four, a K = 128 x 6-plane trailing update spent 9 us per tile draining 8 chunks per warp
//     against 5.3 us of MMA (tensor pipe 44 % in ncu); halving the chunks per warp puts the drain
//     back under the main loop.
constexpr int kPairBM = 256, kPairBN = 256;
constexpr int kPairStages = 6;
constexpr uint32_t kPairStageBytes = 2 * kABytes;   // 16 KB of A + 16 KB of B per CTA
constexpr int kPairEpiWarps = 8;
constexpr int kPairEpiThreads = kPairEpiWarps * 32;
constexpr int kPairThreads = 64 + kPairEpiThreads;
static_assert(kPairThreads == kThreads, "both kernels use two service warps + eight epilogue warps");
constexpr uint32_t kPairEpiBytes = kPairEpiWarps * 4096;   // one 32 x 32 fp32 staging tile per warp
constexpr uint32_t kPairSmem = kPairStages * kPairStageBytes + kPairEpiBytes + 1024 + 512;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
    gemm_tn_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                        const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmD, const KParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* epi_smem = smem + kPairStages * kPairStageBytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(epi_smem + kPairEpiBytes);
  uint64_t* empty = full + kPairStages;
  uint64_t* tmem_full = empty + kPairStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * kPairEpiThreads);   // epilogue threads of BOTH CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_slot, 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < p.total_work; w += n_clusters) {
        const Work wk = decode_work<kPairBN, kPairBM>(p, w);
        if (wk.kb0 >= wk.kb1) continue;
        const int a_col = wk.mi * kPairBM + static_cast<int>(rank) * 128;
        const int b_col = wk.nj * kPairBN + static_cast<int>(rank) * 128;
        for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
          for (int pr = 0; pr < p.npairs; ++pr) {
            mbar_wait(&empty[stage], phase ^ 1u);
            if (leader) mbar_expect_tx(&full[stage], 2 * kPairStageBytes);
            const uint32_t full_leader = map_to_cta(smem_u32(&full[stage]), 0);
            uint8_t* sa = smem + stage * kPairStageBytes;
            uint8_t* sb = sa + kABytes;
#pragma unroll
            for (int c = 0; c < 2; ++c)
              tma_load_3d_pair(sa + c * kBoxBytes, &tmA, full_leader, a_col + c * kChunkCols, kb * BK,
                               p.pair_a[pr]);
#pragma unroll
            for (int c = 0; c < 2; ++c)
              tma_load_3d_pair(sb + c * kBoxBytes, &tmB, full_leader, b_col + c * kChunkCols, kb * BK,
                               p.pair_b[pr]);
            if (++stage == kPairStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader only)
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_bf16(kPairBM, kPairBN, true, true);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int w = cluster_id; w < p.total_work; w += n_clusters) {
        const Work wk = decode_work<kPairBN, kPairBM>(p, w);
        if (wk.kb0 >= wk.kb1) continue;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1u;
        ++it;
        mbar_wait(&tmem_empty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(as * kPairBN);
        uint32_t accum = 0;
        for (int kb = wk.kb0; kb < wk.kb1; ++kb) {
          for (int pr = 0; pr < p.npairs; ++pr) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * kPairStageBytes);
            const uint32_t sb = sa + kABytes;
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t adesc = make_smem_desc_sw128(sa + k * 2048, kBoxBytes, 1024);
              const uint64_t bdesc = make_smem_desc_sw128(sb + k * 2048, kBoxBytes, 1024);
              umma_bf16_pair(d_tmem, adesc, bdesc, idesc, accum);
              accum = 1;
            }
            umma_commit_pair(&empty[stage]);
            if (++stage == kPairStages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        umma_commit_pair(&tmem_full[as]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (both CTAs)
    const int q = warp & 3;                       // TMEM lane quadrant this warp may read
    const int chalf = (warp - 2) >> 2;            // which 128 accumulator columns it drains
    uint8_t* buf = epi_smem + (warp - 2) * 4096;
    int it = 0;
    for (int w = cluster_id; w < p.total_work; w += n_clusters) {
      const Work wk = decode_work<kPairBN, kPairBM>(p, w);
      if (wk.kb0 >= wk.kb1) continue;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1u;
      ++it;
      mbar_wait(&tmem_full[as], aphase);
      tc_fence_after();
      const int64_t warp_row0 = static_cast<int64_t>(wk.mi) * kPairBM + rank * 128 + q * 32;
      // this warp's four 32-column chunks go to registers in one batch, then the accumulator is
      // handed back at once: the next-but-one tile's MMAs no longer wait for the staging / TMA
      // part of the epilogue, only for the TMEM reads
      float v[4][32];
      bool live[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cc = chalf * 4 + j;
        const int64_t gc0 = static_cast<int64_t>(wk.nj) * kPairBN + cc * 32;
        live[j] = warp_row0 < p.M && gc0 < p.N && !(p.tiles == TILES_UPPER && gc0 + 31 < warp_row0);
        if (live[j]) {
          const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                                 static_cast<uint32_t>(as * kPairBN + cc * 32);
          tmem_ld_32x32(taddr, v[j]);
        }
      }
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive_cluster(map_to_cta(smem_u32(&tmem_empty[as]), 0));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (!live[j]) continue;
        const int64_t gc0 = static_cast<int64_t>(wk.nj) * kPairBN + (chalf * 4 + j) * 32;
        if (lane == 0) bulk_wait_read<0>();     // the previous store has read the staging tile
        __syncwarp();
        uint8_t* rowp = buf + lane * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(rowp + ((c ^ (lane & 7)) << 4)) =
              make_float4(p.alpha * v[j][4 * c], p.alpha * v[j][4 * c + 1], p.alpha * v[j][4 * c + 2],
                          p.alpha * v[j][4 * c + 3]);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (p.epi == EPI_STORE) tma_store_2d(&tmD, buf, static_cast<int>(gc0), static_cast<int>(warp_row0));
          else tma_reduce_add_2d(&tmD, buf, static_cast<int>(gc0), static_cast<int>(warp_row0));
          bulk_commit();
        }
      }
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();   // the peer's smem / TMEM stay live until every MMA and remote arrive is done
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 3D map over bf16 planes: dim0 = columns (contiguous), dim1 = rows (K), dim2 = plane.
int make_plane_map(CUtensorMap* tm, const __nv_bfloat16* base, int64_t cols, int64_t rows,
                   int64_t ld, int planes, int64_t plane_stride) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return -90;
  if (planes < 1) return -91;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 8) || ld < cols) return -92;
  if (planes > 1 && (plane_stride % 8)) return -93;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(planes)};
  const int64_t ps = planes > 1 ? plane_stride : rows * ld;
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(ps) * 2};
  cuuint32_t box[3] = {kChunkCols, BK, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                   const_cast<void*>(static_cast<const void*>(base)), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -94;
}

// 2D fp32 map of the output for the TMA epilogue: 32 x 32 boxes, 128B swizzle.
int make_out_map(CUtensorMap* tm, float* base, int64_t cols, int64_t rows, int64_t ld) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return -90;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -94;
}

template <int BN>
int launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmD,
                const KParams& kp, int cta_cap, cudaStream_t stream) {
  using C = Cfg<BN>;
  static PerDeviceOnce once;
  if (int rc = once.run([] {
        return cudaFuncSetAttribute(gemm_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    C::smem_bytes);
      }))
    return rc;
  const int grid = kp.total_work < cta_cap ? kp.total_work : cta_cap;
  gemm_tn_kernel<BN><<<grid, kThreads, C::smem_bytes, stream>>>(tmA, tmB, tmD, kp);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - static_cast<int>(e);
}

}  // namespace

int device_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

int gemm_tn_launch(const GemmArgs& a, cudaStream_t stream) {
  if (!a.A || !a.B || !a.D) return -1;
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return -2;
  if (a.npairs < 1 || a.npairs > kMaxPairs) return -3;
  for (int i = 0; i < a.npairs; ++i)
    if (a.pair_a[i] < 0 || a.pair_a[i] >= a.a_planes || a.pair_b[i] < 0 ||
        a.pair_b[i] >= a.b_planes)
      return -4;
  if (a.tiles == TILES_UPPER && a.M != a.N) return -5;
  if (a.tiles == TILES_DIAG) {
    if (a.M != a.N) return -5;
    if (a.hd != 32 && a.hd != 64 && a.hd != 128) return -6;
    if (a.M % a.hd) return -6;
  } else if (a.ldd < a.N) {
    return -7;
  }
  if (a.epi != EPI_STORE && a.epi != EPI_ADD) return -8;

  const bool bn256 = (a.tiles != TILES_DIAG) && a.N > 128;
  const int BN = bn256 ? 256 : 128;
  // CTA-pair path: plain 2-D fp32 output reachable by the TMA epilogue and at least one full
  // 256-row tile; MG_DISABLE_2CTA=1 forces the single-CTA kernel (A/B measurements).
  static const bool pair_disabled = [] {
    const char* e = std::getenv("MG_DISABLE_2CTA");
    return e && e[0] == '1';
  }();
  const bool pair = !pair_disabled && a.tiles != TILES_DIAG && a.M >= 256 && a.N >= 256 &&
                    ((reinterpret_cast<uintptr_t>(a.D) & 15) == 0) && (a.ldd % 4 == 0);
  const int BMT = pair ? kPairBM : BM;

  KParams kp{};
  kp.M = a.M;
  kp.N = a.N;
  kp.K = a.K;
  kp.D = a.D;
  kp.ldd = a.ldd;
  kp.alpha = a.alpha;
  kp.tiles = a.tiles;
  kp.epi = a.epi;
  kp.hd = a.hd > 0 ? a.hd : 128;
  kp.npairs = a.npairs;
  for (int i = 0; i < a.npairs; ++i) {
    kp.pair_a[i] = a.pair_a[i];
    kp.pair_b[i] = a.pair_b[i];
  }
  kp.MT = static_cast<int>((a.M + BMT - 1) / BMT);
  kp.NT = static_cast<int>((a.N + BN - 1) / BN);
  if (a.tiles == TILES_FULL) {
    kp.ntiles = kp.MT * kp.NT;
  } else if (a.tiles == TILES_DIAG) {
    kp.ntiles = kp.MT;
  } else {
    const int R = BN / BMT;
    const int last = (R * kp.NT < kp.MT) ? R * kp.NT : kp.MT;
    kp.ntiles = R * (kp.NT - 1) * kp.NT / 2 + last;
  }
  kp.kblocks = static_cast<int>((a.K + BK - 1) / BK);
  int ksplit = a.ksplit;
  if (ksplit <= 0) {
    const int at_least = ksplit < 0 ? -ksplit : 1;
    ksplit = 1;
    if (a.epi == EPI_ADD && kp.ntiles < device_sm_count()) {
      ksplit = (2 * device_sm_count() + kp.ntiles - 1) / kp.ntiles;
      // keep at least 8 k-blocks per split so the pipeline fill is amortised
      const int max_split = (kp.kblocks + 7) / 8;
      if (ksplit > max_split) ksplit = max_split;
      if (ksplit < 1) ksplit = 1;
    }
    if (ksplit < at_least) ksplit = at_least;
  }
  if (ksplit > 1 && a.epi != EPI_ADD) return -9;
  if (ksplit > kp.kblocks) ksplit = kp.kblocks;
  kp.kb_per_split = (kp.kblocks + ksplit - 1) / ksplit;
  kp.ksplit = (kp.kblocks + kp.kb_per_split - 1) / kp.kb_per_split;
  kp.atomic = kp.ksplit > 1;
  kp.klo_from_n = a.klo_from_n;
  kp.total_work = kp.ntiles * kp.ksplit;
  if (pair && a.tiles == TILES_UPPER && kp.MT == kp.NT) {
    static const int band_env = [] {
      const char* e = std::getenv("MG_SYRK_BAND");
      return e ? std::atoi(e) : 8;
    }();
    kp.band = band_env;
  }
  kp.vec_ok = ((reinterpret_cast<uintptr_t>(a.D) & 15) == 0) &&
              (a.tiles == TILES_DIAG ? (kp.hd % 4 == 0) : (a.ldd % 4 == 0));

  int cta_cap = device_sm_count();
  if (a.max_ctas >= 2 && a.max_ctas < cta_cap) cta_cap = a.max_ctas;
  CUtensorMap tmA, tmB;
  int rc = make_plane_map(&tmA, a.A, a.M, a.K, a.lda, a.a_planes, a.a_plane_stride);
  if (rc) return rc;
  rc = make_plane_map(&tmB, a.B, a.N, a.K, a.ldb, a.b_planes, a.b_plane_stride);
  if (rc) return rc;
  // TMA epilogue whenever the output is a plain 2D fp32 tile grid with 16-byte aligned rows
  CUtensorMap tmD = tmA;
  const bool diag128 = a.tiles == TILES_DIAG && kp.hd == 128;
  if (((reinterpret_cast<uintptr_t>(a.D) & 15) == 0) &&
      (diag128 || (a.tiles != TILES_DIAG && a.ldd % 4 == 0))) {
    rc = diag128 ? make_out_map(&tmD, a.D, 128, a.M, 128) : make_out_map(&tmD, a.D, a.N, a.M, a.ldd);
    if (rc) return rc;
    kp.tma_epi = 1;
  }
  if (pair) {
    static PerDeviceOnce once;
    if (int orc = once.run([] {
          return cudaFuncSetAttribute(gemm_tn_pair_kernel,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem);
        }))
      return orc;
    const int max_clusters = cta_cap / 2;
    const int clusters = kp.total_work < max_clusters ? kp.total_work : max_clusters;
    gemm_tn_pair_kernel<<<2 * clusters, kPairThreads, kPairSmem, stream>>>(tmA, tmB, tmD, kp);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : -1000 - static_cast<int>(e);
  }
  return bn256 ? launch_impl<256>(tmA, tmB, tmD, kp, cta_cap, stream)
               : launch_impl<128>(tmA, tmB, tmD, kp, cta_cap, stream);
}

}  // namespace mg
