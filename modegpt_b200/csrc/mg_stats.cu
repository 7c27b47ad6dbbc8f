// mg_stats.cu — C ABI for the calibration statistics.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/modegpt_b200.h"
#include "mg_gemm.cuh"

namespace {

inline int cuda_rc() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - static_cast<int>(e);
}

// ---------------------------------------------------------------------------------------------
// Block-Influence partial sum: one warp per row, 16-byte loads, fp32 products folded into fp64
// running sums every 8 elements; one fp64 atomic per block.
// Algorithmic bytes: rows * d * 2 tensors * 2 B.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bi_cosine_kernel(const __nv_bfloat16* __restrict__ xin,
                                                        int64_t ld_in,
                                                        const __nv_bfloat16* __restrict__ xout,
                                                        int64_t ld_out, int64_t rows, int64_t d,
                                                        double* __restrict__ acc) {
  __shared__ double warp_sums[8];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = static_cast<int64_t>(gridDim.x) * 8;
  double local = 0.0;
  const int64_t nvec = d >> 3;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + warp; r < rows; r += warps_total) {
    const uint4* a = reinterpret_cast<const uint4*>(xin + r * ld_in);
    const uint4* b = reinterpret_cast<const uint4*>(xout + r * ld_out);
    double dot = 0.0, na = 0.0, nb = 0.0;
    for (int64_t v = lane; v < nvec; v += 32) {
      const uint4 ua = __ldg(a + v);
      const uint4 ub = __ldg(b + v);
      const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&ua);
      const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&ub);
      float fd = 0.f, fa = 0.f, fb = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 x = __bfloat1622float2(pa[j]);
        const float2 y = __bfloat1622float2(pb[j]);
        fd = fmaf(x.x, y.x, fd);
        fd = fmaf(x.y, y.y, fd);
        fa = fmaf(x.x, x.x, fa);
        fa = fmaf(x.y, x.y, fa);
        fb = fmaf(y.x, y.x, fb);
        fb = fmaf(y.y, y.y, fb);
      }
      dot += fd;
      na += fa;
      nb += fb;
    }
    // scalar tail (d not a multiple of 8)
    for (int64_t j = (nvec << 3) + lane; j < d; j += 32) {
      const double x = __bfloat162float(xin[r * ld_in + j]);
      const double y = __bfloat162float(xout[r * ld_out + j]);
      dot += x * y;
      na += x * x;
      nb += y * y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dot += __shfl_xor_sync(0xffffffffu, dot, o);
      na += __shfl_xor_sync(0xffffffffu, na, o);
      nb += __shfl_xor_sync(0xffffffffu, nb, o);
    }
    if (lane == 0) {
      // torch.cosine_similarity semantics: each norm clamped at eps = 1e-8
      const double da = fmax(sqrt(na), 1e-8), db = fmax(sqrt(nb), 1e-8);
      local += 1.0 - dot / (da * db);
    }
  }
  if (lane == 0) warp_sums[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += warp_sums[i];
    atomicAdd(acc, s);
  }
}

// ---------------------------------------------------------------------------------------------
// scale the upper triangle and mirror it: 32x32 tiles through padded shared memory so both the
// read of tile (bi,bj) and the transposed write to (bj,bi) are coalesced.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) finalize_sym_kernel(float* __restrict__ C, int64_t n,
                                                           int64_t ldc, float scale) {
  __shared__ float tile[32][33];
  const int bj = blockIdx.x, bi = blockIdx.y;
  if (bi > bj) return;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int64_t r0 = static_cast<int64_t>(bi) * 32, c0 = static_cast<int64_t>(bj) * 32;
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + tx;
    float v = 0.f;
    if (r < n && c < n && c >= r) {
      v = C[r * ldc + c] * scale;
      C[r * ldc + c] = v;
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    // element (row = c0 + i, col = r0 + tx) of the lower triangle <- tile[tx][i]
    const int64_t r = c0 + i, c = r0 + tx;
    if (r < n && c < n && r > c) C[r * ldc + c] = tile[tx][i];
  }
}

// ---------------------------------------------------------------------------------------------
// packed upper triangle (row-major, row i holds columns i..n-1 at offset i*n - i*(i-1)/2): the
// wire format of the cross-rank reduction — half the bytes of the square accumulator.
// One block row per matrix row; within a row both sides are contiguous, so loads and stores are
// full 128-byte lines apart from the ragged ends.
// ---------------------------------------------------------------------------------------------
template <bool PACK>
__global__ void __launch_bounds__(256) pack_upper_kernel(float* __restrict__ C, int64_t n,
                                                         int64_t ldc, float* __restrict__ packed) {
  for (int64_t r = blockIdx.y; r < n; r += gridDim.y) {
    float* row = C + r * ldc;
    float* prow = packed + (r * n - r * (r - 1) / 2) - r;   // prow[c] is element (r, c), c >= r
    for (int64_t c = r + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < n;
         c += static_cast<int64_t>(gridDim.x) * blockDim.x) {
      if (PACK) prow[c] = row[c];
      else row[c] = prow[c];
    }
  }
}

// heads[h][i][j] += alpha * full[h*hd + min(i,j)][h*hd + max(i,j)]: the diagonal hd x hd blocks of an
// upper-triangular Gram, mirrored into full per-head blocks (head dims the block-diagonal tile set
// of the engine does not cover: anything but 32 / 64 / 128)
__global__ void __launch_bounds__(256) add_diag_blocks_kernel(const float* __restrict__ full, int64_t ld,
                                                              int hd, float alpha,
                                                              float* __restrict__ heads) {
  const int h = blockIdx.x;
  const float* blk = full + (static_cast<int64_t>(h) * hd) * ld + static_cast<int64_t>(h) * hd;
  float* out = heads + static_cast<int64_t>(h) * hd * hd;
  for (int e = threadIdx.x; e < hd * hd; e += blockDim.x) {
    const int i = e / hd, j = e - i * hd;
    const int r = i < j ? i : j, c = i < j ? j : i;
    out[e] += alpha * blk[static_cast<int64_t>(r) * ld + c];
  }
}

__global__ void scale_kernel(float* __restrict__ x, int64_t count, float scale) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < count;
       i += stride)
    x[i] *= scale;
}

}  // namespace

extern "C" {

int mg_version(void) { return 1; }

int mg_device_sm_count(void) { return mg::device_sm_count(); }

const char* mg_error_string(int code) {
  if (code == 0) return "ok";
  if (code > 0) return "numerical failure: non-positive pivot (value = 1-based index)";
  if (code <= -1000) return cudaGetErrorString(static_cast<cudaError_t>(-1000 - code));
  switch (code) {
    case -1: return "null pointer argument";
    case -2: return "non-positive dimension";
    case -5: return "matrix must be square for this tile set";
    case -6: return "head dim must be 32, 64 or 128 and divide n";
    case -7: return "leading dimension smaller than row length";
    case -9: return "split-K requires accumulate mode";
    case -10: return "workspace too small (see the *_ws_bytes query)";
    case -11: return "invalid rank / head dimension / mode";
    case -90: return "cuTensorMapEncodeTiled unavailable (driver too old?)";
    case -92: return "bf16 operand must be 16-byte aligned with ld % 8 == 0";
    case -94: return "cuTensorMapEncodeTiled failed";
    default: return "invalid argument";
  }
}

int mg_syrk_bf16_f32(const void* X, int64_t T, int64_t n, int64_t ldx, float* C, int64_t ldc,
                     float alpha, int accumulate, void* stream) {
  mg::GemmArgs a{};
  a.A = a.B = static_cast<const __nv_bfloat16*>(X);
  a.lda = a.ldb = ldx;
  a.a_planes = a.b_planes = 1;
  a.npairs = 1;
  a.M = a.N = n;
  a.K = T;
  a.D = C;
  a.ldd = ldc;
  a.alpha = alpha;
  a.tiles = mg::TILES_UPPER;
  a.epi = accumulate ? mg::EPI_ADD : mg::EPI_STORE;
  a.hd = 128;
  a.ksplit = accumulate ? -mg::segments_for(T) : 1;   // auto split-K, at least the segment count
  return mg::gemm_tn_launch(a, static_cast<cudaStream_t>(stream));
}

int mg_syrk_heads_bf16_f32(const void* X, int64_t T, int64_t n, int64_t ldx, int hd, float* C,
                           float alpha, int accumulate, void* stream) {
  mg::GemmArgs a{};
  a.A = a.B = static_cast<const __nv_bfloat16*>(X);
  a.lda = a.ldb = ldx;
  a.a_planes = a.b_planes = 1;
  a.npairs = 1;
  a.M = a.N = n;
  a.K = T;
  a.D = C;
  a.ldd = hd;
  a.alpha = alpha;
  a.tiles = mg::TILES_DIAG;
  a.epi = accumulate ? mg::EPI_ADD : mg::EPI_STORE;
  a.hd = hd;
  a.ksplit = accumulate ? -mg::segments_for(T) : 1;   // at least that many; more if SMs are idle
  return mg::gemm_tn_launch(a, static_cast<cudaStream_t>(stream));
}

int mg_bi_cosine_bf16(const void* x_in, int64_t ld_in, const void* x_out, int64_t ld_out,
                      int64_t rows, int64_t d, double* acc, void* stream) {
  if (!x_in || !x_out || !acc) return -1;
  if (rows <= 0 || d <= 0) return -2;
  if (ld_in < d || ld_out < d) return -7;
  if ((ld_in % 8) || (ld_out % 8) || (reinterpret_cast<uintptr_t>(x_in) & 15) ||
      (reinterpret_cast<uintptr_t>(x_out) & 15))
    return -92;
  int64_t blocks = (rows + 7) / 8;
  const int64_t cap = static_cast<int64_t>(mg::device_sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  bi_cosine_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x_in), ld_in, static_cast<const __nv_bfloat16*>(x_out),
      ld_out, rows, d, acc);
  return cuda_rc();
}

int mg_finalize_sym_f32(float* C, int64_t n, int64_t ldc, float scale, void* stream) {
  if (!C) return -1;
  if (n <= 0) return -2;
  if (ldc < n) return -7;
  const int nb = static_cast<int>((n + 31) / 32);
  finalize_sym_kernel<<<dim3(nb, nb), 256, 0, static_cast<cudaStream_t>(stream)>>>(C, n, ldc,
                                                                                    scale);
  return cuda_rc();
}

int mg_add_diag_blocks_f32(const float* full, int64_t n, int64_t ld, int hd, float alpha, float* heads,
                           void* stream) {
  if (!full || !heads) return -1;
  if (n <= 0 || hd <= 0 || n % hd) return -2;
  if (ld < n) return -7;
  add_diag_blocks_kernel<<<static_cast<unsigned>(n / hd), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      full, ld, hd, alpha, heads);
  return cuda_rc();
}

int mg_pack_upper_f32(const float* C, int64_t n, int64_t ldc, float* packed, void* stream) {
  if (!C || !packed) return -1;
  if (n <= 0) return -2;
  if (ldc < n) return -7;
  const unsigned gx = static_cast<unsigned>(n >= 4096 ? 4 : 1);
  const unsigned gy = static_cast<unsigned>(n < 65535 ? n : 65535);
  pack_upper_kernel<true><<<dim3(gx, gy), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      const_cast<float*>(C), n, ldc, packed);
  return cuda_rc();
}

int mg_unpack_upper_f32(const float* packed, int64_t n, float* C, int64_t ldc, void* stream) {
  if (!C || !packed) return -1;
  if (n <= 0) return -2;
  if (ldc < n) return -7;
  const unsigned gx = static_cast<unsigned>(n >= 4096 ? 4 : 1);
  const unsigned gy = static_cast<unsigned>(n < 65535 ? n : 65535);
  pack_upper_kernel<false><<<dim3(gx, gy), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      C, n, ldc, const_cast<float*>(packed));
  return cuda_rc();
}

int mg_scale_f32(float* x, int64_t count, float scale, void* stream) {
  if (!x) return -1;
  if (count <= 0) return -2;
  int64_t blocks = (count + 255) / 256;
  const int64_t cap = static_cast<int64_t>(mg::device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  scale_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, count,
                                                                                        scale);
  return cuda_rc();
}

}  // extern "C"
