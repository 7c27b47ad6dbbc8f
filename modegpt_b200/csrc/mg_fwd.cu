// mg_fwd.cu — fused elementwise kernels for the CALIBRATION FORWARD.
//
// The hooked forward of src/calibration.py:116 runs HF's eager modules; 36 % of a Llama-2-7B
// calibration step is their elementwise kernels (RMSNorm = 7 kernels, RoPE = 8, SiLU*mul = 2 per
// use).  These kernels replace them one-for-one and reproduce HF's arithmetic INCLUDING every
// intermediate bf16 rounding, so the activations the statistics hooks see do not change:
//   rmsnorm : y = w * bf16(x32 * rsqrt(mean(x32^2) + eps))   (LlamaRMSNorm.forward; the only
//             difference is the summation order inside the mean)
//   swiglu  : out = bf16(silu_f32(g)) * u                     (LlamaMLP.forward: act_fn(gate) * up)
//   rope    : out = bf16(x*cos) + bf16(rotate_half(x)*sin)    (apply_rotary_pos_emb)
// All HBM-bound: one read of each operand, one write of the result, 16-byte accesses.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/modegpt_b200.h"
#include "mg_gemm.cuh"

namespace {

using bf16 = __nv_bfloat16;

inline int cuda_rc() {
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -1000 - static_cast<int>(e);
}

__device__ __forceinline__ float bf(const bf16 x) { return __bfloat162float(x); }
__device__ __forceinline__ bf16 rn(const float x) { return __float2bfloat16_rn(x); }

union Vec8 {
  uint4 u;
  bf16 h[8];
};

// ---------------------------------------------------------------------------------------------
// RMSNorm: G lanes per row (G = 32 for d >= 256, else d / 8), row cached in registers.
// Algorithmic bytes: rows * d * 2 B read + the same written.
// ---------------------------------------------------------------------------------------------
template <int kMaxVec>
__global__ void __launch_bounds__(256) rmsnorm_kernel(const bf16* __restrict__ x, int64_t ldx,
                                                      int64_t rows, int d, int lanes_per_row,
                                                      const bf16* __restrict__ w, float eps,
                                                      bf16* __restrict__ y, int64_t ldy) {
  const int lane = threadIdx.x & 31;
  const int rows_per_warp = 32 / lanes_per_row;
  const int sub = lane / lanes_per_row, gl = lane % lanes_per_row;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t row = warp_global * rows_per_warp + sub;
  const int nvec = d >> 3;                       // 16-byte vectors per row
  const int per_lane = nvec / lanes_per_row;     // exact by construction
  const bool active = row < rows;
  Vec8 cache[kMaxVec];
  float ss = 0.f;
  if (active) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + row * ldx);
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i) {
      if (i < per_lane) {
        cache[i].u = __ldg(xr + gl + i * lanes_per_row);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = bf(cache[i].h[j]);
          ss = fmaf(v, v, ss);
        }
      }
    }
  }
  for (int o = lanes_per_row >> 1; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (!active) return;
  const float rs = rsqrtf(ss / static_cast<float>(d) + eps);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  uint4* yr = reinterpret_cast<uint4*>(y + row * ldy);
#pragma unroll
  for (int i = 0; i < kMaxVec; ++i) {
    if (i < per_lane) {
      Vec8 wv, out;
      wv.u = __ldg(wr + gl + i * lanes_per_row);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bf16 t = rn(bf(cache[i].h[j]) * rs);   // hidden_states.to(input_dtype)
        out.h[j] = rn(bf(wv.h[j]) * bf(t));          // self.weight * (...)
      }
      yr[gl + i * lanes_per_row] = out.u;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// SwiGLU: out = bf16(silu(g)) * u.  Bytes: 3 * count * 2 B.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) swiglu_kernel(const uint4* __restrict__ g,
                                                     const uint4* __restrict__ u,
                                                     uint4* __restrict__ out, int64_t nvec) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    Vec8 a, b, o;
    a.u = __ldg(g + i);
    b.u = __ldg(u + i);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = bf(a.h[j]);
      const bf16 s = rn(x / (1.0f + expf(-x)));     // torch silu: opmath fp32, result rounded
      o.h[j] = rn(bf(s) * bf(b.h[j]));
    }
    out[i] = o.u;
  }
}

// ---------------------------------------------------------------------------------------------
// RoPE on a [B, T, H, hd] buffer (HF hands attention a transposed VIEW of it); cos / sin are
// [Bc, T, hd] with Bc in {1, B}.  One thread rotates 8 (j, j + hd/2) pairs.
// Bytes: 2 * B*T*H*hd * 2 B (+ cos/sin, L2-resident).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rope_kernel(const bf16* __restrict__ x, bf16* __restrict__ out,
                                                   const bf16* __restrict__ cs,
                                                   const bf16* __restrict__ sn, int64_t B, int64_t T,
                                                   int H, int hd, int64_t cs_batch_stride) {
  const int half = hd >> 1;
  const int vec_per_head = half >> 3;
  const int64_t total = B * T * H * vec_per_head;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int v = static_cast<int>(i % vec_per_head);
    const int64_t bth = i / vec_per_head;       // (b*T + t)*H + h
    const int64_t bt = bth / H;
    const int64_t b = bt / T, t = bt - b * T;
    const bf16* xp = x + bth * hd + v * 8;
    const bf16* cp = cs + b * cs_batch_stride + t * hd + v * 8;
    const bf16* sp = sn + b * cs_batch_stride + t * hd + v * 8;
    Vec8 x1, x2, c1, c2, s1, s2, o1, o2;
    x1.u = __ldg(reinterpret_cast<const uint4*>(xp));
    x2.u = __ldg(reinterpret_cast<const uint4*>(xp + half));
    c1.u = __ldg(reinterpret_cast<const uint4*>(cp));
    c2.u = __ldg(reinterpret_cast<const uint4*>(cp + half));
    s1.u = __ldg(reinterpret_cast<const uint4*>(sp));
    s2.u = __ldg(reinterpret_cast<const uint4*>(sp + half));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // first half:  x1*cos + (-x2)*sin ;  second half:  x2*cos + x1*sin   (rotate_half = cat(-x2, x1))
      const bf16 a1 = rn(bf(x1.h[j]) * bf(c1.h[j]));
      const bf16 r1 = rn(-bf(x2.h[j]) * bf(s1.h[j]));
      o1.h[j] = rn(bf(a1) + bf(r1));
      const bf16 a2 = rn(bf(x2.h[j]) * bf(c2.h[j]));
      const bf16 r2 = rn(bf(x1.h[j]) * bf(s2.h[j]));
      o2.h[j] = rn(bf(a2) + bf(r2));
    }
    bf16* op = out + bth * hd + v * 8;
    *reinterpret_cast<uint4*>(op) = o1.u;
    *reinterpret_cast<uint4*>(op + half) = o2.u;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" {

int mg_rmsnorm_bf16(const void* x, int64_t ldx, int64_t rows, int64_t d, const void* weight, float eps,
                    void* y, int64_t ldy, void* stream) {
  if (!x || !weight || !y) return -1;
  if (rows <= 0 || d <= 0) return -2;
  if (ldx < d || ldy < d) return -7;
  if (d % 8 || ldx % 8 || ldy % 8 || !aligned16(x) || !aligned16(y) || !aligned16(weight)) return -92;
  const int nvec = static_cast<int>(d / 8);
  int lanes = 32;
  while (lanes > 1 && nvec % lanes) lanes >>= 1;      // largest power of two dividing nvec, <= 32
  const int per_lane = nvec / lanes;
  if (per_lane > 32) return -11;                       // d > 8192 with awkward factors: unsupported
  const int rows_per_warp = 32 / lanes;
  const int64_t warps = (rows + rows_per_warp - 1) / rows_per_warp;
  const unsigned grid = static_cast<unsigned>((warps + 7) / 8);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bf16* xp = static_cast<const bf16*>(x);
  const bf16* wp = static_cast<const bf16*>(weight);
  bf16* yp = static_cast<bf16*>(y);
  const int di = static_cast<int>(d);
  if (per_lane <= 1) rmsnorm_kernel<1><<<grid, 256, 0, s>>>(xp, ldx, rows, di, lanes, wp, eps, yp, ldy);
  else if (per_lane <= 4) rmsnorm_kernel<4><<<grid, 256, 0, s>>>(xp, ldx, rows, di, lanes, wp, eps, yp, ldy);
  else if (per_lane <= 16) rmsnorm_kernel<16><<<grid, 256, 0, s>>>(xp, ldx, rows, di, lanes, wp, eps, yp, ldy);
  else rmsnorm_kernel<32><<<grid, 256, 0, s>>>(xp, ldx, rows, di, lanes, wp, eps, yp, ldy);
  return cuda_rc();
}

int mg_swiglu_bf16(const void* gate, const void* up, void* out, int64_t count, void* stream) {
  if (!gate || !up || !out) return -1;
  if (count <= 0) return -2;
  if (count % 8 || !aligned16(gate) || !aligned16(up) || !aligned16(out)) return -92;
  const int64_t nvec = count / 8;
  int64_t blocks = (nvec + 255) / 256;
  const int64_t cap = static_cast<int64_t>(mg::device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  swiglu_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(gate), static_cast<const uint4*>(up), static_cast<uint4*>(out), nvec);
  return cuda_rc();
}

int mg_rope_bf16(const void* x, void* out, const void* cos, const void* sin, int64_t batch,
                 int64_t seq, int n_heads, int head_dim, int64_t cos_batch_stride, void* stream) {
  if (!x || !out || !cos || !sin) return -1;
  if (batch <= 0 || seq <= 0 || n_heads <= 0 || head_dim <= 0) return -2;
  if (head_dim % 16 || !aligned16(x) || !aligned16(out) || !aligned16(cos) || !aligned16(sin) ||
      cos_batch_stride % 8)
    return -92;
  const int64_t total = batch * seq * n_heads * (head_dim / 16);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(mg::device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  rope_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(out), static_cast<const bf16*>(cos),
      static_cast<const bf16*>(sin), batch, seq, n_heads, head_dim, cos_batch_stride);
  return cuda_rc();
}

}  // extern "C"
