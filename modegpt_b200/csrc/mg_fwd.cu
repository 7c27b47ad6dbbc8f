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

// ---------------------------------------------------------------------------------------------
// Rebuilt-model ops (reference src/patchers/LlamaRebuild.py:119-187,
// DenseQwenRebuild.py:262-286).  A compressed attention layer keeps r of the hd rotary dimensions
// per kv head (mask [KV, r], the order of the type-II selection: first r/2 entries from the first
// half of the head, last r/2 their partners).  The reference gathers cos/sin through the mask on
// every forward — cos[:, :, mask] materialises [B, T, KV, r] twice — then runs the 8 eager RoPE
// kernels; here the gather happens inside the rotation: x is read once, written once, cos/sin
// ([B or 1, T, hd], a few hundred KB) stay in L2.  Arithmetic = the eager sequence incl. its bf16
// roundings: out = bf16(bf16(x * cos_g) + bf16(rotate_half(x) * sin_g)).
// One thread per (token, head, j < r/2) handles the pair (j, j + r/2); any even r.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rope_masked_kernel(const bf16* __restrict__ x,
                                                          bf16* __restrict__ out,
                                                          const bf16* __restrict__ cs,
                                                          const bf16* __restrict__ sn,
                                                          const int64_t* __restrict__ mask,
                                                          int64_t batch, int64_t seq, int heads,
                                                          int group, int r, int hd,
                                                          int64_t cs_batch_stride) {
  const int half = r >> 1;
  const int64_t total = batch * seq * heads * half;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int j = static_cast<int>(e % half);
    const int64_t bth = e / half;                 // (b * seq + t) * heads + h
    const int h = static_cast<int>(bth % heads);
    const int64_t bt = bth / heads;
    const int64_t t = bt % seq, b = bt / seq;
    const int64_t* m = mask + static_cast<int64_t>(h / group) * r;
    const int64_t m1 = m[j], m2 = m[j + half];
    const bf16* cp = cs + b * cs_batch_stride + t * hd;
    const bf16* sp = sn + b * cs_batch_stride + t * hd;
    const bf16* xp = x + bth * r;
    const float x1 = bf(xp[j]), x2 = bf(xp[j + half]);
    const bf16 a1 = rn(x1 * bf(cp[m1]));
    const bf16 r1 = rn(-x2 * bf(sp[m1]));
    const bf16 a2 = rn(x2 * bf(cp[m2]));
    const bf16 r2 = rn(x1 * bf(sp[m2]));
    bf16* op = out + bth * r;
    op[j] = rn(bf(a1) + bf(r1));
    op[j + half] = rn(bf(a2) + bf(r2));
  }
}

// Qwen3 q_norm / k_norm on a compressed head: RMS over the r kept dimensions, weight gathered
// through the mask:  out = bf16(float(w[mask[h / group, j]]) * (x32 * rsqrt(mean(x32^2) + eps)))
// (DenseQwenRebuild.py:262-286 semantics as re-authored in patchers/DenseQwenRebuild.py).
// One warp per (token, head).
__global__ void __launch_bounds__(256) rmsnorm_masked_kernel(const bf16* __restrict__ x,
                                                             int64_t rows_heads, int heads, int group,
                                                             int r, const bf16* __restrict__ w,
                                                             const int64_t* __restrict__ mask,
                                                             float eps, bf16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t wg = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (wg >= rows_heads) return;
  const int h = static_cast<int>(wg % heads);
  const bf16* xp = x + wg * r;
  const int64_t* m = mask + static_cast<int64_t>(h / group) * r;
  float v[4];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = lane + 32 * i;
    v[i] = j < r ? bf(xp[j]) : 0.f;
    ss = fmaf(v[i], v[i], ss);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rs = rsqrtf(ss / static_cast<float>(r) + eps);
  bf16* op = out + wg * r;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = lane + 32 * i;
    if (j < r) op[j] = rn(bf(w[m[j]]) * (v[i] * rs));
  }
}

// ---------------------------------------------------------------------------------------------
// Perplexity (reference src/eval.py:192-220): per-row negative log-likelihood
// straight from bf16 logits, nll[row] = logsumexp(logits[row, :]) - logits[row, label[row]] in
// fp32, one pass (online max / sum), one CTA per row.  The evaluator feeds it row chunks of
// hidden @ W_lm^T, so neither [B, T, vocab] logits nor an fp32 copy of them ever exist.
// HBM-bound: rows * vocab * 2 B read.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ce_rows_kernel(const bf16* __restrict__ logits, int64_t ld,
                                                      int64_t vocab, const int64_t* __restrict__ labels,
                                                      float* __restrict__ nll, int vec_ok) {
  __shared__ float sm[8], ss[8];
  const int64_t row = blockIdx.x;
  const bf16* p = logits + row * ld;
  float m = -INFINITY, s = 0.f;
  auto push = [&](float xv) {
    if (xv > m) {
      s = s * __expf(m - xv) + 1.f;
      m = xv;
    } else {
      s += __expf(xv - m);
    }
  };
  if (vec_ok) {
    const uint4* p4 = reinterpret_cast<const uint4*>(p);
    for (int64_t i = threadIdx.x; i < (vocab >> 3); i += 256) {
      Vec8 v;
      v.u = __ldg(p4 + i);
      float loc = bf(v.h[0]);
#pragma unroll
      for (int j = 1; j < 8; ++j) loc = fmaxf(loc, bf(v.h[j]));
      if (loc > m) {
        s *= __expf(m - loc);
        m = loc;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) s += __expf(bf(v.h[j]) - m);
    }
  } else {
    for (int64_t i = threadIdx.x; i < vocab; i += 256) push(bf(p[i]));
  }
  // merge (m, s) pairs: warp, then block
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mm = fmaxf(m, m2);
    s = (m == -INFINITY ? 0.f : s * __expf(m - mm)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mm));
    m = mm;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
    sm[warp] = m;
    ss[warp] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mm = sm[0], tot = ss[0];
    for (int i = 1; i < 8; ++i) {
      const float m2 = sm[i], s2 = ss[i];
      const float mx = fmaxf(mm, m2);
      tot = (mm == -INFINITY ? 0.f : tot * __expf(mm - mx)) + (m2 == -INFINITY ? 0.f : s2 * __expf(m2 - mx));
      mm = mx;
    }
    const int64_t lab = labels[row];
    nll[row] = (lab >= 0 && lab < vocab) ? (logf(tot) + mm - bf(p[lab])) : 0.f;
  }
}

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

int mg_rope_masked_bf16(const void* x, void* out, const void* cos, const void* sin,
                        const int64_t* mask, int64_t batch, int64_t seq, int n_heads, int group, int r,
                        int head_dim, int64_t cos_batch_stride, void* stream) {
  if (!x || !out || !cos || !sin || !mask) return -1;
  if (batch <= 0 || seq <= 0 || n_heads <= 0 || group <= 0 || head_dim <= 0) return -2;
  if (r <= 0 || r % 2 || r > head_dim || n_heads % group) return -11;
  const int64_t total = batch * seq * n_heads * (r / 2);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = static_cast<int64_t>(mg::device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  rope_masked_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(out), static_cast<const bf16*>(cos),
      static_cast<const bf16*>(sin), mask, batch, seq, n_heads, group, r, head_dim, cos_batch_stride);
  return cuda_rc();
}

int mg_rmsnorm_masked_bf16(const void* x, int64_t rows, int n_heads, int group, int r,
                           const void* weight, const int64_t* mask, float eps, void* out,
                           void* stream) {
  if (!x || !weight || !mask || !out) return -1;
  if (rows <= 0 || n_heads <= 0 || group <= 0) return -2;
  if (r <= 0 || r > 128 || n_heads % group) return -11;
  const int64_t warps = rows * n_heads;
  const unsigned grid = static_cast<unsigned>((warps + 7) / 8);
  rmsnorm_masked_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), warps, n_heads, group, r, static_cast<const bf16*>(weight), mask,
      eps, static_cast<bf16*>(out));
  return cuda_rc();
}

int mg_ce_rows_bf16(const void* logits, int64_t ld, int64_t rows, int64_t vocab,
                    const int64_t* labels, float* nll, void* stream) {
  if (!logits || !labels || !nll) return -1;
  if (rows <= 0 || vocab <= 0) return -2;
  if (ld < vocab) return -7;
  const int vec = (vocab % 8 == 0) && (ld % 8 == 0) && aligned16(logits);
  ce_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(logits), ld, vocab, labels, nll, vec);
  return cuda_rc();
}

}  // extern "C"
