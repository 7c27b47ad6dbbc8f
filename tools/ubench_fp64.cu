// Micro-benchmark behind the potrf128 pivot chain: accuracy of the 64-bit MUFU
// reciprocal-square-root seed and of the corrections built on it, and the dependent-issue
// latencies of the instructions the chain is made of.  Stand-alone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ubench_fp64 tools/ubench_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ double seed(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double order2(double x) {
  const double y0 = seed(x), h = 0.5 * y0, e = fma(-(x * y0), y0, 1.0);
  return fma(h, e, y0);
}
__device__ __forceinline__ double order3(double x) {
  const double y0 = seed(x), e = fma(-(x * y0), y0, 1.0);
  return fma(y0 * e, fma(0.375, e, 0.5), y0);
}

__global__ void accuracy(double* out) {   // out[0..2] = max relative error of seed / order2 / order3
  const int t = blockIdx.x * blockDim.x + threadIdx.x, n = gridDim.x * blockDim.x;
  double m0 = 0, m2 = 0, m3 = 0;
  for (int i = t; i < (1 << 24); i += n) {
    // mantissas sweep [1, 4) densely; exponents from 1e-60 to 1e60
    const double mant = 1.0 + 3.0 * (static_cast<double>(i) + 0.37) / (1 << 24);
    const double x = mant * exp2(static_cast<double>((i % 401) - 200));
    const double ref = rsqrt(x);
    m0 = fmax(m0, fabs(seed(x) - ref) / ref);
    m2 = fmax(m2, fabs(order2(x) - ref) / ref);
    m3 = fmax(m3, fabs(order3(x) - ref) / ref);
  }
  __shared__ double s[3][256];
  s[0][threadIdx.x] = m0; s[1][threadIdx.x] = m2; s[2][threadIdx.x] = m3;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 0; q < 3; ++q) {
      double m = 0;
      for (int i = 0; i < 256; ++i) m = fmax(m, s[q][i]);
      atomicMax(reinterpret_cast<unsigned long long*>(out + q), static_cast<unsigned long long>(__double_as_longlong(m)));
    }
  }
}

constexpr int kN = 512;
__global__ void latency(long long* cyc, double* sink, double x0) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.0000001;
  long long t0, t1;
  __shared__ double sm[64];
  sm[threadIdx.x & 63] = x;
  __syncwarp();
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < kN; ++i) x = fma(x, y, 1e-9);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < kN; ++i) x = x * y;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < kN; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  x = fabs(x) + 1.5;
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < kN; ++i) x = seed(x) + 1.5;   // MUFU.RSQ64H + one DADD
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < kN; ++i) x = order2(x) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < kN; ++i) x = order3(x) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < kN; ++i) {   // the float-seeded form the kernel used before
    const double y0 = static_cast<double>(rsqrtf(static_cast<float>(x)));
    const double e = fma(-x * y0, y0, 1.0);
    x = fma(y0 * e, fma(0.375, e, 0.5), y0) + 1.5;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  double d0 = x, d1 = y;
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < kN; ++i)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1) : "d"(y), "d"(y));
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = t1 - t0;
  int idx = threadIdx.x & 63;
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < kN; ++i) idx = static_cast<int>(reinterpret_cast<volatile double*>(sm)[idx & 63]) & 63;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[8] = t1 - t0;
  float f = static_cast<float>(x);
  t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < kN; ++i) f = fmaf(f, 1.0000001f, 1e-9f);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[9] = t1 - t0;
  sink[threadIdx.x] = x + d0 + d1 + idx + f;
}

int main() {
  double* out; long long* cyc; double* sink;
  cudaMallocManaged(&out, 3 * sizeof(double));
  cudaMallocManaged(&cyc, 16 * sizeof(long long));
  cudaMallocManaged(&sink, 64 * sizeof(double));
  out[0] = out[1] = out[2] = 0;
  accuracy<<<296, 256>>>(out);
  latency<<<1, 32>>>(cyc, sink, 1.25);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); return 1; }
  printf("rsqrt.approx.ftz.f64 max rel err %.3e (2^%.1f); +2nd order %.3e; +3rd order %.3e\n", out[0],
         log2(out[0]), out[1], out[2]);
  const char* names[] = {"DFMA", "DMUL", "SHFL f64", "MUFU.RSQ64H + DADD", "rsqrt order2 + DADD",
                         "rsqrt order3 + DADD", "float-seeded rsqrt + DADD", "DMMA m8n8k4", "LDS.64 -> index",
                         "FFMA"};
  for (int i = 0; i < 10; ++i) printf("%-28s %6.1f cycles per dependent op\n", names[i], cyc[i] / double(kN));
  return 0;
}
