// Throughput companion of ubench_fp64.cu: one CTA of 512 threads (16 warps) on one SM, independent
// accumulation chains per warp; cycles per warp-instruction per SM for DFMA and the fp64 mma shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/ubench_fp64_tp tools/ubench_fp64_tp.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 256;

__global__ void __launch_bounds__(512, 1) tp(long long* cyc, double* sink, double y, int nwarps) {
  const int warp = threadIdx.x >> 5;
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
  long long t0 = 0, t1 = 0;
  // ---- DFMA, 8 independent chains per thread
  __syncthreads();
  t0 = clock64();
  if (warp < nwarps) {
#pragma unroll 4
    for (int i = 0; i < kIters; ++i)
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] = fma(a[q], y, 1e-9);
  }
  __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // ---- m8n8k4, 4 independent accumulator pairs
  __syncthreads();
  t0 = clock64();
  if (warp < nwarps) {
#pragma unroll 4
    for (int i = 0; i < kIters; ++i)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                     : "+d"(a[2 * q]), "+d"(a[2 * q + 1]) : "d"(y), "d"(y));
  }
  __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // ---- m16n8k4: C 4 regs, A 2, B 1; 2 independent accumulators
  __syncthreads();
  t0 = clock64();
  if (warp < nwarps) {
#pragma unroll 4
    for (int i = 0; i < kIters; ++i)
#pragma unroll
      for (int q = 0; q < 2; ++q)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
                     : "+d"(a[4 * q]), "+d"(a[4 * q + 1]), "+d"(a[4 * q + 2]), "+d"(a[4 * q + 3])
                     : "d"(y), "d"(y), "d"(y));
  }
  __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // ---- m16n8k8: A 4, B 2
  __syncthreads();
  t0 = clock64();
  if (warp < nwarps) {
#pragma unroll 4
    for (int i = 0; i < kIters; ++i)
#pragma unroll
      for (int q = 0; q < 2; ++q)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                     : "+d"(a[4 * q]), "+d"(a[4 * q + 1]), "+d"(a[4 * q + 2]), "+d"(a[4 * q + 3])
                     : "d"(y), "d"(y), "d"(y), "d"(y), "d"(y), "d"(y));
  }
  __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // ---- m16n8k16: A 8, B 4
  __syncthreads();
  t0 = clock64();
  if (warp < nwarps) {
#pragma unroll 4
    for (int i = 0; i < kIters; ++i)
#pragma unroll
      for (int q = 0; q < 2; ++q)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7, %8, %9, %10, %11}, {%12, %13, %14, %15}, {%0, %1, %2, %3};"
                     : "+d"(a[4 * q]), "+d"(a[4 * q + 1]), "+d"(a[4 * q + 2]), "+d"(a[4 * q + 3])
                     : "d"(y), "d"(y), "d"(y), "d"(y), "d"(y), "d"(y), "d"(y), "d"(y), "d"(y), "d"(y), "d"(y), "d"(y));
  }
  __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  sink[threadIdx.x] = s;
}

int main() {
  long long* cyc; double* sink;
  cudaMallocManaged(&cyc, 16 * sizeof(long long));
  cudaMallocManaged(&sink, 512 * sizeof(double));
  const char* names[] = {"DFMA", "mma m8n8k4", "mma m16n8k4", "mma m16n8k8", "mma m16n8k16"};
  const int per_iter[] = {8, 4, 2, 2, 2};
  const double fma_per_op[] = {32, 256, 512, 1024, 2048};
  for (int nw : {1, 4, 16}) {
    tp<<<1, 512>>>(cyc, sink, 1.0000001, nw);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("cuda error\n"); return 1; }
    printf("%d warp(s) on one SM:\n", nw);
    for (int i = 0; i < 5; ++i) {
      const double ops = double(kIters) * per_iter[i] * nw;
      printf("  %-14s %7.2f cycles per warp-instruction per SM, %6.1f FMA/clk/SM\n", names[i], cyc[i] / ops,
             fma_per_op[i] * ops / cyc[i]);
    }
  }
  return 0;
}
