// mg_once.cuh — "once per device" guard for per-function attributes.
//
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-DEVICE property of a kernel: a
// process-wide `static bool` would leave the opt-in unset on every GPU but the first one a
// process touches (and is a data race between host threads).  One flag per device ordinal;
// concurrent first calls may both run `f` (idempotent), later calls are one relaxed load.
#pragma once
#include <cuda_runtime.h>

#include <atomic>

namespace mg {

struct PerDeviceOnce {
  static constexpr int kMaxDevices = 64;
  std::atomic<bool> done[kMaxDevices] = {};

  // f() -> cudaError_t.  Returns 0 or -1000 - cudaError_t.
  template <class F>
  int run(F&& f) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = -1;
    const bool tracked = dev >= 0 && dev < kMaxDevices;
    if (tracked && done[dev].load(std::memory_order_acquire)) return 0;
    const cudaError_t e = f();
    if (e != cudaSuccess) return -1000 - static_cast<int>(e);
    if (tracked) done[dev].store(true, std::memory_order_release);
    return 0;
  }
};

}  // namespace mg
