// mg_prof.cuh — opt-in in-situ timing of the kernels a blocked driver enqueues (MG_PROFILE=1).
// Records a CUDA event before and after each launch on the launching stream; report() synchronises
// and prints per-label totals to stderr.  Off by default: zero events, zero synchronisation.
// MG_PROFILE_TIMELINE=<file> additionally appends one line per launch — call, label, lane (stream),
// start and end in ms since the call's first launch — from which tools/timeline_summary.py derives
// what each lane was doing and which lane the call was waiting for.
#pragma once
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace mg {

class Prof {
 public:
  static Prof& get() {
    static Prof p;
    return p;
  }
  bool on() const { return on_; }
  void tic(cudaStream_t s) {
    if (!on_) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    start_ = e;
  }
  void toc(cudaStream_t s, const char* label) {
    if (!on_) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, s);
    spans_.push_back({label, start_, e, s});
  }
  void report(cudaStream_t s, const char* title) {
    if (!on_ || spans_.empty()) return;
    (void)s;
    cudaDeviceSynchronize();   // spans live on several lanes
    std::map<std::string, std::pair<double, int>> tot;
    float first_to_last = 0.f;
    cudaEventElapsedTime(&first_to_last, spans_.front().a, spans_.back().b);
    std::FILE* tl = nullptr;
    if (const char* path = std::getenv("MG_PROFILE_TIMELINE")) tl = std::fopen(path, "a");
    for (auto& sp : spans_) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, sp.a, sp.b);
      if (tl) {
        float t0 = 0.f;
        cudaEventElapsedTime(&t0, spans_.front().a, sp.a);
        std::fprintf(tl, "%s,%s,%p,%.4f,%.4f\n", title, sp.label, static_cast<void*>(sp.s), t0, t0 + ms);
      }
      tot[sp.label].first += ms;
      tot[sp.label].second += 1;
    }
    for (auto& sp : spans_) {
      cudaEventDestroy(sp.a);
      cudaEventDestroy(sp.b);
    }
    if (tl) std::fclose(tl);
    double sum = 0;
    for (auto& kv : tot) sum += kv.second.first;
    std::fprintf(stderr, "[mg-prof] %s: %.3f ms wall, %.3f ms in kernels\n", title, first_to_last, sum);
    for (auto& kv : tot)
      std::fprintf(stderr, "[mg-prof]   %-28s n=%5d total %8.3f ms avg %7.1f us\n", kv.first.c_str(),
                   kv.second.second, kv.second.first, 1e3 * kv.second.first / kv.second.second);
    spans_.clear();
  }

 private:
  Prof() {
    const char* e = std::getenv("MG_PROFILE");
    on_ = e && e[0] == '1';
  }
  struct Span {
    const char* label;
    cudaEvent_t a, b;
    cudaStream_t s;
  };
  bool on_ = false;
  cudaEvent_t start_{};
  std::vector<Span> spans_;
};

#define MG_TIMED(stream, label, call)  \
  do {                                 \
    ::mg::Prof::get().tic(stream);     \
    call;                              \
    ::mg::Prof::get().toc(stream, label); \
  } while (0)

}  // namespace mg
