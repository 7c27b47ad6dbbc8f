// mg_lanes.cuh — the internal "lanes" (CUDA streams) a blocked factorisation is spread over.
//
// A right-looking blocked Cholesky is a chain of small, latency-bound panel kernels (potrf128 ->
// trsm128) followed by wide updates; run on one stream, the GPU idles behind the chain.  The
// drivers in mg_linalg.cu / mg_type1.cu therefore fork the caller's stream into
//   chain  : high priority — only what the next potrf128 needs: potrf128, the solve of the first
//            384 columns of the block row, one 128 x 384 tile update;
//   chain2 : high priority — the rest of the block row's solve and of the outer block's row
//            updates (overlaps the next potrf128);
//   upd    : the K = 512 update of everything below an outer block (once per 4 panels), the
//            2-CTA diagonal solves of the triangular inverse, the back-substitution updates;
//   tri    : work that only consumes finished block rows (triangular-inverse rows, the Nystrom
//            cross term and its forward substitution);
//   tri2   : the part of a triangular-inverse row that does not depend on the previous row
//            (look-ahead for the tri lane's own chain);
// and join them back before returning, so to the caller every entry point is still an ordinary
// stream-ordered call.  Hand-offs are CUDA events; adds into the same rows are ordered, so results
// do not depend on timing.  Streams and events are created once per device, two sets of them, so
// two host threads can run two factorisations side by side (no memory is allocated).
// MG_SERIAL=1 collapses all lanes onto the caller's stream (A/B measurements, debugging).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mg {

struct Lanes {
  cudaStream_t chain = nullptr, chain2 = nullptr, upd = nullptr, tri = nullptr, tri2 = nullptr;
  cudaStream_t user = nullptr;
  cudaEvent_t fork = nullptr, join[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t trsm = nullptr;              // re-recorded by every panel: "block row pj is final"
  cudaEvent_t upd_done[2] = {nullptr, nullptr};   // trailing update of panel pj (parity pj & 1)
  cudaEvent_t potrf = nullptr;                      // diagonal block of panel pj factored
  cudaEvent_t first[2] = {nullptr, nullptr};       // first 256 columns of block row pj solved (chain)
  cudaEvent_t row_rest[2] = {nullptr, nullptr};    // chain2's row updates of panel pj landed
  cudaEvent_t next_done[2] = {nullptr, nullptr};  // block update of the next outer block's rows 2..G
  cudaEvent_t diag_done[2] = {nullptr, nullptr};  // triangular-inverse diagonal block of panel pj
  cudaEvent_t row_done[2] = {nullptr, nullptr};   // triangular-inverse block row pj is final
  cudaEvent_t bulk_done[2] = {nullptr, nullptr};  // look-ahead part of the row's cross product
  cudaEvent_t misc[2] = {nullptr, nullptr};
  bool serial = true;
  int reserve_sms = 48;   // SMs kept free of persistent bulk CTAs (see LaneScope)
  bool reserve_from_env = false;   // MG_BULK_RESERVE_SMS given: use it as it is

  void record(cudaEvent_t e, cudaStream_t on) const {
    if (!serial) cudaEventRecord(e, on);
  }
  void wait(cudaStream_t who, cudaEvent_t e) const {
    if (!serial) cudaStreamWaitEvent(who, e, 0);
  }
  int bulk_cta_cap() const;   // persistent-grid cap for GEMMs on upd / tri (0 = none)
};

// The bulk GEMMs of `n` concurrent factorisations share the SMs left after the chain reserve.
void set_lane_share(int n);

// RAII: fork the caller's stream into the lanes (holding the per-device lane mutex), join on
// destruction.  `ok()` is false if the streams could not be created; the drivers then run serial.
class LaneScope {
 public:
  // `n` = order of the matrix being factored: small problems are bound by the panel chain (keep
  // a third of the SMs free for it), large ones by the bulk GEMMs (give them nearly everything).
  LaneScope(cudaStream_t user, int64_t n);
  ~LaneScope();
  LaneScope(const LaneScope&) = delete;
  LaneScope& operator=(const LaneScope&) = delete;
  Lanes& lanes() { return lanes_; }

 private:
  Lanes lanes_;
  void* lock_ = nullptr;
};

}  // namespace mg
