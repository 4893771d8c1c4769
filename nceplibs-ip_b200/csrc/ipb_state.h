// ipb200 — process-wide runtime state shared by the per-kind translation units.
//
// The reference keeps one single-entry SAVE cache per interpolation routine
// (src/bilinear_interp_mod.F90:104-108 etc.) and is not re-entrant.  Here all calls are
// serialised by one mutex; plans live in a multi-entry LRU cache on the device.
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <mutex>
#include <vector>

#include "ipb_grid.h"

namespace ipb {

struct DBuf {  // growable device scratch
  void* p = nullptr; size_t cap = 0;
  void* get(size_t bytes);
  void release();
};
struct Work { DBuf gi, vi, li, go, vo, lo, ints; };

struct CacheEntry { std::vector<i64> key; std::shared_ptr<void> plan; int kind; size_t bytes; unsigned long long stamp; };

struct Runtime {
  std::mutex mu;
  bool init = false;
  cudaStream_t stream[2] = {nullptr, nullptr};
  Work work[2];
  std::vector<CacheEntry> cache;
  unsigned long long stamp = 0;
  size_t cache_cap = (size_t)24 << 30;  // bytes of plans kept (180 GB of HBM: generous by default)
  size_t chunk_bytes = (size_t)768 << 20;
  double last_plan_ms = 0, last_apply_ms = 0;
  int last_plan_hit = 0;
  long long last_h2d = 0, last_d2h = 0;  // field bytes moved over PCIe by the last host-pointer call
  void ensure_init();  // aborts loudly when there is no CUDA device: the library has no CPU path
};
Runtime& rt();

}  // namespace ipb
