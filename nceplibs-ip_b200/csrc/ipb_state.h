// ipb200 — process-wide runtime state shared by the per-kind translation units.
//
// The reference keeps one single-entry SAVE cache per interpolation routine
// (src/bilinear_interp_mod.F90:104-108 etc.) and is not re-entrant.  Here every CUDA device the process uses has its
// own state (streams, staging buffers, timing events, a multi-entry LRU plan cache), created on first use and keyed by
// the device that is current when a call enters — or, for host-pointer calls after ipgpu_set_devices(n), by the
// devices the field batch is sharded over.  Calls on one device are serialised by that device's mutex; calls on
// different devices run concurrently.
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
struct HBuf {  // growable pinned host staging
  void* p = nullptr; size_t cap = 0;
  void* get(size_t bytes);
  void release();
};
struct Work { DBuf gi, vi, li, go, vo, lo, ints, lbits; HBuf hbits, hin, hout; };

struct CacheEntry { std::vector<i64> key; std::shared_ptr<void> plan; int kind; size_t bytes; unsigned long long stamp; };

// restores the caller's current device on scope exit
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (dev != prev) cudaSetDevice(dev); else prev = -1; }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct DevState {
  int device = 0;
  std::mutex mu;                         // serialises the calls that use this device's streams and staging buffers
  cudaStream_t stream[2] = {nullptr, nullptr};
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};   // timing events, reused by every call
  Work work[2];
  HBuf geo;                              // pinned staging of rlat / rlon / crot / srot for callers with pageable arrays
  std::vector<CacheEntry> cache;
  unsigned long long stamp = 0;
};

struct Runtime {
  std::mutex mu;                         // guards the device table and the settings below
  std::vector<std::unique_ptr<DevState>> dev;
  std::vector<int> shard;                // devices a host-pointer call shards its field batch over (empty: the current device)
  size_t cache_cap = (size_t)24 << 30;   // bytes of plans kept per device (180 GB of HBM: generous by default)
  size_t chunk_bytes = (size_t)768 << 20;
  double last_plan_ms = 0, last_apply_ms = 0;
  int last_plan_hit = 0;
  long long last_h2d = 0, last_d2h = 0;  // field bytes moved over PCIe by the last host-pointer call
  DevState& state(int device);           // created on first use; aborts loudly when there is no CUDA device: the library has no CPU path
  int current_device();
};
Runtime& rt();

// An explicit plan handle (ipgpu_plan_get_<k>): a reference on a device-resident plan plus the decoded options of the call
struct PlanHandle {
  int kind_bytes = 0, device = 0, method = 0;
  bool vec = false;
  std::shared_ptr<void> plan;
  i64 mi = 0, mo = 0, no = 0, iret = 0;
  double pmp = 0; int mcon = 0, mspiral = 0;
};

// Stream-ordered allocation scope of the calling thread (plan arrays come from the device's memory pool while a plan is
// built or destroyed; outside a scope dalloc / dfree fall back to cudaMalloc / cudaFree).
struct AllocCtx { bool async = false; cudaStream_t stream = nullptr; };
AllocCtx& alloc_ctx();
struct AsyncAlloc {
  AllocCtx saved;
  explicit AsyncAlloc(cudaStream_t s) { saved = alloc_ctx(); alloc_ctx().async = true; alloc_ctx().stream = s; }
  ~AsyncAlloc() { alloc_ctx() = saved; }
};

}  // namespace ipb
