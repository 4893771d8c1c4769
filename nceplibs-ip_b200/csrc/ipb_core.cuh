// ipb200 — host orchestration of one ipolates / ipolatev / gdswzd call, templated on the real kind.
// Instantiated once per kind in ipb_kind_f32.cu / ipb_kind_f64.cu; the C ABI lives in ipb_api.cu.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <thread>

#include "ipb_apply.cuh"
#include "ipb_fast.cuh"
#include "ipb_tiles.cuh"
#include "ipb_budget.cuh"
#include "ipb_gtiles.cuh"
#include "ipb_box.cuh"
#include "ipb_expand.cuh"
#include "ipb_state.h"
#include "ipb_hostcopy.h"

namespace ipb {

template <class R>
std::vector<i64> plan_key(int method, bool vec, const i64* ipopt, const GridHost<R>& gin, const GridHost<R>& gout, i64 mi, i64 mo) {
  std::vector<i64> k;
  k.push_back(sizeof(R)); k.push_back(method); k.push_back(vec); k.push_back(mi); k.push_back(mo);
  k.push_back((i64)gin.key.size());
  k.insert(k.end(), gin.key.begin(), gin.key.end());
  k.insert(k.end(), gout.key.begin(), gout.key.end());
  if (method == M_BUDGET || method == M_NBUDGET) k.insert(k.end(), ipopt, ipopt + 19);  // sub-sample geometry and weights
  return k;
}

template <class R> std::shared_ptr<Plan<R>> cache_find(DevState& DS, const std::vector<i64>& key) {
  for (auto& e : DS.cache)
    if (e.kind == (int)sizeof(R) && e.key == key) { e.stamp = ++DS.stamp; return std::static_pointer_cast<Plan<R>>(e.plan); }
  return nullptr;
}
template <class R> void cache_put(DevState& DS, const std::vector<i64>& key, std::shared_ptr<Plan<R>> p) {
  const size_t cap = rt().cache_cap;
  size_t tot = p->bytes;
  for (auto& e : DS.cache) tot += e.bytes;
  while (tot > cap && !DS.cache.empty()) {
    auto it = std::min_element(DS.cache.begin(), DS.cache.end(), [](const CacheEntry& a, const CacheEntry& b) { return a.stamp < b.stamp; });
    tot -= it->bytes;
    DS.cache.erase(it);
  }
  DS.cache.push_back({key, std::static_pointer_cast<void>(p), (int)sizeof(R), p->bytes, ++DS.stamp});
}
// the cache entry's size follows a plan that grows after it was cached (tables added on first need)
template <class R> void cache_resize(DevState& DS, const std::shared_ptr<Plan<R>>& p) {
  for (auto& e : DS.cache) if (e.plan.get() == (void*)p.get()) e.bytes = p->bytes;
}

// ---- one launch of the apply stage over kc fields already resident on the device ------------------
template <class R>
void launch_apply(const Plan<R>& P, bool vec, FieldArgs<R> a, cudaStream_t st) {
  const int no = P.no;
  if (no <= 0 || a.kc <= 0) return;
  const GridP<R>& gi = P.gin.p;
  const int method = P.method;
  if (gtiles_ok(P, a)) { launch_gtiles(P, a, st); return; }
  if (vec && gtiles_vec_ok(P, a)) { launch_gtiles_vec(P, a, st); return; }
  if (method == M_BILINEAR || method == M_BICUBIC) {
    if (method == M_BILINEAR && !vec && box_ok(P, a) && launch_bilinear_box(P, a, st)) return;
    if (method == M_BILINEAR && !vec && tiles_ok(P, a)) { launch_bilinear_tiles(P, a, st); return; }
    if (method == M_BILINEAR && !vec && fast_bilinear_ok(P, a)) { launch_fast_bilinear(P, a, st); return; }
    const int kpb = std::max(1, std::min(a.kc, 64));   // fields per thread: the plan entry is re-read once per kpb fields
    const int rowlen = patch_rowlen(P.gout.p, no);
    dim3 grid = patch_grid(rowlen, no, 4);
    grid.z = nblk(a.kc, kpb);
    if (method == M_BILINEAR) {
      if (vec) k_apply_stencil<R, 2, true><<<grid, 128, 0, st>>>(a, gi, kpb, P.nxy, P.wxy, P.cxy, P.sxy, P.nc, P.x, P.y, P.crot, P.srot, rowlen);
      else k_apply_stencil<R, 2, false><<<grid, 128, 0, st>>>(a, gi, kpb, P.nxy, P.wxy, nullptr, nullptr, P.nc, P.x, P.y, nullptr, nullptr, rowlen);
    } else {
      if (vec) k_apply_stencil<R, 4, true><<<grid, 128, 0, st>>>(a, gi, kpb, P.nxy, P.wxy, P.cxy, P.sxy, P.nc, P.x, P.y, P.crot, P.srot, rowlen);
      else k_apply_stencil<R, 4, false><<<grid, 128, 0, st>>>(a, gi, kpb, P.nxy, P.wxy, nullptr, nullptr, P.nc, P.x, P.y, nullptr, nullptr, rowlen);
    }
    g_launches++;
  } else if (method == M_NEIGHBOR) {
    if (!vec && fast_neighbor_ok(P, a)) { launch_fast_neighbor(P, a, st); return; }
    const int kpb = std::max(1, std::min(a.kc, 32));
    dim3 grid(nblk(no, 256), nblk(a.kc, kpb));
    if (vec) k_apply_neighbor<R, true><<<grid, 256, 0, st>>>(a, gi, kpb, P.nxy, P.cxy, P.sxy, P.x, P.y, P.rlat, P.rlon, P.crot, P.srot);
    else k_apply_neighbor<R, false><<<grid, 256, 0, st>>>(a, gi, kpb, P.nxy, nullptr, nullptr, P.x, P.y, P.rlat, P.rlon, nullptr, nullptr);
    g_launches++;
  } else {
    constexpr int KF = 8;   // fields per thread: the sub-sample taps are re-read once per KF fields
    dim3 grid(nblk(no, 128), nblk(a.kc, KF));
    const int ns = P.bo.nsamp;
    if (budget_collapsed_ok(P, a)) { launch_budget_collapsed(P, a, st); return; }
    if (method == M_BUDGET && !vec && a.mspiral <= 1 && ns > 0 && (size_t)a.kc * (size_t)a.mi < ((size_t)1 << 32)) {   // 32-bit element offsets inside
      k_budget_scalar<R, KF><<<dim3(nblk(no, 32), nblk(a.kc, 8 * KF)), 256, 0, st>>>(a, ns, P.bn, P.bw, P.d_wb);
      g_launches++;
      return;
    }
    if (method == M_BUDGET) {
      if (vec) k_apply_budget<R, false, true, KF><<<grid, 128, 0, st>>>(a, gi, ns, P.bn, P.bw, P.bc, P.bs, P.d_wb, P.rlat, P.rlon, P.crot, P.srot);
      else k_apply_budget<R, false, false, KF><<<grid, 128, 0, st>>>(a, gi, ns, P.bn, P.bw, nullptr, nullptr, P.d_wb, P.rlat, P.rlon, nullptr, nullptr);
    } else {
      if (vec) k_apply_budget<R, true, true, KF><<<grid, 128, 0, st>>>(a, gi, ns, P.bn, nullptr, P.bc, P.bs, P.d_wb, P.rlat, P.rlon, P.crot, P.srot);
      else k_apply_budget<R, true, false, KF><<<grid, 128, 0, st>>>(a, gi, ns, P.bn, nullptr, nullptr, nullptr, P.d_wb, P.rlat, P.rlon, nullptr, nullptr);
    }
    g_launches++;
  }
}

// ibo(k) = ibi(k), forced to 1 where any lo is false; then the pole fix for lat-lon outputs.
// (not cudaMemsetAsync: a memset may be given to a copy engine, where it queues behind the other chunk's 8 ms go copy
// and holds this chunk's kernels that long)
static __global__ void k_zero_ints(int n, int* p) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = 0; }
static __global__ void k_ibo(int kc, const int* __restrict__ ibi, const int* __restrict__ flag, int* ibo) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < kc) ibo[k] = flag[k] ? 1 : ibi[k];
}
template <class R>
void launch_post(const Plan<R>& P, bool vec, const FieldArgs<R>& a, int* d_ibo, cudaStream_t st) {
  if (a.kc <= 0) return;
  k_ibo<<<nblk(a.kc, 128), 128, 0, st>>>(a.kc, a.ibi, a.flag, d_ibo);
  g_launches++;
  if (P.gout.p.type == P_EQUID && P.no > 0) {
    const int wpb = 4;
    for (int south = 0; south < 2; south++) {
      const int np = south ? P.nps : P.npn;
      if (np <= 1) continue;
      const int* idx = south ? P.pole_s : P.pole_n;
      if (vec) k_polfix<R, true><<<nblk(a.kc, wpb), wpb * 32, 0, st>>>(a.kc, a.mo, d_ibo, np, idx, south ? P.pclon_s : P.pclon_n, south ? P.pslon_s : P.pslon_n, south, a.go, a.vo, a.lo);
      else k_polfix<R, false><<<nblk(a.kc, wpb), wpb * 32, 0, st>>>(a.kc, a.mo, d_ibo, np, idx, nullptr, nullptr, south, a.go, nullptr, a.lo);
      g_launches++;
    }
  }
}

// ---- lo across PCIe as one bit per point -----------------------------------------------------------------------------
// The host-pointer call is bound by the device -> host link (go + lo: 5 bytes per point·field in the _4 kind), and lo is a
// fifth of that while it carries one bit.  The device packs lo (bit i of word w = point 32 w + i) and notes per field
// whether any point is false; the host thread(s) write the caller's LOGICAL*1 array from that: memset for a field
// without a false point, a table lookup per byte of bits otherwise.  The bytes the caller sees are the same.
static __global__ void __launch_bounds__(256)
k_pack_lo(const u8* __restrict__ lo, int mo, int no, int nw, unsigned* __restrict__ bits, int* __restrict__ hasfalse) {
  const int lane = threadIdx.x & 31, k = blockIdx.y;
  const int wbase = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32;
  if (wbase >= nw) return;
  const u8* __restrict__ p = lo + (size_t)k * mo;
  unsigned mine = 0, bad = 0;
#pragma unroll 4
  for (int i = 0; i < 32; i++) {
    if (wbase + i >= nw) break;
    const int n = (wbase + i) * 32 + lane;
    const bool in = n < no;
    const bool t = in && p[n] != 0;
    const unsigned word = __ballot_sync(0xffffffffu, t);
    bad |= __ballot_sync(0xffffffffu, in && !t);
    if (lane == i) mine = word;
  }
  if (wbase + lane < nw) bits[(size_t)k * nw + wbase + lane] = mine;
  if (bad != 0 && lane == 0 && hasfalse[k] == 0) hasfalse[k] = 1;
}

// is p memory the copy engines can reach directly (cudaHostAlloc / cudaHostRegister / managed)?  An ordinary Fortran or
// C array is not: the driver would stage every copy through its own small buffer and hold the host meanwhile.
inline bool host_ptr_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged || at.type == cudaMemoryTypeDevice;
}

// ---- options and plans -----------------------------------------------------------------------------
// option decode of the stencil methods (bilinear_interp_mod.F90:110-115, bicubic :117-122, neighbor :135); the budget
// family's threshold comes from its plan (decode_budget_opt).  Returns iret (0 or 32).
template <class R> struct ApplyOpts { R pmp = 0; int mcon = 0, mspiral = 0; };
template <class R> i64 decode_opts(int method, const i64* ipopt, ApplyOpts<R>& o) {
  i64 mp = 50, mcon = 0, mspiral = 0, iret = 0;
  if (method == M_BILINEAR) { mp = ipopt[0]; mspiral = std::max<i64>(ipopt[1], 0); }
  else if (method == M_BICUBIC) { mcon = ipopt[0]; mp = ipopt[1]; }
  else if (method == M_NEIGHBOR) { mspiral = std::max<i64>(ipopt[0], 1); }
  else mspiral = std::max<i64>(ipopt[19], 1);
  if (method == M_BILINEAR || method == M_BICUBIC) {
    if (mp == -1 || mp == 0) mp = 50;
    if (mp < 0 || mp > 100) iret = 32;
  }
  o.pmp = (R)mp * IPB_RC(0.01); o.mcon = (int)mcon; o.mspiral = (int)mspiral;
  return iret;
}

// The plan of a call: fetched from the device's cache or built on `st` (never cached for station points, whose
// coordinates are call arguments).  Caller holds DS.mu and has DS.device current.
template <class R>
std::shared_ptr<Plan<R>> obtain_plan(DevState& DS, int method, bool vec, const i64* ipopt, const GridHost<R>& gin, const GridHost<R>& gout,
                                     i64 mi, i64 mo, i64 no_in, const R* rlat, const R* rlon, const R* crot, const R* srot, bool any_mask,
                                     cudaStream_t st, int& hit) {
  const bool station = gout.p.type == P_STATION;
  std::shared_ptr<Plan<R>> P;
  std::vector<i64> key = plan_key<R>(method, vec, ipopt, gin, gout, mi, mo);
  hit = 0;
  if (!station) { P = cache_find<R>(DS, key); if (P) hit = 1; }
  if (!P) {
    AsyncAlloc scope(st);   // plan arrays come from the memory pool, ordered on the stream that builds and first uses them
    P.reset(build_plan<R>(method, vec, ipopt, gin, gout, mi, mo, no_in, rlat, rlon, crot, srot, st));
    P->free_stream = DS.stream[0];
    if (P->iret == 0 && !station) { build_tile_tables<R>(*P, st); P->bytes += P->tiles.bytes; build_box_tables<R>(*P, st); }
    if (P->iret == 0) build_budget_collapse<R>(*P, st);
    if (P->iret == 0 && !station && (method != M_BILINEAR || vec)) { build_gtiles<R>(*P, st); P->bytes += P->gt.bytes; }
    if (!station) cache_put<R>(DS, key, P);
  }
  // bilinear calls that carry a bitmap use the point-staged kernel; its tables are added to the plan on first need
  if (P->iret == 0 && !station && method == M_BILINEAR && !vec && !P->gt.ok && !P->gt_tried && any_mask) {
    AsyncAlloc scope(st);
    P->gt_tried = true;
    build_gtiles<R>(*P, st);
    P->bytes += P->gt.bytes;
    cache_resize<R>(DS, P);
  }
  return P;
}

// ---- the driver shared by ipolates / ipolatev, on one device ---------------------------------------
// dev = 1: li/gi/vi/lo/go/vo are device pointers and everything runs on `user_stream`.
template <class R>
void run_on(int device, bool stats, bool vec, int dev, void* user_stream, i64 ip, const i64* ipopt, const GridHost<R>& gin,
            const GridHost<R>& gout, i64 mi, i64 mo, i64 km, const i64* ibi, const u8* li, const R* gi, const R* vi, i64& no, R* rlat,
            R* rlon, R* crot, R* srot, i64* ibo, u8* lo, R* go, R* vo, i64& iret) {
  Runtime& RT = rt();
  DeviceGuard guard(device);
  DevState& DS = RT.state(device);
  std::lock_guard<std::mutex> lock(DS.mu);
  iret = 0;
  const int method = (int)ip;
  const bool budget = method == M_BUDGET || method == M_NBUDGET;
  const bool station = gout.p.type == P_STATION;
  cudaStream_t st0 = dev ? (cudaStream_t)user_stream : DS.stream[0];
  ApplyOpts<R> o;
  iret = decode_opts<R>(method, ipopt, o);
  R pmp = o.pmp;
  const i64 mcon = o.mcon, mspiral = o.mspiral;
  if (!budget && iret != 0) { if (!station) no = 0; return; }

  bool any_mask = false;
  for (i64 k = 0; k < km; k++) if (ibi[k] != 0) any_mask = true;
  cudaEventRecord(DS.ev[0], st0);
  int hit = 0;
  std::shared_ptr<Plan<R>> P = obtain_plan<R>(DS, method, vec, ipopt, gin, gout, mi, mo, no, rlat, rlon, crot, srot, any_mask && km > 0, st0, hit);
  long long h2d = 0, d2h = 0;
  cudaEventRecord(DS.ev[1], st0);
  const int NO = P->no;
  // lat/lon (and rotations) are outputs for real grids; for stations the budget-family vector
  // routines overwrite crot/srot (see k_out_geo)
  // (in the device-pointer entry points these four are device arrays, or null to skip the copy)
  // A pageable destination would hold the host for the length of the copy (and of whatever is queued before it): such
  // arrays are filled from a pinned staging buffer once the stream has drained (finish_geo).
  struct GeoLate { void* dst; const void* src; size_t bytes; };
  std::vector<GeoLate> geo_late;
  auto copy_geo = [&](cudaStream_t s) {
    if (mo <= 0) return;
    const size_t nb = (size_t)mo * sizeof(R);
    unsigned char* stage = nullptr;
    int used = 0;
    auto one = [&](R* dst, const R* src) {
      if (!dst) return;
      if (!dev && !host_ptr_pinned(dst)) {
        if (!stage) stage = (unsigned char*)DS.geo.get(4 * nb);
        IPB_CUDA(cudaMemcpyAsync(stage + used * nb, src, nb, cudaMemcpyDeviceToHost, s));
        geo_late.push_back({dst, stage + used * nb, nb});
        used++;
      } else {
        IPB_CUDA(cudaMemcpyAsync(dst, src, nb, cudaMemcpyDefault, s));
      }
    };
    if (!station) { one(rlat, P->rlat); one(rlon, P->rlon); }
    if (vec && (!station || budget)) { one(crot, P->crot); one(srot, P->srot); }
  };
  auto finish_geo = [&]() { for (const GeoLate& g : geo_late) memcpy(g.dst, g.src, g.bytes); geo_late.clear(); };
  if (budget) {
    // the budget family has no early exit: with bad options it still runs its finish loop
    // (budget_interp_mod.F90:178-181,288-344)
    iret = P->iret;
    pmp = ((R)P->bo.mp * IPB_RC(0.01)) * (R)P->bo.nb4;
    no = NO;
  } else if (P->iret != 0) {
    iret = P->iret;
    copy_geo(st0);
    IPB_CUDA(cudaStreamSynchronize(st0));
    finish_geo();
    if (!station) no = 0;
    return;
  } else {
    no = station ? mo : NO;
  }
  copy_geo(st0);

  std::vector<int> ibi32(std::max<i64>(km, 1));
  for (i64 k = 0; k < km; k++) ibi32[k] = ibi[k] != 0 ? (int)ibi[k] : 0;
  std::vector<int> ibo32(std::max<i64>(km, 1), 0);

  if (NO > 0 && km > 0) {
    if (dev) {
      // whole batch in one pass on the caller's stream
      int* d_int = (int*)DS.work[0].ints.get((size_t)km * 3 * sizeof(int));
      IPB_CUDA(cudaMemcpyAsync(d_int, ibi32.data(), km * sizeof(int), cudaMemcpyHostToDevice, st0));
      IPB_CUDA(cudaMemsetAsync(d_int + km, 0, km * sizeof(int), st0));
      FieldArgs<R> a{NO, (int)mi, (int)mo, (int)km, gi, vi, any_mask ? li : nullptr, d_int, go, vo, lo, d_int + km, pmp, (int)mcon, (int)mspiral};
      launch_apply<R>(*P, vec, a, st0);
      launch_post<R>(*P, vec, a, d_int + 2 * km, st0);
      IPB_CUDA(cudaMemcpyAsync(ibo32.data(), d_int + 2 * km, km * sizeof(int), cudaMemcpyDeviceToHost, st0));
      IPB_CUDA(cudaStreamSynchronize(st0));
    } else {
      // host arrays: stream the batch through the device in chunks, two in flight
      const size_t per_field = ((size_t)mi + (size_t)mo) * sizeof(R) * (vec ? 2 : 1) + (any_mask ? mi : 0) + mo;
      i64 kc = std::max<i64>(1, std::min<i64>(km, (i64)(RT.chunk_bytes / std::max<size_t>(per_field, 1))));
      if (km > 1 && kc >= km) kc = (km + 1) / 2;  // at least two chunks so copies and kernels overlap
      kc = std::min<i64>(kc, 32768);               // fields are a grid dimension of the per-chunk kernels
      // Pageable caller arrays (what a Fortran caller has) go through the slot's own pinned staging, copied by host threads:
      // left to the driver such copies run at a fraction of the link rate and hold the host, so nothing overlaps.
      static const bool no_stage = getenv("IPB_NO_STAGE") && atoi(getenv("IPB_NO_STAGE")) != 0;
      const bool stage_in = !no_stage && !(host_ptr_pinned(gi) && host_ptr_pinned(vi) && (!any_mask || host_ptr_pinned(li)));
      const bool stage_out = !no_stage && !(host_ptr_pinned(go) && host_ptr_pinned(vo));
      if (stage_in || stage_out) kc = std::max<i64>(1, std::min<i64>(kc, (i64)(((size_t)192 << 20) / std::max<size_t>(per_field, 1))));
      IPB_CUDA(cudaStreamSynchronize(st0));        // plan ready before the second stream uses it
      size_t win_lo = 0, win_n = (size_t)mi;
      static const bool h2d_full = getenv("IPB_H2D_FULL") && atoi(getenv("IPB_H2D_FULL")) != 0;   // A/B: whole fields, plain copies
      if (!h2d_full && mspiral <= 1 && !(method == M_BILINEAR && mspiral > 0)) {
        if (P->src_hi >= P->src_lo) { win_lo = (size_t)P->src_lo - 1; win_n = (size_t)(P->src_hi - P->src_lo + 1); }
        else { win_lo = 0; win_n = 0; }
      }
      // lo crosses the link packed (k_pack_lo) unless IPB_LO_BYTES=1 asks for the plain byte copy
      static const bool lo_bytes = getenv("IPB_LO_BYTES") && atoi(getenv("IPB_LO_BYTES")) != 0;
      static const int host_threads = getenv("IPB_HOST_THREADS") ? std::max(1, atoi(getenv("IPB_HOST_THREADS")))
                                                                 : (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
      const int nshard = (int)std::max<size_t>(1, RT.shard.size());
      const int lo_threads = std::max(1, host_threads / nshard);
      // the staged copies of pageable arrays are plain memcpy work the call waits for: every core the shard may use
      const int copy_threads = getenv("IPB_HOST_THREADS") ? lo_threads : std::max(lo_threads, (int)std::min(16u, std::thread::hardware_concurrency()) / nshard);
      const i64 nw = ((i64)NO + 31) / 32;
      // Everything a chunk brings back besides go sits in the slot's PINNED staging (ibo, packed lo): a device -> host copy
      // into pageable memory would hold the host until the chunk has drained, and the two chunks would never overlap.
      struct Pending { i64 k0 = 0, kn = 0; bool on = false; const int* ints = nullptr; const unsigned* bits = nullptr; const R* hgo = nullptr; } pend[2];
      auto finish_chunk = [&](Pending q) {   // the chunk's stream has drained
        if (!q.on) return;
        for (i64 k = 0; k < q.kn; k++) ibo32[q.k0 + k] = q.ints[k];
        if (q.hgo) {   // staged outputs: [kn][NO] of go, then of vo
          copy_rows(go + (size_t)q.k0 * mo, mo * sizeof(R), q.hgo, (size_t)NO * sizeof(R), (size_t)NO * sizeof(R), (size_t)q.kn, copy_threads);
          if (vec) copy_rows(vo + (size_t)q.k0 * mo, mo * sizeof(R), q.hgo + (size_t)q.kn * NO, (size_t)NO * sizeof(R), (size_t)NO * sizeof(R), (size_t)q.kn, copy_threads);
        }
        if (q.bits) expand_lo(lo + (size_t)q.k0 * mo, mo, NO, q.kn, q.bits, nw, q.ints + q.kn, lo_threads);
      };
      // two staging halves per slot: the next chunk is queued on the slot BEFORE the host turns to the finished one's lo
      const size_t stage_bytes = (size_t)3 * kc * sizeof(int) + (lo_bytes ? 0 : (size_t)kc * nw * sizeof(unsigned));
      i64 chunk_no = 0;
      int ci = 0;
      try {
      for (i64 k0 = 0; k0 < km; k0 += kc, ci ^= 1) {
        const i64 kn = std::min(kc, km - k0);
        cudaStream_t s = DS.stream[ci];
        Work& w = DS.work[ci];
        IPB_CUDA(cudaStreamSynchronize(s));  // buffers of this slot are free again
        const Pending done = pend[ci];
        pend[ci].on = false;
        R* d_gi = (R*)w.gi.get((size_t)kn * mi * sizeof(R));
        R* d_vi = vec ? (R*)w.vi.get((size_t)kn * mi * sizeof(R)) : nullptr;
        u8* d_li = any_mask ? (u8*)w.li.get((size_t)kn * mi) : nullptr;
        R* d_go = (R*)w.go.get((size_t)kn * mo * sizeof(R));
        R* d_vo = vec ? (R*)w.vo.get((size_t)kn * mo * sizeof(R)) : nullptr;
        u8* d_lo = (u8*)w.lo.get((size_t)kn * mo);
        int* d_int = (int*)w.ints.get((size_t)kn * 4 * sizeof(int));
        // only the window of the input field the plan refers to crosses PCIe (the spiral searches can leave it)
        h2d += (long long)win_n * kn * (sizeof(R) * (vec ? 2 : 1) + (any_mask ? 1 : 0));
        if (stage_in && win_n > 0) {
          const size_t fb = (size_t)kn * win_n * sizeof(R);
          unsigned char* h = (unsigned char*)w.hin.get(fb * (vec ? 2 : 1) + (any_mask ? (size_t)kn * win_n : 0));
          copy_rows(h, win_n * sizeof(R), gi + (size_t)k0 * mi + win_lo, mi * sizeof(R), win_n * sizeof(R), (size_t)kn, copy_threads);
          IPB_CUDA(cudaMemcpy2DAsync(d_gi + win_lo, mi * sizeof(R), h, win_n * sizeof(R), win_n * sizeof(R), kn, cudaMemcpyHostToDevice, s));
          if (vec) {
            copy_rows(h + fb, win_n * sizeof(R), vi + (size_t)k0 * mi + win_lo, mi * sizeof(R), win_n * sizeof(R), (size_t)kn, copy_threads);
            IPB_CUDA(cudaMemcpy2DAsync(d_vi + win_lo, mi * sizeof(R), h + fb, win_n * sizeof(R), win_n * sizeof(R), kn, cudaMemcpyHostToDevice, s));
          }
          if (any_mask) {
            unsigned char* hl = h + fb * (vec ? 2 : 1);
            copy_rows(hl, win_n, li + (size_t)k0 * mi + win_lo, mi, win_n, (size_t)kn, copy_threads);
            IPB_CUDA(cudaMemcpy2DAsync(d_li + win_lo, mi, hl, win_n, win_n, kn, cudaMemcpyHostToDevice, s));
          }
        } else if (win_lo == 0 && win_n == (size_t)mi) {
          IPB_CUDA(cudaMemcpyAsync(d_gi, gi + (size_t)k0 * mi, (size_t)kn * mi * sizeof(R), cudaMemcpyHostToDevice, s));
          if (vec) IPB_CUDA(cudaMemcpyAsync(d_vi, vi + (size_t)k0 * mi, (size_t)kn * mi * sizeof(R), cudaMemcpyHostToDevice, s));
          if (any_mask) IPB_CUDA(cudaMemcpyAsync(d_li, li + (size_t)k0 * mi, (size_t)kn * mi, cudaMemcpyHostToDevice, s));
        } else if (win_n > 0) {
          IPB_CUDA(cudaMemcpy2DAsync(d_gi + win_lo, mi * sizeof(R), gi + (size_t)k0 * mi + win_lo, mi * sizeof(R), win_n * sizeof(R), kn, cudaMemcpyHostToDevice, s));
          if (vec) IPB_CUDA(cudaMemcpy2DAsync(d_vi + win_lo, mi * sizeof(R), vi + (size_t)k0 * mi + win_lo, mi * sizeof(R), win_n * sizeof(R), kn, cudaMemcpyHostToDevice, s));
          if (any_mask) IPB_CUDA(cudaMemcpy2DAsync(d_li + win_lo, mi, li + (size_t)k0 * mi + win_lo, mi, win_n, kn, cudaMemcpyHostToDevice, s));
        }
        const size_t nbits = lo_bytes ? 0 : (size_t)kn * nw * sizeof(unsigned);
        int* h_int = (int*)((unsigned char*)w.hbits.get(2 * stage_bytes) + ((chunk_no >> 1) & 1) * stage_bytes);   // [ibi | ibo | field has a false point][packed lo]
        for (i64 k = 0; k < kn; k++) h_int[k] = ibi32[k0 + k];
        IPB_CUDA(cudaMemcpyAsync(d_int, h_int, kn * sizeof(int), cudaMemcpyHostToDevice, s));
        k_zero_ints<<<nblk(3 * kn, 128), 128, 0, s>>>((int)(3 * kn), d_int + kn);   // flag, ibo, has-a-false-point
        FieldArgs<R> a{NO, (int)mi, (int)mo, (int)kn, d_gi, d_vi, d_li, d_int, d_go, d_vo, d_lo, d_int + kn, pmp, (int)mcon, (int)mspiral};
        launch_apply<R>(*P, vec, a, s);
        launch_post<R>(*P, vec, a, d_int + 2 * kn, s);
        // The small results first (packed lo, ibo), the big go copy last: a small copy queued behind it would wait for the
        // OTHER chunk's go on the copy engine and hold this slot — and the next chunk's kernels — that long.
        if (!lo_bytes) {
          unsigned* d_bits = (unsigned*)w.lbits.get(nbits);
          k_pack_lo<<<dim3(nblk(nblk(nw, 32), 8), (unsigned)kn), 256, 0, s>>>(d_lo, (int)mo, NO, (int)nw, d_bits, d_int + 3 * kn);
          g_launches++;
          IPB_CUDA(cudaMemcpyAsync(h_int + 3 * kn, d_bits, nbits, cudaMemcpyDeviceToHost, s));
          d2h += (long long)nbits;
        }
        // ibo and the has-a-false-point flags are neighbours on the device (d_int + 2 kn, + 3 kn): one copy
        IPB_CUDA(cudaMemcpyAsync(h_int + kn, d_int + 2 * kn, (lo_bytes ? 1 : 2) * kn * sizeof(int), cudaMemcpyDeviceToHost, s));
        d2h += (long long)((lo_bytes ? 1 : 2) * kn * sizeof(int));
        if (lo_bytes) {
          d2h += (long long)NO * kn;
          if (NO == mo) IPB_CUDA(cudaMemcpyAsync(lo + (size_t)k0 * mo, d_lo, (size_t)kn * mo, cudaMemcpyDeviceToHost, s));
          else IPB_CUDA(cudaMemcpy2DAsync(lo + (size_t)k0 * mo, mo, d_lo, mo, NO, kn, cudaMemcpyDeviceToHost, s));
        }
        // only the first NO points of each field are defined (the reference leaves the rest untouched)
        d2h += (long long)NO * kn * (sizeof(R) * (vec ? 2 : 1));
        pend[ci].hgo = nullptr;
        if (stage_out) {
          const size_t out_bytes = (size_t)kc * NO * sizeof(R) * (vec ? 2 : 1);   // two halves, like the small staging above
          R* h = (R*)((unsigned char*)w.hout.get(2 * out_bytes) + ((chunk_no >> 1) & 1) * out_bytes);
          IPB_CUDA(cudaMemcpy2DAsync(h, (size_t)NO * sizeof(R), d_go, mo * sizeof(R), (size_t)NO * sizeof(R), kn, cudaMemcpyDeviceToHost, s));
          if (vec) IPB_CUDA(cudaMemcpy2DAsync(h + (size_t)kn * NO, (size_t)NO * sizeof(R), d_vo, mo * sizeof(R), (size_t)NO * sizeof(R), kn, cudaMemcpyDeviceToHost, s));
          pend[ci].hgo = h;
        } else if (NO == mo) {
          IPB_CUDA(cudaMemcpyAsync(go + (size_t)k0 * mo, d_go, (size_t)kn * mo * sizeof(R), cudaMemcpyDeviceToHost, s));
          if (vec) IPB_CUDA(cudaMemcpyAsync(vo + (size_t)k0 * mo, d_vo, (size_t)kn * mo * sizeof(R), cudaMemcpyDeviceToHost, s));
        } else {
          IPB_CUDA(cudaMemcpy2DAsync(go + (size_t)k0 * mo, mo * sizeof(R), d_go, mo * sizeof(R), NO * sizeof(R), kn, cudaMemcpyDeviceToHost, s));
          if (vec) IPB_CUDA(cudaMemcpy2DAsync(vo + (size_t)k0 * mo, mo * sizeof(R), d_vo, mo * sizeof(R), NO * sizeof(R), kn, cudaMemcpyDeviceToHost, s));
        }
        pend[ci].k0 = k0; pend[ci].kn = kn; pend[ci].on = true;
        pend[ci].ints = h_int + kn; pend[ci].bits = lo_bytes ? nullptr : reinterpret_cast<const unsigned*>(h_int + 3 * kn);
        chunk_no++;
        finish_chunk(done);
      }
      // the older chunk first: its lo is written while the newer one's go is still on the link
      IPB_CUDA(cudaStreamSynchronize(DS.stream[ci]));
      finish_chunk(pend[ci]);
      IPB_CUDA(cudaStreamSynchronize(DS.stream[ci ^ 1]));
      finish_chunk(pend[ci ^ 1]);
      } catch (...) {
        // no copy may still be writing into the caller's arrays (or reading the staging) when the error reaches the caller
        cudaStreamSynchronize(DS.stream[0]);
        cudaStreamSynchronize(DS.stream[1]);
        throw;
      }
    }
    for (i64 k = 0; k < km; k++) ibo[k] = ibo32[k];
  } else {
    IPB_CUDA(cudaStreamSynchronize(st0));
    for (i64 k = 0; k < km; k++) ibo[k] = ibi[k];
  }
  finish_geo();   // every branch above has drained st0
  cudaEventRecord(DS.ev[2], st0);
  cudaEventSynchronize(DS.ev[2]);
  if (stats) {
    std::lock_guard<std::mutex> g(RT.mu);
    float ms = 0;
    cudaEventElapsedTime(&ms, DS.ev[0], DS.ev[1]); RT.last_plan_ms = ms;
    cudaEventElapsedTime(&ms, DS.ev[1], DS.ev[2]); RT.last_apply_ms = ms;
    RT.last_plan_hit = hit;
  }
  {
    std::lock_guard<std::mutex> g(RT.mu);
    RT.last_h2d += h2d; RT.last_d2h += d2h;
  }
}

// ---- the entry: argument checks, then one device or — host pointers after ipgpu_set_devices(n) — the field batch
// sharded over n devices, one host thread and one PCIe link each (fields are independent given the
// plan; every device builds or fetches its own copy of the plan, bit-identical by construction) -------------------------
template <class R>
void run(bool vec, int dev, void* user_stream, i64 ip, const i64* ipopt, const GridHost<R>& gin, const GridHost<R>& gout, i64 mi,
         i64 mo, i64 km, const i64* ibi, const u8* li, const R* gi, const R* vi, i64& no, R* rlat, R* rlon, R* crot, R* srot,
         i64* ibo, u8* lo, R* go, R* vo, i64& iret) {
  Runtime& RT = rt();
  iret = 0;
  if (!(ip == 0 || ip == 1 || ip == 2 || ip == 3 || ip == 6)) {
    // ip=4 (spectral) is outside this library's scope; other values make the reference error-stop.
    fprintf(stderr, "ipb200: interpolation method ip=%lld is not supported (0,1,2,3,6 are)\n", (long long)ip);
    iret = 1; return;
  }
  if (!gin.ok || !gout.ok) { fprintf(stderr, "ipb200: unrecognised grid definition\n"); iret = 1; return; }
  if (mi >= ((i64)1 << 31) || mo >= ((i64)1 << 31) || km < 0) { iret = 1; return; }
  std::vector<int> devs;
  {
    std::lock_guard<std::mutex> g(RT.mu);
    RT.last_h2d = 0; RT.last_d2h = 0; RT.last_plan_hit = 0;
    if (!dev && gout.p.type != P_STATION) devs = RT.shard;
  }
  const int nd = (int)std::min<i64>((i64)devs.size(), vec ? km : km / 2);
  if (nd <= 1) {
    run_on<R>(devs.size() == 1 ? devs[0] : RT.current_device(), true, vec, dev, user_stream, ip, ipopt, gin, gout, mi, mo, km, ibi, li, gi, vi, no,
              rlat, rlon, crot, srot, ibo, lo, go, vo, iret);
    return;
  }
  // contiguous field slices (vector pairs stay together: u and v share k); slice 0 also delivers the geometry
  std::vector<i64> nos(nd, no), irets(nd, 0);
  std::vector<int> errs(nd, 0);
  std::vector<std::thread> th;
  for (int d = 0; d < nd; d++) {
    const i64 k0 = km * d / nd, k1 = km * (d + 1) / nd;
    th.emplace_back([&, d, k0, k1]() {
      try {
        run_on<R>(devs[d], d == 0, vec, 0, nullptr, ip, ipopt, gin, gout, mi, mo, k1 - k0, ibi + k0, li ? li + (size_t)k0 * mi : nullptr,
                  gi + (size_t)k0 * mi, vi ? vi + (size_t)k0 * mi : nullptr, nos[d], d == 0 ? rlat : nullptr, d == 0 ? rlon : nullptr,
                  d == 0 ? crot : nullptr, d == 0 ? srot : nullptr, ibo + k0, lo + (size_t)k0 * mo, go + (size_t)k0 * mo,
                  vo ? vo + (size_t)k0 * mo : nullptr, irets[d]);
      } catch (cudaError_t e) { errs[d] = (int)e ? (int)e : 1; } catch (...) { errs[d] = -1; }
    });
  }
  for (auto& t : th) t.join();
  for (int d = 0; d < nd; d++) {
    if (errs[d] > 0) throw (cudaError_t)errs[d];
    if (errs[d] < 0) throw std::runtime_error("ipb200: worker failed");
  }
  no = nos[0]; iret = irets[0];
}

// ---- explicit plan handles and the non-draining apply ------------------------------------------
struct IntPack { static constexpr int N = 512; int v[N]; };
static __global__ void k_set_ints(IntPack pk, int n, int* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = pk.v[i];
}

template <class R>
PlanHandle* plan_get(bool vec, i64 ip, const i64* ipopt, const GridHost<R>& gin, const GridHost<R>& gout, i64 mi, i64 mo, i64& no, i64& iret) {
  Runtime& RT = rt();
  iret = 0; no = 0;
  if (!(ip == 0 || ip == 1 || ip == 2 || ip == 3 || ip == 6) || !gin.ok || !gout.ok || gout.p.type == P_STATION ||
      mi >= ((i64)1 << 31) || mo >= ((i64)1 << 31)) { iret = 1; return nullptr; }
  const int method = (int)ip;
  const bool budget = method == M_BUDGET || method == M_NBUDGET;
  ApplyOpts<R> o;
  iret = decode_opts<R>(method, ipopt, o);
  if (!budget && iret != 0) return nullptr;
  const int device = RT.current_device();
  DevState& DS = RT.state(device);
  std::lock_guard<std::mutex> lock(DS.mu);
  int hit = 0;
  std::shared_ptr<Plan<R>> P = obtain_plan<R>(DS, method, vec, ipopt, gin, gout, mi, mo, 0, nullptr, nullptr, nullptr, nullptr, false, DS.stream[0], hit);
  IPB_CUDA(cudaStreamSynchronize(DS.stream[0]));
  if (budget) { iret = P->iret; o.pmp = ((R)P->bo.mp * IPB_RC(0.01)) * (R)P->bo.nb4; }
  else if (P->iret != 0) { iret = P->iret; return nullptr; }
  no = P->no;
  PlanHandle* h = new PlanHandle();
  h->kind_bytes = (int)sizeof(R); h->device = device; h->method = method; h->vec = vec; h->plan = std::static_pointer_cast<void>(P);
  h->mi = mi; h->mo = mo; h->no = P->no; h->iret = iret; h->pmp = (double)o.pmp; h->mcon = o.mcon; h->mspiral = o.mspiral;
  return h;
}

// rlat / rlon / crot / srot of the plan's output grid into host or device arrays (any may be null), queued on `stream`
template <class R> int plan_geometry(PlanHandle* h, void* stream, R* rlat, R* rlon, R* crot, R* srot) {
  if (!h || h->kind_bytes != (int)sizeof(R)) return 1;
  DeviceGuard guard(h->device);
  Plan<R>* P = static_cast<Plan<R>*>(h->plan.get());
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)h->mo * sizeof(R);
  if (rlat) IPB_CUDA(cudaMemcpyAsync(rlat, P->rlat, n, cudaMemcpyDefault, s));
  if (rlon) IPB_CUDA(cudaMemcpyAsync(rlon, P->rlon, n, cudaMemcpyDefault, s));
  if (h->vec && crot) IPB_CUDA(cudaMemcpyAsync(crot, P->crot, n, cudaMemcpyDefault, s));
  if (h->vec && srot) IPB_CUDA(cudaMemcpyAsync(srot, P->srot, n, cudaMemcpyDefault, s));
  return 0;
}

// Queues the apply stage for km device-resident fields on `stream` and returns without waiting: ibi is read from the
// host before the call returns, d_ibo (km int32 on the device, may be null) receives ibo when the stream gets there.
template <class R>
int plan_apply(PlanHandle* h, void* stream, i64 km, const i64* ibi, const u8* li, const R* gi, const R* vi, int* d_ibo, u8* lo, R* go, R* vo) {
  if (!h || h->kind_bytes != (int)sizeof(R) || km < 0) return 1;
  if (km == 0 || h->no <= 0) return 0;
  DeviceGuard guard(h->device);
  Plan<R>* P = static_cast<Plan<R>*>(h->plan.get());
  cudaStream_t s = (cudaStream_t)stream;
  bool any_mask = false;
  std::vector<int> ibi32((size_t)km);
  for (i64 k = 0; k < km; k++) { ibi32[k] = ibi[k] != 0 ? (int)ibi[k] : 0; any_mask = any_mask || ibi[k] != 0; }
  if (any_mask && P->method == M_BILINEAR && !P->vec && !P->gt.ok && !P->gt_tried) {
    DevState& DS = rt().state(h->device);
    std::lock_guard<std::mutex> lock(DS.mu);
    AsyncAlloc scope(s);
    P->gt_tried = true;
    build_gtiles<R>(*P, s);
    P->bytes += P->gt.bytes;
  }
  int* d_int = nullptr;   // [ibi | flag | ibo], stream-ordered: freed on the same stream after its last use
  IPB_CUDA(cudaMallocAsync((void**)&d_int, (size_t)km * 3 * sizeof(int), s));
  for (i64 k0 = 0; k0 < km; k0 += IntPack::N) {   // ibi travels as kernel arguments: no host buffer outlives the call, no stream drain
    IntPack pk;
    const int n = (int)std::min<i64>(IntPack::N, km - k0);
    for (int i = 0; i < n; i++) pk.v[i] = ibi32[k0 + i];
    k_set_ints<<<nblk(n, 128), 128, 0, s>>>(pk, n, d_int + k0);
    g_launches++;
  }
  IPB_CUDA(cudaMemsetAsync(d_int + km, 0, (size_t)km * sizeof(int), s));
  FieldArgs<R> a{(int)h->no, (int)h->mi, (int)h->mo, (int)km, gi, vi, any_mask ? li : nullptr, d_int, go, vo, lo, d_int + km, (R)h->pmp, h->mcon, h->mspiral};
  launch_apply<R>(*P, h->vec, a, s);
  launch_post<R>(*P, h->vec, a, d_ibo ? d_ibo : d_int + 2 * km, s);
  IPB_CUDA(cudaFreeAsync(d_int, s));
  return 0;
}

// ---- gdswzd ----------------------------------------------------------------------------------------
template <class R>
__global__ void k_gdswzd(GridP<R> g, int iopt, int npts, int nm, R fill, R* x, R* y, R* rlon, R* rlat, int* nret, R* crot,
                         R* srot, R* xlon, R* xlat, R* ylon, R* ylat, R* area) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  bool ok = false;
  if (n < npts) {
    PtExt<R> e;
    if (iopt >= 0) {
      R xx, yy;
      if (iopt == 0) { xx = fill; yy = fill; if (n < nm) enumerate_xy(g, n + 1, xx, yy); x[n] = xx; y[n] = yy; }
      else { xx = x[n]; yy = y[n]; }
      R lo_, la_;
      ok = proj_fwd<R, EXT_ALL>(g, fill, xx, yy, lo_, la_, e);
      rlon[n] = lo_; rlat[n] = la_;
    } else {
      R xx, yy;
      ok = proj_inv<R, EXT_ALL>(g, fill, rlon[n], rlat[n], xx, yy, e);
      x[n] = xx; y[n] = yy;
    }
    if (crot) crot[n] = e.crot;
    if (srot) srot[n] = e.srot;
    if (xlon) xlon[n] = e.xlon;
    if (xlat) xlat[n] = e.xlat;
    if (ylon) ylon[n] = e.ylon;
    if (ylat) ylat[n] = e.ylat;
    if (area) area[n] = e.area;
  }
  const unsigned b = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(nret, __popc(b));
}

template <class R>
void run_gdswzd(GridHost<R> gh, i64 iopt, i64 npts, R fill, R* x, R* y, R* rlon, R* rlat, i64& nret, R* crot, R* srot, R* xlon,
                R* xlat, R* ylon, R* ylat, R* area) {
  Runtime& RT = rt();
  const int device = RT.current_device();
  DevState& DS = RT.state(device);
  std::lock_guard<std::mutex> lock(DS.mu);
  if (!gh.ok) { fprintf(stderr, "ipb200: gdswzd: unrecognised grid definition\n"); nret = 0; return; }
  if (npts <= 0) { nret = 0; return; }
  if (gh.p.type == P_STATION) { nret = npts; if (iopt == 0) for (i64 n = 0; n < npts; n++) { x[n] = fill; y[n] = fill; } return; }
  const i64 nm = (i64)gh.p.im * gh.p.jm;
  if (iopt == 0 && nm > npts) {  // src/gdswzd_mod.F90:137-143 (nret untouched)
    for (i64 n = 0; n < npts; n++) { rlat[n] = fill; rlon[n] = fill; x[n] = fill; y[n] = fill; }
    return;
  }
  cudaStream_t st = DS.stream[0];
  AsyncAlloc scope(st);
  R* tab = nullptr;
  upload_tables(gh, tab);
  R* out[7] = {crot, srot, xlon, xlat, ylon, ylat, area};
  int nopt = 0;
  for (int q = 0; q < 7; q++) if (out[q]) nopt++;
  R* d = dalloc<R>((size_t)npts * (4 + nopt));
  int* d_n = dalloc<int>(1);
  IPB_CUDA(cudaMemsetAsync(d_n, 0, sizeof(int), st));
  R *dx = d, *dy = d + npts, *dlo = d + 2 * npts, *dla = d + 3 * npts;
  if (iopt != 0 && iopt >= 0) {
    IPB_CUDA(cudaMemcpyAsync(dx, x, npts * sizeof(R), cudaMemcpyHostToDevice, st));
    IPB_CUDA(cudaMemcpyAsync(dy, y, npts * sizeof(R), cudaMemcpyHostToDevice, st));
  } else if (iopt < 0) {
    IPB_CUDA(cudaMemcpyAsync(dlo, rlon, npts * sizeof(R), cudaMemcpyHostToDevice, st));
    IPB_CUDA(cudaMemcpyAsync(dla, rlat, npts * sizeof(R), cudaMemcpyHostToDevice, st));
  }
  R* dopt[7];
  int c = 0;
  for (int q = 0; q < 7; q++) dopt[q] = out[q] ? d + (size_t)(4 + c++) * npts : nullptr;
  // the rotated grids' error branch (earth shape unknown) fills only one side and returns nret=0
  const bool rot_err = (gh.p.type == P_ROTB || gh.p.type == P_ROTE) && gh.p.rerth < 0;
  k_gdswzd<R><<<nblk(npts, 128), 128, 0, st>>>(gh.p, (int)iopt, (int)npts, (int)nm, fill, dx, dy, dlo, dla, d_n, dopt[0], dopt[1], dopt[2], dopt[3], dopt[4], dopt[5], dopt[6]);
  g_launches++;
  int hn = 0;
  IPB_CUDA(cudaMemcpyAsync(&hn, d_n, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (iopt >= 0) {
    if (iopt == 0) {
      IPB_CUDA(cudaMemcpyAsync(x, dx, npts * sizeof(R), cudaMemcpyDeviceToHost, st));
      IPB_CUDA(cudaMemcpyAsync(y, dy, npts * sizeof(R), cudaMemcpyDeviceToHost, st));
    }
    IPB_CUDA(cudaMemcpyAsync(rlon, dlo, npts * sizeof(R), cudaMemcpyDeviceToHost, st));
    IPB_CUDA(cudaMemcpyAsync(rlat, dla, npts * sizeof(R), cudaMemcpyDeviceToHost, st));
  }
  if (iopt < 0 || (rot_err && iopt == 0)) {
    IPB_CUDA(cudaMemcpyAsync(x, dx, npts * sizeof(R), cudaMemcpyDeviceToHost, st));
    IPB_CUDA(cudaMemcpyAsync(y, dy, npts * sizeof(R), cudaMemcpyDeviceToHost, st));
  }
  for (int q = 0; q < 7; q++)
    if (out[q]) IPB_CUDA(cudaMemcpyAsync(out[q], dopt[q], npts * sizeof(R), cudaMemcpyDeviceToHost, st));
  IPB_CUDA(cudaStreamSynchronize(st));
  if (rot_err && iopt == 0) for (i64 n = 0; n < npts; n++) { x[n] = fill; y[n] = fill; }
  nret = hn;
  dfree(d); dfree(d_n); dfree(tab);
}


}  // namespace ipb
