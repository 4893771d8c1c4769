// ipb200 — host orchestration and the C ABI (see include/ipb200.h for the contract).
//
// Entry points mirror the reference's bind(c) drivers one-to-one, with the build kind as a
// suffix instead of a separate library:
//   ipolates_grib2_<k> ...   src/ipolates.F90:587-636      ipolatev_grib2_<k> ...  src/ipolatev.F90:382-434
//   ipolates_grib1_<k> ...   src/ipolates.F90:293-334      ipolatev_grib1_<k> ...  src/ipolatev.F90:565-628
//   *_single_field_<k>       src/ipolates.F90:158,808       src/ipolatev.F90:680,832
//   gdswzd_<k>, gdswzd_grib1_<k>   src/gdswzd_c.F90:173,259
// <k> = 4 (int32/float), d (int32/double), 8 (int64/double).
// There is no CPU fallback: every call needs a CUDA device and fails loudly without one.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>

#include "ipb_apply.cuh"
#include "ipb_fast.cuh"

namespace ipb {

std::atomic<long long> g_launches{0};

namespace {

std::mutex g_mu;
cudaStream_t g_stream[2] = {nullptr, nullptr};
cudaEvent_t g_ev[2] = {nullptr, nullptr};
bool g_init = false;
double g_last_plan_ms = 0, g_last_apply_ms = 0;
int g_last_plan_hit = 0;
size_t g_chunk_bytes = (size_t)768 << 20;

void ensure_init() {
  if (g_init) return;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fprintf(stderr, "ipb200: no CUDA device available (%s); this library has no CPU path\n", cudaGetErrorString(e));
    abort();
  }
  for (int i = 0; i < 2; i++) {
    IPB_CUDA(cudaStreamCreateWithFlags(&g_stream[i], cudaStreamNonBlocking));
    IPB_CUDA(cudaEventCreateWithFlags(&g_ev[i], cudaEventDisableTiming));
  }
  g_init = true;
}

struct DBuf {  // growable device scratch
  void* p = nullptr; size_t cap = 0;
  void* get(size_t bytes) {
    if (bytes > cap) {
      if (p) cudaFree(p);
      IPB_CUDA(cudaMalloc(&p, bytes));
      cap = bytes;
    }
    return p;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct Work { DBuf gi, vi, li, go, vo, lo, ints; };
Work g_work[2];

// ---- plan cache --------------------------------------------------------------------------------
struct CacheEntry { std::vector<i64> key; std::shared_ptr<void> plan; int kind; size_t bytes; unsigned long long stamp; };
std::vector<CacheEntry> g_cache;
unsigned long long g_stamp = 0;
size_t g_cache_cap = (size_t)24 << 30;  // bytes of plans kept (180 GB HBM: generous by default)

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

template <class R> std::shared_ptr<Plan<R>> cache_find(const std::vector<i64>& key) {
  for (auto& e : g_cache)
    if (e.kind == (int)sizeof(R) && e.key == key) { e.stamp = ++g_stamp; return std::static_pointer_cast<Plan<R>>(e.plan); }
  return nullptr;
}
template <class R> void cache_put(const std::vector<i64>& key, std::shared_ptr<Plan<R>> p) {
  size_t tot = p->bytes;
  for (auto& e : g_cache) tot += e.bytes;
  while (tot > g_cache_cap && !g_cache.empty()) {
    auto it = std::min_element(g_cache.begin(), g_cache.end(), [](const CacheEntry& a, const CacheEntry& b) { return a.stamp < b.stamp; });
    tot -= it->bytes;
    g_cache.erase(it);
  }
  g_cache.push_back({key, std::static_pointer_cast<void>(p), (int)sizeof(R), p->bytes, ++g_stamp});
}

bool is_device_or_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// ---- one launch of the apply stage over kc fields already resident on the device ------------------
template <class R>
void launch_apply(const Plan<R>& P, bool vec, FieldArgs<R> a, cudaStream_t st) {
  const int no = P.no;
  if (no <= 0 || a.kc <= 0) return;
  const GridP<R>& gi = P.gin.p;
  const int method = P.method;
  if (method == M_BILINEAR || method == M_BICUBIC) {
    if (method == M_BILINEAR && !vec && fast_bilinear_ok(P, a)) { launch_fast_bilinear(P, a, st); return; }
    const int kpb = std::max(1, std::min(a.kc, 16));
    dim3 grid(nblk(no, 128), nblk(a.kc, kpb));
    if (method == M_BILINEAR) {
      if (vec) k_apply_stencil<R, 2, true><<<grid, 128, 0, st>>>(a, gi, kpb, P.nxy, P.wxy, P.cxy, P.sxy, P.nc, P.x, P.y, P.crot, P.srot);
      else k_apply_stencil<R, 2, false><<<grid, 128, 0, st>>>(a, gi, kpb, P.nxy, P.wxy, nullptr, nullptr, P.nc, P.x, P.y, nullptr, nullptr);
    } else {
      if (vec) k_apply_stencil<R, 4, true><<<grid, 128, 0, st>>>(a, gi, kpb, P.nxy, P.wxy, P.cxy, P.sxy, P.nc, P.x, P.y, P.crot, P.srot);
      else k_apply_stencil<R, 4, false><<<grid, 128, 0, st>>>(a, gi, kpb, P.nxy, P.wxy, nullptr, nullptr, P.nc, P.x, P.y, nullptr, nullptr);
    }
    g_launches++;
  } else if (method == M_NEIGHBOR) {
    const int kpb = std::max(1, std::min(a.kc, 32));
    dim3 grid(nblk(no, 256), nblk(a.kc, kpb));
    if (vec) k_apply_neighbor<R, true><<<grid, 256, 0, st>>>(a, gi, kpb, P.nxy, P.cxy, P.sxy, P.x, P.y, P.rlat, P.rlon, P.crot, P.srot);
    else k_apply_neighbor<R, false><<<grid, 256, 0, st>>>(a, gi, kpb, P.nxy, nullptr, nullptr, P.x, P.y, P.rlat, P.rlon, nullptr, nullptr);
    g_launches++;
  } else {
    constexpr int KF = 8;
    dim3 grid(nblk(no, 128), nblk(a.kc, KF));
    const int ns = P.bo.nsamp;
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
__global__ void k_ibo(int kc, const int* __restrict__ ibi, const int* __restrict__ flag, int* ibo) {
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

// ---- the driver shared by ipolates / ipolatev ------------------------------------------------------
// dev = 1: li/gi/vi/lo/go/vo are device pointers and everything runs on `user_stream`.
template <class R>
void run(bool vec, int dev, void* user_stream, i64 ip, const i64* ipopt, const GridHost<R>& gin, const GridHost<R>& gout, i64 mi,
         i64 mo, i64 km, const i64* ibi, const u8* li, const R* gi, const R* vi, i64& no, R* rlat, R* rlon, R* crot, R* srot,
         i64* ibo, u8* lo, R* go, R* vo, i64& iret) {
  std::lock_guard<std::mutex> lock(g_mu);
  ensure_init();
  iret = 0;
  if (!(ip == 0 || ip == 1 || ip == 2 || ip == 3 || ip == 6)) {
    // ip=4 (spectral) is outside this library's scope; other values make the reference error-stop.
    fprintf(stderr, "ipb200: interpolation method ip=%lld is not supported (0,1,2,3,6 are)\n", (long long)ip);
    iret = 1; return;
  }
  if (!gin.ok || !gout.ok) { fprintf(stderr, "ipb200: unrecognised grid definition\n"); iret = 1; return; }
  if (mi >= ((i64)1 << 31) || mo >= ((i64)1 << 31) || km < 0) { iret = 1; return; }
  const int method = (int)ip;
  const bool budget = method == M_BUDGET || method == M_NBUDGET;
  const bool station = gout.p.type == P_STATION;
  cudaStream_t st0 = dev ? (cudaStream_t)user_stream : g_stream[0];

  // option decode of the stencil methods (bilinear_interp_mod.F90:110-115, bicubic :117-122, neighbor :135)
  i64 mp = 50, mcon = 0, mspiral = 0;
  if (method == M_BILINEAR) { mp = ipopt[0]; mspiral = std::max<i64>(ipopt[1], 0); }
  else if (method == M_BICUBIC) { mcon = ipopt[0]; mp = ipopt[1]; }
  else if (method == M_NEIGHBOR) { mspiral = std::max<i64>(ipopt[0], 1); }
  else mspiral = std::max<i64>(ipopt[19], 1);
  if (method == M_BILINEAR || method == M_BICUBIC) {
    if (mp == -1 || mp == 0) mp = 50;
    if (mp < 0 || mp > 100) iret = 32;
  }
  R pmp = (R)mp * IPB_RC(0.01);
  if (!budget && iret != 0) { if (!station) no = 0; return; }

  // plan: cached per grid pair (never for station points, whose coordinates are call arguments)
  std::shared_ptr<Plan<R>> P;
  std::vector<i64> key = plan_key<R>(method, vec, ipopt, gin, gout, mi, mo);
  cudaEvent_t t0, t1, t2;
  cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventCreate(&t2);
  cudaEventRecord(t0, st0);
  g_last_plan_hit = 0;
  if (!station) { P = cache_find<R>(key); if (P) g_last_plan_hit = 1; }
  if (!P) {
    P.reset(build_plan<R>(method, vec, ipopt, gin, gout, mi, mo, no, rlat, rlon, crot, srot, st0));
    if (!station) cache_put<R>(key, P);
  }
  cudaEventRecord(t1, st0);
  const int NO = P->no;
  // lat/lon (and rotations) are outputs for real grids; for stations the budget-family vector
  // routines overwrite crot/srot (see k_out_geo)
  auto copy_geo = [&](cudaStream_t s) {
    if (mo <= 0) return;
    if (!station) {
      if (rlat) IPB_CUDA(cudaMemcpyAsync(rlat, P->rlat, mo * sizeof(R), cudaMemcpyDeviceToHost, s));
      if (rlon) IPB_CUDA(cudaMemcpyAsync(rlon, P->rlon, mo * sizeof(R), cudaMemcpyDeviceToHost, s));
    }
    if (vec && (!station || budget)) {
      if (crot) IPB_CUDA(cudaMemcpyAsync(crot, P->crot, mo * sizeof(R), cudaMemcpyDeviceToHost, s));
      if (srot) IPB_CUDA(cudaMemcpyAsync(srot, P->srot, mo * sizeof(R), cudaMemcpyDeviceToHost, s));
    }
  };
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
    if (!station) no = 0;
    cudaEventDestroy(t0); cudaEventDestroy(t1); cudaEventDestroy(t2);
    return;
  } else {
    no = station ? mo : NO;
  }
  copy_geo(st0);

  bool any_mask = false;
  for (i64 k = 0; k < km; k++) if (ibi[k] != 0) any_mask = true;
  std::vector<int> ibi32(std::max<i64>(km, 1));
  for (i64 k = 0; k < km; k++) ibi32[k] = ibi[k] != 0 ? (int)ibi[k] : 0;
  std::vector<int> ibo32(std::max<i64>(km, 1), 0);

  if (NO > 0 && km > 0) {
    if (dev) {
      // whole batch in one pass on the caller's stream
      int* d_int = (int*)g_work[0].ints.get((size_t)km * 3 * sizeof(int));
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
      i64 kc = std::max<i64>(1, std::min<i64>(km, (i64)(g_chunk_bytes / std::max<size_t>(per_field, 1))));
      if (km > 1 && kc >= km) kc = (km + 1) / 2;  // at least two chunks so copies and kernels overlap
      IPB_CUDA(cudaStreamSynchronize(st0));        // plan ready before the second stream uses it
      int ci = 0;
      for (i64 k0 = 0; k0 < km; k0 += kc, ci ^= 1) {
        const i64 kn = std::min(kc, km - k0);
        cudaStream_t s = g_stream[ci];
        Work& w = g_work[ci];
        IPB_CUDA(cudaStreamSynchronize(s));  // buffers of this slot are free again
        R* d_gi = (R*)w.gi.get((size_t)kn * mi * sizeof(R));
        R* d_vi = vec ? (R*)w.vi.get((size_t)kn * mi * sizeof(R)) : nullptr;
        u8* d_li = any_mask ? (u8*)w.li.get((size_t)kn * mi) : nullptr;
        R* d_go = (R*)w.go.get((size_t)kn * mo * sizeof(R));
        R* d_vo = vec ? (R*)w.vo.get((size_t)kn * mo * sizeof(R)) : nullptr;
        u8* d_lo = (u8*)w.lo.get((size_t)kn * mo);
        int* d_int = (int*)w.ints.get((size_t)kn * 3 * sizeof(int));
        IPB_CUDA(cudaMemcpyAsync(d_gi, gi + (size_t)k0 * mi, (size_t)kn * mi * sizeof(R), cudaMemcpyHostToDevice, s));
        if (vec) IPB_CUDA(cudaMemcpyAsync(d_vi, vi + (size_t)k0 * mi, (size_t)kn * mi * sizeof(R), cudaMemcpyHostToDevice, s));
        if (any_mask) IPB_CUDA(cudaMemcpyAsync(d_li, li + (size_t)k0 * mi, (size_t)kn * mi, cudaMemcpyHostToDevice, s));
        IPB_CUDA(cudaMemcpyAsync(d_int, ibi32.data() + k0, kn * sizeof(int), cudaMemcpyHostToDevice, s));
        IPB_CUDA(cudaMemsetAsync(d_int + kn, 0, kn * sizeof(int), s));
        FieldArgs<R> a{NO, (int)mi, (int)mo, (int)kn, d_gi, d_vi, d_li, d_int, d_go, d_vo, d_lo, d_int + kn, pmp, (int)mcon, (int)mspiral};
        launch_apply<R>(*P, vec, a, s);
        launch_post<R>(*P, vec, a, d_int + 2 * kn, s);
        // only the first NO points of each field are defined (the reference leaves the rest untouched)
        if (NO == mo) {
          IPB_CUDA(cudaMemcpyAsync(go + (size_t)k0 * mo, d_go, (size_t)kn * mo * sizeof(R), cudaMemcpyDeviceToHost, s));
          if (vec) IPB_CUDA(cudaMemcpyAsync(vo + (size_t)k0 * mo, d_vo, (size_t)kn * mo * sizeof(R), cudaMemcpyDeviceToHost, s));
          IPB_CUDA(cudaMemcpyAsync(lo + (size_t)k0 * mo, d_lo, (size_t)kn * mo, cudaMemcpyDeviceToHost, s));
        } else {
          IPB_CUDA(cudaMemcpy2DAsync(go + (size_t)k0 * mo, mo * sizeof(R), d_go, mo * sizeof(R), NO * sizeof(R), kn, cudaMemcpyDeviceToHost, s));
          if (vec) IPB_CUDA(cudaMemcpy2DAsync(vo + (size_t)k0 * mo, mo * sizeof(R), d_vo, mo * sizeof(R), NO * sizeof(R), kn, cudaMemcpyDeviceToHost, s));
          IPB_CUDA(cudaMemcpy2DAsync(lo + (size_t)k0 * mo, mo, d_lo, mo, NO, kn, cudaMemcpyDeviceToHost, s));
        }
        IPB_CUDA(cudaMemcpyAsync(ibo32.data() + k0, d_int + 2 * kn, kn * sizeof(int), cudaMemcpyDeviceToHost, s));
      }
      IPB_CUDA(cudaStreamSynchronize(g_stream[0]));
      IPB_CUDA(cudaStreamSynchronize(g_stream[1]));
    }
    for (i64 k = 0; k < km; k++) ibo[k] = ibo32[k];
  } else {
    IPB_CUDA(cudaStreamSynchronize(st0));
    for (i64 k = 0; k < km; k++) ibo[k] = ibi[k];
  }
  cudaEventRecord(t2, st0);
  cudaEventSynchronize(t2);
  float ms = 0;
  cudaEventElapsedTime(&ms, t0, t1); g_last_plan_ms = ms;
  cudaEventElapsedTime(&ms, t1, t2); g_last_apply_ms = ms;
  cudaEventDestroy(t0); cudaEventDestroy(t1); cudaEventDestroy(t2);
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
  std::lock_guard<std::mutex> lock(g_mu);
  ensure_init();
  if (!gh.ok) { fprintf(stderr, "ipb200: gdswzd: unrecognised grid definition\n"); nret = 0; return; }
  if (npts <= 0) { nret = 0; return; }
  if (gh.p.type == P_STATION) { nret = npts; if (iopt == 0) for (i64 n = 0; n < npts; n++) { x[n] = fill; y[n] = fill; } return; }
  const i64 nm = (i64)gh.p.im * gh.p.jm;
  if (iopt == 0 && nm > npts) {  // src/gdswzd_mod.F90:137-143 (nret untouched)
    for (i64 n = 0; n < npts; n++) { rlat[n] = fill; rlon[n] = fill; x[n] = fill; y[n] = fill; }
    return;
  }
  cudaStream_t st = g_stream[0];
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

template <class I> std::vector<i64> widen(const I* p, i64 n) {
  std::vector<i64> v((size_t)std::max<i64>(n, 0));
  for (i64 i = 0; i < n; i++) v[i] = (i64)p[i];
  return v;
}

template <class R, class I>
void api_interp(bool vec, bool grib2, int dev, void* stream, const I* ip, const I* ipopt, const I* numi, const I* dsci, const I* leni,
                const I* numo, const I* dsco, const I* leno, const I* mi, const I* mo, i64 km, const I* ibi, const u8* li,
                const R* gi, const R* vi, I* no, R* rlat, R* rlon, R* crot, R* srot, I* ibo, u8* lo, R* go, R* vo, I* iret) {
  try {
    GridHost<R> gin, gout;
    if (grib2) {
      auto ti = widen(dsci, *leni), to = widen(dsco, *leno);
      gin = decode_grib2<R>(*numi, ti.data(), *leni);
      gout = decode_grib2<R>(*numo, to.data(), *leno);
    } else {
      auto ti = widen(dsci, 200), to = widen(dsco, 200);
      if (vec) {  // src/ipolatev.F90:602-609: E-grids are wind points for the duration of the call
        if (ti[0] == 203) ti[10] |= 256;
        if (to[0] == 203) to[10] |= 256;
      }
      gin = decode_grib1<R>(ti.data());
      gout = decode_grib1<R>(to.data());
    }
    auto opt = widen(ipopt, 20), ib = widen(ibi, km);
    std::vector<i64> ibo64((size_t)std::max<i64>(km, 1), 0);
    i64 no64 = *no, iret64 = 0;
    run<R>(vec, dev, stream, *ip, opt.data(), gin, gout, *mi, *mo, km, ib.data(), li, gi, vi, no64, rlat, rlon, crot, srot,
           ibo64.data(), lo, go, vo, iret64);
    *no = (I)no64; *iret = (I)iret64;
    const bool budget = (*ip == 3 || *ip == 6);
    if ((iret64 == 0 || budget) && iret64 != 1) for (i64 k = 0; k < km; k++) ibo[k] = (I)ibo64[k];
  } catch (cudaError_t e) {
    *iret = (I)(1000 + (int)e);  // CUDA failures surface as iret >= 1000 (never produced by the reference)
  }
}

}  // namespace
}  // namespace ipb

using namespace ipb;

#define IPB_KIND(SFX, R, I)                                                                                              \
  extern "C" {                                                                                                           \
  void ipolates_grib2_##SFX(const I* ip, const I* ipopt, const I* numi, const I* tmpli, const I* leni, const I* numo,    \
                            const I* tmplo, const I* leno, const I* mi, const I* mo, const I* km, const I* ibi,          \
                            const u8* li, const R* gi, I* no, R* rlat, R* rlon, I* ibo, u8* lo, R* go, I* iret) {        \
    api_interp<R, I>(false, true, 0, nullptr, ip, ipopt, numi, tmpli, leni, numo, tmplo, leno, mi, mo, *km, ibi, li, gi, \
                     nullptr, no, rlat, rlon, nullptr, nullptr, ibo, lo, go, nullptr, iret);                             \
  }                                                                                                                      \
  void ipolates_grib2_single_field_##SFX(const I* ip, const I* ipopt, const I* numi, const I* tmpli, const I* leni,      \
                                         const I* numo, const I* tmplo, const I* leno, const I* mi, const I* mo,         \
                                         const I* km, const I* ibi, const u8* li, const R* gi, I* no, R* rlat, R* rlon,  \
                                         I* ibo, u8* lo, R* go, I* iret) {                                               \
    api_interp<R, I>(false, true, 0, nullptr, ip, ipopt, numi, tmpli, leni, numo, tmplo, leno, mi, mo, 1, ibi, li, gi,   \
                     nullptr, no, rlat, rlon, nullptr, nullptr, ibo, lo, go, nullptr, iret);                             \
  }                                                                                                                      \
  void ipolates_grib1_##SFX(const I* ip, const I* ipopt, const I* kgdsi, const I* kgdso, const I* mi, const I* mo,       \
                            const I* km, const I* ibi, const u8* li, const R* gi, I* no, R* rlat, R* rlon, I* ibo,       \
                            u8* lo, R* go, I* iret) {                                                                    \
    api_interp<R, I>(false, false, 0, nullptr, ip, ipopt, nullptr, kgdsi, nullptr, nullptr, kgdso, nullptr, mi, mo, *km, \
                     ibi, li, gi, nullptr, no, rlat, rlon, nullptr, nullptr, ibo, lo, go, nullptr, iret);                \
  }                                                                                                                      \
  void ipolates_grib1_single_field_##SFX(const I* ip, const I* ipopt, const I* kgdsi, const I* kgdso, const I* mi,       \
                                         const I* mo, const I* km, const I* ibi, const u8* li, const R* gi, I* no,       \
                                         R* rlat, R* rlon, I* ibo, u8* lo, R* go, I* iret) {                             \
    api_interp<R, I>(false, false, 0, nullptr, ip, ipopt, nullptr, kgdsi, nullptr, nullptr, kgdso, nullptr, mi, mo, 1,   \
                     ibi, li, gi, nullptr, no, rlat, rlon, nullptr, nullptr, ibo, lo, go, nullptr, iret);                \
  }                                                                                                                      \
  void ipolatev_grib2_##SFX(const I* ip, const I* ipopt, const I* numi, const I* tmpli, const I* leni, const I* numo,    \
                            const I* tmplo, const I* leno, const I* mi, const I* mo, const I* km, const I* ibi,          \
                            const u8* li, const R* ui, const R* vi, I* no, R* rlat, R* rlon, R* crot, R* srot, I* ibo,   \
                            u8* lo, R* uo, R* vo, I* iret) {                                                             \
    api_interp<R, I>(true, true, 0, nullptr, ip, ipopt, numi, tmpli, leni, numo, tmplo, leno, mi, mo, *km, ibi, li, ui,  \
                     vi, no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret);                                             \
  }                                                                                                                      \
  void ipolatev_grib2_single_field_##SFX(const I* ip, const I* ipopt, const I* numi, const I* tmpli, const I* leni,      \
                                         const I* numo, const I* tmplo, const I* leno, const I* mi, const I* mo,         \
                                         const I* km, const I* ibi, const u8* li, const R* ui, const R* vi, I* no,       \
                                         R* rlat, R* rlon, R* crot, R* srot, I* ibo, u8* lo, R* uo, R* vo, I* iret) {    \
    api_interp<R, I>(true, true, 0, nullptr, ip, ipopt, numi, tmpli, leni, numo, tmplo, leno, mi, mo, 1, ibi, li, ui,    \
                     vi, no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret);                                             \
  }                                                                                                                      \
  void ipolatev_grib1_##SFX(const I* ip, const I* ipopt, const I* kgdsi, const I* kgdso, const I* mi, const I* mo,       \
                            const I* km, const I* ibi, const u8* li, const R* ui, const R* vi, I* no, R* rlat, R* rlon,  \
                            R* crot, R* srot, I* ibo, u8* lo, R* uo, R* vo, I* iret) {                                   \
    api_interp<R, I>(true, false, 0, nullptr, ip, ipopt, nullptr, kgdsi, nullptr, nullptr, kgdso, nullptr, mi, mo, *km,  \
                     ibi, li, ui, vi, no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret);                                \
  }                                                                                                                      \
  void ipolatev_grib1_single_field_##SFX(const I* ip, const I* ipopt, const I* kgdsi, const I* kgdso, const I* mi,       \
                                         const I* mo, const I* km, const I* ibi, const u8* li, const R* ui, const R* vi, \
                                         I* no, R* rlat, R* rlon, R* crot, R* srot, I* ibo, u8* lo, R* uo, R* vo,        \
                                         I* iret) {                                                                      \
    api_interp<R, I>(true, false, 0, nullptr, ip, ipopt, nullptr, kgdsi, nullptr, nullptr, kgdso, nullptr, mi, mo, 1,    \
                     ibi, li, ui, vi, no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret);                                \
  }                                                                                                                      \
  /* device-resident variants: li/gi(/vi)/lo/go(/vo) are device pointers, work is queued on `stream` */                 \
  void ipgpu_ipolates_grib2_dev_##SFX(void* stream, const I* ip, const I* ipopt, const I* numi, const I* tmpli,          \
                                      const I* leni, const I* numo, const I* tmplo, const I* leno, const I* mi,          \
                                      const I* mo, const I* km, const I* ibi, const u8* li, const R* gi, I* no, R* rlat, \
                                      R* rlon, I* ibo, u8* lo, R* go, I* iret) {                                         \
    api_interp<R, I>(false, true, 1, stream, ip, ipopt, numi, tmpli, leni, numo, tmplo, leno, mi, mo, *km, ibi, li, gi,  \
                     nullptr, no, rlat, rlon, nullptr, nullptr, ibo, lo, go, nullptr, iret);                             \
  }                                                                                                                      \
  void ipgpu_ipolatev_grib2_dev_##SFX(void* stream, const I* ip, const I* ipopt, const I* numi, const I* tmpli,          \
                                      const I* leni, const I* numo, const I* tmplo, const I* leno, const I* mi,          \
                                      const I* mo, const I* km, const I* ibi, const u8* li, const R* ui, const R* vi,    \
                                      I* no, R* rlat, R* rlon, R* crot, R* srot, I* ibo, u8* lo, R* uo, R* vo,           \
                                      I* iret) {                                                                         \
    api_interp<R, I>(true, true, 1, stream, ip, ipopt, numi, tmpli, leni, numo, tmplo, leno, mi, mo, *km, ibi, li, ui,   \
                     vi, no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret);                                             \
  }                                                                                                                      \
  void gdswzd_##SFX(I num, const I* tmpl, I len, I iopt, I npts, R fill, R* xpts, R* ypts, R* rlon, R* rlat, I* nret,    \
                    R* crot, R* srot, R* xlon, R* xlat, R* ylon, R* ylat, R* area) {                                     \
    try {                                                                                                                \
      auto t = widen(tmpl, len);                                                                                         \
      i64 nr = *nret;                                                                                                    \
      run_gdswzd<R>(decode_grib2<R>(num, t.data(), len), iopt, npts, fill, xpts, ypts, rlon, rlat, nr, crot, srot, xlon, \
                    xlat, ylon, ylat, area);                                                                             \
      *nret = (I)nr;                                                                                                     \
    } catch (cudaError_t) { *nret = 0; }                                                                                 \
  }                                                                                                                      \
  void gdswzd_grib1_##SFX(const I* kgds, I iopt, I npts, R fill, R* xpts, R* ypts, R* rlon, R* rlat, I* nret, R* crot,   \
                          R* srot, R* xlon, R* xlat, R* ylon, R* ylat, R* area) {                                        \
    try {                                                                                                                \
      auto t = widen(kgds, 200);                                                                                         \
      i64 nr = *nret;                                                                                                    \
      run_gdswzd<R>(decode_grib1<R>(t.data()), iopt, npts, fill, xpts, ypts, rlon, rlat, nr, crot, srot, xlon, xlat,     \
                    ylon, ylat, area);                                                                                   \
      *nret = (I)nr;                                                                                                     \
    } catch (cudaError_t) { *nret = 0; }                                                                                 \
  }                                                                                                                      \
  }

IPB_KIND(4, float, int32_t)
IPB_KIND(d, double, int32_t)
IPB_KIND(8, double, int64_t)

extern "C" {
// library management (new; no reference counterpart)
int ipgpu_device_count() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
int ipgpu_set_device(int d) { return (int)cudaSetDevice(d); }
void ipgpu_plan_clear() {
  std::lock_guard<std::mutex> lock(g_mu);
  g_cache.clear();
  for (int i = 0; i < 2; i++) { Work& w = g_work[i]; w.gi.release(); w.vi.release(); w.li.release(); w.go.release(); w.vo.release(); w.lo.release(); w.ints.release(); }
}
long long ipgpu_plan_count() { std::lock_guard<std::mutex> lock(g_mu); return (long long)g_cache.size(); }
long long ipgpu_plan_bytes() { std::lock_guard<std::mutex> lock(g_mu); size_t t = 0; for (auto& e : g_cache) t += e.bytes; return (long long)t; }
void ipgpu_set_plan_cache_bytes(long long b) { std::lock_guard<std::mutex> lock(g_mu); g_cache_cap = (size_t)b; }
void ipgpu_set_chunk_bytes(long long b) { std::lock_guard<std::mutex> lock(g_mu); g_chunk_bytes = (size_t)std::max<long long>(b, 1 << 20); }
long long ipgpu_launch_count() { return g_launches.load(); }
// timings of the last ipolates/ipolatev call on this process, CUDA-event milliseconds
void ipgpu_last_timing(double* plan_ms, double* apply_ms, int* plan_cache_hit) {
  if (plan_ms) *plan_ms = g_last_plan_ms;
  if (apply_ms) *apply_ms = g_last_apply_ms;
  if (plan_cache_hit) *plan_cache_hit = g_last_plan_hit;
}
void* ipgpu_host_alloc(long long bytes) { void* p = nullptr; if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void ipgpu_host_free(void* p) { if (p) cudaFreeHost(p); }
const char* ipgpu_version() { return "ipb200 0.1 (sm_100a)"; }
}
