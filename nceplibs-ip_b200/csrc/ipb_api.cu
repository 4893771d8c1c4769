// ipb200 — the C ABI (contract: include/ipb200.h).
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
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ipb_decode.h"
#include "ipb_state.h"

namespace ipb {

std::atomic<long long> g_launches{0};
bool g_async_alloc = false;
cudaStream_t g_alloc_stream = nullptr, g_free_stream = nullptr;

#define IPB_CUDA_API(x)                                                                          \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess) {                                                                     \
      fprintf(stderr, "ipb200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      throw e_;                                                                                  \
    }                                                                                            \
  } while (0)

void* DBuf::get(size_t bytes) {
  if (bytes > cap) {
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    IPB_CUDA_API(cudaMalloc(&p, bytes));
    cap = bytes;
  }
  return p;
}
void DBuf::release() { if (p) cudaFree(p); p = nullptr; cap = 0; }

Runtime& rt() { static Runtime r; return r; }

void Runtime::ensure_init() {
  if (init) return;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    fprintf(stderr, "ipb200: no CUDA device available (%s); this library has no CPU path\n", cudaGetErrorString(e));
    abort();
  }
  for (int i = 0; i < 2; i++) IPB_CUDA_API(cudaStreamCreateWithFlags(&stream[i], cudaStreamNonBlocking));
  // keep freed plan memory in the pool (stream-ordered allocations of the plan builder, ipb_plan.cuh: dalloc)
  int devid = 0;
  cudaMemPool_t pool;
  if (cudaGetDevice(&devid) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, devid) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    g_free_stream = stream[0];
  }
  init = true;
}

// implemented once per real kind (ipb_kind_f32.cu / ipb_kind_f64.cu)
template <class R>
void run(bool vec, int dev, void* user_stream, i64 ip, const i64* ipopt, const GridHost<R>& gin, const GridHost<R>& gout, i64 mi,
         i64 mo, i64 km, const i64* ibi, const u8* li, const R* gi, const R* vi, i64& no, R* rlat, R* rlon, R* crot, R* srot,
         i64* ibo, u8* lo, R* go, R* vo, i64& iret);
template <class R>
void run_gdswzd(GridHost<R> gh, i64 iopt, i64 npts, R fill, R* x, R* y, R* rlon, R* rlat, i64& nret, R* crot, R* srot, R* xlon,
                R* xlat, R* ylon, R* ylat, R* area);
extern template void run<float>(bool, int, void*, i64, const i64*, const GridHost<float>&, const GridHost<float>&, i64, i64, i64,
                                const i64*, const u8*, const float*, const float*, i64&, float*, float*, float*, float*, i64*, u8*,
                                float*, float*, i64&);
extern template void run<double>(bool, int, void*, i64, const i64*, const GridHost<double>&, const GridHost<double>&, i64, i64, i64,
                                 const i64*, const u8*, const double*, const double*, i64&, double*, double*, double*, double*, i64*,
                                 u8*, double*, double*, i64&);
extern template void run_gdswzd<float>(GridHost<float>, i64, i64, float, float*, float*, float*, float*, i64&, float*, float*,
                                       float*, float*, float*, float*, float*);
extern template void run_gdswzd<double>(GridHost<double>, i64, i64, double, double*, double*, double*, double*, i64&, double*,
                                        double*, double*, double*, double*, double*, double*);

namespace {

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
  Runtime& RT = rt();
  std::lock_guard<std::mutex> lock(RT.mu);
  RT.cache.clear();
  for (int i = 0; i < 2; i++) { Work& w = RT.work[i]; w.gi.release(); w.vi.release(); w.li.release(); w.go.release(); w.vo.release(); w.lo.release(); w.ints.release(); }
}
long long ipgpu_plan_count() { Runtime& RT = rt(); std::lock_guard<std::mutex> lock(RT.mu); return (long long)RT.cache.size(); }
long long ipgpu_plan_bytes() { Runtime& RT = rt(); std::lock_guard<std::mutex> lock(RT.mu); size_t t = 0; for (auto& e : RT.cache) t += e.bytes; return (long long)t; }
void ipgpu_set_plan_cache_bytes(long long b) { Runtime& RT = rt(); std::lock_guard<std::mutex> lock(RT.mu); RT.cache_cap = (size_t)b; }
void ipgpu_set_chunk_bytes(long long b) { Runtime& RT = rt(); std::lock_guard<std::mutex> lock(RT.mu); RT.chunk_bytes = (size_t)std::max<long long>(b, 1 << 20); }
long long ipgpu_launch_count() { return g_launches.load(); }
// timings of the last ipolates/ipolatev call on this process, CUDA-event milliseconds
void ipgpu_last_timing(double* plan_ms, double* apply_ms, int* plan_cache_hit) {
  Runtime& RT = rt();
  if (plan_ms) *plan_ms = RT.last_plan_ms;
  if (apply_ms) *apply_ms = RT.last_apply_ms;
  if (plan_cache_hit) *plan_cache_hit = RT.last_plan_hit;
}
void ipgpu_last_transfer(long long* h2d_bytes, long long* d2h_bytes) {
  Runtime& RT = rt();
  if (h2d_bytes) *h2d_bytes = RT.last_h2d;
  if (d2h_bytes) *d2h_bytes = RT.last_d2h;
}
void* ipgpu_host_alloc(long long bytes) { void* p = nullptr; if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void ipgpu_host_free(void* p) { if (p) cudaFreeHost(p); }
const char* ipgpu_version() { return "ipb200 0.1 (sm_100a)"; }
}
