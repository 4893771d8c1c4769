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
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstring>
#include <mutex>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <vector>

#include "ipb_decode.h"
#include "ipb_state.h"
#include "ipb_hostcopy.h"

namespace ipb {

std::atomic<long long> g_launches{0};

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
void* HBuf::get(size_t bytes) {
  if (bytes > cap) {
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
    IPB_CUDA_API(cudaHostAlloc(&p, bytes, cudaHostAllocPortable));
    cap = bytes;
  }
  return p;
}
void HBuf::release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }

Runtime& rt() { static Runtime r; return r; }
AllocCtx& alloc_ctx() { static thread_local AllocCtx c; return c; }

int Runtime::current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess) {
    fprintf(stderr, "ipb200: no CUDA device available (%s); this library has no CPU path\n", cudaGetErrorString(cudaGetLastError()));
    abort();
  }
  return d;
}

DevState& Runtime::state(int device) {
  std::lock_guard<std::mutex> lock(mu);
  if (dev.empty()) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
      fprintf(stderr, "ipb200: no CUDA device available (%s); this library has no CPU path\n", cudaGetErrorString(e));
      abort();
    }
    dev.resize((size_t)ndev);
  }
  if (device < 0 || device >= (int)dev.size()) { fprintf(stderr, "ipb200: device %d does not exist\n", device); abort(); }
  if (!dev[device]) {
    DeviceGuard guard(device);
    std::unique_ptr<DevState> d(new DevState());
    d->device = device;
    for (int i = 0; i < 2; i++) IPB_CUDA_API(cudaStreamCreateWithFlags(&d->stream[i], cudaStreamNonBlocking));
    for (int i = 0; i < 3; i++) IPB_CUDA_API(cudaEventCreate(&d->ev[i]));
    // keep freed plan memory in the pool (stream-ordered allocations of the plan builder, ipb_plan.cuh: dalloc)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    dev[device] = std::move(d);
  }
  return *dev[device];
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

template <class R> PlanHandle* plan_get(bool, i64, const i64*, const GridHost<R>&, const GridHost<R>&, i64, i64, i64&, i64&);
template <class R> int plan_geometry(PlanHandle*, void*, R*, R*, R*, R*);
template <class R> int plan_apply(PlanHandle*, void*, i64, const i64*, const u8*, const R*, const R*, int*, u8*, R*, R*);
extern template PlanHandle* plan_get<float>(bool, i64, const i64*, const GridHost<float>&, const GridHost<float>&, i64, i64, i64&, i64&);
extern template PlanHandle* plan_get<double>(bool, i64, const i64*, const GridHost<double>&, const GridHost<double>&, i64, i64, i64&, i64&);
extern template int plan_geometry<float>(PlanHandle*, void*, float*, float*, float*, float*);
extern template int plan_geometry<double>(PlanHandle*, void*, double*, double*, double*, double*);
extern template int plan_apply<float>(PlanHandle*, void*, i64, const i64*, const u8*, const float*, const float*, int*, u8*, float*, float*);
extern template int plan_apply<double>(PlanHandle*, void*, i64, const i64*, const u8*, const double*, const double*, int*, u8*, double*, double*);

template <class R> void run_xwafs(int, bool, i64, const i64*, i64, i64, i64, i64, const i64*, const R*, const u8*, R*, u8*, i64*);
extern template void run_xwafs<float>(int, bool, i64, const i64*, i64, i64, i64, i64, const i64*, const float*, const u8*, float*, u8*, i64*);
extern template void run_xwafs<double>(int, bool, i64, const i64*, i64, i64, i64, i64, const i64*, const double*, const u8*, double*, u8*, i64*);

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
  } catch (const std::exception& e) {   // host allocation failures etc.: never let a C++ exception cross the C / Fortran frames
    fprintf(stderr, "ipb200: %s\n", e.what());
    *iret = (I)1001;
  } catch (...) {
    *iret = (I)1001;
  }
}

// shared by the handle entry points
template <class R, class I>
PlanHandle* api_plan_get(bool vec, const I* ip, const I* ipopt, const I* numi, const I* dsci, const I* leni, const I* numo, const I* dsco,
                         const I* leno, const I* mi, const I* mo, I* no, I* iret) {
  try {
    auto ti = widen(dsci, *leni), to = widen(dsco, *leno);
    GridHost<R> gin = decode_grib2<R>(*numi, ti.data(), *leni), gout = decode_grib2<R>(*numo, to.data(), *leno);
    auto opt = widen(ipopt, 20);
    i64 no64 = 0, iret64 = 0;
    PlanHandle* h = plan_get<R>(vec, *ip, opt.data(), gin, gout, *mi, *mo, no64, iret64);
    *no = (I)no64; *iret = (I)iret64;
    return h;
  } catch (cudaError_t e) { *iret = (I)(1000 + (int)e); } catch (...) { *iret = (I)1001; }
  return nullptr;
}
template <class R, class I>
void api_plan_apply(PlanHandle* h, void* stream, i64 km, const I* ibi, const u8* li, const R* gi, const R* vi, int32_t* d_ibo, u8* lo, R* go,
                    R* vo, I* iret) {
  try {
    auto ib = widen(ibi, km);
    const int rc = plan_apply<R>(h, stream, km, ib.data(), li, gi, vi, d_ibo, lo, go, vo);
    *iret = rc ? (I)1 : (I)(h ? h->iret : 1);
  } catch (cudaError_t e) { *iret = (I)(1000 + (int)e); } catch (...) { *iret = (I)1001; }
}


// ---- grid expanders ---------------------------------------------------------------------
// rows of the thinned WAFS grid from the equator to the pole (src/ipxwafs.F90:120-126)
const int kWafsRow[73] = {73, 73, 73, 73, 73, 73, 73, 73, 72, 72, 72, 71, 71, 71, 70, 70, 69, 69, 68, 67, 67, 66, 65, 65, 64,
                          63, 62, 61, 60, 60, 59, 58, 57, 56, 55, 54, 52, 51, 50, 49, 48, 47, 45, 44, 43, 42, 40, 39, 38, 36,
                          35, 33, 32, 30, 29, 28, 26, 25, 23, 22, 20, 19, 17, 16, 14, 12, 11, 9,  8,  6,  5,  3,  2};
template <class R> i64 nint_r(R x) { return (i64)llround((double)x); }   // Fortran NINT: half away from zero

// ipxwafs (variant 1), ipxwafs2 (2), ipxwafs3 (3): template bookkeeping on the host, the field loop on the device
template <class R, class I>
void api_xwafs(int variant, const I* idir, const I* nthin, const I* nfull, const I* km, const I* num_opt, I* opt, const I* len, I* tthin, R* dthin,
               I* ibthin, u8* bthin, I* tfull, R* dfull, I* ibfull, u8* bfull, I* iret) {
  try {
    *iret = 0;
    if (*len != 19 || *nthin != 3447 || *nfull != 5329) { *iret = 1; return; }
    const i64 dir = *idir;
    if (dir > 0) {          // thin -> full: the full grid's template follows from the thin one (ipxwafs.F90:135-156)
      i64 iscale = (i64)tthin[9] * (i64)tthin[10];
      if (iscale == 0) iscale = 1000000;
      const i64 idlat = nint_r<R>((R)1.25 * (R)iscale);
      bool up = *num_opt == 73, down = *num_opt == 73;
      for (int j = 0; j < 73 && (up || down); j++) { up = up && opt[j] == kWafsRow[j]; down = down && opt[j] == kWafsRow[72 - j]; }
      if (!(tthin[18] == 64 && tthin[8] == 73 && idlat == (i64)tthin[17] && (up || down))) { *iret = 1; return; }
      for (int q = 0; q < 19; q++) tfull[q] = tthin[q];
      tfull[7] = 73;
      const R rlon1 = (R)tfull[12] / (R)iscale, rlon2 = (R)tfull[15] / (R)iscale;
      const R hi = ((tfull[18] / 128) % 2) ? (R)-1 : (R)1;
      const R dlon = hi * (std::fmod(hi * (rlon2 - rlon1) - (R)1 + (R)3600, (R)360) + (R)1) / (R)72;
      tfull[16] = (I)nint_r<R>(dlon * (R)iscale);
    } else if (dir < 0) {   // full -> thin (ipxwafs.F90:157-177)
      i64 iscale = (i64)tfull[9] * (i64)tfull[10];
      if (iscale == 0) iscale = 1000000;
      const i64 idl = nint_r<R>((R)1.25 * (R)iscale);
      if (!(tfull[18] == 64 && tfull[7] == 73 && tfull[8] == 73 && *num_opt == 73 && idl == (i64)tfull[17] && idl == (i64)tfull[16])) { *iret = 1; return; }
      for (int q = 0; q < 19; q++) tthin[q] = tfull[q];
      tthin[7] = (I)-1; tthin[16] = (I)-1;
      for (int j = 0; j < 73; j++) opt[j] = (I)(tthin[11] == 0 ? kWafsRow[j] : kWafsRow[72 - j]);
    }
    if (dir != 1 && dir != -1) return;
    const bool expand = dir == 1;
    auto o = widen(opt, 73);
    std::vector<i64> ibin((size_t)std::max<i64>(*km, 1), 0), ibout((size_t)std::max<i64>(*km, 1), 0);
    if (variant != 1) for (i64 k = 0; k < *km; k++) ibin[k] = expand ? ibthin[k] : ibfull[k];
    run_xwafs<R>(variant, expand, *km, o.data(), tfull[8], tfull[7], *nthin, *nfull, ibin.data(), expand ? dthin : dfull, expand ? bthin : bfull,
                 expand ? dfull : dthin, expand ? bfull : bthin, ibout.data());
    if (variant != 1) for (i64 k = 0; k < *km; k++) (expand ? ibfull : ibthin)[k] = (I)ibout[k];
  } catch (cudaError_t e) { *iret = (I)(1000 + (int)e); } catch (...) { *iret = (I)1001; }
}

}  // namespace
}  // namespace ipb

using namespace ipb;

namespace {
struct HostMap { void* p; size_t len; };
std::mutex g_host_mu;
std::vector<HostMap> g_host_maps;
}  // namespace

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
  void ipgpu_ipolates_grib1_dev_##SFX(void* stream, const I* ip, const I* ipopt, const I* kgdsi, const I* kgdso,         \
                                      const I* mi, const I* mo, const I* km, const I* ibi, const u8* li, const R* gi,    \
                                      I* no, R* rlat, R* rlon, I* ibo, u8* lo, R* go, I* iret) {                          \
    api_interp<R, I>(false, false, 1, stream, ip, ipopt, nullptr, kgdsi, nullptr, nullptr, kgdso, nullptr, mi, mo, *km,  \
                     ibi, li, gi, nullptr, no, rlat, rlon, nullptr, nullptr, ibo, lo, go, nullptr, iret);                \
  }                                                                                                                      \
  void ipgpu_ipolatev_grib1_dev_##SFX(void* stream, const I* ip, const I* ipopt, const I* kgdsi, const I* kgdso,         \
                                      const I* mi, const I* mo, const I* km, const I* ibi, const u8* li, const R* ui,    \
                                      const R* vi, I* no, R* rlat, R* rlon, R* crot, R* srot, I* ibo, u8* lo, R* uo,     \
                                      R* vo, I* iret) {                                                                  \
    api_interp<R, I>(true, false, 1, stream, ip, ipopt, nullptr, kgdsi, nullptr, nullptr, kgdso, nullptr, mi, mo, *km,   \
                     ibi, li, ui, vi, no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret);                                \
  }                                                                                                                      \
  /* explicit plan handles: build / fetch once, then queue applies without draining the stream */                       \
  void* ipgpu_plan_get_##SFX(const I* vector, const I* ip, const I* ipopt, const I* numi, const I* tmpli, const I* leni, \
                             const I* numo, const I* tmplo, const I* leno, const I* mi, const I* mo, I* no, I* iret) {   \
    return api_plan_get<R, I>(*vector != 0, ip, ipopt, numi, tmpli, leni, numo, tmplo, leno, mi, mo, no, iret);          \
  }                                                                                                                      \
  void ipgpu_plan_geometry_##SFX(void* plan, void* stream, R* rlat, R* rlon, R* crot, R* srot, I* iret) {                \
    try { *iret = (I)plan_geometry<R>((PlanHandle*)plan, stream, rlat, rlon, crot, srot); }                             \
    catch (cudaError_t e) { *iret = (I)(1000 + (int)e); } catch (...) { *iret = (I)1001; }                               \
  }                                                                                                                      \
  void ipgpu_plan_apply_##SFX(void* plan, void* stream, const I* km, const I* ibi, const u8* li, const R* gi,            \
                              const R* vi, int32_t* ibo_dev, u8* lo, R* go, R* vo, I* iret) {                            \
    api_plan_apply<R, I>((PlanHandle*)plan, stream, *km, ibi, li, gi, vi, ibo_dev, lo, go, vo, iret);                    \
  }                                                                                                                      \
  /* grid expanders (src/ipxwafs.F90:81, ipxwafs2.F90:92, ipxwafs3.F90:95, ipxetas.F90:90): every argument by reference */ \
  void ipxwafs_##SFX(const I* idir, const I* nthin, const I* nfull, const I* km, const I* num_opt, I* opt, const I* len, \
                     I* tthin, R* dthin, I* tfull, R* dfull, I* iret) {                                                  \
    api_xwafs<R, I>(1, idir, nthin, nfull, km, num_opt, opt, len, tthin, dthin, nullptr, nullptr, tfull, dfull, nullptr, \
                    nullptr, iret);                                                                                      \
  }                                                                                                                      \
  void ipxwafs2_##SFX(const I* idir, const I* nthin, const I* nfull, const I* km, const I* num_opt, I* opt, const I* len, \
                      I* tthin, R* dthin, I* ibthin, u8* bthin, I* tfull, R* dfull, I* ibfull, u8* bfull, I* iret) {     \
    api_xwafs<R, I>(2, idir, nthin, nfull, km, num_opt, opt, len, tthin, dthin, ibthin, bthin, tfull, dfull, ibfull,     \
                    bfull, iret);                                                                                        \
  }                                                                                                                      \
  void ipxwafs3_##SFX(const I* idir, const I* nthin, const I* nfull, const I* km, const I* num_opt, I* opt, const I* len, \
                      I* tthin, R* dthin, I* ibthin, u8* bthin, I* tfull, R* dfull, I* ibfull, u8* bfull, I* iret) {     \
    api_xwafs<R, I>(3, idir, nthin, nfull, km, num_opt, opt, len, tthin, dthin, ibthin, bthin, tfull, dfull, ibfull,     \
                    bfull, iret);                                                                                        \
  }                                                                                                                      \
  void ipxetas_##SFX(const I* idir, const I* numi, const I* len, const I* tin, const I* npi, const u8* bi, const R* di,  \
                     I* numo, I* tout, const I* npo, u8* bo, R* dout, I* iret) {                                         \
    /* staggered E grid (H / V points, scan 68 / 72) <-> full E grid (scan 64): the output template, then ipolates */   \
    *iret = 0;                                                                                                           \
    if (*numi != 1) { *iret = 1; return; }                                                                               \
    const I scan = tin[18];                                                                                              \
    int mode = 0; /* 1: staggered -> full, 2: full -> H, 3: full -> V */                                                 \
    if ((scan == 68 || scan == 72) && (*idir < -2 || *idir > -1)) mode = 1;                                              \
    else if (scan == 64 && *idir == -1) mode = 2;                                                                        \
    else if (scan == 64 && *idir == -2) mode = 3;                                                                        \
    *numo = *numi;                                                                                                       \
    for (I q = 0; q < *len; q++) tout[q] = tin[q];                                                                       \
    if (mode == 0) { *iret = 2; return; }                                                                                \
    tout[18] = mode == 1 ? 64 : (mode == 2 ? 68 : 72);                                                                   \
    tout[7] = mode == 1 ? tout[7] * 2 - 1 : (tout[7] + 1) / 2;                                                           \
    if (tout[7] * tout[8] != *npo) { *iret = 3; return; }                                                                \
    {                                                                                                                    \
      i64 iscale = (i64)tout[9] * (i64)tout[10];                                                                         \
      if (iscale == 0) iscale = 1000000;                                                                                 \
      R dlons = (R)tout[16] / (R)iscale;                                                                                 \
      dlons = dlons * (mode == 1 ? (R)0.5 : (R)2);                                                                       \
      tout[16] = (I)nint_r<R>(dlons * (R)iscale);                                                                        \
    }                                                                                                                    \
    const I ip = 0, km = 1, ibi = 1;                                                                                     \
    I ipopt[20] = {0}, ibo = 0, no = 0;                                                                                  \
    try {                                                                                                                \
      std::vector<R> rlat((size_t)*npo), rlon((size_t)*npo);                                                             \
      ipolates_grib2_##SFX(&ip, ipopt, numi, tin, len, numo, tout, len, npi, npo, &km, &ibi, bi, di, &no, rlat.data(),   \
                           rlon.data(), &ibo, bo, dout, iret);                                                           \
    } catch (...) { *iret = (I)1001; }                                                                                   \
    if (*iret != 0) return;                                                                                              \
    const i64 im = (i64)tout[7], jm = (i64)tout[8];                                                                      \
    for (i64 j = 1; j <= jm; j++) { /* the row ends copy their neighbours (ipxetas.F90:195-200) */                       \
      bo[j * im - 1] = bo[j * im - 2]; dout[j * im - 1] = dout[j * im - 2];                                              \
      bo[(j - 1) * im] = bo[(j - 1) * im + 1]; dout[(j - 1) * im] = dout[(j - 1) * im + 1];                              \
    }                                                                                                                    \
  }                                                                                                                      \
  void gdswzd_##SFX(I num, const I* tmpl, I len, I iopt, I npts, R fill, R* xpts, R* ypts, R* rlon, R* rlat, I* nret,    \
                    R* crot, R* srot, R* xlon, R* xlat, R* ylon, R* ylat, R* area) {                                     \
    try {                                                                                                                \
      auto t = widen(tmpl, len);                                                                                         \
      i64 nr = *nret;                                                                                                    \
      run_gdswzd<R>(decode_grib2<R>(num, t.data(), len), iopt, npts, fill, xpts, ypts, rlon, rlat, nr, crot, srot, xlon, \
                    xlat, ylon, ylat, area);                                                                             \
      *nret = (I)nr;                                                                                                     \
    } catch (...) { *nret = 0; }                                                                                         \
  }                                                                                                                      \
  void gdswzd_grib1_##SFX(const I* kgds, I iopt, I npts, R fill, R* xpts, R* ypts, R* rlon, R* rlat, I* nret, R* crot,   \
                          R* srot, R* xlon, R* xlat, R* ylon, R* ylat, R* area) {                                        \
    try {                                                                                                                \
      auto t = widen(kgds, 200);                                                                                         \
      i64 nr = *nret;                                                                                                    \
      run_gdswzd<R>(decode_grib1<R>(t.data()), iopt, npts, fill, xpts, ypts, rlon, rlat, nr, crot, srot, xlon, xlat,     \
                    ylon, ylat, area);                                                                                   \
      *nret = (I)nr;                                                                                                     \
    } catch (...) { *nret = 0; }                                                                                         \
  }                                                                                                                      \
  }

IPB_KIND(4, float, int32_t)
IPB_KIND(d, double, int32_t)
IPB_KIND(8, double, int64_t)

extern "C" {
// library management (new; no reference counterpart)
int ipgpu_device_count() { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }
int ipgpu_set_device(int d) { return (int)cudaSetDevice(d); }
// Host-pointer calls shard their field batch over the first n devices (one host thread, one PCIe link and one plan copy
// each); n <= 1 restores the default (the calling thread's current device).  Returns the number of devices in use.
int ipgpu_set_devices(int n) {
  int have = ipgpu_device_count();
  Runtime& RT = rt();
  std::lock_guard<std::mutex> lock(RT.mu);
  RT.shard.clear();
  if (n > have) n = have;
  if (n > 1) for (int d = 0; d < n; d++) RT.shard.push_back(d);
  return n > 1 ? n : 1;
}
void ipgpu_plan_release(void* plan) { delete (PlanHandle*)plan; }
void ipgpu_plan_clear() {
  Runtime& RT = rt();
  std::vector<DevState*> all;
  { std::lock_guard<std::mutex> lock(RT.mu); for (auto& d : RT.dev) if (d) all.push_back(d.get()); }
  for (DevState* d : all) {
    std::lock_guard<std::mutex> lock(d->mu);
    DeviceGuard guard(d->device);
    d->cache.clear();
    for (int i = 0; i < 2; i++) { Work& w = d->work[i]; w.gi.release(); w.vi.release(); w.li.release(); w.go.release(); w.vo.release(); w.lo.release(); w.ints.release(); w.lbits.release(); w.hbits.release(); w.hin.release(); w.hout.release(); }
    d->geo.release();
  }
}
long long ipgpu_plan_count() {
  Runtime& RT = rt(); long long n = 0;
  std::lock_guard<std::mutex> lock(RT.mu);
  for (auto& d : RT.dev) if (d) { std::lock_guard<std::mutex> l2(d->mu); n += (long long)d->cache.size(); }
  return n;
}
long long ipgpu_plan_bytes() {
  Runtime& RT = rt(); size_t t = 0;
  std::lock_guard<std::mutex> lock(RT.mu);
  for (auto& d : RT.dev) if (d) { std::lock_guard<std::mutex> l2(d->mu); for (auto& e : d->cache) t += e.bytes; }
  return (long long)t;
}
void ipgpu_set_plan_cache_bytes(long long b) { Runtime& RT = rt(); std::lock_guard<std::mutex> lock(RT.mu); RT.cache_cap = (size_t)b; }
void ipgpu_set_chunk_bytes(long long b) { Runtime& RT = rt(); std::lock_guard<std::mutex> lock(RT.mu); RT.chunk_bytes = (size_t)std::max<long long>(b, 1 << 20); }
long long ipgpu_launch_count() { return g_launches.load(); }
// The host half of the host-pointer pipeline, callable without a GPU (tests -m "not gpu"): what = 0 rebuilds lo[kn][mo]
// (first no points of each field) from bits[kn][ceil(no/32)] and hasfalse[kn]; what = 1 copies kn rows of `width` bytes
// from src (pitch spitch) to dst (pitch dpitch).  Returns 0, or 1 for an unknown `what`.
int ipgpu_selftest_host(int what, int threads, unsigned char* dst, long long dpitch, const unsigned char* src, long long spitch, long long width,
                        long long kn, const int* hasfalse) {
  if (what == 0) { expand_lo(dst, dpitch, width, kn, reinterpret_cast<const unsigned*>(src), (width + 31) / 32, hasfalse, threads); return 0; }
  if (what == 1) { copy_rows(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)kn, threads); return 0; }
  return 1;
}
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
// Pinned host memory for the host-pointer API.  The pages are placed on the NUMA node the current device's PCIe link
// hangs off (mbind + first touch, then cudaHostRegister): a buffer on the far socket halves the link rate.  Falls back to
// cudaMallocHost when the node is unknown or the policy call is not permitted.
void* ipgpu_host_alloc(long long bytes) {
  if (bytes <= 0) return nullptr;
  int dev = 0, node = -1;
  char bus[64] = {0};
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) == cudaSuccess) {
    for (char* c = bus; *c; c++) *c = (char)tolower(*c);
    char path[160];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    if (FILE* f = fopen(path, "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
  }
  static const bool numa_off = getenv("IPB_NO_NUMA") != nullptr;
  if (node >= 0 && node < 1024 && !numa_off) {
    const size_t len = ((size_t)bytes + 4095) & ~(size_t)4095;
    void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p != MAP_FAILED) {
      unsigned long mask[16] = {0};
      mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
      (void)syscall(SYS_mbind, p, len, 2 /* MPOL_BIND */, mask, (unsigned long)(8 * sizeof(mask)), 0u);   // best effort
      memset(p, 0, len);                                                                                   // first touch places the pages
      if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {
        std::lock_guard<std::mutex> lock(g_host_mu);
        g_host_maps.push_back({p, len});
        return p;
      }
      cudaGetLastError();
      munmap(p, len);
    }
  }
  void* p = nullptr;
  if (cudaMallocHost(&p, (size_t)bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void ipgpu_host_free(void* p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lock(g_host_mu);
    for (size_t i = 0; i < g_host_maps.size(); i++)
      if (g_host_maps[i].p == p) {
        cudaHostUnregister(p);
        munmap(p, g_host_maps[i].len);
        g_host_maps.erase(g_host_maps.begin() + (long)i);
        return;
      }
  }
  cudaFreeHost(p);
}
// Page-lock arrays the caller already owns (a Fortran allocatable, a C malloc): host-pointer calls then copy them directly at
// link rate instead of staging them.  Returns 0, or the cudaError (e.g. the range is registered already).
int ipgpu_host_register(void* p, long long bytes) {
  if (!p || bytes <= 0) return (int)cudaErrorInvalidValue;
  const cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) cudaGetLastError();
  return (int)e;
}
int ipgpu_host_unregister(void* p) {
  const cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) cudaGetLastError();
  return (int)e;
}
// NUMA node of the current device's PCIe link (-1 unknown)
int ipgpu_device_numa_node() {
  int dev = 0, node = -1;
  char bus[64] = {0};
  if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetPCIBusId(bus, sizeof(bus), dev) == cudaSuccess) {
    for (char* c = bus; *c; c++) *c = (char)tolower(*c);
    char path[160];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    if (FILE* f = fopen(path, "r")) { if (fscanf(f, "%d", &node) != 1) node = -1; fclose(f); }
  }
  return node;
}
// Measured ceiling of the current device's host link with this library's own pinned allocator: GB/s of one large
// host->device and one device->host copy, and of both directions at once (what a streamed call can reach at best).
void ipgpu_link_probe(long long bytes, double* h2d_gbs, double* d2h_gbs, double* both_gbs) {
  if (h2d_gbs) *h2d_gbs = 0; if (d2h_gbs) *d2h_gbs = 0; if (both_gbs) *both_gbs = 0;
  if (bytes < (1 << 20)) bytes = 1 << 20;
  void *h0 = ipgpu_host_alloc(bytes), *h1 = ipgpu_host_alloc(bytes), *d0 = nullptr, *d1 = nullptr;
  cudaStream_t s0 = nullptr, s1 = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h0 && h1 && cudaMalloc(&d0, (size_t)bytes) == cudaSuccess && cudaMalloc(&d1, (size_t)bytes) == cudaSuccess &&
      cudaStreamCreate(&s0) == cudaSuccess && cudaStreamCreate(&s1) == cudaSuccess && cudaEventCreate(&e0) == cudaSuccess &&
      cudaEventCreate(&e1) == cudaSuccess) {
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {   // second repetition is the measurement
      cudaEventRecord(e0, s0); cudaMemcpyAsync(d0, h0, (size_t)bytes, cudaMemcpyHostToDevice, s0); cudaEventRecord(e1, s0); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1); if (h2d_gbs) *h2d_gbs = bytes / (ms * 1e6);
      cudaEventRecord(e0, s0); cudaMemcpyAsync(h1, d1, (size_t)bytes, cudaMemcpyDeviceToHost, s0); cudaEventRecord(e1, s0); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1); if (d2h_gbs) *d2h_gbs = bytes / (ms * 1e6);
      cudaStreamSynchronize(s1);
      cudaEventRecord(e0, s0);
      cudaMemcpyAsync(d0, h0, (size_t)bytes, cudaMemcpyHostToDevice, s0);
      cudaMemcpyAsync(h1, d1, (size_t)bytes, cudaMemcpyDeviceToHost, s1);
      cudaStreamSynchronize(s1);
      cudaEventRecord(e1, s0); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1); if (both_gbs) *both_gbs = 2.0 * bytes / (ms * 1e6);
    }
  }
  cudaGetLastError();
  if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1);
  if (s0) cudaStreamDestroy(s0); if (s1) cudaStreamDestroy(s1);
  if (d0) cudaFree(d0); if (d1) cudaFree(d1);
  ipgpu_host_free(h0); ipgpu_host_free(h1);
}
const char* ipgpu_version() { return "ipb200 0.2 (sm_100a)"; }
}
