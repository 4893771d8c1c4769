// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
//
// C entry points of the CPU oracle.  One set per reference build kind:
//   _4 : int32 / float   (libip_4)      _d : int32 / double (libip_d)      _8 : int64 / double (libip_8)
// Argument lists are those of the reference's bind(c) drivers
// (src/ipolates.F90:158,293,587,808; src/ipolatev.F90:382,565,680,832; src/gdswzd_c.F90:173,259),
// every argument by reference except the by-value scalars of gdswzd / gdswzd_grib1.
#include "ip_oracle_interp.hpp"
#include "ip_oracle_expand.hpp"
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace ipo;

namespace {

template <class I> std::vector<i64> widen(const I* p, i64 n) {
  std::vector<i64> v(n);
  for (i64 i = 0; i < n; i++) v[i] = (i64)p[i];
  return v;
}

template <class R, class I>
void ipolates_any(bool grib2, const I* ip, const I* ipopt, const I* numi, const I* dsci, const I* leni, const I* numo,
                  const I* dsco, const I* leno, const I* mi, const I* mo, const I* km, const I* ibi, const u8* li,
                  const R* gi, I* no, R* rlat, R* rlon, I* ibo, u8* lo, R* go, I* iret) {
  Grid<R> gin, gout;
  bool ok;
  if (grib2) {
    auto ti = widen(dsci, *leni), to = widen(dsco, *leno);
    ok = init_grid_grib2<R>(gin, *numi, ti.data(), *leni) && init_grid_grib2<R>(gout, *numo, to.data(), *leno);
  } else {
    auto ti = widen(dsci, 200), to = widen(dsco, 200);
    ok = init_grid_grib1<R>(gin, ti.data()) && init_grid_grib1<R>(gout, to.data());
  }
  if (!ok) { *iret = 1; return; }  // reference: error stop
  auto opt = widen(ipopt, 20), ib = widen(ibi, *km);
  std::vector<i64> ibo64(*km > 0 ? *km : 1, 0);
  i64 no64 = *no, iret64 = 0;
  ipolates_grid<R>(*ip, opt.data(), gin, gout, *mi, *mo, *km, ib.data(), li, gi, no64, rlat, rlon, ibo64.data(), lo, go, iret64);
  *no = (I)no64; *iret = (I)iret64;
  if (iret64 == 0 || *ip == 3 || *ip == 6) for (i64 k = 0; k < *km; k++) ibo[k] = (I)ibo64[k];
}

template <class R, class I>
void ipolatev_any(bool grib2, const I* ip, const I* ipopt, const I* numi, const I* dsci, const I* leni, const I* numo,
                  const I* dsco, const I* leno, const I* mi, const I* mo, const I* km, const I* ibi, const u8* li,
                  const R* ui, const R* vi, I* no, R* rlat, R* rlon, R* crot, R* srot, I* ibo, u8* lo, R* uo, R* vo,
                  I* iret) {
  Grid<R> gin, gout;
  bool ok;
  if (grib2) {
    auto ti = widen(dsci, *leni), to = widen(dsco, *leno);
    ok = init_grid_grib2<R>(gin, *numi, ti.data(), *leni) && init_grid_grib2<R>(gout, *numo, to.data(), *leno);
  } else {
    // src/ipolatev.F90:602-609: E-grids are treated as wind points for the duration of the call
    auto ti = widen(dsci, 200), to = widen(dsco, 200);
    if (ti[0] == 203) ti[10] |= 256;
    if (to[0] == 203) to[10] |= 256;
    ok = init_grid_grib1<R>(gin, ti.data()) && init_grid_grib1<R>(gout, to.data());
  }
  if (!ok) { *iret = 1; return; }
  auto opt = widen(ipopt, 20), ib = widen(ibi, *km);
  std::vector<i64> ibo64(*km > 0 ? *km : 1, 0);
  i64 no64 = *no, iret64 = 0;
  ipolatev_grid<R>(*ip, opt.data(), gin, gout, *mi, *mo, *km, ib.data(), li, ui, vi, no64, rlat, rlon, crot, srot,
                   ibo64.data(), lo, uo, vo, iret64);
  *no = (I)no64; *iret = (I)iret64;
  if (iret64 == 0 || *ip == 3 || *ip == 6) for (i64 k = 0; k < *km; k++) ibo[k] = (I)ibo64[k];
}

template <class R, class I>
void gdswzd_any(bool grib2, I num, const I* dsc, I len, I iopt, I npts, R fill, R* xpts, R* ypts, R* rlon, R* rlat,
                I* nret, R* crot, R* srot, R* xlon, R* xlat, R* ylon, R* ylat, R* area) {
  Grid<R> g;
  bool ok;
  if (grib2) { auto t = widen(dsc, len); ok = init_grid_grib2<R>(g, num, t.data(), len); }
  else { auto t = widen(dsc, 200); ok = init_grid_grib1<R>(g, t.data()); }
  if (!ok) { *nret = 0; return; }
  GdsOpt<R> o; o.crot = crot; o.srot = srot; o.xlon = xlon; o.xlat = xlat; o.ylon = ylon; o.ylat = ylat; o.area = area;
  i64 nr = *nret;
  gdswzd_grid<R>(g, (int)iopt, npts, fill, xpts, ypts, rlon, rlat, nr, o);
  *nret = (I)nr;
}

}  // namespace

template <class R, class I>
void ipxwafs_c(int variant, const I* idir, const I* nthin, const I* nfull, const I* km, const I* num_opt, I* opt, const I* len, I* tthin, R* dthin,
               I* ibthin, u8* bthin, I* tfull, R* dfull, I* ibfull, u8* bfull, I* iret) {
  auto o = widen(opt, *num_opt), tt = widen(tthin, *len), tf = widen(tfull, *len);
  std::vector<i64> it(*km > 0 ? *km : 1, 0), ifu(*km > 0 ? *km : 1, 0);
  if (variant != 1) for (i64 k = 0; k < *km; k++) { it[k] = ibthin[k]; ifu[k] = ibfull[k]; }
  i64 ir = 0;
  ipxwafs_any<R>(variant, *idir, *nthin, *nfull, *km, *num_opt, o.data(), *len, tt.data(), dthin, it.data(), bthin, tf.data(), dfull, ifu.data(), bfull, ir);
  *iret = (I)ir;
  for (i64 q = 0; q < *num_opt; q++) opt[q] = (I)o[q];
  for (i64 q = 0; q < *len; q++) { tthin[q] = (I)tt[q]; tfull[q] = (I)tf[q]; }
  if (variant != 1) for (i64 k = 0; k < *km; k++) { ibthin[k] = (I)it[k]; ibfull[k] = (I)ifu[k]; }
}
template <class R, class I>
void ipxetas_c(const I* idir, const I* numi, const I* len, const I* tin, const I* npi, const u8* bi, const R* di, I* numo, I* tout, const I* npo,
               u8* bo, R* dout, I* iret) {
  auto ti = widen(tin, *len);
  std::vector<i64> to(*len, 0);
  i64 no = 0, ir = 0;
  ipxetas<R>(*idir, *numi, *len, ti.data(), *npi, bi, di, no, to.data(), *npo, bo, dout, ir);
  *iret = (I)ir;
  if (ir != 1 || *numi == 1) { *numo = (I)no; for (i64 q = 0; q < *len; q++) tout[q] = (I)to[q]; }
}

#define ORACLE_KIND(SFX, R, I)                                                                                          \
  extern "C" {                                                                                                          \
  void oracle_ipxwafs_##SFX(const I* idir, const I* nthin, const I* nfull, const I* km, const I* num_opt, I* opt, const I* len, I* tthin,  \
                            R* dthin, I* tfull, R* dfull, I* iret) {                                                    \
    ipxwafs_c<R, I>(1, idir, nthin, nfull, km, num_opt, opt, len, tthin, dthin, nullptr, nullptr, tfull, dfull, nullptr, nullptr, iret);   \
  }                                                                                                                     \
  void oracle_ipxwafs2_##SFX(const I* idir, const I* nthin, const I* nfull, const I* km, const I* num_opt, I* opt, const I* len, I* tthin, \
                             R* dthin, I* ibthin, u8* bthin, I* tfull, R* dfull, I* ibfull, u8* bfull, I* iret) {       \
    ipxwafs_c<R, I>(2, idir, nthin, nfull, km, num_opt, opt, len, tthin, dthin, ibthin, bthin, tfull, dfull, ibfull, bfull, iret);         \
  }                                                                                                                     \
  void oracle_ipxwafs3_##SFX(const I* idir, const I* nthin, const I* nfull, const I* km, const I* num_opt, I* opt, const I* len, I* tthin, \
                             R* dthin, I* ibthin, u8* bthin, I* tfull, R* dfull, I* ibfull, u8* bfull, I* iret) {       \
    ipxwafs_c<R, I>(3, idir, nthin, nfull, km, num_opt, opt, len, tthin, dthin, ibthin, bthin, tfull, dfull, ibfull, bfull, iret);         \
  }                                                                                                                     \
  void oracle_ipxetas_##SFX(const I* idir, const I* numi, const I* len, const I* tin, const I* npi, const u8* bi, const R* di, I* numo,    \
                            I* tout, const I* npo, u8* bo, R* dout, I* iret) {                                          \
    ipxetas_c<R, I>(idir, numi, len, tin, npi, bi, di, numo, tout, npo, bo, dout, iret);                                \
  }                                                                                                                     \
  void oracle_ipolates_grib2_##SFX(const I* ip, const I* ipopt, const I* numi, const I* tmpli, const I* leni,           \
                                   const I* numo, const I* tmplo, const I* leno, const I* mi, const I* mo, const I* km, \
                                   const I* ibi, const u8* li, const R* gi, I* no, R* rlat, R* rlon, I* ibo, u8* lo,   \
                                   R* go, I* iret) {                                                                    \
    ipolates_any<R, I>(true, ip, ipopt, numi, tmpli, leni, numo, tmplo, leno, mi, mo, km, ibi, li, gi, no, rlat, rlon,  \
                       ibo, lo, go, iret);                                                                              \
  }                                                                                                                     \
  void oracle_ipolates_grib1_##SFX(const I* ip, const I* ipopt, const I* kgdsi, const I* kgdso, const I* mi,            \
                                   const I* mo, const I* km, const I* ibi, const u8* li, const R* gi, I* no, R* rlat,   \
                                   R* rlon, I* ibo, u8* lo, R* go, I* iret) {                                           \
    ipolates_any<R, I>(false, ip, ipopt, nullptr, kgdsi, nullptr, nullptr, kgdso, nullptr, mi, mo, km, ibi, li, gi, no, \
                       rlat, rlon, ibo, lo, go, iret);                                                                  \
  }                                                                                                                     \
  void oracle_ipolatev_grib2_##SFX(const I* ip, const I* ipopt, const I* numi, const I* tmpli, const I* leni,           \
                                   const I* numo, const I* tmplo, const I* leno, const I* mi, const I* mo, const I* km, \
                                   const I* ibi, const u8* li, const R* ui, const R* vi, I* no, R* rlat, R* rlon,      \
                                   R* crot, R* srot, I* ibo, u8* lo, R* uo, R* vo, I* iret) {                           \
    ipolatev_any<R, I>(true, ip, ipopt, numi, tmpli, leni, numo, tmplo, leno, mi, mo, km, ibi, li, ui, vi, no, rlat,    \
                       rlon, crot, srot, ibo, lo, uo, vo, iret);                                                        \
  }                                                                                                                     \
  void oracle_ipolatev_grib1_##SFX(const I* ip, const I* ipopt, const I* kgdsi, const I* kgdso, const I* mi,            \
                                   const I* mo, const I* km, const I* ibi, const u8* li, const R* ui, const R* vi,     \
                                   I* no, R* rlat, R* rlon, R* crot, R* srot, I* ibo, u8* lo, R* uo, R* vo, I* iret) { \
    ipolatev_any<R, I>(false, ip, ipopt, nullptr, kgdsi, nullptr, nullptr, kgdso, nullptr, mi, mo, km, ibi, li, ui, vi, \
                       no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret);                                              \
  }                                                                                                                     \
  void oracle_gdswzd_##SFX(I num, const I* tmpl, I len, I iopt, I npts, R fill, R* xpts, R* ypts, R* rlon, R* rlat,     \
                           I* nret, R* crot, R* srot, R* xlon, R* xlat, R* ylon, R* ylat, R* area) {                    \
    gdswzd_any<R, I>(true, num, tmpl, len, iopt, npts, fill, xpts, ypts, rlon, rlat, nret, crot, srot, xlon, xlat,      \
                     ylon, ylat, area);                                                                                 \
  }                                                                                                                     \
  void oracle_gdswzd_grib1_##SFX(const I* kgds, I iopt, I npts, R fill, R* xpts, R* ypts, R* rlon, R* rlat, I* nret,    \
                                 R* crot, R* srot, R* xlon, R* xlat, R* ylon, R* ylat, R* area) {                       \
    gdswzd_any<R, I>(false, 0, kgds, 200, iopt, npts, fill, xpts, ypts, rlon, rlat, nret, crot, srot, xlon, xlat, ylon, \
                     ylat, area);                                                                                       \
  }                                                                                                                     \
  }

ORACLE_KIND(4, float, int32_t)
ORACLE_KIND(d, double, int32_t)
ORACLE_KIND(8, double, int64_t)

// launchers such as torchrun export OMP_NUM_THREADS=1; the timing harness sets the thread count explicitly
extern "C" void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#endif
}

extern "C" int oracle_num_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// src/earth_radius_mod.F90:40 (exposed for tests/test_oracle_golden.py::test_earth_radius)
extern "C" void oracle_earth_radius_4(const int32_t* t, int32_t len, float* radius, float* e2) {
  auto w = widen(t, len); w.resize(7, 0); earth_radius<float>(w.data(), *radius, *e2);
}
extern "C" void oracle_earth_radius_d(const int32_t* t, int32_t len, double* radius, double* e2) {
  auto w = widen(t, len); w.resize(7, 0); earth_radius<double>(w.data(), *radius, *e2);
}
