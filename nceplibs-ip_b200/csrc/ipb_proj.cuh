// ipb200 — device-side map projections: one thread, one point.
//
// proj_fwd : grid (x,y)      -> earth (rlon,rlat)      [gdswzd iopt=+1]
// proj_inv : earth (rlon,rlat) -> grid (x,y)           [gdswzd iopt=-1]
// each optionally producing the vector rotation (crot,srot), the map Jacobian and the cell area.
// Behaviour follows the type-bound gdswzd of the reference's eight grid classes
// (src/ip_equid_cylind_grid_mod.F90:203, ip_mercator_grid_mod.F90:198, ip_lambert_conf_grid_mod.F90:219,
//  ip_gaussian_grid_mod.F90:195, ip_polar_stereo_grid_mod.F90:239, ip_rot_equid_cylind_grid_mod.F90:256,
//  ip_rot_equid_cylind_egrid_mod.F90:267, ip_station_points_grid_mod.F90:75) with the per-call
// constants precomputed in GridP (ipb_decode.h).
//
// Arithmetic contract: compiled with -fmad=false; R=float rounds every +,-,*,/ to binary32 and
// defines each transcendental as "correctly rounded binary64 evaluation (ipb_crmath.h) rounded once
// more" (DM<float>); R=double uses the correctly rounded functions directly; the rotated
// grids compute in binary64 for both kinds, like the reference's real64 internals.
#pragma once
#include <cuda_runtime.h>

#include "ipb_crmath.h"
#include "ipb_grid.h"

namespace ipb {

#define IPB_DL(x) (sizeof(R) == 4 ? (R)(x##f) : (R)(x))

template <class R> struct DM {
  static __device__ __forceinline__ R sin(R x) { return (R)cr::sin((double)x); }
  static __device__ __forceinline__ R cos(R x) { return (R)cr::cos((double)x); }
  static __device__ __forceinline__ R tan(R x) { return (R)cr::tan((double)x); }
  static __device__ __forceinline__ R asin(R x) { return (R)cr::asin((double)x); }
  static __device__ __forceinline__ R atan(R x) { return (R)cr::atan((double)x); }
  static __device__ __forceinline__ R atan2(R y, R x) { return (R)cr::atan2((double)y, (double)x); }
  static __device__ __forceinline__ R log(R x) { return (R)cr::log((double)x); }
  static __device__ __forceinline__ R exp(R x) { return (R)cr::exp((double)x); }
  static __device__ __forceinline__ R pow(R a, R b) { return (R)cr::pow((double)a, (double)b); }
  static __device__ __forceinline__ R sqrt(R x) { return (R)::sqrt((double)x); }
  static __device__ __forceinline__ R fmod(R a, R b) { return (R)::fmod((double)a, (double)b); }
  static __device__ __forceinline__ R abs(R x) { return x < 0 ? -x : x; }
  static __device__ __forceinline__ R tiny() { return sizeof(R) == 4 ? (R)1.17549435e-38f : (R)2.2250738585072014e-308; }
  static __device__ __forceinline__ R huge() { return sizeof(R) == 4 ? (R)3.40282347e+38f : (R)1.7976931348623157e+308; }
  static __device__ __forceinline__ R pi() { return IPB_DL(3.14159265358979); }
  static __device__ __forceinline__ R dpr() { return IPB_DL(180.0) / pi(); }
  static __device__ __forceinline__ R vmin(R a, R b) { return a < b ? a : b; }
  static __device__ __forceinline__ R vmax(R a, R b) { return a > b ? a : b; }
};

// Fortran intrinsics on device
template <class R> __device__ __forceinline__ int f_nint(R x) { return (int)::llround((double)x); }
template <class R> __device__ __forceinline__ int f_int(R x) { return (int)x; }
template <class R> __device__ __forceinline__ int f_floor(R x) { return (int)::floor((double)x); }
template <class R> __device__ __forceinline__ R f_sign1(R b) { return (b < 0 || (b == 0 && ::signbit((double)b))) ? (R)-1 : (R)1; }

// What a caller wants besides the coordinates.
enum { EXT_NONE = 0, EXT_ROT = 1, EXT_ALL = 2 };
template <class R> struct PtExt { R crot, srot, xlon, xlat, ylon, ylat, area; };

// (i,j) -> 1-based position in the field, 0 = outside.  src/ip_grid_mod.F90:199-249
template <class R> __device__ __forceinline__ int field_pos(const GridP<R>& g, int i, int j) {
  int ii = i, jj = j;
  const int iw = g.iwrap;
  if (iw > 0) {
    ii = (i - 1 + iw) % iw + 1;
    if (j < 1 && g.jwrap1 > 0) { jj = g.jwrap1 - j; ii = (ii - 1 + iw / 2) % iw + 1; }
    else if (j > g.jm && g.jwrap2 > 0) { jj = g.jwrap2 - j; ii = (ii - 1 + iw / 2) % iw + 1; }
  }
  if (g.nscan_fp <= 1) {
    if (ii < 1 || ii > g.im || jj < 1 || jj > g.jm) return 0;
    return g.nscan_fp == 0 ? ii + (jj - 1) * g.im : jj + (ii - 1) * g.jm;
  }
  const int is1 = (g.jm + 1 - g.kscan) / 2;
  const int iif = jj + (ii - is1), jjf = jj - (ii - is1) + g.kscan;
  if (iif < 1 || iif > 2 * g.im - 1 || jjf < 1 || jjf > g.jm) return 0;
  return g.nscan_fp == 2 ? (iif + (jjf - 1) * (2 * g.im - 1) + 1 - g.kscan) / 2 : (iif + 1) / 2 + (jjf - 1) * g.im;
}

// Grid coordinates of the n-th (1-based) point of the grid's own enumeration: gdswzd iopt=0,
// src/gdswzd_mod.F90:146-183 (E-grids use the staggered diagonal numbering).
template <class R> __device__ __forceinline__ void enumerate_xy(const GridP<R>& g, int n, R& x, R& y) {
  const int im = g.im, jm = g.jm, ks = g.kscan;
  int i, j;
  if (g.type == P_ROTE) {
    const int is1 = (ks == 0) ? (jm + 1) / 2 : jm / 2;
    if (g.nscan == 0) { j = (n - 1) / im + 1; i = (n - im * (j - 1)) * 2 - (j + ks) % 2; }
    else {
      const int nn = n * 2 - 1 + ks;
      i = (nn - 1) / jm + 1; j = (nn - 1) % jm + 1;
      if (jm % 2 == 0 && i % 2 == 0) j += (ks == 0) ? 1 : -1;
    }
    x = (R)(is1 + (i - (j - ks)) / 2);
    y = (R)((i + (j - ks)) / 2);
    return;
  }
  if (g.nscan == 0) { j = (n - 1) / im + 1; i = n - im * (j - 1); }
  else { i = (n - 1) / jm + 1; j = n - jm * (i - 1); }
  x = (R)i; y = (R)j;
}

template <class R> __device__ __forceinline__ bool in_window(const GridP<R>& g, R x, R y) {
  return x >= g.xmin && x <= g.xmax && y >= g.ymin && y <= g.ymax;
}

namespace pd {  // per-projection pieces

// ---- cylindrical family: longitude <-> x is shared by lat-lon, Mercator, Gaussian ----------
template <class R> __device__ __forceinline__ R cyl_lon(const GridP<R>& g, R x) {
  return DM<R>::fmod(g.rlon1 + g.dlon * (x - IPB_DL(1.)) + IPB_DL(3600.), IPB_DL(360.));
}
template <class R> __device__ __forceinline__ R cyl_x(const GridP<R>& g, R rlon) {
  return IPB_DL(1.) + g.hi * DM<R>::fmod(g.hi * (rlon - g.rlon1) + IPB_DL(3600.), IPB_DL(360.)) / g.dlon;
}

template <class R, int EXT> __device__ __forceinline__ void equid_ext(const GridP<R>& g, R rlat, PtExt<R>& e) {
  if (EXT >= EXT_ROT) { e.crot = IPB_DL(1.0); e.srot = IPB_DL(0.0); }
  if (EXT >= EXT_ALL) {
    const R dpr = DM<R>::dpr();
    e.xlon = IPB_DL(1.0) / g.dlon; e.xlat = IPB_DL(0.); e.ylon = IPB_DL(0.); e.ylat = IPB_DL(1.0) / g.dlat;
    const R up = DM<R>::vmin(DM<R>::vmax(rlat + g.dlat / IPB_DL(2.), IPB_DL(-90.)), IPB_DL(90.));
    const R dn = DM<R>::vmin(DM<R>::vmax(rlat - g.dlat / IPB_DL(2.), IPB_DL(-90.)), IPB_DL(90.));
    const R ds = DM<R>::sin(up / dpr) - DM<R>::sin(dn / dpr);
    e.area = g.rerth * g.rerth * DM<R>::abs(ds * g.dlon) / dpr;
  }
}
template <class R, int EXT> __device__ __forceinline__ void merc_ext(const GridP<R>& g, R rlat, PtExt<R>& e) {
  if (EXT >= EXT_ROT) { e.crot = IPB_DL(1.0); e.srot = IPB_DL(0.0); }
  if (EXT >= EXT_ALL) {
    const R dpr = DM<R>::dpr();
    const R c = DM<R>::cos(rlat / dpr);
    e.xlon = IPB_DL(1.) / g.dlon; e.xlat = IPB_DL(0.); e.ylon = IPB_DL(0.);
    e.ylat = IPB_DL(1.) / g.dphi / c / dpr;
    e.area = g.rerth * g.rerth * (c * c) * g.dphi * g.dlon / dpr;
  }
}
template <class R, int EXT> __device__ __forceinline__ void gauss_ext(const GridP<R>& g, R y, PtExt<R>& e) {
  if (EXT >= EXT_ROT) { e.crot = IPB_DL(1.0); e.srot = IPB_DL(0.0); }
  if (EXT >= EXT_ALL) {
    const int top = g.jg + 1;
    e.xlon = IPB_DL(1.) / g.dlon; e.xlat = IPB_DL(0.); e.ylon = IPB_DL(0.);
    const int row = f_nint<R>(y);
    e.ylat = (row >= 0 && row <= top) ? g.ylat_row[row] : (R)0;
    const int j = f_int<R>(y);
    const R wb = y - (R)j;
    const int qa = g.j1 + g.jh * (j - 1), qb = g.j1 + g.jh * j;
    const R wa = (qa >= 0 && qa <= top) ? g.blat[qa] : (R)0, wbb = (qb >= 0 && qb <= top) ? g.blat[qb] : (R)0;
    const R wlat = wa + wb * (wbb - wa);
    e.area = (R)(g.rerth * g.rerth * wlat * g.dlon / DM<R>::dpr());
  }
}

// ---- Lambert conformal ------------------------------------------------------------------------
template <class R, int EXT>
__device__ __forceinline__ void lambert_ext(const GridP<R>& g, R fill, R rlat, R dlon, R dr, PtExt<R>& e) {
  const R dpr = DM<R>::dpr(), an = g.an;
  if (EXT >= EXT_ROT) {
    if (g.irot == 1) { e.crot = DM<R>::cos(an * dlon / dpr); e.srot = DM<R>::sin(an * dlon / dpr); }
    else { e.crot = IPB_DL(1.); e.srot = IPB_DL(0.); }
  }
  if (EXT >= EXT_ALL) {
    const R clat = DM<R>::cos(rlat / dpr);
    if (clat <= IPB_DL(0.) || dr <= IPB_DL(0.)) { e.xlon = e.xlat = e.ylon = e.ylat = fill; e.area = fill; }
    else {
      const R ca = DM<R>::cos(an * dlon / dpr), sa = DM<R>::sin(an * dlon / dpr);
      e.xlon = g.h * ca * an / dpr * dr / g.dxs;
      e.xlat = -(g.h * sa * an / dpr * dr / g.dxs / clat);
      e.ylon = g.h * sa * an / dpr * dr / g.dys;
      e.ylat = g.h * ca * an / dpr * dr / g.dys / clat;
      e.area = g.rerth * g.rerth * (clat * clat) * DM<R>::abs(g.dxs) * DM<R>::abs(g.dys) / ((an * dr) * (an * dr));
    }
  }
}

// ---- polar stereographic ----------------------------------------------------------------------
template <class R, int EXT>
__device__ __forceinline__ void polar_ext(const GridP<R>& g, R rlon, R rlat, R dr2, bool sphere, PtExt<R>& e) {
  const R dpr = DM<R>::dpr();
  if (EXT >= EXT_ROT) {
    if (g.irot == 1) { e.crot = g.h * DM<R>::cos((rlon - g.orient) / dpr); e.srot = DM<R>::sin((rlon - g.orient) / dpr); }
    else { e.crot = IPB_DL(1.); e.srot = IPB_DL(0.); }
  }
  if (EXT >= EXT_ALL && sphere) {
    const R cl = DM<R>::cos((rlon - g.orient) / dpr), sl = DM<R>::sin((rlon - g.orient) / dpr);
    if (dr2 < g.de2 * IPB_DL(1.E-6)) {
      const R de = DM<R>::sqrt(g.de2);
      e.xlon = IPB_DL(0.); e.xlat = -(sl / dpr * de / g.dxs / IPB_DL(2.));
      e.ylon = IPB_DL(0.); e.ylat = g.h * cl / dpr * de / g.dys / IPB_DL(2.);
      e.area = g.rerth * g.rerth * DM<R>::abs(g.dxs) * DM<R>::abs(g.dys) * IPB_DL(4.) / g.de2;
    } else {
      const R dr = DM<R>::sqrt(dr2), clat = DM<R>::cos(rlat / dpr);
      e.xlon = g.h * cl / dpr * dr / g.dxs; e.xlat = -(sl / dpr * dr / g.dxs / clat);
      e.ylon = sl / dpr * dr / g.dys; e.ylat = g.h * cl / dpr * dr / g.dys / clat;
      e.area = g.rerth * g.rerth * (clat * clat) * DM<R>::abs(g.dxs) * DM<R>::abs(g.dys) / dr2;
    }
  }
}

// ---- rotated lat-lon (B and E staggers), binary64 inside ----------------------------------------
struct RotPt { double clat, slat, clon, clatr, slatr; };

template <class R, int EXT>
__device__ __forceinline__ void rot_ext(const GridP<R>& g, R fill, R rlon, const RotPt& s, PtExt<R>& e) {
  const double dpr = (double)DM<R>::dpr();
  if (EXT >= EXT_ROT) {
    if (g.irot == 1) {
      if (s.clatr <= 0.) { e.crot = (R)(-::copysign(1., s.slatr * g.slat0)); e.srot = IPB_DL(0.); }
      else {
        const double slon = cr::sin(((double)rlon - g.rlon0) / dpr);
        e.crot = (R)((g.clat0 * s.clat + g.slat0 * s.slat * s.clon) / s.clatr);
        e.srot = (R)(g.slat0 * slon / s.clatr);
      }
    } else { e.crot = IPB_DL(1.); e.srot = IPB_DL(0.); }
  }
  if (EXT >= EXT_ALL) {
    if (s.clatr <= 0.) { e.xlon = e.xlat = e.ylon = e.ylat = fill; e.area = fill; return; }
    const double slon = cr::sin(((double)rlon - g.rlon0) / dpr);
    const double t1 = (g.clat0 * s.clat + g.slat0 * s.slat * s.clon) / s.clatr;
    const double t2 = g.slat0 * slon / s.clatr;
    const double xlo = t1 * s.clat / (g.dlons * s.clatr), xla = -t2 / (g.dlons * s.clatr);
    const double ylo = t2 * s.clat / g.dlats, yla = t1 / g.dlats;
    const double re = (double)g.rerth;
    if (g.type == P_ROTB) {
      e.xlon = (R)xlo; e.xlat = (R)xla; e.ylon = (R)ylo; e.ylat = (R)yla;
      e.area = (R)(2. * (re * re) * s.clatr * (g.dlons / dpr) * cr::sin(0.5 * g.dlats / dpr));
    } else {
      e.xlon = (R)(xlo - ylo); e.xlat = (R)(xla - yla); e.ylon = (R)(xlo + ylo); e.ylat = (R)(xla + yla);
      e.area = (R)(re * re * s.clatr * g.dlats * g.dlons) * IPB_DL(2.) / (DM<R>::dpr() * DM<R>::dpr());
    }
  }
}

}  // namespace pd

// =================================================================================================
// forward: (x,y) -> (rlon,rlat).  Returns false (and rlon=rlat=fill) outside the grid window.
// =================================================================================================
template <class R, int EXT>
__device__ bool proj_fwd(const GridP<R>& g, R fill, R x, R y, R& rlon, R& rlat, PtExt<R>& e) {
  const R dpr = DM<R>::dpr();
  if (EXT >= EXT_ROT) { e.crot = fill; e.srot = fill; }
  if (EXT >= EXT_ALL) { e.xlon = e.xlat = e.ylon = e.ylat = fill; e.area = fill; }
  switch (g.type) {
    case P_EQUID:
      if (!in_window(g, x, y)) break;
      rlon = pd::cyl_lon(g, x);
      rlat = DM<R>::vmin(DM<R>::vmax(g.rlat1 + g.dlat * (y - IPB_DL(1.)), IPB_DL(-90.)), IPB_DL(90.));
      pd::equid_ext<R, EXT>(g, rlat, e);
      return true;
    case P_MERC:
      if (!in_window(g, x, y)) break;
      rlon = pd::cyl_lon(g, x);
      rlat = IPB_DL(2.) * DM<R>::atan(DM<R>::exp(g.dphi * (y - g.ye))) * dpr - IPB_DL(90.);
      pd::merc_ext<R, EXT>(g, rlat, e);
      return true;
    case P_GAUSS: {
      if (!in_window(g, x, y)) break;
      rlon = pd::cyl_lon(g, x);
      const int j = f_int<R>(y), top = g.jg + 1;
      const R wb = y - (R)j;
      const int qa = g.j1 + g.jh * (j - 1), qb = g.j1 + g.jh * j;
      const R la = (qa >= 0 && qa <= top) ? g.alat[qa] : (R)0, lb = (qb >= 0 && qb <= top) ? g.alat[qb] : (R)0;
      rlat = la + wb * (lb - la);
      pd::gauss_ext<R, EXT>(g, y, e);
      return true;
    }
    case P_LAMBERT: {
      if (!in_window(g, x, y)) break;
      const R di = g.h * (x - g.xp) * g.dxs, dj = g.h * (y - g.yp) * g.dys;
      const R dr2 = di * di + dj * dj;
      if (dr2 < g.de2 * IPB_DL(1.E-6)) { rlon = IPB_DL(0.); rlat = g.h * IPB_DL(90.); }
      else {
        rlon = DM<R>::fmod(g.orient + IPB_DL(1.) / g.an * dpr * DM<R>::atan2(di, -dj) + IPB_DL(3600.), IPB_DL(360.));
        rlat = IPB_DL(2.) * dpr * DM<R>::atan(DM<R>::pow(g.de2 / dr2, g.antr)) - IPB_DL(90.);
      }
      if (EXT >= EXT_ROT) {
        const R dlon = DM<R>::fmod(rlon - g.orient + IPB_DL(180.) + IPB_DL(3600.), IPB_DL(360.)) - IPB_DL(180.);
        pd::lambert_ext<R, EXT>(g, fill, rlat, dlon, DM<R>::sqrt(dr2), e);
      }
      return true;
    }
    case P_POLAR: {
      if (!in_window(g, x, y)) break;
      const R di = (x - g.xp) * g.dxs, dj = (y - g.yp) * g.dys;
      if (!g.elliptical) {
        const R dr2 = di * di + dj * dj;
        if (dr2 < g.de2 * IPB_DL(1.E-6)) { rlon = IPB_DL(0.); rlat = g.h * IPB_DL(90.); }
        else {
          rlon = DM<R>::fmod(g.orient + g.h * dpr * DM<R>::atan2(di, -dj) + IPB_DL(3600.), IPB_DL(360.));
          rlat = g.h * dpr * DM<R>::asin((g.de2 - dr2) / (g.de2 + dr2));
        }
        pd::polar_ext<R, EXT>(g, rlon, rlat, dr2, true, e);
      } else {
        const R pi2 = DM<R>::pi() / IPB_DL(2.0);
        const R rho = DM<R>::sqrt(di * di + dj * dj);
        const R t = (rho * g.tc) / (g.rerth * g.mc);
        R along;
        if (DM<R>::abs(y - g.yp) < IPB_DL(0.01)) along = (di > IPB_DL(0.0)) ? g.orient + g.h * IPB_DL(90.0) : g.orient - g.h * IPB_DL(90.0);
        else {
          along = g.orient + g.h * DM<R>::atan(di / (-dj)) * dpr;
          if (dj > 0) along = along + IPB_DL(180.);
        }
        R prev = pi2 - IPB_DL(2.0) * DM<R>::atan(t), alat = 0;
        for (int it = 0; it < 10; it++) {
          const R s = DM<R>::sin(prev);
          alat = pi2 - IPB_DL(2.0) * DM<R>::atan(t * DM<R>::pow((IPB_DL(1.0) - g.ee * s) / (IPB_DL(1.0) + g.ee * s), g.e_over_2));
          if (DM<R>::abs(alat - prev) * dpr < IPB_DL(0.000001)) break;
          prev = alat;
        }
        rlat = g.h * alat * dpr;
        rlon = along;
        if (rlon < IPB_DL(0.0)) rlon = rlon + IPB_DL(360.);
        if (rlon > IPB_DL(360.0)) rlon = rlon - IPB_DL(360.0);
        pd::polar_ext<R, EXT>(g, rlon, rlat, (R)0, false, e);
      }
      return true;
    }
    case P_ROTB:
    case P_ROTE: {
      if (g.rerth < IPB_DL(0.)) break;
      const bool eg = g.type == P_ROTE;
      const double dprd = (double)dpr;
      R xf = x, yf = y;
      int IM = g.im, is1 = 0;
      if (eg) {
        IM = g.im * 2 - 1;
        is1 = (g.kscan == 0) ? (g.jm + 1) / 2 : g.jm / 2;
        xf = y + (x - (R)is1);
        yf = y - (x - (R)is1) + (R)g.kscan;
      }
      if (!in_window(g, xf, yf)) break;
      double rlonr, rlatr, hs;
      if (!eg) {
        rlonr = g.wbd + ((double)xf - 1.) * g.dlons;
        rlatr = g.sbd + ((double)yf - 1.) * g.dlats;
        hs = (rlonr <= 0.) ? -1.0 : 1.0;
      } else {
        const R mid = (R)((IM + 1) / 2);
        hs = g.hid * (double)f_sign1<R>(xf - mid);
        if (!g.is_grib2) {  // the two editions use different origins (ip_rot_equid_cylind_egrid_mod.F90:365-372)
          rlonr = (double)(xf - mid) * g.dlons;
          rlatr = (double)(yf - (R)((g.jm + 1) / 2)) * g.dlats;
        } else {
          rlonr = ((double)xf - 1.0) * g.dlons + g.wbd;
          rlatr = ((double)yf - 1.0) * g.dlats + g.sbd;
        }
      }
      pd::RotPt s;
      const double clonr = cr::cos(rlonr / dprd);
      s.slatr = cr::sin(rlatr / dprd);
      s.clatr = cr::cos(rlatr / dprd);
      s.slat = g.clat0 * s.slatr + g.slat0 * s.clatr * clonr;
      if (s.slat <= -1 || s.slat >= 1) {
        s.clat = 0.; s.clon = cr::cos(g.rlon0 / dprd);
        rlon = IPB_DL(0.); rlat = (s.slat <= -1) ? IPB_DL(-90.) : IPB_DL(90.);
      } else {
        s.clat = ::sqrt(1 - s.slat * s.slat);
        s.clon = (g.clat0 * s.clatr * clonr - g.slat0 * s.slatr) / s.clat;
        s.clon = ::fmin(::fmax(s.clon, -1.), 1.);
        rlon = (R)::fmod(g.rlon0 + hs * dprd * cr::acos(s.clon) + 3600, 360.);
        rlat = (R)(dprd * cr::asin(s.slat));
      }
      pd::rot_ext<R, EXT>(g, fill, rlon, s, e);
      return true;
    }
    default: break;
  }
  rlon = fill; rlat = fill;
  return false;
}

// =================================================================================================
// inverse: (rlon,rlat) -> (x,y).  Returns false (and x=y=fill) when the point is off the grid.
// =================================================================================================
template <class R, int EXT>
__device__ bool proj_inv(const GridP<R>& g, R fill, R rlon, R rlat, R& x, R& y, PtExt<R>& e) {
  const R dpr = DM<R>::dpr(), tiny = DM<R>::tiny();
  if (EXT >= EXT_ROT) { e.crot = fill; e.srot = fill; }
  if (EXT >= EXT_ALL) { e.xlon = e.xlat = e.ylon = e.ylat = fill; e.area = fill; }
  switch (g.type) {
    case P_EQUID:
      if (!(DM<R>::abs(rlon) <= IPB_DL(360.) && DM<R>::abs(rlat) <= IPB_DL(90.))) break;
      x = pd::cyl_x(g, rlon);
      y = IPB_DL(1.) + (rlat - g.rlat1) / g.dlat;
      if (!in_window(g, x, y)) break;
      pd::equid_ext<R, EXT>(g, rlat, e);
      return true;
    case P_MERC:
      if (!(DM<R>::abs(rlon) <= IPB_DL(360.) && DM<R>::abs(rlat) < IPB_DL(90.))) break;
      x = pd::cyl_x(g, rlon);
      y = g.ye + DM<R>::log(DM<R>::tan((rlat + IPB_DL(90.)) / IPB_DL(2.) / dpr)) / g.dphi;
      if (!in_window(g, x, y)) break;
      pd::merc_ext<R, EXT>(g, rlat, e);
      return true;
    case P_GAUSS: {
      if (!(DM<R>::abs(rlon) <= IPB_DL(360.) && DM<R>::abs(rlat) <= IPB_DL(90.))) break;
      x = pd::cyl_x(g, rlon);
      const int jg = g.jg;
      int ja = min(f_int<R>((R)(jg + 1) / IPB_DL(180.) * (IPB_DL(90.) - rlat)), jg);
      if (rlat > g.alat[ja]) ja = max(ja - 2, 0);
      if (rlat < g.alat[ja + 1]) ja = min(ja + 2, jg);
      if (rlat > g.alat[ja]) ja = ja - 1;
      if (rlat < g.alat[ja + 1]) ja = ja + 1;
      const R ya = (R)(1 + g.jh * (ja - g.j1)), yb = (R)(1 + g.jh * (ja + 1 - g.j1));
      const R wb = (g.alat[ja] - rlat) / (g.alat[ja] - g.alat[ja + 1]);
      y = ya + wb * (yb - ya);
      if (!in_window(g, x, y)) break;
      pd::gauss_ext<R, EXT>(g, y, e);
      return true;
    }
    case P_LAMBERT: {
      if (!(DM<R>::abs(rlon) < (IPB_DL(360.) + tiny) && DM<R>::abs(rlat) < (IPB_DL(90.) + tiny) &&
            DM<R>::abs(g.h * rlat + IPB_DL(90.)) > tiny)) break;
      const R dr = g.h * g.de * DM<R>::pow(DM<R>::tan((IPB_DL(90.) - rlat) / IPB_DL(2.) / dpr), g.an);
      const R dlon = DM<R>::fmod(rlon - g.orient + IPB_DL(180.) + IPB_DL(3600.), IPB_DL(360.)) - IPB_DL(180.);
      x = g.xp + g.h * DM<R>::sin(g.an * dlon / dpr) * dr / g.dxs;
      y = g.yp - g.h * DM<R>::cos(g.an * dlon / dpr) * dr / g.dys;
      if (!in_window(g, x, y)) break;
      pd::lambert_ext<R, EXT>(g, fill, rlat, dlon, dr, e);
      return true;
    }
    case P_POLAR: {
      if (!(DM<R>::abs(rlon) < (IPB_DL(360.) + tiny) && DM<R>::abs(rlat) < (IPB_DL(90.) + tiny) &&
            DM<R>::abs(g.h * rlat + IPB_DL(90.)) > tiny)) break;
      if (!g.elliptical) {
        const R dr = g.de * DM<R>::tan((IPB_DL(90.) - g.h * rlat) / IPB_DL(2.) / dpr);
        x = g.xp + g.h * DM<R>::sin((rlon - g.orient) / dpr) * dr / g.dxs;
        y = g.yp - DM<R>::cos((rlon - g.orient) / dpr) * dr / g.dys;
        if (!in_window(g, x, y)) break;
        pd::polar_ext<R, EXT>(g, rlon, rlat, dr * dr, true, e);
      } else {
        const R pi4 = DM<R>::pi() / IPB_DL(4.0);
        const R alat = g.h * rlat / dpr, along = (rlon - g.orient) / dpr;
        const R s = DM<R>::sin(alat);
        const R t = DM<R>::tan(pi4 - alat * IPB_DL(0.5)) / DM<R>::pow((IPB_DL(1.) - g.ee * s) / (IPB_DL(1.) + g.ee * s), g.e_over_2);
        const R rho = g.rerth * g.mc * t / g.tc;
        x = g.xp + rho * DM<R>::sin(g.h * along) / g.dxs;
        y = g.yp - rho * DM<R>::cos(g.h * along) / g.dys;
        if (!in_window(g, x, y)) break;
        pd::polar_ext<R, EXT>(g, rlon, rlat, (R)0, false, e);
      }
      return true;
    }
    case P_ROTB:
    case P_ROTE: {
      if (g.rerth < IPB_DL(0.)) break;
      if (!(DM<R>::abs(rlon) <= IPB_DL(360.) && DM<R>::abs(rlat) <= IPB_DL(90.))) break;
      const bool eg = g.type == P_ROTE;
      const double dprd = (double)dpr;
      pd::RotPt s;
      const double hs = ::copysign(1., ::fmod((double)rlon - g.rlon0 + 180 + 3600, 360.) - 180);
      s.clon = cr::cos(((double)rlon - g.rlon0) / dprd);
      s.slat = cr::sin((double)rlat / dprd);
      s.clat = cr::cos((double)rlat / dprd);
      s.slatr = g.clat0 * s.slat - g.slat0 * s.clat * s.clon;
      double rlonr, rlatr;
      if (s.slatr <= -1 || s.slatr >= 1) { s.clatr = 0.; rlonr = 0.; rlatr = (s.slatr <= -1) ? -90. : 90.; }
      else {
        s.clatr = ::sqrt(1 - s.slatr * s.slatr);
        double clonr = (g.clat0 * s.clat * s.clon + g.slat0 * s.slat) / s.clatr;
        clonr = ::fmin(::fmax(clonr, -1.), 1.);
        rlonr = hs * dprd * cr::acos(clonr);
        rlatr = dprd * cr::asin(s.slatr);
      }
      R xf, yf;
      if (eg && g.is_grib2) {
        const int IM = g.im * 2 - 1;
        xf = (R)((double)((IM + 1) / 2) + rlonr / g.dlons);
        yf = (R)((double)((g.jm + 1) / 2) + rlatr / g.dlats);
      } else {
        xf = (R)(((rlonr - g.wbd) / g.dlons) + 1.0);
        yf = (R)(((rlatr - g.sbd) / g.dlats) + 1.0);
      }
      if (!in_window(g, xf, yf)) break;
      if (eg) {
        const int is1 = (g.kscan == 0) ? (g.jm + 1) / 2 : g.jm / 2;
        x = (R)is1 + (xf - (yf - (R)g.kscan)) / IPB_DL(2.);
        y = (xf + (yf - (R)g.kscan)) / IPB_DL(2.);
      } else { x = xf; y = yf; }
      pd::rot_ext<R, EXT>(g, fill, rlon, s, e);
      return true;
    }
    default: break;
  }
  x = fill; y = fill;
  return false;
}

// "valid coordinate" test used by every interpolator: ABS(X-FILL) > TINY (e.g. bilinear_interp_mod.F90:171)
template <class R> __device__ __forceinline__ bool valid_xy(R x, R y) {
  const R fill = IPB_DL(-9999.);
  return DM<R>::abs(x - fill) > DM<R>::tiny() && DM<R>::abs(y - fill) > DM<R>::tiny();
}

// Great-circle vector transport rotation, always binary64 inside: src/movect.F90:25-68.
// The limits are default-kind literals widened to binary64 (so they differ between kinds).
template <class R> __device__ __forceinline__ void movect(R flat, R flon, R tlat, R tlon, R& crot, R& srot) {
  const double crdlim = (double)IPB_DL(0.9999999);
  const double dpr = 180. / (double)IPB_DL(3.14159265358979);
  const double ctlat = cr::cos((double)tlat / dpr), stlat = cr::sin((double)tlat / dpr);
  const double cflat = cr::cos((double)flat / dpr), sflat = cr::sin((double)flat / dpr);
  const double dl = (double)(flon - tlon) / dpr;
  const double cdlon = cr::cos(dl), sdlon = cr::sin(dl);
  const double crd = stlat * sflat + ctlat * cflat * cdlon;
  if (::fabs(crd) <= crdlim) {
    const double srd2rn = -1 / (1 - crd * crd);
    const double str = cflat * sdlon, ctr = cflat * stlat * cdlon - sflat * ctlat;
    const double sfr = ctlat * sdlon, cfr = ctlat * sflat * cdlon - stlat * cflat;
    crot = (R)(srd2rn * (ctr * cfr - str * sfr));
    srot = (R)(srd2rn * (ctr * sfr + str * cfr));
  } else {
    crot = (R)cdlon;
    srot = (R)(sdlon * stlat);
  }
}

}  // namespace ipb
