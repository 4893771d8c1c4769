// ipb200 — host-side grid decode: GRIB1 KGDS(200) / GRIB2 (igdtnum, igdtmpl) -> GridP<R>.
//
// Replaces the reference's descriptor -> factory -> init_grib1/init_grib2 chain
// (src/ip_grid_descriptor_mod.F90:74-106, src/ip_grid_factory_mod.F90:53-127 and the
// init_grib1/init_grib2 of each src/ip_*_grid_mod.F90) plus the set-up block at the head of
// every gdswzd_* routine.  Edition-specific code only extracts raw numbers ("Raw"); the
// per-projection finishing step is common to both editions.  Rounding is part of the
// contract: every operation below is done in R exactly where the reference does it in the
// default real kind, and in binary64 where the reference uses real64 (the rotated grids).
#pragma once
#include <cmath>
#include <cstdio>
#include <limits>
#include <type_traits>
#include <vector>

#include "ipb_crmath.h"
#include "ipb_grid.h"

namespace ipb {

// Default-kind real literal: a float literal for R=float, a double literal for R=double.
#define IPB_RC(x) (std::is_same<R, float>::value ? (R)(x##f) : (R)(x))

// Host math in kind R.  For R=float every transcendental is evaluated in binary64 and rounded
// once to binary32 — the same definition the device code uses; binary64 evaluations are correctly rounded (ipb_crmath.h).
template <class R> struct HM {
  static R sin(R x) { return (R)cr::sin((double)x); }
  static R cos(R x) { return (R)cr::cos((double)x); }
  static R tan(R x) { return (R)cr::tan((double)x); }
  static R asin(R x) { return (R)cr::asin((double)x); }
  static R log(R x) { return (R)cr::log((double)x); }
  static R pow(R a, R b) { return (R)cr::pow((double)a, (double)b); }
  static R sqrt(R x) { return (R)::sqrt((double)x); }  // exact for float via double
  static R fmod(R a, R b) { return (R)::fmod((double)a, (double)b); }  // exact
  static R abs(R x) { return x < 0 ? -x : x; }
  static R tiny() { return std::numeric_limits<R>::min(); }
  static i64 nint(R x) { return (i64)::llround((double)x); }
};

template <class R> struct GridHost {
  GridP<R> p;
  std::vector<R> alat, blat, ylat_row;  // Gaussian tables (host copies)
  std::vector<i64> key;                 // descriptor words: the plan-cache key
  bool ok = false;
};

namespace detail {

inline i64 bit(i64 word, i64 div) { return (word / div) % 2; }
template <class R> inline R pm1(i64 odd) { return odd ? IPB_RC(-1.) : IPB_RC(1.); }

// Raw numbers pulled out of a descriptor, in the units the reference converts them to.
template <class R> struct Raw {
  i64 im = 0, jm = 0;
  R rlat1 = 0, rlon1 = 0, rlat2 = 0, rlon2 = 0, rlati = 0, rlati1 = 0, rlati2 = 0, orient = 0, dx = 0, dy = 0, slat = 0;
  i64 scan = 0, iproj = 0, irot = 0, jgh = 0;
  R rerth = 0, e2 = 0;
};

// src/earth_radius_mod.F90:40-95
template <class R> void shape_of_earth(const i64* t, R& radius, R& e2) {
  auto p10 = [](i64 e) { i64 r = 1; while (e-- > 0) r *= 10; return r; };
  R flat, amaj, amin;
  switch (t[0]) {
    case 0: radius = IPB_RC(6367470.0); e2 = 0; return;
    case 1: radius = (R)t[2] / (R)p10(t[1]); e2 = 0; return;
    case 2: radius = IPB_RC(6378160.0); flat = IPB_RC(1.0) / IPB_RC(297.0); e2 = (IPB_RC(2.0) * flat) - (flat * flat); return;
    case 3:
      amaj = (R)t[4] / (R)p10(t[3]); amaj = amaj * IPB_RC(1000.0);
      amin = (R)t[6] / (R)p10(t[5]); amin = amin * IPB_RC(1000.0);
      e2 = IPB_RC(1.0) - ((amin * amin) / (amaj * amaj)); radius = amaj; return;
    case 4: radius = IPB_RC(6378137.0); flat = IPB_RC(1.0) / IPB_RC(298.2572); e2 = (IPB_RC(2.0) * flat) - (flat * flat); return;
    case 5: radius = IPB_RC(6378137.0); e2 = IPB_RC(0.00669437999013); return;
    case 6: radius = IPB_RC(6371229.0); e2 = 0; return;
    case 7:
      amaj = (R)t[4] / (R)p10(t[3]); amin = (R)t[6] / (R)p10(t[5]);
      e2 = IPB_RC(1.0) - ((amin * amin) / (amaj * amaj)); radius = amaj; return;
    case 8: radius = IPB_RC(6371200.0); e2 = 0; return;
    default: radius = IPB_RC(-9999.); e2 = IPB_RC(-9999.); return;
  }
}

template <class R> struct K {  // src/ip_constants_mod.F90:14-19
  static R pi() { return IPB_RC(3.14159265358979); }
  static R dpr() { return IPB_RC(180.0) / pi(); }
};

// i-direction increment and x wrap shared by lat-lon, Mercator and Gaussian grids
// (src/ip_equid_cylind_grid_mod.F90:68,77-79; ip_mercator_grid_mod.F90:75,86-87; ip_gaussian_grid_mod.F90:77,86-87)
template <class R> void cyl_x(GridP<R>& p, R rlon2) {
  R span = p.hi * (rlon2 - p.rlon1) - IPB_RC(1.) + IPB_RC(3600.);
  span = HM<R>::fmod(span, IPB_RC(360.)) + IPB_RC(1.);
  p.dlon = p.hi * span / (R)(p.im - 1);
  i64 full = HM<R>::nint(IPB_RC(360.) / HM<R>::abs(p.dlon));
  p.iwrap = (p.im < full) ? 0 : (int)full;
  p.xmin = 0;
  p.xmax = (R)(p.im + ((i64)p.im == full ? 2 : 1));
}
template <class R> void plain_window(GridP<R>& p, i64 xcount) {
  p.xmin = 0; p.xmax = (R)(xcount + 1); p.ymin = 0; p.ymax = (R)(p.jm + 1);
}
template <class R> void scan_bits(GridP<R>& p, i64 scan) { p.nscan = (int)bit(scan, 32); p.nscan_fp = p.nscan; p.kscan = 0; }

template <class R> void finish_equid(GridP<R>& p, const Raw<R>& r) {
  p.type = P_EQUID; p.rlat1 = r.rlat1; p.rlon1 = r.rlon1;
  p.hi = pm1<R>(bit(r.scan, 128));
  cyl_x(p, r.rlon2);
  p.dlat = (r.rlat2 - r.rlat1) / (R)(p.jm - 1);
  scan_bits(p, r.scan);
  // pole fold pivots; note the *signed* dlat (ip_equid_cylind_grid_mod.F90:84-95): N->S grids never fold
  p.jwrap1 = p.jwrap2 = 0;
  if (p.iwrap > 0 && p.iwrap % 2 == 0) {
    const R q = IPB_RC(90.) - IPB_RC(0.25) * p.dlat, h3 = IPB_RC(90.) - IPB_RC(0.75) * p.dlat;
    if (HM<R>::abs(r.rlat1) > q) p.jwrap1 = 2; else if (HM<R>::abs(r.rlat1) > h3) p.jwrap1 = 1;
    if (HM<R>::abs(r.rlat2) > q) p.jwrap2 = 2 * p.jm; else if (HM<R>::abs(r.rlat2) > h3) p.jwrap2 = 2 * p.jm + 1;
  }
  p.ymin = 0; p.ymax = (R)(p.jm + 1);
}

template <class R> void finish_mercator(GridP<R>& p, const Raw<R>& r) {
  const R dpr = K<R>::dpr();
  p.type = P_MERC; p.rlat1 = r.rlat1; p.rlon1 = r.rlon1;
  p.hi = pm1<R>(bit(r.scan, 128));
  const R hj = pm1<R>(1 - bit(r.scan, 64));
  cyl_x(p, r.rlon2);
  p.dphi = hj * r.dy / (p.rerth * HM<R>::cos(r.rlati / dpr));
  scan_bits(p, r.scan);
  p.jwrap1 = p.jwrap2 = 0;
  p.ye = IPB_RC(1.) - HM<R>::log(HM<R>::tan((r.rlat1 + IPB_RC(90.)) / IPB_RC(2.) / dpr)) / p.dphi;  // ip_mercator_grid_mod.F90:254
  p.ymin = 0; p.ymax = (R)(p.jm + 1);
}

template <class R> void finish_lambert(GridP<R>& p, const Raw<R>& r) {
  const R dpr = K<R>::dpr(), tiny = HM<R>::tiny();
  p.type = P_LAMBERT; p.rlat1 = r.rlat1; p.rlon1 = r.rlon1; p.orient = r.orient; p.irot = (int)r.irot;
  p.h = pm1<R>(r.iproj);
  p.dxs = r.dx * pm1<R>(bit(r.scan, 128));
  p.dys = r.dy * pm1<R>(1 - bit(r.scan, 64));
  scan_bits(p, r.scan);
  p.iwrap = p.jwrap1 = p.jwrap2 = 0;
  // cone constant, scale and pole position: ip_lambert_conf_grid_mod.F90:273-296
  if (HM<R>::abs(r.rlati1 - r.rlati2) < tiny) p.an = HM<R>::sin(r.rlati1 / dpr);
  else
    p.an = HM<R>::log(HM<R>::cos(r.rlati1 / dpr) / HM<R>::cos(r.rlati2 / dpr)) /
           HM<R>::log(HM<R>::tan((IPB_RC(90.) - r.rlati1) / IPB_RC(2.) / dpr) / HM<R>::tan((IPB_RC(90.) - r.rlati2) / IPB_RC(2.) / dpr));
  p.de = p.rerth * HM<R>::cos(r.rlati1 / dpr) * HM<R>::pow(HM<R>::tan((r.rlati1 + IPB_RC(90.)) / IPB_RC(2.) / dpr), p.an) / p.an;
  if (HM<R>::abs(p.h * r.rlat1 - IPB_RC(90.)) < tiny) { p.xp = 1; p.yp = 1; }
  else {
    const R dr = p.de / HM<R>::pow(HM<R>::tan((r.rlat1 + IPB_RC(90.)) / IPB_RC(2.) / dpr), p.an);
    const R dlon1 = HM<R>::fmod(r.rlon1 - r.orient + IPB_RC(180.) + IPB_RC(3600.), IPB_RC(360.)) - IPB_RC(180.);
    p.xp = IPB_RC(1.) - HM<R>::sin(p.an * dlon1 / dpr) * dr / p.dxs;
    p.yp = IPB_RC(1.) + HM<R>::cos(p.an * dlon1 / dpr) * dr / p.dys;
  }
  p.antr = IPB_RC(1.) / (IPB_RC(2.) * p.an);
  p.de2 = p.de * p.de;
  plain_window(p, p.im);
}

template <class R> void finish_polar(GridP<R>& p, const Raw<R>& r, bool grib1) {
  const R dpr = K<R>::dpr();
  p.type = P_POLAR; p.rlat1 = r.rlat1; p.rlon1 = r.rlon1; p.orient = r.orient; p.irot = (int)r.irot;
  p.h = pm1<R>(r.iproj);
  if (grib1 && HM<R>::abs(p.h + IPB_RC(1.)) < HM<R>::tiny()) p.orient = p.orient + IPB_RC(180.);  // ip_polar_stereo_grid_mod.F90:113
  p.dxs = r.dx * pm1<R>(bit(r.scan, 128));
  p.dys = r.dy * pm1<R>(1 - bit(r.scan, 64));
  scan_bits(p, r.scan);
  p.iwrap = p.jwrap1 = p.jwrap2 = 0;
  const R slatr = r.slat / dpr;
  if (!p.elliptical) {  // ip_polar_stereo_grid_mod.F90:300-305
    p.de = (IPB_RC(1.) + HM<R>::sin(slatr)) * p.rerth;
    const R dr = p.de * HM<R>::cos(r.rlat1 / dpr) / (IPB_RC(1.) + p.h * HM<R>::sin(r.rlat1 / dpr));
    p.xp = IPB_RC(1.) - p.h * HM<R>::sin((r.rlon1 - p.orient) / dpr) * dr / p.dxs;
    p.yp = IPB_RC(1.) + HM<R>::cos((r.rlon1 - p.orient) / dpr) * dr / p.dys;
    p.de2 = p.de * p.de;
  } else {  // :306-320
    const R pi4 = K<R>::pi() / IPB_RC(4.0);
    p.ee = HM<R>::sqrt(p.e2);
    p.e_over_2 = p.ee * IPB_RC(0.5);
    const R alat = p.h * r.rlat1 / dpr, along = (r.rlon1 - p.orient) / dpr;
    auto tfun = [&](R phi, R half) {
      return HM<R>::tan(pi4 - half) / HM<R>::pow((IPB_RC(1.) - p.ee * HM<R>::sin(phi)) / (IPB_RC(1.) + p.ee * HM<R>::sin(phi)), p.e_over_2);
    };
    const R t = tfun(alat, alat / IPB_RC(2.));
    p.tc = tfun(slatr, slatr / IPB_RC(2.));
    p.mc = HM<R>::cos(slatr) / HM<R>::sqrt(IPB_RC(1.0) - p.e2 * (HM<R>::sin(slatr) * HM<R>::sin(slatr)));
    const R rho = p.rerth * p.mc * t / p.tc;
    p.yp = IPB_RC(1.0) + rho * HM<R>::cos(p.h * along) / p.dys;
    p.xp = IPB_RC(1.0) - rho * HM<R>::sin(p.h * along) / p.dxs;
    p.de2 = 0;
  }
  plain_window(p, p.im);
}

// Gaussian latitudes: SPLAT idrt=4 (src/splat.F).  First guesses in the default kind, Newton
// iteration on the Legendre polynomial in binary64, results rounded to the default kind.
template <class R> void gaussian_latitudes(i64 jmax, std::vector<R>& slat, std::vector<R>& wlat) {
  static const char* bz_txt[50] = {
      "2.4048255577",  "5.5200781103",  "8.6537279129",  "11.7915344391", "14.9309177086", "18.0710639679", "21.2116366299",
      "24.3524715308", "27.4934791320", "30.6346064684", "33.7758202136", "36.9170983537", "40.0584257646", "43.1997917132",
      "46.3411883717", "49.4826098974", "52.6240518411", "55.7655107550", "58.9069839261", "62.0484691902", "65.1899648002",
      "68.3314693299", "71.4729816036", "74.6145006437", "77.7560256304", "80.8975558711", "84.0390907769", "87.1806298436",
      "90.3221726372", "93.4637187819", "96.6052679510", "99.7468198587", "102.888374254", "106.029930916", "109.171489649",
      "112.313050280", "115.454612653", "118.596176630", "121.737742088", "124.879308913", "128.020877005", "131.162446275",
      "134.304016638", "137.445588020", "140.587160352", "143.728733573", "146.870307625", "150.011882457", "153.153458019",
      "156.295034268"};
  auto bz = [&](int j) { return std::is_same<R, float>::value ? (R)strtof(bz_txt[j], nullptr) : (R)strtod(bz_txt[j], nullptr); };
  slat.assign(jmax, 0); wlat.assign(jmax, 0);
  const R pi = K<R>::pi();
  const R c = (IPB_RC(1.) - (IPB_RC(2.) / pi) * (IPB_RC(2.) / pi)) * IPB_RC(0.25);
  const i64 jh = jmax / 2, jhe = (jmax + 1) / 2;
  const R rr = IPB_RC(1.) / HM<R>::sqrt(((R)jmax + IPB_RC(0.5)) * ((R)jmax + IPB_RC(0.5)) + c);
  std::vector<double> z(jh), pk(jh), pkm1(jh), pkm2(jh);
  for (i64 j = 0; j < jh; j++) {
    const R arg = (j < 50) ? bz((int)j) * rr : (bz(49) + (R)(j + 1 - 50) * pi) * rr;
    z[j] = (double)(R)cr::cos((double)arg);
  }
  const double eps = 10. * std::numeric_limits<double>::epsilon();
  for (double spmax = 1.; spmax > eps;) {
    spmax = 0.;
    for (i64 j = 0; j < jh; j++) { pkm1[j] = 1.; pk[j] = z[j]; }
    for (i64 n = 2; n <= jmax; n++)
      for (i64 j = 0; j < jh; j++) {
        pkm2[j] = pkm1[j]; pkm1[j] = pk[j];
        pk[j] = ((double)(2 * n - 1) * z[j] * pkm1[j] - (double)(n - 1) * pkm2[j]) / (double)n;
      }
    for (i64 j = 0; j < jh; j++) {
      const double sp = pk[j] * (1. - z[j] * z[j]) / ((double)jmax * (pkm1[j] - z[j] * pk[j]));
      z[j] = z[j] - sp;
      if (std::fabs(sp) > spmax) spmax = std::fabs(sp);
    }
  }
  for (i64 j = 0; j < jh; j++) {
    const double d = (double)jmax * pkm1[j];
    slat[j] = (R)z[j];
    wlat[j] = (R)((2. * (1. - z[j] * z[j])) / (d * d));
    slat[jmax - 1 - j] = -slat[j];
    wlat[jmax - 1 - j] = wlat[j];
  }
  if (jhe > jh) {
    R w = IPB_RC(2.) / (R)(jmax * jmax);
    for (i64 n = 2; n <= jmax; n += 2) w = w * (R)(n * n) / (R)((n - 1) * (n - 1));
    slat[jhe - 1] = 0; wlat[jhe - 1] = w;
  }
}

template <class R> void finish_gaussian(GridHost<R>& gh, const Raw<R>& r) {
  GridP<R>& p = gh.p;
  const R dpr = K<R>::dpr();
  p.type = P_GAUSS; p.rlat1 = r.rlat1; p.rlon1 = r.rlon1;
  p.jg = (int)(r.jgh * 2);
  p.hi = pm1<R>(bit(r.scan, 128));
  p.jh = bit(r.scan, 64) ? -1 : 1;
  cyl_x(p, r.rlon2);
  scan_bits(p, r.scan);
  p.jwrap1 = p.jwrap2 = 0;
  if (p.iwrap > 0 && p.iwrap % 2 == 0 && p.jm == p.jg) { p.jwrap1 = 1; p.jwrap2 = 2 * p.jm + 1; }
  // latitude tables: ip_gaussian_grid_mod.F90:262-303
  const i64 jg = p.jg;
  std::vector<R> s, w;
  gaussian_latitudes<R>(jg, s, w);
  gh.alat.assign(jg + 2, 0); gh.blat.assign(jg + 2, 0); gh.ylat_row.assign(jg + 2, 0);
  for (i64 j = 1; j <= jg; j++) { gh.alat[j] = (R)(dpr * HM<R>::asin(s[j - 1])); gh.blat[j] = w[j - 1]; }
  gh.alat[0] = IPB_RC(180.) - gh.alat[1]; gh.alat[jg + 1] = -gh.alat[0];
  gh.blat[0] = -gh.blat[1]; gh.blat[jg + 1] = gh.blat[0];
  i64 j1 = 1;
  while (j1 < jg && r.rlat1 < (gh.alat[j1] + gh.alat[j1 + 1]) / IPB_RC(2.)) j1++;
  p.j1 = (int)j1;
  std::vector<R> byrow(jg + 2, 0);
  for (i64 j = 1; j <= jg; j++) { const i64 q = j1 + p.jh * (j - 1); if (q >= 0 && q <= jg + 1) byrow[q] = gh.alat[j]; }
  if (jg >= 2) {
    for (i64 j = 2; j <= jg - 1; j++) gh.ylat_row[j] = IPB_RC(2.0) / (byrow[j + 1] - byrow[j - 1]);
    gh.ylat_row[1] = IPB_RC(1.0) / (byrow[2] - byrow[1]); gh.ylat_row[0] = gh.ylat_row[1];
    gh.ylat_row[jg] = IPB_RC(1.0) / (byrow[jg] - byrow[jg - 1]); gh.ylat_row[jg + 1] = gh.ylat_row[jg];
  }
  p.ymin = IPB_RC(0.5); p.ymax = (R)p.jm + IPB_RC(0.5);
}

// Rotate a geographic point into the rotated system (used by the GRIB1 decode of both
// rotated grids: ip_rot_equid_cylind_grid_mod.F90:95-120, ip_rot_equid_cylind_egrid_mod.F90:116-128).
inline void to_rotated(double rlat, double rlon, double rlon0, double slat0, double clat0, double dpr, double& rlatr, double& rlonr) {
  const double slat = cr::sin(rlat / dpr), clat = cr::cos(rlat / dpr);
  const double hs = std::copysign(1., ::fmod(rlon - rlon0 + 180 + 3600, 360.) - 180);
  const double clon = cr::cos((rlon - rlon0) / dpr);
  const double slatr = clat0 * slat - slat0 * clat * clon;
  const double clatr = ::sqrt(1 - slatr * slatr);
  const double clonr = (clat0 * clat * clon + slat0 * slat) / clatr;
  rlatr = dpr * cr::asin(slatr);
  rlonr = hs * dpr * cr::acos(clonr);
}

}  // namespace detail

// ---- GRIB1 ------------------------------------------------------------------------------------
template <class R> GridHost<R> decode_grib1(const i64* k) {
  using namespace detail;
  GridHost<R> gh;
  GridP<R>& p = gh.p;
  p = GridP<R>();
  p.grid_num = k[0]; p.is_grib2 = 0; p.hi = 1; p.h = 1; p.jh = 1; p.j1 = 1; p.hid = 1;
  gh.key.assign(k, k + 200); gh.key.push_back(1);
  if (k[0] < 0) { p.type = P_STATION; gh.ok = true; return gh; }
  const R milli = IPB_RC(1.E-3);
  const double dprd = (double)K<R>::dpr();
  Raw<R> r;
  p.im = (int)k[1]; p.jm = (int)k[2];
  p.rerth = IPB_RC(6.3712E6); p.e2 = 0;
  r.rlat1 = (R)k[3] * milli; r.rlon1 = (R)k[4] * milli; r.scan = k[10];
  switch (k[0]) {
    case 0: r.rlat2 = (R)k[6] * milli; r.rlon2 = (R)k[7] * milli; finish_equid(p, r); break;
    case 1: r.rlon2 = (R)k[7] * milli; r.rlati = (R)k[8] * milli; r.dy = (R)k[12]; finish_mercator(p, r); break;
    case 3:
      r.irot = bit(k[5], 8); r.orient = (R)k[6] * milli; r.dx = (R)k[7]; r.dy = (R)k[8]; r.iproj = bit(k[9], 128);
      r.rlati1 = (R)k[11] * milli; r.rlati2 = (R)k[12] * milli; finish_lambert(p, r); break;
    case 4: r.rlon2 = (R)k[7] * milli; r.jgh = k[9]; finish_gaussian(gh, r); break;
    case 5:
      p.elliptical = bit(k[5], 64) == 1;
      if (p.elliptical) { p.rerth = IPB_RC(6.378137E6); p.e2 = IPB_RC(0.00669437999013); }
      r.irot = bit(k[5], 8); r.slat = IPB_RC(60.0); r.orient = (R)k[6] * milli; r.dx = (R)k[7]; r.dy = (R)k[8];
      r.iproj = bit(k[9], 128); finish_polar(p, r, true); break;
    case 203: {  // E stagger: increments from rotating the first point
      p.type = P_ROTE; p.nscan = (int)bit(k[10], 32); p.nscan_fp = 3; p.kscan = (int)bit(k[10], 256);
      p.irot = (int)bit(k[5], 8); p.hid = bit(k[10], 128) ? -1. : 1.;
      const double rlat1 = (double)k[3] * 1.E-3, rlon1 = (double)k[4] * 1.E-3, rlat0 = (double)k[6] * 1.E-3;
      p.rlon0 = (double)k[7] * 1.E-3;
      p.slat0 = cr::sin(rlat0 / dprd); p.clat0 = cr::cos(rlat0 / dprd);
      double rlatr, rlonr;
      to_rotated(rlat1, rlon1, p.rlon0, p.slat0, p.clat0, dprd, rlatr, rlonr);
      p.dlats = rlatr / (double)(-((p.jm - 1) / 2));
      p.dlons = rlonr / (double)(-(((p.im * 2 - 1) - 1) / 2));
      p.sbd = rlat1; p.wbd = rlon1; if (p.wbd > 180.0) p.wbd -= 360.0;
      plain_window(p, (i64)p.im * 2 - 1 + 1);
      break;
    }
    case 205: {  // B stagger: boundaries from rotating both corner points
      p.type = P_ROTB; scan_bits(p, k[10]); p.irot = (int)bit(k[5], 8);
      const double rlat1 = (double)k[3] * 1.E-3, rlon1 = (double)k[4] * 1.E-3, rlat0 = (double)k[6] * 1.E-3;
      const double rlat2 = (double)k[11] * 1.E-3, rlon2 = (double)k[12] * 1.E-3;
      p.rlon0 = (double)k[7] * 1.E-3;
      p.slat0 = cr::sin(rlat0 / dprd); p.clat0 = cr::cos(rlat0 / dprd);
      double nbd, ebd;
      to_rotated(rlat1, rlon1, p.rlon0, p.slat0, p.clat0, dprd, p.sbd, p.wbd);
      to_rotated(rlat2, rlon2, p.rlon0, p.slat0, p.clat0, dprd, nbd, ebd);
      p.dlats = (nbd - p.sbd) / (double)(R)(p.jm - 1);
      p.dlons = (ebd - p.wbd) / (double)(R)(p.im - 1);
      plain_window(p, p.im);
      break;
    }
    default: p.type = P_BAD; return gh;  // the reference leaves the grid unallocated
  }
  gh.ok = true;
  return gh;
}

// ---- GRIB2 ------------------------------------------------------------------------------------
template <class R> GridHost<R> decode_grib2(i64 num, const i64* t, i64 len) {
  using namespace detail;
  GridHost<R> gh;
  GridP<R>& p = gh.p;
  p = GridP<R>();
  p.grid_num = num; p.is_grib2 = 1; p.hi = 1; p.h = 1; p.jh = 1; p.j1 = 1; p.hid = 1;
  if (len < 0 || len > 65536) { p.type = P_BAD; return gh; }   // not a template length: nothing is read
  gh.key.assign(t, t + len); gh.key.push_back(num); gh.key.push_back(len); gh.key.push_back(2);
  if (num < 0) { p.type = P_STATION; gh.ok = true; return gh; }
  int type;
  switch (num) {
    case 0: type = P_EQUID; break;
    case 1: type = (len > 18 && bit(t[18], 8) != bit(t[18], 4)) ? P_ROTE : P_ROTB; break;  // ip_grid_factory_mod.F90:100-107
    case 10: type = P_MERC; break;
    case 20: type = P_POLAR; break;
    case 30: type = P_LAMBERT; break;
    case 40: type = P_GAUSS; break;
    case 32768: type = P_ROTE; break;
    case 32769: type = P_ROTB; break;
    default: p.type = P_BAD; return gh;  // the reference error-stops here
  }
  {  // the template must hold every word this decode reads (the reference would read past the caller's array)
    static const int need[] = {/*P_EQUID*/ 19, /*P_MERC*/ 19, /*P_LAMBERT*/ 20, /*P_GAUSS*/ 19, /*P_POLAR*/ 18, /*P_ROTB*/ 21, /*P_ROTE*/ 21};
    int n = 0;
    switch (type) {
      case P_EQUID: n = need[0]; break; case P_MERC: n = need[1]; break; case P_LAMBERT: n = need[2]; break;
      case P_GAUSS: n = need[3]; break; case P_POLAR: n = need[4]; break; case P_ROTB: n = need[5]; break; default: n = need[6];
    }
    if (len < n) { fprintf(stderr, "ipb200: grid definition template 3.%lld needs %d words, got %lld\n", (long long)num, n, (long long)len); p.type = P_BAD; return gh; }
  }
  const R micro = IPB_RC(1.0E-6), milli = IPB_RC(1.0E-3);
  const double dprd = (double)K<R>::dpr();
  shape_of_earth<R>(t, p.rerth, p.e2);
  p.im = (int)t[7]; p.jm = (int)t[8];
  Raw<R> r;
  auto scaled = [&](i64 v) {  // FLOAT(v)/FLOAT(iscale) in the default kind
    i64 s = t[9] * t[10];
    if (s == 0) s = 1000000;
    return (R)v / (R)s;
  };
  switch (type) {
    case P_EQUID:
      r.rlat1 = scaled(t[11]); r.rlon1 = scaled(t[12]); r.rlat2 = scaled(t[14]); r.rlon2 = scaled(t[15]); r.scan = t[18];
      finish_equid(p, r); break;
    case P_GAUSS:
      r.rlat1 = scaled(t[11]); r.rlon1 = scaled(t[12]); r.rlon2 = scaled(t[15]); r.jgh = t[17]; r.scan = t[18];
      finish_gaussian(gh, r); break;
    case P_MERC:
      r.rlat1 = (R)t[9] * micro; r.rlon1 = (R)t[10] * micro; r.rlon2 = (R)t[14] * micro; r.rlati = (R)t[12] * micro;
      r.scan = t[15]; r.dy = (R)t[18] * milli;
      finish_mercator(p, r); break;
    case P_LAMBERT:
      r.rlat1 = (R)t[9] * micro; r.rlon1 = (R)t[10] * micro; r.irot = bit(t[11], 8); r.orient = (R)t[13] * micro;
      r.dx = (R)t[14] * milli; r.dy = (R)t[15] * milli; r.iproj = bit(t[16], 128); r.scan = t[17];
      r.rlati1 = (R)t[18] * micro; r.rlati2 = (R)t[19] * micro;
      finish_lambert(p, r); break;
    case P_POLAR:
      p.elliptical = p.e2 > IPB_RC(0.0);
      r.rlat1 = (R)t[9] * micro; r.rlon1 = (R)t[10] * micro; r.irot = bit(t[11], 8);
      r.slat = (R)(t[12] < 0 ? -t[12] : t[12]) * micro; r.orient = (R)t[13] * micro;
      r.dx = (R)t[14] * milli; r.dy = (R)t[15] * milli; r.iproj = bit(t[16], 128); r.scan = t[17];
      finish_polar(p, r, false); break;
    case P_ROTB: {  // ip_rot_equid_cylind_grid_mod.F90:139-202
      p.type = P_ROTB; scan_bits(p, t[18]); p.irot = (int)bit(t[13], 8);
      const double rlat1 = (double)scaled(t[11]), rlon1 = (double)scaled(t[12]);
      const double rlat0 = (double)scaled(t[19]) + 90.0;
      p.rlon0 = (double)scaled(t[20]);
      const double rlat2 = (double)scaled(t[14]), rlon2 = (double)scaled(t[15]);
      p.slat0 = cr::sin(rlat0 / dprd); p.clat0 = cr::cos(rlat0 / dprd);
      p.wbd = rlon1; if (p.wbd > 180.0) p.wbd = p.wbd - 360.0;
      p.sbd = rlat1;
      p.dlats = (rlat2 - p.sbd) / (double)(R)(p.jm - 1);
      p.dlons = (rlon2 - p.wbd) / (double)(R)(p.im - 1);
      if (bit(t[18], 8) == 1) p.wbd = p.wbd + (0.5 * p.dlons);
      if (bit(t[18], 2) == 1) p.sbd = p.sbd + (0.5 * p.dlats);
      plain_window(p, p.im);
      break;
    }
    case P_ROTE: {  // ip_rot_equid_cylind_egrid_mod.F90:163-206; NSCAN comes from element 16 (sic)
      p.type = P_ROTE; p.nscan = (int)bit(t[15], 32); p.nscan_fp = 3; p.kscan = (int)bit(t[18], 8);
      p.irot = (int)bit(t[13], 8); p.hid = bit(t[18], 128) ? -1. : 1.;
      p.rlon0 = (double)scaled(t[20]);
      p.dlats = (double)scaled(t[17]);
      p.dlons = (double)scaled(t[16]) * 0.5;
      const double rlat0 = (double)scaled(t[19]) + 90.0;
      p.slat0 = cr::sin(rlat0 / dprd); p.clat0 = cr::cos(rlat0 / dprd);
      p.sbd = (double)scaled(t[11]);
      p.wbd = (double)scaled(t[12]); if (p.wbd > 180.0) p.wbd -= 360.0;
      plain_window(p, (i64)p.im * 2 - 1 + 1);
      break;
    }
  }
  gh.ok = true;
  return gh;
}

}  // namespace ipb
