// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
//
// CPU restatement of the NCEPLIBS-ip 5.0.0 geometry layer (grid decode, gdswzd forward /
// inverse map projections, field_pos, SPLAT idrt=4), written from the behaviour of the
// reference sources cited on each function (paths relative to /root/reference).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may include, link or execute anything in this directory.
//
// Parity pin: validated against the reference's own golden binaries
// (tests/data/baseline_data/...), see tests/test_oracle_golden.py.
//
// Arithmetic contract ("kind"): R = float reproduces the _4 build (every +,-,*,/ rounded to
// fp32, no FMA contraction: compile with -ffp-contract=off), R = double the _d/_8 builds.
// Transcendentals in the _4 kind are defined as "evaluate in binary64, round once to
// binary32" (SURVEY.md §7 hard part 1, option a); binary64 evaluations go through
// ip_oracle_libm.hpp (correctly rounded by default, plain glibc in the _glibc flavour) so that
// the CPU oracle and the CUDA kernels can agree bit-for-bit.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <type_traits>
#include <vector>

#include "ip_oracle_libm.hpp"

namespace ipo {

typedef long long i64;

// ---- literal helper: a Fortran default-real literal in kind R --------------------------
#define RL(x) (std::is_same<R, float>::value ? R(x##f) : R(x))

// ---- math in kind R ---------------------------------------------------------------------
template <class R> struct M;
template <> struct M<float> {
  static float sin(float x) { return (float)lm::sin((double)x); }
  static float cos(float x) { return (float)lm::cos((double)x); }
  static float tan(float x) { return (float)lm::tan((double)x); }
  static float asin(float x) { return (float)lm::asin((double)x); }
  static float acos(float x) { return (float)lm::acos((double)x); }
  static float atan(float x) { return (float)lm::atan((double)x); }
  static float atan2(float y, float x) { return (float)lm::atan2((double)y, (double)x); }
  static float log(float x) { return (float)lm::log((double)x); }
  static float exp(float x) { return (float)lm::exp((double)x); }
  static float pow(float a, float b) { return (float)lm::pow((double)a, (double)b); }
  static float sqrt(float x) { return ::sqrtf(x); }
  static float fmod(float a, float b) { return ::fmodf(a, b); }
  static float fabs(float x) { return ::fabsf(x); }
  static float floor(float x) { return ::floorf(x); }
  static float tiny() { return std::numeric_limits<float>::min(); }
  static float huge() { return std::numeric_limits<float>::max(); }
};
template <> struct M<double> {
  static double sin(double x) { return lm::sin(x); }
  static double cos(double x) { return lm::cos(x); }
  static double tan(double x) { return lm::tan(x); }
  static double asin(double x) { return lm::asin(x); }
  static double acos(double x) { return lm::acos(x); }
  static double atan(double x) { return lm::atan(x); }
  static double atan2(double y, double x) { return lm::atan2(y, x); }
  static double log(double x) { return lm::log(x); }
  static double exp(double x) { return lm::exp(x); }
  static double pow(double a, double b) { return lm::pow(a, b); }
  static double sqrt(double x) { return ::sqrt(x); }
  static double fmod(double a, double b) { return ::fmod(a, b); }
  static double fabs(double x) { return ::fabs(x); }
  static double floor(double x) { return ::floor(x); }
  static double tiny() { return std::numeric_limits<double>::min(); }
  static double huge() { return std::numeric_limits<double>::max(); }
};

// Fortran intrinsics
template <class R> inline i64 f_nint(R x) { return (i64)::llround((double)x); }  // half away from zero
template <class R> inline i64 f_int(R x) { return (i64)x; }                      // truncate
template <class R> inline i64 f_floor(R x) { return (i64)M<R>::floor(x); }
template <class R> inline R f_sign(R a, R b) { return std::signbit(b) ? -M<R>::fabs(a) : M<R>::fabs(a); }
template <class R> inline R f_min(R a, R b) { return a < b ? a : b; }
template <class R> inline R f_max(R a, R b) { return a > b ? a : b; }
inline i64 imod(i64 a, i64 p) { return a % p; }  // Fortran MOD on integers == C %
inline i64 ipow10(i64 e) { i64 r = 1; for (i64 i = 0; i < e; i++) r *= 10; return r; }

// src/ip_constants_mod.F90:14-19 (default-kind reals)
template <class R> struct C {
  static R pi() { return RL(3.14159265358979); }
  static R dpr() { return RL(180.0) / pi(); }
  static R pi2() { return pi() / RL(2.0); }
  static R pi4() { return pi() / RL(4.0); }
};

enum GridType { G_EQUID = 0, G_MERC, G_LAMBERT, G_GAUSS, G_POLAR, G_ROTB, G_ROTE, G_STATION, G_BAD };

// Decoded grid.  Mirrors ip_grid (src/ip_grid_mod.F90:54-87) plus each concrete type's
// members; the per-call set-up constants of each gdswzd_* are computed once here.
template <class R> struct Grid {
  int type = G_BAD;
  int is_grib2 = 0;
  i64 grid_num = 0;
  std::vector<i64> desc;  // raw descriptor words (equality == cache key)
  i64 im = 0, jm = 0;
  int nscan = 0, kscan = 0, nscan_field_pos = 0;
  i64 iwrap = 0, jwrap1 = 0, jwrap2 = 0;
  R rerth = 0, eccen_squared = 0;
  // lat-lon / mercator / gaussian
  R hi = 1, rlat1 = 0, rlon1 = 0, rlat2 = 0, rlon2 = 0, dlat = 0, dlon = 0, rlati = 0, dphi = 0, ye = 0;
  i64 jg = 0; int jscan = 0; i64 jh = 1, j1 = 1;
  std::vector<R> alat, blat, ylat_row;  // gaussian tables, index 0..jg+1
  // lambert / polar
  R rlati1 = 0, rlati2 = 0, orient = 0, dxs = 0, dys = 0, h = 1, slatr = 0;
  int irot = 0; bool elliptical = false;
  R an = 0, de = 0, de2 = 0, xp = 0, yp = 0, antr = 0;      // lambert & polar-sphere
  R e = 0, e_over_2 = 0, tc = 0, mc = 0;                     // polar ellipsoid
  // rotated (always real64 inside)
  double clat0 = 0, slat0 = 0, rlon0 = 0, dlats = 0, dlons = 0, wbd = 0, sbd = 0, rlat1d = 0, rlon1d = 0, hid = 1;
  // windows
  R xmin = 0, xmax = 0, ymin = 0, ymax = 0;
};

// ---- src/earth_radius_mod.F90:40-95 ------------------------------------------------------
template <class R> void earth_radius(const i64* t, R& radius, R& e2) {
  R flat, major_axis, minor_axis;
  switch (t[0]) {
    case 0: radius = RL(6367470.0); e2 = RL(0.0); break;
    case 1: radius = R(t[2]) / R(ipow10(t[1])); e2 = RL(0.0); break;
    case 2: radius = RL(6378160.0); flat = RL(1.0) / RL(297.0); e2 = (RL(2.0) * flat) - (flat * flat); break;
    case 3:
      major_axis = R(t[4]) / R(ipow10(t[3])); major_axis = major_axis * RL(1000.0);
      minor_axis = R(t[6]) / R(ipow10(t[5])); minor_axis = minor_axis * RL(1000.0);
      e2 = RL(1.0) - ((minor_axis * minor_axis) / (major_axis * major_axis)); radius = major_axis; break;
    case 4: radius = RL(6378137.0); flat = RL(1.0) / RL(298.2572); e2 = (RL(2.0) * flat) - (flat * flat); break;
    case 5: radius = RL(6378137.0); e2 = RL(0.00669437999013); break;
    case 6: radius = RL(6371229.0); e2 = RL(0.0); break;
    case 7:
      major_axis = R(t[4]) / R(ipow10(t[3])); minor_axis = R(t[6]) / R(ipow10(t[5]));
      e2 = RL(1.0) - ((minor_axis * minor_axis) / (major_axis * major_axis)); radius = major_axis; break;
    case 8: radius = RL(6371200.0); e2 = RL(0.0); break;
    default: radius = RL(-9999.); e2 = RL(-9999.); break;
  }
}

// ---- src/splat.F, IDRT=4 branch (Gaussian latitudes) --------------------------------------
// R, PI, C and BZ are default REAL; the Newton iteration runs in real(15,45) = binary64.
template <class R> void splat4(i64 jmax, std::vector<R>& slat, std::vector<R>& wlat) {
  static const double bzd[50] = {
      2.4048255577,  5.5200781103,  8.6537279129,  11.7915344391, 14.9309177086, 18.0710639679, 21.2116366299,
      24.3524715308, 27.4934791320, 30.6346064684, 33.7758202136, 36.9170983537, 40.0584257646, 43.1997917132,
      46.3411883717, 49.4826098974, 52.6240518411, 55.7655107550, 58.9069839261, 62.0484691902, 65.1899648002,
      68.3314693299, 71.4729816036, 74.6145006437, 77.7560256304, 80.8975558711, 84.0390907769, 87.1806298436,
      90.3221726372, 93.4637187819, 96.6052679510, 99.7468198587, 102.888374254, 106.029930916, 109.171489649,
      112.313050280, 115.454612653, 118.596176630, 121.737742088, 124.879308913, 128.020877005, 131.162446275,
      134.304016638, 137.445588020, 140.587160352, 143.728733573, 146.870307625, 150.011882457, 153.153458019,
      156.295034268};
  const int JZ = 50;
  slat.assign(jmax, 0); wlat.assign(jmax, 0);
  const R PI = RL(3.14159265358979);
  const R Cc = (RL(1.) - (RL(2.) / PI) * (RL(2.) / PI)) * RL(0.25);
  i64 jh = jmax / 2, jhe = (jmax + 1) / 2;
  R r = RL(1.) / M<R>::sqrt((R(jmax) + RL(0.5)) * (R(jmax) + RL(0.5)) + Cc);
  std::vector<double> slatd(jh), pk(jh), pkm1(jh), pkm2(jh);
  for (i64 j = 1; j <= std::min<i64>(jh, JZ); j++) slatd[j - 1] = (double)M<R>::cos(R(bzd[j - 1]) * r);
  for (i64 j = JZ + 1; j <= jh; j++) slatd[j - 1] = (double)M<R>::cos((R(bzd[JZ - 1]) + R(j - JZ) * PI) * r);
  const double eps = 10. * std::numeric_limits<double>::epsilon();
  double spmax = 1.;
  while (spmax > eps) {
    spmax = 0.;
    for (i64 j = 0; j < jh; j++) { pkm1[j] = 1.; pk[j] = slatd[j]; }
    for (i64 n = 2; n <= jmax; n++)
      for (i64 j = 0; j < jh; j++) {
        pkm2[j] = pkm1[j]; pkm1[j] = pk[j];
        pk[j] = ((double)(2 * n - 1) * slatd[j] * pkm1[j] - (double)(n - 1) * pkm2[j]) / (double)n;
      }
    for (i64 j = 0; j < jh; j++) {
      double sp = pk[j] * (1. - slatd[j] * slatd[j]) / ((double)jmax * (pkm1[j] - slatd[j] * pk[j]));
      slatd[j] = slatd[j] - sp;
      spmax = std::max(spmax, std::fabs(sp));
    }
  }
  for (i64 j = 0; j < jh; j++) {
    slat[j] = R(slatd[j]);
    double d = (double)jmax * pkm1[j];
    wlat[j] = R((2. * (1. - slatd[j] * slatd[j])) / (d * d));
    slat[jmax - 1 - j] = -slat[j];
    wlat[jmax - 1 - j] = wlat[j];
  }
  if (jhe > jh) {
    slat[jhe - 1] = 0;
    R w = RL(2.) / R(jmax * jmax);
    for (i64 n = 2; n <= jmax; n += 2) w = w * R(n * n) / R((n - 1) * (n - 1));
    wlat[jhe - 1] = w;
  }
}

// ---- shared decode pieces ------------------------------------------------------------------
// dlon expression shared by lat-lon / mercator / gaussian (e.g. ip_equid_cylind_grid_mod.F90:68)
template <class R> R dlon_of(R hi, R rlon1, R rlon2, i64 im) {
  R t = hi * (rlon2 - rlon1) - RL(1.) + RL(3600.);
  t = M<R>::fmod(t, RL(360.)) + RL(1.);
  return hi * t / R(im - 1);
}
template <class R> void iwrap_of(Grid<R>& g) {
  g.iwrap = f_nint<R>(RL(360.) / M<R>::fabs(g.dlon));
  if (g.im < g.iwrap) g.iwrap = 0;
}
template <class R> R sgn_pow(i64 n) { return (n % 2) ? RL(-1.) : RL(1.); }  // (-1.)**n

// ip_equid_cylind_grid_mod.F90:80-95 / 138-152 (pole fold flags; signed dlat)
template <class R> void equid_jwrap(Grid<R>& g) {
  g.jwrap1 = 0; g.jwrap2 = 0;
  if (g.iwrap > 0 && g.iwrap % 2 == 0) {
    if (M<R>::fabs(g.rlat1) > RL(90.) - RL(0.25) * g.dlat) g.jwrap1 = 2;
    else if (M<R>::fabs(g.rlat1) > RL(90.) - RL(0.75) * g.dlat) g.jwrap1 = 1;
    if (M<R>::fabs(g.rlat2) > RL(90.) - RL(0.25) * g.dlat) g.jwrap2 = 2 * g.jm;
    else if (M<R>::fabs(g.rlat2) > RL(90.) - RL(0.75) * g.dlat) g.jwrap2 = 2 * g.jm + 1;
  }
}

// Per-projection set-up constants (first lines of each gdswzd_* routine).
template <class R> void setup_equid(Grid<R>& g) {
  g.xmin = 0; g.xmax = R(g.im + 1);
  if (g.im == f_nint<R>(RL(360.) / M<R>::fabs(g.dlon))) g.xmax = R(g.im + 2);
  g.ymin = 0; g.ymax = R(g.jm + 1);
}
template <class R> void setup_mercator(Grid<R>& g) {  // ip_mercator_grid_mod.F90:254-260
  const R dpr = C<R>::dpr();
  g.ye = RL(1.) - M<R>::log(M<R>::tan((g.rlat1 + RL(90.)) / RL(2.) / dpr)) / g.dphi;
  setup_equid(g);
}
template <class R> void setup_lambert(Grid<R>& g) {  // ip_lambert_conf_grid_mod.F90:273-296
  const R dpr = C<R>::dpr(), tiny = M<R>::tiny();
  if (M<R>::fabs(g.rlati1 - g.rlati2) < tiny) {
    g.an = M<R>::sin(g.rlati1 / dpr);
  } else {
    g.an = M<R>::log(M<R>::cos(g.rlati1 / dpr) / M<R>::cos(g.rlati2 / dpr)) /
           M<R>::log(M<R>::tan((RL(90.) - g.rlati1) / RL(2.) / dpr) / M<R>::tan((RL(90.) - g.rlati2) / RL(2.) / dpr));
  }
  g.de = g.rerth * M<R>::cos(g.rlati1 / dpr) * M<R>::pow(M<R>::tan((g.rlati1 + RL(90.)) / RL(2.) / dpr), g.an) / g.an;
  if (M<R>::fabs(g.h * g.rlat1 - RL(90.)) < tiny) {
    g.xp = 1; g.yp = 1;
  } else {
    R dr = g.de / M<R>::pow(M<R>::tan((g.rlat1 + RL(90.)) / RL(2.) / dpr), g.an);
    R dlon1 = M<R>::fmod(g.rlon1 - g.orient + RL(180.) + RL(3600.), RL(360.)) - RL(180.);
    g.xp = RL(1.) - M<R>::sin(g.an * dlon1 / dpr) * dr / g.dxs;
    g.yp = RL(1.) + M<R>::cos(g.an * dlon1 / dpr) * dr / g.dys;
  }
  g.antr = RL(1.) / (RL(2.) * g.an);
  g.de2 = g.de * g.de;
  g.xmin = 0; g.xmax = R(g.im + 1); g.ymin = 0; g.ymax = R(g.jm + 1);
}
template <class R> void setup_polar(Grid<R>& g) {  // ip_polar_stereo_grid_mod.F90:300-328
  const R dpr = C<R>::dpr();
  if (!g.elliptical) {
    g.de = (RL(1.) + M<R>::sin(g.slatr)) * g.rerth;
    R dr = g.de * M<R>::cos(g.rlat1 / dpr) / (RL(1.) + g.h * M<R>::sin(g.rlat1 / dpr));
    g.xp = RL(1.) - g.h * M<R>::sin((g.rlon1 - g.orient) / dpr) * dr / g.dxs;
    g.yp = RL(1.) + M<R>::cos((g.rlon1 - g.orient) / dpr) * dr / g.dys;
    g.de2 = g.de * g.de;
  } else {
    const R e2 = g.eccen_squared;
    g.e = M<R>::sqrt(e2);
    g.e_over_2 = g.e * RL(0.5);
    R alat = g.h * g.rlat1 / dpr;
    R along = (g.rlon1 - g.orient) / dpr;
    R t = M<R>::tan(C<R>::pi4() - alat / RL(2.)) /
          M<R>::pow((RL(1.) - g.e * M<R>::sin(alat)) / (RL(1.) + g.e * M<R>::sin(alat)), g.e_over_2);
    g.tc = M<R>::tan(C<R>::pi4() - g.slatr / RL(2.)) /
           M<R>::pow((RL(1.) - g.e * M<R>::sin(g.slatr)) / (RL(1.) + g.e * M<R>::sin(g.slatr)), g.e_over_2);
    g.mc = M<R>::cos(g.slatr) / M<R>::sqrt(RL(1.0) - e2 * (M<R>::sin(g.slatr) * M<R>::sin(g.slatr)));
    R rho = g.rerth * g.mc * t / g.tc;
    g.yp = RL(1.0) + rho * M<R>::cos(g.h * along) / g.dys;
    g.xp = RL(1.0) - rho * M<R>::sin(g.h * along) / g.dxs;
    g.de2 = 0;  // never set on this branch in the reference (only used by Jacobian/area, not computed here)
  }
  g.xmin = 0; g.xmax = R(g.im + 1); g.ymin = 0; g.ymax = R(g.jm + 1);
}
template <class R> void setup_gaussian(Grid<R>& g) {  // ip_gaussian_grid_mod.F90:262-303
  const R dpr = C<R>::dpr();
  std::vector<R> s, w;
  splat4<R>(g.jg, s, w);
  g.alat.assign(g.jg + 2, 0); g.blat.assign(g.jg + 2, 0);
  for (i64 ja = 1; ja <= g.jg; ja++) { g.alat[ja] = R(dpr * M<R>::asin(s[ja - 1])); g.blat[ja] = w[ja - 1]; }
  g.alat[0] = RL(180.) - g.alat[1];
  g.alat[g.jg + 1] = -g.alat[0];
  g.blat[0] = -g.blat[1];
  g.blat[g.jg + 1] = g.blat[0];
  g.j1 = 1;
  while (g.j1 < g.jg && g.rlat1 < (g.alat[g.j1] + g.alat[g.j1 + 1]) / RL(2.)) g.j1++;
  // dy/dlat rows for the Jacobian (LMAP only)
  std::vector<R> aj(g.jg + 2, 0);
  for (i64 ja = 1; ja <= g.jg; ja++) {
    i64 idx = g.j1 + g.jh * (ja - 1);
    if (idx >= 0 && idx <= g.jg + 1) aj[idx] = g.alat[ja];
  }
  g.ylat_row.assign(g.jg + 2, 0);
  if (g.jg >= 2) {
    for (i64 ja = 2; ja <= g.jg - 1; ja++) g.ylat_row[ja] = RL(2.0) / (aj[ja + 1] - aj[ja - 1]);
    g.ylat_row[1] = RL(1.0) / (aj[2] - aj[1]);
    g.ylat_row[0] = g.ylat_row[1];
    g.ylat_row[g.jg] = RL(1.0) / (aj[g.jg] - aj[g.jg - 1]);
    g.ylat_row[g.jg + 1] = g.ylat_row[g.jg];
  }
  g.xmin = 0; g.xmax = R(g.im + 1);
  if (g.im == f_nint<R>(RL(360.) / M<R>::fabs(g.dlon))) g.xmax = R(g.im + 2);
  g.ymin = RL(0.5); g.ymax = R(g.jm) + RL(0.5);
}
template <class R> void setup_rot(Grid<R>& g) {
  if (g.type == G_ROTB) { g.xmin = 0; g.xmax = R(g.im + 1); g.ymin = 0; g.ymax = R(g.jm + 1); }
  else { g.xmin = 0; g.xmax = R((g.im * 2 - 1) + 2); g.ymin = 0; g.ymax = R(g.jm + 1); }
}

// ---- GRIB1 decode: init_grib1 of each grid type + src/ip_grid_factory_mod.F90:53-83 -------
template <class R> bool init_grid_grib1(Grid<R>& g, const i64* k /*200 words*/) {
  const R dpr = C<R>::dpr();
  g = Grid<R>();
  g.is_grib2 = 0; g.grid_num = k[0]; g.desc.assign(k, k + 200);
  const R m3 = RL(1.E-3);
  if (k[0] < 0) { g.type = G_STATION; return true; }
  switch (k[0]) {
    case 0: {  // ip_equid_cylind_grid_mod.F90:53-107
      g.type = G_EQUID; g.im = k[1]; g.jm = k[2];
      g.rlat1 = R(k[3]) * m3; g.rlon1 = R(k[4]) * m3; g.rlat2 = R(k[6]) * m3; g.rlon2 = R(k[7]) * m3;
      g.hi = sgn_pow<R>(imod(k[10] / 128, 2));
      g.dlon = dlon_of<R>(g.hi, g.rlon1, g.rlon2, g.im);
      g.dlat = (g.rlat2 - g.rlat1) / R(g.jm - 1);
      g.nscan = (int)imod(k[10] / 32, 2); g.nscan_field_pos = g.nscan; g.kscan = 0;
      iwrap_of(g); equid_jwrap(g);
      g.rerth = RL(6.3712E6); g.eccen_squared = RL(0.0);
      setup_equid(g); return true;
    }
    case 1: {  // ip_mercator_grid_mod.F90:52-98
      g.type = G_MERC; g.rerth = RL(6.3712E6); g.eccen_squared = RL(0.0);
      g.im = k[1]; g.jm = k[2];
      g.rlat1 = R(k[3]) * m3; g.rlon1 = R(k[4]) * m3; g.rlon2 = R(k[7]) * m3; g.rlati = R(k[8]) * m3;
      i64 iscan = imod(k[10] / 128, 2), jscan = imod(k[10] / 64, 2);
      R dy = R(k[12]);
      g.hi = sgn_pow<R>(iscan); R hj = sgn_pow<R>(1 - jscan);
      g.dlon = dlon_of<R>(g.hi, g.rlon1, g.rlon2, g.im);
      g.dphi = hj * dy / (g.rerth * M<R>::cos(g.rlati / dpr));
      g.nscan = (int)imod(k[10] / 32, 2); g.nscan_field_pos = g.nscan; g.kscan = 0;
      iwrap_of(g); g.jwrap1 = g.jwrap2 = 0;
      setup_mercator(g); return true;
    }
    case 3: {  // ip_lambert_conf_grid_mod.F90:60-110
      g.type = G_LAMBERT; g.rerth = RL(6.3712E6); g.eccen_squared = RL(0.0);
      g.im = k[1]; g.jm = k[2];
      g.rlat1 = R(k[3]) * m3; g.rlon1 = R(k[4]) * m3;
      g.irot = (int)imod(k[5] / 8, 2); g.orient = R(k[6]) * m3;
      R dx = R(k[7]), dy = R(k[8]);
      i64 iproj = imod(k[9] / 128, 2), iscan = imod(k[10] / 128, 2), jscan = imod(k[10] / 64, 2);
      g.rlati1 = R(k[11]) * m3; g.rlati2 = R(k[12]) * m3;
      g.h = sgn_pow<R>(iproj);
      g.dxs = dx * sgn_pow<R>(iscan); g.dys = dy * sgn_pow<R>(1 - jscan);
      g.nscan = (int)imod(k[10] / 32, 2); g.nscan_field_pos = g.nscan; g.kscan = 0;
      setup_lambert(g); return true;
    }
    case 4: {  // ip_gaussian_grid_mod.F90:60-107
      g.type = G_GAUSS; g.rerth = RL(6.3712E6); g.eccen_squared = RL(0.0);
      g.im = k[1]; g.jm = k[2];
      g.rlat1 = R(k[3]) * m3; g.rlon1 = R(k[4]) * m3; g.rlon2 = R(k[7]) * m3;
      g.jg = k[9] * 2;
      g.jscan = (int)imod(k[10] / 64, 2);
      g.hi = sgn_pow<R>(imod(k[10] / 128, 2)); g.jh = g.jscan ? -1 : 1;
      g.dlon = dlon_of<R>(g.hi, g.rlon1, g.rlon2, g.im);
      g.nscan = (int)imod(k[10] / 32, 2); g.nscan_field_pos = g.nscan; g.kscan = 0;
      iwrap_of(g); g.jwrap1 = g.jwrap2 = 0;
      if (g.iwrap > 0 && g.iwrap % 2 == 0 && g.jm == g.jg) { g.jwrap1 = 1; g.jwrap2 = 2 * g.jm + 1; }
      setup_gaussian(g); return true;
    }
    case 5: {  // ip_polar_stereo_grid_mod.F90:63-126
      g.type = G_POLAR;
      g.elliptical = imod(k[5] / 64, 2) == 1;
      if (!g.elliptical) { g.rerth = RL(6.3712E6); g.eccen_squared = 0; }
      else { g.rerth = RL(6.378137E6); g.eccen_squared = RL(0.00669437999013); }
      g.im = k[1]; g.jm = k[2];
      g.rlat1 = R(k[3]) * m3; g.rlon1 = R(k[4]) * m3;
      g.irot = (int)imod(k[5] / 8, 2);
      g.slatr = RL(60.0) / dpr;
      g.orient = R(k[6]) * m3;
      R dx = R(k[7]), dy = R(k[8]);
      i64 iproj = imod(k[9] / 128, 2), iscan = imod(k[10] / 128, 2), jscan = imod(k[10] / 64, 2);
      g.h = sgn_pow<R>(iproj);
      if (M<R>::fabs(g.h + RL(1.)) < M<R>::tiny()) g.orient = g.orient + RL(180.);
      g.dxs = dx * sgn_pow<R>(iscan); g.dys = dy * sgn_pow<R>(1 - jscan);
      g.nscan = (int)imod(k[10] / 32, 2); g.nscan_field_pos = g.nscan; g.kscan = 0;
      setup_polar(g); return true;
    }
    case 203: {  // ip_rot_equid_cylind_egrid_mod.F90:86-161 (all real64; DPR is the default-kind value widened)
      g.type = G_ROTE; g.rerth = R(6.3712E6); g.eccen_squared = RL(0.0);
      const double dprd = (double)dpr;
      i64 im = k[1], jm = k[2];
      g.nscan_field_pos = 3; g.nscan = (int)imod(k[10] / 32, 2);
      double rlat1 = (double)k[3] * 1.E-3, rlon1 = (double)k[4] * 1.E-3, rlat0 = (double)k[6] * 1.E-3, rlon0 = (double)k[7] * 1.E-3;
      g.irot = (int)imod(k[5] / 8, 2);
      g.kscan = (int)imod(k[10] / 256, 2);
      i64 iscan = imod(k[10] / 128, 2);
      double hi = iscan ? -1. : 1.;
      double slat1 = lm::sin(rlat1 / dprd), clat1 = lm::cos(rlat1 / dprd), slat0 = lm::sin(rlat0 / dprd), clat0 = lm::cos(rlat0 / dprd);
      double hs = std::copysign(1., ::fmod(rlon1 - rlon0 + 180 + 3600, 360.) - 180);
      double clon1 = lm::cos((rlon1 - rlon0) / dprd);
      double slatr = clat0 * slat1 - slat0 * clat1 * clon1;
      double clatr = ::sqrt(1 - slatr * slatr);
      double clonr = (clat0 * clat1 * clon1 + slat0 * slat1) / clatr;
      double rlatr = dprd * lm::asin(slatr);
      double rlonr = hs * dprd * lm::acos(clonr);
      g.dlats = rlatr / (double)(-((jm - 1) / 2));
      g.dlons = rlonr / (double)(-(((im * 2 - 1) - 1) / 2));
      g.im = im; g.jm = jm; g.rlon0 = rlon0; g.rlon1d = rlon1; g.rlat1d = rlat1; g.clat0 = clat0; g.slat0 = slat0; g.hid = hi;
      setup_rot(g); return true;
    }
    case 205: {  // ip_rot_equid_cylind_grid_mod.F90:73-137
      g.type = G_ROTB; g.rerth = R(6.3712E6); g.eccen_squared = 0;
      const double dprd = (double)dpr;
      double rlat1 = (double)k[3] * 1.E-3, rlon1 = (double)k[4] * 1.E-3, rlat0 = (double)k[6] * 1.E-3;
      g.rlon0 = (double)k[7] * 1.E-3;
      double rlat2 = (double)k[11] * 1.E-3, rlon2 = (double)k[12] * 1.E-3;
      g.irot = (int)imod(k[5] / 8, 2); g.im = k[1]; g.jm = k[2];
      double slat1 = lm::sin(rlat1 / dprd), clat1 = lm::cos(rlat1 / dprd);
      g.slat0 = lm::sin(rlat0 / dprd); g.clat0 = lm::cos(rlat0 / dprd);
      double hs = std::copysign(1., ::fmod(rlon1 - g.rlon0 + 180 + 3600, 360.) - 180);
      double clon1 = lm::cos((rlon1 - g.rlon0) / dprd);
      double slatr = g.clat0 * slat1 - g.slat0 * clat1 * clon1;
      double clatr = ::sqrt(1 - slatr * slatr);
      double clonr = (g.clat0 * clat1 * clon1 + g.slat0 * slat1) / clatr;
      double rlatr = dprd * lm::asin(slatr);
      double rlonr = hs * dprd * lm::acos(clonr);
      g.wbd = rlonr; g.sbd = rlatr;
      double slat2 = lm::sin(rlat2 / dprd), clat2 = lm::cos(rlat2 / dprd);
      double hs2 = std::copysign(1., ::fmod(rlon2 - g.rlon0 + 180 + 3600, 360.) - 180);
      double clon2 = lm::cos((rlon2 - g.rlon0) / dprd);
      slatr = g.clat0 * slat2 - g.slat0 * clat2 * clon2;
      clatr = ::sqrt(1 - slatr * slatr);
      clonr = (g.clat0 * clat2 * clon2 + g.slat0 * slat2) / clatr;
      double nbd = dprd * lm::asin(slatr);
      double ebd = hs2 * dprd * lm::acos(clonr);
      g.dlats = (nbd - g.sbd) / (double)R(g.jm - 1);
      g.dlons = (ebd - g.wbd) / (double)R(g.im - 1);
      g.nscan = (int)imod(k[10] / 32, 2); g.nscan_field_pos = g.nscan; g.kscan = 0;
      setup_rot(g); return true;
    }
    default: g.type = G_BAD; return false;  // reference: grid left unallocated (crash)
  }
}

// ---- GRIB2 decode: init_grib2 of each grid type + src/ip_grid_factory_mod.F90:89-127 -------
template <class R> bool init_grid_grib2(Grid<R>& g, i64 num, const i64* t, i64 len) {
  const R dpr = C<R>::dpr();
  g = Grid<R>();
  g.is_grib2 = 1; g.grid_num = num;
  g.desc.assign(t, t + len); g.desc.push_back(num); g.desc.push_back(len);
  if (num < 0) { g.type = G_STATION; return true; }
  int type = G_BAD;
  switch (num) {
    case 0: type = G_EQUID; break;
    case 1: type = (imod(t[18] / 8, 2) != imod(t[18] / 4, 2)) ? G_ROTE : G_ROTB; break;
    case 10: type = G_MERC; break;
    case 20: type = G_POLAR; break;
    case 30: type = G_LAMBERT; break;
    case 40: type = G_GAUSS; break;
    case 32768: type = G_ROTE; break;
    case 32769: type = G_ROTB; break;
    default: return false;  // reference: error stop
  }
  g.type = type;
  const R m6 = RL(1.0E-6), m3 = RL(1.0E-3);
  auto iscale_of = [&]() { i64 s = t[9] * t[10]; if (s == 0) s = 1000000; return s; };
  switch (type) {
    case G_EQUID: {  // ip_equid_cylind_grid_mod.F90:109-157
      g.im = t[7]; g.jm = t[8];
      i64 iscale = iscale_of();
      g.rlat1 = R(t[11]) / R(iscale); g.rlon1 = R(t[12]) / R(iscale);
      g.rlat2 = R(t[14]) / R(iscale); g.rlon2 = R(t[15]) / R(iscale);
      g.hi = sgn_pow<R>(imod(t[18] / 128, 2));
      g.dlon = dlon_of<R>(g.hi, g.rlon1, g.rlon2, g.im);
      g.dlat = (g.rlat2 - g.rlat1) / R(g.jm - 1);
      g.nscan = (int)imod(t[18] / 32, 2); g.nscan_field_pos = g.nscan; g.kscan = 0;
      iwrap_of(g); equid_jwrap(g);
      earth_radius<R>(t, g.rerth, g.eccen_squared);
      setup_equid(g); break;
    }
    case G_MERC: {  // ip_mercator_grid_mod.F90:100-138
      earth_radius<R>(t, g.rerth, g.eccen_squared);
      g.im = t[7]; g.jm = t[8];
      g.rlat1 = R(t[9]) * m6; g.rlon1 = R(t[10]) * m6; g.rlon2 = R(t[14]) * m6; g.rlati = R(t[12]) * m6;
      i64 iscan = imod(t[15] / 128, 2), jscan = imod(t[15] / 64, 2);
      R dy = R(t[18]) * m3;
      g.hi = sgn_pow<R>(iscan); R hj = sgn_pow<R>(1 - jscan);
      g.dlon = dlon_of<R>(g.hi, g.rlon1, g.rlon2, g.im);
      g.dphi = hj * dy / (g.rerth * M<R>::cos(g.rlati / dpr));
      g.jwrap1 = g.jwrap2 = 0; g.kscan = 0;
      g.nscan = (int)imod(t[15] / 32, 2); g.nscan_field_pos = g.nscan;
      iwrap_of(g);
      setup_mercator(g); break;
    }
    case G_LAMBERT: {  // ip_lambert_conf_grid_mod.F90:112-155
      earth_radius<R>(t, g.rerth, g.eccen_squared);
      g.im = t[7]; g.jm = t[8];
      g.rlat1 = R(t[9]) * m6; g.rlon1 = R(t[10]) * m6;
      g.irot = (int)imod(t[11] / 8, 2); g.orient = R(t[13]) * m6;
      R dx = R(t[14]) * m3, dy = R(t[15]) * m3;
      i64 iproj = imod(t[16] / 128, 2), iscan = imod(t[17] / 128, 2), jscan = imod(t[17] / 64, 2);
      g.rlati1 = R(t[18]) * m6; g.rlati2 = R(t[19]) * m6;
      g.h = sgn_pow<R>(iproj);
      g.dxs = dx * sgn_pow<R>(iscan); g.dys = dy * sgn_pow<R>(1 - jscan);
      g.nscan = (int)imod(t[17] / 32, 2); g.nscan_field_pos = g.nscan; g.kscan = 0;
      setup_lambert(g); break;
    }
    case G_POLAR: {  // ip_polar_stereo_grid_mod.F90:128-178
      earth_radius<R>(t, g.rerth, g.eccen_squared);
      g.elliptical = g.eccen_squared > RL(0.0);
      g.im = t[7]; g.jm = t[8];
      g.rlat1 = R(t[9]) * m6; g.rlon1 = R(t[10]) * m6;
      g.irot = (int)imod(t[11] / 8, 2);
      R slat = R(t[12] < 0 ? -t[12] : t[12]) * m6;
      g.slatr = slat / dpr;
      g.orient = R(t[13]) * m6;
      R dx = R(t[14]) * m3, dy = R(t[15]) * m3;
      i64 iproj = imod(t[16] / 128, 2), iscan = imod(t[17] / 128, 2), jscan = imod(t[17] / 64, 2);
      g.h = sgn_pow<R>(iproj);
      g.dxs = dx * sgn_pow<R>(iscan); g.dys = dy * sgn_pow<R>(1 - jscan);
      g.nscan = (int)imod(t[17] / 32, 2); g.nscan_field_pos = g.nscan; g.kscan = 0;
      setup_polar(g); break;
    }
    case G_GAUSS: {  // ip_gaussian_grid_mod.F90:109-155
      earth_radius<R>(t, g.rerth, g.eccen_squared);
      g.im = t[7]; g.jm = t[8];
      i64 iscale = iscale_of();
      g.rlat1 = R(t[11]) / R(iscale); g.rlon1 = R(t[12]) / R(iscale); g.rlon2 = R(t[15]) / R(iscale);
      g.jg = t[17] * 2;
      g.jscan = (int)imod(t[18] / 64, 2);
      g.hi = sgn_pow<R>(imod(t[18] / 128, 2)); g.jh = g.jscan ? -1 : 1;
      g.dlon = dlon_of<R>(g.hi, g.rlon1, g.rlon2, g.im);
      iwrap_of(g); g.jwrap1 = g.jwrap2 = 0;
      if (g.iwrap > 0 && g.iwrap % 2 == 0 && g.jm == g.jg) { g.jwrap1 = 1; g.jwrap2 = 2 * g.jm + 1; }
      g.nscan = (int)imod(t[18] / 32, 2); g.nscan_field_pos = g.nscan; g.kscan = 0;
      setup_gaussian(g); break;
    }
    case G_ROTB: {  // ip_rot_equid_cylind_grid_mod.F90:139-202 (FLOAT()/FLOAT() in default kind, then real64)
      earth_radius<R>(t, g.rerth, g.eccen_squared);
      const double dprd = (double)dpr;
      i64 i_offset_odd = imod(t[18] / 8, 2), j_offset = imod(t[18] / 2, 2);
      i64 iscale = iscale_of();
      double rlat1 = (double)(R(t[11]) / R(iscale)), rlon1 = (double)(R(t[12]) / R(iscale));
      double rlat0 = (double)(R(t[19]) / R(iscale)); rlat0 = rlat0 + 90.0;
      g.rlon0 = (double)(R(t[20]) / R(iscale));
      double rlat2 = (double)(R(t[14]) / R(iscale)), rlon2 = (double)(R(t[15]) / R(iscale));
      g.irot = (int)imod(t[13] / 8, 2); g.im = t[7]; g.jm = t[8];
      g.slat0 = lm::sin(rlat0 / dprd); g.clat0 = lm::cos(rlat0 / dprd);
      g.wbd = rlon1; if (g.wbd > 180.0) g.wbd = g.wbd - 360.0;
      g.sbd = rlat1;
      g.dlats = (rlat2 - g.sbd) / (double)R(g.jm - 1);
      g.dlons = (rlon2 - g.wbd) / (double)R(g.im - 1);
      if (i_offset_odd == 1) g.wbd = g.wbd + (0.5 * g.dlons);
      if (j_offset == 1) g.sbd = g.sbd + (0.5 * g.dlats);
      g.kscan = 0; g.nscan = (int)imod(t[18] / 32, 2); g.nscan_field_pos = g.nscan;
      setup_rot(g); break;
    }
    case G_ROTE: {  // ip_rot_equid_cylind_egrid_mod.F90:163-206 (NSCAN read from element 16, sic)
      earth_radius<R>(t, g.rerth, g.eccen_squared);
      const double dprd = (double)dpr;
      g.im = t[7]; g.jm = t[8];
      g.nscan = (int)imod(t[15] / 32, 2); g.nscan_field_pos = 3;
      i64 iscale = iscale_of();
      g.rlon0 = (double)(R(t[20]) / R(iscale));
      g.dlats = (double)(R(t[17]) / R(iscale));
      g.dlons = (double)(R(t[16]) / R(iscale)) * 0.5;
      g.irot = (int)imod(t[13] / 8, 2);
      g.kscan = (int)imod(t[18] / 8, 2);
      g.hid = imod(t[18] / 128, 2) ? -1. : 1.;
      double rlat0 = (double)(R(t[19]) / R(iscale)); rlat0 = rlat0 + 90.0;
      g.slat0 = lm::sin(rlat0 / dprd); g.clat0 = lm::cos(rlat0 / dprd);
      g.rlat1d = (double)(R(t[11]) / R(iscale)); g.rlon1d = (double)(R(t[12]) / R(iscale));
      setup_rot(g); break;
    }
  }
  return true;
}

// ---- src/ip_grid_mod.F90:199-249 -----------------------------------------------------------
template <class R> inline i64 field_pos(const Grid<R>& g, i64 i, i64 j) {
  i64 im = g.im, jm = g.jm, iwrap = g.iwrap, ii = i, jj = j;
  if (iwrap > 0) {
    ii = imod(i - 1 + iwrap, iwrap) + 1;
    if (j < 1 && g.jwrap1 > 0) { jj = g.jwrap1 - j; ii = imod(ii - 1 + iwrap / 2, iwrap) + 1; }
    else if (j > jm && g.jwrap2 > 0) { jj = g.jwrap2 - j; ii = imod(ii - 1 + iwrap / 2, iwrap) + 1; }
  }
  i64 fp = 0;
  int nscan = g.nscan_field_pos, kscan = g.kscan;
  if (nscan == 0) { if (ii >= 1 && ii <= im && jj >= 1 && jj <= jm) fp = ii + (jj - 1) * im; }
  else if (nscan == 1) { if (ii >= 1 && ii <= im && jj >= 1 && jj <= jm) fp = jj + (ii - 1) * jm; }
  else if (nscan == 2) {
    i64 is1 = (jm + 1 - kscan) / 2, iif = jj + (ii - is1), jjf = jj - (ii - is1) + kscan;
    if (iif >= 1 && iif <= 2 * im - 1 && jjf >= 1 && jjf <= jm) fp = (iif + (jjf - 1) * (2 * im - 1) + 1 - kscan) / 2;
  } else if (nscan == 3) {
    i64 is1 = (jm + 1 - kscan) / 2, iif = jj + (ii - is1), jjf = jj - (ii - is1) + kscan;
    if (iif >= 1 && iif <= 2 * im - 1 && jjf >= 1 && jjf <= jm) fp = (iif + 1) / 2 + (jjf - 1) * im;
  }
  return fp;
}

// Optional outputs of gdswzd.
template <class R> struct GdsOpt {
  R* crot = nullptr; R* srot = nullptr;
  R* xlon = nullptr; R* xlat = nullptr; R* ylon = nullptr; R* ylat = nullptr;
  R* area = nullptr;
  bool lrot() const { return crot && srot; }
  bool lmap() const { return xlon && xlat && ylon && ylat; }
  bool larea() const { return area != nullptr; }
};

// State of one rotated-grid transform needed by the rotation / Jacobian / area helpers.
struct RotState { double clat = 0, slat = 0, clon = 0, clatr = 0, slatr = 0; };

template <class R> inline void rot_vect_rot(const Grid<R>& g, const RotState& s, R rlon, R& crot, R& srot) {
  // ip_rot_equid_cylind_grid_mod.F90:507-530, ip_rot_equid_cylind_egrid_mod.F90:514-536
  const double dprd = (double)C<R>::dpr();
  if (g.irot == 1) {
    if (s.clatr <= 0.) { crot = R(-std::copysign(1., s.slatr * g.slat0)); srot = 0; }
    else {
      double slon = lm::sin(((double)rlon - g.rlon0) / dprd);
      crot = R((g.clat0 * s.clat + g.slat0 * s.slat * s.clon) / s.clatr);
      srot = R(g.slat0 * slon / s.clatr);
    }
  } else { crot = RL(1.); srot = RL(0.); }
}
template <class R> inline void rot_map_jacob(const Grid<R>& g, const RotState& s, R fill, R rlon, R& xlon, R& xlat, R& ylon, R& ylat) {
  const double dprd = (double)C<R>::dpr();
  if (s.clatr <= 0.) { xlon = xlat = ylon = ylat = fill; return; }
  double slon = lm::sin(((double)rlon - g.rlon0) / dprd);
  double term1 = (g.clat0 * s.clat + g.slat0 * s.slat * s.clon) / s.clatr;
  double term2 = g.slat0 * slon / s.clatr;
  double xlonf = term1 * s.clat / (g.dlons * s.clatr), xlatf = -term2 / (g.dlons * s.clatr);
  double ylonf = term2 * s.clat / g.dlats, ylatf = term1 / g.dlats;
  if (g.type == G_ROTB) { xlon = R(xlonf); xlat = R(xlatf); ylon = R(ylonf); ylat = R(ylatf); }
  else { xlon = R(xlonf - ylonf); xlat = R(xlatf - ylatf); ylon = R(xlonf + ylonf); ylat = R(xlatf + ylatf); }
}
template <class R> inline void rot_area(const Grid<R>& g, const RotState& s, R fill, R& area) {
  const double dprd = (double)C<R>::dpr();
  const double rerth = (double)g.rerth;
  if (s.clatr <= 0.) { area = fill; return; }
  if (g.type == G_ROTB) area = R(2. * (rerth * rerth) * s.clatr * (g.dlons / dprd) * lm::sin(0.5 * g.dlats / dprd));
  else area = R(rerth * rerth * s.clatr * g.dlats * g.dlons) * RL(2.) / (C<R>::dpr() * C<R>::dpr());
}

// ---- the projections: type-bound gdswzd of each grid (iopt = +1 forward, -1 inverse) -------
template <class R>
void gdswzd_proj(const Grid<R>& g, int iopt, i64 npts, R fill, R* xpts, R* ypts, R* rlon, R* rlat, i64& nret,
                 const GdsOpt<R>& o) {
  const R dpr = C<R>::dpr();
  const R tiny = M<R>::tiny();
  const bool lrot = o.lrot(), lmap = o.lmap(), larea = o.larea();
  if (g.type == G_STATION) { nret = npts; return; }  // ip_station_points_grid_mod.F90:75-92
  for (i64 n = 0; n < npts; n++) {
    if (o.crot) o.crot[n] = fill;
    if (o.srot) o.srot[n] = fill;
    if (o.xlon) o.xlon[n] = fill;
    if (o.xlat) o.xlat[n] = fill;
    if (o.ylon) o.ylon[n] = fill;
    if (o.ylat) o.ylat[n] = fill;
    if (o.area) o.area[n] = fill;
  }
  nret = 0;
  if ((g.type == G_ROTB || g.type == G_ROTE) && g.rerth < RL(0.)) {  // *_ERROR routines
    for (i64 n = 0; n < npts; n++) {
      if (iopt >= 0) { rlon[n] = fill; rlat[n] = fill; }
      if (iopt <= 0) { xpts[n] = fill; ypts[n] = fill; }
    }
    return;
  }
  const R xmin = g.xmin, xmax = g.xmax, ymin = g.ymin, ymax = g.ymax;
  auto inwin = [&](R x, R y) { return x >= xmin && x <= xmax && y >= ymin && y <= ymax; };
  const bool fwd = (iopt == 0 || iopt == 1), inv = (iopt == -1);
  if (!fwd && !inv) return;

  switch (g.type) {
    case G_EQUID: {  // ip_equid_cylind_grid_mod.F90:203-313
      auto extras = [&](i64 n) {
        if (lrot) { o.crot[n] = RL(1.0); o.srot[n] = RL(0.0); }
        if (lmap) { o.xlon[n] = RL(1.0) / g.dlon; o.xlat[n] = RL(0.); o.ylon[n] = RL(0.); o.ylat[n] = RL(1.0) / g.dlat; }
        if (larea) {
          R rlatu = f_min(f_max(rlat[n] + g.dlat / RL(2.), RL(-90.)), RL(90.));
          R rlatd = f_min(f_max(rlat[n] - g.dlat / RL(2.), RL(-90.)), RL(90.));
          R dslat = M<R>::sin(rlatu / dpr) - M<R>::sin(rlatd / dpr);
          o.area[n] = g.rerth * g.rerth * M<R>::fabs(dslat * g.dlon) / dpr;
        }
      };
      for (i64 n = 0; n < npts; n++) {
        if (fwd) {
          if (inwin(xpts[n], ypts[n])) {
            rlon[n] = M<R>::fmod(g.rlon1 + g.dlon * (xpts[n] - RL(1.)) + RL(3600.), RL(360.));
            rlat[n] = f_min(f_max(g.rlat1 + g.dlat * (ypts[n] - RL(1.)), RL(-90.)), RL(90.));
            nret++; extras(n);
          } else { rlon[n] = fill; rlat[n] = fill; }
        } else {
          if (M<R>::fabs(rlon[n]) <= RL(360.) && M<R>::fabs(rlat[n]) <= RL(90.)) {
            xpts[n] = RL(1.) + g.hi * M<R>::fmod(g.hi * (rlon[n] - g.rlon1) + RL(3600.), RL(360.)) / g.dlon;
            ypts[n] = RL(1.) + (rlat[n] - g.rlat1) / g.dlat;
            if (inwin(xpts[n], ypts[n])) { nret++; extras(n); }
            else { xpts[n] = fill; ypts[n] = fill; }
          } else { xpts[n] = fill; ypts[n] = fill; }
        }
      }
      break;
    }
    case G_MERC: {  // ip_mercator_grid_mod.F90:198-313
      auto extras = [&](i64 n) {
        if (lrot) { o.crot[n] = RL(1.0); o.srot[n] = RL(0.0); }
        if (lmap) {
          o.xlon[n] = RL(1.) / g.dlon; o.xlat[n] = RL(0.); o.ylon[n] = RL(0.);
          o.ylat[n] = RL(1.) / g.dphi / M<R>::cos(rlat[n] / dpr) / dpr;
        }
        if (larea) {
          R c = M<R>::cos(rlat[n] / dpr);
          o.area[n] = g.rerth * g.rerth * (c * c) * g.dphi * g.dlon / dpr;
        }
      };
      for (i64 n = 0; n < npts; n++) {
        if (fwd) {
          if (inwin(xpts[n], ypts[n])) {
            rlon[n] = M<R>::fmod(g.rlon1 + g.dlon * (xpts[n] - RL(1.)) + RL(3600.), RL(360.));
            rlat[n] = RL(2.) * M<R>::atan(M<R>::exp(g.dphi * (ypts[n] - g.ye))) * dpr - RL(90.);
            nret++; extras(n);
          } else { rlon[n] = fill; rlat[n] = fill; }
        } else {
          if (M<R>::fabs(rlon[n]) <= RL(360.) && M<R>::fabs(rlat[n]) < RL(90.)) {
            xpts[n] = RL(1.) + g.hi * M<R>::fmod(g.hi * (rlon[n] - g.rlon1) + RL(3600.), RL(360.)) / g.dlon;
            ypts[n] = g.ye + M<R>::log(M<R>::tan((rlat[n] + RL(90.)) / RL(2.) / dpr)) / g.dphi;
            if (inwin(xpts[n], ypts[n])) { nret++; extras(n); }
            else { xpts[n] = fill; ypts[n] = fill; }
          } else { xpts[n] = fill; ypts[n] = fill; }
        }
      }
      break;
    }
    case G_LAMBERT: {  // ip_lambert_conf_grid_mod.F90:219-370 (+ VECT_ROT :390, MAP_JACOB :427, GRID_AREA :468)
      const R an = g.an, h = g.h, dxs = g.dxs, dys = g.dys;
      auto extras = [&](i64 n, R dlon, R dr) {
        if (lrot) {
          if (g.irot == 1) { o.crot[n] = M<R>::cos(an * dlon / dpr); o.srot[n] = M<R>::sin(an * dlon / dpr); }
          else { o.crot[n] = RL(1.); o.srot[n] = RL(0.); }
        }
        if (lmap || larea) {
          R clat = M<R>::cos(rlat[n] / dpr);
          bool bad = (clat <= RL(0.) || dr <= RL(0.));
          if (lmap) {
            if (bad) { o.xlon[n] = o.xlat[n] = o.ylon[n] = o.ylat[n] = fill; }
            else {
              o.xlon[n] = h * M<R>::cos(an * dlon / dpr) * an / dpr * dr / dxs;
              o.xlat[n] = -h * M<R>::sin(an * dlon / dpr) * an / dpr * dr / dxs / clat;
              o.ylon[n] = h * M<R>::sin(an * dlon / dpr) * an / dpr * dr / dys;
              o.ylat[n] = h * M<R>::cos(an * dlon / dpr) * an / dpr * dr / dys / clat;
            }
          }
          if (larea) {
            if (bad) o.area[n] = fill;
            else o.area[n] = g.rerth * g.rerth * (clat * clat) * M<R>::fabs(dxs) * M<R>::fabs(dys) / ((an * dr) * (an * dr));
          }
        }
      };
      for (i64 n = 0; n < npts; n++) {
        if (fwd) {
          if (inwin(xpts[n], ypts[n])) {
            R di = h * (xpts[n] - g.xp) * dxs, dj = h * (ypts[n] - g.yp) * dys;
            R dr2 = di * di + dj * dj;
            R dr = M<R>::sqrt(dr2);
            if (dr2 < g.de2 * RL(1.E-6)) { rlon[n] = RL(0.); rlat[n] = h * RL(90.); }
            else {
              rlon[n] = M<R>::fmod(g.orient + RL(1.) / an * dpr * M<R>::atan2(di, -dj) + RL(3600.), RL(360.));
              rlat[n] = (RL(2.) * dpr * M<R>::atan(M<R>::pow(g.de2 / dr2, g.antr)) - RL(90.));
            }
            nret++;
            R dlon = M<R>::fmod(rlon[n] - g.orient + RL(180.) + RL(3600.), RL(360.)) - RL(180.);
            extras(n, dlon, dr);
          } else { rlon[n] = fill; rlat[n] = fill; }
        } else {
          if (M<R>::fabs(rlon[n]) < (RL(360.) + tiny) && M<R>::fabs(rlat[n]) < (RL(90.) + tiny) &&
              M<R>::fabs(h * rlat[n] + RL(90.)) > tiny) {
            R dr = h * g.de * M<R>::pow(M<R>::tan((RL(90.) - rlat[n]) / RL(2.) / dpr), an);
            R dlon = M<R>::fmod(rlon[n] - g.orient + RL(180.) + RL(3600.), RL(360.)) - RL(180.);
            xpts[n] = g.xp + h * M<R>::sin(an * dlon / dpr) * dr / dxs;
            ypts[n] = g.yp - h * M<R>::cos(an * dlon / dpr) * dr / dys;
            if (inwin(xpts[n], ypts[n])) { nret++; extras(n, dlon, dr); }
            else { xpts[n] = fill; ypts[n] = fill; }
          } else { xpts[n] = fill; ypts[n] = fill; }
        }
      }
      break;
    }
    case G_POLAR: {  // ip_polar_stereo_grid_mod.F90:239-461 (+ VECT_ROT :482, MAP_JACOB :520, GRID_AREA :563)
      const R h = g.h, dxs = g.dxs, dys = g.dys, orient = g.orient, de2 = g.de2;
      auto rot = [&](i64 n) {
        if (!lrot) return;
        if (g.irot == 1) { o.crot[n] = h * M<R>::cos((rlon[n] - orient) / dpr); o.srot[n] = M<R>::sin((rlon[n] - orient) / dpr); }
        else { o.crot[n] = RL(1.); o.srot[n] = RL(0.); }
      };
      auto jac_area = [&](i64 n, R dr2) {
        if (lmap) {
          if (dr2 < de2 * RL(1.E-6)) {
            R de = M<R>::sqrt(de2);
            o.xlon[n] = RL(0.);
            o.xlat[n] = -M<R>::sin((rlon[n] - orient) / dpr) / dpr * de / dxs / RL(2.);
            o.ylon[n] = RL(0.);
            o.ylat[n] = h * M<R>::cos((rlon[n] - orient) / dpr) / dpr * de / dys / RL(2.);
          } else {
            R dr = M<R>::sqrt(dr2), clat = M<R>::cos(rlat[n] / dpr);
            o.xlon[n] = h * M<R>::cos((rlon[n] - orient) / dpr) / dpr * dr / dxs;
            o.xlat[n] = -M<R>::sin((rlon[n] - orient) / dpr) / dpr * dr / dxs / clat;
            o.ylon[n] = M<R>::sin((rlon[n] - orient) / dpr) / dpr * dr / dys;
            o.ylat[n] = h * M<R>::cos((rlon[n] - orient) / dpr) / dpr * dr / dys / clat;
          }
        }
        if (larea) {
          if (dr2 < de2 * RL(1.E-6)) o.area[n] = g.rerth * g.rerth * M<R>::fabs(dxs) * M<R>::fabs(dys) * RL(4.) / de2;
          else { R clat = M<R>::cos(rlat[n] / dpr); o.area[n] = g.rerth * g.rerth * (clat * clat) * M<R>::fabs(dxs) * M<R>::fabs(dys) / dr2; }
        }
      };
      for (i64 n = 0; n < npts; n++) {
        if (fwd) {
          if (!inwin(xpts[n], ypts[n])) { rlon[n] = fill; rlat[n] = fill; continue; }
          if (!g.elliptical) {
            R di = (xpts[n] - g.xp) * dxs, dj = (ypts[n] - g.yp) * dys;
            R dr2 = di * di + dj * dj;
            if (dr2 < de2 * RL(1.E-6)) { rlon[n] = RL(0.); rlat[n] = h * RL(90.); }
            else {
              rlon[n] = M<R>::fmod(orient + h * dpr * M<R>::atan2(di, -dj) + RL(3600.), RL(360.));
              rlat[n] = h * dpr * M<R>::asin((de2 - dr2) / (de2 + dr2));
            }
            nret++; rot(n); jac_area(n, dr2);
          } else {
            R di = (xpts[n] - g.xp) * dxs, dj = (ypts[n] - g.yp) * dys;
            R rho = M<R>::sqrt(di * di + dj * dj);
            R t = (rho * g.tc) / (g.rerth * g.mc);
            R along;
            if (M<R>::fabs(ypts[n] - g.yp) < RL(0.01)) {
              if (di > RL(0.0)) along = orient + h * RL(90.0);
              else along = orient - h * RL(90.0);
            } else {
              along = orient + h * M<R>::atan(di / (-dj)) * dpr;
              if (dj > 0) along = along + RL(180.);
            }
            R alat1 = C<R>::pi2() - RL(2.0) * M<R>::atan(t), alat = 0;
            for (int iter = 1; iter <= 10; iter++) {
              alat = C<R>::pi2() - RL(2.0) * M<R>::atan(t * M<R>::pow((RL(1.0) - g.e * M<R>::sin(alat1)) / (RL(1.0) + g.e * M<R>::sin(alat1)), g.e_over_2));
              R diff = M<R>::fabs(alat - alat1) * dpr;
              if (diff < RL(0.000001)) break;
              alat1 = alat;
            }
            rlat[n] = h * alat * dpr;
            rlon[n] = along;
            if (rlon[n] < RL(0.0)) rlon[n] = rlon[n] + RL(360.);
            if (rlon[n] > RL(360.0)) rlon[n] = rlon[n] - RL(360.0);
            nret++; rot(n);
          }
        } else {
          if (!(M<R>::fabs(rlon[n]) < (RL(360.) + tiny) && M<R>::fabs(rlat[n]) < (RL(90.) + tiny) &&
                M<R>::fabs(h * rlat[n] + RL(90.)) > tiny)) { xpts[n] = fill; ypts[n] = fill; continue; }
          if (!g.elliptical) {
            R dr = g.de * M<R>::tan((RL(90.) - h * rlat[n]) / RL(2.) / dpr);
            R dr2 = dr * dr;
            xpts[n] = g.xp + h * M<R>::sin((rlon[n] - orient) / dpr) * dr / dxs;
            ypts[n] = g.yp - M<R>::cos((rlon[n] - orient) / dpr) * dr / dys;
            if (inwin(xpts[n], ypts[n])) { nret++; rot(n); jac_area(n, dr2); }
            else { xpts[n] = fill; ypts[n] = fill; }
          } else {
            R alat = h * rlat[n] / dpr;
            R along = (rlon[n] - orient) / dpr;
            R t = M<R>::tan(C<R>::pi4() - alat * RL(0.5)) /
                  M<R>::pow((RL(1.) - g.e * M<R>::sin(alat)) / (RL(1.) + g.e * M<R>::sin(alat)), g.e_over_2);
            R rho = g.rerth * g.mc * t / g.tc;
            xpts[n] = g.xp + rho * M<R>::sin(h * along) / dxs;
            ypts[n] = g.yp - rho * M<R>::cos(h * along) / dys;
            if (inwin(xpts[n], ypts[n])) { nret++; rot(n); }
            else { xpts[n] = fill; ypts[n] = fill; }
          }
        }
      }
      break;
    }
    case G_GAUSS: {  // ip_gaussian_grid_mod.F90:195-363
      const i64 jg = g.jg, j1 = g.j1, jh = g.jh;
      auto AL = [&](i64 idx) -> R { return (idx >= 0 && idx <= jg + 1) ? g.alat[idx] : R(0); };
      auto BL = [&](i64 idx) -> R { return (idx >= 0 && idx <= jg + 1) ? g.blat[idx] : R(0); };
      auto extras = [&](i64 n) {
        if (lrot) { o.crot[n] = RL(1.0); o.srot[n] = RL(0.0); }
        if (lmap) {
          o.xlon[n] = RL(1.) / g.dlon; o.xlat[n] = RL(0.); o.ylon[n] = RL(0.);
          i64 idx = f_nint<R>(ypts[n]);
          o.ylat[n] = (idx >= 0 && idx <= jg + 1) ? g.ylat_row[idx] : R(0);
        }
        if (larea) {
          i64 j = f_int<R>(ypts[n]);
          R wb = ypts[n] - R(j);
          R wlata = BL(j1 + jh * (j - 1)), wlatb = BL(j1 + jh * j);
          R wlat = wlata + wb * (wlatb - wlata);
          o.area[n] = R(g.rerth * g.rerth * wlat * g.dlon / dpr);
        }
      };
      for (i64 n = 0; n < npts; n++) {
        if (fwd) {
          if (inwin(xpts[n], ypts[n])) {
            rlon[n] = M<R>::fmod(g.rlon1 + g.dlon * (xpts[n] - RL(1.)) + RL(3600.), RL(360.));
            i64 j = f_int<R>(ypts[n]);
            R wb = ypts[n] - R(j);
            R rlata = AL(j1 + jh * (j - 1)), rlatb = AL(j1 + jh * j);
            rlat[n] = rlata + wb * (rlatb - rlata);
            nret++; extras(n);
          } else { rlon[n] = fill; rlat[n] = fill; }
        } else {
          xpts[n] = fill; ypts[n] = fill;
          if (M<R>::fabs(rlon[n]) <= RL(360.) && M<R>::fabs(rlat[n]) <= RL(90.)) {
            xpts[n] = RL(1.) + g.hi * M<R>::fmod(g.hi * (rlon[n] - g.rlon1) + RL(3600.), RL(360.)) / g.dlon;
            i64 ja = std::min<i64>(f_int<R>(R(jg + 1) / RL(180.) * (RL(90.) - rlat[n])), jg);
            if (rlat[n] > g.alat[ja]) ja = std::max<i64>(ja - 2, 0);
            if (rlat[n] < g.alat[ja + 1]) ja = std::min<i64>(ja + 2, jg);
            if (rlat[n] > g.alat[ja]) ja = ja - 1;
            if (rlat[n] < g.alat[ja + 1]) ja = ja + 1;
            R yptsa = R(1 + jh * (ja - j1)), yptsb = R(1 + jh * (ja + 1 - j1));
            R wb = (g.alat[ja] - rlat[n]) / (g.alat[ja] - g.alat[ja + 1]);
            ypts[n] = yptsa + wb * (yptsb - yptsa);
            if (inwin(xpts[n], ypts[n])) { nret++; extras(n); }
            else { xpts[n] = fill; ypts[n] = fill; }
          }
        }
      }
      break;
    }
    case G_ROTB:
    case G_ROTE: {
      // ip_rot_equid_cylind_grid_mod.F90:256-443, ip_rot_equid_cylind_egrid_mod.F90:267-457
      const bool eg = (g.type == G_ROTE);
      const double dprd = (double)dpr;
      const double clat0 = g.clat0, slat0 = g.slat0, rlon0 = g.rlon0, dlats = g.dlats, dlons = g.dlons;
      const i64 IM = eg ? g.im * 2 - 1 : g.im, JM = g.jm;
      double wbd = g.wbd, sbd = g.sbd;
      i64 is1 = 0;
      if (eg) {
        sbd = g.rlat1d; wbd = g.rlon1d;
        if (wbd > 180.0) wbd = wbd - 360.0;
        is1 = (g.kscan == 0) ? (JM + 1) / 2 : JM / 2;
      }
      const R kscan = R(g.kscan);
      for (i64 n = 0; n < npts; n++) {
        RotState s;
        if (fwd) {
          R xf = xpts[n], yf = ypts[n];
          if (eg) { xf = ypts[n] + (xpts[n] - R(is1)); yf = ypts[n] - (xpts[n] - R(is1)) + kscan; }
          if (!inwin(xf, yf)) { rlon[n] = fill; rlat[n] = fill; continue; }
          double rlonr, rlatr, hs;
          if (!eg) {
            rlonr = wbd + ((double)xf - 1.) * dlons;
            rlatr = sbd + ((double)yf - 1.) * dlats;
            hs = (rlonr <= 0.) ? -1.0 : 1.0;
          } else {
            hs = g.hid * (double)f_sign<R>(RL(1.), xf - R((IM + 1) / 2));
            if (!g.is_grib2) {
              rlonr = (double)(xf - R((IM + 1) / 2)) * dlons;
              rlatr = (double)(yf - R((JM + 1) / 2)) * dlats;
            } else {
              rlonr = ((double)xf - 1.0) * dlons + wbd;
              rlatr = ((double)yf - 1.0) * dlats + sbd;
            }
          }
          double clonr = lm::cos(rlonr / dprd);
          s.slatr = lm::sin(rlatr / dprd);
          s.clatr = lm::cos(rlatr / dprd);
          s.slat = clat0 * s.slatr + slat0 * s.clatr * clonr;
          if (s.slat <= -1) { s.clat = 0.; s.clon = lm::cos(rlon0 / dprd); rlon[n] = RL(0.); rlat[n] = RL(-90.); }
          else if (s.slat >= 1) { s.clat = 0.; s.clon = lm::cos(rlon0 / dprd); rlon[n] = RL(0.); rlat[n] = RL(90.); }
          else {
            s.clat = ::sqrt(1 - s.slat * s.slat);
            s.clon = (clat0 * s.clatr * clonr - slat0 * s.slatr) / s.clat;
            s.clon = std::min(std::max(s.clon, -1.), 1.);
            rlon[n] = R(::fmod(rlon0 + hs * dprd * lm::acos(s.clon) + 3600, 360.));
            rlat[n] = R(dprd * lm::asin(s.slat));
          }
          nret++;
          if (lrot) rot_vect_rot(g, s, rlon[n], o.crot[n], o.srot[n]);
          if (lmap) rot_map_jacob(g, s, fill, rlon[n], o.xlon[n], o.xlat[n], o.ylon[n], o.ylat[n]);
          if (larea) rot_area(g, s, fill, o.area[n]);
        } else {
          if (!(M<R>::fabs(rlon[n]) <= RL(360.) && M<R>::fabs(rlat[n]) <= RL(90.))) { xpts[n] = fill; ypts[n] = fill; continue; }
          double hs = std::copysign(1., ::fmod((double)rlon[n] - rlon0 + 180 + 3600, 360.) - 180);
          s.clon = lm::cos(((double)rlon[n] - rlon0) / dprd);
          s.slat = lm::sin((double)rlat[n] / dprd);
          s.clat = lm::cos((double)rlat[n] / dprd);
          s.slatr = clat0 * s.slat - slat0 * s.clat * s.clon;
          double rlonr, rlatr;
          if (s.slatr <= -1) { s.clatr = 0.; rlonr = 0.; rlatr = -90.; }
          else if (s.slatr >= 1) { s.clatr = 0.; rlonr = 0.; rlatr = 90.; }
          else {
            s.clatr = ::sqrt(1 - s.slatr * s.slatr);
            double clonr = (clat0 * s.clat * s.clon + slat0 * s.slat) / s.clatr;
            clonr = std::min(std::max(clonr, -1.), 1.);
            rlonr = hs * dprd * lm::acos(clonr);
            rlatr = dprd * lm::asin(s.slatr);
          }
          R xf, yf;
          if (!eg || !g.is_grib2) {
            xf = R(((rlonr - wbd) / dlons) + 1.0);
            yf = R(((rlatr - sbd) / dlats) + 1.0);
          } else {
            xf = R((double)((IM + 1) / 2) + rlonr / dlons);
            yf = R((double)((JM + 1) / 2) + rlatr / dlats);
          }
          if (inwin(xf, yf)) {
            if (eg) {
              xpts[n] = R(is1) + (xf - (yf - kscan)) / RL(2.);
              ypts[n] = (xf + (yf - kscan)) / RL(2.);
            } else { xpts[n] = xf; ypts[n] = yf; }
            nret++;
            if (lrot) rot_vect_rot(g, s, rlon[n], o.crot[n], o.srot[n]);
            if (lmap) rot_map_jacob(g, s, fill, rlon[n], o.xlon[n], o.xlat[n], o.ylon[n], o.ylat[n]);
            if (larea) rot_area(g, s, fill, o.area[n]);
          } else { xpts[n] = fill; ypts[n] = fill; }
        }
      }
      break;
    }
    default: break;
  }
}

// ---- src/gdswzd_mod.F90:105-199 (gdswzd_grid; iopt=0 enumerates the grid's own points) -----
// Returns false when iopt==0 and nm>npts (the reference returns without touching nret).
template <class R>
bool gdswzd_grid(const Grid<R>& g, int iopt, i64 npts, R fill, R* xpts, R* ypts, R* rlon, R* rlat, i64& nret,
                 const GdsOpt<R>& o) {
  int iopf = iopt;
  if (iopt == 0) {
    iopf = 1;
    i64 im = g.im, jm = g.jm, nm = (g.grid_num == -1) ? npts : im * jm;
    // NB: the reference tests grid_num == -1 exactly (gdswzd_mod.F90:127); other negative
    // station grid numbers have im=jm=0 => nm=0.
    int nscan = g.nscan, kscan = g.kscan;
    if (nm > npts) {
      for (i64 n = 0; n < npts; n++) { rlat[n] = fill; rlon[n] = fill; xpts[n] = fill; ypts[n] = fill; }
      return false;
    }
    if (g.type == G_ROTE) {
      i64 is1 = (kscan == 0) ? (jm + 1) / 2 : jm / 2;
      for (i64 n = 1; n <= nm; n++) {
        i64 i, j;
        if (nscan == 0) { j = (n - 1) / im + 1; i = (n - im * (j - 1)) * 2 - imod(j + kscan, 2); }
        else {
          i64 nn = (n * 2) - 1 + kscan;
          i = (nn - 1) / jm + 1; j = imod(nn - 1, jm) + 1;
          if (imod(jm, 2) == 0 && imod(i, 2) == 0 && kscan == 0) j = j + 1;
          if (imod(jm, 2) == 0 && imod(i, 2) == 0 && kscan == 1) j = j - 1;
        }
        xpts[n - 1] = R(is1 + (i - (j - kscan)) / 2);
        ypts[n - 1] = R((i + (j - kscan)) / 2);
      }
    } else if (g.type == G_STATION) {
      for (i64 n = 0; n < nm; n++) { xpts[n] = fill; ypts[n] = fill; }
    } else {
      for (i64 n = 1; n <= nm; n++) {
        i64 i, j;
        if (nscan == 0) { j = (n - 1) / im + 1; i = n - im * (j - 1); }
        else { i = (n - 1) / jm + 1; j = n - jm * (i - 1); }
        xpts[n - 1] = R(i); ypts[n - 1] = R(j);
      }
    }
    for (i64 n = nm; n < npts; n++) { xpts[n] = fill; ypts[n] = fill; }
  }
  gdswzd_proj<R>(g, iopf, npts, fill, xpts, ypts, rlon, rlat, nret, o);
  return true;
}

}  // namespace ipo
