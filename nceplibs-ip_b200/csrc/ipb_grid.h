// ipb200 — grid description shared by host and device code.
//
// GridP<R> is the POD a kernel receives by value: every scalar the eight map projections of
// NCEPLIBS-ip need, already rounded in the build kind's arithmetic (R = float for the _4 build,
// double for _d/_8).  It replaces the reference's class hierarchy (src/ip_grid_mod.F90:54-87 and
// the eight src/ip_*_grid_mod.F90 types) and the per-call set-up block at the top of each
// gdswzd_* routine, which is hoisted into the host-side decode (ipb_decode.h).
#pragma once
#include <cstdint>

namespace ipb {

typedef long long i64;
typedef unsigned char u8;

enum ProjType : int {
  P_EQUID = 0,    // equidistant cylindrical (lat-lon)      GRIB2 0      / GRIB1 0
  P_MERC = 1,     // Mercator                                GRIB2 10     / GRIB1 1
  P_LAMBERT = 2,  // Lambert conformal                       GRIB2 30     / GRIB1 3
  P_GAUSS = 3,    // Gaussian                                GRIB2 40     / GRIB1 4
  P_POLAR = 4,    // polar stereographic                     GRIB2 20     / GRIB1 5
  P_ROTB = 5,     // rotated lat-lon, B / non-E stagger      GRIB2 1,32769/ GRIB1 205
  P_ROTE = 6,     // rotated lat-lon, E stagger              GRIB2 1,32768/ GRIB1 203
  P_STATION = 7,  // station points (grid number < 0)
  P_BAD = 8
};

template <class R> struct GridP {
  int type, is_grib2, irot, elliptical;
  int nscan, kscan, nscan_fp;
  int im, jm, iwrap, jwrap1, jwrap2;
  int jg, jh, j1;
  i64 grid_num;
  R rerth, e2;
  R hi, rlat1, rlon1, dlat, dlon, dphi, ye;
  R orient, dxs, dys, h, an, de, de2, xp, yp, antr, ee, e_over_2, tc, mc;
  R xmin, xmax, ymin, ymax;
  double clat0, slat0, rlon0, dlats, dlons, wbd, sbd, hid;
  const R* alat;      // Gaussian latitudes, index 0..jg+1 (device pointer once uploaded)
  const R* blat;      // Gaussian weights
  const R* ylat_row;  // Gaussian dy/dlat per row
};

}  // namespace ipb
