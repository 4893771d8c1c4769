// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
//
// CPU restatement of the grid expanders that usually sit right before ipolates:
//   ipxwafs / ipxwafs2 / ipxwafs3   /root/reference/src/ipxwafs.F90:81-225, ipxwafs2.F90:92-253, ipxwafs3.F90:95-274
//   ipxetas                         /root/reference/src/ipxetas.F90:90-204
// Pinned by the known answers of /root/reference/tests/test_ipxwafs.F90:59-68 and its round-trip check (:94)
// (tests/test_oracle_expand.py).
#pragma once
#include <cmath>
#include <vector>

#include "ip_oracle_interp.hpp"

namespace ipo {

// points per row of the thinned WAFS grid, equator to pole (ipxwafs.F90:120-126)
static const int NPWAFS[73] = {73, 73, 73, 73, 73, 73, 73, 73, 72, 72, 72, 71, 71, 71, 70, 70, 69, 69, 68, 67, 67, 66, 65, 65, 64,
                               63, 62, 61, 60, 60, 59, 58, 57, 56, 55, 54, 52, 51, 50, 49, 48, 47, 45, 44, 43, 42, 40, 39, 38, 36,
                               35, 33, 32, 30, 29, 28, 26, 25, 23, 22, 20, 19, 17, 16, 14, 12, 11, 9,  8,  6,  5,  3,  2};

// variant 1: ipxwafs (no bitmaps), 2: ipxwafs2 (bilinear along the row, both neighbours must be valid),
// 3: ipxwafs3 (nearest neighbour along the row).  Arrays are Fortran (npts, km); templates and opt_pts are in/out.
template <class R>
void ipxwafs_any(int variant, i64 idir, i64 nthin, i64 nfull, i64 km, i64 num_opt, i64* opt, i64 len, i64* tthin, R* dthin, i64* ibthin,
                 u8* bthin, i64* tfull, R* dfull, i64* ibfull, u8* bfull, i64& iret) {
  iret = 0;
  if (len != 19 || nthin != 3447 || nfull != 5329) { iret = 1; return; }   // ipxwafs.F90:130-133
  if (idir > 0) {                                                          // :135-156
    const i64 scan = tthin[18];
    i64 iscale = tthin[9] * tthin[10];
    if (iscale == 0) iscale = 1000000;
    const i64 idlat = f_nint<R>(RL(1.25) * (R)iscale);
    bool t1 = num_opt == 73, t2 = num_opt == 73;
    for (int j = 0; j < 73 && (t1 || t2); j++) { t1 = t1 && opt[j] == NPWAFS[j]; t2 = t2 && opt[j] == NPWAFS[72 - j]; }
    if (scan == 64 && tthin[8] == 73 && idlat == tthin[17] && (t1 || t2)) {
      for (int q = 0; q < 19; q++) tfull[q] = tthin[q];
      const i64 im = 73;
      tfull[7] = im;
      const R rlon1 = (R)tfull[12] / (R)iscale, rlon2 = (R)tfull[15] / (R)iscale;
      const i64 iscan = (tfull[18] / 128) % 2;
      const R hi = iscan ? RL(-1.) : RL(1.);
      const R dlon = hi * (std::fmod(hi * (rlon2 - rlon1) - R(1) + R(3600), RL(360.)) + R(1)) / (R)(im - 1);
      tfull[16] = f_nint<R>(dlon * (R)iscale);
    } else iret = 1;
  } else if (idir < 0) {                                                   // :157-177
    const i64 scan = tfull[18];
    i64 iscale = tfull[9] * tfull[10];
    if (iscale == 0) iscale = 1000000;
    const i64 idlat = f_nint<R>(RL(1.25) * (R)iscale), idlon = idlat;
    if (scan == 64 && tfull[7] == 73 && tfull[8] == 73 && num_opt == 73 && idlat == tfull[17] && idlon == tfull[16]) {
      for (int q = 0; q < 19; q++) tthin[q] = tfull[q];
      tthin[7] = -1; tthin[16] = -1;
      for (int j = 0; j < 73; j++) opt[j] = tthin[11] == 0 ? NPWAFS[j] : NPWAFS[72 - j];
    } else iret = 1;
  }
  if (iret != 0 || (idir != 1 && idir != -1)) return;
  const bool expand = idir == 1;
  const i64 jm = tfull[8], imf = tfull[7];
  for (i64 k = 0; k < km; k++) {
    i64 is1 = 0, is2 = 0;
    if (variant != 1) (expand ? ibfull : ibthin)[k] = 0;
    const i64 ibin = variant != 1 ? (expand ? ibthin : ibfull)[k] : 0;
    const R* din = (expand ? dthin + k * nthin : dfull + k * nfull);
    R* dout = (expand ? dfull + k * nfull : dthin + k * nthin);
    const u8* bin = variant != 1 ? (expand ? bthin + k * nthin : bfull + k * nfull) : nullptr;
    u8* bout = variant != 1 ? (expand ? bfull + k * nfull : bthin + k * nthin) : nullptr;
    for (i64 j = 0; j < jm; j++) {
      const i64 im1 = opt[j], im2 = imf;
      const i64 nin = expand ? im1 : im2, nout = expand ? im2 : im1, sin_ = expand ? is1 : is2, sout = expand ? is2 : is1;
      const R rat = (R)(nin - 1) / (R)(nout - 1);                          // RAT1 / RAT2
      for (i64 i = 1; i <= nout; i++) {
        const R x = (R)(i - 1) * rat + R(1);
        i64 ia = (i64)x;
        ia = std::min(std::max(ia, (i64)1), nin - 1);
        const i64 ib = ia + 1;
        const R wa = (R)ib - x, wb = x - (R)ia;
        const i64 pa = sin_ + ia - 1, pb = sin_ + ib - 1, po = sout + i - 1;
        if (variant == 1) dout[po] = wa * din[pa] + wb * din[pb];
        else if (variant == 2) {
          if (ibin == 0 || (bin[pa] && bin[pb])) { dout[po] = wa * din[pa] + wb * din[pb]; bout[po] = 1; }
          else { dout[po] = 0; bout[po] = 0; (expand ? ibfull : ibthin)[k] = 1; }
        } else {
          const i64 pn = wa >= wb ? pa : pb;
          if (ibin == 0 || bin[pn]) { dout[po] = din[pn]; bout[po] = 1; }
          else { dout[po] = 0; bout[po] = 0; (expand ? ibfull : ibthin)[k] = 1; }
        }
      }
      is1 += im1; is2 += im2;
    }
  }
}

// ipxetas.F90:90-204: staggered E grid (H or V points) <-> full E grid, through ipolates bilinear (ip=0, ibi=1, km=1)
template <class R>
void ipxetas(i64 idir, i64 numi, i64 len, const i64* tin, i64 npi, const u8* bi, const R* di, i64& numo, i64* tout, i64 npo, u8* bo, R* dout,
             i64& iret) {
  iret = 0;
  if (numi != 1) { iret = 1; return; }
  const i64 scan = tin[18];
  auto half_dlons = [&](double f) {
    i64 iscale = tout[9] * tout[10];
    if (iscale == 0) iscale = 1000000;
    R dlons = (R)tout[16] / (R)iscale;
    dlons = dlons * (R)f;
    tout[16] = f_nint<R>(dlons * (R)iscale);
  };
  numo = numi;
  for (i64 q = 0; q < len; q++) tout[q] = tin[q];
  if ((scan == 68 || scan == 72) && (idir < -2 || idir > -1)) {            // staggered -> full
    tout[18] = 64; tout[7] = tout[7] * 2 - 1;
    if (tout[7] * tout[8] != npo) { iret = 3; return; }
    half_dlons(0.5);
  } else if (scan == 64 && idir == -1) {                                    // full -> H
    tout[18] = 68; tout[7] = (tout[7] + 1) / 2;
    if (tout[7] * tout[8] != npo) { iret = 3; return; }
    half_dlons(2.0);
  } else if (scan == 64 && idir == -2) {                                    // full -> V
    tout[18] = 72; tout[7] = (tout[7] + 1) / 2;
    if (tout[7] * tout[8] != npo) { iret = 3; return; }
    half_dlons(2.0);
  } else { iret = 2; return; }
  Grid<R> gin, gout;
  if (!init_grid_grib2<R>(gin, numi, tin, len) || !init_grid_grib2<R>(gout, numo, tout, len)) { iret = 1; return; }
  i64 ipopt[20] = {0}, ibi = 1, ibo = 0, no = 0;
  std::vector<R> rlat((size_t)npo), rlon((size_t)npo);
  ipolates_grid<R>(0, ipopt, gin, gout, npi, npo, 1, &ibi, bi, di, no, rlat.data(), rlon.data(), &ibo, bo, dout, iret);
  if (iret != 0) return;
  const i64 im = tout[7], jm = tout[8];
  for (i64 j = 1; j <= jm; j++) {                                           // :195-200: the row ends copy their neighbours
    bo[j * im - 1] = bo[j * im - 2]; dout[j * im - 1] = dout[j * im - 2];
    bo[(j - 1) * im] = bo[(j - 1) * im + 1]; dout[(j - 1) * im] = dout[(j - 1) * im + 1];
  }
}

}  // namespace ipo
