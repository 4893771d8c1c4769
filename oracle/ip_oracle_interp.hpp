// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
//
// CPU restatement of the NCEPLIBS-ip 5.0.0 interpolators (bilinear, bicubic, neighbor, budget,
// neighbor-budget; scalar and vector), MOVECT and POLFIXS/POLFIXV.  Loop structure, option
// decoding, accumulation order and error behaviour follow the reference files cited on each
// function (paths relative to /root/reference).  OpenMP pragmas sit on the same loops the
// reference parallelises, so this file doubles as the "port" CPU baseline.
//
// Deliberate differences from the reference (documented in DESIGN.md):
//  * the single-entry SAVE caches are keyed on (descriptor_in, descriptor_out) like the
//    reference's, but a failed or skipped build never leaves a stale plan behind;
//  * x/y, input lat/lon and input rotations are kept with the cached plan (the reference
//    reads unsaved locals on a cache hit: bilinear_interp_mod.F90:98,223-227);
//  * "mo smaller than the output grid" gives iret=3, no=0 (the reference leaves no untouched).
#pragma once
#include "ip_oracle_grid.hpp"

namespace ipo {

typedef unsigned char u8;

// ---- src/movect.F90:25-68 (always real64 inside; constants are default-real literals) ------
template <class R> inline void movect(R flat, R flon, R tlat, R tlon, R& crot, R& srot) {
  const double crdlim = (double)RL(0.9999999);
  const double pi = (double)RL(3.14159265358979);
  const double dpr = 180. / pi;
  double ctlat = lm::cos((double)tlat / dpr), stlat = lm::sin((double)tlat / dpr);
  double cflat = lm::cos((double)flat / dpr), sflat = lm::sin((double)flat / dpr);
  // FLON-TLON is formed in the default kind before the promotion
  double cdlon = lm::cos((double)(flon - tlon) / dpr), sdlon = lm::sin((double)(flon - tlon) / dpr);
  double crd = stlat * sflat + ctlat * cflat * cdlon;
  if (std::fabs(crd) <= crdlim) {
    double srd2rn = -1 / (1 - crd * crd);
    double str = cflat * sdlon;
    double ctr = cflat * stlat * cdlon - sflat * ctlat;
    double sfr = ctlat * sdlon;
    double cfr = ctlat * sflat * cdlon - stlat * cflat;
    crot = R(srd2rn * (ctr * cfr - str * sfr));
    srot = R(srd2rn * (ctr * sfr + str * cfr));
  } else {
    crot = R(cdlon);
    srot = R(sdlon * stlat);
  }
}

// ---- src/polfix_mod.F90:29-105 ---------------------------------------------------------------
template <class R> void polfixs(i64 nm, i64 nx, i64 km, const R* rlat, const i64* ib, u8* lo, R* go) {
  const R rlatnp = RL(89.9995), rlatsp = -rlatnp;
  for (i64 k = 0; k < km; k++) {
    R wnp = 0, gnp = 0, tnp = 0, wsp = 0, gsp = 0, tsp = 0;
    for (i64 n = 0; n < nm; n++) {
      if (rlat[n] >= rlatnp) {
        wnp = wnp + 1;
        if (ib[k] == 0 || lo[n + k * nx]) { gnp = gnp + go[n + k * nx]; tnp = tnp + 1; }
      } else if (rlat[n] <= rlatsp) {
        wsp = wsp + 1;
        if (ib[k] == 0 || lo[n + k * nx]) { gsp = gsp + go[n + k * nx]; tsp = tsp + 1; }
      }
    }
    if (wnp > 1) {
      if (tnp >= wnp / 2) gnp = gnp / tnp; else gnp = 0;
      for (i64 n = 0; n < nm; n++)
        if (rlat[n] >= rlatnp) { if (ib[k] != 0) lo[n + k * nx] = tnp >= wnp / 2; go[n + k * nx] = gnp; }
    }
    if (wsp > 1) {
      if (tsp >= wsp / 2) gsp = gsp / tsp; else gsp = 0;
      for (i64 n = 0; n < nm; n++)
        if (rlat[n] <= rlatsp) { if (ib[k] != 0) lo[n + k * nx] = tsp >= wsp / 2; go[n + k * nx] = gsp; }
    }
  }
}

// ---- src/polfix_mod.F90:124-220 (south pole divides by WSP, sic) ------------------------------
template <class R> void polfixv(i64 nm, i64 nx, i64 km, const R* rlat, const R* rlon, const i64* ib, u8* lo, R* uo, R* vo) {
  const R rlatnp = RL(89.9995), rlatsp = -rlatnp;
  const R dpr = RL(180.) / RL(3.14159265358979);
  std::vector<R> clon(nm), slon(nm);
  for (i64 n = 0; n < nm; n++) { clon[n] = M<R>::cos(rlon[n] / dpr); slon[n] = M<R>::sin(rlon[n] / dpr); }
  for (i64 k = 0; k < km; k++) {
    R wnp = 0, unp = 0, vnp = 0, tnp = 0, wsp = 0, usp = 0, vsp = 0, tsp = 0;
    for (i64 n = 0; n < nm; n++) {
      const i64 a = n + k * nx;
      if (rlat[n] >= rlatnp) {
        wnp = wnp + 1;
        if (ib[k] == 0 || lo[a]) {
          unp = unp + (clon[n] * uo[a] - slon[n] * vo[a]);
          vnp = vnp + (slon[n] * uo[a] + clon[n] * vo[a]);
          tnp = tnp + 1;
        }
      } else if (rlat[n] <= rlatsp) {
        wsp = wsp + 1;
        if (ib[k] == 0 || lo[a]) {
          usp = usp + (clon[n] * uo[a] + slon[n] * vo[a]);
          vsp = vsp + (-slon[n] * uo[a] + clon[n] * vo[a]);
          tsp = tsp + 1;
        }
      }
    }
    if (wnp > 1) {
      if (tnp >= wnp / 2) { unp = unp / tnp; vnp = vnp / tnp; } else { unp = 0; vnp = 0; }
      for (i64 n = 0; n < nm; n++)
        if (rlat[n] >= rlatnp) {
          const i64 a = n + k * nx;
          if (ib[k] != 0) lo[a] = tnp >= wnp / 2;
          uo[a] = clon[n] * unp + slon[n] * vnp;
          vo[a] = -slon[n] * unp + clon[n] * vnp;
        }
    }
    if (wsp > 1) {
      if (tsp >= wsp / 2) { usp = usp / wsp; vsp = vsp / wsp; } else { usp = 0; vsp = 0; }
      for (i64 n = 0; n < nm; n++)
        if (rlat[n] <= rlatsp) {
          const i64 a = n + k * nx;
          if (ib[k] != 0) lo[a] = tsp >= wsp / 2;
          uo[a] = clon[n] * usp - slon[n] * vsp;
          vo[a] = slon[n] * usp + clon[n] * vsp;
        }
    }
  }
}

// ---- spiral search step (e.g. bilinear_interp_mod.F90:228-246) -------------------------------
template <class R> inline void spiral_step(i64 mx, i64 i1, i64 j1, i64 ixs, i64 jxs, i64& ix, i64& jx) {
  i64 kxs = f_int<R>(M<R>::sqrt(R(4 * mx) - RL(2.5)));
  i64 kxt = mx - (kxs * kxs / 4 + 1);
  switch (imod(kxs, 4)) {
    case 1: ix = i1 - ixs * (kxs / 4 - kxt); jx = j1 - jxs * kxs / 4; break;
    case 2: ix = i1 + ixs * (1 + kxs / 4); jx = j1 - jxs * (kxs / 4 - kxt); break;
    case 3: ix = i1 + ixs * (1 + kxs / 4 - kxt); jx = j1 + jxs * (1 + kxs / 4); break;
    default: ix = i1 - ixs * kxs / 4; jx = j1 + jxs * (kxs / 4 - kxt); break;
  }
}
template <class R> inline bool valid_xy(R x, R y) {
  const R fill = RL(-9999.), tiny = M<R>::tiny();
  return M<R>::fabs(x - fill) > tiny && M<R>::fabs(y - fill) > tiny;
}

// ---- plan shared by the cached interpolators --------------------------------------------------
template <class R> struct Plan {
  bool built = false;
  std::vector<i64> din, dout;
  i64 nox = -1, iretx = -1;
  int nt = 0;  // taps per point
  std::vector<i64> nxy;   // [nt*no]
  std::vector<R> wxy, cxy, sxy;
  std::vector<int> nc;
  std::vector<R> rlatx, rlonx, crotx, srotx, xptsx, yptsx;
  std::vector<R> rlai, rloi, croi, sroi;  // input-grid lat/lon/rotation (vector methods)
};

template <class R> inline bool same_pair(const Plan<R>& p, const Grid<R>& gin, const Grid<R>& gout) {
  return p.built && p.din == gin.desc && p.dout == gout.desc;
}

// Locate step common to the cached interpolators (e.g. bilinear_interp_mod.F90:146-150).
template <class R>
void locate(const Grid<R>& gin, const Grid<R>& gout, i64 mo, i64& no, R* rlat, R* rlon, R* crot, R* srot,
            std::vector<R>& xpts, std::vector<R>& ypts, i64& iret) {
  const R fill = RL(-9999.);
  xpts.assign(mo, fill); ypts.assign(mo, fill);
  GdsOpt<R> o; o.crot = crot; o.srot = srot;
  if (!gdswzd_grid<R>(gout, 0, mo, fill, xpts.data(), ypts.data(), rlon, rlat, no, o)) no = 0;
  if (no == 0) iret = 3;
  i64 nv = 0;
  GdsOpt<R> none;
  gdswzd_grid<R>(gin, -1, no, fill, xpts.data(), ypts.data(), rlon, rlat, nv, none);
  if (iret == 0 && nv == 0) iret = 2;
}
template <class R> void input_geo(const Grid<R>& gin, i64 mi, Plan<R>& p) {
  const R fill = RL(-9999.);
  std::vector<R> xi(mi, fill), yi(mi, fill);
  p.rloi.assign(mi, fill); p.rlai.assign(mi, fill); p.croi.assign(mi, fill); p.sroi.assign(mi, fill);
  GdsOpt<R> o; o.crot = p.croi.data(); o.srot = p.sroi.data();
  i64 nv = 0;
  gdswzd_grid<R>(gin, 0, mi, fill, xi.data(), yi.data(), p.rloi.data(), p.rlai.data(), nv, o);
}

template <class R> inline void set_ibo(i64 no, i64 mo, i64 km, const i64* ibi, const u8* lo, i64* ibo) {
  for (i64 k = 0; k < km; k++) {
    ibo[k] = ibi[k];
    bool all = true;
    for (i64 n = 0; n < no; n++) if (!lo[n + k * mo]) { all = false; break; }
    if (!all) ibo[k] = 1;
  }
}

// =================================================================================================
// bilinear: src/bilinear_interp_mod.F90:72-275 (scalar), :329-520 (vector)
// bicubic : src/bicubic_interp_mod.F90:80-279 (scalar), :338-568 (vector)
// One body for both: NT=2 is bilinear, NT=4 bicubic.
// =================================================================================================
template <class R, int NT, bool VEC>
void interp_stencil(const i64* ipopt, const Grid<R>& gin, const Grid<R>& gout, i64 mi, i64 mo, i64 km, const i64* ibi,
                    const u8* li, const R* gi, const R* vi, i64& no, R* rlat, R* rlon, R* crot, R* srot, i64* ibo,
                    u8* lo, R* go, R* vo, i64& iret) {
  static Plan<R> plan;
  Plan<R>& p = plan;
  const int T2 = NT * NT;
  iret = 0;
  i64 mcon = 0, mp, mspiral = 0;
  if (NT == 2) { mp = ipopt[0]; mspiral = std::max<i64>(ipopt[1], 0); }
  else { mcon = ipopt[0]; mp = ipopt[1]; }
  if (mp == -1 || mp == 0) mp = 50;
  if (mp < 0 || mp > 100) iret = 32;
  const R pmp = R(mp) * RL(0.01);
  const bool to_station = gout.type == G_STATION;
  const bool same = same_pair(p, gin, gout);
  if (iret == 0 && (to_station || !same)) {
    p.built = false;
    std::vector<R> xpts, ypts;
    locate<R>(gin, gout, mo, no, rlat, rlon, VEC ? crot : nullptr, VEC ? srot : nullptr, xpts, ypts, iret);
    if (VEC) input_geo<R>(gin, mi, p);
    p.nox = no; p.iretx = iret; p.nt = T2;
    if (iret == 0) {
      p.nxy.assign((size_t)T2 * no, 0); p.wxy.assign((size_t)T2 * no, 0); p.nc.assign(no, 0);
      if (VEC) { p.cxy.assign((size_t)T2 * no, 0); p.sxy.assign((size_t)T2 * no, 0); p.crotx.assign(crot, crot + no); p.srotx.assign(srot, srot + no); }
      p.rlatx.assign(rlat, rlat + no); p.rlonx.assign(rlon, rlon + no);
      p.xptsx = xpts; p.yptsx = ypts;
#pragma omp parallel for schedule(static)
      for (i64 n = 0; n < no; n++) {
        const R xij = xpts[n], yij = ypts[n];
        i64* nx = &p.nxy[(size_t)T2 * n];
        R* wx_ = &p.wxy[(size_t)T2 * n];
        if (valid_xy(xij, yij)) {
          i64 ijx[NT], ijy[NT];
          R wx[NT], wy[NT];
          if (NT == 2) {
            for (int a = 0; a < 2; a++) { ijx[a] = f_floor<R>(xij) + a; ijy[a] = f_floor<R>(yij) + a; }
            R xf = xij - R(ijx[0]), yf = yij - R(ijy[0]);
            wx[0] = (RL(1.) - xf); wx[1] = xf; wy[0] = (RL(1.) - yf); wy[1] = yf;
            for (int j = 0; j < 2; j++) for (int i = 0; i < 2; i++) nx[i + 2 * j] = field_pos(gin, ijx[i], ijy[j]);
            p.nc[n] = 1;
          } else {
            for (int a = 0; a < 4; a++) { ijx[a] = f_floor<R>(xij - RL(1.)) + a; ijy[a] = f_floor<R>(yij - RL(1.)) + a; }
            R xf = xij - R(ijx[1]), yf = yij - R(ijy[1]);
            i64 mn = 1;
            for (int j = 0; j < 4; j++) for (int i = 0; i < 4; i++) { nx[i + 4 * j] = field_pos(gin, ijx[i], ijy[j]); if (nx[i + 4 * j] <= 0) mn = 0; }
            if (mn > 0) {
              p.nc[n] = 1;
              wx[0] = xf * (RL(1.) - xf) * (RL(2.) - xf) / RL(-6.);
              wx[1] = (xf + RL(1.)) * (RL(1.) - xf) * (RL(2.) - xf) / RL(2.);
              wx[2] = (xf + RL(1.)) * xf * (RL(2.) - xf) / RL(2.);
              wx[3] = (xf + RL(1.)) * xf * (RL(1.) - xf) / RL(-6.);
              wy[0] = yf * (RL(1.) - yf) * (RL(2.) - yf) / RL(-6.);
              wy[1] = (yf + RL(1.)) * (RL(1.) - yf) * (RL(2.) - yf) / RL(2.);
              wy[2] = (yf + RL(1.)) * yf * (RL(2.) - yf) / RL(2.);
              wy[3] = (yf + RL(1.)) * yf * (RL(1.) - yf) / RL(-6.);
            } else {
              p.nc[n] = 2;
              wx[0] = 0; wx[1] = (RL(1.) - xf); wx[2] = xf; wx[3] = 0;
              wy[0] = 0; wy[1] = (RL(1.) - yf); wy[2] = yf; wy[3] = 0;
            }
          }
          for (int j = 0; j < NT; j++)
            for (int i = 0; i < NT; i++) {
              const int t = i + NT * j;
              wx_[t] = wx[i] * wy[j];
              if (VEC && nx[t] > 0) {
                R cm, sm;
                movect<R>(p.rlai[nx[t] - 1], p.rloi[nx[t] - 1], rlat[n], rlon[n], cm, sm);
                p.cxy[(size_t)T2 * n + t] = cm * p.croi[nx[t] - 1] + sm * p.sroi[nx[t] - 1];
                p.sxy[(size_t)T2 * n + t] = sm * p.croi[nx[t] - 1] - cm * p.sroi[nx[t] - 1];
              }
            }
        } else {
          for (int t = 0; t < T2; t++) nx[t] = 0;
          p.nc[n] = (NT == 2) ? 1 : 0;
        }
      }
      p.din = gin.desc; p.dout = gout.desc; p.built = true;
    }
  }
  if (iret == 0 && p.built && p.iretx == 0) {
    if (!to_station) {
      no = p.nox;
      for (i64 n = 0; n < no; n++) {
        rlon[n] = p.rlonx[n]; rlat[n] = p.rlatx[n];
        if (VEC) { crot[n] = p.crotx[n]; srot[n] = p.srotx[n]; }
      }
    }
    const i64 nktot = no * km;
#pragma omp parallel for schedule(static)
    for (i64 nk = 0; nk < nktot; nk++) {
      const i64 k = nk / no, n = nk - no * k;
      const i64* nx = &p.nxy[(size_t)T2 * n];
      const R* w_ = &p.wxy[(size_t)T2 * n];
      const size_t oa = n + k * mo;
      const int nc = p.nc[n];
      if (nc > 0) {
        R g = 0, v = 0, w = 0;
        R gmin = 0, gmax = 0, vmin = 0, vmax = 0;
        if (mcon > 0) { gmin = M<R>::huge(); gmax = -M<R>::huge(); vmin = M<R>::huge(); vmax = -M<R>::huge(); }
        const int lo_t = (NT == 2) ? 0 : nc - 1, hi_t = (NT == 2) ? 1 : 4 - nc;
        for (int j = lo_t; j <= hi_t; j++)
          for (int i = lo_t; i <= hi_t; i++) {
            const int t = i + NT * j;
            if (nx[t] > 0) {
              const size_t ia = (size_t)(nx[t] - 1) + (size_t)k * mi;
              if (ibi[k] == 0 || li[ia]) {
                if (!VEC) {
                  g = g + w_[t] * gi[ia];
                  w = w + w_[t];
                  if (mcon > 0) { gmin = f_min(gmin, gi[ia]); gmax = f_max(gmax, gi[ia]); }
                } else {
                  const R c = p.cxy[(size_t)T2 * n + t], s = p.sxy[(size_t)T2 * n + t];
                  R urot = c * gi[ia] - s * vi[ia];
                  R vrot = s * gi[ia] + c * vi[ia];
                  g = g + w_[t] * urot;
                  v = v + w_[t] * vrot;
                  w = w + w_[t];
                  if (mcon > 0) { gmin = f_min(gmin, urot); gmax = f_max(gmax, urot); vmin = f_min(vmin, vrot); vmax = f_max(vmax, vrot); }
                }
              }
            }
          }
        lo[oa] = w >= pmp;
        if (lo[oa]) {
          if (!VEC) {
            go[oa] = g / w;
            if (mcon > 0) go[oa] = f_min(f_max(go[oa], gmin), gmax);
          } else {
            R urot = crot[n] * g - srot[n] * v;
            R vrot = srot[n] * g + crot[n] * v;
            go[oa] = urot / w; vo[oa] = vrot / w;
            if (mcon > 0) { go[oa] = f_min(f_max(go[oa], gmin), gmax); vo[oa] = f_min(f_max(vo[oa], vmin), vmax); }
          }
        } else if (NT == 2 && !VEC && mspiral > 0 && valid_xy(p.xptsx[n], p.yptsx[n])) {
          const R x = p.xptsx[n], y = p.yptsx[n];
          i64 i1 = f_nint<R>(x), j1 = f_nint<R>(y);
          i64 ixs = f_int<R>(f_sign<R>(RL(1.), x - R(i1))), jxs = f_int<R>(f_sign<R>(RL(1.), y - R(j1)));
          for (i64 mx = 1; mx <= mspiral * mspiral; mx++) {
            i64 ix, jx;
            spiral_step<R>(mx, i1, j1, ixs, jxs, ix, jx);
            i64 nxx = field_pos(gin, ix, jx);
            if (nxx > 0) {
              const size_t ia = (size_t)(nxx - 1) + (size_t)k * mi;
              if (li[ia] || ibi[k] == 0) { go[oa] = gi[ia]; lo[oa] = 1; break; }
            }
          }
          if (!lo[oa]) go[oa] = 0;
        } else {
          go[oa] = 0;
          if (VEC) vo[oa] = 0;
        }
      } else {
        lo[oa] = 0; go[oa] = 0;
        if (VEC) vo[oa] = 0;
      }
    }
    set_ibo<R>(no, mo, km, ibi, lo, ibo);
    if (gout.type == G_EQUID) {
      if (!VEC) polfixs<R>(no, mo, km, rlat, ibo, lo, go);
      else polfixv<R>(no, mo, km, rlat, rlon, ibo, lo, go, vo);
    }
  } else {
    if (iret == 0) iret = (p.built ? p.iretx : (p.iretx > 0 ? p.iretx : 1));
    if (!to_station) no = 0;
  }
}

// =================================================================================================
// neighbor: src/neighbor_interp_mod.F90:100-274 (scalar), :352-557 (vector)
// =================================================================================================
template <class R, bool VEC>
void interp_neighbor(const i64* ipopt, const Grid<R>& gin, const Grid<R>& gout, i64 mi, i64 mo, i64 km, const i64* ibi,
                     const u8* li, const R* gi, const R* vi, i64& no, R* rlat, R* rlon, R* crot, R* srot, i64* ibo,
                     u8* lo, R* go, R* vo, i64& iret) {
  static Plan<R> plan;
  Plan<R>& p = plan;
  iret = 0;
  const i64 mspiral = std::max<i64>(ipopt[0], 1);
  const bool to_station = gout.type == G_STATION;
  const bool same = same_pair(p, gin, gout);
  if (to_station || !same) {
    p.built = false;
    std::vector<R> xpts, ypts;
    locate<R>(gin, gout, mo, no, rlat, rlon, VEC ? crot : nullptr, VEC ? srot : nullptr, xpts, ypts, iret);
    if (VEC) input_geo<R>(gin, mi, p);
    p.nox = no; p.iretx = iret;
    if (iret == 0) {
      p.nxy.assign(no, 0);
      if (VEC) { p.cxy.assign(no, 0); p.sxy.assign(no, 0); p.crotx.assign(crot, crot + no); p.srotx.assign(srot, srot + no); }
      p.rlatx.assign(rlat, rlat + no); p.rlonx.assign(rlon, rlon + no);
      p.xptsx = xpts; p.yptsx = ypts;
#pragma omp parallel for schedule(static)
      for (i64 n = 0; n < no; n++) {
        if (valid_xy(xpts[n], ypts[n])) {
          p.nxy[n] = field_pos(gin, f_nint<R>(xpts[n]), f_nint<R>(ypts[n]));
          if (VEC && p.nxy[n] > 0) {
            R cm, sm;
            const i64 s = p.nxy[n] - 1;
            movect<R>(p.rlai[s], p.rloi[s], rlat[n], rlon[n], cm, sm);
            p.cxy[n] = cm * p.croi[s] + sm * p.sroi[s];
            p.sxy[n] = sm * p.croi[s] - cm * p.sroi[s];
          }
        } else p.nxy[n] = 0;
      }
      p.din = gin.desc; p.dout = gout.desc; p.built = true;
    }
  }
  if (iret == 0 && p.built && p.iretx == 0) {
    if (!to_station) {
      no = p.nox;
      for (i64 n = 0; n < no; n++) {
        rlon[n] = p.rlonx[n]; rlat[n] = p.rlatx[n];
        if (VEC) { crot[n] = p.crotx[n]; srot[n] = p.srotx[n]; }
      }
    }
    const i64 nktot = no * km;
#pragma omp parallel for schedule(static)
    for (i64 nk = 0; nk < nktot; nk++) {
      const i64 k = nk / no, n = nk - no * k;
      const size_t oa = n + k * mo;
      go[oa] = 0; if (VEC) vo[oa] = 0;
      lo[oa] = 0;
      if (p.nxy[n] > 0) {
        const size_t ia = (size_t)(p.nxy[n] - 1) + (size_t)k * mi;
        if (ibi[k] == 0 || li[ia]) {
          if (!VEC) go[oa] = gi[ia];
          else {
            R urot = p.cxy[n] * gi[ia] - p.sxy[n] * vi[ia];
            R vrot = p.sxy[n] * gi[ia] + p.cxy[n] * vi[ia];
            go[oa] = crot[n] * urot - srot[n] * vrot;
            vo[oa] = srot[n] * urot + crot[n] * vrot;
          }
          lo[oa] = 1;
        } else if (mspiral > 1) {
          const R x = p.xptsx[n], y = p.yptsx[n];
          i64 i1 = f_nint<R>(x), j1 = f_nint<R>(y);
          i64 ixs = f_int<R>(f_sign<R>(RL(1.), x - R(i1))), jxs = f_int<R>(f_sign<R>(RL(1.), y - R(j1)));
          for (i64 mx = 2; mx <= mspiral * mspiral; mx++) {
            i64 ix, jx;
            spiral_step<R>(mx, i1, j1, ixs, jxs, ix, jx);
            i64 nxx = field_pos(gin, ix, jx);
            if (nxx > 0) {
              const size_t ib_ = (size_t)(nxx - 1) + (size_t)k * mi;
              if (li[ib_]) {
                if (!VEC) go[oa] = gi[ib_];
                else {
                  R cm, sm;
                  movect<R>(p.rlai[nxx - 1], p.rloi[nxx - 1], rlat[n], rlon[n], cm, sm);
                  R cx = cm * p.croi[nxx - 1] + sm * p.sroi[nxx - 1];
                  R sx = sm * p.croi[nxx - 1] - cm * p.sroi[nxx - 1];
                  R urot = cx * gi[ib_] - sx * vi[ib_];
                  R vrot = sx * gi[ib_] + cx * vi[ib_];
                  go[oa] = crot[n] * urot - srot[n] * vrot;
                  vo[oa] = srot[n] * urot + crot[n] * vrot;
                }
                lo[oa] = 1;
                break;
              }
            }
          }
        }
      }
    }
    set_ibo<R>(no, mo, km, ibi, lo, ibo);
    if (gout.type == G_EQUID) {
      if (!VEC) polfixs<R>(no, mo, km, rlat, ibo, lo, go);
      else polfixv<R>(no, mo, km, rlat, rlon, ibo, lo, go, vo);
    }
  } else {
    if (iret == 0) iret = p.iretx > 0 ? p.iretx : 1;
    if (!to_station) no = 0;
  }
}

// =================================================================================================
// budget: src/budget_interp_mod.F90:94-351 (scalar), :423-712 (vector)
// neighbor-budget: src/neighbor_budget_interp_mod.F90:107-253 (scalar), :351-549 (vector)
// NB (neighbor-budget) = true selects the nearest-tap variant.
// =================================================================================================
template <class R, bool NBUD, bool VEC>
void interp_budget(const i64* ipopt, const Grid<R>& gin, const Grid<R>& gout, i64 mi, i64 mo, i64 km, const i64* ibi,
                   const u8* li, const R* gi, const R* vi, i64& no, R* rlat, R* rlon, R* crot, R* srot, i64* ibo,
                   u8* lo, R* go, R* vo, i64& iret) {
  const R fill = RL(-9999.), tiny = M<R>::tiny();
  const bool to_station = gout.type == G_STATION;
  iret = 0;
  std::vector<R> xpts(mo, fill), ypts(mo, fill);
  {
    GdsOpt<R> o; if (VEC) { o.crot = crot; o.srot = srot; }
    if (!gdswzd_grid<R>(gout, 0, mo, fill, xpts.data(), ypts.data(), rlon, rlat, no, o)) no = 0;
    // budget vector (non-station) has no NO==0 test (budget_interp_mod.F90:489-491)
    if (no == 0 && !(VEC && !NBUD && !to_station)) iret = 3;
    if (to_station) {
      i64 nv = 0;
      gdswzd_grid<R>(gin, -1, no, fill, xpts.data(), ypts.data(), rlon, rlat, nv, o);  // vector: overwrites crot/srot
      if (nv == 0) iret = 2;
    }
  }
  Plan<R> geo;
  if (VEC) input_geo<R>(gin, mi, geo);
  if (!NBUD && !VEC && ipopt[0] > 16) iret = 32;
  const i64 mspiral = std::max<i64>(ipopt[19], 1);
  i64 nb1 = ipopt[0];
  if (nb1 == -1) nb1 = 2;
  if (iret == 0 && nb1 < 0) iret = 32;
  int lsw = 1;
  if (!NBUD && ipopt[1] == -2) lsw = 2;
  if (ipopt[0] == -1 || ipopt[1] == -1) lsw = 0;
  if (iret == 0 && lsw == 1 && nb1 > 15) iret = 32;
  i64 mpi = 3 + ipopt[0];  // 1-based index into ipopt
  i64 mp = (mpi >= 1 && mpi <= 20) ? ipopt[mpi - 1] : 50;
  if (mp == -1 || mp == 0) mp = 50;
  if (mp < 0 || mp > 100) iret = 32;
  const R pmp = R(mp) * RL(0.01);
  i64 nb2 = 0, nb3 = 0, nb4 = 0;
  R rb2 = 0;
  if (iret == 0) {
    nb2 = 2 * nb1 + 1;
    rb2 = RL(1.) / R(nb2);
    nb3 = nb2 * nb2;
    nb4 = nb3;
    if (lsw == 2) { rb2 = RL(1.) / R(nb1 + 1); nb4 = (nb1 + 1) * (nb1 + 1) * (nb1 + 1) * (nb1 + 1); }
    else if (lsw == 1) { nb4 = ipopt[1]; for (i64 ib = 1; ib <= nb1; ib++) nb4 = nb4 + 8 * ib * ipopt[1 + ib]; }
  } else { nb2 = 0; nb3 = 0; nb4 = NBUD ? 0 : 1; }
  std::vector<R> wo((size_t)mo * km, 0);
  for (i64 k = 0; k < km; k++)
    for (i64 n = 0; n < no; n++) { go[n + k * mo] = 0; if (VEC) vo[n + k * mo] = 0; wo[n + k * mo] = 0; }
  std::vector<R> xptb(no > 0 ? no : 1), yptb(no > 0 ? no : 1), rlob(no > 0 ? no : 1), rlab(no > 0 ? no : 1);
  const int NT = NBUD ? 1 : 4;
  std::vector<i64> nn((size_t)NT * (no > 0 ? no : 1));
  std::vector<R> ww((size_t)NT * (no > 0 ? no : 1)), cc, ss;
  if (VEC) { cc.resize(nn.size()); ss.resize(nn.size()); }
  for (i64 nb = 1; nb <= nb3; nb++) {
    i64 jb = (nb - 1) / nb2 - nb1;
    i64 ib = nb - (jb + nb1) * nb2 - nb1 - 1;
    i64 lb = std::max(std::llabs(ib), std::llabs(jb));
    R wb = 1;
    if (!NBUD && !VEC) {
      if (lsw == 2) wb = R((nb1 + 1 - std::llabs(ib)) * (nb1 + 1 - std::llabs(jb)));
      else if (lsw == 1) wb = R(ipopt[1 + lb]);
    } else if (!NBUD && VEC) {  // budget_interp_mod.F90:571-575 tests IPOPT(2) directly
      if (ipopt[1] == -2) wb = R((nb1 + 1 - std::llabs(ib)) * (nb1 + 1 - std::llabs(jb)));
      else if (ipopt[1] != -1) wb = R(ipopt[1 + lb]);
    } else {
      if (lsw == 1) wb = R(ipopt[1 + lb]);
    }
    if (!(M<R>::fabs(wb) > tiny)) continue;
#pragma omp parallel for schedule(static)
    for (i64 n = 0; n < no; n++) {
      if (!NBUD) { xptb[n] = xpts[n] + R(ib) * rb2; yptb[n] = ypts[n] + R(jb) * rb2; }
      else { xptb[n] = xpts[n] + R(ib) / R(nb2); yptb[n] = ypts[n] + R(jb) / R(nb2); }
    }
    i64 nv = 0;
    GdsOpt<R> none;
    gdswzd_grid<R>(to_station ? gin : gout, 1, no, fill, xptb.data(), yptb.data(), rlob.data(), rlab.data(), nv, none);
    gdswzd_grid<R>(gin, -1, no, fill, xptb.data(), yptb.data(), rlob.data(), rlab.data(), nv, none);
    if (iret == 0 && nv == 0 && lb == 0) iret = 2;
#pragma omp parallel for schedule(static)
    for (i64 n = 0; n < no; n++) {
      const R xi = xptb[n], yi = yptb[n];
      i64* nx = &nn[(size_t)NT * n];
      R* w_ = &ww[(size_t)NT * n];
      if (valid_xy(xi, yi)) {
        if (!NBUD) {
          i64 i1 = f_int<R>(xi), i2 = i1 + 1, j1 = f_int<R>(yi), j2 = j1 + 1;
          R xf = xi - R(i1), yf = yi - R(j1);
          nx[0] = field_pos(gin, i1, j1); nx[1] = field_pos(gin, i2, j1);
          nx[2] = field_pos(gin, i1, j2); nx[3] = field_pos(gin, i2, j2);
          if (std::min(std::min(nx[0], nx[1]), std::min(nx[2], nx[3])) > 0) {
            if (!VEC) {
              w_[0] = (RL(1.) - xf) * (RL(1.) - yf); w_[1] = xf * (RL(1.) - yf);
              w_[2] = (RL(1.) - xf) * yf; w_[3] = xf * yf;
            } else {
              R wi2 = xf, wi1 = RL(1.) - wi2, wj2 = yf, wj1 = RL(1.) - wj2;
              w_[0] = wi1 * wj1; w_[1] = wi2 * wj1; w_[2] = wi1 * wj2; w_[3] = wi2 * wj2;
              for (int t = 0; t < 4; t++) {
                R cm, sm; const i64 s = nx[t] - 1;
                movect<R>(geo.rlai[s], geo.rloi[s], rlat[n], rlon[n], cm, sm);
                cc[(size_t)4 * n + t] = cm * geo.croi[s] + sm * geo.sroi[s];
                ss[(size_t)4 * n + t] = sm * geo.croi[s] - cm * geo.sroi[s];
              }
            }
          } else { nx[0] = nx[1] = nx[2] = nx[3] = 0; }
        } else {
          nx[0] = field_pos(gin, f_nint<R>(xi), f_nint<R>(yi));
          if (VEC && nx[0] > 0) {
            R cm, sm; const i64 s = nx[0] - 1;
            movect<R>(geo.rlai[s], geo.rloi[s], rlat[n], rlon[n], cm, sm);
            cc[n] = cm * geo.croi[s] + sm * geo.sroi[s];
            ss[n] = sm * geo.croi[s] - cm * geo.sroi[s];
          }
        }
      } else { for (int t = 0; t < NT; t++) nx[t] = 0; }
    }
#pragma omp parallel for schedule(static)
    for (i64 k = 0; k < km; k++) {
      for (i64 n = 0; n < no; n++) {
        const i64* nx = &nn[(size_t)NT * n];
        if (nx[0] <= 0) continue;
        const R* w_ = &ww[(size_t)NT * n];
        const size_t oa = n + k * mo, kb = (size_t)k * mi;
        if (NBUD) {
          const size_t ia = (size_t)(nx[0] - 1) + kb;
          if (ibi[k] == 0 || li[ia]) {
            if (!VEC) go[oa] = go[oa] + wb * gi[ia];
            else {
              R u11 = cc[n] * gi[ia] - ss[n] * vi[ia], v11 = ss[n] * gi[ia] + cc[n] * vi[ia];
              go[oa] = go[oa] + wb * u11; vo[oa] = vo[oa] + wb * v11;
            }
            wo[oa] = wo[oa] + wb;
          }
        } else if (!VEC) {
          if (ibi[k] == 0) {
            R gb = w_[0] * gi[nx[0] - 1 + kb] + w_[1] * gi[nx[1] - 1 + kb] + w_[2] * gi[nx[2] - 1 + kb] + w_[3] * gi[nx[3] - 1 + kb];
            go[oa] = go[oa] + wb * gb;
            wo[oa] = wo[oa] + wb;
          } else {
            for (int t = 0; t < 4; t++)
              if (li[nx[t] - 1 + kb]) {
                go[oa] = go[oa] + wb * w_[t] * gi[nx[t] - 1 + kb];
                wo[oa] = wo[oa] + wb * w_[t];
              }
          }
        } else {
          const R* c_ = &cc[(size_t)4 * n]; const R* s_ = &ss[(size_t)4 * n];
          if (ibi[k] == 0) {
            R u[4], v[4];
            for (int t = 0; t < 4; t++) {
              const size_t ia = nx[t] - 1 + kb;
              u[t] = c_[t] * gi[ia] - s_[t] * vi[ia];
              v[t] = s_[t] * gi[ia] + c_[t] * vi[ia];
            }
            R ub = w_[0] * u[0] + w_[1] * u[1] + w_[2] * u[2] + w_[3] * u[3];
            R vb = w_[0] * v[0] + w_[1] * v[1] + w_[2] * v[2] + w_[3] * v[3];
            go[oa] = go[oa] + wb * ub; vo[oa] = vo[oa] + wb * vb; wo[oa] = wo[oa] + wb;
          } else {
            for (int t = 0; t < 4; t++) {
              const size_t ia = nx[t] - 1 + kb;
              if (li[ia]) {
                R u = c_[t] * gi[ia] - s_[t] * vi[ia], v = s_[t] * gi[ia] + c_[t] * vi[ia];
                go[oa] = go[oa] + wb * w_[t] * u;
                vo[oa] = vo[oa] + wb * w_[t] * v;
                wo[oa] = wo[oa] + wb * w_[t];
              }
            }
          }
        }
      }
    }
  }
  // finish (budget_interp_mod.F90:288-344, :686-705; neighbor_budget_interp_mod.F90:235-246, :524-541)
  for (i64 k = 0; k < km; k++) {
    ibo[k] = ibi[k];
    for (i64 n = 0; n < no; n++) {
      const size_t oa = n + k * mo;
      lo[oa] = wo[oa] >= pmp * R(nb4);
      if (lo[oa]) {
        go[oa] = go[oa] / wo[oa];
        if (VEC) {
          vo[oa] = vo[oa] / wo[oa];
          R urot = crot[n] * go[oa] - srot[n] * vo[oa];
          R vrot = srot[n] * go[oa] + crot[n] * vo[oa];
          go[oa] = urot; vo[oa] = vrot;
        }
      } else if (!NBUD && !VEC && mspiral > 1) {
        R lat1 = rlat[n], lon1 = rlon[n], xx = fill, yy = fill;
        i64 nv = 0;
        GdsOpt<R> none;
        gdswzd_grid<R>(gin, -1, 1, fill, &xx, &yy, &lon1, &lat1, nv, none);
        bool found = false;
        if (nv == 1) {
          i64 i1 = f_nint<R>(xx), j1 = f_nint<R>(yy);
          i64 ixs = f_int<R>(f_sign<R>(RL(1.), xx - R(i1))), jxs = f_int<R>(f_sign<R>(RL(1.), yy - R(j1)));
          for (i64 mx = 2; mx <= mspiral * mspiral; mx++) {
            i64 ix, jx;
            spiral_step<R>(mx, i1, j1, ixs, jxs, ix, jx);
            i64 nxx = field_pos(gin, ix, jx);
            if (nxx > 0) {
              const size_t ia = (size_t)(nxx - 1) + (size_t)k * mi;
              if (li[ia] || ibi[k] == 0) { go[oa] = gi[ia]; lo[oa] = 1; found = true; break; }
            }
          }
        }
        if (!found) { ibo[k] = 1; go[oa] = 0; }
      } else {
        ibo[k] = 1; go[oa] = 0; if (VEC) vo[oa] = 0;
      }
    }
  }
  if (gout.type == G_EQUID) {
    if (!VEC) polfixs<R>(no, mo, km, rlat, ibo, lo, go);
    else polfixv<R>(no, mo, km, rlat, rlon, ibo, lo, go, vo);
  }
}

// ---- ipolates_grid / ipolatev_grid dispatch: src/ipolates.F90:63-112, src/ipolatev.F90:67-119 --
template <class R>
void ipolates_grid(i64 ip, const i64* ipopt, const Grid<R>& gin, const Grid<R>& gout, i64 mi, i64 mo, i64 km,
                   const i64* ibi, const u8* li, const R* gi, i64& no, R* rlat, R* rlon, i64* ibo, u8* lo, R* go, i64& iret) {
  switch (ip) {
    case 0: interp_stencil<R, 2, false>(ipopt, gin, gout, mi, mo, km, ibi, li, gi, nullptr, no, rlat, rlon, nullptr, nullptr, ibo, lo, go, nullptr, iret); break;
    case 1: interp_stencil<R, 4, false>(ipopt, gin, gout, mi, mo, km, ibi, li, gi, nullptr, no, rlat, rlon, nullptr, nullptr, ibo, lo, go, nullptr, iret); break;
    case 2: interp_neighbor<R, false>(ipopt, gin, gout, mi, mo, km, ibi, li, gi, nullptr, no, rlat, rlon, nullptr, nullptr, ibo, lo, go, nullptr, iret); break;
    case 3: interp_budget<R, false, false>(ipopt, gin, gout, mi, mo, km, ibi, li, gi, nullptr, no, rlat, rlon, nullptr, nullptr, ibo, lo, go, nullptr, iret); break;
    case 6: interp_budget<R, true, false>(ipopt, gin, gout, mi, mo, km, ibi, li, gi, nullptr, no, rlat, rlon, nullptr, nullptr, ibo, lo, go, nullptr, iret); break;
    default: iret = 1; break;  // ip=4 (spectral) is out of scope; others: reference error-stops
  }
}
template <class R>
void ipolatev_grid(i64 ip, const i64* ipopt, const Grid<R>& gin, const Grid<R>& gout, i64 mi, i64 mo, i64 km,
                   const i64* ibi, const u8* li, const R* ui, const R* vi, i64& no, R* rlat, R* rlon, R* crot, R* srot,
                   i64* ibo, u8* lo, R* uo, R* vo, i64& iret) {
  switch (ip) {
    case 0: interp_stencil<R, 2, true>(ipopt, gin, gout, mi, mo, km, ibi, li, ui, vi, no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret); break;
    case 1: interp_stencil<R, 4, true>(ipopt, gin, gout, mi, mo, km, ibi, li, ui, vi, no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret); break;
    case 2: interp_neighbor<R, true>(ipopt, gin, gout, mi, mo, km, ibi, li, ui, vi, no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret); break;
    case 3: interp_budget<R, false, true>(ipopt, gin, gout, mi, mo, km, ibi, li, ui, vi, no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret); break;
    case 6: interp_budget<R, true, true>(ipopt, gin, gout, mi, mo, km, ibi, li, ui, vi, no, rlat, rlon, crot, srot, ibo, lo, uo, vo, iret); break;
    default: iret = 1; break;
  }
}

}  // namespace ipo
