// ipb200 — device-resident interpolation plans and the kernels that build them.
//
// A plan is everything about a (grid_in, grid_out, method, options) tuple that does not depend on
// the fields: output lat/lon (+ rotations), source x/y, tap indices, tap weights, vector tap
// rotations, budget sub-sample taps, pole-point lists.  It replaces the reference's per-routine
// single-entry SAVE caches (src/bilinear_interp_mod.F90:104-108, bicubic_interp_mod.F90:111-115,
// neighbor_interp_mod.F90:129-132) and — new — also caches the budget family, which the reference
// rebuilds on every call (src/budget_interp_mod.F90:190-249).
//
// Layout in HBM: structure-of-arrays, tap-major ([tap][point]) so that a warp of consecutive
// output points reads each tap array coalesced.
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <memory>
#include <vector>

#include "ipb_decode.h"
#include "ipb_proj.cuh"
#include "ipb_state.h"

namespace ipb {

extern std::atomic<long long> g_launches;  // kernels launched by this library (bench.py's gpu_launches)

#define IPB_CUDA(x)                                                                              \
  do {                                                                                           \
    cudaError_t e_ = (x);                                                                        \
    if (e_ != cudaSuccess) {                                                                     \
      fprintf(stderr, "ipb200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      throw e_;                                                                                  \
    }                                                                                            \
  } while (0)

// Plan arrays are allocated stream-ordered from the device's memory pool while a plan is being built (an AsyncAlloc
// scope of the calling thread, ipb_state.h): a plan is ~40 arrays, and cudaMalloc's round trips were most of the 20 ms
// a first call on a grid pair took; station-point outputs, whose plans are rebuilt on every call, pay it every time.
template <class T> T* dalloc(size_t n) {
  T* p = nullptr;
  const AllocCtx& c = alloc_ctx();
  if (c.async) IPB_CUDA(cudaMallocAsync((void**)&p, (n ? n : 1) * sizeof(T), c.stream));
  else IPB_CUDA(cudaMalloc((void**)&p, (n ? n : 1) * sizeof(T)));
  return p;
}
// Frees go back to the pool stream-ordered as well: on the scope's stream (while a plan is built: the building stream;
// when a plan is destroyed: the owning device's first stream — a plan is only destroyed after every call that used it
// has synchronised its streams).
template <class T> void dfree(T*& p) {
  if (p) {
    const AllocCtx& c = alloc_ctx();
    if (c.async) cudaFreeAsync((void*)p, c.stream);
    else cudaFree((void*)p);
  }
  p = nullptr;
}

inline int nblk(long long n, int t) { return (int)((n + t - 1) / t); }

enum Method { M_BILINEAR = 0, M_BICUBIC = 1, M_NEIGHBOR = 2, M_BUDGET = 3, M_NBUDGET = 6 };

// Option block of the budget family (decoded on the host exactly as budget_interp_mod.F90:151-181,
// neighbor_budget_interp_mod.F90:158-180).
struct BudgetOpt {
  int nb1 = 0, nb2 = 0, nb3 = 0, lsw = 0;
  long long nb4 = 1, mp = 50;
  int nsamp = 0;                     // samples with non-zero weight
  std::vector<int> ib, jb, lb;       // per kept sample
  std::vector<double> wb;            // per kept sample (small integers; exact in either kind)
};

// Per-tile cell tables of the staged bilinear kernel (ipb_tiles.cuh), one set per alignment variant.
struct TileTables {
  bool ok = false;
  int ntile = 0, cstride = 0, max_cells = 0;
  u8* cid[4] = {nullptr, nullptr, nullptr, nullptr};     // [ntile*128] cell number of each point inside its tile (0xff: none)
  int* ncell[4] = {nullptr, nullptr, nullptr, nullptr};  // [ntile]
  int* cells[4] = {nullptr, nullptr, nullptr, nullptr};  // [ntile][cstride] staging list: 0-based source position of each entry, sorted (-1 = padding)
  u8* edst[4] = {nullptr, nullptr, nullptr, nullptr};    // [ntile][cstride] shared-memory slot (cell*4 + tap) of each entry
  size_t bytes = 0;
};
// Point-staged kernels (ipb_gtiles.cuh): tile shape and the per-tile lists of distinct source positions.
struct GtGeom { int mode = 0, rowlen = 0, ntx = 1, ntile = 0; };   // mode 0: 128 consecutive points; 1: 32x4 patch; 2: 16x8 patch
struct GtTables {
  bool ok = false;
  GtGeom geom;
  int U = 0, estride = 0, maxent = 0, novf = 0;
  unsigned long long entries = 0;   // distinct positions summed over tiles (= staging copies per field)
  int* esrc = nullptr;              // [ntile][estride] sorted 0-based source positions, -1 = padding
  u8* slot = nullptr;               // [U][no] position of every tap in its tile's list
  unsigned short* slot16 = nullptr; // the same for lists of up to 512 positions (vector bilinear)
  size_t bytes = 0;
};
struct TileDev { const u8* cid[4]; const int* ncell[4]; const int* cells[4]; const u8* edst[4]; int cstride; int ntile; };

template <class R> struct Plan {
  std::vector<i64> key;
  int method = 0; bool vec = false; bool to_station = false;
  GridHost<R> gin, gout;
  R *d_tab_in = nullptr, *d_tab_out = nullptr;  // Gaussian tables (3 x (jg+2))
  int mi = 0, mo = 0, no = 0, nv = 0, iret = 0;
  R *rlat = nullptr, *rlon = nullptr, *crot = nullptr, *srot = nullptr;  // [mo]
  R *x = nullptr, *y = nullptr;                                          // [mo] source-grid coordinates
  // stencil / neighbor
  int nt = 0;                 // taps per point: 4 (bilinear), 16 (bicubic), 1 (neighbor)
  int* nxy = nullptr;         // [nt][no]
  R* wxy = nullptr;           // [nt][no]
  R *cxy = nullptr, *sxy = nullptr;  // [nt][no] (vector)
  u8* nc = nullptr;           // [no] bicubic stencil class (0 none, 1 bicubic, 2 bilinear fallback)
  // budget family
  BudgetOpt bo;
  int* bn = nullptr;          // [nsamp][bt][no], bt = 4 (budget) or 1 (neighbor-budget)
  R* bw = nullptr;            // [nsamp][4][no] (budget only)
  R *bc = nullptr, *bs = nullptr;  // [nsamp][bt][no] (vector)
  R* d_wb = nullptr;          // [nsamp]
  // collapsed budget taps of the binary64 kinds (ipb_budget.cuh): cu = taps per point in use (0 = not available)
  int cu = 0;
  int* cn = nullptr;          // [16][no] distinct source positions (padded with a valid position)
  R *cw = nullptr, *cwsum = nullptr;  // [16][no] summed weights (padding 0); [no] sum of the sample weights
  // poles (lat-lon outputs): POLFIXS / POLFIXV
  int npn = 0, nps = 0;
  int *pole_n = nullptr, *pole_s = nullptr;  // 0-based output indices, ascending
  R *pclon_n = nullptr, *pslon_n = nullptr, *pclon_s = nullptr, *pslon_s = nullptr;
  // 1-based range of input-field positions any tap refers to (host calls copy only this window)
  int src_lo = 1, src_hi = 0;
  TileTables tiles;
  GtTables gt;
  bool gt_tried = false;   // bilinear: the point-staged tables are built on the first call that carries a bitmap
  R *wq = nullptr, *ws2 = nullptr;   // staged bilinear kernel: weights per point [no][4], (sum, 1/sum) [no][2]
  bool box_ok = false;               // TMA-staged bilinear kernel (ipb_box.cuh)
  double box_wmin = 0;               // smallest weight sum over the existing taps of any point (lo = 1 everywhere iff >= pmp)
  int* bxy = nullptr;                // [no] source cell of every point (iy << 16 | ix), -1 = irregular stencil
  size_t bytes = 0;
  double build_ms = 0;
  int device = 0;                    // the device the arrays live on, and the stream their frees are ordered on
  cudaStream_t free_stream = nullptr;

  ~Plan() {
    DeviceGuard guard(device);
    AsyncAlloc scope(free_stream);
    dfree(d_tab_in); dfree(d_tab_out); dfree(rlat); dfree(rlon); dfree(crot); dfree(srot); dfree(x); dfree(y);
    dfree(nxy); dfree(wxy); dfree(cxy); dfree(sxy); dfree(nc); dfree(bn); dfree(bw); dfree(bc); dfree(bs); dfree(d_wb); dfree(cn); dfree(cw); dfree(cwsum);
    dfree(pole_n); dfree(pole_s); dfree(pclon_n); dfree(pslon_n); dfree(pclon_s); dfree(pslon_s);
    for (int v = 0; v < 4; v++) { dfree(tiles.cid[v]); dfree(tiles.ncell[v]); dfree(tiles.cells[v]); dfree(tiles.edst[v]); }
    dfree(wq); dfree(ws2); dfree(bxy);
    dfree(gt.esrc); dfree(gt.slot); dfree(gt.slot16);
  }
};

// ---- kernels ------------------------------------------------------------------------------------

// Output-point geometry + source coordinates: gdswzd(out,0) followed by gdswzd(in,-1)
// (e.g. bilinear_interp_mod.F90:146-150).  mode 0: x/y = inverse on the input grid (stencil
// methods, and station points in the budget family); mode 1: x/y = the output grid's own
// coordinates (budget family, budget_interp_mod.F90:145).  One thread per output point.
template <class R, bool VEC>
__global__ void k_out_geo(GridP<R> gin, GridP<R> gout, int mo, int nm, int no, int station, int mode, int vec_inv_rot,
                          R* rlat, R* rlon, R* crot, R* srot, R* xo, R* yo, int* nv) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const R fill = IPB_DL(-9999.);
  bool ok = false;
  if (n < mo) {
    R x = fill, y = fill, lon, lat;
    PtExt<R> e;
    if (station) { lon = rlon[n]; lat = rlat[n]; }
    else {
      if (n < nm) enumerate_xy(gout, n + 1, x, y);
      proj_fwd<R, VEC ? EXT_ROT : EXT_NONE>(gout, fill, x, y, lon, lat, e);
      rlon[n] = lon; rlat[n] = lat;
      if (VEC) { crot[n] = e.crot; srot[n] = e.srot; }
    }
    if (mode == 0 || station) {
      R xi = fill, yi = fill;
      if (n < no) {
        ok = proj_inv<R, VEC ? EXT_ROT : EXT_NONE>(gin, fill, lon, lat, xi, yi, e);
        // budget-family vector calls pass crot/srot to the inverse on station points
        // (budget_interp_mod.F90:484, neighbor_budget_interp_mod.F90:424): they are overwritten.
        if (VEC && vec_inv_rot) { crot[n] = e.crot; srot[n] = e.srot; }
      }
      xo[n] = xi; yo[n] = yi;
    } else {
      xo[n] = x; yo[n] = y;
    }
  }
  const unsigned b = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(nv, __popc(b));
}

// lat/lon and rotation of an input-grid point given its 1-based field position — what the
// reference tabulates with gdswzd(grid_in,0,mi,...,croi,sroi) (bilinear_interp_mod.F90:414).
template <class R> __device__ __forceinline__ void in_point_geo(const GridP<R>& gin, int npos, R& rla, R& rlo, R& cro, R& sro) {
  const R fill = IPB_DL(-9999.);
  R x, y;
  PtExt<R> e;
  enumerate_xy(gin, npos, x, y);
  proj_fwd<R, EXT_ROT>(gin, fill, x, y, rlo, rla, e);
  cro = e.crot; sro = e.srot;
}
template <class R>
__device__ __forceinline__ void tap_rotation(const GridP<R>& gin, int npos, R tlat, R tlon, R& c, R& s) {
  R rla, rlo, cro, sro, cm, sm;
  in_point_geo(gin, npos, rla, rlo, cro, sro);
  movect<R>(rla, rlo, tlat, tlon, cm, sm);
  c = cm * cro + sm * sro;
  s = sm * cro - cm * sro;
}

// bilinear (NT=2) / bicubic (NT=4) taps and weights: bilinear_interp_mod.F90:165-189 (:427-459 vector),
// bicubic_interp_mod.F90:174-222 (:437-495 vector).
template <class R, int NT, bool VEC>
__global__ void k_plan_stencil(GridP<R> gin, int no, const R* __restrict__ xs, const R* __restrict__ ys,
                               const R* __restrict__ rlat, const R* __restrict__ rlon, int* nxy, R* wxy, R* cxy, R* sxy, u8* nc) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= no) return;
  constexpr int T2 = NT * NT;
  const R x = xs[n], y = ys[n];
  int idx[T2];
  R wx[NT], wy[NT];
  int cls = (NT == 2) ? 1 : 0;
  if (valid_xy(x, y)) {
    if (NT == 2) {
      const int i0 = f_floor<R>(x), j0 = f_floor<R>(y);
      const R xf = x - (R)i0, yf = y - (R)j0;
      wx[0] = IPB_DL(1.) - xf; wx[1] = xf; wy[0] = IPB_DL(1.) - yf; wy[1] = yf;
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int a = 0; a < 2; a++) idx[a + 2 * b] = field_pos(gin, i0 + a, j0 + b);
    } else {
      const int i0 = f_floor<R>(x - IPB_DL(1.)), j0 = f_floor<R>(y - IPB_DL(1.));
      const R xf = x - (R)(i0 + 1), yf = y - (R)(j0 + 1);
      bool full = true;
#pragma unroll
      for (int b = 0; b < 4; b++)
#pragma unroll
        for (int a = 0; a < 4; a++) { idx[a + 4 * b] = field_pos(gin, i0 + a, j0 + b); full = full && idx[a + 4 * b] > 0; }
      if (full) {
        cls = 1;
        wx[0] = xf * (IPB_DL(1.) - xf) * (IPB_DL(2.) - xf) / IPB_DL(-6.);
        wx[1] = (xf + IPB_DL(1.)) * (IPB_DL(1.) - xf) * (IPB_DL(2.) - xf) / IPB_DL(2.);
        wx[2] = (xf + IPB_DL(1.)) * xf * (IPB_DL(2.) - xf) / IPB_DL(2.);
        wx[3] = (xf + IPB_DL(1.)) * xf * (IPB_DL(1.) - xf) / IPB_DL(-6.);
        wy[0] = yf * (IPB_DL(1.) - yf) * (IPB_DL(2.) - yf) / IPB_DL(-6.);
        wy[1] = (yf + IPB_DL(1.)) * (IPB_DL(1.) - yf) * (IPB_DL(2.) - yf) / IPB_DL(2.);
        wy[2] = (yf + IPB_DL(1.)) * yf * (IPB_DL(2.) - yf) / IPB_DL(2.);
        wy[3] = (yf + IPB_DL(1.)) * yf * (IPB_DL(1.) - yf) / IPB_DL(-6.);
      } else {
        cls = 2;
        wx[0] = 0; wx[1] = IPB_DL(1.) - xf; wx[2] = xf; wx[3] = 0;
        wy[0] = 0; wy[1] = IPB_DL(1.) - yf; wy[2] = yf; wy[3] = 0;
      }
    }
  } else {
#pragma unroll
    for (int t = 0; t < T2; t++) idx[t] = 0;
#pragma unroll
    for (int a = 0; a < NT; a++) { wx[a] = 0; wy[a] = 0; }
  }
  if (nc) nc[n] = (u8)cls;
  const R tlat = VEC ? rlat[n] : (R)0, tlon = VEC ? rlon[n] : (R)0;
#pragma unroll
  for (int b = 0; b < NT; b++)
#pragma unroll
    for (int a = 0; a < NT; a++) {
      const int t = a + NT * b;
      nxy[(size_t)t * no + n] = idx[t];
      wxy[(size_t)t * no + n] = wx[a] * wy[b];
      if (VEC) {
        R c = 0, s = 0;
        if (idx[t] > 0) tap_rotation(gin, idx[t], tlat, tlon, c, s);
        cxy[(size_t)t * no + n] = c; sxy[(size_t)t * no + n] = s;
      }
    }
}

// neighbor: neighbor_interp_mod.F90:185-199 (:447-470 vector)
template <class R, bool VEC>
__global__ void k_plan_neighbor(GridP<R> gin, int no, const R* __restrict__ xs, const R* __restrict__ ys,
                                const R* __restrict__ rlat, const R* __restrict__ rlon, int* nxy, R* cxy, R* sxy) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= no) return;
  const R x = xs[n], y = ys[n];
  int p = 0;
  if (valid_xy(x, y)) p = field_pos(gin, f_nint<R>(x), f_nint<R>(y));
  nxy[n] = p;
  if (VEC) {
    R c = 0, s = 0;
    if (p > 0) tap_rotation(gin, p, rlat[n], rlon[n], c, s);
    cxy[n] = c; sxy[n] = s;
  }
}

// budget / neighbor-budget sub-sample taps.  blockIdx.y = kept sample.
// budget_interp_mod.F90:202-249 (:563-625 vector), neighbor_budget_interp_mod.F90:195-219 (:482-505 vector).
// pgrid is the grid the shifted sample is projected forward on: the output grid, or the input
// grid when interpolating to station points.
template <class R, bool NBUD, bool VEC>
__global__ void k_plan_budget(GridP<R> gin, GridP<R> pgrid, int no, const R* __restrict__ xs, const R* __restrict__ ys,
                              const R* __restrict__ rlat, const R* __restrict__ rlon, const int* __restrict__ sib,
                              const int* __restrict__ sjb, R rb2, int nb2, int* bn, R* bw, R* bc, R* bs, int* nv_center) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int s = blockIdx.y;
  constexpr int BT = NBUD ? 1 : 4;
  const R fill = IPB_DL(-9999.);
  bool ok = false;
  const int ib = sib[s], jb = sjb[s];
  if (n < no) {
    R xb, yb;
    if (!NBUD) { xb = xs[n] + (R)ib * rb2; yb = ys[n] + (R)jb * rb2; }
    else { xb = xs[n] + (R)ib / (R)nb2; yb = ys[n] + (R)jb / (R)nb2; }
    R lon, lat, xi, yi;
    PtExt<R> e;
    proj_fwd<R, EXT_NONE>(pgrid, fill, xb, yb, lon, lat, e);
    ok = proj_inv<R, EXT_NONE>(gin, fill, lon, lat, xi, yi, e);
    int idx[BT];
    R w[BT];
#pragma unroll
    for (int t = 0; t < BT; t++) { idx[t] = 0; w[t] = 0; }
    if (valid_xy(xi, yi)) {
      if (!NBUD) {
        const int i1 = f_int<R>(xi), j1 = f_int<R>(yi);
        const R xf = xi - (R)i1, yf = yi - (R)j1;
        idx[0] = field_pos(gin, i1, j1); idx[1] = field_pos(gin, i1 + 1, j1);
        idx[2] = field_pos(gin, i1, j1 + 1); idx[3] = field_pos(gin, i1 + 1, j1 + 1);
        if (min(min(idx[0], idx[1]), min(idx[2], idx[3])) > 0) {
          const R a1 = IPB_DL(1.) - xf, b1 = IPB_DL(1.) - yf;
          w[0] = a1 * b1; w[1] = xf * b1; w[2] = a1 * yf; w[3] = xf * yf;
        } else { idx[0] = idx[1] = idx[2] = idx[3] = 0; }
      } else {
        idx[0] = field_pos(gin, f_nint<R>(xi), f_nint<R>(yi));
      }
    }
    const size_t base = ((size_t)s * BT) * no + n;
#pragma unroll
    for (int t = 0; t < BT; t++) {
      bn[base + (size_t)t * no] = idx[t];
      if (!NBUD) bw[base + (size_t)t * no] = w[t];
      if (VEC) {
        R c = 0, sn = 0;
        if (idx[t] > 0) tap_rotation(gin, idx[t], rlat[n], rlon[n], c, sn);
        bc[base + (size_t)t * no] = c; bs[base + (size_t)t * no] = sn;
      }
    }
  }
  if (ib == 0 && jb == 0) {
    const unsigned b = __ballot_sync(0xffffffffu, ok);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(nv_center, __popc(b));
  }
}

// cos/sin of longitude for the pole points of a lat-lon output (polfix_mod.F90:145-149)
template <class R> __global__ void k_pole_trig(int np, const int* __restrict__ idx, const R* __restrict__ rlon, R* clon, R* slon) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= np) return;
  const R dpr = IPB_DL(180.) / IPB_DL(3.14159265358979);
  const R l = rlon[idx[q]];
  clon[q] = DM<R>::cos(l / dpr);
  slon[q] = DM<R>::sin(l / dpr);
}

// Range of the 1-based input positions referenced by a tap array (zeros = "outside" are skipped).
static __global__ void k_idx_range(const int* __restrict__ idx, size_t n, int* lohi) {
  int lo = 0x7fffffff, hi = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int v = idx[i];
    if (v > 0) { lo = min(lo, v); hi = max(hi, v); }
  }
  for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
  if ((threadIdx.x & 31) == 0 && hi > 0) { atomicMin(lohi, lo); atomicMax(lohi + 1, hi); }
}

// ---- host: option decode ---------------------------------------------------------------------
// Returns iret contribution (0 or 32) and fills bo.  `vec_wb_quirk`: budget vector picks WB from
// IPOPT(2) directly (budget_interp_mod.F90:571-575) instead of LSW.
inline int decode_budget_opt(const i64* ipopt, bool nbud, bool vec, int iret_in, BudgetOpt& bo) {
  int iret = iret_in;
  if (!nbud && !vec && ipopt[0] > 16) iret = 32;
  i64 nb1 = ipopt[0];
  if (nb1 == -1) nb1 = 2;
  if (iret == 0 && nb1 < 0) iret = 32;
  int lsw = 1;
  if (!nbud && ipopt[1] == -2) lsw = 2;
  if (ipopt[0] == -1 || ipopt[1] == -1) lsw = 0;
  if (iret == 0 && lsw == 1 && nb1 > 15) iret = 32;
  const i64 q = 3 + ipopt[0];  // MP=IPOPT(3+IPOPT(1)): indexes IPOPT(2) when ipopt(1)=-1
  i64 mp = (q >= 1 && q <= 20) ? ipopt[q - 1] : 50;
  if (mp == -1 || mp == 0) mp = 50;
  if (mp < 0 || mp > 100) iret = 32;
  bo = BudgetOpt();
  bo.lsw = lsw; bo.mp = mp;
  if (iret != 0) { bo.nb1 = 0; bo.nb2 = 0; bo.nb3 = 0; bo.nb4 = nbud ? 0 : 1; return iret; }
  bo.nb1 = (int)nb1; bo.nb2 = 2 * bo.nb1 + 1; bo.nb3 = bo.nb2 * bo.nb2; bo.nb4 = bo.nb3;
  if (lsw == 2) { i64 q = nb1 + 1; bo.nb4 = q * q * q * q; }
  else if (lsw == 1) { bo.nb4 = ipopt[1]; for (i64 ib = 1; ib <= nb1; ib++) bo.nb4 += 8 * ib * ipopt[1 + ib]; }
  for (int nb = 1; nb <= bo.nb3; nb++) {
    const int jb = (nb - 1) / bo.nb2 - bo.nb1;
    const int ib = nb - (jb + bo.nb1) * bo.nb2 - bo.nb1 - 1;
    const int lb = std::max(std::abs(ib), std::abs(jb));
    double wb = 1;
    if (!nbud && !vec) {
      if (lsw == 2) wb = (double)((bo.nb1 + 1 - std::abs(ib)) * (bo.nb1 + 1 - std::abs(jb)));
      else if (lsw == 1) wb = (double)ipopt[1 + lb];
    } else if (!nbud && vec) {
      if (ipopt[1] == -2) wb = (double)((bo.nb1 + 1 - std::abs(ib)) * (bo.nb1 + 1 - std::abs(jb)));
      else if (ipopt[1] != -1) wb = (double)ipopt[1 + lb];
    } else if (lsw == 1) wb = (double)ipopt[1 + lb];
    if (wb == 0) continue;  // ABS(WB) > TINY
    bo.ib.push_back(ib); bo.jb.push_back(jb); bo.lb.push_back(lb); bo.wb.push_back(wb);
  }
  bo.nsamp = (int)bo.ib.size();
  return iret;
}

template <class R> void upload_tables(GridHost<R>& gh, R*& dptr) {
  if (gh.p.type != P_GAUSS) return;
  const size_t n = gh.alat.size();
  dptr = dalloc<R>(3 * n);
  // on the allocating stream (the array is stream-ordered); the host vectors are pageable, so each copy has left the
  // host buffer when the call returns
  cudaStream_t st = alloc_ctx().async ? alloc_ctx().stream : (cudaStream_t)0;
  IPB_CUDA(cudaMemcpyAsync(dptr, gh.alat.data(), n * sizeof(R), cudaMemcpyHostToDevice, st));
  IPB_CUDA(cudaMemcpyAsync(dptr + n, gh.blat.data(), n * sizeof(R), cudaMemcpyHostToDevice, st));
  IPB_CUDA(cudaMemcpyAsync(dptr + 2 * n, gh.ylat_row.data(), n * sizeof(R), cudaMemcpyHostToDevice, st));
  gh.p.alat = dptr; gh.p.blat = dptr + n; gh.p.ylat_row = dptr + 2 * n;
}

// Number of points the output grid enumerates for gdswzd iopt=0 (src/gdswzd_mod.F90:127-143).
template <class R> inline i64 grid_nm(const GridP<R>& g, i64 npts) {
  if (g.type == P_STATION) return g.grid_num == -1 ? npts : 0;
  return (i64)g.im * g.jm;
}

// ---- host: build a plan -----------------------------------------------------------------------
// station_lat/lon(/crot/srot): host arrays of the caller (only read when the output is station points).
template <class R>
Plan<R>* build_plan(int method, bool vec, const i64* ipopt, const GridHost<R>& gin_h, const GridHost<R>& gout_h, i64 mi,
                    i64 mo, i64 no_in, const R* st_lat, const R* st_lon, const R* st_crot, const R* st_srot,
                    cudaStream_t st) {
  struct Ev { cudaEvent_t e = nullptr; Ev() { cudaEventCreate(&e); } ~Ev() { if (e) cudaEventDestroy(e); } } e0, e1;   // released on every path
  cudaEventRecord(e0.e, st);
  std::unique_ptr<Plan<R>> holder(new Plan<R>());   // a CUDA failure below throws: the plan and its device arrays go with it
  Plan<R>* P = holder.get();
  cudaGetDevice(&P->device); P->free_stream = st;
  P->method = method; P->vec = vec;
  P->gin = gin_h; P->gout = gout_h;
  upload_tables(P->gin, P->d_tab_in);
  upload_tables(P->gout, P->d_tab_out);
  const GridP<R>& gi = P->gin.p;
  const GridP<R>& go = P->gout.p;
  const bool station = go.type == P_STATION;
  const bool budget = (method == M_BUDGET || method == M_NBUDGET);
  const bool nbud = method == M_NBUDGET;
  P->to_station = station;
  P->mi = (int)mi; P->mo = (int)mo;
  const int T = 256;
  int iret = 0;

  // number of output points (gdswzd iopt=0 on the output grid)
  i64 nm = grid_nm(go, mo);
  i64 no;
  if (nm > mo) no = 0;  // reference: arrays filled, NO untouched; we report 0
  else if (station) no = mo;
  else if ((go.type == P_ROTB || go.type == P_ROTE) && go.rerth < 0) no = 0;
  else no = nm;
  (void)no_in;
  P->no = (int)no;
  // the budget vector routine has no NO==0 test for non-station outputs (budget_interp_mod.F90:489-491)
  if (no == 0 && !(vec && method == M_BUDGET && !station)) iret = 3;

  P->rlat = dalloc<R>(mo); P->rlon = dalloc<R>(mo); P->x = dalloc<R>(mo); P->y = dalloc<R>(mo);
  if (vec) { P->crot = dalloc<R>(mo); P->srot = dalloc<R>(mo); }
  P->bytes += (size_t)mo * sizeof(R) * (vec ? 6 : 4);
  if (station) {
    IPB_CUDA(cudaMemcpyAsync(P->rlat, st_lat, mo * sizeof(R), cudaMemcpyDefault, st));
    IPB_CUDA(cudaMemcpyAsync(P->rlon, st_lon, mo * sizeof(R), cudaMemcpyDefault, st));
    if (vec) {
      IPB_CUDA(cudaMemcpyAsync(P->crot, st_crot, mo * sizeof(R), cudaMemcpyDefault, st));
      IPB_CUDA(cudaMemcpyAsync(P->srot, st_srot, mo * sizeof(R), cudaMemcpyDefault, st));
    }
  }
  int* d_cnt = dalloc<int>(2);
  IPB_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(int), st));
  if (mo > 0) {
    const int mode = (budget && !station) ? 1 : 0;
    const int vec_inv_rot = (budget && station && vec) ? 1 : 0;
    if (vec) k_out_geo<R, true><<<nblk(mo, T), T, 0, st>>>(gi, go, (int)mo, (int)(nm > mo ? 0 : nm), (int)no, station, mode, vec_inv_rot, P->rlat, P->rlon, P->crot, P->srot, P->x, P->y, d_cnt);
    else k_out_geo<R, false><<<nblk(mo, T), T, 0, st>>>(gi, go, (int)mo, (int)(nm > mo ? 0 : nm), (int)no, station, mode, 0, P->rlat, P->rlon, nullptr, nullptr, P->x, P->y, d_cnt);
    g_launches++;
  }
  int h_cnt[2] = {0, 0};
  if (!budget || station) {
    IPB_CUDA(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(int), cudaMemcpyDeviceToHost, st));
    IPB_CUDA(cudaStreamSynchronize(st));
    P->nv = h_cnt[0];
    if (budget) { if (P->nv == 0) iret = 2; }          // budget_interp_mod.F90:137-138 (station branch)
    else if (iret == 0 && P->nv == 0) iret = 2;        // bilinear_interp_mod.F90:150
  }

  if (!budget) {
    if (iret == 0) {
      if (method == M_NEIGHBOR) {
        P->nt = 1;
        P->nxy = dalloc<int>(no);
        if (vec) { P->cxy = dalloc<R>(no); P->sxy = dalloc<R>(no); }
        P->bytes += (size_t)no * (4 + (vec ? 2 * sizeof(R) : 0));
        if (vec) k_plan_neighbor<R, true><<<nblk(no, T), T, 0, st>>>(gi, (int)no, P->x, P->y, P->rlat, P->rlon, P->nxy, P->cxy, P->sxy);
        else k_plan_neighbor<R, false><<<nblk(no, T), T, 0, st>>>(gi, (int)no, P->x, P->y, P->rlat, P->rlon, P->nxy, nullptr, nullptr);
        g_launches++;
      } else {
        const int nt = method == M_BILINEAR ? 4 : 16;
        P->nt = nt;
        P->nxy = dalloc<int>((size_t)nt * no); P->wxy = dalloc<R>((size_t)nt * no);
        if (method == M_BICUBIC) P->nc = dalloc<u8>(no);
        if (vec) { P->cxy = dalloc<R>((size_t)nt * no); P->sxy = dalloc<R>((size_t)nt * no); }
        P->bytes += (size_t)nt * no * (4 + sizeof(R) + (vec ? 2 * sizeof(R) : 0)) + no;
        const int TB = 128;
        if (method == M_BILINEAR) {
          if (vec) k_plan_stencil<R, 2, true><<<nblk(no, TB), TB, 0, st>>>(gi, (int)no, P->x, P->y, P->rlat, P->rlon, P->nxy, P->wxy, P->cxy, P->sxy, P->nc);
          else k_plan_stencil<R, 2, false><<<nblk(no, TB), TB, 0, st>>>(gi, (int)no, P->x, P->y, P->rlat, P->rlon, P->nxy, P->wxy, nullptr, nullptr, P->nc);
        } else {
          if (vec) k_plan_stencil<R, 4, true><<<nblk(no, TB), TB, 0, st>>>(gi, (int)no, P->x, P->y, P->rlat, P->rlon, P->nxy, P->wxy, P->cxy, P->sxy, P->nc);
          else k_plan_stencil<R, 4, false><<<nblk(no, TB), TB, 0, st>>>(gi, (int)no, P->x, P->y, P->rlat, P->rlon, P->nxy, P->wxy, nullptr, nullptr, P->nc);
        }
        g_launches++;
      }
    }
  } else {
    iret = decode_budget_opt(ipopt, nbud, vec, iret, P->bo);
    BudgetOpt& bo = P->bo;
    if (bo.nsamp > 0 && no > 0) {
      const int bt = nbud ? 1 : 4;
      const size_t tot = (size_t)bo.nsamp * bt * no;
      P->bn = dalloc<int>(tot);
      if (!nbud) P->bw = dalloc<R>(tot);
      if (vec) { P->bc = dalloc<R>(tot); P->bs = dalloc<R>(tot); }
      P->bytes += tot * (4 + (nbud ? 0 : sizeof(R)) + (vec ? 2 * sizeof(R) : 0));
      std::vector<R> wbr(bo.nsamp);
      for (int s = 0; s < bo.nsamp; s++) wbr[s] = (R)bo.wb[s];
      P->d_wb = dalloc<R>(bo.nsamp);
      int* d_ij = dalloc<int>(2 * bo.nsamp);
      IPB_CUDA(cudaMemcpyAsync(P->d_wb, wbr.data(), bo.nsamp * sizeof(R), cudaMemcpyHostToDevice, st));
      IPB_CUDA(cudaMemcpyAsync(d_ij, bo.ib.data(), bo.nsamp * sizeof(int), cudaMemcpyHostToDevice, st));
      IPB_CUDA(cudaMemcpyAsync(d_ij + bo.nsamp, bo.jb.data(), bo.nsamp * sizeof(int), cudaMemcpyHostToDevice, st));
      const R rb2 = (bo.lsw == 2) ? IPB_RC(1.) / (R)(bo.nb1 + 1) : IPB_RC(1.) / (R)bo.nb2;
      const GridP<R>& pg = station ? gi : go;
      dim3 grid(nblk(no, 128), bo.nsamp);
      if (!nbud) {
        if (vec) k_plan_budget<R, false, true><<<grid, 128, 0, st>>>(gi, pg, (int)no, P->x, P->y, P->rlat, P->rlon, d_ij, d_ij + bo.nsamp, rb2, bo.nb2, P->bn, P->bw, P->bc, P->bs, d_cnt + 1);
        else k_plan_budget<R, false, false><<<grid, 128, 0, st>>>(gi, pg, (int)no, P->x, P->y, P->rlat, P->rlon, d_ij, d_ij + bo.nsamp, rb2, bo.nb2, P->bn, P->bw, nullptr, nullptr, d_cnt + 1);
      } else {
        if (vec) k_plan_budget<R, true, true><<<grid, 128, 0, st>>>(gi, pg, (int)no, P->x, P->y, P->rlat, P->rlon, d_ij, d_ij + bo.nsamp, rb2, bo.nb2, P->bn, nullptr, P->bc, P->bs, d_cnt + 1);
        else k_plan_budget<R, true, false><<<grid, 128, 0, st>>>(gi, pg, (int)no, P->x, P->y, P->rlat, P->rlon, d_ij, d_ij + bo.nsamp, rb2, bo.nb2, P->bn, nullptr, nullptr, nullptr, d_cnt + 1);
      }
      g_launches++;
      IPB_CUDA(cudaMemcpyAsync(h_cnt + 1, d_cnt + 1, sizeof(int), cudaMemcpyDeviceToHost, st));
      IPB_CUDA(cudaStreamSynchronize(st));
      dfree(d_ij);
      bool has_center = false;
      for (int s = 0; s < bo.nsamp; s++) if (bo.lb[s] == 0) has_center = true;
      if (iret == 0 && has_center && h_cnt[1] == 0) iret = 2;  // budget_interp_mod.F90:215
    }
  }

  // pole lists for lat-lon outputs
  if (go.type == P_EQUID && no > 0) {
    std::vector<R> hl(no);
    IPB_CUDA(cudaMemcpyAsync(hl.data(), P->rlat, no * sizeof(R), cudaMemcpyDeviceToHost, st));
    IPB_CUDA(cudaStreamSynchronize(st));
    const R np_ = IPB_RC(89.9995), sp_ = -np_;
    std::vector<int> pn, ps;
    for (i64 n = 0; n < no; n++) { if (hl[n] >= np_) pn.push_back((int)n); else if (hl[n] <= sp_) ps.push_back((int)n); }
    P->npn = (int)pn.size(); P->nps = (int)ps.size();
    auto up = [&](std::vector<int>& v, int*& d, R*& c, R*& s) {
      if (v.empty()) return;
      d = dalloc<int>(v.size());
      IPB_CUDA(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, st));
      if (vec) {
        c = dalloc<R>(v.size()); s = dalloc<R>(v.size());
        k_pole_trig<R><<<nblk(v.size(), 128), 128, 0, st>>>((int)v.size(), d, P->rlon, c, s);
        g_launches++;
      }
      IPB_CUDA(cudaStreamSynchronize(st));
    };
    up(pn, P->pole_n, P->pclon_n, P->pslon_n);
    up(ps, P->pole_s, P->pclon_s, P->pslon_s);
  }
  dfree(d_cnt);
  {  // source window actually referenced by the plan
    const int* taps = budget ? P->bn : P->nxy;
    const size_t ntap = budget ? (size_t)P->bo.nsamp * (nbud ? 1 : 4) * no : (size_t)P->nt * no;
    P->src_lo = 1; P->src_hi = 0;
    if (taps && ntap > 0) {
      int* d_lohi = dalloc<int>(2);
      const int init[2] = {0x7fffffff, 0};
      IPB_CUDA(cudaMemcpyAsync(d_lohi, init, sizeof(init), cudaMemcpyHostToDevice, st));
      k_idx_range<<<std::min(nblk((long long)ntap, 256), 148 * 8), 256, 0, st>>>(taps, ntap, d_lohi);
      g_launches++;
      int lohi[2];
      IPB_CUDA(cudaMemcpyAsync(lohi, d_lohi, sizeof(lohi), cudaMemcpyDeviceToHost, st));
      IPB_CUDA(cudaStreamSynchronize(st));
      dfree(d_lohi);
      if (lohi[1] > 0) { P->src_lo = lohi[0]; P->src_hi = lohi[1]; }
    }
  }
  P->iret = iret;
  cudaEventRecord(e1.e, st);
  cudaEventSynchronize(e1.e);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0.e, e1.e);
  P->build_ms = ms;
  return holder.release();
}

}  // namespace ipb
