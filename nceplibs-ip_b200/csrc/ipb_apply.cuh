// ipb200 — the km-field apply kernels.
//
// Output-point-major: a thread owns one (or, in the fast bilinear path, four consecutive) output
// points, keeps that point's taps and weights in registers and walks the field batch, so the plan
// is read from HBM once per launch instead of once per field as in the reference's flattened
// NK loop (src/bilinear_interp_mod.F90:202-261).  A warp covers 32 consecutive output points, so
// go/lo stores are fully coalesced and the source gathers of neighbouring points hit the same
// L1/L2 lines.  Accumulation order inside a point is exactly the reference's (tap (1,1),(2,1),
// (1,2),(2,2); budget sub-samples NB=1..NB3), so results are bit-identical to the oracle.
//
// Field layout (column-major Fortran arrays seen from C): field k of the input starts at
// gi + k*mi, of the output at go + k*mo; li/lo are one byte per point.
#pragma once
#include "ipb_plan.cuh"

namespace ipb {

template <class R> struct FieldArgs {
  int no, mi, mo, kc;            // points, strides, fields in this launch
  const R* gi; const R* vi;      // [kc][mi]  (vi: second component for vectors)
  const u8* li;                  // [kc][mi] or nullptr when every ibi==0
  const int* ibi;                // [kc] device
  R* go; R* vo; u8* lo;          // [kc][mo]
  int* flag;                     // [kc] set to 1 when any lo is false (-> ibo)
  R pmp;                         // mask threshold (fraction, or fraction*nb4 for the budget family)
  int mcon, mspiral;
};

// ibo bookkeeping: flag[k] != 0 means some lo(:,k) is false.  Hundreds of thousands of threads may want to
// set the same word; writing only when the (L1-cached, possibly stale) value is still zero keeps that down to
// about one store per SM and field instead of one per warp, which would serialise in a single L2 slice.
__device__ __forceinline__ void mark_bad(int* flag, int k) { if (flag[k] == 0) flag[k] = 1; }

// Spiral search around (x,y): e.g. bilinear_interp_mod.F90:224-253.  Returns the 1-based
// position of the first acceptable point, 0 if none.  first = 1 (bilinear) or 2 (others).
template <class R>
__device__ int spiral_find(const GridP<R>& gin, R x, R y, int first, int mspiral, const u8* l, bool any_ok) {
  const int i1 = f_nint<R>(x), j1 = f_nint<R>(y);
  const int ixs = (int)f_sign1<R>(x - (R)i1), jxs = (int)f_sign1<R>(y - (R)j1);
  const int last = mspiral * mspiral;
  for (int mx = first; mx <= last; mx++) {
    const int kxs = (int)DM<R>::sqrt((R)(4 * mx) - IPB_DL(2.5));
    const int kxt = mx - (kxs * kxs / 4 + 1);
    int ix, jx;
    switch (kxs % 4) {
      case 1: ix = i1 - ixs * (kxs / 4 - kxt); jx = j1 - jxs * kxs / 4; break;
      case 2: ix = i1 + ixs * (1 + kxs / 4); jx = j1 - jxs * (kxs / 4 - kxt); break;
      case 3: ix = i1 + ixs * (1 + kxs / 4 - kxt); jx = j1 + jxs * (1 + kxs / 4); break;
      default: ix = i1 - ixs * kxs / 4; jx = j1 + jxs * (kxs / 4 - kxt); break;
    }
    const int p = field_pos(gin, ix, jx);
    if (p > 0 && (any_ok || (l && l[p - 1]))) return p;
  }
  return 0;
}

// Output point of a thread in a block that covers a 32 x ROWS patch of the output grid (rows of `rowlen`
// consecutive points, one warp per row): the warps of a block gather from the same few source rows, so the
// sectors one warp pulls into L1 are hits for the others.  Returns -1 outside the grid.
template <int ROWS> __device__ __forceinline__ int patch_point(int rowlen, int no) {
  const int col = blockIdx.x * 32 + (threadIdx.x & 31);
  const long long n = (long long)(blockIdx.y * ROWS + (threadIdx.x >> 5)) * rowlen + col;
  return (col < rowlen && n < no) ? (int)n : -1;
}
inline dim3 patch_grid(int rowlen, int no, int rows) { return dim3(nblk(rowlen, 32), nblk(nblk(no, rowlen), rows)); }
// Row length of a grid's own point enumeration (gdswzd iopt=0); 256 (plain linear blocks) for station points.
template <class R> inline int patch_rowlen(const GridP<R>& g, int no) {
  if (g.type == P_STATION || g.im <= 0 || g.jm <= 0) return 256;
  const int r = g.nscan == 0 ? g.im : g.jm;
  return (r > 0 && r <= no) ? r : 256;
}

// ---- bilinear (NT=2) / bicubic (NT=4), scalar and vector -------------------------------------------
// KPB = fields walked by one thread (grid.y covers ceil(kc/KPB)).
template <class R, int NT, bool VEC>
__global__ void __launch_bounds__(128)
k_apply_stencil(FieldArgs<R> a, GridP<R> gin, int kpb, const int* __restrict__ nxy, const R* __restrict__ wxy,
                const R* __restrict__ cxy, const R* __restrict__ sxy, const u8* __restrict__ nc,
                const R* __restrict__ xs, const R* __restrict__ ys, const R* __restrict__ crot, const R* __restrict__ srot, int rowlen) {
  const int n = patch_point<4>(rowlen, a.no);
  if (n < 0) return;
  constexpr int T2 = NT * NT;
  int idx[T2];
  R w[T2], c[VEC ? T2 : 1], s[VEC ? T2 : 1];
#pragma unroll
  for (int t = 0; t < T2; t++) {
    idx[t] = nxy[(size_t)t * a.no + n];
    w[t] = wxy[(size_t)t * a.no + n];
    if (VEC) { c[t] = cxy[(size_t)t * a.no + n]; s[t] = sxy[(size_t)t * a.no + n]; }
  }
  const int cls = (NT == 4) ? nc[n] : 1;
  const R cr = VEC ? crot[n] : (R)1, sr = VEC ? srot[n] : (R)0;
  const int k0 = blockIdx.z * kpb, k1 = min(k0 + kpb, a.kc);
  // complete stencil: every tap present (bicubic: the true 4x4 case).  Without a bitmap the tap set and
  // the weight sum are then the same for every field, and the field loop is straight-line code.
  bool complete = cls == 1;
  R Wall = 0;
#pragma unroll
  for (int t = 0; t < T2; t++) { complete = complete && idx[t] > 0; Wall = Wall + w[t]; }
  complete = complete && Wall >= a.pmp;
  for (int k = k0; k < k1; k++) {
    const R* __restrict__ g = a.gi + (size_t)k * a.mi;
    const R* __restrict__ gv = VEC ? a.vi + (size_t)k * a.mi : nullptr;
    // a.li is null when no field of the call carries a bitmap: the ibi loads then stay off the critical path
    const bool masked = a.li != nullptr && a.ibi[k] != 0;
    const u8* __restrict__ l = a.li ? a.li + (size_t)k * a.mi : nullptr;
    const size_t oa = (size_t)k * a.mo + n;
    if (complete && !masked) {   // complete also means the weight sum passes the threshold
      // KU fields per iteration with all their gathers issued first: this loop is bound by gather latency
      // (ncu, vector bilinear S -> P: 43 % of the warps' time waiting on the gathers with two pairs in flight)
      constexpr int KU = (NT == 2) ? 4 : 1;
      int ku = 1;
      if (KU > 1) while (ku < KU && k + ku < k1 && (a.li == nullptr || a.ibi[k + ku] == 0)) ku++;
      R tv[KU][T2], tu[KU][VEC ? T2 : 1];
#pragma unroll
      for (int q = 0; q < KU; q++) {
        if (q < ku) {
          const R* __restrict__ gq = g + (size_t)q * a.mi;
          const R* __restrict__ gvq = VEC ? gv + (size_t)q * a.mi : nullptr;
#pragma unroll
          for (int t = 0; t < T2; t++) { tv[q][t] = __ldg(gq + (idx[t] - 1)); if (VEC) tu[q][t] = __ldg(gvq + (idx[t] - 1)); }
        }
      }
#pragma unroll
      for (int q = 0; q < KU; q++) {
        if (q >= ku) break;
        R G = 0, V = 0;
        R gmin = DM<R>::huge(), gmax = -DM<R>::huge(), vmin = DM<R>::huge(), vmax = -DM<R>::huge();
        const bool clampit = NT == 4 && a.mcon > 0;
#pragma unroll
        for (int t = 0; t < T2; t++) {
          if (!VEC) {
            G = G + w[t] * tv[q][t];
          } else {
            const R u = tv[q][t], v = tu[q][t];
            const R ur = c[t] * u - s[t] * v;
            const R vr = s[t] * u + c[t] * v;
            G = G + w[t] * ur;
            V = V + w[t] * vr;
            if (clampit) {
              gmin = DM<R>::vmin(gmin, ur); gmax = DM<R>::vmax(gmax, ur);
              vmin = DM<R>::vmin(vmin, vr); vmax = DM<R>::vmax(vmax, vr);
            }
          }
        }
        if (!VEC && clampit) {   // separate pass: the common unconstrained case pays nothing for it
#pragma unroll
          for (int t = 0; t < T2; t++) { gmin = DM<R>::vmin(gmin, tv[q][t]); gmax = DM<R>::vmax(gmax, tv[q][t]); }
        }
        R og, ov = 0;
        if (!VEC) {
          og = G / Wall;
          if (clampit) og = DM<R>::vmin(DM<R>::vmax(og, gmin), gmax);
        } else {
          const R ur = cr * G - sr * V;
          const R vr = sr * G + cr * V;
          og = ur / Wall; ov = vr / Wall;
          if (clampit) { og = DM<R>::vmin(DM<R>::vmax(og, gmin), gmax); ov = DM<R>::vmin(DM<R>::vmax(ov, vmin), vmax); }
        }
        const size_t ob = oa + (size_t)q * a.mo;
        __stcs(a.go + ob, og);
        if (VEC) __stcs(a.vo + ob, ov);
        __stcs(a.lo + ob, (u8)1);
      }
      k += ku - 1;
      continue;
    }
    R G = 0, V = 0, W = 0;
    R gmin = DM<R>::huge(), gmax = -DM<R>::huge(), vmin = DM<R>::huge(), vmax = -DM<R>::huge();
    bool good = false;
    if (cls > 0) {
#pragma unroll
      for (int t = 0; t < T2; t++) {
        if (NT == 4 && cls == 2) {  // bilinear fallback visits the inner 2x2 only (bicubic_interp_mod.F90:245-246)
          const int ta = t & 3, tb = t >> 2;
          if (ta < 1 || ta > 2 || tb < 1 || tb > 2) continue;
        }
        const int p = idx[t];
        if (p > 0 && (!masked || l[p - 1])) {
          if (!VEC) {
            const R v = g[p - 1];
            G = G + w[t] * v;
            W = W + w[t];
            if (NT == 4 && a.mcon > 0) { gmin = DM<R>::vmin(gmin, v); gmax = DM<R>::vmax(gmax, v); }
          } else {
            const R u = g[p - 1], v = gv[p - 1];
            const R ur = c[t] * u - s[t] * v;
            const R vr = s[t] * u + c[t] * v;
            G = G + w[t] * ur;
            V = V + w[t] * vr;
            W = W + w[t];
            if (NT == 4 && a.mcon > 0) {
              gmin = DM<R>::vmin(gmin, ur); gmax = DM<R>::vmax(gmax, ur);
              vmin = DM<R>::vmin(vmin, vr); vmax = DM<R>::vmax(vmax, vr);
            }
          }
        }
      }
      good = W >= a.pmp;
    }
    R og = 0, ov = 0;
    if (good) {
      if (!VEC) {
        og = G / W;
        if (NT == 4 && a.mcon > 0) og = DM<R>::vmin(DM<R>::vmax(og, gmin), gmax);
      } else {
        const R ur = cr * G - sr * V;
        const R vr = sr * G + cr * V;
        og = ur / W; ov = vr / W;
        if (NT == 4 && a.mcon > 0) { og = DM<R>::vmin(DM<R>::vmax(og, gmin), gmax); ov = DM<R>::vmin(DM<R>::vmax(ov, vmin), vmax); }
      }
    } else if (NT == 2 && !VEC && a.mspiral > 0) {
      const R x = xs[n], y = ys[n];
      if (valid_xy(x, y)) {
        const int p = spiral_find<R>(gin, x, y, 1, a.mspiral, l, !masked);
        if (p > 0) { og = g[p - 1]; good = true; }
      }
    }
    a.go[oa] = og;
    if (VEC) a.vo[oa] = ov;
    a.lo[oa] = good ? 1 : 0;
    if (!good) mark_bad(a.flag, k);
  }
}

// ---- neighbor ------------------------------------------------------------------------------------
// neighbor_interp_mod.F90:213-263 (:484-546 vector)
template <class R, bool VEC>
__global__ void __launch_bounds__(256)
k_apply_neighbor(FieldArgs<R> a, GridP<R> gin, int kpb, const int* __restrict__ nxy, const R* __restrict__ cxy,
                 const R* __restrict__ sxy, const R* __restrict__ xs, const R* __restrict__ ys,
                 const R* __restrict__ rlat, const R* __restrict__ rlon, const R* __restrict__ crot, const R* __restrict__ srot) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.no) return;
  const int p0 = nxy[n];
  const R c0 = VEC ? cxy[n] : (R)1, s0 = VEC ? sxy[n] : (R)0;
  const R cr = VEC ? crot[n] : (R)1, sr = VEC ? srot[n] : (R)0;
  const int k0 = blockIdx.y * kpb, k1 = min(k0 + kpb, a.kc);
  for (int k = k0; k < k1; k++) {
    const R* __restrict__ g = a.gi + (size_t)k * a.mi;
    const R* __restrict__ gv = VEC ? a.vi + (size_t)k * a.mi : nullptr;
    const bool masked = a.ibi[k] != 0;
    const u8* __restrict__ l = a.li ? a.li + (size_t)k * a.mi : nullptr;
    const size_t oa = (size_t)k * a.mo + n;
    int p = 0;
    R c = c0, s = s0;
    if (p0 > 0) {
      if (!masked || l[p0 - 1]) p = p0;
      else if (a.mspiral > 1) {
        p = spiral_find<R>(gin, xs[n], ys[n], 2, a.mspiral, l, false);
        if (VEC && p > 0) tap_rotation(gin, p, rlat[n], rlon[n], c, s);
      }
    }
    R og = 0, ov = 0;
    if (p > 0) {
      if (!VEC) og = g[p - 1];
      else {
        const R u = g[p - 1], v = gv[p - 1];
        const R ur = c * u - s * v;
        const R vr = s * u + c * v;
        og = cr * ur - sr * vr;
        ov = sr * ur + cr * vr;
      }
    }
    a.go[oa] = og;
    if (VEC) a.vo[oa] = ov;
    a.lo[oa] = p > 0 ? 1 : 0;
    if (p <= 0) mark_bad(a.flag, k);
  }
}

// ---- budget / neighbor-budget --------------------------------------------------------------------
// budget_interp_mod.F90:252-344 (:630-705 vector); neighbor_budget_interp_mod.F90:221-246 (:507-541).
// A thread owns one output point and KF fields; sub-samples are the outer loop so each sample's
// taps are read once per KF fields.  Sample order (and therefore rounding) is the reference's.
template <class R, bool NBUD, bool VEC, int KF>
__global__ void __launch_bounds__(128)
k_apply_budget(FieldArgs<R> a, GridP<R> gin, int nsamp, const int* __restrict__ bn, const R* __restrict__ bw,
               const R* __restrict__ bc, const R* __restrict__ bs, const R* __restrict__ wbs,
               const R* __restrict__ rlat, const R* __restrict__ rlon, const R* __restrict__ crot, const R* __restrict__ srot) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= a.no) return;
  constexpr int BT = NBUD ? 1 : 4;
  const int k0 = blockIdx.y * KF;
  R acc[KF], acv[VEC ? KF : 1], wo[KF];
  bool masked[KF];
#pragma unroll
  for (int f = 0; f < KF; f++) {
    acc[f] = 0; wo[f] = 0; if (VEC) acv[f] = 0;
    masked[f] = (k0 + f < a.kc) ? (a.ibi[k0 + f] != 0) : false;
  }
  for (int sidx = 0; sidx < nsamp; sidx++) {
    const size_t base = ((size_t)sidx * BT) * a.no + n;
    int idx[BT];
    R w[BT], c[VEC ? BT : 1], s[VEC ? BT : 1];
#pragma unroll
    for (int t = 0; t < BT; t++) {
      idx[t] = bn[base + (size_t)t * a.no];
      w[t] = NBUD ? (R)1 : bw[base + (size_t)t * a.no];
      if (VEC) { c[t] = bc[base + (size_t)t * a.no]; s[t] = bs[base + (size_t)t * a.no]; }
    }
    if (idx[0] <= 0) continue;
    const R wb = wbs[sidx];
#pragma unroll
    for (int f = 0; f < KF; f++) {
      const int k = k0 + f;
      if (k >= a.kc) break;
      const R* __restrict__ g = a.gi + (size_t)k * a.mi;
      const R* __restrict__ gv = VEC ? a.vi + (size_t)k * a.mi : nullptr;
      const u8* __restrict__ l = a.li ? a.li + (size_t)k * a.mi : nullptr;
      if (NBUD) {
        const int p = idx[0] - 1;
        if (!masked[f] || l[p]) {
          if (!VEC) acc[f] = acc[f] + wb * g[p];
          else {
            const R u = g[p], v = gv[p];
            const R u11 = c[0] * u - s[0] * v, v11 = s[0] * u + c[0] * v;
            acc[f] = acc[f] + wb * u11; acv[f] = acv[f] + wb * v11;
          }
          wo[f] = wo[f] + wb;
        }
      } else if (!masked[f]) {
        if (!VEC) {
          const R gb = w[0] * g[idx[0] - 1] + w[1] * g[idx[1] - 1] + w[2] * g[idx[2] - 1] + w[3] * g[idx[3] - 1];
          acc[f] = acc[f] + wb * gb;
        } else {
          R ur[4], vr[4];
#pragma unroll
          for (int t = 0; t < 4; t++) {
            const R u = g[idx[t] - 1], v = gv[idx[t] - 1];
            ur[t] = c[t] * u - s[t] * v;
            vr[t] = s[t] * u + c[t] * v;
          }
          const R ub = w[0] * ur[0] + w[1] * ur[1] + w[2] * ur[2] + w[3] * ur[3];
          const R vb = w[0] * vr[0] + w[1] * vr[1] + w[2] * vr[2] + w[3] * vr[3];
          acc[f] = acc[f] + wb * ub; acv[f] = acv[f] + wb * vb;
        }
        wo[f] = wo[f] + wb;
      } else {
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const int p = idx[t] - 1;
          if (l[p]) {
            if (!VEC) acc[f] = acc[f] + wb * w[t] * g[p];
            else {
              const R u = g[p], v = gv[p];
              const R ur = c[t] * u - s[t] * v, vr = s[t] * u + c[t] * v;
              acc[f] = acc[f] + wb * w[t] * ur; acv[f] = acv[f] + wb * w[t] * vr;
            }
            wo[f] = wo[f] + wb * w[t];
          }
        }
      }
    }
  }
  const R cr = VEC ? crot[n] : (R)1, sr = VEC ? srot[n] : (R)0;
#pragma unroll
  for (int f = 0; f < KF; f++) {
    const int k = k0 + f;
    if (k >= a.kc) break;
    const size_t oa = (size_t)k * a.mo + n;
    bool good = wo[f] >= a.pmp;
    R og = 0, ov = 0;
    if (good) {
      og = acc[f] / wo[f];
      if (VEC) {
        ov = acv[f] / wo[f];
        const R ur = cr * og - sr * ov, vr = sr * og + cr * ov;
        og = ur; ov = vr;
      }
    } else if (!NBUD && !VEC && a.mspiral > 1) {
      // fresh inverse of the output point on the input grid (budget_interp_mod.F90:295-299)
      const R fill = IPB_DL(-9999.);
      R xx, yy;
      PtExt<R> e;
      if (proj_inv<R, EXT_NONE>(gin, fill, rlon[n], rlat[n], xx, yy, e)) {
        const R* __restrict__ g = a.gi + (size_t)k * a.mi;
        const u8* __restrict__ l = a.li ? a.li + (size_t)k * a.mi : nullptr;
        const int p = spiral_find<R>(gin, xx, yy, 2, a.mspiral, l, !masked[f]);
        if (p > 0) { og = g[p - 1]; good = true; }
      }
    }
    a.go[oa] = og;
    if (VEC) a.vo[oa] = ov;
    a.lo[oa] = good ? 1 : 0;
    if (!good) mark_bad(a.flag, k);
  }
}

// if (on) { a += x; b += y; } as two predicated adds (the compiler's own choice is an unconditional add plus
// register selects, four more instructions per tap in binary64)
template <class R> __device__ __forceinline__ void add2_if(unsigned on, R& a, R x, R& b, R y);
template <> __device__ __forceinline__ void add2_if<double>(unsigned on, double& a, double x, double& b, double y) {
  asm("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p add.rn.f64 %0, %0, %3;\n @p add.rn.f64 %1, %1, %4;\n}\n" : "+d"(a), "+d"(b) : "r"(on), "d"(x), "d"(y));
}
template <> __device__ __forceinline__ void add2_if<float>(unsigned on, float& a, float x, float& b, float y) {
  asm("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p add.rn.f32 %0, %0, %3;\n @p add.rn.f32 %1, %1, %4;\n}\n" : "+f"(a), "+f"(b) : "r"(on), "f"(x), "f"(y));
}

// ---- budget, scalar, without spiral search: the batched kernel ------------------------------------------
// budget_interp_mod.F90:252-344.  The sub-sample plan is large (25 samples x 4 taps x (index + weight) =
// 1200 bytes per output point in the _d kind) and does not fit L2, so what bounds this method is how often
// the plan is streamed.  A block owns 32 consecutive output points (one per lane) and 8 x KF fields (warp w
// owns fields w*KF .. w*KF+KF-1): every chunk of samples is copied once into shared memory — coalesced, the
// plan is tap-major — and used by all eight warps, so the plan crosses HBM once per 64 fields instead of once
// per 8.  Accumulation order and association are the reference's: per sample either the 4-tap sum times the
// sample weight (field without bitmap) or tap by tap with the product wb*w formed first (field with bitmap).
template <class R, int KF>
__global__ void __launch_bounds__(256, 2)
k_budget_scalar(FieldArgs<R> a, int nsamp, const int* __restrict__ bn, const R* __restrict__ bw, const R* __restrict__ wbs) {
  constexpr int SCH = 16;                              // samples per shared-memory chunk
  __shared__ int s_idx[SCH * 4 * 32];
  __shared__ R s_w[SCH * 4 * 32];
  __shared__ R s_wb[SCH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + lane;
  const bool live = n < a.no;
  const int k0 = (blockIdx.y * 8 + warp) * KF;         // first field of this warp
  R acc[KF], wo[KF];
  unsigned maskbits = 0;
#pragma unroll
  for (int f = 0; f < KF; f++) {
    acc[f] = 0; wo[f] = 0;
    if (k0 + f < a.kc && a.ibi[k0 + f] != 0) maskbits |= 1u << f;
  }
  R wo_plain = 0;                                      // weight sum of the fields without a bitmap (the same for all of them)
  // element offset of every field of this warp, formed once (the launcher checks that kc * mi fits 32 bits): the kernel is
  // bound by instruction issue, and a 64-bit k * mi per gather was a good part of it
  unsigned koff[KF];
#pragma unroll
  for (int f = 0; f < KF; f++) koff[f] = (unsigned)min(k0 + f, a.kc - 1) * (unsigned)a.mi;
  for (int s0 = 0; s0 < nsamp; s0 += SCH) {
    const int ns = min(SCH, nsamp - s0);
    __syncthreads();
    for (int r = warp; r < ns * 4; r += 8) {           // row r = (sample, tap): 32 consecutive points
      const size_t src = (size_t)(s0 * 4 + r) * a.no + n;
      s_idx[r * 32 + lane] = live ? bn[src] : 0;
      s_w[r * 32 + lane] = live ? bw[src] : (R)0;
    }
    if (threadIdx.x < ns) s_wb[threadIdx.x] = wbs[s0 + threadIdx.x];
    __syncthreads();
    if (k0 >= a.kc) continue;                          // a warp without fields only helps with the copies
    for (int s = 0; s < ns; s++) {
      const int i0 = s_idx[(s * 4 + 0) * 32 + lane];
      if (i0 <= 0) continue;                           // sample outside the input grid for this point
      const int i1 = s_idx[(s * 4 + 1) * 32 + lane], i2 = s_idx[(s * 4 + 2) * 32 + lane], i3 = s_idx[(s * 4 + 3) * 32 + lane];
      const R w0 = s_w[(s * 4 + 0) * 32 + lane], w1 = s_w[(s * 4 + 1) * 32 + lane];
      const R w2 = s_w[(s * 4 + 2) * 32 + lane], w3 = s_w[(s * 4 + 3) * 32 + lane];
      const R wb = s_wb[s];
      const R b0 = wb * w0, b1 = wb * w1, b2 = wb * w2, b3 = wb * w3;   // WB*W11 ... formed first, as the reference's left-to-right product
      wo_plain = wo_plain + wb;
#pragma unroll
      const unsigned j0 = (unsigned)(i0 - 1), j1 = (unsigned)(i1 - 1), j2 = (unsigned)(i2 - 1), j3 = (unsigned)(i3 - 1);
#pragma unroll
      for (int f = 0; f < KF; f++) {
        const R* __restrict__ g = a.gi;
        const unsigned o = koff[f];
        if (!((maskbits >> f) & 1u)) {
          const R gb = w0 * __ldg(g + (o + j0)) + w1 * __ldg(g + (o + j1)) + w2 * __ldg(g + (o + j2)) + w3 * __ldg(g + (o + j3));
          acc[f] = acc[f] + wb * gb;
        } else {
          // straight-line: every tap is loaded (the positions are valid), the bitmap only predicates the adds
          const u8* __restrict__ l = a.li;
          const u8 l0 = __ldg(l + (o + j0)), l1 = __ldg(l + (o + j1)), l2 = __ldg(l + (o + j2)), l3 = __ldg(l + (o + j3));
          const R g0 = __ldg(g + (o + j0)), g1 = __ldg(g + (o + j1)), g2 = __ldg(g + (o + j2)), g3 = __ldg(g + (o + j3));
          R ac = acc[f], wc = wo[f];
          add2_if<R>(l0, ac, b0 * g0, wc, b0);
          add2_if<R>(l1, ac, b1 * g1, wc, b1);
          add2_if<R>(l2, ac, b2 * g2, wc, b2);
          add2_if<R>(l3, ac, b3 * g3, wc, b3);
          acc[f] = ac; wo[f] = wc;
        }
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int f = 0; f < KF; f++) {
    const int k = k0 + f;
    if (k >= a.kc) break;
    const R w = ((maskbits >> f) & 1u) ? wo[f] : wo_plain;
    const bool good = w >= a.pmp;
    const size_t oa = (size_t)k * a.mo + n;
    __stcs(a.go + oa, good ? acc[f] / w : (R)0);
    __stcs(a.lo + oa, (u8)(good ? 1 : 0));
    if (!good) mark_bad(a.flag, k);
  }
}

// ---- pole fix for lat-lon outputs: one warp per field ------------------------------------------------
// polfix_mod.F90:29-105 (scalar), :124-220 (vector).  Sums run in ascending N order (the reference's
// serial order; its OpenMP reduction order is thread-count dependent).  All lanes carry the same
// running sum: lane q loads point base+q, the partial sums are replayed through shuffles in order.
template <class R, bool VEC>
__global__ void k_polfix(int kc, int mo, const int* __restrict__ ibo, int np, const int* __restrict__ pidx,
                         const R* __restrict__ pcl, const R* __restrict__ psl, int south, R* go, R* vo, u8* lo) {
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= kc) return;
  const bool ib = ibo[k] != 0;
  R* g = go + (size_t)k * mo;
  R* v = VEC ? vo + (size_t)k * mo : nullptr;
  u8* l = lo + (size_t)k * mo;
  R wsum = 0, tsum = 0, gs = 0, vs = 0;
  for (int base = 0; base < np; base += 32) {
    const int q = base + lane;
    bool in = q < np, use = false;
    R a = 0, b = 0;
    if (in) {
      const int p = pidx[q];
      use = !ib || l[p];
      if (use) {
        if (!VEC) a = g[p];
        else {
          const R cl = pcl[q], sl = psl[q], u = g[p], w = v[p];
          if (!south) { a = cl * u - sl * w; b = sl * u + cl * w; }
          else { a = cl * u + sl * w; b = -sl * u + cl * w; }
        }
      }
    }
    const int cnt = min(32, np - base);
    for (int j = 0; j < cnt; j++) {
      const bool uj = __shfl_sync(0xffffffffu, (int)use, j);
      const R aj = __shfl_sync(0xffffffffu, a, j);
      const R bj = VEC ? __shfl_sync(0xffffffffu, b, j) : (R)0;
      wsum = wsum + 1;
      if (uj) { gs = gs + aj; if (VEC) vs = vs + bj; tsum = tsum + 1; }
    }
  }
  if (!(wsum > 1)) return;
  const bool keep = tsum >= wsum / 2;
  R gm = 0, vm = 0;
  if (keep) {
    // the vector south pole divides by the point count, not the valid count (polfix_mod.F90:202-203)
    const R den = (VEC && south) ? wsum : tsum;
    gm = gs / den;
    if (VEC) vm = vs / den;
  }
  for (int q = lane; q < np; q += 32) {
    const int p = pidx[q];
    if (ib) l[p] = keep ? 1 : 0;
    if (!VEC) g[p] = gm;
    else {
      const R cl = pcl[q], sl = psl[q];
      if (!south) { g[p] = cl * gm + sl * vm; v[p] = -sl * gm + cl * vm; }
      else { g[p] = cl * gm - sl * vm; v[p] = sl * gm + cl * vm; }
    }
  }
}

}  // namespace ipb
