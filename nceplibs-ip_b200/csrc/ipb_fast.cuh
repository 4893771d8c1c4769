// ipb200 — the streaming bilinear apply kernel (ipolates ip=0, the BASELINE headline path).
//
// What the reference does per (output point, field) — src/bilinear_interp_mod.F90:202-261 — is a
// 4-tap masked weighted sum followed by a divide and a 5-byte store (go + lo).  At the headline
// configuration the output grid oversamples the source ~66x, so the loop is a *write-bound
// stream*: 5.09 algorithmic bytes per point-field, of which 5 are stores.
// This kernel is built around that:
//
//  * a thread owns FOUR consecutive output points and walks a slice of the field batch with the
//    16 tap offsets / weights, the per-point weight sums, their reciprocals and the packed lo
//    word held in registers: the plan is read once per slice, not once per field;
//  * go is stored as one 16-byte (fp32) / 32-byte (fp64) vector and lo as one 32-bit word per
//    thread and field, with streaming (evict-first) hints so the 10 GB output stream does not
//    push the plan and the touched source window out of L2;
//  * field k of go starts at element k*mo, which is only vector-aligned for some k.  Fields are
//    therefore split into four alignment classes (k mod 4); a block works on one class, and the
//    thread->point mapping is shifted per class so every vector store is aligned;
//  * when no bitmap is in play the weight sum W and lo do not depend on the field, and G/W is
//    formed as a correctly rounded quotient from the stored correctly rounded reciprocal
//    (two fused residual corrections, Markstein) instead of a full IEEE division sequence;
//  * blockIdx.x runs over field slices, blockIdx.y over point blocks, so blocks that are
//    resident together share the same plan segment (L2 hits) while streaming different fields.
//
// Accumulation order and every rounding are the reference's: taps (1,1),(2,1),(1,2),(2,2), one
// multiply and one add each (the library is compiled with -fmad=false).
#pragma once
#include <cstdlib>
#include "ipb_apply.cuh"

namespace ipb {

template <class R> __device__ __forceinline__ R fma_r(R a, R b, R c);
template <> __device__ __forceinline__ float fma_r<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <> __device__ __forceinline__ double fma_r<double>(double a, double b, double c) { return __fma_rn(a, b, c); }

// g / w, correctly rounded, given rw = RN(1/w).  q0 = RN(g*rw) is within 2 ulp; one fused residual
// step makes it faithful and the second one correctly rounded (Markstein's theorem needs rw
// correctly rounded and a faithful q).  Outside the exponent window where the residuals are exact
// the plain division is used.
template <class R> __device__ __forceinline__ R quot(R g, R w, R rw) {
  const R ag = g < 0 ? -g : g;
  const R lo = sizeof(R) == 4 ? (R)1e-25f : (R)1e-250, hi = sizeof(R) == 4 ? (R)1e25f : (R)1e250;
  if (!(ag < hi) || (ag < lo && ag != 0)) return g / w;
  R q = g * rw;
  R r = fma_r<R>(-q, w, g);
  q = fma_r<R>(r, rw, q);
  r = fma_r<R>(-q, w, g);
  return fma_r<R>(r, rw, q);
}

template <class R> __device__ __forceinline__ void store4(R* p, R a, R b, R c, R d);
template <> __device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
  __stcs(reinterpret_cast<float4*>(p), make_float4(a, b, c, d));
}
template <> __device__ __forceinline__ void store4<double>(double* p, double a, double b, double c, double d) {
  __stcs(reinterpret_cast<double2*>(p), make_double2(a, b));
  __stcs(reinterpret_cast<double2*>(p) + 1, make_double2(c, d));
}

// One output point, one field, any mask state: the reference loop body without the spiral search.
template <class R>
__device__ __forceinline__ bool bilinear_point(const R* __restrict__ g, const u8* __restrict__ l, bool masked, const int (&idx)[4],
                                               const R (&w)[4], R pmp, R& out) {
  R G = 0, W = 0;
#pragma unroll
  for (int t = 0; t < 4; t++) {
    const int p = idx[t];
    if (p > 0 && (!masked || l[p - 1])) {
      G = G + w[t] * __ldg(g + (p - 1));
      W = W + w[t];
    }
  }
  const bool good = W >= pmp;
  out = good ? G / W : (R)0;
  return good;
}

// KPT = fields walked by one thread.  blockIdx.x = 4*slice + class, blockIdx.y = point block.
template <class R, int KPT>
__global__ void __launch_bounds__(128)
k_bilinear_stream(FieldArgs<R> a, const int* __restrict__ nxy, const R* __restrict__ wxy, int base_mis) {
  const int cls = blockIdx.x & 3, slice = blockIdx.x >> 2;
  // first output point whose element offset is a multiple of 4 for fields of this class
  const int first = (4 - (int)(((long long)base_mis + (long long)cls * a.mo) & 3)) & 3;
  const int n0 = (first ? first - 4 : 0) + 4 * (int)(blockIdx.y * blockDim.x + threadIdx.x);
  if (n0 >= a.no) return;
  int idx[4][4];
  R w[4][4], W[4], rW[4];
  bool full = true;        // all four points inside, every tap present, every point passes the mask threshold
  unsigned goodbits = 0;   // bit p: point p passes when no bitmap is involved
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int n = n0 + p;
    const bool in = n >= 0 && n < a.no;
    R ws = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) {
      idx[p][t] = in ? nxy[(size_t)t * a.no + n] : 0;
      w[p][t] = in ? wxy[(size_t)t * a.no + n] : (R)0;
      if (idx[p][t] > 0) ws = ws + w[p][t]; else full = false;
    }
    W[p] = ws;
    rW[p] = (R)1 / ws;
    const bool good = in && ws >= a.pmp;
    if (good) goodbits |= 1u << p;
    full = full && good;
  }
  const unsigned lo_word = (goodbits & 1u) | ((goodbits & 2u) << 7) | ((goodbits & 4u) << 14) | ((goodbits & 8u) << 21);
  const bool all_in = n0 >= 0 && n0 + 3 < a.no;
  const int kbeg = cls + 4 * slice * KPT;
#pragma unroll 1
  for (int j = 0; j < KPT; j++) {
    const int k = kbeg + 4 * j;
    if (k >= a.kc) break;
    const R* __restrict__ g = a.gi + (size_t)k * a.mi;
    const size_t ob = (size_t)k * a.mo;
    const bool masked = a.ibi[k] != 0;
    if (!masked && full) {
      R o[4];
#pragma unroll
      for (int p = 0; p < 4; p++) {
        R G = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) G = G + w[p][t] * __ldg(g + (idx[p][t] - 1));
        o[p] = quot<R>(G, W[p], rW[p]);
      }
      store4<R>(a.go + ob + n0, o[0], o[1], o[2], o[3]);
      __stcs(reinterpret_cast<unsigned*>(a.lo + ob + n0), lo_word);
    } else {
      const u8* __restrict__ l = a.li ? a.li + (size_t)k * a.mi : nullptr;
      R o[4];
      unsigned lw = 0;
      bool bad = false;
#pragma unroll
      for (int p = 0; p < 4; p++) {
        const bool good = bilinear_point<R>(g, l, masked, idx[p], w[p], a.pmp, o[p]);
        if (good) lw |= 1u << (8 * p);
        else if (n0 + p >= 0 && n0 + p < a.no) bad = true;
      }
      if (all_in) {
        store4<R>(a.go + ob + n0, o[0], o[1], o[2], o[3]);
        __stcs(reinterpret_cast<unsigned*>(a.lo + ob + n0), lw);
      } else {
#pragma unroll
        for (int p = 0; p < 4; p++)
          if (n0 + p >= 0 && n0 + p < a.no) { a.go[ob + n0 + p] = o[p]; a.lo[ob + n0 + p] = (u8)((lw >> (8 * p)) & 1u); }
      }
      if (bad) mark_bad(a.flag, k);
    }
  }
}

// The streaming kernel needs go and lo to share their 4-element phase (true for any pair of
// allocations aligned to 16 bytes) and does not implement the spiral search.
template <class R> bool fast_bilinear_ok(const Plan<R>& P, const FieldArgs<R>& a) {
  if (P.method != M_BILINEAR || P.vec || a.mspiral > 0 || P.no <= 0) return false;
  const uintptr_t g = reinterpret_cast<uintptr_t>(a.go), l = reinterpret_cast<uintptr_t>(a.lo);
  if (g % sizeof(R)) return false;
  return ((g / sizeof(R)) & 3) == (l & 3);
}

template <class R> void launch_fast_bilinear(const Plan<R>& P, const FieldArgs<R>& a, cudaStream_t st) {
  constexpr int KPT = 32;
  const int base_mis = (int)(reinterpret_cast<uintptr_t>(a.lo) & 3);
  const int nthreads = a.no / 4 + 2;
  const int per_class = (a.kc + 3) / 4;
  dim3 grid(4 * nblk(per_class, KPT), nblk(nthreads, 128));
  k_bilinear_stream<R, KPT><<<grid, 128, 0, st>>>(a, P.nxy, P.wxy, base_mis);
  g_launches++;
}

// ---- neighbor, scalar (ipolates ip=2, neighbor_interp_mod.F90:213-263) --------------------------------
// One output point per lane, so that the 32 gathers of a warp instruction fall on neighbouring source
// points (a few 32-byte sectors) whatever the resolution ratio of the two grids; go is stored as one
// coalesced 128-byte line per warp and field, lo as 32 bytes.  A thread walks a slice of the field
// batch with its source position in a register; eight fields are in flight at a time.
template <class R>
__global__ void __launch_bounds__(256)
k_neighbor_lane(FieldArgs<R> a, const int* __restrict__ nxy, int kpb, int rowlen) {
  const int n = patch_point<8>(rowlen, a.no);
  if (n < 0) return;
  const int p0 = nxy[n];
  const int k0 = blockIdx.z * kpb, k1 = min(k0 + kpb, a.kc);
  const R* __restrict__ g = a.gi + (size_t)k0 * a.mi + (p0 > 0 ? p0 - 1 : 0);
  R* go = a.go + (size_t)k0 * a.mo + n;
  u8* lo = a.lo + (size_t)k0 * a.mo + n;
  if (a.li == nullptr) {   // no bitmap anywhere in this call: lo depends on the point only
    const u8 ok = p0 > 0 ? 1 : 0;
    int k = k0;
    for (; k + 8 <= k1; k += 8) {
      R v[8];
#pragma unroll
      for (int f = 0; f < 8; f++) v[f] = ok ? __ldg(g + (size_t)f * a.mi) : (R)0;
#pragma unroll
      for (int f = 0; f < 8; f++) { __stcs(go + (size_t)f * a.mo, v[f]); __stcs(lo + (size_t)f * a.mo, ok); }
      g += 8 * (size_t)a.mi; go += 8 * (size_t)a.mo; lo += 8 * (size_t)a.mo;
    }
    for (; k < k1; k++) {
      __stcs(go, ok ? __ldg(g) : (R)0); __stcs(lo, ok);
      g += a.mi; go += a.mo; lo += a.mo;
    }
    if (!ok) for (int kk = k0; kk < k1; kk++) mark_bad(a.flag, kk);
    return;
  }
  const u8* __restrict__ l = a.li + (size_t)k0 * a.mi + (p0 > 0 ? p0 - 1 : 0);
  for (int k = k0; k < k1; k++) {
    const bool ok = p0 > 0 && (a.ibi[k] == 0 || *l);
    __stcs(go, ok ? __ldg(g) : (R)0); __stcs(lo, (u8)(ok ? 1 : 0));
    if (!ok) mark_bad(a.flag, k);
    g += a.mi; l += a.mi; go += a.mo; lo += a.mo;
  }
}

// ---- neighbor, scalar: gather in 8 x 4 sub-patches, store 32-wide ---------------------------------------------
// k_neighbor_lane is bound by L1 wavefronts (93 % of peak): a warp of 32 consecutive output points pulls ~15 sectors
// per gather when the output rows cut across the source rows, on top of the ~6 its two stores cost.  Here a WARP owns
// a 32 x 4 patch of the output grid, four points per lane.  For the gathers the lane's points are (8s + lane%8,
// lane/8), s = 0..3: each warp instruction covers an 8 x 4 sub-patch, a compact footprint on the source grid (~6
// sectors).  The values cross the warp's own shared-memory tile (row pitch 40 words: the four rows of a sub-patch land
// in distinct banks) and are stored row by row, 32 consecutive points per instruction.  Warps never meet a block
// barrier; four fields (16 gathers per lane) are in flight, the next round's are issued before this round's stores.
// MASKED: the batch has fields with a bitmap; their li bytes are gathered next to the values, in the same sub-patch
// mapping, and the usable bits travel to the storing lanes by shuffle.
template <class R, bool MASKED>
__global__ void __launch_bounds__(128)
k_neighbor_patch(FieldArgs<R> a, const int* __restrict__ nxy, int kpb, int rowlen, int ntx, int npatch, int amask, int period, int base_e) {
  constexpr int FB = 4, PITCH = 40;
  __shared__ R tile_all[4][2][FB][4 * PITCH];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int patch = blockIdx.x * 4 + w;
  if (patch >= npatch) return;
  R (*tile)[FB][4 * PITCH] = tile_all[w];
  const int col0 = (patch % ntx) * 32, row0 = (patch / ntx) * 4;
  // Fields with the same element phase of go(:,k) form a class (k = cls + period * j); the 32-point run of every output
  // row is shifted per class and row so that it starts on a 128-byte line of the caller's go array and on a 32-byte
  // sector of lo (stores alone, tools/wr_bench2.cu: runs that straddle lines cost 2-3x).  amask = 0: no classes.
  const int nslice = gridDim.y / period;
  const int cls = blockIdx.y / nslice, slice = blockIdx.y % nslice;
  const int phase = (base_e + (int)(((long long)cls * a.mo) & amask)) & amask;
  auto shift_of = [&](int row) { return (int)(((long long)phase + (long long)row * rowlen) & amask); };
  int pg[4];                   // gather points: 0-based source position, -1 = none
  long long ns[4];             // store points: output index, -1 = outside the grid
  u8 oks[4];
  bool anybad = false;
  {
    // the eight plan loads of the patch are issued together, before any is consumed (one memory latency instead of four:
    // ncu showed a quarter of the warps' time parked on these loads when they were taken one sub-patch at a time)
    long long ng[4];
    int vg[4], vs[4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
      const int gr = row0 + (lane >> 3), gc = col0 - shift_of(gr) + 8 * s + (lane & 7);
      ng[s] = (long long)gr * rowlen + gc;
      if (!(gc >= 0 && gc < rowlen && ng[s] < a.no)) ng[s] = -1;
      const int sr = row0 + s, sc = col0 - shift_of(sr) + lane;
      const long long n = (long long)sr * rowlen + sc;
      ns[s] = (sc >= 0 && sc < rowlen && n < a.no) ? n : -1;
    }
#pragma unroll
    for (int s = 0; s < 4; s++) { vg[s] = __ldg(nxy + (ng[s] >= 0 ? ng[s] : 0)); vs[s] = __ldg(nxy + (ns[s] >= 0 ? ns[s] : 0)); }
#pragma unroll
    for (int s = 0; s < 4; s++) {
      pg[s] = ng[s] >= 0 ? vg[s] - 1 : -1;
      oks[s] = (ns[s] >= 0 && vs[s] > 0) ? 1 : 0;
      anybad = anybad || (ns[s] >= 0 && !oks[s]);
    }
  }
  // fields of this block: kf(j) = cls + period * j, j in [k0, k1)   (period = 1, cls = 0 without classes)
  const int k0 = slice * kpb, k1 = min(k0 + kpb, (a.kc - cls + period - 1) / period);
  if (k0 >= k1) return;
  const size_t gstep = (size_t)period * a.mi, ostep = (size_t)period * a.mo;
  const R* __restrict__ g = a.gi + (size_t)(cls + period * k0) * a.mi;
  const u8* __restrict__ l = MASKED ? a.li + (size_t)(cls + period * k0) * a.mi : nullptr;
  R* go = a.go + (size_t)(cls + period * k0) * a.mo;
  u8* lo = a.lo + (size_t)(cls + period * k0) * a.mo;
  const int wslot = (lane >> 3) * PITCH + (lane & 7);
  // values (and, MASKED, bitmap bytes: 1 where the field has none) of one round at this lane's gather points
  auto gather = [&](R (&v)[FB][4], u8 (&m)[FB][4], int k) {
#pragma unroll
    for (int f = 0; f < FB; f++) {
      const bool mk = MASKED && k + f < k1 && a.ibi[cls + period * (k + f)] != 0;
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const bool on = pg[s] >= 0 && k + f < k1;
        v[f][s] = on ? __ldg(g + (size_t)f * gstep + pg[s]) : (R)0;
        m[f][s] = (MASKED && mk && on) ? __ldg(l + (size_t)f * gstep + pg[s]) : (u8)1;
      }
    }
  };
  R v[FB][4], vn[FB][4];
  u8 m[FB][4], mn[FB][4];
  gather(v, m, k0);
  int buf = 0;
  for (int k = k0; k < k1; k += FB) {
    unsigned goodbits = 0;     // MASKED: bit 4f+s: this lane's gather point s has a usable source point in field k+f
#pragma unroll
    for (int f = 0; f < FB; f++) {
#pragma unroll
      for (int s = 0; s < 4; s++) {
        const bool ok = pg[s] >= 0 && (!MASKED || m[f][s]);
        if (MASKED && ok) goodbits |= 1u << (4 * f + s);
        tile[buf][f][wslot + 8 * s] = (!MASKED || ok) ? v[f][s] : (R)0;
      }
    }
    g += (size_t)FB * gstep;
    if (MASKED) l += (size_t)FB * gstep;
    if (k + FB < k1) gather(vn, mn, k + FB);     // in flight while this round is stored
    __syncwarp();
    const int nf = min(FB, k1 - k);
    unsigned gbs[4];           // MASKED: the bits of the lanes that computed this lane's four store points
#pragma unroll
    for (int r = 0; r < 4; r++) gbs[r] = MASKED ? __shfl_sync(0xffffffffu, goodbits, r * 8 + (lane & 7)) : 0u;
#pragma unroll
    for (int f = 0; f < FB; f++) {
      if (f < nf) {
        bool bad = false;
#pragma unroll
        for (int r = 0; r < 4; r++) {
          if (ns[r] >= 0) {
            // store point (row r, column lane) is sub-patch point s = lane/8 of lane r*8 + lane%8
            const u8 ok = MASKED ? (u8)((gbs[r] >> (4 * f + (lane >> 3))) & 1u) : oks[r];
            __stcs(go + (size_t)f * ostep + ns[r], tile[buf][f][r * PITCH + lane]);
            __stcs(lo + (size_t)f * ostep + ns[r], ok);
            bad = bad || !ok;
          }
        }
        if (MASKED && bad) mark_bad(a.flag, cls + period * (k + f));
      }
    }
    go += (size_t)FB * ostep; lo += (size_t)FB * ostep;
#pragma unroll
    for (int f = 0; f < FB; f++) {
#pragma unroll
      for (int s = 0; s < 4; s++) { v[f][s] = vn[f][s]; m[f][s] = mn[f][s]; }
    }
    buf ^= 1;   // the other buffer was last read two rounds ago, before the previous __syncwarp
  }
  if (!MASKED && anybad) for (int k = k0; k < k1; k++) mark_bad(a.flag, cls + period * k);
}

template <class R> bool fast_neighbor_ok(const Plan<R>& P, const FieldArgs<R>& a) {
  return P.method == M_NEIGHBOR && !P.vec && a.mspiral <= 1 && P.no > 0;
}

template <class R> void launch_fast_neighbor(const Plan<R>& P, const FieldArgs<R>& a, cudaStream_t st) {
  static const bool no_patch = getenv("IPB_NPATCH") && atoi(getenv("IPB_NPATCH")) == 0;
  const GridP<R>& go = P.gout.p;
  int orow = 0;
  if (go.type != P_STATION && go.im > 0 && go.jm > 0) orow = go.nscan == 0 ? go.im : go.jm;
  if (!no_patch && orow >= 32 && orow < a.no) {
    // few fields per block: the blocks resident together then write neighbouring patches of the same fields
    // (measured: 12 best, profiles/README.md)
    static const int kpb_env = getenv("IPB_NKPB") ? atoi(getenv("IPB_NKPB")) : 32;
    // line-aligned runs: alignment unit = 128 bytes of go (32 / 16 elements); needs go and lo to share their element phase
    static const int nalign_env = getenv("IPB_NALIGN") ? atoi(getenv("IPB_NALIGN")) : 128;
    int unit = std::max(1, nalign_env / (int)sizeof(R));
    const uintptr_t gp = reinterpret_cast<uintptr_t>(a.go), lp = reinterpret_cast<uintptr_t>(a.lo);
    if (gp % sizeof(R) || ((gp / sizeof(R)) & (unit - 1)) != (lp & (unit - 1)) || (unit & (unit - 1))) unit = 1;
    int s_ = (int)(a.mo & (unit - 1)), g_ = unit;
    while (s_) { const int t = g_ % s_; g_ = s_; s_ = t; }
    const int period = unit / g_;
    const int per_class = nblk(a.kc, period);
    const int kpb = std::max(4, std::min(per_class, kpb_env));
    const int ntx = nblk(orow + unit - 1, 32), npatch = ntx * nblk(nblk(a.no, orow), 4);
    dim3 grid(nblk(npatch, 4), period * nblk(per_class, kpb));
    const int base_e = (int)((gp / sizeof(R)) & (unit - 1));
    if (a.li) k_neighbor_patch<R, true><<<grid, 128, 0, st>>>(a, P.nxy, kpb, orow, ntx, npatch, unit - 1, period, base_e);
    else k_neighbor_patch<R, false><<<grid, 128, 0, st>>>(a, P.nxy, kpb, orow, ntx, npatch, unit - 1, period, base_e);
    g_launches++;
    return;
  }
  // station points (no rows to cut patches from): one point per lane.  Measured on S -> E (km = 1024): plain linear blocks of
  // 256 points with 32 fields per thread are as fast as 32 x 8 patches of one point per lane (2.04 vs 2.25 ms); the
  // kernel sits at 93 % of the L1TEX wavefront peak (profiles/README.md)
  const int kpb = std::max(1, std::min(a.kc, 32));
  const int rowlen = 256;
  dim3 grid = patch_grid(rowlen, a.no, 8);
  grid.z = nblk(a.kc, kpb);
  k_neighbor_lane<R><<<grid, 256, 0, st>>>(a, P.nxy, kpb, rowlen);
  g_launches++;
}

}  // namespace ipb
