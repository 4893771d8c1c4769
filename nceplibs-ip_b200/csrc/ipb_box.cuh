// ipb200 — the TMA-staged bilinear apply kernel (ipolates ip=0, the BASELINE headline path; round 2).
//
// What the reference does per (output point, field) — src/bilinear_interp_mod.F90:202-261 — is a 4-tap weighted sum, a
// divide and a 5-byte store.  With an output grid that oversamples the source (3 km Lambert from 0.25 degree: 66 output
// points per touched source point) the loop is a write-bound stream, 5.09 algorithmic bytes per point-field.  The
// round-1 kernel (ipb_tiles.cuh) gathered each tile's source cells with 4-byte cp.async (LDGSTS) and kept four points
// per lane; ncu showed it bound by the load/store unit (L1TEX 81 %) and by issue slots at 16 resident warps per SM.
// This kernel changes the data movement:
//
//  * a warp owns a PATCH of the output grid: 4 consecutive output rows x 32 consecutive points, one point per lane and
//    row.  The source footprint of such a patch is a small rectangle of the (regular, row-major) source grid, so per
//    pass of 8 fields the rectangle is fetched with ONE `cp.async.bulk.tensor` (TMA; 4-D tensor map x / y / class /
//    field-of-class over the caller's gi, box 16 x 4 x 1 x 8) into the warp's shared-memory slice, completion
//    signalled on an mbarrier, double buffered.  No per-lane staging instructions, no address lists in the plan: a
//    point only knows the (ix, iy) of its cell, packed in one word, and the tables do not depend on the store alignment;
//  * the four taps of a point are four shared-memory loads at register + immediate addresses
//    (base, +1, +pitch, +pitch+1), the weighted sum is the reference's 4 multiplies and 3 adds in its order, the
//    quotient G/W comes from the stored correctly rounded 1/W by two fused residual steps (Markstein: the correctly
//    rounded quotient, i.e. the IEEE division's bits; sums outside the exact-residual window take the IEEE division);
//  * go is stored 32 lanes wide — one full 128-byte line per warp instruction: the 32-point run of every row is
//    shifted per field class so that its first element sits on a 128-byte boundary of the caller's go array
//    (field k starts at element k*mo, rows at j*rowlen: the shift is (phase_k + j*rowlen) mod 32 elements).  Fields
//    with the same phase form a class; a block works on one class.  lo goes out as one 32-bit word per 4 points.
//    Anything but whole, line-aligned 128-byte runs is expensive on this part (stores alone, tools/wr_bench2.cu:
//    128-byte aligned runs 1.4-2.0 ms for the 9.75 GB of a km = 1024 call, 64-byte aligned 2.1-3.5 ms, 16-byte 4-6 ms);
//  * blockIdx.x runs over (class, field slice), blockIdx.y / z over patches, so blocks that are resident together share
//    a patch's plan data through L2 while streaming different fields.
//
// Patches whose cells do not fit the box (16 x 4 source points), that contain an irregular stencil (grid edge,
// x-wrap, pole fold), a point failing the mask threshold, and fields that carry a bitmap are computed by the general
// per-point code (global gathers) inside the same kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "ipb_tiles.cuh"
#include "ipb_tma.cuh"

namespace ipb {

constexpr int BOX_BX = 16, BOX_BY = 4;       // source points per staged rectangle (x, y)

struct BoxGeom {
  int rowlen, nrow, ntx, nty;   // output row length, rows, patches per row / column
  int amask;                    // alignment unit of the 32-point runs in elements, minus one (31: 128-byte lines in the _4 kind)
  int period, plog;             // fields k and k + period share their phase (number of classes, a power of two), log2
  int base_e;                   // element phase of the caller's go pointer
  int jfull;                    // kc / period: field rows j < jfull exist in every class (bound of the 4-D tensor map)
};

// ---- plan time: the cell of every output point -----------------------------------------------------------------------
// bxy[n] = iy << 16 | ix (0-based source column / row of tap (1,1)) when the four taps are n, n+1, n+im, n+im+1;
// -1 otherwise (missing tap, x-wrap, pole fold, column-major source).
// wmin = the smallest weight sum over the taps that exist, over all points (bit pattern of a non-negative binary64):
// without a bitmap every point passes the mask threshold pmp iff wmin >= pmp, and then lo is 1 everywhere.
template <class R>
__global__ void k_box_index(int no, int im, const int* __restrict__ nxy, const R* __restrict__ wxy, int* __restrict__ bxy, unsigned long long* wmin) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  double ws = 1e300;
  if (n < no) {
    const int t0 = nxy[n], t1 = nxy[(size_t)no + n], t2 = nxy[2 * (size_t)no + n], t3 = nxy[3 * (size_t)no + n];
    int v = -1;
    if (t0 > 0 && t1 == t0 + 1 && t2 == t0 + im && t3 == t2 + 1) {
      const int ix = (t0 - 1) % im, iy = (t0 - 1) / im;
      if (ix + 1 < im && iy < 32767) v = (iy << 16) | ix;
    }
    bxy[n] = v;
    R s = 0;   // bilinear_interp_mod.F90:210-219: W accumulates the weights of the taps that exist, in tap order
    if (t0 > 0) s = s + wxy[n];
    if (t1 > 0) s = s + wxy[(size_t)no + n];
    if (t2 > 0) s = s + wxy[2 * (size_t)no + n];
    if (t3 > 0) s = s + wxy[3 * (size_t)no + n];
    ws = s > (R)0 ? (double)s : 0.0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ws = fmin(ws, __shfl_xor_sync(0xffffffffu, ws, o));
  if ((threadIdx.x & 31) == 0) atomicMin(wmin, (unsigned long long)__double_as_longlong(ws));
}

template <class R> void build_box_tables(Plan<R>& P, cudaStream_t st) {
  P.box_ok = false;
  if (P.method != M_BILINEAR || P.vec || P.no <= 0 || !P.nxy || !P.wq) return;
  const GridP<R>& gi = P.gin.p;
  if (gi.im <= 1 || gi.im > 65535 || gi.jm <= 1 || gi.jm > 32767) return;
  if ((long long)gi.im * gi.jm > P.mi) return;
  if (((size_t)gi.im * sizeof(R)) % 16 || ((size_t)P.mi * sizeof(R)) % 16) return;   // tensor-map strides are multiples of 16 bytes
  P.bxy = dalloc<int>(P.no);
  unsigned long long* d_wmin = dalloc<unsigned long long>(1);
  const unsigned long long big = 0x7fefffffffffffffull;
  IPB_CUDA(cudaMemcpyAsync(d_wmin, &big, sizeof(big), cudaMemcpyHostToDevice, st));
  k_box_index<R><<<nblk(P.no, 256), 256, 0, st>>>(P.no, gi.im, P.nxy, P.wxy, P.bxy, d_wmin);
  g_launches++;
  unsigned long long hmin = 0;
  IPB_CUDA(cudaMemcpyAsync(&hmin, d_wmin, sizeof(hmin), cudaMemcpyDeviceToHost, st));
  IPB_CUDA(cudaStreamSynchronize(st));
  dfree(d_wmin);
  memcpy(&P.box_wmin, &hmin, sizeof(double));
  P.bytes += (size_t)P.no * 4;
  P.box_ok = true;
}

template <class R> __device__ __forceinline__ R lds_r(unsigned addr);
template <> __device__ __forceinline__ float lds_r<float>(unsigned addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
template <> __device__ __forceinline__ double lds_r<double>(unsigned addr) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr)); return v; }

// the NR weighted sums and quotients of one lane and field
template <class R, int NR> struct BoxRows {
  static __device__ __forceinline__ void run(const R (&v)[NR][4], const R (&w)[NR][4], const R (&W)[NR], const R (&rW)[NR], R (&G)[NR], R (&o)[NR]) {
#pragma unroll
    for (int r = 0; r < NR; r++) {
      // the reference starts from G=0; 0 + x only changes x = -0, which the quotient maps to +0 again
      R g = w[r][0] * v[r][0];
#pragma unroll
      for (int t = 1; t < 4; t++) g = g + w[r][t] * v[r][t];
      G[r] = g;
      o[r] = quot_near1<R>(g, W[r], rW[r]);
    }
  }
};
// binary32: the same operations with the products and the quotient steps issued two at a time (FMUL2 / FFMA2)
template <int NR> struct BoxRows<float, NR> {
  static __device__ __forceinline__ void run(const float (&v)[NR][4], const float (&w)[NR][4], const float (&W)[NR], const float (&rW)[NR],
                                             float (&G)[NR], float (&o)[NR]) {
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const float2 pa = mul2(make_float2(w[r][0], w[r][1]), make_float2(v[r][0], v[r][1]));
      const float2 pb = mul2(make_float2(w[r][2], w[r][3]), make_float2(v[r][2], v[r][3]));
      G[r] = ((pa.x + pa.y) + pb.x) + pb.y;
    }
#pragma unroll
    for (int h = 0; h < NR / 2; h++) {
      const float2 g = make_float2(G[2 * h], G[2 * h + 1]);
      const float2 nw = make_float2(-W[2 * h], -W[2 * h + 1]), rw = make_float2(rW[2 * h], rW[2 * h + 1]);
      float2 r = fma2(g, nw, g);
      const float2 q1 = fma2(r, rw, g);
      r = fma2(q1, nw, g);
      const float2 q = fma2(r, rw, q1);
      o[2 * h] = q.x; o[2 * h + 1] = q.y;
    }
  }
};

// NR = output rows of a patch (one point per lane and row), F = fields staged per pass, THREADS = block size (the warps
// of a block own neighbouring patches of one output row band), MINB = resident blocks per SM the register budget allows.
// tm4: (x, y, class, field-of-class), box BX x BY x 1 x F — whole passes; tm3: (x, y, field), box BX x BY x 1 — the
// ragged last pass of a batch whose size is not a multiple of period x F.
template <class R, int NR, int F, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_bilinear_box(const __grid_constant__ CUtensorMap tm4, const __grid_constant__ CUtensorMap tm3, FieldArgs<R> a, const int* __restrict__ nxy,
               const int* __restrict__ bxy, const R* __restrict__ wq, const R* __restrict__ ws2, BoxGeom G, int kpt) {
  constexpr int BX = BOX_BX, BY = BOX_BY;
  constexpr int BOXB = BX * BY * (int)sizeof(R);   // bytes of one staged rectangle (256 / 512: a multiple of 128)
  constexpr int PITCH = BX * (int)sizeof(R);
  constexpr int WARPS = THREADS / 32;
  extern __shared__ __align__(128) unsigned char sm_box[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned sm_base = (unsigned)__cvta_generic_to_shared(sm_box);
  const unsigned sbuf = sm_base + warp * (2 * F * BOXB);                       // two buffers of F rectangles
  const unsigned sbar = sm_base + WARPS * (2 * F * BOXB) + warp * 16;          // two mbarriers
  const int cls = blockIdx.x & (G.period - 1), slice = blockIdx.x >> G.plog;
  const int tx = blockIdx.y * WARPS + warp, ty = blockIdx.z;
  // fields of this block: k = cls + period * j, j in [jbeg, jend)
  const int jbeg = kpt * slice, jend = min(kpt * (slice + 1), (a.kc - cls + G.period - 1) >> G.plog);
  if (jbeg >= jend || tx >= G.ntx) return;
  const int phase = (G.base_e + cls * (a.mo & G.amask)) & G.amask;

  // ---- the patch: one point per lane and row ----
  int npt[NR];          // output point of (row r, lane), -1 outside
  unsigned soff[NR];    // byte offset of its cell inside a staged rectangle
  R w[NR][4], W[NR], rW[NR];
  int minx = 1 << 30, miny = 1 << 30, maxx = -1, maxy = -1;
  bool regular = true, full = true, whole = true, some = false;   // whole: every lane of every row has a point
  int cxy[NR];
  float4 q4[sizeof(R) == 4 ? NR : 1]; float2 s4[sizeof(R) == 4 ? NR : 1];
  double2 q8a[sizeof(R) == 8 ? NR : 1], q8b[sizeof(R) == 8 ? NR : 1], s8[sizeof(R) == 8 ? NR : 1];
  {
    // all plan loads of the patch are issued before anything consumes one (twelve independent loads, one memory
    // latency; the compiler otherwise waits row by row — ncu showed a third of the warps' time parked there)
    size_t nn[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const int j = ty * NR + r;
      const int n0 = j * G.rowlen;                       // first point of the output row (n0 + rowlen <= 2^31: checked by the launcher)
      const int i = tx * 32 - ((phase + n0) & G.amask) + lane;
      const bool in = j < G.nrow && i >= 0 && i < G.rowlen && n0 + i < a.no;
      npt[r] = in ? n0 + i : -1;
      nn[r] = in ? (size_t)(n0 + i) : 0;                 // lanes outside the grid read point 0 and ignore it
    }
#pragma unroll
    for (int r = 0; r < NR; r++) cxy[r] = __ldg(bxy + nn[r]);
    if (sizeof(R) == 4) {
#pragma unroll
      for (int r = 0; r < NR; r++) { q4[r] = __ldg(reinterpret_cast<const float4*>(wq) + nn[r]); s4[r] = __ldg(reinterpret_cast<const float2*>(ws2) + nn[r]); }
    } else {
#pragma unroll
      for (int r = 0; r < NR; r++) {
        q8a[r] = __ldg(reinterpret_cast<const double2*>(wq) + 2 * nn[r]); q8b[r] = __ldg(reinterpret_cast<const double2*>(wq) + 2 * nn[r] + 1);
        s8[r] = __ldg(reinterpret_cast<const double2*>(ws2) + nn[r]);
      }
    }
    // the cells come first: box origin, fit, and the first TMA boxes leave before the weights are looked at
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const bool in = npt[r] >= 0;
      whole = whole && in;
      some = some || in;
      const int c = cxy[r];
      if (in) {
        if (c < 0) regular = false;
        else { minx = min(minx, c & 0xffff); maxx = max(maxx, c & 0xffff); miny = min(miny, c >> 16); maxy = max(maxy, c >> 16); }
      }
    }
  }
  if (!__any_sync(0xffffffffu, some)) return;   // patch entirely outside the grid
  minx = __reduce_min_sync(0xffffffffu, minx); miny = __reduce_min_sync(0xffffffffu, miny);
  maxx = __reduce_max_sync(0xffffffffu, maxx); maxy = __reduce_max_sync(0xffffffffu, maxy);
  // the box starts on a 16-byte boundary of the source row (a TMA requirement for the innermost coordinate: measured,
  // tools/tma_probe.cu — an unaligned start raises "illegal instruction")
  minx &= ~(16 / (int)sizeof(R) - 1);
  const bool fit = __all_sync(0xffffffffu, regular) && maxx >= 0 && (maxx + 1 - minx) < BX && (maxy + 1 - miny) < BY;
  auto issue = [&](int buf, int j0) {   // lane 0 only
    const unsigned bar = sbar + 8 * buf, dst = sbuf + (unsigned)(buf * F * BOXB);
    if (j0 + F <= G.jfull) {
      mbar_expect_tx(bar, (unsigned)(F * BOXB));
      tma_box_4d(dst, &tm4, minx, miny, cls, j0, bar);
    } else {          // the batch's ragged end: field by field
      const int nf = min(F, jend - j0);
      mbar_expect_tx(bar, (unsigned)(nf * BOXB));
      for (int f = 0; f < nf; f++) tma_box_3d(dst + (unsigned)(f * BOXB), &tm3, minx, miny, cls + G.period * (j0 + f), bar);
    }
  };
  if (fit) {
    if (lane == 0) { mbar_init(sbar, 1); mbar_init(sbar + 8, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    if (lane == 0) issue(0, jbeg);
  }
  // the weights (their loads have been in flight since before the cells were reduced)
#pragma unroll
  for (int r = 0; r < NR; r++) {
    if (npt[r] >= 0) {
      if (sizeof(R) == 4) { w[r][0] = (R)q4[r].x; w[r][1] = (R)q4[r].y; w[r][2] = (R)q4[r].z; w[r][3] = (R)q4[r].w; W[r] = (R)s4[r].x; rW[r] = (R)s4[r].y; }
      else { w[r][0] = (R)q8a[r].x; w[r][1] = (R)q8a[r].y; w[r][2] = (R)q8b[r].x; w[r][3] = (R)q8b[r].y; W[r] = (R)s8[r].x; rW[r] = (R)s8[r].y; }
      const R dev = W[r] - (R)1;
      full = full && W[r] >= a.pmp && dev <= (R)9.5e-7 && dev >= (R)-9.5e-7;
    } else {
      w[r][0] = w[r][1] = w[r][2] = w[r][3] = (R)0; W[r] = (R)1; rW[r] = (R)1;
    }
  }
  // staged: every stencil regular and complete, every point passing the threshold with W ~ 1, the cells inside one box
  const bool staged = fit && __all_sync(0xffffffffu, full);
  const bool whole_w = __all_sync(0xffffffffu, whole);
#pragma unroll
  for (int r = 0; r < NR; r++) {
    const int c = cxy[r];
    soff[r] = (npt[r] >= 0 && staged) ? (unsigned)((((c >> 16) - miny) * BX + ((c & 0xffff) - minx)) * (int)sizeof(R)) : 0u;
  }
  const size_t ostep = (size_t)G.period * a.mo;
  const unsigned ostep_l = (unsigned)ostep, ostep_b = (unsigned)(ostep * sizeof(R));   // F * ostep_b < 2^32 (checked by the launcher)

  auto general_field = [&](int k) {
    // general path, one field: bitmaps, partial stencils, threshold failures, boxes that do not fit
    const R* __restrict__ g = a.gi + (size_t)k * a.mi;
    const bool masked = a.ibi[k] != 0;
    const u8* __restrict__ l = (masked && a.li) ? a.li + (size_t)k * a.mi : nullptr;
    bool bad = false;
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const int n = npt[r];
      if (n < 0) continue;
      int idx[4];
#pragma unroll
      for (int t = 0; t < 4; t++) idx[t] = nxy[(size_t)t * a.no + n];
      R o;
      const bool good = bilinear_point<R>(g, l, masked && l, idx, w[r], a.pmp, o);
      a.go[(size_t)k * a.mo + n] = o;
      a.lo[(size_t)k * a.mo + n] = good ? 1 : 0;
      bad = bad || !good;
    }
    if (bad) mark_bad(a.flag, k);
  };
  if (!staged) {
    if (fit) mbar_wait(sbar, 0);   // boxes were requested before the weights said no: let them land before the block may retire
    for (int j = jbeg; j < jend; j++) general_field(cls + G.period * j);
    return;
  }

  // ---- staged path ----
  // does any field of this warp carry a bitmap?  (uniform; the common answer is no, and the pass loop then skips the bookkeeping)
  bool anymask = false;
  for (int j = jbeg + lane; j < jend; j += 32) anymask = anymask || a.ibi[cls + G.period * j] != 0;
  anymask = __any_sync(0xffffffffu, anymask);
  // per-row output pointers of this lane (field jbeg) and the lo word pointer; advanced pass by pass
  R* gp[NR];
#pragma unroll
  for (int r = 0; r < NR; r++) gp[r] = a.go + (size_t)(cls + G.period * jbeg) * a.mo + (npt[r] >= 0 ? npt[r] : 0);
  u8* lop = a.lo + (size_t)(cls + G.period * jbeg) * a.mo;
  // lo words: lane 8r + q stores the word of points 4q .. 4q+3 of row r (whole patches only)
  u8* lwp;
  {
    const int r = lane >> 3, j = ty * NR + r, n0 = j * G.rowlen;
    lwp = lop + (n0 + tx * 32 - ((phase + n0) & G.amask) + 4 * (lane & 7));
  }
  const bool lo_lane = lane < 8 * NR;
  unsigned par = 0;   // bit b: parity to wait for on buffer b
  int buf = 0;
  for (int j0 = jbeg; j0 < jend; j0 += F) {
    if (lane == 0 && j0 + F < jend) issue(buf ^ 1, j0 + F);
    mbar_wait(sbar + 8 * buf, (par >> buf) & 1u);
    par ^= 1u << buf;
    unsigned passmask = 0;   // bit f: field f of the pass carries a bitmap
    if (anymask) passmask = __ballot_sync(0xffffffffu, lane < F && j0 + lane < jend && a.ibi[cls + G.period * (j0 + lane)] != 0);
    const unsigned sb = sbuf + (unsigned)(buf * F * BOXB);
    unsigned sr[NR];
#pragma unroll
    for (int r = 0; r < NR; r++) sr[r] = sb + soff[r];
    bool done = false;
    if (j0 + F <= jend && passmask == 0 && whole_w) {
      // whole pass, no bitmaps, no row edge: straight-line code.  amax / amin collect the window test of the fast quotient.
      R amax = win_lo<R>(), amin = win_lo<R>();
#pragma unroll
      for (int f = 0; f < F; f++) {
        R v[NR][4], Gs[NR], o[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) {
          v[r][0] = lds_r<R>(sr[r] + f * BOXB);
          v[r][1] = lds_r<R>(sr[r] + f * BOXB + (int)sizeof(R));
          v[r][2] = lds_r<R>(sr[r] + f * BOXB + PITCH);
          v[r][3] = lds_r<R>(sr[r] + f * BOXB + PITCH + (int)sizeof(R));
        }
        BoxRows<R, NR>::run(v, w, W, rW, Gs, o);
#pragma unroll
        for (int r = 0; r < NR; r++)
          __stcs(reinterpret_cast<R*>(reinterpret_cast<char*>(gp[r]) + (size_t)((unsigned)f * ostep_b)), o[r]);
#pragma unroll
        for (int h = 0; h < NR; h += 2) {
          amax = max_r<R>(max_r<R>(amax, abs_r<R>(Gs[h])), abs_r<R>(Gs[h + 1]));
          amin = min_r<R>(min_r<R>(amin, abs_r<R>(Gs[h])), abs_r<R>(Gs[h + 1]));
        }
        if (lo_lane) __stcs(reinterpret_cast<unsigned*>(lwp + (size_t)((unsigned)f * ostep_l)), 0x01010101u);
      }
      // rare: a sum that is zero, tiny, huge or not finite lies outside the exact-residual window of the fast
      // quotient: the pass is redone below with the IEEE division
      const bool ok = amax < win_hi<R>() && amin >= win_lo<R>();
      done = __all_sync(0xffffffffu, ok);
    }
    if (!done) {
#pragma unroll 1
      for (int f = 0; f < F && j0 + f < jend; f++) {
        const int k = cls + G.period * (j0 + f);
        if ((passmask >> f) & 1u) { general_field(k); continue; }
#pragma unroll
        for (int r = 0; r < NR; r++) {
          const unsigned s = sr[r] + (unsigned)(f * BOXB);
          R g = w[r][0] * lds_r<R>(s);
          g = g + w[r][1] * lds_r<R>(s + (int)sizeof(R));
          g = g + w[r][2] * lds_r<R>(s + PITCH);
          g = g + w[r][3] * lds_r<R>(s + PITCH + (int)sizeof(R));
          const R q = ((R)0 + g) / W[r];
          if (npt[r] >= 0) {
            __stcs(gp[r] + (size_t)f * ostep, q);
            __stcs(lop + (size_t)f * ostep + npt[r], (u8)1);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < NR; r++) gp[r] += (size_t)F * ostep;
    lop += (size_t)F * ostep; lwp += (size_t)F * ostep;
    __syncwarp();
    buf ^= 1;
  }
}

template <class R> bool box_ok(const Plan<R>& P, const FieldArgs<R>& a) {
  static const int off = getenv("IPB_NO_BOX") ? atoi(getenv("IPB_NO_BOX")) : 0;
  if (off || !P.box_ok || !fast_bilinear_ok(P, a)) return false;
  if (reinterpret_cast<uintptr_t>(a.gi) % 16) return false;   // tensor-map base address
  return true;
}

template <class R, int F> bool launch_bilinear_box_f(const Plan<R>& P, const FieldArgs<R>& a, cudaStream_t st) {
  PFN_tmap_encode enc = tmap_encoder();
  if (!enc) return false;
  const GridP<R>& gi = P.gin.p;
  static const int align_env = getenv("IPB_BOX_ALIGN") ? atoi(getenv("IPB_BOX_ALIGN")) : 0;   // alignment unit of the store runs, bytes
  static const int kpt_env = getenv("IPB_BOX_KPT") ? atoi(getenv("IPB_BOX_KPT")) : 32;
  static const int warps_env = getenv("IPB_BOX_WARPS") ? atoi(getenv("IPB_BOX_WARPS")) : 4;
  static const int minb_env = getenv("IPB_BOX_MINB") ? atoi(getenv("IPB_BOX_MINB")) : 6;
  constexpr int NR = 4;
  int unit = (align_env >= 16 && align_env <= 128 ? align_env : 128) / (int)sizeof(R);       // elements; a power of two >= 4 (lo words)
  if (unit < 4) unit = 4;
  const int WARPS = warps_env == 8 ? 8 : 4;
  BoxGeom G;
  G.rowlen = patch_rowlen(P.gout.p, P.no);
  G.nrow = nblk(P.no, G.rowlen);
  if ((long long)G.nrow * G.rowlen + G.rowlen >= (1ll << 31)) return false;
  G.amask = unit - 1;
  G.ntx = (G.rowlen + G.amask + 31) / 32;
  G.nty = nblk(G.nrow, NR);
  int s = (int)(a.mo & G.amask), g = unit;   // period = unit / gcd(mo mod unit, unit)
  while (s) { const int t = g % s; g = s; s = t; }
  G.period = unit / g;
  G.plog = 0;
  while ((1 << G.plog) < G.period) G.plog++;
  G.base_e = (int)((reinterpret_cast<uintptr_t>(a.go) / sizeof(R)) & G.amask);
  G.jfull = a.kc / G.period;
  const int per_class = (a.kc + G.period - 1) / G.period;
  const int kpt = std::max(F, std::min((kpt_env / F) * F, ((per_class + F - 1) / F) * F));   // fields of one class per warp: whole passes
  dim3 grid(G.period * nblk(per_class, kpt), nblk(G.ntx, WARPS), G.nty);
  if (grid.z > 65535 || grid.y > 65535) return false;
  if ((unsigned long long)F * G.period * a.mo * sizeof(R) >= (1ull << 32)) return false;   // 32-bit byte offsets inside a pass
  // tensor maps over the caller's gi: (x, y, class, field-of-class) for whole passes, (x, y, field) for the ragged end
  CUtensorMap tm4, tm3;
  const CUtensorMapDataType dt = sizeof(R) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  {
    const cuuint64_t dims[3] = {(cuuint64_t)gi.im, (cuuint64_t)gi.jm, (cuuint64_t)a.kc};
    const cuuint64_t strides[2] = {(cuuint64_t)gi.im * sizeof(R), (cuuint64_t)a.mi * sizeof(R)};
    const cuuint32_t box[3] = {BOX_BX, BOX_BY, 1}, es[3] = {1, 1, 1};
    if (enc(&tm3, dt, 3, const_cast<R*>(a.gi), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
  }
  if (G.jfull > 0) {
    const cuuint64_t dims[4] = {(cuuint64_t)gi.im, (cuuint64_t)gi.jm, (cuuint64_t)G.period, (cuuint64_t)G.jfull};
    const cuuint64_t strides[3] = {(cuuint64_t)gi.im * sizeof(R), (cuuint64_t)a.mi * sizeof(R), (cuuint64_t)G.period * a.mi * sizeof(R)};
    const cuuint32_t box[4] = {BOX_BX, BOX_BY, 1, F}, es[4] = {1, 1, 1, 1};
    if (strides[2] >= (1ull << 40)) return false;
    if (enc(&tm4, dt, 4, const_cast<R*>(a.gi), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
  } else {
    tm4 = tm3;   // never used: every pass is ragged
  }
#define IPB_BOX_LAUNCH(TH, MB)                                                                                               \
  {                                                                                                                          \
    constexpr int SM = ((TH) / 32) * (2 * F * BOX_BX * BOX_BY * (int)sizeof(R) + 16);                                        \
    static bool once[64] = {};                                                                                               \
    int dv_ = 0; cudaGetDevice(&dv_); dv_ &= 63;                                                                             \
    if (!once[dv_]) { IPB_CUDA(cudaFuncSetAttribute(k_bilinear_box<R, NR, F, TH, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM)); once[dv_] = true; } \
    k_bilinear_box<R, NR, F, TH, MB><<<grid, TH, SM, st>>>(tm4, tm3, a, P.nxy, P.bxy, P.wq, P.ws2, G, kpt);                  \
  }
  if (WARPS == 8) IPB_BOX_LAUNCH(256, 3)
  else if (minb_env >= 6) IPB_BOX_LAUNCH(128, 6)
  else if (minb_env == 5) IPB_BOX_LAUNCH(128, 5)
  else IPB_BOX_LAUNCH(128, 4)
#undef IPB_BOX_LAUNCH
  g_launches++;
  return true;
}

template <class R> bool launch_bilinear_box(const Plan<R>& P, const FieldArgs<R>& a, cudaStream_t st) {
  // fields staged per pass: 8 (measured at km = 1024: 4 -> 2.18 ms, 8 -> 2.12 ms, 16 -> 2.24 ms; profiles/README.md)
  return launch_bilinear_box_f<R, 8>(P, a, st);
}

}  // namespace ipb
