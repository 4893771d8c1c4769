// ipb200 — the TMA-staged bilinear apply kernel (ipolates ip=0, the BASELINE headline path; round 2).
//
// What the reference does per (output point, field) — src/bilinear_interp_mod.F90:202-261 — is a 4-tap weighted sum, a
// divide and a 5-byte store.  With an output grid that oversamples the source (3 km Lambert from 0.25 degree: 66 output
// points per touched source point) the loop is a write-bound stream, 5.09 algorithmic bytes per point-field.  The
// round-1 kernel (ipb_tiles.cuh) gathered each tile's source cells with 4-byte cp.async (LDGSTS) and kept four points
// per lane; ncu showed it bound by the load/store unit (L1TEX 81 %) and by issue slots at 16 resident warps per SM.
// This kernel changes the data movement:
//
//  * a warp owns a PATCH of the output grid: NR consecutive output rows x 32 consecutive points, one point per lane
//    and row.  The source footprint of such a patch is a small rectangle of the (regular, row-major) source grid, so
//    per pass of F fields the rectangle is fetched with ONE `cp.async.bulk.tensor` box per field (TMA, 3-D tensor map
//    x / y / field over the caller's gi) into the warp's shared-memory slice, completion signalled on an mbarrier,
//    double buffered.  No per-lane staging instructions, no address lists in the plan: a point only knows the
//    (ix, iy) of its cell, packed in one word, and the tables do not depend on the store alignment;
//  * the four taps of a point are four shared-memory loads at register + immediate addresses
//    (base, +1, +pitch, +pitch+1), the weighted sum is the reference's 4 multiplies and 3 adds in its order, the
//    quotient G/W comes from the stored correctly rounded 1/W by two fused residual steps (Markstein: the correctly
//    rounded quotient, i.e. the IEEE division's bits; sums outside the exact-residual window take the IEEE division);
//  * go is stored 32 lanes wide — one full 128-byte line per warp instruction: the 32-point run of every row is
//    shifted per field class so that its first element sits on a 128-byte boundary of the caller's go array
//    (field k starts at element k*mo, rows at j*rowlen: the shift is (phase_k + j*rowlen) mod 32 elements).  Fields
//    with the same phase form a class; a block works on one class.  lo goes out as one 32-bit word per 4 points;
//  * blockIdx.x runs over (class, field slice), blockIdx.y over patches, so blocks that are resident together share
//    a patch's plan data through L2 while streaming different fields.
//
// Patches whose cells do not fit the box (BX x BY source points), that contain an irregular stencil (grid edge,
// x-wrap, pole fold), a point failing the mask threshold, and fields that carry a bitmap are computed by the general
// per-point code (global gathers) inside the same kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>

#include "ipb_tiles.cuh"

namespace ipb {

constexpr int BOX_WARPS = 4;                 // independent warp-patches per block
constexpr int BOX_THREADS = 32 * BOX_WARPS;
constexpr int BOX_BX = 16, BOX_BY = 4;       // source points per staged rectangle (x, y)

struct BoxGeom {
  int rowlen, nrow, ntx, nty;   // output row length, rows, patches per row / column
  int amask;                    // alignment unit of the 32-point runs in elements, minus one (31: 128-byte lines in the _4 kind)
  int period;                   // fields k and k + period share their phase (number of classes)
  int base_e;                   // element phase of the caller's go pointer
  int im_src;
};

// ---- plan time: the cell of every output point -----------------------------------------------------------------------
// bxy[n] = iy << 16 | ix (0-based source column / row of tap (1,1)) when the four taps are n, n+1, n+im, n+im+1;
// -1 otherwise (missing tap, x-wrap, pole fold, column-major source).
static __global__ void k_box_index(int no, int im, const int* __restrict__ nxy, int* __restrict__ bxy) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= no) return;
  const int t0 = nxy[n], t1 = nxy[(size_t)no + n], t2 = nxy[2 * (size_t)no + n], t3 = nxy[3 * (size_t)no + n];
  int v = -1;
  if (t0 > 0 && t1 == t0 + 1 && t2 == t0 + im && t3 == t2 + 1) {
    const int ix = (t0 - 1) % im, iy = (t0 - 1) / im;
    if (ix + 1 < im && iy < 32767) v = (iy << 16) | ix;
  }
  bxy[n] = v;
}

template <class R> void build_box_tables(Plan<R>& P, cudaStream_t st) {
  P.box_ok = false;
  if (P.method != M_BILINEAR || P.vec || P.no <= 0 || !P.nxy || !P.wq) return;
  const GridP<R>& gi = P.gin.p;
  if (gi.im <= 1 || gi.im > 65535 || gi.jm <= 1 || gi.jm > 32767) return;
  if ((long long)gi.im * gi.jm > P.mi) return;
  if (((size_t)gi.im * sizeof(R)) % 16 || ((size_t)P.mi * sizeof(R)) % 16) return;   // tensor-map strides are multiples of 16 bytes
  P.bxy = dalloc<int>(P.no);
  k_box_index<<<nblk(P.no, 256), 256, 0, st>>>(P.no, gi.im, P.nxy, P.bxy);
  g_launches++;
  P.bytes += (size_t)P.no * 4;
  P.box_ok = true;
}

// ---- device helpers: mbarrier + TMA ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// one BX x BY x 1 box of the tensor (x, y, field) -> shared memory, completion bytes on `bar`
__device__ __forceinline__ void tma_box_3d(unsigned dst, const CUtensorMap* tm, int x, int y, int k, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(tm), "r"(x), "r"(y), "r"(k), "r"(bar)
               : "memory");
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

#ifndef IPB_BOX_MINBLOCKS
#define IPB_BOX_MINBLOCKS 4
#endif

// NR = output rows per patch, F = fields staged per pass.
template <class R, int NR, int F>
__global__ void __launch_bounds__(BOX_THREADS, IPB_BOX_MINBLOCKS)
k_bilinear_box(const __grid_constant__ CUtensorMap tmap, FieldArgs<R> a, const int* __restrict__ nxy, const R* __restrict__ wxy,
               const int* __restrict__ bxy, const R* __restrict__ wq, const R* __restrict__ ws2, BoxGeom G, int kpt) {
  constexpr int BOXB = BOX_BX * BOX_BY * (int)sizeof(R);   // bytes of one staged rectangle (256 / 512: a multiple of 128)
  constexpr int PITCH = BOX_BX * (int)sizeof(R);
  extern __shared__ __align__(128) unsigned char sm_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned sm_base = (unsigned)__cvta_generic_to_shared(sm_raw);
  const unsigned sbuf = sm_base + warp * (2 * F * BOXB);                       // two buffers of F rectangles
  const unsigned sbar = sm_base + BOX_WARPS * (2 * F * BOXB) + warp * 16;      // two mbarriers
  const int cls = blockIdx.x % G.period, slice = blockIdx.x / G.period;
  const int tile = blockIdx.y * BOX_WARPS + warp;
  if (tile >= G.ntx * G.nty) return;
  const int tx = tile % G.ntx, ty = tile / G.ntx;
  // fields of this warp: k = cls + period * j, j in [jbeg, jend)
  const int jbeg = kpt * slice, jend = min(kpt * (slice + 1), (a.kc - cls + G.period - 1) / G.period);
  if (jbeg >= jend) return;
  const int phase = (int)(((long long)G.base_e + (long long)cls * a.mo) & G.amask);

  // ---- the patch: one point per lane and row ----
  int npt[NR];          // output point of (row r, lane), -1 outside
  unsigned soff[NR];    // byte offset of its cell inside a staged rectangle
  R w[NR][4], W[NR], rW[NR];
  int minx = 1 << 30, miny = 1 << 30, maxx = -1, maxy = -1;
  bool regular = true, full = true, whole = true;   // whole: every lane of every row has a point (no row edge inside the patch)
  int cxy[NR];
#pragma unroll
  for (int r = 0; r < NR; r++) {
    const int j = ty * NR + r;
    const int shift = (int)(((long long)phase + (long long)j * G.rowlen) & G.amask);
    const int i = tx * 32 - shift + lane;
    const long long n = (long long)j * G.rowlen + i;
    const bool in = j < G.nrow && i >= 0 && i < G.rowlen && n < a.no;
    npt[r] = in ? (int)n : -1;
    whole = whole && in;
    const size_t nn = in ? (size_t)n : 0;
    const int c = in ? __ldg(bxy + nn) : 0;
    cxy[r] = c;
    if (in) {
      if (c < 0) regular = false;
      else { minx = min(minx, c & 0xffff); maxx = max(maxx, c & 0xffff); miny = min(miny, c >> 16); maxy = max(maxy, c >> 16); }
    }
    if (sizeof(R) == 4) {
      const float4 q = in ? __ldg(reinterpret_cast<const float4*>(wq) + nn) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float2 s2 = in ? __ldg(reinterpret_cast<const float2*>(ws2) + nn) : make_float2(1.f, 1.f);
      w[r][0] = (R)q.x; w[r][1] = (R)q.y; w[r][2] = (R)q.z; w[r][3] = (R)q.w; W[r] = (R)s2.x; rW[r] = (R)s2.y;
    } else {
      const double2 z = make_double2(0., 0.), one = make_double2(1., 1.);
      const double2 q0 = in ? __ldg(reinterpret_cast<const double2*>(wq) + 2 * nn) : z, q1 = in ? __ldg(reinterpret_cast<const double2*>(wq) + 2 * nn + 1) : z;
      const double2 s2 = in ? __ldg(reinterpret_cast<const double2*>(ws2) + nn) : one;
      w[r][0] = (R)q0.x; w[r][1] = (R)q0.y; w[r][2] = (R)q1.x; w[r][3] = (R)q1.y; W[r] = (R)s2.x; rW[r] = (R)s2.y;
    }
    const R dev = W[r] - (R)1;
    if (in) full = full && W[r] >= a.pmp && dev <= (R)9.5e-7 && dev >= (R)-9.5e-7;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    minx = min(minx, __shfl_xor_sync(0xffffffffu, minx, o)); miny = min(miny, __shfl_xor_sync(0xffffffffu, miny, o));
    maxx = max(maxx, __shfl_xor_sync(0xffffffffu, maxx, o)); maxy = max(maxy, __shfl_xor_sync(0xffffffffu, maxy, o));
  }
  const bool any_pt = maxx >= 0 || !__all_sync(0xffffffffu, regular);
  if (!any_pt && !__any_sync(0xffffffffu, npt[0] >= 0 || npt[NR - 1] >= 0)) {
    bool some = false;
#pragma unroll
    for (int r = 0; r < NR; r++) some = some || npt[r] >= 0;
    if (!__any_sync(0xffffffffu, some)) return;   // patch entirely outside the grid
  }
  // staged: every stencil regular and complete, every point passing the threshold with W ~ 1, the cells inside one box
  const bool staged = __all_sync(0xffffffffu, regular && full) && maxx >= 0 && (maxx + 1 - minx) < BOX_BX && (maxy + 1 - miny) < BOX_BY;
  const bool whole_w = __all_sync(0xffffffffu, whole);
#pragma unroll
  for (int r = 0; r < NR; r++) {
    const int c = cxy[r];
    soff[r] = (npt[r] >= 0 && staged) ? (unsigned)((((c >> 16) - miny) * BOX_BX + ((c & 0xffff) - minx)) * (int)sizeof(R)) : 0u;
  }
  // bit b set: field j = jm0 + b of this warp carries a bitmap; refreshed every 32 fields
  unsigned maskbits = __ballot_sync(0xffffffffu, (jbeg + lane < jend) && a.ibi[cls + G.period * (jbeg + lane)] != 0);
  int jm0 = jbeg;
  const size_t ostep = (size_t)G.period * a.mo;
  // lo words: lane 8r + q stores the word of points 4q .. 4q+3 of row r (whole patches only)
  long long lo_off = -1;
  if (NR * 8 >= 32 || lane < NR * 8) {
    const int r = lane >> 3, q = lane & 7;
    const int first = __shfl_sync(0xffffffffu, npt[0], 0);   // placeholder to keep the shuffle uniform; replaced below
    (void)first;
    lo_off = 0;
    (void)r; (void)q;
  }
  // point index of lane 0 of every row (uniform), for the lo words
  int row0[NR];
#pragma unroll
  for (int r = 0; r < NR; r++) row0[r] = __shfl_sync(0xffffffffu, npt[r], 0);
  {
    const int r = lane >> 3, q = lane & 7;
    int base = -1;
#pragma unroll
    for (int rr = 0; rr < NR; rr++) if (rr == r) base = row0[rr];
    lo_off = (r < NR && base >= 0) ? (long long)base + 4 * q : -1;
  }

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
    for (int j = jbeg; j < jend; j++) general_field(cls + G.period * j);
    return;
  }

  // ---- staged path ----
  if (lane == 0) { mbar_init(sbar, 1); mbar_init(sbar + 8, 1); }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  auto issue = [&](int buf, int j0) {
    const int nf = min(F, jend - j0);
    if (lane == 0) mbar_expect_tx(sbar + 8 * buf, (unsigned)(nf * BOXB));
    __syncwarp();
    if (lane < nf) tma_box_3d(sbuf + (unsigned)((buf * F + lane) * BOXB), &tmap, minx, miny, cls + G.period * (j0 + lane), sbar + 8 * buf);
  };
  issue(0, jbeg);
  R* gop = a.go + (size_t)(cls + G.period * jbeg) * a.mo;    // field base (uniform); lanes add their point
  u8* lop = a.lo + (size_t)(cls + G.period * jbeg) * a.mo;
  unsigned par = 0;   // bit b: parity to wait for on buffer b
  int buf = 0;
  for (int j0 = jbeg; j0 < jend; j0 += F) {
    if (j0 + F < jend) issue(buf ^ 1, j0 + F);
    mbar_wait(sbar + 8 * buf, (par >> buf) & 1u);
    par ^= 1u << buf;
    if (j0 - jm0 >= 32) {
      jm0 = j0;
      maskbits = __ballot_sync(0xffffffffu, (j0 + lane < jend) && a.ibi[cls + G.period * (j0 + lane)] != 0);
    }
    const unsigned passmask = (maskbits >> (j0 - jm0)) & ((1u << F) - 1u);
    const unsigned sb = sbuf + (unsigned)(buf * F * BOXB);
    if (j0 + F <= jend && passmask == 0 && whole_w) {
      // whole pass, no bitmaps, no row edge: straight-line code
#pragma unroll
      for (int f = 0; f < F; f++) {
        R v[NR][4], Gs[NR], o[NR];
#pragma unroll
        for (int r = 0; r < NR; r++) {
          const unsigned s = sb + soff[r];
          v[r][0] = lds_r<R>(s + f * BOXB);
          v[r][1] = lds_r<R>(s + f * BOXB + (int)sizeof(R));
          v[r][2] = lds_r<R>(s + f * BOXB + PITCH);
          v[r][3] = lds_r<R>(s + f * BOXB + PITCH + (int)sizeof(R));
        }
        BoxRows<R, NR>::run(v, w, W, rW, Gs, o);
#pragma unroll
        for (int r = 0; r < NR; r++) {
          const R ag = abs_r<R>(Gs[r]);
          // rare: a sum that is zero, tiny, huge or not finite is outside the exact-residual window: IEEE division
          if (!(ag < win_hi<R>() && ag >= win_lo<R>())) o[r] = ((R)0 + Gs[r]) / W[r];
          __stcs(gop + (size_t)f * ostep + npt[r], o[r]);
        }
        if (lo_off >= 0) __stcs(reinterpret_cast<unsigned*>(lop + (size_t)f * ostep + lo_off), 0x01010101u);
      }
    } else {
      for (int f = 0; f < F && j0 + f < jend; f++) {
        const int k = cls + G.period * (j0 + f);
        if ((passmask >> f) & 1u) { general_field(k); continue; }
#pragma unroll
        for (int r = 0; r < NR; r++) {
          const unsigned s = sb + soff[r] + (unsigned)(f * BOXB);
          R g = w[r][0] * lds_r<R>(s);
          g = g + w[r][1] * lds_r<R>(s + (int)sizeof(R));
          g = g + w[r][2] * lds_r<R>(s + PITCH);
          g = g + w[r][3] * lds_r<R>(s + PITCH + (int)sizeof(R));
          const R q = ((R)0 + g) / W[r];
          if (npt[r] >= 0) {
            __stcs(gop + (size_t)f * ostep + npt[r], q);
            __stcs(lop + (size_t)f * ostep + npt[r], (u8)1);
          }
        }
      }
    }
    gop += (size_t)F * ostep; lop += (size_t)F * ostep;
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

typedef CUresult (*PFN_tmap_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_tmap_encode tmap_encoder() {
  static PFN_tmap_encode fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (PFN_tmap_encode)p;
  }();
  return fn;
}

template <class R> bool launch_bilinear_box(const Plan<R>& P, const FieldArgs<R>& a, cudaStream_t st) {
  PFN_tmap_encode enc = tmap_encoder();
  if (!enc) return false;
  const GridP<R>& gi = P.gin.p;
  CUtensorMap tm;
  const cuuint64_t dims[3] = {(cuuint64_t)gi.im, (cuuint64_t)gi.jm, (cuuint64_t)a.kc};
  const cuuint64_t strides[2] = {(cuuint64_t)gi.im * sizeof(R), (cuuint64_t)a.mi * sizeof(R)};
  const cuuint32_t box[3] = {BOX_BX, BOX_BY, 1}, es[3] = {1, 1, 1};
  const CUresult rc = enc(&tm, sizeof(R) == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<R*>(a.gi), dims, strides, box,
                          es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) return false;
  static const int align_env = getenv("IPB_BOX_ALIGN") ? atoi(getenv("IPB_BOX_ALIGN")) : 0;   // alignment unit of the store runs, bytes
  static const int kpt_env = getenv("IPB_BOX_KPT") ? atoi(getenv("IPB_BOX_KPT")) : 32;
  static const int nr_env = getenv("IPB_BOX_NR") ? atoi(getenv("IPB_BOX_NR")) : 4;
  int unit = (align_env >= 16 && align_env <= 128 ? align_env : 128) / (int)sizeof(R);       // elements; a power of two >= 4 (lo words)
  if (unit < 4) unit = 4;
  BoxGeom G;
  G.rowlen = patch_rowlen(P.gout.p, P.no);
  G.nrow = nblk(P.no, G.rowlen);
  G.amask = unit - 1;
  G.ntx = (G.rowlen + G.amask + 31) / 32;
  const int NR = nr_env == 8 ? 8 : nr_env == 2 ? 2 : 4;
  G.nty = nblk(G.nrow, NR);
  int s = (int)(a.mo & G.amask), g = unit;   // period = unit / gcd(mo mod unit, unit)
  while (s) { const int t = g % s; g = s; s = t; }
  G.period = unit / g;
  G.base_e = (int)((reinterpret_cast<uintptr_t>(a.go) / sizeof(R)) & G.amask);
  G.im_src = gi.im;
  const int per_class = (a.kc + G.period - 1) / G.period;
  const int kpt = std::max(8, std::min(kpt_env, ((per_class + 7) / 8) * 8));   // fields of one class per warp
  dim3 grid(G.period * nblk(per_class, kpt), nblk(G.ntx * G.nty, BOX_WARPS));
#define IPB_BOX_LAUNCH(NRR, FF)                                                                                              \
  {                                                                                                                          \
    constexpr int SM = BOX_WARPS * (2 * (FF) * BOX_BX * BOX_BY * (int)sizeof(R) + 16);                                       \
    static bool once[64] = {};                                                                                               \
    int dv_ = 0; cudaGetDevice(&dv_); dv_ &= 63;                                                                             \
    if (!once[dv_]) { IPB_CUDA(cudaFuncSetAttribute(k_bilinear_box<R, NRR, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM)); once[dv_] = true; } \
    k_bilinear_box<R, NRR, FF><<<grid, BOX_THREADS, SM, st>>>(tm, a, P.nxy, P.wxy, P.bxy, P.wq, P.ws2, G, kpt);              \
  }
  if (NR == 8) IPB_BOX_LAUNCH(8, 4)
  else if (NR == 2) IPB_BOX_LAUNCH(2, 8)
  else IPB_BOX_LAUNCH(4, 8)
#undef IPB_BOX_LAUNCH
  g_launches++;
  return true;
}

}  // namespace ipb
