// ipb200 — the tile-staged bilinear apply kernel (ipolates ip=0; BASELINE's headline path) and the
// tile tables it needs in the plan.
//
// What the reference does per (output point, field) — src/bilinear_interp_mod.F90:202-261 — is a
// 4-tap masked weighted sum, a divide and a 5-byte store (go + lo).  When the output grid oversamples
// the source (3 km Lambert from 0.25 degree: ~66 output points per touched source point) the loop
// is a write-bound stream, 5.09 algorithmic bytes per point-field, and what limits a GPU is not HBM
// but *instruction issue*: four global gathers with 64-bit address arithmetic, eight unfused
// multiply/adds and an IEEE division per 5 bytes stored.  This kernel removes the gathers from the
// inner loop:
//
//  * the output points are cut into TILES of 512 consecutive points (128 threads x 4 points).  At
//    plan time the distinct source CELLS of a tile — a cell is the 4-tuple of taps
//    (n11,n21,n12,n22) of an output point — are listed once (~85 per tile for the headline pair),
//    and every output point stores the one-byte number of its cell;
//  * per batch of F fields the block copies the tile's cells from the fields into shared memory
//    with 4-byte cp.async (the only global reads of field data), laid out as one 16-byte (fp32) /
//    32-byte (fp64) record per cell and field;
//  * the inner loop is then ONE shared-memory vector load per point-field (address = per-point
//    register + compile-time field offset), the reference's 4 multiplies and 4 adds in its order,
//    the division, and one 16-byte go store + one 4-byte lo store per 4 points and field, with
//    evict-first hints so the 10 GB output stream does not push the plan out of L2;
//  * field k of go starts at element k*mo, which is only vector-aligned for some k.  Fields are
//    split into four alignment classes (k mod 4); a block works on one class, and the
//    thread -> point mapping is shifted per class ("variant" a = 0..3, group g covers points
//    4g-a .. 4g-a+3) so that every vector store is aligned.  The tile tables exist per variant;
//  * the division G/W is formed from the stored correctly rounded 1/W by two fused residual
//    corrections (Markstein), which gives the correctly rounded quotient — the same bits as the
//    IEEE division the reference executes — at 4 FMAs instead of a ~10-instruction divide;
//  * blockIdx.x runs over (class, field slice), blockIdx.y over tiles, so blocks that are
//    resident together share a tile's plan data through L2 while streaming different fields.
//
// Points with a missing tap, a mask threshold failure, fields with a bitmap, and tiles whose cell
// count exceeds the table fall back — per thread and field — to the general code path
// (bilinear_point, global gathers), which is also what the other kernels use.
#pragma once
#include <cstdlib>

#include "ipb_fast.cuh"

namespace ipb {

constexpr int TILE_PTS = 128;      // output points per tile: one warp x 4 points
constexpr int TILE_WARPS = 4;      // independent warp-tiles per block
constexpr int TILE_THREADS = 32 * TILE_WARPS;

// ---- plan-time: cells of every tile, for one alignment variant ------------------------------------
// pass 1: per point the number of its cell inside the tile (first-appearance order; 0xff = not a
// complete 4-tap stencil), per tile the cell count.  One warp per tile.
static __global__ void __launch_bounds__(TILE_THREADS)
k_tile_index(int variant, int no, int ntile, const int* __restrict__ nxy, u8* __restrict__ cid, int* __restrict__ ncell, int* gmax) {
  __shared__ int4 tup_all[TILE_WARPS][TILE_PTS];
  __shared__ int rank_all[TILE_WARPS][TILE_PTS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * TILE_WARPS + warp;
  if (tile >= ntile) return;
  int4* tup = tup_all[warp];
  int* rank = rank_all[warp];
  const int nbase = 4 * (tile * 32 + lane) - variant;
  int4 mine[4];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int n = nbase + p;
    int4 t = make_int4(0, 0, 0, 0);
    if (n >= 0 && n < no) {
      t = make_int4(nxy[n], nxy[(size_t)no + n], nxy[2 * (size_t)no + n], nxy[3 * (size_t)no + n]);
      if (t.x <= 0 || t.y <= 0 || t.z <= 0 || t.w <= 0) t = make_int4(0, 0, 0, 0);
    }
    mine[p] = t;
    tup[4 * lane + p] = t;
  }
  __syncwarp();
  int first[4], nrep = 0;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int q = 4 * lane + p;
    first[p] = -1;
    if (mine[p].x > 0) {
      int r = 0;
      for (; r < q; r++) {
        const int4 o = tup[r];
        if (o.x == mine[p].x && o.y == mine[p].y && o.z == mine[p].z && o.w == mine[p].w) break;
      }
      first[p] = r;
      if (r == q) nrep++;
    }
  }
  int incl = nrep;   // inclusive prefix of representative counts over lanes
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
  int run = incl - nrep;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int q = 4 * lane + p;
    rank[q] = -1;
    if (first[p] == q) rank[q] = run++;
  }
  __syncwarp();
  unsigned packed = 0;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    int c = 0xff;
    if (first[p] >= 0) { const int r = rank[first[p]]; c = r < 0xff ? r : 0xff; }
    packed |= (unsigned)c << (8 * p);
  }
  reinterpret_cast<unsigned*>(cid)[(size_t)tile * 32 + lane] = packed;
  if (lane == 31) { ncell[tile] = incl; atomicMax(gmax, incl); }
}
// pass 2: the staging list of every tile.  Entry e = cell*4 + tap is one 4-byte value to copy per field;
// entries are sorted by source position so that the 32 copies of one warp instruction fall into a
// few 32-byte sectors (the copy engine works sector by sector).  esrc = 0-based source position
// (-1 = padding), edst = shared-memory slot e of the entry.
static __global__ void __launch_bounds__(TILE_THREADS)
k_tile_entries(int variant, int no, int ntile, int estride, const int* __restrict__ nxy, const u8* __restrict__ cid,
               const int* __restrict__ ncell, int* __restrict__ esrc, u8* __restrict__ edst) {
  __shared__ int key_all[TILE_WARPS][4 * 64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x * TILE_WARPS + warp;
  if (tile >= ntile) return;
  int* key = key_all[warp];
  const int nbase = 4 * (tile * 32 + lane) - variant;
  const unsigned packed = reinterpret_cast<const unsigned*>(cid)[(size_t)tile * 32 + lane];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int n = nbase + p;
    const int c = (packed >> (8 * p)) & 0xff;
    if (c == 0xff || n < 0 || n >= no) continue;
#pragma unroll
    for (int t = 0; t < 4; t++) key[c * 4 + t] = nxy[(size_t)t * no + n] - 1;   // every point of a cell writes the same values
  }
  __syncwarp();
  const int ne = 4 * ncell[tile];
  for (int e = lane; e < estride; e += 32) {
    if (e >= ne) { esrc[(size_t)tile * estride + e] = -1; edst[(size_t)tile * estride + e] = 0; }
  }
  for (int e = lane; e < ne; e += 32) {
    const int k = key[e];
    int rank = 0;
    for (int o = 0; o < ne; o++) { const int ko = key[o]; rank += (ko < k || (ko == k && o < e)) ? 1 : 0; }
    esrc[(size_t)tile * estride + rank] = k;
    edst[(size_t)tile * estride + rank] = (u8)e;
  }
}

// Per-point weight record of the staged kernel: the four tap weights contiguous (one 16/32-byte load per
// point instead of four scattered ones), and the weight sum with its correctly rounded reciprocal.
template <class R>
__global__ void k_tile_weights(int no, const R* __restrict__ wxy, R* __restrict__ wq, R* __restrict__ ws2) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= no) return;
  R s = 0;
#pragma unroll
  for (int t = 0; t < 4; t++) { const R w = wxy[(size_t)t * no + n]; wq[(size_t)n * 4 + t] = w; s = s + w; }   // reference order (1,1),(2,1),(1,2),(2,2)
  ws2[(size_t)n * 2] = s;
  ws2[(size_t)n * 2 + 1] = (R)1 / s;
}

constexpr int TILE_MAX_CELLS = 64;   // staging covers 8*NI cells per tile, NI <= 8

template <class R> void build_tile_tables(Plan<R>& P, cudaStream_t st) {
  TileTables& T = P.tiles;
  T = TileTables();
  if (P.method != M_BILINEAR || P.vec || P.no <= 0 || !P.nxy) return;
  const int no = P.no;
  const int ng = (no + 3) / 4 + 1;                 // groups of 4 points, any variant
  const int ntile = nblk(ng, 32);
  const int nb = nblk(ntile, TILE_WARPS);
  int* d_max = dalloc<int>(1);
  IPB_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), st));
  for (int v = 0; v < 4; v++) {
    T.cid[v] = dalloc<u8>((size_t)ntile * TILE_PTS);
    T.ncell[v] = dalloc<int>(ntile);
    k_tile_index<<<nb, TILE_THREADS, 0, st>>>(v, no, ntile, P.nxy, T.cid[v], T.ncell[v], d_max);
    g_launches++;
  }
  int hmax = 0;
  IPB_CUDA(cudaMemcpyAsync(&hmax, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
  IPB_CUDA(cudaStreamSynchronize(st));
  dfree(d_max);
  T.ntile = ntile; T.max_cells = hmax;
  if (hmax <= 0 || hmax > TILE_MAX_CELLS) {        // too little reuse for staging to pay: general kernel
    for (int v = 0; v < 4; v++) { dfree(T.cid[v]); dfree(T.ncell[v]); }
    return;
  }
  T.cstride = ((4 * hmax + 31) / 32) * 32;          // staging entries per tile, padded to whole warp instructions
  for (int v = 0; v < 4; v++) {
    T.cells[v] = dalloc<int>((size_t)ntile * T.cstride);
    T.edst[v] = dalloc<u8>((size_t)ntile * T.cstride);
    k_tile_entries<<<nb, TILE_THREADS, 0, st>>>(v, no, ntile, T.cstride, P.nxy, T.cid[v], T.ncell[v], T.cells[v], T.edst[v]);
    g_launches++;
  }
  P.wq = dalloc<R>((size_t)no * 4);
  P.ws2 = dalloc<R>((size_t)no * 2);
  k_tile_weights<R><<<nblk(no, 256), 256, 0, st>>>(no, P.wxy, P.wq, P.ws2);
  g_launches++;
  T.bytes = 4 * ((size_t)ntile * TILE_PTS + (size_t)ntile * 4 + (size_t)ntile * T.cstride * 5) + (size_t)no * 6 * sizeof(R);
  T.ok = true;
}

// ---- apply ---------------------------------------------------------------------------------------------
template <class R> struct CellRec;   // the four taps of one cell for one field
template <> struct CellRec<float> { float4 v; __device__ __forceinline__ float get(int t) const { return t == 0 ? v.x : t == 1 ? v.y : t == 2 ? v.z : v.w; } };
template <> struct CellRec<double> { double2 a, b; __device__ __forceinline__ double get(int t) const { return t == 0 ? a.x : t == 1 ? a.y : t == 2 ? b.x : b.y; } };

// predicated 4- / 8-byte asynchronous copy global -> shared (straight-line code, no divergent branch)
template <int BYTES> __device__ __forceinline__ void cp_async_if(bool on, unsigned smem_addr, const void* gmem) {
  if (BYTES == 4)
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p cp.async.ca.shared.global [%0], [%1], 4;\n}\n" ::"r"(smem_addr), "l"(gmem), "r"((int)on));
  else
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %2, 0;\n @p cp.async.ca.shared.global [%0], [%1], 8;\n}\n" ::"r"(smem_addr), "l"(gmem), "r"((int)on));
}

// Correctly rounded g / w for |w - 1| <= 2^-20, given rw = RN(1/w): q0 = g is within a few ulp,
// one fused correction makes it faithful, the second correctly rounded (Markstein).  Valid while the
// residuals are exact, i.e. g = 0 or |g| inside in_window(); g = -0 gives +0, which is what the
// reference's sum (it starts from +0) produces.
template <class R> __device__ __forceinline__ R quot_near1(R g, R w, R rw) {
  R r = fma_r<R>(-g, w, g);
  const R q1 = fma_r<R>(r, rw, g);
  r = fma_r<R>(-q1, w, g);
  return fma_r<R>(r, rw, q1);
}
// ---- packed binary32 arithmetic (sm_100 FMUL2 / FFMA2: two IEEE operations per issued instruction) ----------
// Used where the reference's operation is exactly a pair of independent multiplies or a pair of independent
// fused steps of the quotient.  Multiplies are never followed by a packed add (ptxas 12.9 fuses a packed
// mul.rn + add.rn pair into FFMA2 even with --fmad false, which would change the rounding): the adds of the
// tap sum stay scalar.
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*reinterpret_cast<unsigned long long*>(&r)) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*reinterpret_cast<unsigned long long*>(&r)) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)), "l"(*reinterpret_cast<unsigned long long*>(&c)));
  return r;
}
// the four tap sums and quotients of one thread and field: generic form
template <class R> struct Quad {
  static __device__ __forceinline__ void run(const CellRec<R> (&rec)[4], const R (&w)[4][4], const R (&W)[4], const R (&rW)[4], R (&G)[4], R (&o)[4]) {
#pragma unroll
    for (int p = 0; p < 4; p++) {
      // the reference starts from G=0; 0 + x only changes x = -0, which the quotient maps to +0 again
      R g = w[p][0] * rec[p].get(0);
#pragma unroll
      for (int t = 1; t < 4; t++) g = g + w[p][t] * rec[p].get(t);
      G[p] = g;
      o[p] = quot_near1<R>(g, W[p], rW[p]);
    }
  }
};
// binary32: the same operations with the products and the quotient steps issued two at a time
template <> struct Quad<float> {
  static __device__ __forceinline__ void run(const CellRec<float> (&rec)[4], const float (&w)[4][4], const float (&W)[4], const float (&rW)[4],
                                             float (&G)[4], float (&o)[4]) {
#pragma unroll
    for (int p = 0; p < 4; p++) {
      const float2 pa = mul2(make_float2(w[p][0], w[p][1]), make_float2(rec[p].v.x, rec[p].v.y));
      const float2 pb = mul2(make_float2(w[p][2], w[p][3]), make_float2(rec[p].v.z, rec[p].v.w));
      G[p] = ((pa.x + pa.y) + pb.x) + pb.y;
    }
#pragma unroll
    for (int h = 0; h < 2; h++) {
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

template <class R> __device__ __forceinline__ R win_lo() { return sizeof(R) == 4 ? (R)6.5e-27f : (R)1e-187; }   // ~2^-87 / 2^-620
template <class R> __device__ __forceinline__ R win_hi() { return sizeof(R) == 4 ? (R)1.5e26f : (R)1e187; }
template <class R> __device__ __forceinline__ R abs_r(R x);
template <> __device__ __forceinline__ float abs_r<float>(float x) { return fabsf(x); }
template <> __device__ __forceinline__ double abs_r<double>(double x) { return fabs(x); }
template <class R> __device__ __forceinline__ R max_r(R a, R b);
template <> __device__ __forceinline__ float max_r<float>(float a, float b) { return fmaxf(a, b); }
template <> __device__ __forceinline__ double max_r<double>(double a, double b) { return fmax(a, b); }
template <class R> __device__ __forceinline__ R min_r(R a, R b);
template <> __device__ __forceinline__ float min_r<float>(float a, float b) { return fminf(a, b); }
template <> __device__ __forceinline__ double min_r<double>(double a, double b) { return fmin(a, b); }

#ifndef IPB_TILES_NBUF
#define IPB_TILES_NBUF 2
#endif
#ifndef IPB_TILES_MINBLOCKS
#define IPB_TILES_MINBLOCKS (sizeof(R) == 4 ? 4 : 3)
#endif
// F = fields staged per pass, NI = staging copies per lane and field (covers 8*NI cells).
// Each warp is an independent tile: it stages its own cells into its own shared-memory slice
// (double buffered: the copies of pass s+1 are in flight while pass s is computed) and never
// meets a block-wide barrier.
// MASKED = the batch contains fields with a bitmap: their li bytes are staged next to the values (one packed
// word per cell) and those fields are computed from shared memory too, tap by tap with the IEEE division,
// instead of falling back to global gathers.
template <class R, int F, int NI, bool MASKED>
__global__ void __launch_bounds__(TILE_THREADS, IPB_TILES_MINBLOCKS)
k_bilinear_tiles(FieldArgs<R> a, const int* __restrict__ nxy, const R* __restrict__ wq, const R* __restrict__ ws2, TileDev T, int base_mis, int kpt) {
  constexpr int CMAX = 8 * NI;                      // cells a tile may have
  constexpr int BUF = F * CMAX * 4;                 // elements per buffer
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NBUF = IPB_TILES_NBUF;               // staging depth: passes in flight ahead of the one being computed + 1
  R* sm = reinterpret_cast<R*>(sm_raw) + warp * NBUF * BUF;
  u8* smask = sm_raw + (size_t)TILE_WARPS * NBUF * BUF * sizeof(R) + (size_t)warp * F * CMAX * 4;   // [F][cell*4 + tap] (MASKED only)
  const int cls = blockIdx.x & 3, slice = blockIdx.x >> 2, tile = blockIdx.y * TILE_WARPS + warp;
  if (tile >= T.ntile) return;
  const int variant = (base_mis + cls * (a.mo & 3)) & 3;
  const int n0 = 4 * (tile * 32 + lane) - variant;
  // fields of this warp: k = cls + 4 j, j in [jbeg, jend)
  const int jbeg = kpt * slice, jend = min(kpt * (slice + 1), (a.kc - cls + 3) / 4);
  if (jbeg >= jend) return;
  // ---- staging assignment: entry q = cell*4 + tap of this tile's cell table ----
  const int* __restrict__ esrc = T.cells[variant] + (size_t)tile * T.cstride;
  const u8* __restrict__ edst = T.edst[variant] + (size_t)tile * T.cstride;
  const size_t fstep = 4 * (size_t)a.mi;            // elements between consecutive fields of one class
  const R* gp[NI];                                  // source address of entry i in the next field to stage
  unsigned sdst[NI];                                // its shared-memory address in buffer 0, field 0
  bool act[NI];
  int srci[MASKED ? NI : 1], slot[MASKED ? NI : 1];   // MASKED: source position and shared-memory slot of each entry, for the bitmap bytes
#pragma unroll
  for (int i = 0; i < NI; i++) {
    const int q = lane + 32 * i;
    const int src = q < T.cstride ? esrc[q] : -1;
    act[i] = src >= 0;
    if (MASKED) { srci[i] = act[i] ? src : 0; slot[i] = act[i] ? (int)edst[q] : 0; }
    gp[i] = a.gi + (size_t)(cls + 4 * jbeg) * a.mi + (act[i] ? src : 0);
    sdst[i] = (unsigned)__cvta_generic_to_shared(sm + (act[i] ? (int)edst[q] : 0));
  }
  auto stage = [&](int buf, int j0) {
    const unsigned boff = (unsigned)(buf * BUF * sizeof(R));
    if (j0 + F <= jend) {          // whole pass: straight-line copies, addresses formed in temporaries
#pragma unroll
      for (int f = 0; f < F; f++) {
#pragma unroll
        for (int i = 0; i < NI; i++)
          cp_async_if<sizeof(R)>(act[i], sdst[i] + boff + (unsigned)(f * CMAX * 4 * sizeof(R)), gp[i] + (size_t)f * fstep);
      }
#pragma unroll
      for (int i = 0; i < NI; i++) gp[i] += (size_t)F * fstep;
    } else {                       // ragged last pass
      for (int f = 0; j0 + f < jend; f++) {
#pragma unroll
        for (int i = 0; i < NI; i++) {
          cp_async_if<sizeof(R)>(act[i], sdst[i] + boff + (unsigned)(f * CMAX * 4 * sizeof(R)), gp[i]);
          gp[i] += fstep;
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  stage(0, jbeg);
  // bit b set: field j = jm0 + b of this warp carries a bitmap; refreshed every 32 fields
  unsigned maskbits = __ballot_sync(0xffffffffu, (jbeg + lane < jend) && a.ibi[cls + 4 * (jbeg + lane)] != 0);
  int jm0 = jbeg;
  // ---- per-thread plan data: weights, weight sums, cell numbers ----
  R w[4][4], W[4], rW[4];
  int coff[4];
  bool full = true, incell = true;   // full: the fast quotient path applies; incell: all four points have a staged cell
  const unsigned packed = reinterpret_cast<const unsigned*>(T.cid[variant])[(size_t)tile * 32 + lane];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int n = n0 + p;
    const bool in = n >= 0 && n < a.no;
    const int c = (packed >> (8 * p)) & 0xff;
    const size_t nn = in ? (size_t)n : 0;
    if (sizeof(R) == 4) {
      const float4 q = __ldg(reinterpret_cast<const float4*>(wq) + nn);
      const float2 s2 = __ldg(reinterpret_cast<const float2*>(ws2) + nn);
      w[p][0] = (R)q.x; w[p][1] = (R)q.y; w[p][2] = (R)q.z; w[p][3] = (R)q.w; W[p] = (R)s2.x; rW[p] = (R)s2.y;
    } else {
      const double2 q0 = __ldg(reinterpret_cast<const double2*>(wq) + 2 * nn), q1 = __ldg(reinterpret_cast<const double2*>(wq) + 2 * nn + 1);
      const double2 s2 = __ldg(reinterpret_cast<const double2*>(ws2) + nn);
      w[p][0] = (R)q0.x; w[p][1] = (R)q0.y; w[p][2] = (R)q1.x; w[p][3] = (R)q1.y; W[p] = (R)s2.x; rW[p] = (R)s2.y;
    }
    const R ws = W[p];
    coff[p] = c * 4;
    const R dev = ws - (R)1;
    full = full && in && c != 0xff && ws >= a.pmp && dev <= (R)9.5e-7 && dev >= (R)-9.5e-7;
    incell = incell && in && c != 0xff;
  }
  const bool same1a = coff[1] == coff[0], same2a = coff[2] == coff[0];
  const bool own1 = !same1a && coff[1] != coff[3], own2 = !same2a && coff[2] != coff[3];
  R* gop = a.go + ((size_t)(cls + 4 * jbeg) * a.mo + n0);   // this thread's 4 points in the next field to write
  u8* lop = a.lo + ((size_t)(cls + 4 * jbeg) * a.mo + n0);
  const size_t ostep = 4 * (size_t)a.mo;
  // prologue: NBUF-1 passes in flight (an empty group is committed where there is no pass, so that the
  // group count below is uniform)
#pragma unroll
  for (int b = 1; b < NBUF - 1; b++) {
    if (jbeg + b * F < jend) stage(b, jbeg + b * F);
    else asm volatile("cp.async.commit_group;\n" ::: "memory");
  }
  int buf = 0, nxt = NBUF - 1;
  for (int j0 = jbeg; j0 < jend; j0 += F) {
    if (j0 + (NBUF - 1) * F < jend) stage(nxt, j0 + (NBUF - 1) * F);
    else asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group %0;\n" ::"n"(NBUF - 1) : "memory");
    __syncwarp();
    const R* __restrict__ sb = sm + buf * BUF;
    if (j0 - jm0 >= 32) {
      jm0 = j0;
      maskbits = __ballot_sync(0xffffffffu, (j0 + lane < jend) && a.ibi[cls + 4 * (j0 + lane)] != 0);
    }
    // ---- compute ----
    const bool plain = full && (j0 + F <= jend) && ((maskbits >> (j0 - jm0)) & ((1u << F) - 1u)) == 0;
    const unsigned passmask = MASKED ? ((maskbits >> (j0 - jm0)) & ((1u << F) - 1u)) : 0u;   // bit f: field f of the pass has a bitmap
    if (MASKED && passmask) {
      // the bitmap bytes of this pass's masked fields, at the positions of the staging list
      for (int f = 0; f < F && j0 + f < jend; f++) {
        if (!((passmask >> f) & 1u)) continue;
        const u8* __restrict__ l = a.li + (size_t)(cls + 4 * (j0 + f)) * a.mi;
#pragma unroll
        for (int i = 0; i < NI; i++)
          if (act[i]) smask[f * CMAX * 4 + slot[i]] = __ldg(l + srci[i]);
      }
      __syncwarp();
    }
    if (plain) {
#pragma unroll
      for (int f = 0; f < F; f++) {
        R G[4], o[4];
        // the 4 consecutive points of a thread usually lie in one or two source cells: the records of
        // the first and the last point are loaded, the middle points reuse them (a third load only
        // where a middle point sits in a cell of its own) — shared-memory bandwidth is what bounds
        // this loop, not instruction issue
        CellRec<R> rec[4];
        rec[0] = *reinterpret_cast<const CellRec<R>*>(&sb[f * CMAX * 4 + coff[0]]);
        rec[3] = *reinterpret_cast<const CellRec<R>*>(&sb[f * CMAX * 4 + coff[3]]);
        rec[1] = same1a ? rec[0] : rec[3];
        rec[2] = same2a ? rec[0] : rec[3];
        if (own1) rec[1] = *reinterpret_cast<const CellRec<R>*>(&sb[f * CMAX * 4 + coff[1]]);
        if (own2) rec[2] = *reinterpret_cast<const CellRec<R>*>(&sb[f * CMAX * 4 + coff[2]]);
        Quad<R>::run(rec, w, W, rW, G, o);
        const R amax = max_r<R>(max_r<R>(abs_r<R>(G[0]), abs_r<R>(G[1])), max_r<R>(abs_r<R>(G[2]), abs_r<R>(G[3])));
        const R amin = min_r<R>(min_r<R>(abs_r<R>(G[0]), abs_r<R>(G[1])), min_r<R>(abs_r<R>(G[2]), abs_r<R>(G[3])));
        if (!(amax < win_hi<R>() && amin >= win_lo<R>())) {
          // rare: a sum that is zero, tiny, huge or not finite.  Zero is fine as computed; the rest
          // is outside the exact-residual window and is redone with the IEEE division.
#pragma unroll
          for (int p = 0; p < 4; p++) {
            const R ag = abs_r<R>(G[p]);
            if (!(ag == (R)0 || (ag < win_hi<R>() && ag >= win_lo<R>()))) o[p] = ((R)0 + G[p]) / W[p];
          }
        }
        store4<R>(gop, o[0], o[1], o[2], o[3]);
        __stcs(reinterpret_cast<unsigned*>(lop), 0x01010101u);
        gop += ostep; lop += ostep;
      }
    } else {
      // general path, per field: bitmaps, partial stencils, threshold failures, ragged ends
      for (int f = 0; f < F && j0 + f < jend; f++) {
        const int k = cls + 4 * (j0 + f);
        const R* __restrict__ g = a.gi + (size_t)k * a.mi;
        const u8* __restrict__ l = a.li ? a.li + (size_t)k * a.mi : nullptr;
        const bool masked = (maskbits >> (j0 + f - jm0)) & 1u;
        R o[4];
        unsigned lw = 0;
        bool bad = false;
        if (MASKED && incell) {
          // from the staged records: the reference's loop body, taps in its order, IEEE division
#pragma unroll
          for (int p = 0; p < 4; p++) {
            const CellRec<R> c = *reinterpret_cast<const CellRec<R>*>(&sb[f * CMAX * 4 + coff[p]]);
            const unsigned mw = masked ? *reinterpret_cast<const unsigned*>(&smask[f * CMAX * 4 + coff[p]]) : 0x01010101u;
            R Gp = 0, Wp = 0;
#pragma unroll
            for (int t = 0; t < 4; t++)
              if ((mw >> (8 * t)) & 0xffu) { Gp = Gp + w[p][t] * c.get(t); Wp = Wp + w[p][t]; }
            const bool good = Wp >= a.pmp;
            o[p] = good ? Gp / Wp : (R)0;
            if (good) lw |= 1u << (8 * p); else bad = true;
          }
          store4<R>(gop, o[0], o[1], o[2], o[3]);
          __stcs(reinterpret_cast<unsigned*>(lop), lw);
          if (bad) mark_bad(a.flag, k);
          gop += ostep; lop += ostep;
          continue;
        }
#pragma unroll
        for (int p = 0; p < 4; p++) {
          const int n = n0 + p;
          o[p] = 0;
          if (n < 0 || n >= a.no) continue;
          int idx[4];
#pragma unroll
          for (int t = 0; t < 4; t++) idx[t] = nxy[(size_t)t * a.no + n];
          const bool good = bilinear_point<R>(g, l, masked, idx, w[p], a.pmp, o[p]);
          if (good) lw |= 1u << (8 * p); else bad = true;
        }
        if (n0 >= 0 && n0 + 3 < a.no) {
          store4<R>(gop, o[0], o[1], o[2], o[3]);
          __stcs(reinterpret_cast<unsigned*>(lop), lw);
        } else {
#pragma unroll
          for (int p = 0; p < 4; p++)
            if (n0 + p >= 0 && n0 + p < a.no) { gop[p] = o[p]; lop[p] = (u8)((lw >> (8 * p)) & 1u); }
        }
        if (bad) mark_bad(a.flag, k);
        gop += ostep; lop += ostep;
      }
    }
    __syncwarp();
    buf = buf + 1 == NBUF ? 0 : buf + 1;
    nxt = nxt + 1 == NBUF ? 0 : nxt + 1;
  }
}

template <class R> bool tiles_ok(const Plan<R>& P, const FieldArgs<R>& a) {
  return P.tiles.ok && fast_bilinear_ok(P, a);
}

template <class R> void launch_bilinear_tiles(const Plan<R>& P, const FieldArgs<R>& a, cudaStream_t st) {
  const TileTables& T = P.tiles;
  TileDev D;
  for (int v = 0; v < 4; v++) { D.cid[v] = T.cid[v]; D.ncell[v] = T.ncell[v]; D.cells[v] = T.cells[v]; D.edst[v] = T.edst[v]; }
  D.cstride = T.cstride; D.ntile = T.ntile;
  constexpr int F = 8, F8 = sizeof(R) == 8 ? 4 : 8;   // fields staged per pass
  const int base_mis = (int)((reinterpret_cast<uintptr_t>(a.go) / sizeof(R)) & 3);
  const int per_class = (a.kc + 3) / 4;
  static const int kpt_env = getenv("IPB_KPT") ? atoi(getenv("IPB_KPT")) : 64;
  const int kpt = std::max(32, std::min(kpt_env, ((per_class + 31) / 32) * 32));   // fields of one class per warp (multiple of 32)
  dim3 grid(4 * nblk(per_class, kpt), nblk(T.ntile, TILE_WARPS));
  const int ni = (4 * T.max_cells + 31) / 32;       // staging copies per lane and field
#define IPB_TILES_LAUNCH_M(FF, NN, MM)                                                                                    \
  {                                                                                                                       \
    constexpr int SM = TILE_WARPS * (IPB_TILES_NBUF * (FF) * (8 * (NN)) * 4 * (int)sizeof(R) + ((MM) ? (FF) * (8 * (NN)) * 4 : 0)); \
    static bool once[64] = {};   /* the attribute is per device (ipgpu_set_device may switch within one process) */           \
    int dv_ = 0; cudaGetDevice(&dv_); dv_ &= 63;                                                                              \
    if (!once[dv_]) { IPB_CUDA(cudaFuncSetAttribute(k_bilinear_tiles<R, FF, NN, MM>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM)); once[dv_] = true; } \
    k_bilinear_tiles<R, FF, NN, MM><<<grid, TILE_THREADS, SM, st>>>(a, P.nxy, P.wq, P.ws2, D, base_mis, kpt);              \
  }
#define IPB_TILES_LAUNCH(FF, NN) { if (a.li) IPB_TILES_LAUNCH_M(FF, NN, true) else IPB_TILES_LAUNCH_M(FF, NN, false) }
  if (ni <= 2) IPB_TILES_LAUNCH(F, 2)
  else if (ni <= 3) IPB_TILES_LAUNCH(F, 3)
  else if (ni <= 4) IPB_TILES_LAUNCH(F, 4)
  else if (ni <= 6) IPB_TILES_LAUNCH(F8, 6)
  else IPB_TILES_LAUNCH(F8, 8)
#undef IPB_TILES_LAUNCH
#undef IPB_TILES_LAUNCH_M
  g_launches++;
}

}  // namespace ipb
