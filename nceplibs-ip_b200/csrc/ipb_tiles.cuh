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
#include "ipb_fast.cuh"

namespace ipb {

constexpr int TILE_PTS = 512;   // output points per tile
constexpr int TILE_THREADS = 128;

// ---- plan-time: cells of every tile, for one alignment variant ------------------------------------
// pass 1: per point the number of its cell inside the tile (first-appearance order; 0xff = not a
// complete 4-tap stencil), per tile the cell count.
static __global__ void __launch_bounds__(TILE_THREADS)
k_tile_index(int variant, int no, const int* __restrict__ nxy, u8* __restrict__ cid, int* __restrict__ ncell, int* gmax) {
  __shared__ int4 tup[TILE_PTS];
  __shared__ int warp_tot[TILE_THREADS / 32];
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int nbase = 4 * (tile * TILE_THREADS + tid) - variant;
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
    tup[4 * tid + p] = t;
  }
  __syncthreads();
  int first[4], nrep = 0;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int q = 4 * tid + p;
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
  // exclusive prefix of representative counts over threads
  int incl = nrep;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if ((tid & 31) >= o) incl += v; }
  if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
  __syncthreads();
  int base = incl - nrep;
  for (int w = 0; w < (tid >> 5); w++) base += warp_tot[w];
  // rank of each representative, published through the (now free) .x slot of its tuple
  __syncthreads();
  __shared__ int rank[TILE_PTS];
  int run = base;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int q = 4 * tid + p;
    rank[q] = -1;
    if (first[p] == q) rank[q] = run++;
  }
  __syncthreads();
  unsigned packed = 0;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    int c = 0xff;
    if (first[p] >= 0) { const int r = rank[first[p]]; c = r < 0xff ? r : 0xff; }
    packed |= (unsigned)c << (8 * p);
  }
  reinterpret_cast<unsigned*>(cid)[(size_t)tile * TILE_THREADS + tid] = packed;
  if (tid == TILE_THREADS - 1) {
    const int tot = base + nrep;
    ncell[tile] = tot;
    atomicMax(gmax, tot);
  }
}
// pass 2: the cell table itself, [tile][cstride][4] (1-based source positions)
static __global__ void __launch_bounds__(TILE_THREADS)
k_tile_cells(int variant, int no, int cstride, const int* __restrict__ nxy, const u8* __restrict__ cid, int* __restrict__ cells) {
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int nbase = 4 * (tile * TILE_THREADS + tid) - variant;
  const unsigned packed = reinterpret_cast<const unsigned*>(cid)[(size_t)tile * TILE_THREADS + tid];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int n = nbase + p;
    const int c = (packed >> (8 * p)) & 0xff;
    if (c == 0xff || n < 0 || n >= no) continue;
    int* dst = cells + ((size_t)tile * cstride + c) * 4;   // every point of a cell writes the same four values
#pragma unroll
    for (int t = 0; t < 4; t++) dst[t] = nxy[(size_t)t * no + n];
  }
}

template <class R> void build_tile_tables(Plan<R>& P, cudaStream_t st) {
  TileTables& T = P.tiles;
  T = TileTables();
  if (P.method != M_BILINEAR || P.vec || P.no <= 0 || !P.nxy) return;
  const int no = P.no;
  const int ng = (no + 3) / 4 + 1;                 // groups, any variant
  const int ntile = nblk(ng, TILE_THREADS);
  int* d_max = dalloc<int>(1);
  IPB_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), st));
  for (int v = 0; v < 4; v++) {
    T.cid[v] = dalloc<u8>((size_t)ntile * TILE_PTS);
    T.ncell[v] = dalloc<int>(ntile);
    k_tile_index<<<ntile, TILE_THREADS, 0, st>>>(v, no, P.nxy, T.cid[v], T.ncell[v], d_max);
    g_launches++;
  }
  int hmax = 0;
  IPB_CUDA(cudaMemcpyAsync(&hmax, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
  IPB_CUDA(cudaStreamSynchronize(st));
  dfree(d_max);
  T.ntile = ntile; T.max_cells = hmax;
  if (hmax <= 0 || hmax > 254) {                   // too little reuse for staging to pay: general kernel
    for (int v = 0; v < 4; v++) { dfree(T.cid[v]); dfree(T.ncell[v]); }
    return;
  }
  T.cstride = hmax;
  for (int v = 0; v < 4; v++) {
    T.cells[v] = dalloc<int>((size_t)ntile * T.cstride * 4);
    k_tile_cells<<<ntile, TILE_THREADS, 0, st>>>(v, no, T.cstride, P.nxy, T.cid[v], T.cells[v]);
    g_launches++;
  }
  T.bytes = 4 * ((size_t)ntile * TILE_PTS + (size_t)ntile * 4 + (size_t)ntile * T.cstride * 16);
  T.ok = true;
}

// ---- apply ---------------------------------------------------------------------------------------------
template <class R> struct CellRec;   // the four taps of one cell for one field
template <> struct CellRec<float> { float4 v; __device__ __forceinline__ float get(int t) const { return t == 0 ? v.x : t == 1 ? v.y : t == 2 ? v.z : v.w; } };
template <> struct CellRec<double> { double2 a, b; __device__ __forceinline__ double get(int t) const { return t == 0 ? a.x : t == 1 ? a.y : t == 2 ? b.x : b.y; } };

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Correctly rounded g / w for |w - 1| <= 2^-20, given rw = RN(1/w): q0 = g is within a few ulp,
// one fused correction makes it faithful, the second correctly rounded (Markstein).  Returns false
// when g is outside the exponent window in which the residuals are exact (caller divides).
template <class R> __device__ __forceinline__ bool quot_near1(R g, R w, R rw, R& q) {
  R r = fma_r<R>(-g, w, g);
  R q1 = fma_r<R>(r, rw, g);
  r = fma_r<R>(-q1, w, g);
  q = fma_r<R>(r, rw, q1);
  if (sizeof(R) == 4) {
    const unsigned b = __float_as_uint((float)g);
    const unsigned t = b + b - 0x28000000u;          // 2*bits - 2*bits(2^-87): exponent window [2^-87, 2^+87)
    return t < 0xAE000000u || g == (R)0;
  } else {
    const unsigned long long b = (unsigned long long)__double_as_longlong((double)g);
    const unsigned long long t = b + b - 0x3200000000000000ull;   // window [2^-623, 2^+623)
    return t < 0x9BC0000000000000ull || g == (R)0;
  }
}

// F = fields staged per pass, NI = staging loads per thread and field (covers 128*NI/4 cells).
template <class R, int F, int NI>
__global__ void __launch_bounds__(TILE_THREADS)
k_bilinear_tiles(FieldArgs<R> a, const int* __restrict__ nxy, const R* __restrict__ wxy, TileDev T, int base_mis, int kpt) {
  constexpr int CMAX = TILE_THREADS * NI / 4;       // cells a tile may have
  __shared__ __align__(16) R sm[F * CMAX * 4];
  const int cls = blockIdx.x & 3, slice = blockIdx.x >> 2, tile = blockIdx.y, tid = threadIdx.x;
  const int variant = (base_mis + cls * (a.mo & 3)) & 3;
  const int n0 = 4 * (tile * TILE_THREADS + tid) - variant;
  // ---- per-thread plan data: weights, weight sums, cell numbers ----
  R w[4][4], W[4], rW[4];
  int coff[4];
  bool full = true;
  const unsigned packed = reinterpret_cast<const unsigned*>(T.cid[variant])[(size_t)tile * TILE_THREADS + tid];
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int n = n0 + p;
    const bool in = n >= 0 && n < a.no;
    const int c = (packed >> (8 * p)) & 0xff;
    R ws = 0;
#pragma unroll
    for (int t = 0; t < 4; t++) { w[p][t] = in ? wxy[(size_t)t * a.no + n] : (R)0; ws = ws + w[p][t]; }
    W[p] = ws;
    rW[p] = (R)1 / ws;
    coff[p] = c * 4;
    const R dev = ws - (R)1;
    full = full && in && c != 0xff && ws >= a.pmp && dev <= (R)9.5e-7 && dev >= (R)-9.5e-7;
  }
  // ---- staging assignment: entry q = cell*4 + tap of this tile's cell table ----
  const int ncell = T.ncell[variant][tile];
  const int* __restrict__ cells = T.cells[variant] + (size_t)tile * T.cstride * 4;
  int sidx[NI];
#pragma unroll
  for (int i = 0; i < NI; i++) {
    const int q = tid + TILE_THREADS * i;
    sidx[i] = (q < 4 * ncell) ? cells[q] - 1 : -1;
  }
  const int jend = min(kpt * (slice + 1), (a.kc - cls + 3) / 4);   // fields of this class: k = cls + 4 j
  for (int j0 = kpt * slice; j0 < jend; j0 += F) {
    // ---- stage F fields' cells ----
#pragma unroll
    for (int f = 0; f < F; f++) {
      if (j0 + f < jend) {
        const R* __restrict__ g = a.gi + (size_t)(cls + 4 * (j0 + f)) * a.mi;
#pragma unroll
        for (int i = 0; i < NI; i++)
          if (sidx[i] >= 0) {
            if (sizeof(R) == 4) cp_async4(&sm[f * CMAX * 4 + tid + TILE_THREADS * i], g + sidx[i]);
            else cp_async8(&sm[f * CMAX * 4 + tid + TILE_THREADS * i], g + sidx[i]);
          }
      }
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- compute ----
    bool plain = full && (j0 + F <= jend);
    if (plain) {
#pragma unroll
      for (int f = 0; f < F; f++) plain = plain && (a.ibi[cls + 4 * (j0 + f)] == 0);
    }
    if (plain) {
#pragma unroll
      for (int f = 0; f < F; f++) {
        const size_t ob = (size_t)(cls + 4 * (j0 + f)) * a.mo + n0;
        R o[4];
        bool ok = true;
#pragma unroll
        for (int p = 0; p < 4; p++) {
          const CellRec<R> c = *reinterpret_cast<const CellRec<R>*>(&sm[f * CMAX * 4 + coff[p]]);
          R G = 0;
#pragma unroll
          for (int t = 0; t < 4; t++) G = G + w[p][t] * c.get(t);
          if (!quot_near1<R>(G, W[p], rW[p], o[p])) { ok = false; o[p] = G; }
        }
        if (!ok) {
#pragma unroll
          for (int p = 0; p < 4; p++) {   // rare: outside the exact-residual window, redo with the IEEE division
            const CellRec<R> c = *reinterpret_cast<const CellRec<R>*>(&sm[f * CMAX * 4 + coff[p]]);
            R G = 0;
#pragma unroll
            for (int t = 0; t < 4; t++) G = G + w[p][t] * c.get(t);
            o[p] = G / W[p];
          }
        }
        store4<R>(a.go + ob, o[0], o[1], o[2], o[3]);
        __stcs(reinterpret_cast<unsigned*>(a.lo + ob), 0x01010101u);
      }
    } else {
      // general path, per field: bitmaps, partial stencils, threshold failures, ragged ends
      for (int f = 0; f < F && j0 + f < jend; f++) {
        const int k = cls + 4 * (j0 + f);
        const R* __restrict__ g = a.gi + (size_t)k * a.mi;
        const u8* __restrict__ l = a.li ? a.li + (size_t)k * a.mi : nullptr;
        const bool masked = a.ibi[k] != 0;
        const size_t ob = (size_t)k * a.mo;
        R o[4];
        unsigned lw = 0;
        bool bad = false;
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
          store4<R>(a.go + ob + n0, o[0], o[1], o[2], o[3]);
          __stcs(reinterpret_cast<unsigned*>(a.lo + ob + n0), lw);
        } else {
#pragma unroll
          for (int p = 0; p < 4; p++)
            if (n0 + p >= 0 && n0 + p < a.no) { a.go[ob + n0 + p] = o[p]; a.lo[ob + n0 + p] = (u8)((lw >> (8 * p)) & 1u); }
        }
        if (bad) a.flag[k] = 1;
      }
    }
    __syncthreads();
  }
}

template <class R> bool tiles_ok(const Plan<R>& P, const FieldArgs<R>& a) {
  return P.tiles.ok && fast_bilinear_ok(P, a);
}

template <class R> void launch_bilinear_tiles(const Plan<R>& P, const FieldArgs<R>& a, cudaStream_t st) {
  const TileTables& T = P.tiles;
  TileDev D;
  for (int v = 0; v < 4; v++) { D.cid[v] = T.cid[v]; D.ncell[v] = T.ncell[v]; D.cells[v] = T.cells[v]; }
  D.cstride = T.cstride;
  constexpr int F = 8, F8 = sizeof(R) == 8 ? 4 : 8;   // fields staged per pass (48 KB of static shared memory at most)
  const int base_mis = (int)((reinterpret_cast<uintptr_t>(a.go) / sizeof(R)) & 3);
  const int per_class = (a.kc + 3) / 4;
  const int kpt = std::max(F, std::min(32, ((per_class + F - 1) / F) * F));   // fields of one class per block
  dim3 grid(4 * nblk(per_class, kpt), T.ntile);
  if (T.max_cells <= 128) k_bilinear_tiles<R, F, 4><<<grid, TILE_THREADS, 0, st>>>(a, P.nxy, P.wxy, D, base_mis, kpt);
  else k_bilinear_tiles<R, F8, 8><<<grid, TILE_THREADS, 0, st>>>(a, P.nxy, P.wxy, D, base_mis, kpt);
  g_launches++;
}

}  // namespace ipb
