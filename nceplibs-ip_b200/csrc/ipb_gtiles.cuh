// ipb200 — point-staged apply kernels: bicubic (bicubic_interp_mod.F90:226-272), the collapsed budget taps of
// the binary64 kinds (budget_interp_mod.F90:252-344 through ipb_budget.cuh), and bilinear batches that carry
// bitmaps (bilinear_interp_mod.F90:202-261; batches without one use the cell-staged kernel of ipb_tiles.cuh).
//
// The general kernels gather every tap of every output point from global memory for every field: 16
// (bicubic) or 9 (+9 bitmap bytes, budget) scattered loads per point and field, each a handful of 32-byte
// sectors per warp instruction.  That is what bounds them (L1 tag-stage wavefronts), not HBM.  Here the
// gathers leave the field loop:
//
//  * the output points are cut into TILES of 128 points — 128 consecutive points, or a 32x4 / 16x8 patch of
//    the output grid, chosen per plan by counting the distinct source points each shape needs;
//  * at plan time every tile gets the sorted list of the DISTINCT source positions its points refer to
//    (at most 256), and every tap of every point the one-byte number ("slot") of its position in that list;
//  * per pass of 8 fields the block copies those positions from each field into shared memory with
//    cp.async — address-sorted, so one warp instruction touches few sectors — double buffered; for fields
//    with a bitmap the li bytes at the same positions come along, masked-out values are zeroed in shared
//    memory and the usable-tap bits are kept as one byte per position (bit f = field f of the pass);
//  * the field loop of a thread (one output point) is then U shared-memory loads at precomputed offsets,
//    the reference's multiply/adds in its order, the division and the go/lo stores — coalesced, 32
//    consecutive points (or two runs of 16) per warp instruction.  For budget fields with a bitmap the weight
//    sum over the usable taps, its reciprocal and lo depend on the point and the bitmap only; they are kept
//    while consecutive passes show the same usable-tap pattern (the usual case: one land mask for all fields).
//
// Arithmetic is the general kernels' (same operations, same order): bicubic and bilinear results are bit-identical
// to k_apply_stencil / bilinear_point; budget values follow k_budget_collapsed's contract (ipb_budget.cuh: 1e-12,
// lo bit-exact).
// Incomplete stencils, bitmap fields of the bicubic method, tiles with more than 256 positions and
// near-threshold budget sums take the general per-point code inside the same kernel.
// (The neighbor method was tried on this scheme and stayed with k_neighbor_lane: profiles/README.md.)
#pragma once
#include <climits>
#include <type_traits>
#include <utility>
#include <cstdlib>

#include "ipb_budget.cuh"
#include "ipb_tiles.cuh"

namespace ipb {

constexpr int GT_PTS = 128;       // output points per tile = threads per block
constexpr int GT_MAXENT = 256;    // distinct source positions a tile may have (one-byte slots)
constexpr int GT_MAXENT_VEC = 512; // the same for vector bilinear plans (two-byte slots: coarse output grids share few taps)
constexpr int GT_F = 8;           // fields per staging pass

__host__ __device__ inline int gt_point(const GtGeom& g, int tile, int r, int lane, int no) {
  if (g.mode == 0) {
    const long long n = (long long)tile * GT_PTS + 32 * r + lane;
    return n < no ? (int)n : -1;
  }
  const int tx = tile % g.ntx, ty = tile / g.ntx;
  int col, row;
  if (g.mode == 1) { col = 32 * tx + lane; row = 4 * ty + r; }
  else { col = 16 * tx + (lane & 15); row = 8 * ty + 2 * r + (lane >> 4); }
  const long long n = (long long)row * g.rowlen + col;
  return (col < g.rowlen && n < no) ? (int)n : -1;
}
inline GtGeom gt_geom(int mode, int rowlen, int no) {
  GtGeom g;
  g.mode = mode; g.rowlen = rowlen; g.ntx = 1;
  if (mode == 0) { g.ntile = nblk(no, GT_PTS); return g; }
  const int nrows = nblk(no, rowlen);
  g.ntx = nblk(rowlen, mode == 1 ? 32 : 16);
  g.ntile = g.ntx * nblk(nrows, mode == 1 ? 4 : 8);
  return g;
}

// ---- plan time: the distinct source positions of every tile ----------------------------------------------
// One block per tile.  write = 0: only count (maximum and total over tiles, to choose the tile shape);
// write = 1: emit the sorted list esrc[tile][estride] (0-based, -1 = padding) and the slot of every tap.
template <int NK, int MAXENT = GT_MAXENT>   // NK: power of two >= 128 * U; MAXENT: positions a tile's list may hold
__global__ void __launch_bounds__(GT_PTS)
k_gt_build(GtGeom g, int no, int U, const int* __restrict__ taps, int write, int estride, int* __restrict__ esrc,
           u8* __restrict__ slot, int* gmax, unsigned long long* gsum, int* govf, unsigned short* __restrict__ slot16 = nullptr) {
  __shared__ int key[NK];
  __shared__ int part[GT_PTS + 1];
  __shared__ int uniq[MAXENT];
  const int tid = threadIdx.x, tile = blockIdx.x;
  const int n = gt_point(g, tile, tid >> 5, tid & 31, no);
  for (int u = 0; u < U; u++) {
    const int p = n >= 0 ? taps[(size_t)u * no + n] : 0;
    key[u * GT_PTS + tid] = p > 0 ? p - 1 : INT_MAX;
  }
  for (int i = U * GT_PTS + tid; i < NK; i += GT_PTS) key[i] = INT_MAX;
  __syncthreads();
  for (int k = 2; k <= NK; k <<= 1) {          // bitonic sort, ascending
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < NK; i += GT_PTS) {
        const int o = i ^ j;
        if (o > i) {
          const int x = key[i], y = key[o];
          if ((x > y) == ((i & k) == 0)) { key[i] = y; key[o] = x; }
        }
      }
      __syncthreads();
    }
  }
  constexpr int PER = NK / GT_PTS;
  int cnt = 0;
#pragma unroll
  for (int q = 0; q < PER; q++) {
    const int i = tid * PER + q, v = key[i];
    if (v != INT_MAX && (i == 0 || key[i - 1] != v)) cnt++;
  }
  part[tid] = cnt;
  __syncthreads();
  if (tid == 0) {
    int s = 0;
    for (int t = 0; t < GT_PTS; t++) { const int c = part[t]; part[t] = s; s += c; }
    part[GT_PTS] = s;
  }
  __syncthreads();
  const int total = part[GT_PTS];
  if (!write) {
    if (tid == 0) { atomicMax(gmax, total); atomicAdd(gsum, (unsigned long long)total); if (total > MAXENT) atomicAdd(govf, 1); }
    return;
  }
  int pos = part[tid];
#pragma unroll
  for (int q = 0; q < PER; q++) {
    const int i = tid * PER + q, v = key[i];
    if (v != INT_MAX && (i == 0 || key[i - 1] != v)) { if (pos < MAXENT) uniq[pos] = v; pos++; }
  }
  __syncthreads();
  // a tile with more positions than the table holds is marked (-2): its points gather from global memory
  const bool ovf = total > MAXENT;
  const int ne = ovf ? 0 : total;
  for (int i = tid; i < estride; i += GT_PTS) esrc[(size_t)tile * estride + i] = i < ne ? uniq[i] : (ovf && i == 0 ? -2 : -1);
  if (n < 0) return;
  for (int u = 0; u < U; u++) {
    const int p = taps[(size_t)u * no + n];
    int s = 0;
    if (p > 0 && ne > 0) {
      int lo = 0, hi = ne - 1;
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (uniq[mid] < p - 1) lo = mid + 1; else hi = mid; }
      s = lo;
    }
    if (slot16) slot16[(size_t)u * no + n] = (unsigned short)s; else slot[(size_t)u * no + n] = (u8)s;
  }
}

template <class R> void build_gtiles(Plan<R>& P, cudaStream_t st) {
  GtTables& T = P.gt;
  T = GtTables();
  static const bool off = getenv("IPB_NO_GTILES") && atoi(getenv("IPB_NO_GTILES")) != 0;
  if (off || P.no <= 0) return;
  const int* taps = nullptr;
  int U = 0;
  const bool vecb = P.vec && P.method == M_BILINEAR && P.nxy;   // vector bilinear: lists of up to 512 positions, 16-bit slots (k_gtiles_vec)
  static const bool vec_off = getenv("IPB_NO_GTILES_VEC") && atoi(getenv("IPB_NO_GTILES_VEC")) != 0;
  if (P.vec && (!vecb || vec_off)) return;
  const int maxent = vecb ? GT_MAXENT_VEC : GT_MAXENT;
  if (vecb) { taps = P.nxy; U = 4; }
  else if (P.method == M_BICUBIC && P.nxy) { taps = P.nxy; U = 16; }
  else if (P.method == M_BILINEAR && P.nxy) { taps = P.nxy; U = 4; }   // built on demand (first call with a bitmap): run()
  else if (P.method == M_BUDGET && P.cu > 0 && P.cn) { taps = P.cn; U = P.cu; }
  else return;
  const int no = P.no;
  const GridP<R>& go = P.gout.p;
  int rowlen = 0;
  if (go.type != P_STATION && go.im > 0 && go.jm > 0) rowlen = go.nscan == 0 ? go.im : go.jm;
  const int nmode = (rowlen >= 16 && rowlen < no) ? 3 : 1;
  int* d_max = dalloc<int>(8);
  int* d_ovf = d_max + 4;
  unsigned long long* d_sum = dalloc<unsigned long long>(4);
  IPB_CUDA(cudaMemsetAsync(d_max, 0, 8 * sizeof(int), st));
  IPB_CUDA(cudaMemsetAsync(d_sum, 0, 4 * sizeof(unsigned long long), st));
  auto launch = [&](const GtGeom& g, int write, int estride, int m) {
    if (vecb) k_gt_build<512, GT_MAXENT_VEC><<<g.ntile, GT_PTS, 0, st>>>(g, no, U, taps, write, estride, T.esrc, nullptr, d_max + m, d_sum + m, d_ovf + m, T.slot16);
    else if (U == 1) k_gt_build<128><<<g.ntile, GT_PTS, 0, st>>>(g, no, U, taps, write, estride, T.esrc, T.slot, d_max + m, d_sum + m, d_ovf + m);
    else if (U <= 4) k_gt_build<512><<<g.ntile, GT_PTS, 0, st>>>(g, no, U, taps, write, estride, T.esrc, T.slot, d_max + m, d_sum + m, d_ovf + m);
    else k_gt_build<2048><<<g.ntile, GT_PTS, 0, st>>>(g, no, U, taps, write, estride, T.esrc, T.slot, d_max + m, d_sum + m, d_ovf + m);
    g_launches++;
  };
  for (int m = 0; m < nmode; m++) launch(gt_geom(m, rowlen, no), 0, 0, m);
  int hmax[8];
  int* hovf = hmax + 4;
  unsigned long long hsum[4];
  IPB_CUDA(cudaMemcpyAsync(hmax, d_max, 8 * sizeof(int), cudaMemcpyDeviceToHost, st));
  IPB_CUDA(cudaMemcpyAsync(hsum, d_sum, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  IPB_CUDA(cudaStreamSynchronize(st));
  if (getenv("IPB_DEBUG"))
    for (int m = 0; m < nmode; m++)
      fprintf(stderr, "ipb200: gtiles method %d U %d mode %d: tiles %d, max %d, mean %.1f distinct positions per tile, %d tiles over %d\n", P.method, U, m,
              gt_geom(m, rowlen, no).ntile, hmax[m], (double)hsum[m] / gt_geom(m, rowlen, no).ntile, hovf[m], maxent);
  // Tile shape: fewest staging copies per field, where a tile whose list does not fit (it gathers from global
  // memory) counts like 2048 copies and the 16x8 patch pays for its short store runs (two 16-point runs per
  // warp instruction instead of one 32-point run).  Measured on the BASELINE pairs: profiles/README.md.
  int best = -1;
  double best_cost = 0;
  static const int force = getenv("IPB_GMODE") ? atoi(getenv("IPB_GMODE")) : -1;
  for (int m = 0; m < nmode; m++) {
    if (hmax[m] <= 0) continue;
    const int nt = gt_geom(m, rowlen, no).ntile;
    // only the bicubic kernel has a per-tile fallback for lists that do not fit
    if (hmax[m] > maxent && !((P.method == M_BICUBIC || vecb) && hovf[m] <= nt / 8)) continue;
    if (force >= 0 && m != force) continue;
    const double cost = (double)hsum[m] + 2048.0 * hovf[m] + (m == 2 ? 64.0 * nt : 0.0);
    if (best < 0 || cost < best_cost) { best = m; best_cost = cost; }
  }
  if (best >= 0) {
    T.geom = gt_geom(best, rowlen, no);
    T.U = U;
    T.maxent = hmax[best];
    T.estride = vecb ? GT_PTS * std::min(4, std::max(1, (std::min(hmax[best], maxent) + GT_PTS - 1) / GT_PTS)) : (hmax[best] <= GT_PTS ? GT_PTS : 2 * GT_PTS);
    T.novf = hovf[best];
    T.esrc = dalloc<int>((size_t)T.geom.ntile * T.estride);
    if (vecb) T.slot16 = dalloc<unsigned short>((size_t)U * no); else T.slot = dalloc<u8>((size_t)U * no);
    launch(T.geom, 1, T.estride, 3);
    T.entries = hsum[best];
    T.bytes = (size_t)T.geom.ntile * T.estride * 4 + (size_t)U * no * (vecb ? 2 : 1);
    T.ok = true;
  }
  dfree(d_max); dfree(d_sum);
}

// ---- apply ---------------------------------------------------------------------------------------------
enum GtKind { GT_STENCIL = 0, GT_BILIN = 1, GT_BUDGET = 2 };   // bicubic; bilinear with bitmaps; collapsed budget

template <class R> struct GtArgs {
  GtGeom g;
  int estride;
  const int* esrc;       // [ntile][estride]
  const u8* slot;        // [U][no]
  const int* taps;       // [U][no] 1-based source positions (stencil: nxy; budget: cn)
  const R* w;            // [U][no] weights (stencil: wxy; budget: cw)
  const u8* nc;          // bicubic stencil class per point
  const R* wsum;         // budget: sum of the sample weights per point
  int nsamp; const int* bn; const R* bw; const R* wbs;   // budget: the sub-sample plan, for the exact fallback
};

// One point, one field of the bicubic method from global memory (k_apply_stencil's general path, scalar):
// incomplete stencils, the bilinear fallback class, bitmaps.
template <class R>
__device__ __noinline__ bool bicubic_point_general(const FieldArgs<R>& a, int k, int n, int cls, const int* __restrict__ nxy,
                                                   const R* __restrict__ wxy, R& og) {
  const R* __restrict__ g = a.gi + (size_t)k * a.mi;
  const bool masked = a.ibi[k] != 0;
  const u8* __restrict__ l = a.li ? a.li + (size_t)k * a.mi : nullptr;
  R G = 0, W = 0, gmin = DM<R>::huge(), gmax = -DM<R>::huge();
  if (cls > 0) {
    for (int t = 0; t < 16; t++) {
      if (cls == 2) {
        const int ta = t & 3, tb = t >> 2;
        if (ta < 1 || ta > 2 || tb < 1 || tb > 2) continue;
      }
      const int p = nxy[(size_t)t * a.no + n];
      if (p > 0 && (!masked || l[p - 1])) {
        const R w = wxy[(size_t)t * a.no + n], v = g[p - 1];
        G = G + w * v;
        W = W + w;
        if (a.mcon > 0) { gmin = DM<R>::vmin(gmin, v); gmax = DM<R>::vmax(gmax, v); }
      }
    }
  }
  const bool good = cls > 0 && W >= a.pmp;
  og = 0;
  if (good) {
    og = G / W;
    if (a.mcon > 0) og = DM<R>::vmin(DM<R>::vmax(og, gmin), gmax);
  }
  return good;
}

// unpredicated 4- / 8-byte asynchronous copy global -> shared
template <int BYTES> __device__ __forceinline__ void cp_async_b(unsigned smem_addr, const void* gmem) {
  if (BYTES == 4) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem_addr), "l"(gmem));
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_addr), "l"(gmem));
}

// shared-memory loads by 32-bit shared address: register + compile-time offset, so that a field of the pass costs no
// address arithmetic (the compiler does not hoist the generic -> shared conversion out of the field loop).  The loads
// are not volatile — the scheduler may interleave the taps of several fields — but their address registers are
// re-formed after every pass's barrier with a `token` (zero, produced by a volatile asm, opaque to the compiler), so none
// of them can move above that barrier; their results feed stores that sit before the next one.
template <class R, int IMM> __device__ __forceinline__ R lds_at(unsigned addr) {
  R v;
  if (sizeof(R) == 4) asm("ld.shared.f32 %0, [%1+%2];" : "=f"(*reinterpret_cast<float*>(&v)) : "r"(addr), "n"(IMM));
  else asm("ld.shared.f64 %0, [%1+%2];" : "=d"(*reinterpret_cast<double*>(&v)) : "r"(addr), "n"(IMM));
  return v;
}
// two adjacent values (the u and v of one staged position) with one shared-memory load
template <class R, int IMM> __device__ __forceinline__ void lds2_at(unsigned addr, R& x, R& y) {
  if (sizeof(R) == 4) asm("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(*reinterpret_cast<float*>(&x)), "=f"(*reinterpret_cast<float*>(&y)) : "r"(addr), "n"(IMM));
  else asm("ld.shared.v2.f64 {%0, %1}, [%2+%3];" : "=d"(*reinterpret_cast<double*>(&x)), "=d"(*reinterpret_cast<double*>(&y)) : "r"(addr), "n"(IMM));
}
__device__ __forceinline__ unsigned pass_token() { unsigned t; asm volatile("mov.u32 %0, 0;" : "=r"(t) :: "memory"); return t; }

template <class Fn, int... I> __device__ __forceinline__ void static_for_impl(Fn&& fn, std::integer_sequence<int, I...>) { (fn(std::integral_constant<int, I>{}), ...); }
template <int N, class Fn> __device__ __forceinline__ void static_for(Fn&& fn) { static_for_impl(fn, std::make_integer_sequence<int, N>{}); }

// num / den with the IEEE quotient's bits, without sending a zero numerator through the division's slow path (FCHK
// rejects it; masked-out points and fields with exact zeros are full of them): 0 / den = +0 for the den > 0 used here,
// and the sums formed from +0 are never -0.
// (the operands are selected in inline asm: the compiler otherwise folds the selects away and divides 0 by 0 again)
__device__ __forceinline__ float one_unless(float x, bool on) {
  float r; asm("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n selp.f32 %0, %1, 0f3F800000, p;\n}\n" : "=f"(r) : "f"(x), "r"((unsigned)on)); return r;
}
__device__ __forceinline__ double one_unless(double x, bool on) {
  double r; asm("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n selp.f64 %0, %1, 0d3FF0000000000000, p;\n}\n" : "=d"(r) : "d"(x), "r"((unsigned)on)); return r;
}
template <class R> __device__ __forceinline__ R div_or_zero(R num, R den, bool on) {
  const bool nz = on && num != (R)0;
  const R q = one_unless(num, nz) / one_unless(den, nz);
  return nz ? q : (R)0;
}

constexpr int GT_MAXKPB = 256;   // fields per block at most (bitmap flags of the block's fields: 8 words)

// NE = staging entries per thread (the tile's list has up to 128*NE positions); MASKED = the batch has
// fields with a bitmap: the li bytes at the tile's positions are fetched with the values, masked-out values
// are zeroed in shared memory (w*0 adds nothing, whatever the caller left in a masked-out point) and the
// usable-tap bits go to shared memory as one byte per position (bit f = field f of the pass).
#ifndef IPB_GT_MINB_BUDGET
#define IPB_GT_MINB_BUDGET 4
#endif
template <class R, int U, int KIND, int NE, bool MASKED, int FP = GT_F, int MB = 0>   // FP < 8 only without bitmaps (the usable-tap bytes hold 8 fields)
__global__ void __launch_bounds__(GT_PTS, MB > 0 ? MB : (KIND == GT_BUDGET ? IPB_GT_MINB_BUDGET : (sizeof(R) == 4 ? 5 : 4)))
k_gtiles(FieldArgs<R> a, GtArgs<R> T, int kpb) {
  static_assert(FP == 8 || !MASKED, "bitmap passes are 8 fields wide");
  constexpr int F = FP, ESTR = GT_PTS * NE, BUF = F * ESTR;
  constexpr int NW = (U + 3) / 4;
  constexpr bool SUMK = KIND == GT_BUDGET || KIND == GT_BILIN;   // weighted sum over the usable taps, bitmaps from shared memory
  constexpr bool EXACT = KIND == GT_BILIN;                        // bit-exact contract: IEEE division by the weight sum in tap order
  extern __shared__ __align__(16) unsigned char gt_sm[];
  R* sm = reinterpret_cast<R*>(gt_sm);                    // [2][F][ESTR] field values at the tile's positions
  u8* smk = gt_sm + (size_t)2 * BUF * sizeof(R);          // [ESTR] MASKED: bit f = position usable in field f of the pass
  __shared__ unsigned s_fm[GT_MAXKPB / 32];               // bit j: field k0+j carries a bitmap
  const int tid = threadIdx.x, r = tid >> 5, lane = tid & 31, tile = blockIdx.x;
  const int k0 = blockIdx.y * kpb, k1 = min(k0 + kpb, a.kc);
  if (k0 >= k1) return;
  const int no = a.no;
  const int n = gt_point(T.g, tile, r, lane, no);
  const bool live = n >= 0;
  const size_t nn = live ? (size_t)n : 0;
  const size_t mi = (size_t)a.mi;
  if (a.li) {
    for (int w = r; w * 32 < k1 - k0; w += GT_PTS / 32) {
      const int k = k0 + w * 32 + lane;
      const unsigned b = __ballot_sync(0xffffffffu, k < k1 && a.ibi[k] != 0);
      if (lane == 0) s_fm[w] = b;
    }
    __syncthreads();
  }
  auto field_mask = [&](int kb) -> unsigned { return a.li ? (s_fm[(kb - k0) >> 5] >> ((kb - k0) & 31)) & 0xffu : 0u; };
  // ---- staging assignment: entry q of the tile's sorted list goes to shared-memory slot q ----
  const int* __restrict__ es = T.esrc + (size_t)tile * T.estride;
  const bool ovf = es[0] == -2;      // more than 256 distinct positions: this tile gathers from global memory
  int src[NE];
  bool act[NE], wact[NE];            // wact: some lane of the warp has an entry (warp-uniform)
#pragma unroll
  for (int i = 0; i < NE; i++) {
    const int s = es[tid + GT_PTS * i];
    act[i] = s >= 0;
    src[i] = act[i] ? s : 0;
    wact[i] = __any_sync(0xffffffffu, act[i]);
  }
  const unsigned sm_base = (unsigned)__cvta_generic_to_shared(sm);
  auto stage = [&](int buf, int kb) {
    const int nf = min(F, k1 - kb);
#pragma unroll
    for (int i = 0; i < NE; i++) {
      if (act[i]) {
        const R* __restrict__ gp = a.gi + (size_t)kb * mi + src[i];
        const unsigned d = sm_base + (unsigned)((buf * BUF + tid + GT_PTS * i) * sizeof(R));
        if (nf == F) {
#pragma unroll
          for (int f = 0; f < F; f++) cp_async_b<sizeof(R)>(d + (unsigned)(f * ESTR * sizeof(R)), gp + (size_t)f * mi);
        } else {
          for (int f = 0; f < nf; f++) cp_async_b<sizeof(R)>(d + (unsigned)(f * ESTR * sizeof(R)), gp + (size_t)f * mi);
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  u8 mraw[NE][F];
  auto load_masks = [&](int kb, unsigned m) {   // li bytes of pass kb at this thread's entries
#pragma unroll
    for (int i = 0; i < NE; i++) {
#pragma unroll
      for (int f = 0; f < F; f++) mraw[i][f] = 1;
      if (MASKED && m != 0 && wact[i]) {
        const u8* __restrict__ lp = a.li + (size_t)kb * mi + src[i];
        if (m == 0xffu) {
#pragma unroll
          for (int f = 0; f < F; f++) mraw[i][f] = __ldg(lp + (size_t)f * mi);
        } else {
#pragma unroll
          for (int f = 0; f < F; f++) if ((m >> f) & 1u) mraw[i][f] = __ldg(lp + (size_t)f * mi);
        }
      }
    }
  };
  auto fix_masks = [&](int buf) {               // after this thread's own copies of the pass have landed
#pragma unroll
    for (int i = 0; i < NE; i++) {
      if (wact[i]) {
        unsigned b = 0;
#pragma unroll
        for (int f = 0; f < F; f++) b |= (mraw[i][f] ? 1u : 0u) << f;
        smk[tid + GT_PTS * i] = (u8)b;
        if (b != 0xffu) {
#pragma unroll
          for (int f = 0; f < F; f++) if (!((b >> f) & 1u)) sm[buf * BUF + f * ESTR + tid + GT_PTS * i] = (R)0;
        }
      }
    }
  };
  // ---- per-point plan data ----
  // every plan load of the point is issued before any is consumed (unconditional loads at point 0 for lanes without a
  // point): one memory latency instead of one per tap group — ncu had 16 % (bicubic) / 7 % (budget) of the warps' time here
  R w[U];
  unsigned soff[U];
  int tapv[(KIND == GT_BUDGET) ? 1 : U];
  {
    u8 sl[U];
#pragma unroll
    for (int u = 0; u < U; u++) sl[u] = __ldg(T.slot + (size_t)u * no + nn);
#pragma unroll
    for (int u = 0; u < U; u++) w[u] = __ldg(T.w + (size_t)u * no + nn);
    if (KIND != GT_BUDGET) {
#pragma unroll
      for (int u = 0; u < U; u++) tapv[(KIND == GT_BUDGET) ? 0 : u] = __ldg(T.taps + (size_t)u * no + nn);
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      // shared address of the tap in field 0 of the buffer in use (moved from buffer to buffer after every pass)
      soff[u] = sm_base + (live ? (unsigned)sl[u] : 0u) * (unsigned)sizeof(R);
      if (KIND == GT_STENCIL && ovf) { const int p = live ? tapv[(KIND == GT_BUDGET) ? 0 : u] : 0; soff[u] = p > 0 ? (unsigned)(p - 1) : 0u; }   // element offset in the field
      if (!live) w[u] = (R)0;
    }
  }
  bool complete = live;       // stencil: the straight-line path applies
  int cls = 0;
  R Wall = 0, rWall = 0;      // stencil: weight sum; budget: sample-weight sum and its reciprocal
  if (KIND == GT_STENCIL) {
    cls = live ? T.nc[nn] : 0;
    complete = complete && cls == 1;
#pragma unroll
    for (int u = 0; u < U; u++) { complete = complete && (live ? tapv[(KIND == GT_BUDGET) ? 0 : u] : 0) > 0; Wall = Wall + w[u]; }
    complete = complete && Wall >= a.pmp;
  } else if (KIND == GT_BILIN) {
    // bilinear_interp_mod.F90:210-219: W accumulates the usable weights in tap order; a missing tap (grid edge) sends
    // the point to the general per-point code
#pragma unroll
    for (int u = 0; u < U; u++) { complete = complete && (live ? tapv[(KIND == GT_BUDGET) ? 0 : u] : 0) > 0; Wall = Wall + w[u]; }
  } else {
    Wall = live ? T.wsum[nn] : (R)0;
    rWall = Wall > 0 ? (R)1 / Wall : (R)0;
  }
  int utaps = U;              // budget: taps up to the last non-zero weight, maximum over the warp's points
  if (KIND == GT_BUDGET) {
    int cnt = 0;
#pragma unroll
    for (int u = 0; u < U; u++) if (w[u] != (R)0) cnt = u + 1;
    utaps = (int)__reduce_max_sync(0xffffffffu, (unsigned)cnt);
  }
  R* go = a.go + (size_t)k0 * a.mo + nn;
  u8* lo = a.lo + (size_t)k0 * a.mo + nn;
  // budget with bitmaps: weight sum of the usable taps, remembered while the usable-tap pattern stays the same
  unsigned lastp[NW];
  R wsel = Wall, rsel = rWall;       // the weight sum in force (all taps: the exact sample-weight sum) and its reciprocal
  bool near_c = false, ones_c = true;
#pragma unroll
  for (int q = 0; q < NW; q++) lastp[q] = 0xffffffffu;   // not a pattern (patterns have bits 0, 8, 16, 24 only)
  // ---- pipeline ----
  unsigned fm = field_mask(k0);
  stage(0, k0);
  if (MASKED) load_masks(k0, fm);
  int buf = 0;
  for (int kb = k0; kb < k1; kb += F) {
    const bool more = kb + F < k1;
    if (more) stage(buf ^ 1, kb + F);
    else asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    if (MASKED) {
      if (fm) fix_masks(buf);
      if (more) load_masks(kb + F, field_mask(kb + F));
    }
    __syncthreads();
    // ---- compute the pass from shared memory ----
    {
      // move the tap addresses to this pass's buffer (and make them depend on something produced after the barrier)
      const unsigned tok = pass_token() + (kb == k0 ? 0u : (buf ? (unsigned)(BUF * sizeof(R)) : 0u - (unsigned)(BUF * sizeof(R))));
      if (!(KIND == GT_STENCIL && ovf)) {
#pragma unroll
        for (int u = 0; u < U; u++) soff[u] += tok;
      }
    }
    const int nf = min(F, k1 - kb);
    const unsigned sbuf = sm_base + (unsigned)(buf * BUF * sizeof(R));
    unsigned mbp[NW];            // per tap one byte, bit f = tap usable in field f of the pass
    unsigned allm = 0xffu;
#pragma unroll
    for (int q = 0; q < NW; q++) mbp[q] = 0;
    if (MASKED && fm) {
#pragma unroll
      for (int u = 0; u < U; u++) {
        const unsigned b = live ? (unsigned)smk[(soff[u] - sbuf) / (unsigned)sizeof(R)] : 0xffu;
        mbp[u >> 2] |= b << (8 * (u & 3));
        allm &= b;
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; u++) mbp[u >> 2] |= 0xffu << (8 * (u & 3));
    }
    // The straight-line form of the pass (warp-uniform choice): all F fields present and
    //   stencil  — every live point of the warp has its complete stencil, no field of the pass has a bitmap;
    //   budget   — every field of the pass sees the same usable taps at every point of the warp (then the weight
    //              sum, its reciprocal and lo are per point and pass, remembered while the pattern stays), and
    //              no point's weight sum sits next to the threshold.
    bool fast = nf == F;
    if (KIND == GT_STENCIL) fast = fast && fm == 0 && __all_sync(0xffffffffu, complete || !live);
    if (KIND == GT_BILIN) fast = fast && __all_sync(0xffffffffu, complete || !live);
    if (SUMK && MASKED && fm) {
      bool uni = true;
#pragma unroll
      for (int q = 0; q < NW; q++) uni = uni && (((mbp[q] ^ (mbp[q] >> 1)) & 0x7f7f7f7fu) == 0);
      fast = fast && __all_sync(0xffffffffu, uni);
      if (fast) {
        bool same = true;
#pragma unroll
        for (int q = 0; q < NW; q++) same = same && (mbp[q] & 0x01010101u) == lastp[q];
        if (!same) {
          R wo = 0;
#pragma unroll
          for (int u = 0; u < U; u++) if ((mbp[u >> 2] >> (8 * (u & 3))) & 1u) wo = wo + w[u];
          ones_c = (allm & 1u) != 0;
          // a field with a bitmap accumulates WB*W per tap in floating point (budget_interp_mod.F90:262-277) even when
          // every tap is usable, a field without one the integer WB: next to the threshold (mp = 100) the two can
          // differ, so the reference's own order and association decide there
          const R d = (ones_c ? Wall : wo) - a.pmp;
          near_c = !EXACT && live && (d < 0 ? -d : d) <= a.pmp * (R)1e-9;
          wsel = ones_c ? Wall : wo;
          rsel = ones_c ? rWall : (wo > 0 ? (R)1 / wo : (R)0);
#pragma unroll
          for (int q = 0; q < NW; q++) lastp[q] = mbp[q] & 0x01010101u;
        }
        fast = !__any_sync(0xffffffffu, near_c);
      }
    } else if (SUMK) {
      if (lastp[0] != 0x01010101u || !ones_c) {   // back to "every tap usable"
        ones_c = true; near_c = false; wsel = Wall; rsel = rWall;
#pragma unroll
        for (int q = 0; q < NW; q++) lastp[q] = 0xffffffffu;
        lastp[0] = 0x01010101u;
      }
    }
    if (fast) {
      const R* __restrict__ gk = a.gi + (size_t)kb * mi;
      if (SUMK) {
        const bool good = wsel >= a.pmp;
        const R rs = EXACT ? wsel : rsel;   // EXACT divides by the weight sum, budget multiplies by its reciprocal
        // ALL: every lane of the warp has a point (all tiles but those at the grid's edge): unconditional stores,
        // one basic block for the whole pass, so that the taps of several fields can be in flight together.
        // UU: taps actually in use by the warp's points.  The collapsed taps are padded to U with weight zero, and
        // most points need fewer — 4 when all sub-samples fall into one source cell, 6 when they straddle one cell
        // edge, 9 at a corner; a sum that stops at the last non-zero weight has the same bits (acc + 0 = acc).
        auto pass = [&](auto allc, auto uuc) {
          constexpr bool ALL = decltype(allc)::value;
          constexpr int UU = decltype(uuc)::value < U ? decltype(uuc)::value : U;
          static_for<F>([&](auto fc) {
            constexpr int f = decltype(fc)::value;
            R v[UU];
#pragma unroll
            for (int u = 0; u < UU; u++) v[u] = lds_at<R, f * ESTR * (int)sizeof(R)>(soff[u]);
            R acc = 0;   // masked-out values are zero in shared memory: the plain sum is the sum over the usable taps
#pragma unroll
            for (int u = 0; u < UU; u++) acc = acc + w[u] * v[u];
            const R og = EXACT ? div_or_zero<R>(acc, rs, good) : (good ? acc * rs : (R)0);
            if (ALL || live) { __stcs(go, og); __stcs(lo, (u8)(good ? 1 : 0)); }
            go += a.mo; lo += a.mo;
          });
        };
        const bool all_live = __all_sync(0xffffffffu, live);
        auto pick = [&](auto uuc) { if (all_live) pass(std::true_type{}, uuc); else pass(std::false_type{}, uuc); };
        if (U > 4 && utaps <= 4) pick(std::integral_constant<int, 4>{});
        else if (U > 6 && utaps <= 6) pick(std::integral_constant<int, 6>{});
        else if (U > 9 && utaps <= 9) pick(std::integral_constant<int, 9>{});
        else pick(std::integral_constant<int, U>{});
        if (live && !good) for (int f = 0; f < F; f++) mark_bad(a.flag, kb + f);
      } else if (KIND == GT_STENCIL) {
        // the tap sums of the whole pass first, as one basic block (OVF and CLAMP are compile-time here), so that
        // the loads of several fields are in flight together; divisions, clamps and stores afterwards
        auto pass = [&](auto ovfc, auto clampc) {
          constexpr bool OVF = decltype(ovfc)::value, CLAMP = decltype(clampc)::value;
          R G[F], gmin[CLAMP ? F : 1], gmax[CLAMP ? F : 1];
          static_for<F>([&](auto fc) {
            constexpr int f = decltype(fc)::value;
            R v[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
              if (!OVF) v[u] = lds_at<R, f * ESTR * (int)sizeof(R)>(soff[u]);
              else v[u] = __ldg(gk + (size_t)f * mi + soff[u]);
            }
            R g = 0;
#pragma unroll
            for (int u = 0; u < U; u++) g = g + w[u] * v[u];
            G[f] = g;
            if (CLAMP) {
              R lo_ = DM<R>::huge(), hi_ = -DM<R>::huge();
#pragma unroll
              for (int u = 0; u < U; u++) { lo_ = DM<R>::vmin(lo_, v[u]); hi_ = DM<R>::vmax(hi_, v[u]); }
              gmin[CLAMP ? f : 0] = lo_; gmax[CLAMP ? f : 0] = hi_;
            }
          });
          static_for<F>([&](auto fc) {
            constexpr int f = decltype(fc)::value;
            R og = div_or_zero<R>(G[f], Wall, true);
            if (CLAMP) og = DM<R>::vmin(DM<R>::vmax(og, gmin[CLAMP ? f : 0]), gmax[CLAMP ? f : 0]);
            if (live) { __stcs(go, og); __stcs(lo, (u8)1); }
            go += a.mo; lo += a.mo;
          });
        };
        if (!ovf) { if (a.mcon > 0) pass(std::false_type{}, std::true_type{}); else pass(std::false_type{}, std::false_type{}); }
        else { if (a.mcon > 0) pass(std::true_type{}, std::true_type{}); else pass(std::true_type{}, std::false_type{}); }
      }
    } else {
      // general form, per field: ragged last pass, incomplete stencils, bicubic bitmap fields, budget fields that
      // differ in their usable taps or have a weight sum next to the threshold
#pragma unroll 1
      for (int f = 0; f < nf; f++) {
        const int k = kb + f;
        const unsigned fo = (unsigned)(f * ESTR * sizeof(R));
        R og = 0;
        bool good = false;
        if (KIND == GT_STENCIL) {
          if (complete && !((fm >> f) & 1u)) {
            R v[U];
            if (!ovf) {
#pragma unroll
              for (int u = 0; u < U; u++) v[u] = lds_at<R, 0>(soff[u] + fo);
            } else {
#pragma unroll
              for (int u = 0; u < U; u++) v[u] = __ldg(a.gi + (size_t)k * mi + soff[u]);
            }
            R G = 0;
#pragma unroll
            for (int u = 0; u < U; u++) G = G + w[u] * v[u];
            og = G / Wall;
            if (a.mcon > 0) {
              R gmin = DM<R>::huge(), gmax = -DM<R>::huge();
#pragma unroll
              for (int u = 0; u < U; u++) { gmin = DM<R>::vmin(gmin, v[u]); gmax = DM<R>::vmax(gmax, v[u]); }
              og = DM<R>::vmin(DM<R>::vmax(og, gmin), gmax);
            }
            good = true;
          } else if (live) {
            good = bicubic_point_general<R>(a, k, n, cls, T.taps, T.w, og);
          }
        } else if (KIND == GT_BILIN) {
          if (!complete) {
            if (live) {   // a tap missing at the grid's edge: the reference loop body from global memory
              if constexpr (U == 4) {
                int idx[4];
#pragma unroll
                for (int u = 0; u < 4; u++) idx[u] = T.taps[(size_t)u * no + nn];
                good = bilinear_point<R>(a.gi + (size_t)k * mi, a.li ? a.li + (size_t)k * mi : nullptr, ((fm >> f) & 1u) != 0, idx, w, a.pmp, og);
              }
            }
          } else {
            R v[U];
#pragma unroll
            for (int u = 0; u < U; u++) v[u] = lds_at<R, 0>(soff[u] + fo);
            R acc = 0, wo = 0;
#pragma unroll
            for (int u = 0; u < U; u++) add2_if<R>((mbp[u >> 2] >> (8 * (u & 3) + f)) & 1u, acc, w[u] * v[u], wo, w[u]);
            good = wo >= a.pmp;
            og = div_or_zero<R>(acc, wo, good);
          }
        } else {
          R v[U];
#pragma unroll
          for (int u = 0; u < U; u++) v[u] = lds_at<R, 0>(soff[u] + fo);
          R acc = 0, wo = 0;
#pragma unroll
          for (int u = 0; u < U; u++) add2_if<R>((mbp[u >> 2] >> (8 * (u & 3) + f)) & 1u, acc, w[u] * v[u], wo, w[u]);
          const R dall = Wall - a.pmp;
          const bool near_all = ((fm >> f) & 1u) != 0 && (dall < 0 ? -dall : dall) <= a.pmp * (R)1e-9;   // bitmap field at the threshold
          if (((allm >> f) & 1u) && !near_all) {
            good = Wall >= a.pmp;
            og = good ? acc * rWall : (R)0;
          } else {
            const R d = wo - a.pmp;
            if (live && (near_all || (d < 0 ? -d : d) <= a.pmp * (R)1e-9)) {   // the reference's order and association decide lo
              acc = budget_exact_point<R>(a.gi + (size_t)k * mi, a.li + (size_t)k * mi, no, n, true, T.nsamp, T.bn, T.bw, T.wbs, wo);
              good = wo >= a.pmp;
              og = good ? acc / wo : (R)0;
            } else {
              good = wo >= a.pmp;
              og = good ? acc * (wo > 0 ? (R)1 / wo : (R)0) : (R)0;
            }
          }
        }
        if (live) {
          __stcs(go, og);
          __stcs(lo, (u8)(good ? 1 : 0));
          if (!good) mark_bad(a.flag, k);
        }
        go += a.mo; lo += a.mo;
      }
    }
    __syncthreads();
    fm = field_mask(kb + F < k1 ? kb + F : k0);
    buf ^= 1;
  }
}


// ---- vector bilinear (ipolatev ip = 0; bilinear_interp_mod.F90:463-546): both wind components point-staged -------------
// The same scheme for field pairs: u and v at the tile's distinct source positions go to shared memory once per pass, every
// point reads its four taps of both from there, rotates them tap by tap (cxy, sxy from movect), sums in the reference's
// order, rotates to the output grid (crot, srot) and divides by the weight sum — k_apply_stencil<2, true>'s arithmetic,
// bit for bit.  Output grids coarser than the input share few taps (NH polar stereographic 512 x 512 from 0.25 degrees: 232
// distinct positions per 128-point tile on average, 512 at the most), hence lists of up to 512 positions and two-byte
// slots; what is gained is the address-sorted staging: a handful of sectors per warp instruction instead of twenty.
// Batches with bitmaps and points with an incomplete stencil take the general per-point code.
// Fields per pass and resident blocks per SM, measured on BASELINE config 4 (S -> NH polar stereographic, km = 512 pairs;
// profiles/README.md): 8 / 4 / 2 / 1 pairs per pass with 3 / 4 / 7 / 8 blocks -> 1.21 / 1.18 / 0.93 / 1.08 ms.  Short passes
// keep the staging of one block from holding the load queue of the SM (ncu on the 4-pair form: mio_throttle 2.8 per issue).
constexpr int GT_VEC_F = 2;
template <class R> struct GtVecB { static constexpr int v = sizeof(R) == 4 ? 7 : 4; };

template <class R, int NE, int F = GT_VEC_F, int MINB = GtVecB<R>::v>   // F: field pairs per pass
__global__ void __launch_bounds__(GT_PTS, MINB)
k_gtiles_vec(FieldArgs<R> a, GtGeom g, int estride, const int* __restrict__ esrc, const unsigned short* __restrict__ slot16,
             const int* __restrict__ nxy, const R* __restrict__ wxy, const R* __restrict__ cxy, const R* __restrict__ sxy,
             const R* __restrict__ crot, const R* __restrict__ srot, int kpb) {
  constexpr int ESTR = GT_PTS * NE, COMP = F * ESTR, BUF = 2 * COMP;
  extern __shared__ __align__(16) unsigned char gt_sm[];
  R* sm = reinterpret_cast<R*>(gt_sm);                    // [2 buffers][F][ESTR][u, v]: one load brings both components of a tap
  const int tid = threadIdx.x, r = tid >> 5, lane = tid & 31, tile = blockIdx.x;
  const int k0 = blockIdx.y * kpb, k1 = min(k0 + kpb, a.kc);
  if (k0 >= k1) return;
  const int no = a.no;
  const int n = gt_point(g, tile, r, lane, no);
  const bool live = n >= 0;
  const size_t nn = live ? (size_t)n : 0;
  const size_t mi = (size_t)a.mi;
  const int* __restrict__ es = esrc + (size_t)tile * estride;
  int src[NE];
  bool act[NE];
#pragma unroll
  for (int i = 0; i < NE; i++) {
    const int s = es[tid + GT_PTS * i];
    act[i] = s >= 0;
    src[i] = act[i] ? s : 0;
  }
  const unsigned sm_base = (unsigned)__cvta_generic_to_shared(sm);
  auto stage = [&](int buf, int kb) {
    const int nf = min(F, k1 - kb);
#pragma unroll
    for (int i = 0; i < NE; i++) {
      if (act[i]) {
        const R* __restrict__ gp = a.gi + (size_t)kb * mi + src[i];
        const R* __restrict__ vp = a.vi + (size_t)kb * mi + src[i];
        const unsigned d = sm_base + (unsigned)((buf * BUF + 2 * (tid + GT_PTS * i)) * sizeof(R));
        if (nf == F) {
#pragma unroll
          for (int f = 0; f < F; f++) {
            cp_async_b<sizeof(R)>(d + (unsigned)(2 * f * ESTR * sizeof(R)), gp + (size_t)f * mi);
            cp_async_b<sizeof(R)>(d + (unsigned)((2 * f * ESTR + 1) * sizeof(R)), vp + (size_t)f * mi);
          }
        } else {
          for (int f = 0; f < nf; f++) {
            cp_async_b<sizeof(R)>(d + (unsigned)(2 * f * ESTR * sizeof(R)), gp + (size_t)f * mi);
            cp_async_b<sizeof(R)>(d + (unsigned)((2 * f * ESTR + 1) * sizeof(R)), vp + (size_t)f * mi);
          }
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };
  // ---- per-point plan data, every load issued before any is consumed ----
  R w[4], c[4], s[4];
  int idx[4];
  unsigned soff[4];
  {
    unsigned short sl[4];
#pragma unroll
    for (int t = 0; t < 4; t++) sl[t] = __ldg(slot16 + (size_t)t * no + nn);
#pragma unroll
    for (int t = 0; t < 4; t++) { idx[t] = __ldg(nxy + (size_t)t * no + nn); w[t] = __ldg(wxy + (size_t)t * no + nn); }
#pragma unroll
    for (int t = 0; t < 4; t++) { c[t] = __ldg(cxy + (size_t)t * no + nn); s[t] = __ldg(sxy + (size_t)t * no + nn); }
#pragma unroll
    for (int t = 0; t < 4; t++) soff[t] = sm_base + (live ? (unsigned)sl[t] : 0u) * (unsigned)(2 * sizeof(R));
  }
  const R cr = __ldg(crot + nn), sr = __ldg(srot + nn);
  bool complete = live;
  R Wall = 0;
#pragma unroll
  for (int t = 0; t < 4; t++) { complete = complete && idx[t] > 0; Wall = Wall + w[t]; }
  complete = complete && Wall >= a.pmp;
  R* go = a.go + (size_t)k0 * a.mo + nn;
  R* vo = a.vo + (size_t)k0 * a.mo + nn;
  u8* lo = a.lo + (size_t)k0 * a.mo + nn;
  stage(0, k0);
  int buf = 0;
  for (int kb = k0; kb < k1; kb += F) {
    const bool more = kb + F < k1;
    if (more) stage(buf ^ 1, kb + F);
    else asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 1;\n" ::: "memory");
    __syncthreads();
    {
      const unsigned tok = pass_token() + (kb == k0 ? 0u : (buf ? (unsigned)(BUF * sizeof(R)) : 0u - (unsigned)(BUF * sizeof(R))));
#pragma unroll
      for (int t = 0; t < 4; t++) soff[t] += tok;
    }
    const int nf = min(F, k1 - kb);
    const bool fast = nf == F && __all_sync(0xffffffffu, complete || !live);
    if (fast) {
      static_for<F>([&](auto fc) {
        constexpr int f = decltype(fc)::value;
        R u[4], v[4];
#pragma unroll
        for (int t = 0; t < 4; t++) lds2_at<R, 2 * f * ESTR * (int)sizeof(R)>(soff[t], u[t], v[t]);
        R G = 0, V = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
          const R ur = c[t] * u[t] - s[t] * v[t];
          const R vr = s[t] * u[t] + c[t] * v[t];
          G = G + w[t] * ur;
          V = V + w[t] * vr;
        }
        const R ur = cr * G - sr * V;
        const R vr = sr * G + cr * V;
        if (live) { __stcs(go, ur / Wall); __stcs(vo, vr / Wall); __stcs(lo, (u8)1); }
        go += a.mo; vo += a.mo; lo += a.mo;
      });
    } else {
#pragma unroll 1
      for (int f = 0; f < nf; f++) {
        const int k = kb + f;
        R og = 0, ov = 0;
        bool good = false;
        if (complete) {
          const unsigned fo = (unsigned)(2 * f * ESTR * sizeof(R));
          R G = 0, V = 0;
#pragma unroll
          for (int t = 0; t < 4; t++) {
            R u, v;
            lds2_at<R, 0>(soff[t] + fo, u, v);
            const R ur = c[t] * u - s[t] * v;
            const R vr = s[t] * u + c[t] * v;
            G = G + w[t] * ur;
            V = V + w[t] * vr;
          }
          const R ur = cr * G - sr * V;
          const R vr = sr * G + cr * V;
          og = ur / Wall; ov = vr / Wall;
          good = true;
        } else if (live) {   // a tap outside the input grid: the reference's loop body from global memory
          const R* __restrict__ gk = a.gi + (size_t)k * mi;
          const R* __restrict__ vk = a.vi + (size_t)k * mi;
          R G = 0, V = 0, W = 0;
#pragma unroll
          for (int t = 0; t < 4; t++) {
            const int p = idx[t];
            if (p > 0) {
              const R u = gk[p - 1], v = vk[p - 1];
              const R ur = c[t] * u - s[t] * v;
              const R vr = s[t] * u + c[t] * v;
              G = G + w[t] * ur;
              V = V + w[t] * vr;
              W = W + w[t];
            }
          }
          good = W >= a.pmp;
          if (good) {
            const R ur = cr * G - sr * V;
            const R vr = sr * G + cr * V;
            og = ur / W; ov = vr / W;
          }
        }
        if (live) {
          __stcs(go, og); __stcs(vo, ov); __stcs(lo, (u8)(good ? 1 : 0));
          if (!good) mark_bad(a.flag, k);
        }
        go += a.mo; vo += a.mo; lo += a.mo;
      }
    }
    __syncthreads();
    buf ^= 1;
  }
}

template <class R> bool gtiles_vec_ok(const Plan<R>& P, const FieldArgs<R>& a) {
  return P.gt.ok && P.vec && P.method == M_BILINEAR && P.gt.slot16 != nullptr && P.gt.novf == 0 && a.li == nullptr && a.vi != nullptr && a.vo != nullptr;
}
template <class R, int NE> void launch_gtiles_vec_inst(const Plan<R>& P, const FieldArgs<R>& a, int kpb, cudaStream_t st) {
  constexpr int ESTR = GT_PTS * NE, F = GT_VEC_F;
  constexpr int SM = 2 * 2 * F * ESTR * (int)sizeof(R);
  static_assert(SM <= 48 * 1024, "fits the default dynamic shared-memory limit");
  const GtTables& G = P.gt;
  dim3 grid(G.geom.ntile, nblk(a.kc, kpb));
  k_gtiles_vec<R, NE><<<grid, GT_PTS, SM, st>>>(a, G.geom, G.estride, G.esrc, G.slot16, P.nxy, P.wxy, P.cxy, P.sxy, P.crot, P.srot, kpb);
}
template <class R> void launch_gtiles_vec(const Plan<R>& P, const FieldArgs<R>& a, cudaStream_t st) {
  static const int kpb_env = getenv("IPB_GKPB") ? atoi(getenv("IPB_GKPB")) : 0;
  const int kpb = std::max(8, std::min(GT_MAXKPB, (std::min(a.kc, kpb_env > 0 ? kpb_env : 64) + 7) / 8 * 8));   // (fields per block: the point's plan data is re-read once per kpb pairs)
  switch (P.gt.estride / GT_PTS) {
    case 1: launch_gtiles_vec_inst<R, 1>(P, a, kpb, st); break;
    case 2: launch_gtiles_vec_inst<R, 2>(P, a, kpb, st); break;
    case 3: launch_gtiles_vec_inst<R, 3>(P, a, kpb, st); break;
    default: launch_gtiles_vec_inst<R, 4>(P, a, kpb, st); break;
  }
  g_launches++;
}

template <class R> bool gtiles_ok(const Plan<R>& P, const FieldArgs<R>& a) {
  if (!P.gt.ok || P.vec) return false;
  if (P.method == M_BICUBIC) return true;
  if (P.method == M_BILINEAR) return a.li != nullptr && a.mspiral <= 0;   // batches without a bitmap stay with k_bilinear_tiles
  if (P.method == M_BUDGET) return budget_collapsed_ok(P, a);
  return false;
}

template <class R, int U, int KIND, int NE, bool MASKED, int FP = GT_F, int MB = 0>
void launch_gtiles_inst(const FieldArgs<R>& a, const GtArgs<R>& T, int kpb, cudaStream_t st) {
  constexpr int SM = 2 * FP * GT_PTS * NE * (int)sizeof(R) + (MASKED ? GT_PTS * NE : 0);
  static_assert(SM <= 48 * 1024, "fits the default dynamic shared-memory limit: no per-device function attribute needed");
  dim3 grid(T.g.ntile, nblk(a.kc, kpb));
  k_gtiles<R, U, KIND, NE, MASKED, FP, MB><<<grid, GT_PTS, SM, st>>>(a, T, kpb);
}
template <class R, int U, int KIND>
void launch_gtiles_u(const FieldArgs<R>& a, const GtArgs<R>& T, int kpb, bool masked, cudaStream_t st) {
  const bool two = T.estride > GT_PTS;
  if constexpr (KIND == GT_STENCIL) {   // bicubic bitmap fields take the general per-point path: no staged bitmap
    // (4 or 2 fields per pass with 5-7 resident blocks, the vector kernel's recipe, lose here: 4.50-7.15 ms against 4.37)
    if (two) launch_gtiles_inst<R, U, KIND, 2, false>(a, T, kpb, st); else launch_gtiles_inst<R, U, KIND, 1, false>(a, T, kpb, st);
  } else {
    if (two) { if (masked) launch_gtiles_inst<R, U, KIND, 2, true>(a, T, kpb, st); else launch_gtiles_inst<R, U, KIND, 2, false>(a, T, kpb, st); }
    else { if (masked) launch_gtiles_inst<R, U, KIND, 1, true>(a, T, kpb, st); else launch_gtiles_inst<R, U, KIND, 1, false>(a, T, kpb, st); }
  }
}

template <class R> void launch_gtiles(const Plan<R>& P, const FieldArgs<R>& a, cudaStream_t st) {
  const GtTables& G = P.gt;
  GtArgs<R> T;
  T.g = G.geom; T.estride = G.estride; T.esrc = G.esrc; T.slot = G.slot;
  T.nc = P.nc; T.wsum = P.cwsum; T.nsamp = P.bo.nsamp; T.bn = P.bn; T.bw = P.bw; T.wbs = P.d_wb;
  // fields per block: the per-point plan data (slots, weights) is re-read once per kpb fields
  static const int kpb_env = getenv("IPB_GKPB") ? atoi(getenv("IPB_GKPB")) : 0;
  const int kpb_want = kpb_env > 0 ? kpb_env : (P.method == M_BUDGET ? 256 : 128);
  const int kpb = std::max(GT_F, std::min(GT_MAXKPB, (std::min(a.kc, kpb_want) + GT_F - 1) / GT_F * GT_F));
  const bool masked = a.li != nullptr;
  if (P.method == M_BICUBIC) { T.taps = P.nxy; T.w = P.wxy; launch_gtiles_u<R, 16, GT_STENCIL>(a, T, kpb, masked, st); }
  else if (P.method == M_BILINEAR) { T.taps = P.nxy; T.w = P.wxy; launch_gtiles_u<R, 4, GT_BILIN>(a, T, kpb, masked, st); }
  else {
    T.taps = P.cn; T.w = P.cw;
    if constexpr (sizeof(R) == 8) {   // collapsed taps exist in the binary64 kinds only
      if (G.U == 4) launch_gtiles_u<R, 4, GT_BUDGET>(a, T, kpb, masked, st);
      else if (G.U == 9) launch_gtiles_u<R, 9, GT_BUDGET>(a, T, kpb, masked, st);
      else launch_gtiles_u<R, 16, GT_BUDGET>(a, T, kpb, masked, st);
    }
  }
  g_launches++;
}

}  // namespace ipb
