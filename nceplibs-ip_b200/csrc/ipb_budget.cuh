// ipb200 — the collapsed budget apply path of the binary64 kinds (ipolates ip=3, _d / _8).
//
// The reference accumulates, per output point and field, NB3 sub-samples of four bilinear taps each
// (budget_interp_mod.F90:252-281): 100 multiply-adds and 100 bitmap tests at the default options, although
// for an output grid finer than the input all those taps fall on the same 4..9 source points.  Summing the
// tap weights per source point first,
//       cw(p) = sum over (sample s, tap t) landing on p of WB(s) * W(s,t),
// turns the field loop into 4..9 taps:  GO = sum_p cw(p) g(p) / sum_p cw(p)  (p running over the unmasked
// points).  That is the same real number as the reference's sum in another association, so it differs from
// it by rounding only: a few 1e-16 relative in binary64, inside the 1e-12 budget BASELINE.json states for
// budget outputs in the _d kind.  It is therefore used for the binary64 kinds only; the _4 kind (4-ulp
// budget, which a re-associated binary32 sum of 100 terms does not meet) keeps the reference's order
// (k_budget_scalar / k_apply_budget), and so does any plan whose points touch more than 16 source points.
//
// lo must stay bit-exact: lo = (weight sum >= threshold).  Without a bitmap the weight sum is a sum of the
// small integers WB and exact in any order.  With a bitmap, a collapsed weight sum within 1e-9 (relative)
// of the threshold is not trusted: that point-field is recomputed in the reference's order and association
// (budget_exact_point), so lo — and go there — are the reference's bits.
#pragma once
#include <cstdlib>

#include "ipb_apply.cuh"

namespace ipb {

constexpr int BUD_UMAX = 16;

// plan time: one thread per output point gathers its distinct source points and their summed weights
template <class R>
__global__ void k_budget_collapse(int no, int nsamp, const int* __restrict__ bn, const R* __restrict__ bw,
                                  const R* __restrict__ wbs, int* __restrict__ cn, R* __restrict__ cw, R* __restrict__ wsum, int* gmax) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= no) return;
  int idx[BUD_UMAX];
  R w[BUD_UMAX];
  int cnt = 0;
  bool over = false;
  R ws = 0;
  for (int s = 0; s < nsamp; s++) {
    const size_t base = (size_t)s * 4 * no + n;
    if (bn[base] <= 0) continue;
    const R wb = wbs[s];
    ws = ws + wb;
#pragma unroll
    for (int t = 0; t < 4; t++) {
      const int p = bn[base + (size_t)t * no];
      const R v = wb * bw[base + (size_t)t * no];
      int u = 0;
      for (; u < cnt; u++) if (idx[u] == p) break;
      if (u == cnt) {
        if (cnt == BUD_UMAX) { over = true; continue; }
        idx[cnt] = p; w[cnt] = 0; cnt++;
      }
      w[u] = w[u] + v;
    }
  }
  // padding: a valid position with weight zero, so that the apply loop needs no tap count
  const int pad = cnt > 0 ? idx[0] : 1;
  for (int u = 0; u < BUD_UMAX; u++) {
    cn[(size_t)u * no + n] = u < cnt ? idx[u] : pad;
    cw[(size_t)u * no + n] = u < cnt ? w[u] : (R)0;
  }
  wsum[n] = ws;
  atomicMax(gmax, over ? BUD_UMAX + 1 : cnt);
}

// One point, one field, in the reference's order and association (bit-exact; the rare path).
template <class R>
__device__ __noinline__ R budget_exact_point(const R* __restrict__ g, const u8* __restrict__ l, int no, int n, bool masked, int nsamp,
                                             const int* __restrict__ bn, const R* __restrict__ bw, const R* __restrict__ wbs, R& wo_out) {
  R acc = 0, wo = 0;
  for (int s = 0; s < nsamp; s++) {
    const size_t base = (size_t)s * 4 * no + n;
    const int i0 = bn[base];
    if (i0 <= 0) continue;
    const int i1 = bn[base + no], i2 = bn[base + 2 * (size_t)no], i3 = bn[base + 3 * (size_t)no];
    const R w0 = bw[base], w1 = bw[base + no], w2 = bw[base + 2 * (size_t)no], w3 = bw[base + 3 * (size_t)no];
    const R wb = wbs[s];
    if (!masked) {
      const R gb = w0 * g[i0 - 1] + w1 * g[i1 - 1] + w2 * g[i2 - 1] + w3 * g[i3 - 1];
      acc = acc + wb * gb; wo = wo + wb;
    } else {
      if (l[i0 - 1]) { acc = acc + wb * w0 * g[i0 - 1]; wo = wo + wb * w0; }
      if (l[i1 - 1]) { acc = acc + wb * w1 * g[i1 - 1]; wo = wo + wb * w1; }
      if (l[i2 - 1]) { acc = acc + wb * w2 * g[i2 - 1]; wo = wo + wb * w2; }
      if (l[i3 - 1]) { acc = acc + wb * w3 * g[i3 - 1]; wo = wo + wb * w3; }
    }
  }
  wo_out = wo;
  return acc;
}

// apply: one output point per lane, a slice of the field batch per thread.  The U collapsed taps of the
// thread's point live in shared memory (column tid of s_off / s_w: conflict-free), not in registers, so that
// enough warps are resident to hide the gather latency.
template <class R, int U>
__global__ void __launch_bounds__(256, 4)
k_budget_collapsed(FieldArgs<R> a, int kpb, int nsamp, const int* __restrict__ cn, const R* __restrict__ cw, const R* __restrict__ wsum,
                   const int* __restrict__ bn, const R* __restrict__ bw, const R* __restrict__ wbs) {
  __shared__ int s_off[U][256];
  __shared__ R s_w[U][256];
  const int tid = threadIdx.x;
  const int n = blockIdx.x * blockDim.x + tid;
  if (n >= a.no) return;
#pragma unroll
  for (int u = 0; u < U; u++) { s_off[u][tid] = cn[(size_t)u * a.no + n] - 1; s_w[u][tid] = cw[(size_t)u * a.no + n]; }
  const R wplain = wsum[n];
  const R rwplain = wplain > 0 ? (R)1 / wplain : (R)0;   // the collapsed path is held to 1e-12, not to the last bit: multiply instead of divide
  const int k0 = blockIdx.y * kpb, k1 = min(k0 + kpb, a.kc);
  const int no = a.no;
  const R pmp = a.pmp;
  const R* __restrict__ g = a.gi + (size_t)k0 * a.mi;
  const u8* __restrict__ l = a.li ? a.li + (size_t)k0 * a.mi : nullptr;
  R* go = a.go + (size_t)k0 * a.mo + n;
  u8* lo = a.lo + (size_t)k0 * a.mo + n;
#pragma unroll 1
  for (int k = k0; k < k1; k++) {
    const bool masked = a.ibi[k] != 0;
    R acc = 0, wo;
    R v[U];
#pragma unroll
    for (int u = 0; u < U; u++) v[u] = __ldg(g + s_off[u][tid]);
    unsigned mb = (1u << U) - 1u;                      // bit u: tap u takes part
    if (masked) {
      mb = 0;
#pragma unroll
      for (int u = 0; u < U; u++) mb |= (__ldg(l + s_off[u][tid]) ? 1u : 0u) << u;
    }
    // bitmaps are spatially coherent: most warps see all of their taps unmasked and take the plain path
    // (a field WITH a bitmap accumulates WB*W per tap in floating point even when every tap is usable: next to the
    // threshold — mp = 100 — that sum, not the integer sum of the sample weights, decides lo: exact path below)
    const R dpl = wplain - pmp;
    const bool near_plain = masked && (dpl < 0 ? -dpl : dpl) <= pmp * (R)1e-9;
    if (__all_sync(__activemask(), mb == (1u << U) - 1u && !near_plain)) {
#pragma unroll
      for (int u = 0; u < U; u++) acc = acc + s_w[u][tid] * v[u];
      const bool good = wplain >= pmp;
      __stcs(go, good ? acc * rwplain : (R)0);
      __stcs(lo, (u8)(good ? 1 : 0));
      if (!good) mark_bad(a.flag, k);
      g += a.mi; if (l) l += a.mi; go += a.mo; lo += a.mo;
      continue;
    } else {
      wo = 0;
#pragma unroll
      for (int u = 0; u < U; u++) {
        const R wu = s_w[u][tid];
        if ((mb >> u) & 1u) { acc = acc + wu * v[u]; wo = wo + wu; }
      }
      const R d = wo - pmp;
      if (near_plain || (d < 0 ? -d : d) <= pmp * (R)1e-9) acc = budget_exact_point<R>(g, l, no, n, masked, nsamp, bn, bw, wbs, wo);
    }
    const bool good = wo >= pmp;
    __stcs(go, good ? acc / wo : (R)0);
    __stcs(lo, (u8)(good ? 1 : 0));
    if (!good) mark_bad(a.flag, k);
    g += a.mi; if (l) l += a.mi; go += a.mo; lo += a.mo;
  }
}

template <class R> void build_budget_collapse(Plan<R>& P, cudaStream_t st) {
  P.cu = 0;
  if (sizeof(R) != 8 || P.method != M_BUDGET || P.vec || P.no <= 0 || P.bo.nsamp <= 0 || !P.bn || !P.bw) return;
  const int no = P.no;
  P.cn = dalloc<int>((size_t)BUD_UMAX * no);
  P.cw = dalloc<R>((size_t)BUD_UMAX * no);
  P.cwsum = dalloc<R>(no);
  int* d_max = dalloc<int>(1);
  IPB_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int), st));
  k_budget_collapse<R><<<nblk(no, 128), 128, 0, st>>>(no, P.bo.nsamp, P.bn, P.bw, P.d_wb, P.cn, P.cw, P.cwsum, d_max);
  g_launches++;
  int h = 0;
  IPB_CUDA(cudaMemcpyAsync(&h, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
  IPB_CUDA(cudaStreamSynchronize(st));
  dfree(d_max);
  if (h <= 0 || h > BUD_UMAX) { dfree(P.cn); dfree(P.cw); dfree(P.cwsum); return; }
  P.cu = h <= 4 ? 4 : (h <= 9 ? 9 : BUD_UMAX);
  P.bytes += (size_t)BUD_UMAX * no * (4 + sizeof(R)) + (size_t)no * sizeof(R);
}

template <class R> bool budget_collapsed_ok(const Plan<R>& P, const FieldArgs<R>& a) {
  return P.cu > 0 && P.method == M_BUDGET && !P.vec && a.mspiral <= 1;
}

template <class R> void launch_budget_collapsed(const Plan<R>& P, const FieldArgs<R>& a, cudaStream_t st) {
  static const int kpb_env = getenv("IPB_BKPB") ? atoi(getenv("IPB_BKPB")) : 256;
  const int kpb = std::max(1, std::min(a.kc, kpb_env));   // fields per thread: the collapsed taps are re-read once per kpb fields
  dim3 grid(nblk(a.no, 256), nblk(a.kc, kpb));
  const int ns = P.bo.nsamp;
  if (P.cu == 4) k_budget_collapsed<R, 4><<<grid, 256, 0, st>>>(a, kpb, ns, P.cn, P.cw, P.cwsum, P.bn, P.bw, P.d_wb);
  else if (P.cu == 9) k_budget_collapsed<R, 9><<<grid, 256, 0, st>>>(a, kpb, ns, P.cn, P.cw, P.cwsum, P.bn, P.bw, P.d_wb);
  else k_budget_collapsed<R, BUD_UMAX><<<grid, 256, 0, st>>>(a, kpb, ns, P.cn, P.cw, P.cwsum, P.bn, P.bw, P.d_wb);
  g_launches++;
}

}  // namespace ipb
