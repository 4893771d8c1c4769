// ipb200 — the WAFS thin/full grid expander of the reference (src/ipxwafs.F90:81-225, ipxwafs2.F90:92-253,
// ipxwafs3.F90:95-274) on the device: one thread per output point and field.  The grids are tiny (3447 <-> 5329 points);
// the point of having it here is that a caller who keeps its fields on the GPU side of the library does not have to take
// them back to the host for this step.  Arithmetic: the reference's, in the kind's type, no FMA.
#pragma once
#include <cuda_runtime.h>

#include "ipb_plan.cuh"
#include "ipb_state.h"

namespace ipb {

struct WafsRows { int start[74]; int npts[73]; };   // first point and length of every row of the thinned grid

// variant 1: linear along the row, no bitmaps; 2: linear, both neighbours must be valid; 3: nearest neighbour along the row
template <class R>
__global__ void k_xwafs(int variant, int expand, int km, WafsRows rows, int jm, int imf, int nthin, int nfull, const int* __restrict__ ib_in,
                        const R* __restrict__ din, const u8* __restrict__ bin, R* __restrict__ dout, u8* __restrict__ bout, int* ib_out) {
  const int nout = expand ? nfull : nthin, nin = expand ? nthin : nfull;
  const int po = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
  if (po >= nout || k >= km) return;
  int j, i;
  if (expand) { j = po / imf; i = po - j * imf + 1; }
  else {
    int lo = 0, hi = jm - 1;                       // row of the thinned grid that holds point po
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (rows.start[mid] <= po) lo = mid; else hi = mid - 1; }
    j = lo; i = po - rows.start[j] + 1;
  }
  if (j >= jm) return;
  const int im1 = rows.npts[j], im2 = imf;
  const int n_in = expand ? im1 : im2, n_out = expand ? im2 : im1;
  const int s_in = expand ? rows.start[j] : j * imf;
  const R rat = (R)(n_in - 1) / (R)(n_out - 1);                  // RAT1 / RAT2
  const R x = (R)(i - 1) * rat + (R)1;
  int ia = (int)x;
  ia = min(max(ia, 1), n_in - 1);
  const int ib = ia + 1;
  const R wa = (R)ib - x, wb = x - (R)ia;
  const size_t pa = (size_t)k * nin + s_in + ia - 1, pb = pa + 1, o = (size_t)k * nout + po;
  if (variant == 1) { dout[o] = wa * din[pa] + wb * din[pb]; return; }
  const bool masked = ib_in[k] != 0;
  bool ok;
  R v;
  if (variant == 2) { ok = !masked || (bin[pa] && bin[pb]); v = ok ? wa * din[pa] + wb * din[pb] : (R)0; }
  else { const size_t pn = wa >= wb ? pa : pb; ok = !masked || bin[pn]; v = ok ? din[pn] : (R)0; }
  dout[o] = v;
  bout[o] = ok ? 1 : 0;
  if (!ok && ib_out[k] == 0) ib_out[k] = 1;
}

// host arrays in, host arrays out (the direction decides which of thin / full is the input)
template <class R>
void run_xwafs(int variant, bool expand, i64 km, const i64* opt, i64 jm, i64 imf, i64 nthin, i64 nfull, const i64* ib_in, const R* din, const u8* bin,
               R* dout, u8* bout, i64* ib_out) {
  if (km <= 0) return;
  Runtime& RT = rt();
  const int device = RT.current_device();
  DevState& DS = RT.state(device);
  std::lock_guard<std::mutex> lock(DS.mu);
  cudaStream_t st = DS.stream[0];
  AsyncAlloc scope(st);
  WafsRows rows;
  int s = 0;
  for (int j = 0; j < 73; j++) { rows.start[j] = s; rows.npts[j] = j < jm ? (int)opt[j] : 0; s += rows.npts[j]; }
  rows.start[73] = s;
  const i64 nin = expand ? nthin : nfull, nout = expand ? nfull : nthin;
  R* d_in = dalloc<R>((size_t)km * nin);
  R* d_out = dalloc<R>((size_t)km * nout);
  u8 *b_in = nullptr, *b_out = nullptr;
  int* d_ib = dalloc<int>((size_t)2 * km);
  std::vector<int> ib32((size_t)2 * km, 0);
  for (i64 k = 0; k < km; k++) ib32[k] = (variant != 1 && ib_in[k] != 0) ? 1 : 0;
  IPB_CUDA(cudaMemcpyAsync(d_in, din, (size_t)km * nin * sizeof(R), cudaMemcpyHostToDevice, st));
  IPB_CUDA(cudaMemcpyAsync(d_ib, ib32.data(), (size_t)2 * km * sizeof(int), cudaMemcpyHostToDevice, st));
  if (variant != 1) {
    b_in = dalloc<u8>((size_t)km * nin);
    b_out = dalloc<u8>((size_t)km * nout);
    IPB_CUDA(cudaMemcpyAsync(b_in, bin, (size_t)km * nin, cudaMemcpyHostToDevice, st));
  }
  k_xwafs<R><<<dim3(nblk(nout, 256), (unsigned)km), 256, 0, st>>>(variant, expand ? 1 : 0, (int)km, rows, (int)jm, (int)imf, (int)nthin, (int)nfull, d_ib, d_in,
                                                                b_in, d_out, b_out, d_ib + km);
  g_launches++;
  IPB_CUDA(cudaMemcpyAsync(dout, d_out, (size_t)km * nout * sizeof(R), cudaMemcpyDeviceToHost, st));
  if (variant != 1) {
    IPB_CUDA(cudaMemcpyAsync(bout, b_out, (size_t)km * nout, cudaMemcpyDeviceToHost, st));
    IPB_CUDA(cudaMemcpyAsync(ib32.data() + km, d_ib + km, (size_t)km * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  IPB_CUDA(cudaStreamSynchronize(st));
  if (variant != 1) for (i64 k = 0; k < km; k++) ib_out[k] = ib32[km + k];
  dfree(d_in); dfree(d_out); dfree(d_ib);
  if (b_in) dfree(b_in);
  if (b_out) dfree(b_out);
}

}  // namespace ipb
