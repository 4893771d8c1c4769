// Store-pattern microbenchmark for the TMA-staged bilinear kernel (csrc/ipb_box.cuh): no arithmetic, no loads — only the
// go / lo store stream of the headline configuration (mo = 1905141 points in rows of 1799, km = 1024 fields), written
// patch by patch exactly as the kernel does, with the knobs that shape the stream:
//   unit   alignment unit of the 32-point runs in elements (4 / 8 / 16 / 32 -> 16 / 32 / 64 / 128 bytes)
//   kpt    fields of one class walked by a warp
//   W      warps per block (each warp owns one patch of NR rows x 32 points)
//   ppw    patches a warp visits per pass of 8 fields before it moves to the next 8 fields (page locality)
//   hint   0: st.global.cs (evict first)   1: plain st.global   2: st.global.cg
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/wr_bench2 tools/wr_bench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

template <int HINT> __device__ __forceinline__ void st32(float* p, float v) {
  if (HINT == 0) __stcs(p, v); else if (HINT == 1) *p = v; else __stcg(p, v);
}
template <int HINT> __device__ __forceinline__ void st32u(unsigned* p, unsigned v) {
  if (HINT == 0) __stcs(p, v); else if (HINT == 1) *p = v; else __stcg(p, v);
}

template <int NR, int HINT>
__global__ void k_rows(float* go, unsigned char* lo, long long mo, int rowlen, int nrow, int amask, int period, int kc, int kpt, int ppw,
                       int ntx, int nty, int flat, int sync, int swap, int remap, int nolo) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  const unsigned bx = swap ? blockIdx.y : blockIdx.x, by = swap ? blockIdx.x : blockIdx.y;
  const int cls = bx % period, slice = bx / period;
  const int jbeg = kpt * slice, jend = min(kpt * (slice + 1), (kc - cls + period - 1) / period);
  if (jbeg >= jend) return;
  const int phase = (int)(((long long)cls * mo) & amask);
  const size_t ostep = (size_t)period * mo;
  for (int j0 = jbeg; j0 < jend; j0 += 8) {
    for (int p = 0; p < ppw; p++) {
      const int tile = (by * ppw + p) * W + warp;
      const bool inside = tile < ntx * nty;
      const int tx = tile % ntx, ty = tile / ntx;
      int npt[NR];
      bool whole = true;
#pragma unroll
      for (int r = 0; r < NR; r++) {
        const int SX = flat, RY = NR / SX;   // flat = segments per row
        const int j = ty * RY + r / SX;
        const int shift = (int)(((long long)phase + (long long)j * rowlen) & amask);
        const int i = (tx * SX + r % SX) * 32 - shift + lane;
        const bool in = inside && j < nrow && i >= 0 && i < rowlen;
        npt[r] = in ? j * rowlen + i : -1;
        whole = whole && in;
      }
      const bool act = __all_sync(0xffffffffu, whole);
      const int r8 = lane >> 3;
      int base = -1, base2 = -1;
#pragma unroll
      for (int r = 0; r < NR; r++) { const int b = __shfl_sync(0xffffffffu, npt[r], 0); if (r == r8) base = b; if (r == r8 + 4) base2 = b; }
      const long long lo_off = base >= 0 ? base + 4 * (lane & 7) : -1, lo_off2 = base2 >= 0 ? base2 + 4 * (lane & 7) : -1;
      float* gp = go + (size_t)(cls + period * j0) * mo;
      unsigned char* lp = lo + (size_t)(cls + period * j0) * mo;
      if (sync) __syncthreads();
      if (remap) {   // diagnostic layout: the fields of a 128-point block are neighbours in memory (same 2 MB page)
        for (int f = 0; act && f < 8 && j0 + f < jend; f++) {
          const size_t k = (size_t)(cls + period * (j0 + f));
#pragma unroll
          for (int r = 0; r < NR; r++) st32<HINT>(go + ((size_t)(npt[r] >> 7) * kc + k) * 128 + (npt[r] & 127), (float)f);
          if (lo_off >= 0) st32u<HINT>(reinterpret_cast<unsigned*>(lo + ((size_t)(lo_off >> 7) * kc + k) * 128 + (lo_off & 124)), 0x01010101u);
        }
        continue;
      }
      for (int f = 0; act && f < 8 && j0 + f < jend; f++) {
#pragma unroll
        for (int r = 0; r < NR; r++) st32<HINT>(gp + npt[r], (float)f);
        if (!nolo && lo_off >= 0) st32u<HINT>(reinterpret_cast<unsigned*>(lp + lo_off), 0x01010101u);
        if (!nolo && NR > 4 && lo_off2 >= 0) st32u<HINT>(reinterpret_cast<unsigned*>(lp + lo_off2), 0x01010101u);
        gp += ostep; lp += ostep;
      }
    }
  }
}

int main(int argc, char** argv) {
  const long long mo = 1905141; const int rowlen = 1799, nrow = 1059, km = 1024;
  float* go; unsigned char* lo;
  cudaMalloc(&go, (size_t)mo * km * 4 + 256); cudaMalloc(&lo, (size_t)mo * km + 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto run = [&](int unit, int kpt, int W, int ppw, int hint, int NR, int flat = 0, int sync = 0, int swap = 0, int remap = 0, int nolo = 0) {
    const int amask = unit - 1;
    int s = (int)(mo & amask), g = unit; while (s) { int t = g % s; g = s; s = t; }
    const int period = unit / g;
    if (flat == 0) flat = 1; else if (flat == 1) flat = NR;
    const int SXh = flat, RYh = NR / SXh;
    const int ntx = (rowlen + amask + 32 * SXh - 1) / (32 * SXh), nty = (nrow + RYh - 1) / RYh;
    const int per_class = (km + period - 1) / period;
    dim3 grid(period * ((per_class + kpt - 1) / kpt), (ntx * nty + W * ppw - 1) / (W * ppw));
    if (swap) grid = dim3(grid.y, grid.x);
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
#define L(NRR, H) k_rows<NRR, H><<<grid, 32 * W>>>(go, lo, mo, rowlen, nrow, amask, period, km, kpt, ppw, ntx, nty, flat, sync, swap, remap, nolo)
      if (NR == 4) { if (hint == 0) L(4, 0); else if (hint == 1) L(4, 1); else L(4, 2); }
      else { if (hint == 0) L(8, 0); else if (hint == 1) L(8, 1); else L(8, 2); }
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("nolo=%d remap=%d swap=%d sync=%d flat=%d unit=%2d(%3dB) period=%2d kpt=%4d W=%2d ppw=%2d hint=%d NR=%d grid=(%u,%u): %.3f ms  %.0f GB/s  %s\n", nolo, remap, swap, sync, flat, unit, unit * 4, period, kpt, W, ppw, hint,
           NR, grid.x, grid.y, best, mo * km * (nolo ? 4.0 : 5.0) / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  if (argc > 6) {   // seventh series: go only (lo written by a separate linear kernel)
    for (int nl : {1, 0}) for (int kpt : {8, 32, 64}) for (int W : {4, 8}) run(32, kpt, W, 1, 0, 4, 0, 0, 0, 0, nl);
    for (int kpt : {32}) for (int W : {4, 8}) { run(32, kpt, W, 1, 0, 4, 0, 1, 0, 0, 1); run(32, kpt, W, 1, 0, 4, 2, 0, 0, 0, 1); run(32, kpt, W, 1, 0, 4, 1, 0, 0, 0, 1); run(32, kpt, W, 1, 0, 8, 0, 0, 0, 0, 1); run(32, kpt, W, 4, 0, 4, 0, 0, 0, 0, 1); run(32, kpt, W, 1, 0, 4, 0, 0, 1, 0, 1);}
    run(16, 64, 4, 1, 0, 4, 0, 0, 0, 0, 1); run(8, 128, 4, 1, 0, 4, 0, 0, 0, 0, 1);
    return 0;
  }
  if (argc > 5) {   // sixth series: page locality diagnostic
    for (int rm : {1, 0}) for (int kpt : {8, 32, 64}) run(32, kpt, 4, 1, 0, 4, 0, 0, 0, rm);
    for (int rm : {1, 0}) run(4, 256, 4, 1, 0, 4, 0, 0, 0, rm);
    for (int rm : {1, 0}) run(32, 32, 4, 8, 0, 4, 0, 0, 0, rm);
    return 0;
  }
  if (argc > 4) {   // fifth series: patches fastest in the block order
    for (int sw : {1, 0}) for (int kpt : {8, 32, 64}) for (int W : {4, 8}) run(32, kpt, W, 1, 0, 4, 0, 0, sw);
    for (int kpt : {32, 64, 128}) run(16, kpt, 4, 1, 0, 4, 0, 0, 1);
    for (int kpt : {32, 256}) run(4, kpt, 4, 1, 0, 4, 0, 0, 1);
    run(32, 32, 4, 1, 0, 4, 0, 1, 1); run(32, 32, 4, 4, 0, 4, 0, 0, 1); run(32, 32, 4, 1, 0, 4, 1, 0, 1); run(32, 32, 4, 1, 0, 8, 0, 0, 1);
    return 0;
  }
  if (argc > 3) {   // fourth series: patch shapes RY x SX
    for (int sy : {1, 0}) for (int W : {4, 8, 16}) { run(32, 32, W, 1, 0, 4, 2, sy); }
    for (int W : {4, 8}) { run(32, 32, W, 1, 0, 8, 2, 1); run(32, 32, W, 1, 0, 8, 4, 1); }
    for (int W : {8, 16}) run(32, 32, W, 1, 0, 4, 0, 1);
    for (int W : {4, 8}) { run(16, 64, W, 1, 0, 4, 2, 1); run(16, 32, W, 1, 0, 4, 2, 1); }
    return 0;
  }
  if (argc > 2) {   // third series: block-synchronous passes
    for (int sy : {0, 1}) for (int W : {4, 8}) for (int kpt : {8, 32, 64}) run(32, kpt, W, 1, 0, 4, 1, sy);
    for (int sy : {0, 1}) for (int W : {4, 8, 16}) run(32, 32, W, 1, 0, 4, 0, sy);
    for (int sy : {0, 1}) for (int W : {4, 8}) run(16, 64, W, 1, 0, 4, 1, sy);
    for (int sy : {0, 1}) run(32, 32, 4, 1, 0, 8, 1, sy);
    return 0;
  }
  if (argc > 1) {   // second series: contiguity per block
    for (int flat : {1, 0}) for (int W : {2, 4, 8, 16}) for (int kpt : {8, 32}) run(32, kpt, W, 1, 0, 4, flat);
    for (int W : {4, 8}) run(32, 32, W, 1, 0, 8, 1);
    for (int W : {4, 8, 16}) run(16, 64, W, 1, 0, 4, 1);
    for (int W : {4, 8, 16}) run(8, 128, W, 1, 0, 4, 1);
    for (int W : {12, 20, 24}) run(32, 32, W, 1, 0, 4, 0);
    return 0;
  }
  // alignment x fields per warp
  for (int unit : {32, 16, 8, 4}) for (int kpt : {8, 32, 64, 256}) if (kpt * unit <= 1024 * 4 && kpt <= 1024 / (unit > 32 ? 32 : 1)) run(unit, kpt, 4, 1, 0, 4);
  // warps per block
  for (int W : {1, 2, 8, 16, 32}) run(32, 32, W, 1, 0, 4);
  for (int W : {8, 16}) run(16, 64, W, 1, 0, 4);
  // patches per pass (page locality): a warp revisits ppw patches for every 8 fields
  for (int ppw : {2, 4, 8, 16}) run(32, 32, 4, ppw, 0, 4);
  for (int ppw : {4, 16}) run(16, 64, 4, ppw, 0, 4);
  // store hints
  for (int h : {1, 2}) { run(32, 32, 4, 1, h, 4); run(16, 64, 4, 1, h, 4); }
  // rows per patch
  run(32, 32, 4, 1, 0, 8); run(16, 64, 4, 1, 0, 8);
  return 0;
}
