// write-pattern microbenchmark: how fast can B200 absorb (a) one linear stream, (b) km field streams written tile by
// tile as the staged bilinear kernel does, (c) the same with the real, odd field length (alignment classes, 16-byte
// stores that are not 32-byte-sector aligned)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_linear(float4* p, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += st) __stcs(p + i, make_float4(1, 2, 3, 4));
}
// block = 128 thr; tile = 512 floats; each block writes fields [f0, f0+nf) of its tile, go (16B/thread) + lo (4B/thread)
__global__ void k_tiles(float* go, unsigned char* lo, size_t mo, int ntile, int nf, int swap) {
  int tile = swap ? blockIdx.x : blockIdx.y, sl = swap ? blockIdx.y : blockIdx.x;
  size_t n0 = (size_t)tile * 512 + threadIdx.x * 4;
  if (n0 + 3 >= mo) return;
  for (int f = 0; f < nf; f++) {
    size_t k = (size_t)sl * nf + f;
    __stcs(reinterpret_cast<float4*>(go + k * mo + n0), make_float4(1, 2, 3, 4));
    __stcs(reinterpret_cast<unsigned*>(lo + k * mo + n0), 0x01010101u);
  }
}
// the real layout: mo odd, fields of class k mod 4 handled by blocks (class, slice) with the group shift of that class
__global__ void k_classes(float* go, unsigned char* lo, long long mo, int kc, int kpt) {
  const int cls = blockIdx.x & 3, slice = blockIdx.x >> 2, tile = blockIdx.y;
  const int variant = (int)((cls * (mo & 3)) & 3);
  const long long n0 = 4ll * (tile * 128 + threadIdx.x) - variant;
  if (n0 < 0 || n0 + 3 >= mo) return;
  const int jend = min(kpt * (slice + 1), (kc - cls + 3) / 4);
  for (int j = kpt * slice; j < jend; j++) {
    const long long k = cls + 4 * j;
    __stcs(reinterpret_cast<float4*>(go + k * mo + n0), make_float4(1, 2, 3, 4));
    __stcs(reinterpret_cast<unsigned*>(lo + k * mo + n0), 0x01010101u);
  }
}
// as k_classes, but with NC = 8 (or 32) alignment classes: the 512-byte run of a warp starts on a 32-byte sector
// (or 128-byte line) boundary
template <int NC>
__global__ void k_classes_al(float* go, unsigned char* lo, long long mo, int kc, int kpt) {
  const int cls = blockIdx.x % NC, slice = blockIdx.x / NC, tile = blockIdx.y;
  const int variant = (int)((cls * (mo % NC)) % NC);           // element misalignment of field cls modulo NC
  const long long n0 = 4ll * (tile * 128 + threadIdx.x) - variant;   // shift by `variant` elements: warp run starts NC-element aligned
  if (n0 < 0 || n0 + 3 >= mo) return;
  const int jend = min(kpt * (slice + 1), (kc - cls + NC - 1) / NC);
  for (int j = kpt * slice; j < jend; j++) {
    const long long k = cls + (long long)NC * j;
    __stcs(reinterpret_cast<float4*>(go + k * mo + n0), make_float4(1, 2, 3, 4));
    __stcs(reinterpret_cast<unsigned*>(lo + k * mo + n0), 0x01010101u);
  }
}
int main() {
  const size_t mo = 1905152;  // multiple of 4 so every field is aligned
  const int km = 1024, ntile = (int)(mo / 512);
  float* go; unsigned char* lo;
  cudaMalloc(&go, mo * km * 4 + 64); cudaMalloc(&lo, mo * km + 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(e0);
    k_linear<<<148 * 16, 256>>>((float4*)go, mo * km / 4);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("linear 7.8GB: %.3f ms  %.0f GB/s\n", ms, mo * km * 4.0 / ms / 1e6);
    for (int swap = 0; swap < 2; swap++)
      for (int nf : {8, 32, 128, 1024}) {
        dim3 g(km / nf, ntile); if (swap) g = dim3(ntile, km / nf);
        cudaEventRecord(e0);
        k_tiles<<<g, 128>>>(go, lo, mo, ntile, nf, swap);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("tiles swap=%d nf=%4d: %.3f ms  %.0f GB/s\n", swap, nf, ms, mo * km * 5.0 / ms / 1e6);
      }
    const long long mo2 = 1905141;
    for (int kpt : {32, 64, 256}) {
      dim3 g(4 * ((km / 4 + kpt - 1) / kpt), (unsigned)((mo2 / 4 + 2 + 127) / 128));
      cudaEventRecord(e0);
      k_classes<<<g, 128>>>(go, lo, mo2, km, kpt);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      printf("classes (mo=%lld odd) kpt=%3d: %.3f ms  %.0f GB/s\n", mo2, kpt, ms, mo2 * km * 5.0 / ms / 1e6);
    }
    for (int kpt : {16, 32, 64}) {
      dim3 g(8 * ((km / 8 + kpt - 1) / kpt), (unsigned)((mo2 / 4 + 8 + 127) / 128));
      cudaEventRecord(e0);
      k_classes_al<8><<<g, 128>>>(go, lo, mo2, km, kpt);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      printf("8 classes (32B-aligned runs) kpt=%3d: %.3f ms  %.0f GB/s\n", kpt, ms, mo2 * km * 5.0 / ms / 1e6);
    }
    for (int kpt : {8, 16, 32}) {
      dim3 g(32 * ((km / 32 + kpt - 1) / kpt), (unsigned)((mo2 / 4 + 32 + 127) / 128));
      cudaEventRecord(e0);
      k_classes_al<32><<<g, 128>>>(go, lo, mo2, km, kpt);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
      printf("32 classes (128B-aligned runs) kpt=%3d: %.3f ms  %.0f GB/s\n", kpt, ms, mo2 * km * 5.0 / ms / 1e6);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
