// write-pattern microbenchmark: how fast can B200 absorb (a) one linear stream, (b) km field streams written tile by tile
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
int main() {
  const size_t mo = 1905152;  // multiple of 4 so every field is aligned
  const int km = 1024, ntile = (int)(mo / 512);
  float* go; unsigned char* lo;
  cudaMalloc(&go, mo * km * 4); cudaMalloc(&lo, mo * km);
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
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
