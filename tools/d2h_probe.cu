// d2h_probe — where does the device->host rate of a 10 GB result go?  Times D2H copies into pinned memory by size,
// by allocator (cudaHostAlloc | mmap + cudaHostRegister, with and without transparent huge pages) and by stream pattern.
//   nvcc -O2 -o d2h_probe tools/d2h_probe.cu && ./d2h_probe
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

static double timed_copy(void* dst, const void* src, size_t n, cudaStream_t s) {
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  CK(cudaEventRecord(a, s)); CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, s)); CK(cudaEventRecord(b, s));
  CK(cudaEventSynchronize(b)); float ms; CK(cudaEventElapsedTime(&ms, a, b));
  cudaEventDestroy(a); cudaEventDestroy(b);
  return n / (ms * 1e6);
}

int main() {
  const size_t G = (size_t)1 << 30, total = 6 * G;
  char* d; CK(cudaMalloc(&d, total)); CK(cudaMemset(d, 1, total));
  cudaStream_t s[2]; CK(cudaStreamCreate(&s[0])); CK(cudaStreamCreate(&s[1]));
  for (int mode = 0; mode < 3; mode++) {
    char* h = nullptr;
    if (mode == 0) CK(cudaHostAlloc((void**)&h, total, cudaHostAllocPortable));
    else {
      h = (char*)mmap(nullptr, total, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
      if (h == MAP_FAILED) { printf("mmap failed\n"); return 1; }
      if (mode == 2) printf("madvise(MADV_HUGEPAGE) -> %d\n", madvise(h, total, MADV_HUGEPAGE));
      memset(h, 0, total);
      CK(cudaHostRegister(h, total, cudaHostRegisterPortable));
    }
    const char* name = mode == 0 ? "cudaHostAlloc" : (mode == 1 ? "mmap+register" : "mmap+THP+register");
    timed_copy(h, d, 256 << 20, s[0]);
    double r256 = 0; for (int i = 0; i < 4; i++) r256 = timed_copy(h, d, 256 << 20, s[0]);
    const double rbig = timed_copy(h, d, total, s[0]);
    const double rbig2 = timed_copy(h, d, total, s[0]);
    // the library's pattern: chunks of 448 MB + 112 MB alternating on two streams
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a, s[0]));
    size_t off = 0; int ci = 0;
    while (off + (560u << 20) <= total) {
      CK(cudaMemcpyAsync(h + off, d + off, 448u << 20, cudaMemcpyDeviceToHost, s[ci])); off += 448u << 20;
      CK(cudaMemcpyAsync(h + off, d + off, 112u << 20, cudaMemcpyDeviceToHost, s[ci])); off += 112u << 20;
      ci ^= 1;
    }
    CK(cudaStreamSynchronize(s[1]));
    CK(cudaEventRecord(b, s[0])); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    printf("%-20s 256 MiB repeated %.1f GB/s | 6 GiB one copy %.1f, again %.1f | two-stream chunks %.1f GB/s\n", name, r256, rbig, rbig2, off / (ms * 1e6));
    if (mode == 0) CK(cudaFreeHost(h)); else { CK(cudaHostUnregister(h)); munmap(h, total); }
  }
  return 0;
}
