// Probe for the TMA box load used by csrc/ipb_box.cuh: one warp fetches F boxes (BX x BY x 1) of a 3-D tensor
// (x, y, field) into shared memory behind an mbarrier and copies them out for checking.
//   mode 0: lane 0 issues every box (canonical elected-thread form)
//   mode 1: lane f issues box f
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tma_probe tools/tma_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int BX = 16, BY = 4, F = 8, BOXB = BX * BY * 4;

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@!p bra WAIT_%=;\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_box_3d(unsigned dst, const CUtensorMap* tm, int x, int y, int k, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(k), "r"(bar) : "memory");
}

__global__ void k_probe(const __grid_constant__ CUtensorMap tmap, int mode, int x0, int y0, int k0, float* out, int nstr) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31;
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
  const unsigned bar = base + F * BOXB;
  if (lane == 0) mbar_init(bar, 1);
  if (mode & 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  else asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  if (mode & 4) { if (lane == 0) out[0] = 1.f; return; }   // no TMA at all
  if (mode & 8) {
    if (lane == 0) { mbar_expect_tx(bar, nstr * BOXB); tma_box_3d(base, &tmap, x0, y0, k0, bar); }
  } else if ((mode & 1) == 0) {
    if (lane == 0) {
      mbar_expect_tx(bar, F * BOXB);
      for (int f = 0; f < F; f++) tma_box_3d(base + f * BOXB, &tmap, x0, y0, k0 + f, bar);
    }
  } else {
    if (lane == 0) mbar_expect_tx(bar, F * BOXB);
    __syncwarp();
    if (lane < F) tma_box_3d(base + lane * BOXB, &tmap, x0, y0, k0 + lane, bar);
  }
  __syncwarp();
  mbar_wait(bar, 0);
  const float* s = reinterpret_cast<const float*>(sm);
  for (int i = lane; i < F * BX * BY; i += 32) out[i] = s[i];
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  const int im = 360, jm = 180, km = 40;
  const int S = argc > 3 ? atoi(argv[3]) : 1, NS = argc > 4 ? atoi(argv[4]) : F;   // mode 8: element stride and number of fields of the strided box
  std::vector<float> h((size_t)im * jm * km);
  for (size_t i = 0; i < h.size(); i++) h[i] = (float)(i % 100003);
  float *d, *out;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, F * BX * BY * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encoder\n"); return 2; }
  typedef CUresult (*ENC)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                          CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  cuuint64_t dims[3] = {im, jm, km}, strides[2] = {(cuuint64_t)im * 4, (cuuint64_t)im * jm * 4};
  cuuint32_t box[3] = {BX, BY, 1}, es[3] = {1, 1, 1};
  if (mode & 8) { box[2] = (cuuint32_t)(NS * S); es[2] = (cuuint32_t)S; }
  CUresult rc = ((ENC)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("mode %d encode rc=%d\n", mode, (int)rc);
  const int x0 = argc > 2 ? atoi(argv[2]) : 351, y0 = 77, k0 = 3;   // 351: the box hangs over the right edge: zero fill
  k_probe<<<1, 32, F * BOXB + 64>>>(tm, mode, x0, y0, k0, out, NS);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<float> r(F * BX * BY);
  cudaMemcpy(r.data(), out, r.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int f = 0; f < F; f++) for (int y = 0; y < BY; y++) for (int x = 0; x < BX; x++) {
    const int gx = x0 + x, gy = y0 + y;
    const int kf = (mode & 8) ? k0 + f * S : k0 + f;
    if ((mode & 8) && f >= NS) continue;
    const float want = (gx < im && kf < km) ? h[((size_t)kf * jm + gy) * im + gx] : 0.f;
    if (r[(f * BY + y) * BX + x] != want) bad++;
  }
  printf("mismatches: %d\n", bad);
  return bad != 0;
}
