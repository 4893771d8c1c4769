// ipb200 — mbarrier / TMA (cp.async.bulk.tensor) helpers shared by the TMA-staged kernels (ipb_box.cuh, ipb_gtiles.cuh)
// and the host-side tensor-map encoder (driver entry point fetched through the runtime: no link against libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace ipb {

// ---- device helpers: mbarrier + TMA ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@!p bra WAIT_%=;\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}
// one BX x BY x 1 box of the tensor (x, y, field) -> shared memory, completion bytes on `bar`
__device__ __forceinline__ void tma_box_3d(unsigned dst, const CUtensorMap* tm, int x, int y, int k, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
               "l"(tm), "r"(x), "r"(y), "r"(k), "r"(bar)
               : "memory");
}
// one BX x BY x 1 x F box of the 4-D tensor (x, y, class, field-of-class)
__device__ __forceinline__ void tma_box_4d(unsigned dst, const CUtensorMap* tm, int x, int y, int c, int j, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
               "l"(tm), "r"(x), "r"(y), "r"(c), "r"(j), "r"(bar)
               : "memory");
}
typedef CUresult (*PFN_tmap_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline PFN_tmap_encode tmap_encoder() {
  static PFN_tmap_encode fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (PFN_tmap_encode)p;
  }();
  return fn;
}

}  // namespace ipb
