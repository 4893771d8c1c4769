// ipb200 — instantiation of the projection, plan and apply kernels for R = float
// (the _4 build kind of the reference, src/CMakeLists.txt:36-76).
#include "ipb_core.cuh"

namespace ipb {
template void run<float>(bool, int, void*, i64, const i64*, const GridHost<float>&, const GridHost<float>&, i64, i64, i64, const i64*,
                  const u8*, const float*, const float*, i64&, float*, float*, float*, float*, i64*, u8*, float*, float*, i64&);
template void run_gdswzd<float>(GridHost<float>, i64, i64, float, float*, float*, float*, float*, i64&, float*, float*, float*, float*, float*, float*, float*);
}  // namespace ipb
