// ipb200 — instantiation of the projection, plan and apply kernels for R = double
// (the _d and _8 build kinds of the reference, src/CMakeLists.txt:36-76).
#include "ipb_core.cuh"

namespace ipb {
template void run<double>(bool, int, void*, i64, const i64*, const GridHost<double>&, const GridHost<double>&, i64, i64, i64, const i64*,
                  const u8*, const double*, const double*, i64&, double*, double*, double*, double*, i64*, u8*, double*, double*, i64&);
template void run_gdswzd<double>(GridHost<double>, i64, i64, double, double*, double*, double*, double*, i64&, double*, double*, double*, double*, double*, double*, double*);
template void run_xwafs<double>(int, bool, i64, const i64*, i64, i64, i64, i64, const i64*, const double*, const u8*, double*, u8*, i64*);
template PlanHandle* plan_get<double>(bool, i64, const i64*, const GridHost<double>&, const GridHost<double>&, i64, i64, i64&, i64&);
template int plan_geometry<double>(PlanHandle*, void*, double*, double*, double*, double*);
template int plan_apply<double>(PlanHandle*, void*, i64, const i64*, const u8*, const double*, const double*, int*, u8*, double*, double*);
}  // namespace ipb
