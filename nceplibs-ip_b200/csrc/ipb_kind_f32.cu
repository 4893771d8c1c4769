// ipb200 — instantiation of the projection, plan and apply kernels for R = float
// (the _4 build kind of the reference, src/CMakeLists.txt:36-76).
#include "ipb_core.cuh"

namespace ipb {
template void run<float>(bool, int, void*, i64, const i64*, const GridHost<float>&, const GridHost<float>&, i64, i64, i64, const i64*,
                  const u8*, const float*, const float*, i64&, float*, float*, float*, float*, i64*, u8*, float*, float*, i64&);
template void run_gdswzd<float>(GridHost<float>, i64, i64, float, float*, float*, float*, float*, i64&, float*, float*, float*, float*, float*, float*, float*);
template void run_xwafs<float>(int, bool, i64, const i64*, i64, i64, i64, i64, const i64*, const float*, const u8*, float*, u8*, i64*);
template PlanHandle* plan_get<float>(bool, i64, const i64*, const GridHost<float>&, const GridHost<float>&, i64, i64, i64&, i64&);
template int plan_geometry<float>(PlanHandle*, void*, float*, float*, float*, float*);
template int plan_apply<float>(PlanHandle*, void*, i64, const i64*, const u8*, const float*, const float*, int*, u8*, float*, float*);
}  // namespace ipb

// Host-side view of the point-staged kernels' tile geometry (ipb_gtiles.cuh), for the CPU tests: the output point of
// (tile, warp row r, lane) for a tile shape, -1 where the tile hangs over the grid; and the number of tiles.
extern "C" int ipgpu_tile_point(int mode, int rowlen, int no, int tile, int r, int lane) {
  return ipb::gt_point(ipb::gt_geom(mode, rowlen, no), tile, r, lane, no);
}
extern "C" int ipgpu_tile_count(int mode, int rowlen, int no) { return ipb::gt_geom(mode, rowlen, no).ntile; }
