// ipb200 — vectorised bilinear apply path (placeholder until the first parity run is green).
#pragma once
#include "ipb_apply.cuh"

namespace ipb {
template <class R> bool fast_bilinear_ok(const Plan<R>&, const FieldArgs<R>&) { return false; }
template <class R> void launch_fast_bilinear(const Plan<R>&, const FieldArgs<R>&, cudaStream_t) {}
}  // namespace ipb
