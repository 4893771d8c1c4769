#!/bin/bash
cd "$(dirname "$0")/.."
run() { local label="$1"; shift; local out
  out=$(env "$@" python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-secondary $EXTRA 2>/dev/null | tail -1)
  echo "$label $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("ms_per_step=%.4f kernel_ms=%.4f frac=%.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))' 2>/dev/null || echo "FAILED: $out")"; }
EXTRA="$*"
for shape in 41 22; do for w in 4 8 16; do for sy in 1 0; do
run "shape=$shape warps=$w sync=$sy" IPB_BOX_SHAPE=$shape IPB_BOX_WARPS=$w IPB_BOX_SYNC=$sy
done; done; done
run "shape=41 warps=8 sync=1 kpt=16" IPB_BOX_SHAPE=41 IPB_BOX_WARPS=8 IPB_BOX_KPT=16
run "shape=41 warps=8 sync=1 align=64 kpt=64" IPB_BOX_SHAPE=41 IPB_BOX_WARPS=8 IPB_BOX_ALIGN=64 IPB_BOX_KPT=64
run "shape=41 warps=8 sync=1 align=64 kpt=32" IPB_BOX_SHAPE=41 IPB_BOX_WARPS=8 IPB_BOX_ALIGN=64 IPB_BOX_KPT=32
