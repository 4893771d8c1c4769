#!/bin/bash
cd "$(dirname "$0")/.."
run() { local label="$1"; shift; local out
  out=$(env "$@" python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-secondary $EXTRA 2>/dev/null | tail -1)
  echo "$label $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("ms_per_step=%.4f kernel_ms=%.4f frac=%.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))' 2>/dev/null || echo "FAILED: $out")"; }
EXTRA="$*"
for w in 4 8; do for mb in 2 5; do for ppw in 4 8 16; do
run "pipe warps=$w minb=$mb ppw=$ppw" IPB_BOX_SHAPE=41 IPB_BOX_WARPS=$w IPB_BOX_MINB=$mb IPB_BOX_PPW=$ppw
done; done; done
run "pipe warps=4 minb=4 ppw=8 align=64 kpt=64" IPB_BOX_WARPS=4 IPB_BOX_PPW=8 IPB_BOX_ALIGN=64 IPB_BOX_KPT=64
run "pipe warps=4 minb=4 ppw=8 kpt=16" IPB_BOX_WARPS=4 IPB_BOX_PPW=8 IPB_BOX_KPT=16
run "pipe warps=4 minb=4 ppw=32" IPB_BOX_WARPS=4 IPB_BOX_PPW=32
