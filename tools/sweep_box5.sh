#!/bin/bash
cd "$(dirname "$0")/.."
run() { local label="$1"; shift; local out
  out=$(env "$@" python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-secondary $EXTRA 2>/dev/null | tail -1)
  echo "$label $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("ms_per_step=%.4f kernel_ms=%.4f frac=%.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))' 2>/dev/null || echo "FAILED: $out")"; }
EXTRA="$*"
run "lofill=0 swap=0 w=4" IPB_BOX_LOFILL=0
run "lofill=1 swap=0 w=4" IPB_BOX_LOFILL=1
run "lofill=1 swap=1 w=4" IPB_BOX_SWAP=1
run "lofill=1 swap=1 w=4 sync=1" IPB_BOX_SWAP=1 IPB_BOX_SYNC=1
run "lofill=1 swap=1 w=8" IPB_BOX_SWAP=1 IPB_BOX_WARPS=8
run "lofill=1 swap=1 w=8 sync=1" IPB_BOX_SWAP=1 IPB_BOX_WARPS=8 IPB_BOX_SYNC=1
run "lofill=1 swap=0 w=8 sync=1" IPB_BOX_WARPS=8 IPB_BOX_SYNC=1
run "lofill=1 swap=1 w=4 kpt=16" IPB_BOX_SWAP=1 IPB_BOX_KPT=16
run "lofill=1 swap=0 w=4 kpt=16" IPB_BOX_KPT=16
run "lofill=1 swap=1 w=4 align=64 kpt=64" IPB_BOX_SWAP=1 IPB_BOX_ALIGN=64 IPB_BOX_KPT=64
run "lofill=1 swap=0 w=4 align=64 kpt=32" IPB_BOX_ALIGN=64 IPB_BOX_KPT=32
