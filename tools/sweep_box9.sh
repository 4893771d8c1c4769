#!/bin/bash
cd "$(dirname "$0")/.."
run() { local label="$1"; shift; local out
  out=$(env "$@" python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-secondary $EXTRA 2>/dev/null | tail -1)
  echo "$label $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("ms_per_step=%.4f kernel_ms=%.4f frac=%.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))' 2>/dev/null || echo "FAILED: $out")"; }
EXTRA="$*"
for f in 4 8 16; do for mb in 5 6; do run "F=$f minb=$mb" IPB_BOX_F=$f IPB_BOX_MINB=$mb; done; done
run "F=16 minb=6 align=64 kpt=64" IPB_BOX_F=16 IPB_BOX_ALIGN=64 IPB_BOX_KPT=64
run "F=4 minb=6 kpt=16" IPB_BOX_F=4 IPB_BOX_KPT=16
