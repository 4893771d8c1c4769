#!/bin/bash
# A/B runs of the TMA-staged bilinear kernel (csrc/ipb_box.cuh) on one GPU: prints ms per step of the headline workload
# for each setting of the measurement knobs.  usage: tools/sweep_box.sh [extra bench args]
cd "$(dirname "$0")/.."
run() { # label, env...
  local label="$1"; shift
  local out
  out=$(env "$@" python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-secondary $EXTRA 2>/dev/null | tail -1)
  echo "$label $(echo "$out" | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("ms_per_step=%.4f kernel_ms=%.4f frac=%.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))' 2>/dev/null || echo "FAILED: $out")"
}
EXTRA="$*"
run "old-tiles      " IPB_NO_BOX=1
run "box default    " IPB_NO_BOX=0
for mb in 5 6; do run "box minb=$mb     " IPB_BOX_MINB=$mb; done
for al in 16 32 64; do run "box align=$al minb=5" IPB_BOX_ALIGN=$al IPB_BOX_MINB=5; done
for nr in 2 8; do run "box nr=$nr        " IPB_BOX_NR=$nr; done
for kp in 8 16 64; do run "box kpt=$kp minb=5 " IPB_BOX_KPT=$kp IPB_BOX_MINB=5; done
