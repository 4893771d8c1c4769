#!/bin/bash
# A/B of the point-staged kernels against the general ones (profiles/README.md): kernel time per BASELINE configuration
# with IPB_NO_GTILES=1 and with each tile shape.  Run on a B200:  tools_gpurun.sh 900 "bash tools/ab_gtiles.sh"
run() { # workload-args... -- env...
  local args=() ; while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
  env "$@" python bench.py "${args[@]}" --steps 10 --warmup 3 --no-cpu --no-e2e --no-secondary 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('${args[*]}', '$*', 'ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3))"
}
for w in "--workload budget" "--workload bicubic" "--workload bilinear --mask 1 --km 256"; do
  run $w -- IPB_NO_GTILES=1
  run $w -- IPB_DEBUG=0
  for m in 0 1 2; do run $w -- IPB_GMODE=$m; done
done
