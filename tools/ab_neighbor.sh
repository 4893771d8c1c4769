#!/bin/bash
# neighbor patch kernel: fields per block (profiles/README.md).  tools_gpurun.sh 300 "bash tools/ab_neighbor.sh"
run() { env "$@" python bench.py --workload neighbor --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('neighbor', '$*', 'ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3))"; }
for k in 4 8 12 16; do run IPB_NKPB=$k; done
run IPB_NPATCH=0
