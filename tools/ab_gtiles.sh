python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "staged or (scalar_vs_oracle and (d-3 or 4-1 or d-1))" 2>&1 | tail -3
run() { # workload, env...
  w=$1; shift
  env "$@" python bench.py --workload $w --steps 10 --warmup 3 --no-e2e 2>gpurun_out/bench_$w.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$w', '$*', 'ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), d['cpu_baseline']['spot_check'])"
}
run budget A=1
run bicubic A=1
