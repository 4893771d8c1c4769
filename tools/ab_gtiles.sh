python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_staged_kernels_vs_oracle or (scalar_vs_oracle and d-3)" 2>&1 | tail -3
run() { # workload, env...
  w=$1; shift
  env "$@" python bench.py --workload $w --steps 10 --warmup 3 --no-e2e 2>gpurun_out/bench_$w.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$w', '$*', 'ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), d['cpu_baseline']['spot_check'])"
}
run budget A=1
run budget IPB_GMODE=2
run budget IPB_GMODE=0
