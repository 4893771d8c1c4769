python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_staged_kernels_vs_oracle or scalar_vs_oracle" 2>&1 | tail -3
run() { # workload, env...
  w=$1; shift
  env "$@" python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_$w.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$w', '$*', 'ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'plan_ms', round(d['plan_build_ms'],1))"
}
run budget A=1
run bicubic A=1
run bicubic IPB_GKPB=64
