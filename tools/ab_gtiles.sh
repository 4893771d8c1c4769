python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_staged_kernels_vs_oracle or (scalar_vs_oracle and (4-0 or d-0 or 4-1))" 2>&1 | tail -3
run() { # args
  python bench.py --steps 10 --warmup 3 --no-e2e --no-secondary "$@" 2>gpurun_out/bench_x.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$*', 'ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), d['cpu_baseline']['spot_check'])"
}
run --workload bilinear --mask 1 --km 256
run --workload bicubic
