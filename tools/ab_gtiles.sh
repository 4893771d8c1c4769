python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { # workload, env...
  w=$1; shift
  env "$@" python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_$w.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$w', '$*', 'ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'plan_ms', round(d['plan_build_ms'],1))"
}
run budget IPB_DEBUG=1
run bicubic IPB_DEBUG=1
run bicubic IPB_GKPB=256
run neighbor IPB_DEBUG=1
grep gtiles gpurun_out/bench_*.err
