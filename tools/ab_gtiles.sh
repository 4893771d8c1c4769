IPB_GMODE=3 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_staged_kernels_vs_oracle" 2>&1 | tail -3
run() { # workload, env...
  w=$1; shift
  env "$@" python bench.py --workload $w --steps 10 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_$w.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$w', '$*', 'ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), 'plan_ms', round(d['plan_build_ms'],1))"
}
run bicubic IPB_GMODE=2
run bicubic IPB_GMODE=3 IPB_DEBUG=1
grep gtiles gpurun_out/bench_bicubic.err
ncu --set full --clock-control none --import-source on -k regex:k_gtiles -s 3 -c 1 -f -o gpurun_out/prof_gt_bicubic3 env IPB_GMODE=3 python bench.py --workload bicubic --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/ncu_gt2.log 2>&1
python __graft_entry__.py smoke 2>&1 | tail -2
