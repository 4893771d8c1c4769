python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py -m gpu -x -q -k "not every_tile_shape and not golden" 2>&1 | tail -3
run() { # workload, env...
  w=$1; shift
  env IPB_DEBUG=1 "$@" python bench.py --workload $w --steps 10 --warmup 3 --no-e2e --mask 1 --km 256 --no-secondary 2>gpurun_out/bench_$w.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$w masked km256', '$*', 'ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), d['cpu_baseline']['spot_check'])"
}
run bilinear IPB_NO_GTILES=1
run bilinear A=1
grep gtiles gpurun_out/bench_bilinear.err
