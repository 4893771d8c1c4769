python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "vector" 2>&1 | tail -2
run() { env "$@" python bench.py --workload vector --steps 10 --warmup 3 --no-e2e 2>gpurun_out/bench_vector.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('vector', '$*', 'ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), d['cpu_baseline']['spot_check'])"; }
run IPB_VPATCH=0
run IPB_VPATCH=1
run IPB_VKPB=4
run IPB_VKPB=64
