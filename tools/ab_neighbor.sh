run() { env "$@" python bench.py --workload neighbor --steps 10 --warmup 3 --no-e2e 2>gpurun_out/bench_neighbor.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('neighbor', '$*', 'ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3), d['cpu_baseline']['spot_check'])"; }
run IPB_NKPB=4
run A=1
python -m pytest tests/test_gpu_parity.py tests/test_gpu_device_api.py -m gpu -x -q -k "(scalar_vs_oracle and (4-2 or d-2 or 8-2)) or golden_gpu or device" 2>&1 | tail -1
