python bench.py --workload budget --steps 10 --warmup 3 --no-e2e --no-cpu 2>gpurun_out/bench_budget.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('budget minb3', 'ms', round(d['roofline']['kernel_ms'],3), 'frac', round(d['roofline']['frac'],3))"
