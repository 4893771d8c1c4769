for w in bilinear budget bicubic neighbor vector; do python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-e2e 2>gpurun_out/bench_$w.err | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('$w', 'ms', round(d['roofline']['kernel_ms'],3), 'plan_ms cold', round(d['plan_build_ms'],1), 'warm', round(d['plan_rebuild_ms_warm'],2))"; done
