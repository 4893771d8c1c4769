python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
ncu --set full --clock-control none --import-source on -k regex:k_neighbor_patch -s 3 -c 1 -f -o gpurun_out/prof_neighbor_patch python bench.py --workload neighbor --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/ncu_n.log 2>&1
python bench.py --workload neighbor --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_neighbor.json 2> gpurun_out/bench_neighbor.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_final.json').readline())
print('ours', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])
print('secondary', d['secondary']['value'], d['secondary']['ms_per_step'], d['secondary']['roofline']['frac'])
n=json.loads(open('gpurun_out/bench_neighbor.json').readline())
print('neighbor', n['value'], n['roofline']['kernel_ms'], n['roofline']['frac'], n['cpu_baseline']['spot_check'])"
