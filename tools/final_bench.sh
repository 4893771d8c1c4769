python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_final.json').readline())
print('ours', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])
print('secondary', d['secondary']['value'], d['secondary']['ms_per_step'], d['secondary']['roofline']['frac'])
r=json.loads(open('gpurun_out/bench_ref.json').readline())
print('ref', r['value'], r['cpu_baseline']['cores'], r.get('impl'))"
