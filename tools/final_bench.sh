python -m pytest tests -m gpu -x -q 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:k_gtiles -s 3 -c 1 -f -o gpurun_out/prof_gt_budget python bench.py --workload budget --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/ncu_gt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gtiles -s 3 -c 1 -f -o gpurun_out/prof_gt_bicubic python bench.py --workload bicubic --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/ncu_gt2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|ipb' --csv --log-file gpurun_out/launches_budget.csv python bench.py --workload budget --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_l.log 2>&1
python bench.py --workload budget --steps 20 --warmup 3 > gpurun_out/bench_budget.json 2> gpurun_out/bench_budget.err
for w in bicubic neighbor vector; do python bench.py --workload $w --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; done
for f in budget bicubic neighbor vector; do python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$f.json').readline())
print('$f', round(d['ms_per_step'],3), round(d['roofline']['kernel_ms'],3), round(d['roofline']['frac'],3), '%.3g'%d['value'], d['cpu_baseline'] and '%.3g'%d['cpu_baseline']['value'], d['e2e'] and '%.3g'%d['e2e']['value'], d['cpu_baseline'] and d['cpu_baseline'].get('spot_check'))"; done
