#!/bin/bash
# end-of-round evidence: default bench line, the reference arm, launch lists (ncu, after the plain runs), pageable e2e, smoke, full GPU suite
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py > gpurun_out/final_bench_default.json 2> gpurun_out/final_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "ref rc=$?"
python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary --pageable 2>/dev/null | tail -1 > gpurun_out/final_e2e_pageable.json
for w in vector neighbor bicubic; do python bench.py --workload $w --steps 20 --warmup 3 --no-cpu --no-secondary 2>/dev/null | tail -1 > gpurun_out/final_bench_$w.json; done
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bilinear_km1024.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary > gpurun_out/ncu_bil.log 2>&1
python bench.py --workload budget --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_budget_km256.csv python bench.py --workload budget --steps 2 --warmup 3 --no-cpu --no-e2e --no-secondary > gpurun_out/ncu_bud.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/final_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get('e2e') or {}
        print(f, 'ms', d.get('ms_per_step'), 'frac', (d.get('roofline') or {}).get('frac'), 'e2e_ms', e.get('ms_per_step'), 'link', e.get('frac_of_link_ceiling'), 'value', d.get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
(time python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider) > gpurun_out/r2_tests_final.log 2>&1
tail -6 gpurun_out/r2_tests_final.log
