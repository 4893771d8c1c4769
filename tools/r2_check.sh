#!/bin/bash
# one GPU call: D2H probe, vector kernel variants, e2e with lo as bytes / as bits, then the GPU test suite
mkdir -p gpurun_out
./tools/d2h_probe > gpurun_out/d2h_probe.log 2>&1
for k in 4 3 2; do
  IPB_VKU=$k python bench.py --workload vector --steps 20 --warmup 3 --no-cpu --no-e2e --no-secondary 2>/dev/null | tail -1 > gpurun_out/vec_ku$k.json
done
IPB_LO_BYTES=1 python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | tail -1 > gpurun_out/e2e_lo_bytes.json
python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | tail -1 > gpurun_out/e2e_lo_bits.json
IPB_HOST_THREADS=2 python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | tail -1 > gpurun_out/e2e_lo_bits_t2.json
python bench.py --workload budget --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | tail -1 > gpurun_out/e2e_budget_lo_bits.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/vec_ku*.json')+glob.glob('gpurun_out/e2e_*.json')):
    try:
        d=json.loads(open(f).read())
        e=d.get('e2e') or {}
        print(f, 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'e2e_ms', e.get('ms_per_step'), 'd2h', e.get('d2h_bytes_per_step'), 'link', e.get('frac_of_link_ceiling'), 'match', e.get('matches_device_path'))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
cat gpurun_out/d2h_probe.log
(time python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider) > gpurun_out/r2_tests_full.log 2>&1
tail -8 gpurun_out/r2_tests_full.log
