#!/bin/bash
mkdir -p gpurun_out
IPB_LO_BYTES=1 python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | tail -1 > gpurun_out/e2e4_lo_bytes.json
python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | tail -1 > gpurun_out/e2e4_lo_bits.json
IPB_H2D_FULL=1 python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | tail -1 > gpurun_out/e2e4_lo_bits_h2dfull.json
python bench.py --workload budget --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | tail -1 > gpurun_out/e2e4_budget_lo_bits.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e2e4_*.json')):
    try:
        d=json.loads(open(f).read())
        e=d.get('e2e') or {}
        print(f, 'ms', round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3), 'e2e_ms', e.get('ms_per_step'), 'd2h', e.get('d2h_bytes_per_step'), 'h2d', e.get('h2d_bytes_per_step'), 'link', e.get('frac_of_link_ceiling'), 'match', e.get('matches_device_path'))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
python -m pytest tests/test_gpu_device_api.py -m gpu -q -x --timeout 900 -p no:cacheprovider -k "lo_crosses or sharded or ragged" 2>&1 | tail -4
