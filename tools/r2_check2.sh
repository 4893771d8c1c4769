#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary 2>/dev/null | tail -1 > gpurun_out/e2e5_pinned.json
python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary --pageable 2>/dev/null | tail -1 > gpurun_out/e2e5_pageable_staged.json
IPB_NO_STAGE=1 python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary --pageable 2>/dev/null | tail -1 > gpurun_out/e2e5_pageable_driver.json
IPB_HOST_THREADS=4 python bench.py --steps 10 --warmup 3 --no-cpu --no-secondary --pageable 2>/dev/null | tail -1 > gpurun_out/e2e5_pageable_staged_t4.json
python bench.py --workload budget --steps 10 --warmup 3 --no-cpu --no-secondary --pageable 2>/dev/null | tail -1 > gpurun_out/e2e5_budget_pageable_staged.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e2e5_*.json')):
    try:
        d=json.loads(open(f).read())
        e=d.get('e2e') or {}
        print(f, 'ms', round(d['ms_per_step'],3), 'e2e_ms', e.get('ms_per_step'), 'd2h', e.get('d2h_bytes_per_step'), 'link', e.get('frac_of_link_ceiling'), 'match', e.get('matches_device_path'), e.get('host_arrays'))
    except Exception as ex:
        print(f, 'unreadable', ex)
PY
python -m pytest tests/test_gpu_device_api.py tests/test_gpu_parity.py -m gpu -q -x --timeout 900 -p no:cacheprovider 2>&1 | tail -4
