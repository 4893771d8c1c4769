#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload budget4 --km 64 --steps 3 --warmup 3 --no-cpu --no-e2e --no-secondary > gpurun_out/b7.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_budget_scalar -s 2 -c 1 -o gpurun_out/r2_budget4 -f python bench.py --workload budget4 --km 64 --steps 3 --warmup 3 --no-cpu --no-e2e --no-secondary > gpurun_out/b7_ncu.log 2>&1
tail -3 gpurun_out/b7_ncu.log
ls -la gpurun_out/r2_budget4.ncu-rep
