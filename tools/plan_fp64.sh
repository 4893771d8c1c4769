#!/bin/bash
# FP64 work of the plan-building kernels (north_star: "the plan build is reported separately against FP64 peak"):
# ncu SASS op counts + durations of every plan kernel of one build, per workload -> gpurun_out/plan_fp64_<workload>.csv
cd "$(dirname "$0")/.."
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,gpu__time_duration.sum
for w in bilinear budget vector bicubic neighbor; do
  python bench.py --workload $w --steps 1 --warmup 1 --no-cpu --no-e2e --no-secondary > gpurun_out/plain_plan_$w.log 2>&1 &&
  ncu --metrics $M --clock-control none -k regex:"k_out_geo|k_plan|k_tile|k_box|k_gt_build|k_budget_collapse|k_idx|k_pole" --csv --log-file gpurun_out/plan_fp64_$w.csv \
      python bench.py --workload $w --steps 1 --warmup 1 --no-cpu --no-e2e --no-secondary > gpurun_out/ncu_plan_$w.log 2>&1
done
