IPB_DEBUG=1 python bench.py --workload bicubic --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 >/dev/null | grep gtiles
IPB_GKPB=256 ncu --set full --clock-control none --import-source on -k regex:k_gtiles -s 3 -c 1 -f -o gpurun_out/prof_gt_budget python bench.py --workload budget --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/ncu_gt.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gtiles -s 3 -c 1 -f -o gpurun_out/prof_gt_bicubic python bench.py --workload bicubic --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/ncu_gt2.log 2>&1
