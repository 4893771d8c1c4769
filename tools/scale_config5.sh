#!/bin/bash
# BASELINE config 5: ipolates neighbor + bicubic, 0.25deg global -> rotated lat-lon E grid, km = 8192 sharded over N GPUs
# (strong scaling: the field batch is fixed, rank r owns a contiguous slice).  usage: tools/scale_config5.sh N [N ...]
cd "$(dirname "$0")/.."
for n in "$@"; do
  for w in neighbor bicubic; do
    if [ "$n" = 1 ]; then
      python bench.py --workload $w --km 8192 --scaling strong --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/scale5_${w}_n$n.json
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --workload $w --km 8192 \
        --scaling strong --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | tail -1 > gpurun_out/scale5_${w}_n$n.json
    fi
    python -c "import json; d=json.load(open('gpurun_out/scale5_${w}_n$n.json')); print('N=$n', '$w', 'value %.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'km/gpu', d['config']['km_per_gpu'], 'frac %.3f' % d['roofline']['frac'])"
  done
done
