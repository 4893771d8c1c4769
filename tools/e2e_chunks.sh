for c in 128 384 768 1536; do
python bench.py --steps 3 --warmup 3 --no-cpu --chunk-mib $c 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); e=d['e2e']; print('chunk $c MiB: e2e ms', round(e['ms_per_step'],1), 'GB/s d2h', round(e['d2h_bytes_per_step']/e['ms_per_step']/1e6,1), 'value %.3g' % e['value'])"
done
