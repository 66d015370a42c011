for st in 4 16 32; do timeout 300 python bench.py --workload eval --steps $st --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('steps $st', round(d['value'],1), 'layouts/s', round(d['ms_per_step'],2), 'ms per batch')
"; done
