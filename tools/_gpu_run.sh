for v in "" "LD_CONV_IMPLICIT=0" "LD_GEMM_2SM=0" "LD_LANES=0" "LD_UPFIRDN_TILED=0"; do env $v timeout 300 python bench.py --workload eval --steps 4 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v', round(d['value'],1), round(d['ms_per_step'],2))
"; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2_eval_launches_ncu.csv python bench.py --workload eval --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1; wc -l gpurun_out/r2_eval_launches_ncu.csv
