timeout 900 python bench.py --steps 60 --no-cpu-baseline --variants 0 > gpurun_out/r2_bench_steps60.json 2> gpurun_out/r2_bench_steps60.err; python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_steps60.json'):
    if l.startswith("{"):
        d=json.loads(l); print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "loop", d["loop"]["value"], d["loop"]["steps"], d["clocks"])
PY
