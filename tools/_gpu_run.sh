TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
timeout 600 python -m pytest tests/test_train_gpu.py tests/test_model_gpu.py tests/test_conv_implicit_gpu.py tests/test_dp_gpu.py -m gpu -q --timeout=900 2>&1 | tail -3
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err; echo "rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/r2_bench_n2_final.json'):
    if l.startswith("{"):
        d=json.loads(l); print("N=2 value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["replicas"]["identical"], d["exchange"]["G"], d["exchange"]["D"], [v["value"] for v in d["variants"]])
PY
tail -2 gpurun_out/r2_bench_n2_final.err
