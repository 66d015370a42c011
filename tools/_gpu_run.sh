timeout 300 python -m pytest tests/test_conv_implicit_gpu.py -m gpu -q --timeout=300 2>&1 | grep -E "^E|passed|failed" | head -12
timeout 600 python bench.py --no-cpu-baseline --variants 0 --loop-steps 0 > gpurun_out/r2_bench_j.json 2> gpurun_out/r2_bench_j.err
grep -o '"value": [0-9.]*, "unit": "samples/s", "n_gpus": 1, "steps": 10, "warmup": 3, "ms_per_step": [0-9.]*' gpurun_out/r2_bench_j.json; tail -2 gpurun_out/r2_bench_j.err
timeout 600 python -m pytest tests/test_stylegan2_gpu.py tests/test_model_gpu.py tests/test_train_gpu.py -m gpu -q --timeout=900 2>&1 | tail -3
LD_GEMM_LOG=gpurun_out/r2_gemm_log.json timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_step_launches_ncu.csv python bench.py --ncu --graph 0 --no-cpu-baseline --variants 0 > gpurun_out/r2_ncu_launch.log 2>&1; wc -l gpurun_out/r2_step_launches_ncu.csv
