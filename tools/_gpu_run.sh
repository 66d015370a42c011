timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/r2_pytest_gpu_a.log 2>&1; tail -25 gpurun_out/r2_pytest_gpu_a.log
timeout 400 python bench.py --variants 0 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 900 gpurun_out/r2_bench_b.json; tail -5 gpurun_out/r2_bench_b.err
