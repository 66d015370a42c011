timeout 900 python -m pytest tests/test_conv_gradfix_gpu.py tests/test_graph_gpu.py tests/test_lanes_gpu.py -m gpu -q --timeout=600 > gpurun_out/r2_pytest_gpu_b.log 2>&1; tail -12 gpurun_out/r2_pytest_gpu_b.log
timeout 600 python bench.py > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 2500 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
