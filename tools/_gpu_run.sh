timeout 300 python tools/gemm_probe.py > gpurun_out/r2_gemm_epilogue_variants.txt 2>&1; cat gpurun_out/r2_gemm_epilogue_variants.txt
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -m gpu -q -x --timeout=600 2>&1 | tail -5
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -s --deselect tests/test_gemm_gpu.py --deselect tests/test_kernels_gpu.py > gpurun_out/r2_pytest_gpu_c.log 2>&1; tail -15 gpurun_out/r2_pytest_gpu_c.log
timeout 600 python bench.py --no-cpu-baseline --variants 0 --loop-steps 0 > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -c 1800 gpurun_out/r2_bench_c.json; tail -3 gpurun_out/r2_bench_c.err
