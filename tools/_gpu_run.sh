timeout 300 python tools/splitk_probe.py > gpurun_out/r2_splitk_probe.txt 2>&1; cat gpurun_out/r2_splitk_probe.txt
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout=600 2>&1 | tail -3
