timeout 600 python -m pytest tests/test_conv_implicit_gpu.py tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_conv_gradfix_gpu.py -m gpu -q --timeout=600 2>&1 | tail -25
