timeout 900 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/r2_ops python tools/ncu_ops.py > gpurun_out/r2_ncu_ops.log 2>&1; tail -2 gpurun_out/r2_ncu_ops.log
ncu -i /tmp/r2_ops.ncu-rep --page raw --csv > gpurun_out/r2_ops_raw.csv 2>/dev/null; wc -l gpurun_out/r2_ops_raw.csv
ncu -i /tmp/r2_ops.ncu-rep --page details > gpurun_out/r2_ops_details.txt 2>/dev/null; wc -l gpurun_out/r2_ops_details.txt
timeout 600 ncu --set full --clock-control none -k regex:gemm_bf16 -s 6 -c 3 -f -o /tmp/r2_gemm python tools/gemm_ncu.py > gpurun_out/r2_ncu_gemm.log 2>&1; tail -2 gpurun_out/r2_ncu_gemm.log
ncu -i /tmp/r2_gemm.ncu-rep --page raw --csv > gpurun_out/r2_gemm_raw.csv 2>/dev/null; wc -l gpurun_out/r2_gemm_raw.csv
ncu -i /tmp/r2_gemm.ncu-rep --page details > gpurun_out/r2_gemm_details.txt 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_step_launches_ncu.csv python bench.py --ncu --graph 0 --no-cpu-baseline --variants 0 > gpurun_out/r2_ncu_launch.log 2>&1; wc -l gpurun_out/r2_step_launches_ncu.csv
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "deterministic" 2>&1 | tail -3
du -sh gpurun_out
