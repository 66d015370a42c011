nvidia-smi -L
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_dropout_gpu.py -m gpu -q -x > gpurun_out/r2_attn_test3.log 2>&1; tail -15 gpurun_out/r2_attn_test3.log
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_fwd -o gpurun_out/r2_attn_v5 python tools/ncu_ops.py > gpurun_out/r2_ncu_v5.log 2>&1; tail -3 gpurun_out/r2_ncu_v5.log
timeout 250 python bench.py --variants 0 --no-cpu-baseline --dropout 0 > gpurun_out/r2_bench_attn5_nodrop.json 2> gpurun_out/r2_bench_attn5_nodrop.err; tail -c 600 gpurun_out/r2_bench_attn5_nodrop.json; tail -3 gpurun_out/r2_bench_attn5_nodrop.err
