nvidia-smi -L
timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_dropout_gpu.py -m gpu -q -x > gpurun_out/r2_attn_test.log 2>&1; tail -25 gpurun_out/r2_attn_test.log
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention -o gpurun_out/r2_attn_v3 python tools/ncu_ops.py > gpurun_out/r2_ncu_v3.log 2>&1; tail -3 gpurun_out/r2_ncu_v3.log
timeout 250 python bench.py --variants 0 --no-cpu-baseline --dropout 0 > gpurun_out/r2_bench_attn3_nodrop.json 2> gpurun_out/r2_bench_attn3_nodrop.err; tail -c 1500 gpurun_out/r2_bench_attn3_nodrop.json; tail -3 gpurun_out/r2_bench_attn3_nodrop.err
timeout 250 python bench.py --variants 0 --no-cpu-baseline --dropout 1 > gpurun_out/r2_bench_attn3_drop.json 2> gpurun_out/r2_bench_attn3_drop.err; tail -c 1500 gpurun_out/r2_bench_attn3_drop.json; tail -3 gpurun_out/r2_bench_attn3_drop.err
