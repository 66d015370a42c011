timeout 300 python -m pytest tests/test_attention_gpu.py tests/test_dropout_gpu.py -m gpu -q -x > gpurun_out/r2_attn_test5.log 2>&1; tail -3 gpurun_out/r2_attn_test5.log
timeout 120 python tools/attn_trace.py > gpurun_out/r2_attn_trace2.txt 2>&1; head -12 gpurun_out/r2_attn_trace2.txt
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention_fwd -o gpurun_out/r2_attn_v7 python tools/ncu_ops.py > gpurun_out/r2_ncu_v7.log 2>&1; tail -2 gpurun_out/r2_ncu_v7.log
timeout 250 python bench.py --variants 0 --no-cpu-baseline --dropout 0 > gpurun_out/r2_bench_attn7_nodrop.json 2> gpurun_out/r2_bench_attn7_nodrop.err; tail -c 300 gpurun_out/r2_bench_attn7_nodrop.json; tail -3 gpurun_out/r2_bench_attn7_nodrop.err
