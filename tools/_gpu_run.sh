TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
./tools/microbench/fp32x2 > gpurun_out/r2_fp32x2_issue_rates.txt 2>&1; cat gpurun_out/r2_fp32x2_issue_rates.txt
timeout 300 $TR tools/dp_check.py > gpurun_out/r2_dp_check.json 2> gpurun_out/r2_dp_check.err; echo "dp_check rc=$?"; tail -c 1500 gpurun_out/r2_dp_check.json; grep -v "Warning\|warn\|nested" gpurun_out/r2_dp_check.err | tail -30
LD_DP_OVERLAP=1 timeout 300 $TR bench.py --gpus 2 --no-cpu-baseline --variants 0 --loop-steps 0 > gpurun_out/r2_bench_n2_overlap.json 2> gpurun_out/r2_bench_n2_overlap.err; echo "bench rc=$?"; tail -c 900 gpurun_out/r2_bench_n2_overlap.json; tail -5 gpurun_out/r2_bench_n2_overlap.err
