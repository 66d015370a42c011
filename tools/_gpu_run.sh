TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_driver_flags.json 2> gpurun_out/r2_bench_n2_driver_flags.err; echo "rc=$?"; grep "^{" gpurun_out/r2_bench_n2_driver_flags.json | tail -c 1800; tail -4 gpurun_out/r2_bench_n2_driver_flags.err
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2_bench_n2_ref.json 2> gpurun_out/r2_bench_n2_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/r2_bench_n2_ref.json
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q 2>&1 | tail -3
