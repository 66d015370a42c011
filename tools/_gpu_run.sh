timeout 1700 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/r2_pytest_gpu_final.log 2>&1; tail -6 gpurun_out/r2_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; grep "^{" gpurun_out/r2_bench_final.json | tail -c 2600; tail -3 gpurun_out/r2_bench_final.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "ref rc=$?"; tail -c 700 gpurun_out/r2_bench_ref.json
