set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 3000 gpurun_out/r2_bench_b.json; tail -5 gpurun_out/r2_bench_b.err
ARTIC_LOG_GENERIC=1 timeout 600 python bench.py --steps 3 --no-extras 2>&1 | grep "generic" | sort | uniq -c > gpurun_out/r2_generic_shapes.log
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest_gpu.log; tail -5 gpurun_out/r2_pytest_gpu.log
