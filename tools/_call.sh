timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r1_pytest_gpu_32.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_32.log 2>&1
tail -3 gpurun_out/r1_pytest_gpu_32.log; tail -1 gpurun_out/r1_bench_32.log | cut -c1-300
