timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r1_pytest_gpu_23.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_23.log 2>&1
ARTIC_GROUP=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_23_nogroup.log 2>&1
tail -3 gpurun_out/r1_pytest_gpu_23.log; tail -1 gpurun_out/r1_bench_23.log | cut -c1-300; tail -1 gpurun_out/r1_bench_23_nogroup.log | cut -c1-300
