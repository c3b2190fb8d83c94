timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r1_pytest_gpu_9.log
timeout 300 python tools/tc_timeline.py > gpurun_out/r1_tc_timeline_9.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_9.log 2>&1
tail -12 gpurun_out/r1_pytest_gpu_9.log; tail -1 gpurun_out/r1_bench_9.log | cut -c1-300
