timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r1_pytest_gpu_31.log
timeout 100 python tools/tc_timeline.py > gpurun_out/r1_tc_timeline_31.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_31.log 2>&1
timeout 900 python tools/tc_sweep.py fwd dgrad > gpurun_out/r1_tc_sweep_31.log 2>&1
tail -3 gpurun_out/r1_pytest_gpu_31.log; tail -1 gpurun_out/r1_bench_31.log | cut -c1-300
