timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r1_pytest_gpu_17.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_17.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_17.csv python tools/profile_step.py > gpurun_out/r1_profile_step_17.log 2>&1
tail -3 gpurun_out/r1_pytest_gpu_17.log; tail -1 gpurun_out/r1_bench_17.log | cut -c1-300
