timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r1_pytest_gpu_12.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_12.log 2>&1
ARTIC_STREAMS=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_12_nostreams.log 2>&1
timeout 900 python tools/tc_sweep.py fwd > gpurun_out/r1_tc_sweep_12.log 2>&1
timeout 300 python tools/tc_timeline.py > gpurun_out/r1_tc_timeline_12.log 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_12.csv python tools/profile_step.py > gpurun_out/r1_profile_step_12.log 2>&1
tail -4 gpurun_out/r1_pytest_gpu_12.log; tail -1 gpurun_out/r1_bench_12.log | cut -c1-300; tail -1 gpurun_out/r1_bench_12_nostreams.log | cut -c1-300
