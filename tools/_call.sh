timeout 300 python tools/tc_probe.py tc0 > gpurun_out/r1_tcprobe3_tc0.log 2>&1; echo "probe rc=$?"
cut -c1-175 gpurun_out/r1_tcprobe3_tc0.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_bf16_tc3.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches_tc3_bf16.csv python tools/profile_step.py --precision bf16 > gpurun_out/r1_profile_step3.log 2>&1; echo "ncu rc=$?"
