set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2_t14.log; tail -6 gpurun_out/r2_t14.log
timeout 300 python tools/inv_profile.py > gpurun_out/r2_inv_profile_c.log 2>&1; cat gpurun_out/r2_inv_profile_c.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_1gpu_h.json 2>/dev/null; tail -c 400 gpurun_out/r2_bench_1gpu_h.json
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"tapwgrad_tc_kernel" -s 40 -c 6 -o gpurun_out/r2_full_wgrad python tools/profile_step.py --precision bf16 > gpurun_out/r2_prof_full_w.log 2>&1
ncu -i gpurun_out/r2_full_wgrad.ncu-rep --page raw --csv > gpurun_out/r2_full_wgrad_raw.csv 2>/dev/null
