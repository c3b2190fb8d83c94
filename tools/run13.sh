set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x 2>&1 | tail -12 > gpurun_out/r2_t13.log; tail -6 gpurun_out/r2_t13.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bigru" -c 1 -o gpurun_out/r2_full_gru python tools/inv_profile.py > gpurun_out/r2_prof_gru.log 2>&1
ncu -i gpurun_out/r2_full_gru.ncu-rep --page raw --csv > gpurun_out/r2_full_gru_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_full_gru.ncu-rep --page source --csv > gpurun_out/r2_full_gru_source.csv 2>/dev/null
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"tapwgrad_tc_kernel" -s 40 -c 6 -o gpurun_out/r2_full_wgrad python tools/profile_step.py --precision bf16 > gpurun_out/r2_prof_full_w.log 2>&1
ncu -i gpurun_out/r2_full_wgrad.ncu-rep --page raw --csv > gpurun_out/r2_full_wgrad_raw.csv 2>/dev/null
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:"tapconv_tc_kernel" -s 150 -c 6 -o gpurun_out/r2_full_tc_x3 python tools/profile_step.py --precision bf16x3 > gpurun_out/r2_prof_full_x3.log 2>&1
ncu -i gpurun_out/r2_full_tc_x3.ncu-rep --page raw --csv > gpurun_out/r2_full_tc_x3_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_1gpu_h.json 2>/dev/null; tail -c 400 gpurun_out/r2_bench_1gpu_h.json
