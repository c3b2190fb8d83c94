set -x
cd $GRAFT_REPO_ROOT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_bf16.csv python tools/profile_step.py --precision bf16 > gpurun_out/r2_prof_bf16.log 2>&1
timeout 900 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_bf16x3.csv python tools/profile_step.py --precision bf16x3 > gpurun_out/r2_prof_x3.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_bf16.csv 25 > gpurun_out/r2_launch_summary_bf16.txt; head -30 gpurun_out/r2_launch_summary_bf16.txt
python tools/launch_summary.py gpurun_out/r2_launches_bf16x3.csv 25 > gpurun_out/r2_launch_summary_bf16x3.txt; head -30 gpurun_out/r2_launch_summary_bf16x3.txt
python tools/traffic_summary.py gpurun_out/r2_launches_bf16.csv gpurun_out/r2_traffic.json
# tensor-pipe utilisation of the top three kernels inside the step (ncu --set full)
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"tapconv_tc_kernel|tapwgrad_tc_kernel|bigru" -s 100 -c 9 -o gpurun_out/r2_full_tc python tools/profile_step.py --precision bf16 > gpurun_out/r2_prof_full.log 2>&1
ncu -i gpurun_out/r2_full_tc.ncu-rep --page raw --csv > gpurun_out/r2_full_tc_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep
