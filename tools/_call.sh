timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r1_pytest_gpu_24.log
timeout 900 python bench.py > gpurun_out/r1_bench_24.log 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1_bench_24_reference.log 2>&1
timeout 300 python tools/bench_decode.py --cpu > gpurun_out/r1_bench_decode_24.log 2>&1
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r1_launches_24.csv python tools/profile_step.py > gpurun_out/r1_profile_step_24.log 2>&1
timeout 900 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:tapconv_tc_kernel --launch-skip 60 --launch-count 4 -o /tmp/r1_tc_full -f python tools/profile_step.py > gpurun_out/r1_ncu_full_24.log 2>&1
ncu -i /tmp/r1_tc_full.ncu-rep --page details --csv > gpurun_out/r1_tc_full_details_24.csv 2>/dev/null
ncu -i /tmp/r1_tc_full.ncu-rep --page source --csv 2>/dev/null | head -c 3000000 > gpurun_out/r1_tc_full_source_24.csv
ls -la /tmp/r1_tc_full.ncu-rep | tee -a gpurun_out/r1_ncu_full_24.log
tail -3 gpurun_out/r1_pytest_gpu_24.log; tail -1 gpurun_out/r1_bench_24.log | cut -c1-400; tail -1 gpurun_out/r1_bench_24_reference.log | cut -c1-300; tail -1 gpurun_out/r1_bench_decode_24.log | cut -c1-300
