timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r1_pytest_gpu_5.log
timeout 300 python tools/layer_times.py > gpurun_out/r1_layer_times_5.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_bf16_tc5.log 2>&1
timeout 300 python tools/tc_probe.py tc0 > gpurun_out/r1_tcprobe5_tc0.log 2>&1
tail -5 gpurun_out/r1_pytest_gpu_5.log
