timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_21_default.log 2>&1
ARTIC_DEBUG="8=110" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_21_compact110.log 2>&1
ARTIC_DEBUG="8=75" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_21_compact75.log 2>&1
ARTIC_DEBUG="8=110" timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r1_pytest_gpu_21.log
for f in default compact110 compact75; do tail -1 gpurun_out/r1_bench_21_$f.log | cut -c1-260; done; cat gpurun_out/r1_pytest_gpu_21.log
