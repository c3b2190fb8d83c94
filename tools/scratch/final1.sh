set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/final_bench_bf16.json 2> gpurun_out/final_bench_bf16.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err
python bench.py --precision bf16x3 --no-extras --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/final_bench_bf16x3.json 2> gpurun_out/final_bench_bf16x3.err
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/final_launches_bf16.csv python tools/profile_step.py > /dev/null 2>&1
tail -3 gpurun_out/final_pytest_gpu.log; tail -2 gpurun_out/final_smoke.log
