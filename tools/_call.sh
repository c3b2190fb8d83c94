timeout 600 python bench.py > gpurun_out/r1_bench_73.log 2>&1
tail -1 gpurun_out/r1_bench_73.log | cut -c1-260
