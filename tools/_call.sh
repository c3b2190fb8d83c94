timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -q 2>&1 | tail -3 > gpurun_out/r1_pytest_dp2_62.log
tail -2 gpurun_out/r1_pytest_dp2_62.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_2gpu_62.log 2>&1
tail -1 gpurun_out/r1_bench_2gpu_62.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-200
