cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/r2_bench_8gpu.err
tail -c 1500 gpurun_out/r2_bench_8gpu.json; tail -5 gpurun_out/r2_bench_8gpu.err
ARTIC_DP_OVERLAP=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --fp32-wire --no-extras > gpurun_out/r2_bench_8gpu_blocking_fp32.json 2> gpurun_out/r2_bench_8gpu_blocking_fp32.err
tail -c 600 gpurun_out/r2_bench_8gpu_blocking_fp32.json
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,TUNING NCCL_DEBUG_FILE=gpurun_out/nccl_%h_%p.log timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err
tail -c 600 gpurun_out/r2_bench_4gpu.json
grep -h -i "nvls" gpurun_out/nccl_*.log | head -5 > gpurun_out/r2_nccl_nvls.log; rm -f gpurun_out/nccl_*.log; cat gpurun_out/r2_nccl_nvls.log
