set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_dp2_test.log; tail -5 gpurun_out/r2_dp2_test.log
for OV in 1 0; do
ARTIC_DP_OVERLAP=$OV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu_ov$OV.json 2> gpurun_out/r2_bench_2gpu_ov$OV.err
tail -c 1200 gpurun_out/r2_bench_2gpu_ov$OV.json; tail -3 gpurun_out/r2_bench_2gpu_ov$OV.err
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_1gpu_c.json 2>gpurun_out/r2_bench_1gpu_c.err; tail -c 600 gpurun_out/r2_bench_1gpu_c.json
timeout 300 python tools/inv_profile.py > gpurun_out/r2_inv_profile.log 2>&1; cat gpurun_out/r2_inv_profile.log
