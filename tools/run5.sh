set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_t5.log; tail -5 gpurun_out/r2_t5.log
timeout 900 python -m pytest tests/test_gpu_dp.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2_dp2_test.log; tail -3 gpurun_out/r2_dp2_test.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_1gpu_d.json 2>gpurun_out/r2_bench_1gpu_d.err; tail -c 900 gpurun_out/r2_bench_1gpu_d.json; tail -3 gpurun_out/r2_bench_1gpu_d.err
ARTIC_DEBUG=21=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_1gpu_d_nobias.json 2>/dev/null; tail -c 900 gpurun_out/r2_bench_1gpu_d_nobias.json
