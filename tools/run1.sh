set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "conv_layer" 2>&1 | tail -40 > gpurun_out/r2_k1.log
timeout 900 python -m pytest tests/test_gpu_models.py -m gpu -q -k "full_width" 2>&1 | tail -60 > gpurun_out/r2_m1.log
timeout 600 python bench.py --steps 10 --precision bf16x3 --no-cpu-baseline > gpurun_out/r2_bench_x3_a.json 2> gpurun_out/r2_bench_x3_a.err
timeout 600 python bench.py --steps 10 --precision bf16 --no-cpu-baseline > gpurun_out/r2_bench_bf16_a.json 2> gpurun_out/r2_bench_bf16_a.err
tail -5 gpurun_out/r2_k1.log gpurun_out/r2_m1.log
