set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cluster" 2>&1 | tail -25 > gpurun_out/r2_t6.log; tail -12 gpurun_out/r2_t6.log
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_models.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2_t6b.log; tail -3 gpurun_out/r2_t6b.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_1gpu_e.json 2>gpurun_out/r2_bench_1gpu_e.err; tail -c 700 gpurun_out/r2_bench_1gpu_e.json; tail -3 gpurun_out/r2_bench_1gpu_e.err
ARTIC_DEBUG=22=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_1gpu_e_nocluster.json 2>/dev/null; tail -c 700 gpurun_out/r2_bench_1gpu_e_nocluster.json
ARTIC_DEBUG=22=4 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_1gpu_e_cs4.json 2>/dev/null; tail -c 700 gpurun_out/r2_bench_1gpu_e_cs4.json
timeout 300 python - > gpurun_out/r2_stft_ms.log 2>&1 <<'PY'
import torch, bench
print("stft loss ms", bench.stft_loss_gpu_ms(torch.device("cuda",0)))
PY
cat gpurun_out/r2_stft_ms.log
