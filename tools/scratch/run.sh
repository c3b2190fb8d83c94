set -x
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "past_fc" 2>&1 | tail -5
python -m pytest tests/test_gpu_models.py tests/test_gpu_plugins.py tests/test_gpu_inversion.py -q -m gpu -x 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_mlp.json 2> gpurun_out/bench_mlp.err
tail -c 1500 gpurun_out/bench_mlp.json
