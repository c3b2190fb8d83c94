python -m pytest tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -3
python -m pytest tests/test_gpu_models.py tests/test_gpu_plugins.py -q -m gpu -x 2>&1 | tail -3
python tools/weights_bench.py 2>&1 | tail -7
