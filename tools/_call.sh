timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r1_pytest_gpu_39.log
tail -3 gpurun_out/r1_pytest_gpu_39.log
