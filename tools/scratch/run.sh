python -m pytest tests/test_gpu_kernels.py -q -m gpu -x 2>&1 | tail -2
python tools/weights_bench.py 2>&1 | tail -7
ncu --set full --clock-control none --import-source on -k regex:'wprep_rows|wunprep_rows' -c 3 -o gpurun_out/weights_ncu3 -f python tools/weights_bench.py --ncu > /dev/null 2>&1
ncu -i gpurun_out/weights_ncu3.ncu-rep --page raw --csv > gpurun_out/weights_ncu3_raw.csv
ncu -i gpurun_out/weights_ncu3.ncu-rep --page source --csv --kernel-name wprep_rows_kernel > gpurun_out/weights_ncu3_src.csv 2>/dev/null
