ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/decode_launches.csv python tools/bench_decode.py --frames 50 --iters 1 > gpurun_out/decode_ncu.log 2>&1
tail -3 gpurun_out/decode_ncu.log
