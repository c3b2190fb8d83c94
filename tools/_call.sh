timeout 100 python bench.py --steps 10 --warmup 3 > gpurun_out/r1_bench_74.log 2>&1
tail -1 gpurun_out/r1_bench_74.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d.get('stft_loss'), d.get('cpu_baseline',{}).get('value'))"
