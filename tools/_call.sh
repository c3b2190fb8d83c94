timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline >/dev/null 2>&1
for v in "order,spectral" "order,spectral,order2" "order,spectral" "order,spectral,order2"; do
ARTIC_BG=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('BG=$v', d['ms_per_step'])" | tee -a gpurun_out/r1_bg_67.log
done
