timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline >/dev/null 2>&1
for v in "16=6" "16=9" "16=12" "16=18" "14=8"; do
ARTIC_DEBUG=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('DEBUG=$v', d['ms_per_step'])" | tee -a gpurun_out/r1_knobs_69.log
done
