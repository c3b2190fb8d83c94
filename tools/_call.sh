timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for d in "18=0" "18=1"; do
ARTIC_DEBUG="$d" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$d', d['ms_per_step'])" | tee -a gpurun_out/r1_wgepi_53.log
done
