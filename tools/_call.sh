for v in 100 60 30 -1; do
ARTIC_G_OBJECTIVE=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('G objective $v', d['ms_per_step'])" | tee -a gpurun_out/r1_gobj_64.log
done
