cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_plugins.py -m gpu -q -x -k "yaml or mri" 2>&1 | tail -15 > gpurun_out/r2_t11.log; tail -8 gpurun_out/r2_t11.log
for D in 100 75 50 25; do
ARTIC_D_OBJECTIVE=$D timeout 600 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('D_OBJECTIVE $D', round(d['ms_per_step'],3))"
done
for G in 40 80; do
ARTIC_G_OBJECTIVE=$G timeout 600 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('G_OBJECTIVE $G', round(d['ms_per_step'],3))"
done
