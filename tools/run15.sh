cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_inversion.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/inv_profile.py 2>&1 | grep gru
for CFG in "100 60" "60 60" "30 60" "60 30" "30 30"; do
set -- $CFG
ARTIC_D_OBJECTIVE=$1 ARTIC_G_OBJECTIVE=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --precision bf16x3 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('x3 D_OBJ $1 G_OBJ $2', round(d['ms_per_step'],3))"
done
