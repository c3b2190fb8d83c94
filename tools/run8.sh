cd $GRAFT_REPO_ROOT
for D in "22=1" "23=0" "23=600" "23=1100"; do
ARTIC_DEBUG=$D timeout 600 python bench.py --steps 20 --warmup 5 --no-extras 2>/dev/null | python -c "import sys,json; d=json.loads([l for l in sys.stdin if l.startswith('{')][0]); print('$D', round(d['ms_per_step'],3))"
done
