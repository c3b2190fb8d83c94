python tools/weights_bench.py 2>&1 | tail -7
python -m pytest tests/test_gpu_models.py tests/test_gpu_plugins.py -q -m gpu -x 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], round(d['ms_per_step'],3))
"
