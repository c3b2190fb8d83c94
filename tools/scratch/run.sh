python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "past_fc" 2>&1 | tail -3
python -m pytest tests/test_gpu_models.py tests/test_gpu_plugins.py tests/test_gpu_inversion.py -q -m gpu -x 2>&1 | tail -2
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_mlp2.json 2>/dev/null
python - <<'PY'
import json
for l in open('gpurun_out/bench_mlp2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['car_inference']['ms_per_batch'])
PY
