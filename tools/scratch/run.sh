set -x
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "fused_residual" 2>&1 | tail -5
python -m pytest tests/test_gpu_models.py -q -m gpu -x 2>&1 | tail -3
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_k11.json 2> gpurun_out/bench_k11.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_k11.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['gpu_launches'], d['car_inference']['ms_per_batch'] if 'car_inference' in d else [k for k in d])
PY
