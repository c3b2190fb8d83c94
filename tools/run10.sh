set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2_t10.log; tail -12 gpurun_out/r2_t10.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_g.json 2>gpurun_out/r2_bench_g.err; tail -3 gpurun_out/r2_bench_g.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_g.json') if l.startswith('{')][0])
print('ms', d['ms_per_step'], 'x3', d.get('parity_mode',{}).get('ms_per_step'), 'car', d.get('car_inference',{}).get('ms_per_batch'), d.get('car_inference',{}).get('error'), 'inv', d.get('inversion',{}).get('ms_per_batch'), d.get('inversion',{}).get('error'), 'b8', d.get('batch8_per_gpu',{}).get('ms_per_step'))
PY
