set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_inversion.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_inv.log; tail -5 gpurun_out/r2_inv.log
timeout 600 python - > gpurun_out/r2_inv_bench.json 2> gpurun_out/r2_inv_bench.err <<'PY'
import json, torch, bench
print(json.dumps(bench.inversion(torch.device("cuda", 0))))
PY
cat gpurun_out/r2_inv_bench.json | head -c 2000; tail -3 gpurun_out/r2_inv_bench.err
