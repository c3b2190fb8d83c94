python -m pytest tests/test_gpu_dp.py -q -m gpu 2>&1 | tail -8
python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "small_ops or adam" 2>&1 | tail -3
for i in 1 2; do python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], round(d['ms_per_step'],3))
"; done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_timeline.py 2>&1 | grep -A8 "^rank 0"
