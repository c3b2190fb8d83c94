N=$1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL NCCL_DEBUG_FILE=gpurun_out/nccl_${N}_%p.log python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/final_bench_${N}gpu.json 2> gpurun_out/final_bench_${N}gpu.err
grep '^{' gpurun_out/final_bench_${N}gpu.json | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], d['ms_per_step'], d['value'], d.get('batch8_per_gpu',{}).get('ms_per_step'), d.get('exchange'))
"
grep -l NVLS gpurun_out/nccl_${N}_*.log | head -2; grep -h "NVLS" gpurun_out/nccl_${N}_*.log | head -5 > gpurun_out/nccl_nvls_${N}.log; rm -f gpurun_out/nccl_${N}_*.log
