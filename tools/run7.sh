set -x
cd $GRAFT_REPO_ROOT
timeout 600 python tools/cluster_sweep.py > gpurun_out/r2_cluster_sweep.log 2>&1; cat gpurun_out/r2_cluster_sweep.log
timeout 300 python -m pytest tests/test_gpu_inversion.py -m gpu -q 2>&1 | tail -5
timeout 300 python tools/inv_profile.py > gpurun_out/r2_inv_profile_b.log 2>&1; cat gpurun_out/r2_inv_profile_b.log
ARTIC_DEBUG=22=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/r2_bench_1gpu_f_nocluster.json 2>/dev/null; tail -c 500 gpurun_out/r2_bench_1gpu_f_nocluster.json
