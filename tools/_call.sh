# scratch command file for `gpurun -- 'bash tools/_call.sh'` (rewritten per experiment)
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-250
